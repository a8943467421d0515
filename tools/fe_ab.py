"""Front-end A/B on one box: the three presets x {f32, i16} with the library given on the command line
(python tools/fe_ab.py [path/to/lib.so]) -- ms per 148 clips of 10 s, clips/s, fraction of the HBM copy peak."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from sed_b200 import capi
if len(sys.argv) > 1:
    capi.LIB_PATH = os.path.abspath(sys.argv[1])
import torch
from sed_b200 import engine, synth
from tools.profile_layers import timeit
dev = torch.device("cuda:0")
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6537.6
res = {}
for sr in (8000, 16000, 32000):
    n_fft, hop, _, _ = synth.PRESETS[sr]
    sd = synth.synthetic_state_dict("Cnn_9layers_Gru_FrameAtt", sr)
    plan = engine.FrontendPlan(sd["spectrogram_extractor.stft.conv_real.weight"], sd["spectrogram_extractor.stft.conv_imag.weight"],
                               n_fft, hop, sd["logmel_extractor.melW"], dev)
    for dt in ("f32", "i16"):
        w = synth.synthetic_waveform(148, 10 * sr, seed=1).to(dev)
        if dt == "i16":
            w = torch.round(w * 32767).to(torch.int16)
        out = torch.empty((148, 10 * sr // hop + 1, 64), device=dev)
        t = timeit(lambda: engine.logmel_forward(plan, w, out=out), n=30)
        byts = 148 * (4 * 10 * sr + 4 * (10 * sr // hop + 1) * 64)
        res["%d_%s" % (sr, dt)] = t
        print("%s %5d Hz %s: %.4f ms per 148 clips, %.0f k clips/s, %.1f %% of %.1f GB/s (float32-in bytes)" % (
            os.path.basename(capi.LIB_PATH), sr, dt, t, 148 / t, 100 * byts / t / 1e6 / peak, peak), flush=True)
print(json.dumps({"lib": os.path.basename(capi.LIB_PATH), "ms_per_148_clips": res}))
