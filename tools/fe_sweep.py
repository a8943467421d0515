"""Front-end alone at the three presets (ms per 148 clips of 10 s, fraction of the HBM copy peak) -- developer tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sed_b200 import engine, synth
from tools.profile_layers import timeit
dev = torch.device("cuda:0")
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
        t = timeit(lambda: engine.logmel_forward(plan, w, out=out), n=20)
        byts = 148 * (4 * 10 * sr + 4 * (10 * sr // hop + 1) * 64)
        print("%5d Hz %s: %.3f ms per 148 clips, %.0f k clips/s, %.1f %% of 6537.6 GB/s (float32-in bytes)" % (sr, dt, t, 148 / t, 100 * byts / t / 1e6 / 6537.6))
