"""Informational measurements of BASELINE.json configs 3-5 on one GPU (the graded line is bench.py = config 2).

config 3: Cnn_9layers_Transformer_FrameAtt, 16 kHz, the 512-clip per-GPU share of the 4096-clip job
config 4: streaming predictor, 60 s recordings at 32 kHz, 5 s windows / 1 s stride (56 windows per file)
config 5: log-mel front-end alone at 8k / 16k / 32k, resident chunk looped
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from sed_b200 import engine, streaming, synth  # noqa: E402
from tools.profile_layers import timeit  # noqa: E402

dev = torch.device("cuda:0")
res = {}

mt = "Cnn_9layers_Transformer_FrameAtt"
pm = engine.PackedModel(synth.synthetic_state_dict(mt, 16000), mt, 512, 160, dev)
wave = synth.synthetic_waveform(512, 160000).to(dev)
t = timeit(lambda: pm.forward(wave), n=5, warm=2)
res["config3_transformer_16k_b512_per_gpu"] = {"ms": t, "clips_per_s": 512 / t * 1e3}

mt = "Cnn_9layers_Gru_FrameAtt"
pm32 = engine.PackedModel(synth.synthetic_state_dict(mt, 32000), mt, 1024, 320, dev)
rec = synth.synthetic_waveform(1, 60 * 32000, seed=3, kind="events", sample_rate=32000)[0].to(dev)
t = timeit(lambda: streaming.predict_framewise(pm32, rec, 32000, 5, 1), n=10, warm=3)
res["config4_streaming_32k_60s_file"] = {"ms_per_file": t, "windows_per_s": 56 / t * 1e3, "audio_seconds_per_s": 60 / t * 1e3,
                                         "note": "one file per call (56 windows = a partial wave of the persistent grids)"}
recs = [rec] * 16
t = timeit(lambda: streaming.predict_framewise_many(pm32, recs, 32000, 5, 1), n=10, warm=3)
res["config4_streaming_32k_16_files_per_call"] = {"ms_per_call": t, "windows_per_s": 16 * 56 / t * 1e3,
                                                  "audio_seconds_per_s": 16 * 60 / t * 1e3,
                                                  "note": "896 windows per call through the offset table"}
recq = torch.round(rec * 32767).to(torch.int16)
t = timeit(lambda: streaming.predict_framewise(pm32, recq, 32000, 5, 1), n=10, warm=3)
res["config4_streaming_32k_60s_file_int16"] = {"ms_per_file": t, "windows_per_s": 56 / t * 1e3}

for sr in (8000, 16000, 32000):
    n_fft, hop, fmin, fmax = synth.PRESETS[sr]
    pmf = engine.PackedModel(synth.synthetic_state_dict(mt, sr), mt, n_fft, hop, dev)
    B = 592
    w = synth.synthetic_waveform(B, sr * 10).to(dev)
    out = torch.empty((B, sr * 10 // hop + 1, 64), device=dev)
    t = timeit(lambda: engine.logmel_forward(pmf.front, w, out=out), n=10, warm=3)
    byts = B * (4 * sr * 10 + 4 * (sr * 10 // hop + 1) * 64)
    res["config5_frontend_%dk" % (sr // 1000)] = {"ms_per_592_clips": t, "clips_per_s": B / t * 1e3,
                                                 "algorithmic_GBps": byts / t / 1e6, "hbm_frac_of_6537.6": byts / t / 1e6 / 6537.6}
print(json.dumps(res, indent=1))
