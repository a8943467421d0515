"""Front-end alone (for ncu captures): python tools/fe_only.py [sample_rate] [n_calls]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sed_b200 import engine, synth
sr = int(sys.argv[1]) if len(sys.argv) > 1 else 16000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
n_fft, hop, _, _ = synth.PRESETS[sr]
dev = torch.device("cuda:0")
mt = "Cnn_9layers_Gru_FrameAtt"
pm = engine.PackedModel(synth.synthetic_state_dict(mt, sr), mt, n_fft, hop, dev)
wave = synth.synthetic_waveform(148, sr * 10).to(dev)
out = torch.empty((148, 1001, 64), device=dev)
for _ in range(n):
    engine.logmel_forward(pm.front, wave, pm.bn0_scale, pm.bn0_shift, out=out)
torch.cuda.synchronize()
print("ok")
