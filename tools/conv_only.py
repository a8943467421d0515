"""Conv stack alone on one micro-batch (for ncu captures): python tools/conv_only.py [n_calls]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sed_b200 import engine, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda:0")
mt = "Cnn_9layers_Gru_FrameAtt"
pm = engine.PackedModel(synth.synthetic_state_dict(mt, 16000), mt, 512, 160, dev)
wave = synth.synthetic_waveform(148, 160000).to(dev)
feat = torch.empty((148, 125, 512), dtype=pm.tdtype, device=dev)
for _ in range(n):
    pm.conv_stack(wave, feat)
torch.cuda.synchronize()
print("ok")
