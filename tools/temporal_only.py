"""Temporal block + pooling head alone at B=1024 (for ncu captures): python tools/temporal_only.py [model_type]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sed_b200 import engine, synth
mt = sys.argv[1] if len(sys.argv) > 1 else "Cnn_9layers_Gru_FrameAtt"
dev = torch.device("cuda:0")
pm = engine.PackedModel(synth.synthetic_state_dict(mt, 16000), mt, 512, 160, dev)
B, T = 1024, 125
feat_t = (torch.randn(T, B, 512, device=dev) * 0.5).to(pm.tdtype)
for _ in range(3):
    x = pm.gru_tmajor(feat_t, B) if pm.temporal_kind == "gru" else pm.mha_tmajor(feat_t, B)
    pm.head(x, 1000, want_cla=False, n=B)
torch.cuda.synchronize()
print("ok")
