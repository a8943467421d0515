"""GRU recurrence time vs batch (how many 8-CTA clusters run concurrently?) -- developer tool."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["SED_GRU_DBG"] = "1"
import torch
from sed_b200 import capi, engine, synth
capi.use_profile_library()  # experiment switches / stamps exist only in the -DSED_PROFILE build
from tools.profile_layers import timeit
dev = torch.device("cuda:0")
lib = capi.load()
mt = "Cnn_9layers_Gru_FrameAtt"
pm = engine.PackedModel(synth.synthetic_state_dict(mt), mt, 512, 160, dev)
T = 125
for B in (128, 512, 1024, 2048):
    feat = torch.randn(B, T, 512, device=dev).half()
    gi = pm.linear(feat.transpose(0, 1).contiguous().view(-1, 512), pm.gru_wih, pm.gru_bih, out_layout=1)
    out = torch.empty((B, T, 512), device=dev)
    ws = torch.empty((lib.sed_bigru_workspace_bytes(B),), dtype=torch.uint8, device=dev)
    t = timeit(lambda: lib.sed_bigru(capi.ptr(gi), capi.ptr(pm.gru_whh), capi.ptr(pm.gru_bhh), B, T, capi.ptr(out),
                                     capi.ptr(ws), pm.dtype_code, capi.current_stream(dev)))
    os.environ.pop("SED_GRU_DBG", None)
    print("B=%4d clusters=%2d  %.3f ms  (%.2f us/step)" % (B, 2 * ((B + 127) // 128), t, t * 1e3 / T))
