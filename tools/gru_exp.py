"""GRU recurrence timing decomposition (SED_GRU_DBG2 bits) -- developer tool."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sed_b200 import capi, engine, synth
capi.use_profile_library()  # experiment switches / stamps exist only in the -DSED_PROFILE build
from tools.profile_layers import timeit
dev = torch.device("cuda:0")
lib = capi.load()
mt = "Cnn_9layers_Gru_FrameAtt"
pm = engine.PackedModel(synth.synthetic_state_dict(mt), mt, 512, 160, dev)
T, B = 125, 1024
feat = torch.randn(B, T, 512, device=dev).half()
gi = pm.linear(feat.transpose(0, 1).contiguous().view(-1, 512), pm.gru_wih, pm.gru_bih, out_layout=1)
out = torch.empty((B, T, 512), device=dev)
ws = torch.empty((lib.sed_bigru_workspace_bytes(B),), dtype=torch.uint8, device=dev)
for flag in (0, 8, 1, 2, 4, 6, 7):
    os.environ["SED_GRU_DBG2"] = str(flag)
    t = timeit(lambda: lib.sed_bigru(capi.ptr(gi), capi.ptr(pm.gru_whh), capi.ptr(pm.gru_bhh), B, T, capi.ptr(out),
                                     capi.ptr(ws), pm.dtype_code, capi.current_stream(dev)))
    print("dbg=%d  %.3f ms  (%.2f us/step)" % (flag, t, t * 1e3 / T))
