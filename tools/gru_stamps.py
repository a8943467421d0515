"""Where does a GRU recurrence step spend its time?  clock64() stamps from CTA 0 (developer tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sed_b200 import capi, engine, synth
capi.use_profile_library()  # experiment switches / stamps exist only in the -DSED_PROFILE build
dev = torch.device("cuda:0")
lib = capi.load()
mt = "Cnn_9layers_Gru_FrameAtt"
pm = engine.PackedModel(synth.synthetic_state_dict(mt), mt, 512, 160, dev)
B, T = 1024, 125
feat = torch.randn(B, T, 512, device=dev).half()
gi = pm.linear(feat.transpose(0, 1).contiguous().view(-1, 512), pm.gru_wih, pm.gru_bih, out_layout=1)
out = torch.empty((B, T, 512), device=dev)
ws = torch.empty((lib.sed_bigru_workspace_bytes(B),), dtype=torch.uint8, device=dev)
stamps = torch.zeros(8 * 12, dtype=torch.int64, device=dev)
for _ in range(3):
    rc = lib.sed_bigru_profile(capi.ptr(gi), capi.ptr(pm.gru_whh), capi.ptr(pm.gru_bhh), B, T, capi.ptr(out), capi.ptr(ws),
                               pm.dtype_code, capi.ptr(stamps), capi.current_stream(dev))
    capi.check(rc, "profile")
torch.cuda.synchronize()
st = stamps.cpu().view(8, 12)
names = ["issuer: own slice of h seen", "issuer: own-chunk MMAs issued", "issuer: all MMAs committed",
         "issuer: starts waiting", "warp0: acc seen", "warp0: gates done", "(step 12: per-warp top, warps 0-7)",
         "(step 12: per-warp top, warps 8-15)", "warp0: arrive sent", "last warp: slice shipped"]
for s_ in range(1, 5):
    print("step", 8 + s_, " ".join("s%d=%d" % (i, st[s_, i].item() - st[s_, 0].item()) for i in (3, 0, 1, 2, 4, 5, 8, 9)),
          "| step period", st[s_, 0].item() - st[s_ - 1, 0].item())
for i, n in enumerate(names):
    print(i, n)

base12 = st[4, 0].item()
print("step 12 per-warp (relative to own-slice-seen): gi loads issued / finish")
for w in range(16):
    print("  warp %2d: top %6d  finish %6d" % (w, st[w % 8, 6 + w // 8].item() - base12, st[w % 8, 10 + w // 8].item() - base12))
