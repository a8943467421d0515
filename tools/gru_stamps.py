"""Where does a GRU recurrence step spend its time?  clock64() stamps from CTA 0 (developer tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sed_b200 import capi, engine, synth
dev = torch.device("cuda:0")
lib = capi.load()
mt = "Cnn_9layers_Gru_FrameAtt"
pm = engine.PackedModel(synth.synthetic_state_dict(mt), mt, 512, 160, dev)
B, T = 1024, 125
feat = torch.randn(B, T, 512, device=dev).half()
gi = pm.linear(feat.view(-1, 512), pm.gru_wih, pm.gru_bih)
out = torch.empty((B, T, 512), device=dev)
ws = torch.empty((lib.sed_bigru_workspace_bytes(B),), dtype=torch.uint8, device=dev)
stamps = torch.zeros(8 * 12, dtype=torch.int64, device=dev)
for _ in range(3):
    rc = lib.sed_bigru_profile(capi.ptr(gi), capi.ptr(pm.gru_whh), capi.ptr(pm.gru_bhh), B, T, capi.ptr(out), capi.ptr(ws),
                               pm.dtype_code, capi.ptr(stamps), capi.current_stream(dev))
    capi.check(rc, "profile")
torch.cuda.synchronize()
st = stamps.cpu().view(8, 12)
names = ["prod: h_ready seen", "prod: after proxy fence", "mma: a_full seen", "mma: committed", "epi: before acc wait",
         "epi: acc_full seen", "epi: gates done", "epi: stores issued", "epi: after proxy fence", "epi: arrives sent"]
base = st[1, 0].item()
for s in range(1, 5):
    print("step", 8 + s, " ".join("%s=%d" % (n.split(":")[0] + str(i), st[s, i].item() - st[s, 0].item()) for i, n in enumerate(names)),
          "| step period", st[s, 0].item() - st[s - 1, 0].item())
for i, n in enumerate(names):
    print(i, n)
