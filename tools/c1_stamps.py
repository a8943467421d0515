"""Pipeline timeline of the fused conv_block1 kernel (debug build: make -C .../csrc clean all NVFLAGS+=-DSED_C1_STAMPS).
Prints, per work item of CTA 0, the clock of every hand-over point relative to the first stamp."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from sed_b200 import engine, synth, capi
dev = torch.device("cuda:0")
mt = "Cnn_9layers_Gru_FrameAtt"
pm = engine.PackedModel(synth.synthetic_state_dict(mt, 16000), mt, 512, 160, dev)
wave = synth.synthetic_waveform(148, 160000).to(dev)
feat = torch.empty((148, 125, 512), dtype=pm.tdtype, device=dev)
for _ in range(3):
    pm.conv_stack(wave, feat, variant=4)
torch.cuda.synchronize()
lib = ctypes.CDLL(capi.LIB_PATH)
buf = np.zeros((3, 32, 8), dtype=np.int64)
rc = lib.sed_debug_c1_stamps(buf.ctypes.data_as(ctypes.c_void_p))
assert rc == 0
t0 = buf[buf > 0].min()
rel = np.where(buf > 0, buf - t0, -1)
names = {0: "mma : op_full c1_empty t_empty a_full issued", 1: "epi : t_full done",
         2: "prod: b.start b.bar b.op_empty b.done d.start d.c1_full d.a_empty d.done"}
for role in range(3):
    print(names[role])
    for i in range(32):
        print("  %3d " % i + " ".join("%7d" % v for v in rel[role, i] if v >= 0))
per = np.diff(buf[0, :, 4])
print("clk per item (mma issue to issue):", per.tolist(), "mean", per.mean())
