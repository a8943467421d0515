import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from sed_b200 import capi, engine, synth
capi.use_profile_library()  # experiment switches / stamps exist only in the -DSED_PROFILE build
from tools.profile_layers import timeit
dev = torch.device("cuda:0")
sd = synth.synthetic_state_dict("Cnn_9layers_Gru_FrameAtt", 16000)
pm = engine.PackedModel(sd, "Cnn_9layers_Gru_FrameAtt", 512, 160, dev)
wave = synth.synthetic_waveform(148, 160000).to(dev)
out = torch.empty((148, 1001, 64), device=dev)
for flag, name in ((0, "full"), (1, "first pass only"), (2, "no mel/log"), (3, "first pass only, no mel")):
    os.environ["SED_FE_DBG"] = str(flag)
    t = timeit(lambda: engine.logmel_forward(pm.front, wave, pm.bn0_scale, pm.bn0_shift, out=out))
    print("%-28s %.3f ms" % (name, t))
