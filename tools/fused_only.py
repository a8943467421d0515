import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sed_b200 import engine, synth
dev = torch.device("cuda:0")
mt = "Cnn_9layers_Gru_FrameAtt"
pm = engine.PackedModel(synth.synthetic_state_dict(mt, 16000), mt, 512, 160, dev)
wave = synth.synthetic_waveform(148, 160000).to(dev)
feat = torch.empty((148, 125, 512), dtype=pm.tdtype, device=dev)
for _ in range(2):
    pm.conv_stack(wave, feat, variant=int(os.environ.get("SED_VARIANT", "4")))
torch.cuda.synchronize()
print("ok")
