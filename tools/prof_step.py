"""One small forward of the hot path for ncu: `python tools/prof_step.py [--model-type M] [--batch B] [--calls N]`.
(ncu launch list: --metrics gpu__time_duration.sum; top kernels: --set full -k regex:<names>.)"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from sed_b200 import engine, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--model-type", default="Cnn_9layers_Gru_FrameAtt")
ap.add_argument("--batch", type=int, default=148)
ap.add_argument("--calls", type=int, default=2)
ap.add_argument("--sr", type=int, default=16000)
args = ap.parse_args()
dev = torch.device("cuda:0")
n_fft, hop, _, _ = synth.PRESETS[args.sr]
pm = engine.PackedModel(synth.synthetic_state_dict(args.model_type, args.sr), args.model_type, n_fft, hop, dev)
wave = synth.synthetic_waveform(args.batch, 10 * args.sr, seed=9, sample_rate=args.sr).to(dev)
for _ in range(args.calls):
    out = pm.forward(wave)
torch.cuda.synchronize()
print("ok", tuple(out["framewise_output"].shape))
