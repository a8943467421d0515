"""Experiment: temporal block + head of step i on a (high-priority) side stream while the front-end + conv stack of
step i+1 run on the main stream.  Prints ms/step sequential vs overlapped (device-resident, batch 1024)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from sed_b200 import engine, synth  # noqa: E402

dev = torch.device("cuda:0")
mt = sys.argv[1] if len(sys.argv) > 1 else "Cnn_9layers_Gru_FrameAtt"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
pm = engine.PackedModel(synth.synthetic_state_dict(mt, 16000), mt, 512, 160, dev)
wave = synth.synthetic_waveform(B, 160000, seed=3).to(dev)
Tp, frames = 125, 1000
main = torch.cuda.current_stream(dev)


def run(steps, overlap, prio):
    tail = torch.cuda.Stream(dev, priority=prio) if overlap else main
    outs = [(torch.empty((B, 25), device=dev), torch.empty((B, frames, 25), device=dev)) for _ in range(2)]
    done = [None, None]
    for i in range(steps):
        feat16, feat32, slot = pm._alloc_features(B, Tp)
        pm.conv_stack(wave, **slot(0, B))
        ev = torch.cuda.Event()
        ev.record(main)
        if overlap:
            feat16.record_stream(tail)
            tail.wait_event(ev)
            if done[i % 2] is not None:
                tail.wait_event(done[i % 2])
        with torch.cuda.stream(tail):
            x = pm._temporal_or_features(feat16, feat32, B)
            pm.head(x, frames, want_cla=False, n=B, out=outs[i % 2])
            done[i % 2] = torch.cuda.Event()
            done[i % 2].record(tail)
    if overlap:
        main.wait_stream(tail)
    return outs


import statistics
res = {"sequential": [], "overlap, tail high priority": []}
for rep in range(5):
    for name, overlap, prio in (("sequential", False, 0), ("overlap, tail high priority", True, -1)):
        run(3, overlap, prio)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        outs = run(20, overlap, prio)
        e1.record()
        torch.cuda.synchronize()
        res[name].append(e0.elapsed_time(e1) / 20)
for name, v in res.items():
    print("%-30s median %.3f ms/step   runs %s" % (name, statistics.median(v), " ".join("%.2f" % x for x in v)))
