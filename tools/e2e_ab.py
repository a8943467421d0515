"""End-to-end (host buffers in, host results out) A/B of forward_host options on one box."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sed_b200 import engine, synth
dev = torch.device("cuda:0")
mt = "Cnn_9layers_Gru_FrameAtt"
pm = engine.PackedModel(synth.synthetic_state_dict(mt, 16000), mt, 512, 160, dev)
B = 1024
wave = synth.synthetic_waveform(B, 160000).pin_memory()
q = torch.round(wave * 32767).to(torch.int16).pin_memory()
wd = wave.to(dev)
def timeit(fn, n=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3
print("device-resident forward: %.2f ms" % timeit(lambda: pm.forward(wd)))
for rep in range(3):
    for parts in (1, 2):
        for name, w in (("f32", wave), ("int16", q)):
            print("parts=%d %-5s: %.2f ms" % (parts, name, timeit(lambda: pm.forward_host(w, result_parts=parts))))

if len(sys.argv) > 1:
    for parts in (1, 2):
        for name, w in (("f32", wave), ("int16", q)):
            tr = []
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            pm.forward_host(w, result_parts=parts, trace=tr)
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            print("--- parts=%d %s: host call %.2f ms" % (parts, name, (t1 - t0) * 1e3))
            for label, ev in tr[1:]:
                print("   %-24s %7.2f ms" % (label, tr[0][1].elapsed_time(ev)))
    