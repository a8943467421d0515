"""A/B of the EPI_STORE epilogue on one box: TMA-store staging (default) against direct 32-byte global stores
(SED_CONV_DBG=8, read per launch), for the first convs of blocks 2-4."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sed_b200 import capi, engine, synth
dev = torch.device("cuda:0")
mt = "Cnn_9layers_Gru_FrameAtt"
pm = engine.PackedModel(synth.synthetic_state_dict(mt, 16000), mt, 512, 160, dev)
mb = int(sys.argv[1]) if len(sys.argv) > 1 else 444
wave = synth.synthetic_waveform(mb, 160000).to(dev)
feat = torch.empty((mb, 125, 512), dtype=pm.tdtype, device=dev)
pm.conv_stack(wave, feat, variant=4)
ws = pm._workspace(mb, 1001)
lib = capi.load()
stream = capi.current_stream(dev)
layers = {1: ("p1", "a2"), 3: ("p2", "a3"), 5: ("p3", "a4")}
def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for rep in range(3):
    for li, (src, dst) in layers.items():
        cin, cout, mode, wp, s, b = pm.convs[li]
        x, out = ws[src], ws[dst]
        call = lambda: lib.sed_conv3x3_bn_relu(capi.ptr(x), mb, x.shape[1], x.shape[2], cin, capi.ptr(wp), capi.ptr(s),
                                               capi.ptr(b), cout, mode, capi.ptr(out), None, 0, 0, pm.dtype_code, 2, stream)
        res = []
        for mode_env in ("0", "8"):
            os.environ["SED_CONV_DBG"] = mode_env
            res.append(timeit(call))
        print("rep %d  %d->%d  staging %.4f ms  direct %.4f ms  ratio %.3f" % (rep, cin, cout, res[0], res[1], res[1] / res[0]))
os.environ["SED_CONV_DBG"] = "0"
