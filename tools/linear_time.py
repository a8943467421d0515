"""sed_linear at the shapes of the temporal blocks (developer tool): GRU input projection [128000 x 512] x [1536 x 512]^T -> f32
transposed blocks; QKV projection of 512 clips -> split 16-bit; fc -> f32 blocks."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sed_b200 import capi, engine, synth
from tools.profile_layers import timeit
dev = torch.device("cuda:0")
lib = capi.load()
for B, N in ((1024, 1536), (512, 1536), (512, 512), (128, 1536)):
    M = B * 125 if B % 128 else B * 125
    M = (B + 127) // 128 * 128 * 125
    a = (torch.randn(M, 512, device=dev) * 0.5).half()
    w = (torch.randn(N, 512, device=dev) * 0.05).half()
    b = torch.randn(N, device=dev)
    out = torch.empty((M, N), dtype=torch.float32, device=dev)
    t = timeit(lambda: lib.sed_linear(capi.ptr(a), M, 512, capi.ptr(w), capi.ptr(b), N, 0, capi.ptr(out), None, 1, 0, capi.current_stream(dev)))
    ref = (a[:256].float() @ w.float().t() + b)
    lib.sed_linear(capi.ptr(a), M, 512, capi.ptr(w), capi.ptr(b), N, 0, capi.ptr(out), None, 0, 0, capi.current_stream(dev))
    torch.cuda.synchronize()
    err = (out[:256] - ref).abs().max().item()
    err2 = (out[-256:] - (a[-256:].float() @ w.float().t() + b)).abs().max().item()
    print("M=%6d N=%4d: %.3f ms  %.0f TFLOP/s  out %.0f MB   max err %.2e / %.2e" % (M, N, t, 2.0 * M * N * 512 / t / 1e9, M * N * 4 / 1e6, err, err2))
