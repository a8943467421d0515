"""Blocks pooling head: stage 1 (projections) and stage 2 (per-clip pass) timed separately -- developer tool."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from sed_b200 import capi, engine, synth
from tools.profile_layers import timeit
dev = torch.device("cuda:0")
mt = "Cnn_9layers_Gru_FrameAtt"
pm = engine.PackedModel(synth.synthetic_state_dict(mt, 16000), mt, 512, 160, dev)
B, T = 1024, 125
xb = torch.randn(T, B // 128, 128, 128, 4, device=dev)
clip = torch.empty(B, 25, device=dev); frame = torch.empty(B, 1000, 25, device=dev)
scratch = torch.empty((capi.load().sed_attpool_blocks_scratch_bytes(B, T),), dtype=torch.uint8, device=dev)
t1 = timeit(lambda: pm._head_blocks(xb, B, 1000, False, False, (clip, frame), stage=1, scratch=scratch))
t2 = timeit(lambda: pm._head_blocks(xb, B, 1000, False, False, (clip, frame), stage=2, clips=(0, B), scratch=scratch))
t3 = timeit(lambda: pm._head_blocks(xb, B, 1000, True, True, (clip, frame), stage=2, clips=(0, B), scratch=scratch))
print("B=%d: projections %.3f ms, per-clip pass %.3f ms (%.0f GB/s of framewise writes), with cla+norm_att outputs %.3f ms"
      % (B, t1, t2, B * 100100 / t2 / 1e6, t3))
