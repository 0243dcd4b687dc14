"""Descriptor-space neighbour search alone at the CoarseReg shape (32 pairs x 256 x 256 descriptors of 256 dims, K = 8) --
profiling harness:  ncu --set full --import-source on -k regex:knnd_kernel -c 1 -o knnd python tools/knnd_probe.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcd_reg_hregnet_b200 import engine  # noqa: E402

torch.manual_seed(0)
src = torch.randn(32, 256, 256, device="cuda")
dst = torch.randn(32, 256, 256, device="cuda")
for _ in range(3):
    idx, _ = engine.knn_idx(src, dst, 8)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(20):
    idx, _ = engine.knn_idx(src, dst, 8)
e.record()
torch.cuda.synchronize()
print("ok", idx.shape, "us per search", s.elapsed_time(e) / 20 * 1e3)
