"""Level-1 neighbour search alone (64 clouds x 1024 FPS samples x 16384 points, K = 64) -- profiling harness:
   ncu --set full --import-source on -k regex:knn3_sorted_kernel -c 1 -o knn python tools/knn_l1.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcd_reg_hregnet_b200 import engine, ops, synth  # noqa: E402

xyz = torch.stack([synth.make_pair(1000 + b, 16384)[1] for b in range(64)]).cuda()
idx = ops.furthest_point_sample(xyz, 1024)
for _ in range(3):
    out, _ = engine.knn_idx(None, xyz, 64, q_idx=idx)
torch.cuda.synchronize()
print("ok", out.shape)
