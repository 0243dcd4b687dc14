"""Times the weighted samplings of levels 2 / 3 (64 clouds: 1024 -> 512 and 512 -> 256).  HRN_FPS_WARP=0: cluster kernels."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcd_reg_hregnet_b200 import engine
g = torch.Generator().manual_seed(0)
for N, M in ((1024, 512), (512, 256)):
    x = (torch.rand(64, N, 3, generator=g) * 100).cuda()
    w = (torch.rand(64, N, generator=g) + 0.1).cuda()
    for weights in (w, None):
        for _ in range(3):
            engine.fps(x, M, weights)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(20):
            engine.fps(x, M, weights)
        e.record()
        torch.cuda.synchronize()
        print(f"N={N} M={M} weighted={weights is not None}: {s.elapsed_time(e) / 20 * 1e3:.1f} us (HRN_FPS_WARP={os.environ.get('HRN_FPS_WARP', '1')})")
