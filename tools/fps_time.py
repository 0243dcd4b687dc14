"""Times hrn_fps at the bench shape (64 clouds x 16384 points -> 1024 samples) on LiDAR-like and uniform clouds.
HRN_FPS_CULL=0 selects the un-culled cluster kernel.  python tools/fps_time.py [B] [N] [M]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcd_reg_hregnet_b200 import engine, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
M = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
clouds = {"lidar": torch.stack([synth.make_pair(100 + b, N)[1] for b in range(min(B, 8))]).repeat((B + 7) // 8, 1, 1)[:B],
          "uniform": torch.rand(B, N, 3, generator=torch.Generator().manual_seed(1))}
for name, x in clouds.items():
    x = x.cuda().contiguous()
    for _ in range(3):
        idx = engine.fps(x, M)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        idx = engine.fps(x, M)
    e.record()
    torch.cuda.synchronize()
    print(f"{name}: {s.elapsed_time(e) / 10 * 1e3:.1f} us per launch ({B} x {N} -> {M}), cull={os.environ.get('HRN_FPS_CULL', '1')}")
