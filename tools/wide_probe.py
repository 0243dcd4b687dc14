"""CoarseReg convs_1 + attention tail (csrc/chain_wide.cu) alone at the bench shape: 32 pairs x 256 keypoints x 8 candidates
= 65,536 rows, 528 -> 512 -> 512 -> 512.  Timing, and the harness for
   ncu --set full --import-source on -k regex:chain_wide --launch-skip 2 -c 1 -o wide python tools/wide_probe.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcd_reg_hregnet_b200 import engine, engine_tc
from pcd_reg_hregnet_b200.engine import SEG_BROADCAST, SEG_GATHER, RowsView
from pcd_reg_hregnet_b200._lib import ACT_RELU
DEV = "cuda"
B, M, N, C, kseg = 32, 256, 256, 256, 8
g = torch.Generator().manual_seed(1)
rows = B * M * kseg
misc = torch.randn(rows, 16, generator=g).to(DEV)
src = torch.randn(B * M, C, generator=g).to(DEV)
dst = torch.randn(B * N, C, generator=g).to(DEV)
idx = torch.randint(0, N, (B, M, kseg), generator=g).int().to(DEV)
v = RowsView(rows, group=kseg, gather_idx=idx, rows_per_batch=M * kseg, src_rows_per_batch=N)
v.add(misc).add(src, SEG_BROADCAST).add(dst, SEG_GATHER)
widths = [16 + 2 * C, 512, 512, 512]
layers = []
for i in range(3):
    W = (torch.randn(widths[i + 1], widths[i], generator=g) / widths[i] ** 0.5).to(DEV)
    b = (torch.randn(widths[i + 1], generator=g) * 0.1).to(DEV)
    layers.append((W, b, ACT_RELU))
prec = int(sys.argv[1]) if len(sys.argv) > 1 else 3
for _ in range(3):
    engine_tc.chain_wide(v, layers, kseg, prec=prec)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    engine_tc.chain_wide(v, layers, kseg, prec=prec)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
fl = 2.0 * rows * (528 * 512 + 2 * 512 * 512)
print(f"chain_wide prec={prec}: {ms * 1e3:.1f} us per launch, {fl / ms / 1e9:.1f} useful TFLOP/s")

from pcd_reg_hregnet_b200._lib import lib
import ctypes
if hasattr(lib(), "hrn_chain_wide_prof"):          # library built with -DHRN_WIDE_PROF
    buf = (ctypes.c_longlong * 8)()
    lib().hrn_chain_wide_prof.argtypes = [ctypes.c_void_p]
    lib().hrn_chain_wide_prof(buf)
    tot = max(buf[5], 1)
    names = ["input stages", "own blocks", "peer blocks", "weights", "free accumulator"]
    print("MMA thread of CTA 0, share of its time waiting for:",
          ", ".join(f"{n} {100.0 * buf[i] / tot:.1f}%" for i, n in enumerate(names)), f"(total {tot} cycles)")
