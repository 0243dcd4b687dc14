"""FineReg level-2 conv stack + attention (csrc/chain_tc.cu, chain_ws_kernel) alone at the bench shape: 32 pairs x 512
keypoints x 8 candidates = 131,072 rows, 268 -> 256 -> 256 -> 256; with -DHRN_CHAIN_PROF also where its MMA thread waits."""
import os, sys, ctypes
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcd_reg_hregnet_b200 import engine, engine_tc
from pcd_reg_hregnet_b200.engine import SEG_BROADCAST, SEG_GATHER, RowsView
from pcd_reg_hregnet_b200._lib import ACT_RELU, lib
DEV = "cuda"
B, M, N, C, kseg = 32, 512, 512, 128, 8
W = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = torch.Generator().manual_seed(1)
rows = B * M * kseg
misc = torch.randn(rows, 12, generator=g).to(DEV)
src = torch.randn(B * M, C, generator=g).to(DEV)
dst = torch.randn(B * N, C, generator=g).to(DEV)
idx = torch.randint(0, N, (B, M, kseg), generator=g).int().to(DEV)
v = RowsView(rows, group=kseg, gather_idx=idx, rows_per_batch=M * kseg, src_rows_per_batch=N)
v.add(misc).add(src, SEG_BROADCAST).add(dst, SEG_GATHER)
widths = [12 + 2 * C, W, W, W]
layers = []
for i in range(3):
    Wt = (torch.randn(widths[i + 1], widths[i], generator=g) / widths[i] ** 0.5).to(DEV)
    b = (torch.randn(widths[i + 1], generator=g) * 0.1).to(DEV)
    layers.append((Wt, b, ACT_RELU))
run = lambda: engine_tc.chain3(v, layers, engine_tc.EPI_ATTN, kseg, want_rows=False)
for _ in range(3):
    run()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    run()
e.record()
torch.cuda.synchronize()
print(f"chain3 (split first layer: {engine_tc.SPLIT_CHAINS}) width {W}: {s.elapsed_time(e) / 10 * 1e3:.1f} us per call incl. the per-point launches")
if hasattr(lib(), "hrn_chain_prof"):
    buf = (ctypes.c_longlong * 8)()
    lib().hrn_chain_prof.argtypes = [ctypes.c_void_p]
    lib().hrn_chain_prof(buf)
    tot = max(buf[4], 1)
    names = ["input stages", "hidden blocks", "weights", "free accumulator"]
    print("MMA thread of CTA 0 waits:", ", ".join(f"{n} {100.0 * buf[i] / tot:.1f}%" for i, n in enumerate(names)), f"(total {tot} cycles)")
