"""Debug harness for csrc/chain_wide.cu: one small case against fp64, error pattern per column half / tile."""
import sys, os
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcd_reg_hregnet_b200 import engine, engine_tc
from pcd_reg_hregnet_b200.engine import SEG_BROADCAST, SEG_GATHER, RowsView
from pcd_reg_hregnet_b200._lib import ACT_RELU
DEV = "cuda"
B, M = int(sys.argv[1]), int(sys.argv[2])
dims = [512, 512, 512]
MODE = sys.argv[3] if len(sys.argv) > 3 else "full"      # full | id3 (W3 = I) | id23 (W2 = W3 = I)
kseg, N, C = 8, 300, 256
g = torch.Generator().manual_seed(1)
rows = B * M * kseg
misc = torch.randn(rows, 16, generator=g).to(DEV)
src = torch.randn(B * M, C, generator=g).to(DEV)
dst = torch.randn(B * N, C, generator=g).to(DEV)
idx = torch.randint(0, N, (B, M, kseg), generator=g).int().to(DEV)
v = RowsView(rows, group=kseg, gather_idx=idx, rows_per_batch=M * kseg, src_rows_per_batch=N)
v.add(misc).add(src, SEG_BROADCAST).add(dst, SEG_GATHER)
r = torch.arange(rows, device=DEV)
X = torch.cat([misc, src[r // kseg], dst[(r // (M * kseg)) * N + idx.view(-1).long()]], 1).double()
widths = [16 + 2 * C] + dims
layers = []
for i in range(3):
    W = (torch.randn(widths[i + 1], widths[i], generator=g) / widths[i] ** 0.5).to(DEV)
    b = (torch.randn(widths[i + 1], generator=g) * 0.1).to(DEV)
    if (i == 2 and MODE in ("id3", "id23")) or (i == 1 and MODE == "id23"):
        W = torch.eye(512, device=DEV); b = torch.zeros(512, device=DEV)
    layers.append((W, b, ACT_RELU))
    X = torch.relu(X @ W.double().t() + b.double())
from pcd_reg_hregnet_b200._lib import lib
import ctypes
dbg = None
if hasattr(lib(), "hrn_chain_wide_set_debug"):
    dbg = torch.full((2, rows, 512), float("nan"), device=DEV)
    lib().hrn_chain_wide_set_debug.argtypes = [ctypes.c_void_p]
    lib().hrn_chain_wide_set_debug(dbg.data_ptr())
G, a = engine_tc.chain_wide(v, layers, kseg)
torch.cuda.synchronize()
if dbg is not None:
    Xl = torch.cat([misc, src[r // kseg], dst[(r // (M * kseg)) * N + idx.view(-1).long()]], 1).double()
    for li in range(2):
        W, b, _ = layers[li]
        pre = Xl @ W.double().t() + b.double()
        e = (dbg[li].double() - pre).abs()
        for rk in (0, 1):
            ee = e[:, rk * 256:(rk + 1) * 256]
            print(f"layer {li + 1} rank {rk}: nan {int(torch.isnan(ee).sum())} max err {float(ee.nan_to_num(9).max()):.3e}; per 32-col block:",
                  [f"{float(ee[:, 32 * j:32 * j + 32].nan_to_num(9).max()):.0e}" for j in range(8)])
        Xl = torch.relu(pre)
Xg = X.view(-1, kseg, dims[2])
a_ref = torch.softmax(Xg.max(dim=2)[0], dim=1)
G_ref = (a_ref[:, :, None] * Xg).sum(1)
print("a nan:", int(torch.isnan(a).sum()), "of", a.numel(), " G nan:", int(torch.isnan(G).sum()), "of", G.numel())
ea = (a.double().view(-1, kseg) - a_ref).abs()
print("a err max", float(ea.nan_to_num(9).max()))
h = dims[2] // 2
for rk in (0, 1):
    e = (G.double()[:, rk * h:(rk + 1) * h] - G_ref[:, rk * h:(rk + 1) * h]).abs() / float(X.abs().max())
    print(f"rank {rk} cols: nan {int(torch.isnan(e).sum())}, max err {float(e.nan_to_num(9).max()):.3e}, per tile max:",
          [f"{float(t.nan_to_num(9).max()):.1e}" for t in e.view(-1, 16, h)[:8]])
# unnormalised check: G / sum? compare G summed over rows with the un-attended column sums
print("G sample", G[0, :4].tolist(), G[0, h:h + 4].tolist(), "ref", G_ref[0, :4].tolist(), G_ref[0, h:h + 4].tolist())
print("a sample", a[:8].tolist(), "ref", a_ref[0].tolist())
