"""GPU: the tcgen05 shared-MLP layer (csrc/mlp_tc.cu, bf16x3 split) against an fp64 evaluation of the same
virtual-rows GEMM, and against the exact-fp32 CUDA-core layer.  Tolerance: 1e-4 relative to the tensor maximum
(the path's feature gate is 1e-3)."""
import pytest
import torch

from pcd_reg_hregnet_b200 import engine
from pcd_reg_hregnet_b200.engine import SEG_BROADCAST, SEG_GATHER, RowsView
from pcd_reg_hregnet_b200._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SOFTPLUS_EPS

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ref(view_cols, W, b, act):
    y = view_cols.double() @ W.double().t() + b.double()
    if act == ACT_RELU:
        y = torch.relu(y)
    elif act == ACT_SIGMOID:
        y = torch.sigmoid(y)
    elif act == ACT_SOFTPLUS_EPS:
        y = torch.nn.functional.softplus(y) + 0.001
    return y


def _check(view, cols, Cout, act, seed):
    g = torch.Generator().manual_seed(seed)
    K = cols.shape[1]
    W = (torch.randn(Cout, K, generator=g) / K ** 0.5).to(DEV)
    b = torch.randn(Cout, generator=g).to(DEV)
    want = _ref(cols, W, b, act)
    engine.set_precision("tc")
    got = engine.layer(view, W, b, act)
    engine.set_precision("fp32")
    try:
        got32 = engine.layer(view, W, b, act)
    finally:
        engine.set_precision("tc")
    torch.cuda.synchronize()
    scale = float(want.abs().max())
    e_tc = float((got.double() - want).abs().max()) / scale
    e_32 = float((got32.double() - want).abs().max()) / scale
    assert e_32 < 1e-5, e_32
    assert e_tc < 1e-4, (e_tc, e_32)
    return e_tc


@pytest.mark.parametrize("rows,K,Cout,act", [
    (128, 32, 32, ACT_RELU), (256, 64, 64, ACT_RELU), (1000, 4, 32, ACT_RELU), (4096, 132, 128, ACT_RELU),
    (640, 256, 256, ACT_RELU), (384, 528, 512, ACT_RELU), (300, 512, 1, ACT_SIGMOID), (777, 64, 1, ACT_SOFTPLUS_EPS),
    (130, 768, 128, ACT_NONE), (65536, 32, 64, ACT_RELU),
])
def test_direct_rows(rows, K, Cout, act):
    g = torch.Generator().manual_seed(rows + K)
    X = torch.randn(rows, K, generator=g).to(DEV)
    e = _check(RowsView(rows).add(X), X, Cout, act, seed=K + Cout)
    print(f"rows={rows} K={K} Cout={Cout}: tc rel err {e:.2e}")


def test_segments_gather_broadcast_scale():
    g = torch.Generator().manual_seed(0)
    B, M, k, N, C = 3, 40, 8, 100, 64
    rows = B * M * k
    misc = torch.randn(rows, 16, generator=g).to(DEV)
    src = torch.randn(B * M, C, generator=g).to(DEV)
    dst = torch.randn(B * N, C, generator=g).to(DEV)
    a = torch.rand(rows, generator=g).to(DEV)
    idx = torch.randint(0, N, (B, M, k), generator=g).int().to(DEV)
    v = RowsView(rows, group=k, gather_idx=idx, rows_per_batch=M * k, src_rows_per_batch=N)
    v.add(misc, channels=12).add(src, SEG_BROADCAST).add(dst, SEG_GATHER, row_scale=a)
    r = torch.arange(rows, device=DEV)
    gathered = dst[(r // (M * k)) * N + idx.view(-1).long()] * a[:, None]
    cols = torch.cat([misc[:, :12], src[r // k], gathered], dim=1)
    _check(v, cols, 128, ACT_RELU, seed=5)
    # strided source / column offset
    v2 = RowsView(rows).add(misc, channels=5, col0=3).add(misc, channels=4, col0=12)
    _check(v2, torch.cat([misc[:, 3:8], misc[:, 12:16]], 1), 48, ACT_RELU, seed=6)


@pytest.mark.parametrize("kseg,mode", [(8, 2), (16, 2), (16, 1), (32, 1), (32, 2), (8, 0)])
def test_chain3_kernel(kseg, mode):
    """csrc/chain_tc.cu against an fp64 evaluation of the three layers + group reduction."""
    from pcd_reg_hregnet_b200 import engine_tc
    g = torch.Generator().manual_seed(kseg * 10 + mode)
    B, M, N, C = 2, 128 * 8 // kseg, 300, 128
    rows = B * M * kseg
    misc = torch.randn(rows, 12, generator=g).to(DEV)
    src = torch.randn(B * M, C, generator=g).to(DEV)
    dst = torch.randn(B * N, C, generator=g).to(DEV)
    idx = torch.randint(0, N, (B, M, kseg), generator=g).int().to(DEV)
    v = RowsView(rows, group=kseg, gather_idx=idx, rows_per_batch=M * kseg, src_rows_per_batch=N)
    v.add(misc).add(src, SEG_BROADCAST).add(dst, SEG_GATHER)
    r = torch.arange(rows, device=DEV)
    X = torch.cat([misc, src[r // kseg], dst[(r // (M * kseg)) * N + idx.view(-1).long()]], 1).double()
    dims = [12 + 2 * C, 256, 128, 256]
    layers = []
    for i in range(3):
        W = (torch.randn(dims[i + 1], dims[i], generator=g) / dims[i] ** 0.5).to(DEV)
        b = (torch.randn(dims[i + 1], generator=g) * 0.1).to(DEV)
        layers.append((W, b, ACT_RELU))
        X = torch.relu(X @ W.double().t() + b.double())
    assert engine_tc.chain_supported(v, layers)
    Y, G, a = engine_tc.chain3(v, layers, mode, kseg)
    torch.cuda.synchronize()
    Xg = X.view(-1, kseg, 256)
    scale = float(X.abs().max())
    if mode == 0:
        assert float((Y.double() - X).abs().max()) / scale < 1e-4
    elif mode == 1:
        assert float((Y.double() - X).abs().max()) / scale < 1e-4
        assert float((G.double() - Xg.max(dim=1)[0]).abs().max()) / scale < 1e-4
    else:
        a_ref = torch.softmax(Xg.max(dim=2)[0], dim=1)
        assert float((a.double().view(-1, kseg) - a_ref).abs().max()) < 1e-4
        assert float((G.double() - (a_ref[:, :, None] * Xg).sum(1)).abs().max()) / scale < 1e-4
        assert float((Y.double().view(-1, kseg, 256) - a_ref[:, :, None] * Xg).abs().max()) / scale < 1e-4


def test_chain_two_layers_narrow_output_and_activation():
    """Generalised chain: 2 layers, last width 1 (zero-padded to 16), sigmoid; multi-pass K (> 256 input channels)."""
    from pcd_reg_hregnet_b200 import engine_tc
    g = torch.Generator().manual_seed(3)
    rows = 1024
    X = torch.randn(rows, 320, generator=g).to(DEV)
    W1 = (torch.randn(128, 320, generator=g) / 18).to(DEV); b1 = torch.randn(128, generator=g).to(DEV) * 0.1
    W2 = (torch.randn(1, 128, generator=g) / 11).to(DEV); b2 = torch.randn(1, generator=g).to(DEV)
    layers = [(W1, b1, ACT_RELU), (W2, b2, ACT_SIGMOID)]
    v = RowsView(rows).add(X)
    assert engine_tc.chain_supported(v, layers, last_relu_only=False)
    Y, _, _ = engine_tc.chain(v, layers, engine_tc.EPI_STORE)
    want = torch.sigmoid(torch.relu(X.double() @ W1.double().t() + b1.double()) @ W2.double().t() + b2.double())
    assert Y.shape == (rows, 1) and float((Y.double() - want).abs().max()) < 1e-5


@pytest.mark.parametrize("rows,K0,dims,mode,kseg", [
    (128 * 700, 140, [128, 128, 128], 2, 8),      # narrow: 4 TMEM accumulators in rotation, ~5 tiles per CTA
    (128 * 450, 132, [128, 128, 256], 1, 16),     # wide last layer with the row store through the transpose tile
    (128, 64, [64, 32, 48], 0, 8),                # a single tile, odd (multiple-of-16) widths
    (128 * 149, 776, [256, 256, 256], 2, 32),     # one tile more than the grid; 25 input stages (> the operand rings)
    (128 * 300, 96, [128, 64], 1, 8),             # two layers
])
def test_chain_persistent_schedule(rows, K0, dims, mode, kseg):
    """Persistent warp-specialised chain: several tiles per CTA (ring / accumulator hand-offs across tiles), wide
    inputs, two layers -- against fp64 on a strided row sample (every tile is hit)."""
    from pcd_reg_hregnet_b200 import engine_tc
    g = torch.Generator().manual_seed(rows % 1000 + K0)
    X = torch.randn(rows, K0, generator=g).to(DEV)
    layers, kin = [], K0
    for w in dims:
        layers.append(((torch.randn(w, kin, generator=g) / kin ** 0.5).to(DEV), (torch.randn(w, generator=g) * 0.1).to(DEV), ACT_RELU))
        kin = w
    v = RowsView(rows).add(X)
    assert engine_tc.chain_supported(v, layers)
    Y, G, a = engine_tc.chain(v, layers, mode, kseg)
    torch.cuda.synchronize()
    # fp64 on whole groups taken from every tile: groups g = 0, s, 2s, ... with s chosen so that ~4096 rows are checked
    ngrp = rows // kseg
    step = max(1, ngrp // (4096 // kseg))
    gsel = torch.arange(0, ngrp, step, device=DEV)
    rsel = (gsel[:, None] * kseg + torch.arange(kseg, device=DEV)[None, :]).reshape(-1)
    R = X[rsel].double()
    for W, b, _ in layers:
        R = torch.relu(R @ W.double().t() + b.double())
    C = dims[-1]
    Rg = R.view(-1, kseg, C)
    scale = float(R.abs().max())
    if mode == 0:
        assert float((Y[rsel].double() - R).abs().max()) / scale < 1e-4
    elif mode == 1:
        assert float((Y[rsel].double() - R).abs().max()) / scale < 1e-4
        assert float((G[gsel].double() - Rg.max(dim=1)[0]).abs().max()) / scale < 1e-4
    else:
        a_ref = torch.softmax(Rg.max(dim=2)[0], dim=1)
        assert float((a[rsel].double().view(-1, kseg) - a_ref).abs().max()) < 1e-4
        assert float((G[gsel].double() - (a_ref[:, :, None] * Rg).sum(1)).abs().max()) / scale < 1e-4
        assert float((Y[rsel].double().view(-1, kseg, C) - a_ref[:, :, None] * Rg).abs().max()) / scale < 1e-4


@pytest.mark.parametrize("k,rows,K,N", [(16, 4096, 256, 256), (8, 1024 + 128, 64, 512), (32, 2048, 96, 64)])
def test_layer_groupmax_epilogue_equals_layer_then_group_max(k, rows, K, N):
    """hrn_layer_tc_groupmax == hrn_group_max(hrn_layer_tc): max_i act(x_i + b) = act(max_i x_i + b) exactly."""
    from pcd_reg_hregnet_b200 import engine, engine_tc
    from pcd_reg_hregnet_b200.engine import RowsView, ACT_RELU
    g = torch.Generator(device="cuda").manual_seed(k)
    X = torch.randn(rows, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    b = torch.randn(N, device="cuda", generator=g)
    engine.set_precision("tc")
    v = RowsView(rows).add(X)
    assert engine_tc.layer_tc_groupmax_ok(v, W, ACT_RELU, k)
    got = engine_tc.layer_tc_groupmax(v, W, b, ACT_RELU, k)
    want = engine.group_max(engine.layer(RowsView(rows).add(X), W, b, ACT_RELU), k)
    assert got.shape == (rows // k, N)
    assert torch.equal(got, want)


@pytest.mark.parametrize("rows,K,Cout", [(4096, 528, 512), (1024, 64, 64), (640, 256, 256), (300, 512, 1)])
def test_single_pass_fp16_layer(rows, K, Cout):
    """prec = 1 of the per-layer kernel (single-pass fp16 operands): against fp64 within the 11-bit significand's error
    (2^-11 per product, random signs), and against the same kernel in bf16x3 -- layout / descriptor errors would show as
    O(1) differences, not O(1e-3)."""
    from pcd_reg_hregnet_b200 import engine_tc
    g = torch.Generator().manual_seed(rows + Cout)
    X = torch.randn(rows, K, generator=g).to(DEV)
    W = (torch.randn(Cout, K, generator=g) / K ** 0.5).to(DEV)
    b = torch.randn(Cout, generator=g).to(DEV)
    want = torch.relu(X.double() @ W.double().t() + b.double())
    out1 = engine_tc.layer_tc(RowsView(rows).add(X), W, b, ACT_RELU, torch.empty(rows, Cout, device=DEV), prec=1)
    out3 = engine_tc.layer_tc(RowsView(rows).add(X), W, b, ACT_RELU, torch.empty(rows, Cout, device=DEV), prec=3)
    torch.cuda.synchronize()
    scale = float(want.abs().max())
    e1 = float((out1.double() - want).abs().max()) / scale
    e3 = float((out3.double() - want).abs().max()) / scale
    print(f"rows={rows} K={K} Cout={Cout}: fp16 single pass {e1:.1e}, bf16x3 {e3:.1e}")
    assert e3 < 1e-4 and e1 < 2e-3


@pytest.mark.parametrize("kseg,mode,dims", [(8, 2, [256, 256, 256]), (8, 2, [128, 128, 128]), (16, 1, [128, 128, 256]), (8, 0, [64, 32, 48])])
def test_single_pass_fp16_chain(kseg, mode, dims):
    """prec = 1 of the chain kernel on gathered / broadcast segments: every epilogue mode against fp64."""
    from pcd_reg_hregnet_b200 import engine_tc
    g = torch.Generator().manual_seed(kseg + mode + dims[0])
    B, M, N, C = 2, 128 * 8 // kseg * 3, 300, 128
    rows = B * M * kseg
    misc = torch.randn(rows, 12, generator=g).to(DEV)
    src = torch.randn(B * M, C, generator=g).to(DEV)
    dst = torch.randn(B * N, C, generator=g).to(DEV)
    idx = torch.randint(0, N, (B, M, kseg), generator=g).int().to(DEV)
    v = RowsView(rows, group=kseg, gather_idx=idx, rows_per_batch=M * kseg, src_rows_per_batch=N)
    v.add(misc).add(src, SEG_BROADCAST).add(dst, SEG_GATHER)
    r = torch.arange(rows, device=DEV)
    X = torch.cat([misc, src[r // kseg], dst[(r // (M * kseg)) * N + idx.view(-1).long()]], 1).double()
    layers, kin = [], 12 + 2 * C
    for wdt in dims:
        W = (torch.randn(wdt, kin, generator=g) / kin ** 0.5).to(DEV)
        b = (torch.randn(wdt, generator=g) * 0.1).to(DEV)
        layers.append((W, b, ACT_RELU))
        X = torch.relu(X @ W.double().t() + b.double())
        kin = wdt
    Y, G, a = engine_tc.chain(v, layers, mode, kseg, prec=1)
    torch.cuda.synchronize()
    Cl = dims[-1]
    Xg = X.view(-1, kseg, Cl)
    scale = float(X.abs().max())
    tol = 3e-3
    if mode == 0:
        assert float((Y.double() - X).abs().max()) / scale < tol
    elif mode == 1:
        assert float((Y.double() - X).abs().max()) / scale < tol
        assert float((G.double() - Xg.max(dim=1)[0]).abs().max()) / scale < tol
    else:
        a_ref = torch.softmax(Xg.max(dim=2)[0], dim=1)
        assert float((a.double().view(-1, kseg) - a_ref).abs().max()) < 1e-2
        assert float((G.double() - (a_ref[:, :, None] * Xg).sum(1)).abs().max()) / scale < 2 * tol


@pytest.mark.parametrize("split", [True, False])
@pytest.mark.parametrize("B,M,dims,prec,tol", [
    (2, 256, [512, 512, 512], 3, 1e-4),          # CoarseReg convs_1 as in the model (528 -> 512 -> 512 -> 512), 32 tiles
    (1, 16, [512, 512, 512], 3, 1e-4),           # one tile: a single cluster
    (5, 80, [512, 256, 384], 3, 1e-4),           # 25 tiles, unequal widths (blocks per rank 8 / 4 / 6)
    (37, 256, [512, 512, 512], 3, 1e-4),         # 592 tiles: several rounds per cluster, ring / phase wrap-around
    (2, 256, [512, 512, 512], 1, 3e-3),          # single-pass fp16 operands
])
def test_chain_wide_cluster_kernel(B, M, dims, prec, tol, split):
    """csrc/chain_wide.cu (2-CTA cluster, hidden activations through DSMEM) against an fp64 evaluation of the three
    layers + softmax_k(max_c) attention + attentive feature (reference layers.py:364-390).  split: the group-constant and
    the gathered input segments of the first layer applied once per point and added as fp32 rows (the model's path)."""
    from pcd_reg_hregnet_b200 import engine_tc
    kseg, N, C = 8, 300, 256
    g = torch.Generator().manual_seed(B * 1000 + M + dims[1])
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
        layers.append((W, b, ACT_RELU))
        X = torch.relu(X @ W.double().t() + b.double())
    assert engine_tc.chain_wide_supported(v, layers, kseg)
    G, a = engine_tc.chain_wide(v, layers, kseg, prec=prec, split_first=split)
    G2, a2 = engine_tc.chain_wide(v, layers, kseg, prec=prec, split_first=split)     # a second launch: same bits (no race between the CTAs)
    torch.cuda.synchronize()
    Xg = X.view(-1, kseg, dims[2])
    scale = float(X.abs().max())
    a_ref = torch.softmax(Xg.max(dim=2)[0], dim=1)
    e_a = float((a.double().view(-1, kseg) - a_ref).abs().max())
    e_g = float((G.double() - (a_ref[:, :, None] * Xg).sum(1)).abs().max()) / scale
    print(f"chain_wide B={B} M={M} dims={dims} prec={prec} split={split}: a {e_a:.2e}  AF {e_g:.2e}")
    assert e_a < tol and e_g < tol, (e_a, e_g)
    assert torch.equal(G, G2) and torch.equal(a, a2)


@pytest.mark.parametrize("rows,c", [(8192, 512), (128, 512), (1024 + 128, 384)])
def test_chain_wide_head(rows, c):
    """Per-keypoint confidence head c -> c -> c -> 1 + sigmoid (reference layers.py:391-394) as one launch of the cluster
    kernel, against fp64."""
    from pcd_reg_hregnet_b200 import engine_tc
    g = torch.Generator().manual_seed(rows + c)
    X = torch.randn(rows, c, generator=g).to(DEV)
    layers, Y = [], X.double()
    for i, (n, act) in enumerate([(c, ACT_RELU), (c, ACT_RELU), (1, ACT_NONE)]):
        W = (torch.randn(n, c, generator=g) / c ** 0.5).to(DEV)
        b = (torch.randn(n, generator=g) * 0.1).to(DEV)
        layers.append((W, b, act))
        Y = Y @ W.double().t() + b.double()
        if act == ACT_RELU:
            Y = torch.relu(Y)
    want = torch.sigmoid(Y)[:, 0]
    v = RowsView(rows).add(X)
    assert engine_tc.chain_wide_head_supported(v, layers)
    got = engine_tc.chain_wide_head(v, layers, ACT_SIGMOID)
    torch.cuda.synchronize()
    e = float((got.double() - want).abs().max())
    print(f"chain_wide_head rows={rows} c={c}: {e:.2e}")
    assert got.shape == (rows,) and e < 1e-5, e


@pytest.mark.parametrize("dims", [[256, 256, 256], [128, 128, 128]])
def test_chain_first_layer_split_equals_unsplit(dims):
    """FineReg-shaped conv stack through the single-CTA chain kernel with and without the first-layer split (group-constant and
    gathered input segments applied once per point, added as fp32 rows): both within 1e-4 of fp64, and of each other."""
    from pcd_reg_hregnet_b200 import engine_tc
    kseg, B, M, N, C = 8, 3, 256, 200, dims[0] // 2
    g = torch.Generator().manual_seed(dims[0])
    rows = B * M * kseg
    misc = torch.randn(rows, 12, generator=g).to(DEV)
    src = torch.randn(B * M, C, generator=g).to(DEV)
    dst = torch.randn(B * N, C, generator=g).to(DEV)
    idx = torch.randint(0, N, (B, M, kseg), generator=g).int().to(DEV)

    def view():
        v = RowsView(rows, group=kseg, gather_idx=idx, rows_per_batch=M * kseg, src_rows_per_batch=N)
        return v.add(misc).add(src, SEG_BROADCAST).add(dst, SEG_GATHER)
    r = torch.arange(rows, device=DEV)
    X = torch.cat([misc, src[r // kseg], dst[(r // (M * kseg)) * N + idx.view(-1).long()]], 1).double()
    widths = [12 + 2 * C] + dims
    layers = []
    for i in range(3):
        W = (torch.randn(widths[i + 1], widths[i], generator=g) / widths[i] ** 0.5).to(DEV)
        b = (torch.randn(widths[i + 1], generator=g) * 0.1).to(DEV)
        layers.append((W, b, ACT_RELU))
        X = torch.relu(X @ W.double().t() + b.double())
    Xg = X.view(-1, kseg, dims[2])
    a_ref = torch.softmax(Xg.max(dim=2)[0], dim=1)
    af_ref = (a_ref[:, :, None] * Xg).sum(1)
    out = {}
    try:
        for split in (True, False):
            engine_tc.SPLIT_CHAINS = split
            _, af, a = engine_tc.chain3(view(), layers, engine_tc.EPI_ATTN, kseg, want_rows=False)
            torch.cuda.synchronize()
            out[split] = (af, a)
            assert float((a.double().view(-1, kseg) - a_ref).abs().max()) < 1e-4
            assert float((af.double() - af_ref).abs().max()) / float(X.abs().max()) < 1e-4
    finally:
        engine_tc.SPLIT_CHAINS = True
    assert float((out[True][0] - out[False][0]).abs().max()) / float(X.abs().max()) < 1e-4
