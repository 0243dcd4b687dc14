"""GPU parity tests of the point ops, called through the C ABI (pcd_reg_hregnet_b200.ops -> libhregnet_b200.so)
against the CPU oracle (oracle/native_ops.c).  Bar: indices and distances BIT-EXACT."""
import importlib.util
import os

import pytest
import torch

from oracle import native
from pcd_reg_hregnet_b200 import engine, ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _clouds(kind, B, N, seed):
    g = torch.Generator().manual_seed(seed)
    if kind == "lidar":
        return torch.stack([synth.make_pair(seed + b, N)[1] for b in range(B)])
    if kind == "uniform":
        return torch.rand(B, N, 3, generator=g)
    if kind == "lattice":      # many exact ties
        return torch.randint(0, 6, (B, N, 3), generator=g).float()
    if kind == "far":          # coordinates beyond the +-128 m Morton lattice (clamped keys, still exact)
        return (torch.rand(B, N, 3, generator=g) - 0.5) * 2000
    if kind == "dup":          # dataset-style duplicate padding (dataset/dataset_utils.py:203-207)
        return torch.stack([synth.duplicate_padded_cloud(seed + b, N, max(N // 2, 1)) for b in range(B)])
    raise ValueError(kind)


@pytest.mark.parametrize("kind,B,N,M", [
    ("lidar", 2, 16384, 1024), ("uniform", 3, 16384, 1024), ("lattice", 2, 4096, 512), ("dup", 2, 8192, 700),
    ("lidar", 2, 8096, 1024), ("uniform", 2, 10000, 333), ("uniform", 4, 1024, 512), ("lattice", 3, 512, 256),
    ("uniform", 2, 1000, 1000), ("lattice", 2, 300, 64), ("uniform", 2, 37, 20), ("uniform", 1, 5, 5),
    ("uniform", 2, 2048, 1), ("lidar", 1, 20000, 300), ("dup", 1, 32768, 257),
    # cluster-size coverage of the all-in-registers kernel: 64 clouds -> CS=2/P=8 (the bench shape), 20 -> CS=4/P=4,
    # 40000 points -> CS=8/P=5->8, 70000 points -> streaming kernel
    ("uniform", 64, 16384, 1024), ("uniform", 20, 16384, 256), ("uniform", 40, 8192, 128), ("uniform", 2, 40000, 200),
    ("uniform", 1, 70000, 64),
    # more samples than the 10-bit iteration tag of the exchange packets counts (tags wrap, buffers alternate)
    ("uniform", 2, 4096, 3000), ("lidar", 1, 16384, 2500), ("uniform", 2, 1024, 1024),
    # 16-CTA clusters (65536 < N <= 131072), streaming kernel beyond
    ("uniform", 2, 131072, 100), ("lattice", 1, 100000, 64), ("uniform", 1, 140000, 40),
    # the spatially culled kernel (8192 < N <= 16384): exact ties on a lattice, duplicate padding, far coordinates,
    # ragged sizes around its limits, more samples than tags, LiDAR-like sweeps at the bench shape
    ("lattice", 3, 16384, 600), ("lattice", 2, 9000, 400), ("dup", 3, 16384, 1024), ("dup", 2, 12000, 777),
    ("far", 2, 16384, 512), ("uniform", 2, 8193, 300), ("lidar", 3, 16383, 1024), ("lidar", 2, 12345, 2000),
    ("lidar", 64, 16384, 1024),
    # the few-warps-per-cloud kernels of the small power-of-two clouds (256: one warp, 512: two, 1024: four)
    ("lattice", 5, 256, 256), ("uniform", 3, 256, 100), ("lattice", 4, 1024, 700), ("dup", 3, 512, 300),
])
def test_fps_bit_exact(kind, B, N, M):
    xyz = _clouds(kind, B, N, seed=N + M)
    want = native.fps(xyz, M)
    got = ops.furthest_point_sample(xyz.to(DEV), M)
    assert got.dtype == torch.int32 and got.shape == (B, M)
    assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize("kind,B,N,M", [
    ("uniform", 4, 1024, 512), ("lidar", 2, 1024, 512), ("lattice", 3, 512, 256), ("uniform", 2, 512, 256),
    ("dup", 2, 16384, 400), ("uniform", 1, 700, 128), ("lattice", 1, 24000, 100), ("uniform", 64, 1024, 512),
    ("uniform", 80, 9000, 100), ("uniform", 3, 50000, 50), ("lattice", 5, 256, 200), ("dup", 4, 1024, 1000),
])
def test_weighted_fps_bit_exact(kind, B, N, M):
    xyz = _clouds(kind, B, N, seed=7 * N + M)
    g = torch.Generator().manual_seed(N)
    w = torch.rand(B, N, generator=g) * 3 + 0.05
    if kind == "lattice":
        w = torch.randint(1, 3, (B, N), generator=g).float()
    want = native.fps(xyz, M, w)
    got = ops.weighted_furthest_point_sample(xyz.to(DEV), w.to(DEV), M)
    assert torch.equal(got.cpu(), want)


@pytest.mark.parametrize("N,M", [(1024, 300), (3000, 200), (16384, 150)])
def test_weighted_fps_with_negative_and_zero_weights(N, M):
    """The reference multiplies the distance by whatever weight it is given (furthest_point_sampling_gpu.cu:299);
    negative and zero weights make the "distances" negative / zero -- the ordered-key reductions must still agree."""
    xyz = _clouds("uniform", 2, N, seed=N)
    g = torch.Generator().manual_seed(N + 1)
    w = torch.rand(2, N, generator=g) * 3 - 1
    w[:, ::7] = 0.0
    want = native.fps(xyz, M, w)
    got = ops.weighted_furthest_point_sample(xyz.to(DEV), w.to(DEV), M)
    assert torch.equal(got.cpu(), want)


def test_fps_temp_scratch_semantics_via_point_utils_shim():
    """The reference's positional module contract (point_utils_api.cpp:6-13): caller-provided temp is the initial
    min-distance array and is clobbered with the final one."""
    xyz = _clouds("uniform", 2, 3000, 1)
    temp_o = torch.full((2, 3000), 1e10)
    want = native.fps(xyz, 100, None, temp_o)
    temp = torch.full((2, 3000), 1e10, device=DEV)
    out = torch.empty(2, 100, dtype=torch.int32, device=DEV)
    assert ops.point_utils_cuda.furthest_point_sampling_wrapper(2, 3000, 100, xyz.to(DEV), temp, out) == 1
    assert torch.equal(out.cpu(), want) and torch.equal(temp.cpu(), temp_o)


def _load_ref_ext():
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "point_utils_cuda.so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location("point_utils_cuda", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_against_unmodified_reference_kernels():
    """oracle/_ref/point_utils_cuda.so = the reference's own .cu/.cpp compiled for sm_100a (oracle/build_ref.py):
    pins BOTH the new kernels and the CPU oracle to the real reference (rounding + tie-break)."""
    ref = _load_ref_ext()
    if ref is None:
        pytest.skip("oracle/_ref not built")
    for kind, B, N, M in [("lidar", 2, 16384, 1024), ("lattice", 2, 4096, 512), ("dup", 2, 8192, 600),
                          ("uniform", 3, 1024, 512), ("lattice", 2, 512, 256), ("uniform", 1, 10000, 200)]:
        xyz = _clouds(kind, B, N, seed=N).to(DEV)
        temp = torch.full((B, N), 1e10, device=DEV)
        out = torch.empty(B, M, dtype=torch.int32, device=DEV)
        ref.furthest_point_sampling_wrapper(B, N, M, xyz, temp, out)
        assert torch.equal(out, ops.furthest_point_sample(xyz, M)), (kind, N)
        assert torch.equal(out.cpu(), native.fps(xyz.cpu(), M)), ("oracle", kind, N)
        w = (torch.rand(B, N, generator=torch.Generator().manual_seed(N)) * 2 + 0.1).to(DEV)
        temp.fill_(1e10)
        ref.weighted_furthest_point_sampling_wrapper(B, N, M, xyz, w, temp, out)
        assert torch.equal(out, ops.weighted_furthest_point_sample(xyz, w, M)), ("weighted", kind, N)
        assert torch.equal(out.cpu(), native.fps(xyz.cpu(), M, w.cpu())), ("oracle weighted", kind, N)
        feat = torch.rand(B, 7, N, device=DEV)
        g_ref = torch.empty(B, 7, M, device=DEV)
        ref.gather_points_wrapper(B, 7, N, M, feat, out, g_ref)
        assert torch.equal(g_ref, ops.gather_operation(feat, out))


def test_gather_forward_backward():
    g = torch.Generator().manual_seed(0)
    feat = torch.randn(3, 5, 200, generator=g).to(DEV).requires_grad_(True)
    idx = torch.randint(0, 200, (3, 64), generator=g).int().to(DEV)
    out = ops.gather_operation(feat, idx)
    want = torch.gather(feat, 2, idx.long()[:, None, :].expand(-1, 5, -1))
    assert torch.equal(out, want)
    go = torch.randn_like(out)
    out.backward(go)
    gw, = torch.autograd.grad(want, feat, go)
    assert torch.allclose(feat.grad, gw, atol=1e-6)


@pytest.mark.parametrize("kind,B,M,N,K", [
    ("lidar", 2, 1024, 16384, 64), ("uniform", 2, 512, 1024, 32), ("uniform", 2, 256, 512, 16),
    ("lattice", 2, 200, 3000, 64), ("lattice", 2, 256, 256, 8), ("dup", 1, 300, 5000, 32), ("uniform", 1, 7, 9, 9),
    ("uniform", 2, 1024, 1024, 8), ("uniform", 1, 100, 2049, 1),
    # spatially culled path (2048 <= N <= 16384): ties, duplicates, non-power-of-two N, points outside the lattice
    ("lattice", 2, 512, 16384, 64), ("dup", 2, 700, 8096, 32), ("lidar", 1, 64, 10000, 16), ("far", 1, 128, 4096, 8),
    ("uniform", 64, 1024, 16384, 64),
    # 4-byte sort keys (16384 < N <= 32768): full and ragged clouds, ties, duplicates; beyond: brute force
    ("lidar", 2, 512, 32768, 64), ("lattice", 1, 300, 20000, 32), ("dup", 1, 256, 32768, 16), ("far", 1, 100, 25000, 8),
    ("uniform", 1, 64, 40000, 16),
])
def test_knn_xyz_bit_exact(kind, B, M, N, K):
    p2 = _clouds(kind, B, N, seed=N + K)
    p1 = _clouds(kind, B, M, seed=M + 3 * K) if kind != "dup" else p2[:, :M].contiguous()
    d_o, i_o, nn_o = native.knn_points(p1, p2, K=K, return_nn=True)
    d, i, nn = ops.knn_points(p1.to(DEV), p2.to(DEV), K=K, return_nn=True)
    assert i.dtype == torch.int64
    assert torch.equal(i.cpu(), i_o)
    assert torch.equal(d.cpu(), d_o)
    assert torch.equal(nn.cpu(), nn_o)
    assert (d[..., 1:] >= d[..., :-1]).all()


@pytest.mark.parametrize("B,M,N,D,K", [(2, 256, 256, 256, 8), (1, 50, 300, 64, 8), (1, 33, 70, 17, 40)])
def test_knn_descriptor_space_bit_exact(B, M, N, D, K):
    g = torch.Generator().manual_seed(D)
    p1, p2 = torch.rand(B, M, D, generator=g), torch.rand(B, N, D, generator=g)
    d_o, i_o, nn_o = native.knn_points(p1, p2, K=K, return_nn=True)
    d, i, nn = ops.knn_points(p1.to(DEV), p2.to(DEV), K=K, return_nn=True)
    assert torch.equal(i.cpu(), i_o) and torch.equal(d.cpu(), d_o) and torch.equal(nn.cpu(), nn_o)


@pytest.mark.parametrize("B,M,N,D,K", [(2, 100, 512, 32, 16), (1, 40, 130, 8, 5), (2, 64, 256, 256, 8), (1, 37, 513, 12, 16),
                                       (1, 20, 16, 6, 16)])
def test_knn_descriptor_space_ties_and_overflow(B, M, N, D, K):
    """Both selection paths of the descriptor-space search (sorted extract-min for K <= 16 of N <= 512, the running set
    otherwise) on what separates them from a plain sort: exact distance ties (every reference stored twice or more, so
    the order is decided by the index), queries that coincide with references, squared distances that overflow to +inf
    (ordered by index among themselves and in front of nothing), unaligned views (4-byte copy path)."""
    g = torch.Generator().manual_seed(N + K)
    base = torch.rand(B, max(N // 3, 1), D, generator=g)
    p2 = base[:, torch.randint(0, base.shape[1], (N,), generator=g)].contiguous()       # duplicates: exact ties
    p1 = torch.rand(B, M, D, generator=g)
    p1[:, ::3] = p2[:, : p1[:, ::3].shape[1]]                                            # zero distances
    p2[:, 1::7] *= 3e19                                                                  # (3e19)^2 overflows fp32
    d_o, i_o, _ = native.knn_points(p1, p2, K=K)
    d, i, _ = ops.knn_points(p1.to(DEV), p2.to(DEV), K=K)
    assert torch.equal(i.cpu(), i_o) and torch.equal(d.cpu(), d_o)
    # the same through views that start 4 bytes off a 16-byte boundary
    q1 = torch.empty(p1.numel() + 1, device=DEV)[1:].view_as(p1).copy_(p1)
    q2 = torch.empty(p2.numel() + 1, device=DEV)[1:].view_as(p2).copy_(p2)
    d, i, _ = ops.knn_points(q1, q2, K=K)
    assert torch.equal(i.cpu(), i_o) and torch.equal(d.cpu(), d_o)


def test_knn_fused_query_gather_and_knn_gather():
    xyz = _clouds("lidar", 2, 4096, 5)
    fidx = native.fps(xyz, 256)
    q = xyz[torch.arange(2)[:, None], fidx.long()]
    _, i_o, _ = native.knn_points(q, xyz, K=16)
    idx, q_out = engine.knn_idx(None, xyz.to(DEV), 16, q_idx=fidx.to(DEV))
    assert torch.equal(idx.cpu().long(), i_o) and torch.equal(q_out.cpu(), q)
    x = torch.rand(2, 4096, 11)
    got = ops.knn_gather(x.to(DEV), i_o.to(DEV))
    assert torch.equal(got.cpu(), native.knn_gather(x, i_o))


def test_full_size_properties():
    """BASELINE config sizes (16384 points): size-independent invariants."""
    xyz = _clouds("lidar", 4, 16384, 99).to(DEV)
    idx = ops.furthest_point_sample(xyz, 1024)
    assert (idx[:, 0] == 0).all()
    for b in range(4):
        assert idx[b].unique().numel() == 1024                     # no repeats on duplicate-free clouds
    # prefix property: FPS to M' < M is the prefix of FPS to M
    assert torch.equal(ops.furthest_point_sample(xyz, 300), idx[:, :300])
    # translating the cloud by a power-of-two-exact offset leaves squared distances (hence picks) unchanged
    q = xyz[torch.arange(4, device=DEV)[:, None], idx.long()]
    d, i, nn = ops.knn_points(q, xyz, K=64, return_nn=True)
    assert (d[..., 0] == 0).all() and (i[..., 0] == idx.long()).all()   # a sampled point is its own nearest neighbour
    assert (d[..., 1:] >= d[..., :-1]).all()
    assert torch.equal(nn, ops.knn_gather(xyz, i))
    # kNN is idempotent under re-query with K' < K (prefix)
    d2, i2, _ = ops.knn_points(q, xyz, K=16)
    assert torch.equal(i2, i[..., :16]) and torch.equal(d2, d[..., :16])


@pytest.mark.parametrize("B,M,N,D,K", [(2, 300, 2048, 3, 16), (2, 64, 200, 3, 8), (1, 40, 96, 32, 8)])
def test_knn_points_and_knn_gather_backward(B, M, N, D, K):
    """The pytorch3d stand-ins are differentiable like the originals: gradients of the squared distances
    w.r.t. both clouds, of `knn` w.r.t. p2 and of knn_gather w.r.t. x, against fp64 autograd of the same formulas on
    the CPU (the neighbour indices are the forward's, bit-exact)."""
    g = torch.Generator().manual_seed(B * M + K)
    p1 = torch.randn(B, M, D, generator=g)
    p2 = torch.randn(B, N, D, generator=g)
    wd, wn = torch.randn(B, M, K, generator=g), torch.randn(B, M, K, D, generator=g)
    a1, a2 = p1.to(DEV).requires_grad_(True), p2.to(DEV).requires_grad_(True)
    d, i, nn = ops.knn_points(a1, a2, K=K, return_nn=True)
    ((d * wd.to(DEV)).sum() + (nn * wn.to(DEV)).sum()).backward()
    c1, c2 = p1.double().requires_grad_(True), p2.double().requires_grad_(True)
    bidx = torch.arange(B)[:, None, None]
    nn_c = c2[bidx, i.cpu()]
    d_c = ((c1[:, :, None, :] - nn_c) ** 2).sum(-1)
    ((d_c * wd.double()).sum() + (nn_c * wn.double()).sum()).backward()
    assert torch.allclose(d.detach().cpu().double(), d_c.detach(), rtol=1e-5, atol=1e-6)
    assert torch.allclose(a1.grad.cpu().double(), c1.grad, rtol=1e-4, atol=1e-4)
    assert torch.allclose(a2.grad.cpu().double(), c2.grad, rtol=1e-4, atol=1e-4)
    # knn_gather alone
    x = torch.randn(B, N, 7, generator=g)
    w = torch.randn(B, M, K, 7, generator=g)
    ax = x.to(DEV).requires_grad_(True)
    (ops.knn_gather(ax, i) * w.to(DEV)).sum().backward()
    cx = x.double().requires_grad_(True)
    (cx[bidx, i.cpu()] * w.double()).sum().backward()
    assert torch.allclose(ax.grad.cpu().double(), cx.grad, rtol=1e-4, atol=1e-4)
    # no gradient requested -> plain forward, indices unchanged
    d2, i2, _ = ops.knn_points(p1.to(DEV), p2.to(DEV), K=K)
    assert torch.equal(i2, i) and torch.equal(d2, d.detach())
