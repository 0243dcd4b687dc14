"""CPU: the C oracle of the native ops (oracle/native_ops.c) against independent restatements."""
import numpy as np
import pytest
import torch

from oracle import native


def _kernel_sim_fps(xyz, npoint, w=None):
    """Lock-step python simulation of the reference CUDA block (furthest_point_sampling_gpu.cu:84-206): T threads,
    strided per-thread scan with strict '>', then the halving tree of __update (:75-80)."""
    xyz = xyz.numpy().astype(np.float32)
    N = xyz.shape[0]
    T = native.opt_n_threads(N)
    temp = np.full(N, 1e10, np.float32)
    idx = [0]
    old = 0
    f32 = np.float32
    for _ in range(1, npoint):
        best = np.full(T, -1.0, np.float32)
        besti = np.zeros(T, np.int64)
        for tid in range(T):
            for k in range(tid, N, T):
                dx, dy, dz = (xyz[k] - xyz[old]).astype(np.float32)
                t = f32(dy * dy)
                t = f32(np.float64(dx) * np.float64(dx) + np.float64(t))      # fma: exact product, one rounding
                d = f32(np.float64(dz) * np.float64(dz) + np.float64(t))
                if w is not None:
                    d = f32(f32(w[k]) * d)
                d2 = min(d, temp[k])
                temp[k] = d2
                if d2 > best[tid]:
                    best[tid], besti[tid] = d2, k
        s = T // 2
        while s >= 1:
            for tid in range(s):
                v1, v2 = best[tid], best[tid + s]
                i1, i2 = besti[tid], besti[tid + s]
                best[tid] = max(v1, v2)
                besti[tid] = i2 if v2 > v1 else i1
            s //= 2
        old = int(besti[0])
        idx.append(old)
    return np.array(idx, np.int32)


@pytest.mark.parametrize("n,m,grid", [(64, 32, False), (96, 40, True), (37, 37, True), (130, 50, True)])
def test_fps_matches_kernel_simulation(n, m, grid):
    g = torch.Generator().manual_seed(n)
    if grid:   # integer lattice: many exactly equal distances -> exercises the block-reduction tie-break
        xyz = torch.randint(0, 4, (n, 3), generator=g).float()
    else:
        xyz = torch.randn(n, 3, generator=g)
    got = native.fps(xyz[None], m)[0].numpy()
    assert np.array_equal(got, _kernel_sim_fps(xyz, m))


def test_weighted_fps_matches_kernel_simulation():
    g = torch.Generator().manual_seed(5)
    xyz = torch.randint(0, 3, (80, 3), generator=g).float()
    w = torch.randint(1, 3, (80,), generator=g).float()
    got = native.fps(xyz[None], 30, w[None])[0].numpy()
    assert np.array_equal(got, _kernel_sim_fps(xyz, 30, w.numpy()))


def test_fps_tiebreak_is_bit_reversed_thread_order():
    """All points identical: every min-distance is 0 after the first pick, so each pick is decided purely by the
    reduction order -> with T=8 threads the winner is thread 0 first-k... the closed form says key
    (bitrev(k mod T), k div T) is minimised, i.e. index 0 forever."""
    xyz = torch.zeros(1, 8, 3)
    assert native.fps(xyz, 5)[0].tolist() == [0, 0, 0, 0, 0]
    # two far clusters of duplicates: after picking 0 the farthest are indices 4..7 (all equal);
    # T=8 -> thread ids 4,5,6,7 -> bit-reversed 1,5,3,7 -> thread 4 wins
    xyz = torch.zeros(1, 8, 3)
    xyz[0, 4:] = 1.0
    assert native.fps(xyz, 2)[0].tolist() == [0, 4]
    # farthest candidates at indices 1,2,3 (threads 1,2,3 -> bitrev3 = 4,2,6) -> index 2 wins, not the lowest index
    xyz = torch.zeros(1, 8, 3)
    xyz[0, 1:4] = 1.0
    assert native.fps(xyz, 2)[0].tolist() == [0, 2]


def test_fps_edge_cases():
    xyz = torch.randn(2, 50, 3)
    assert native.fps(xyz, 1).tolist() == [[0], [0]]
    idx = native.fps(xyz, 50)
    assert sorted(idx[0].tolist()) == list(range(50))          # all points picked exactly once
    temp = torch.full((2, 50), 1e10)
    native.fps(xyz, 10, None, temp)
    assert float(temp.max()) < 1e9                              # scratch clobbered with the min-distances


def test_opt_n_threads():
    for n, t in [(1, 1), (2, 2), (3, 2), (511, 256), (512, 512), (1023, 512), (1024, 1024), (16384, 1024), (8096, 1024)]:
        assert native.opt_n_threads(n) == t


@pytest.mark.parametrize("D,K", [(3, 8), (3, 64), (16, 5)])
def test_knn_matches_bruteforce_with_ties(D, K):
    g = torch.Generator().manual_seed(D * 100 + K)
    p1 = torch.randint(0, 5, (2, 40, D), generator=g).float()
    p2 = torch.randint(0, 5, (2, 150, D), generator=g).float()       # lattice -> many equal distances
    d, i, nn = native.knn_points(p1, p2, K=K, return_nn=True)
    d_ref, i_ref = native.knn_bruteforce_numpy(p1, p2, K)
    assert np.array_equal(i.numpy(), i_ref)
    assert np.array_equal(d.numpy(), d_ref)
    assert torch.equal(nn, p2[torch.arange(2)[:, None, None], i])
    assert (d[..., 1:] >= d[..., :-1]).all()


def test_knn_random_float_inputs():
    g = torch.Generator().manual_seed(1)
    p1, p2 = torch.randn(1, 64, 3, generator=g) * 30, torch.randn(1, 500, 3, generator=g) * 30
    d, i, _ = native.knn_points(p1, p2, K=16)
    full = ((p1[:, :, None, :] - p2[:, None, :, :]) ** 2).sum(-1)
    assert torch.equal(i, full.topk(16, dim=-1, largest=False)[1]) or \
        torch.allclose(d, full.topk(16, dim=-1, largest=False)[0], rtol=1e-5)


def test_gather_and_grad():
    g = torch.Generator().manual_seed(2)
    pts = torch.randn(2, 5, 30, generator=g)
    idx = torch.randint(0, 30, (2, 12), generator=g).int()
    out = native.gather_points(pts, idx)
    assert torch.equal(out, torch.gather(pts, 2, idx.long()[:, None, :].expand(-1, 5, -1)))
    go = torch.randn(2, 5, 12, generator=g)
    ref = torch.zeros(2, 5, 30).scatter_add_(2, idx.long()[:, None, :].expand(-1, 5, -1), go)
    assert torch.allclose(native.gather_points_grad(go, idx, 30), ref, atol=1e-6)
