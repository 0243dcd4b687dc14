"""CPU model of the two top-K selection schemes of the descriptor-space search (csrc/knn.cu `knnd_kernel`, csrc/knn_select.cuh):
the running unsorted set with replace-the-maximum (WarpSet) and the sorted carried set merged tile by tile with K
extract-min rounds (K <= 16 of N <= 512).  Both must return the K smallest members of

    {(dist, n) : dist not NaN}  U  {(+inf, 0x7fffffff - pos) : pos < K}        ordered by (distance bits, index)

with the surviving padding members mapped to point `pos` -- the contract the GPU tests pin against oracle/native_ops.c
(`oracle_knn`: squared L2, ascending, ties by index; reference call site models/HRegNet/layers.py:278).  Lane-level
emulation in numpy: what the kernel does per lane and per REDUX, not a re-statement of "sort and take K"."""
import numpy as np
import pytest

TAKEN = 0x7FFFFFFF
INF_BITS = 0x7F800000


def _bits(x):
    return np.float32(x).view(np.int32).item()


def _contract(d, K):
    N = len(d)
    members = [(_bits(v), n) for n, v in enumerate(d) if not np.isnan(v)] + [(INF_BITS, TAKEN - p) for p in range(K)]
    members.sort()
    out = []
    for kb, i in members[:K]:
        if i >= N:                      # sanitize(): padding member 0x7fffffff - p -> point p
            p = TAKEN - i
            i = p if 0 <= p < N else 0
        out.append((kb, i))
    return out


def _extract_min(d, K):
    """knnd_kernel's xsel path: 32 lanes, tiles of 128 references (lane l owns n = t0 + l + 32 j), carried set of rank r in
    lane 32 - K + r."""
    N = len(d)
    car_k = [INF_BITS if l >= 32 - K else TAKEN for l in range(32)]
    car_i = [TAKEN - (31 - l) if l >= 32 - K else TAKEN for l in range(32)]
    for t0 in range(0, N, 128):
        ck = [[TAKEN] * 4 for _ in range(32)]
        for l in range(32):
            for j in range(4):
                n = t0 + l + 32 * j
                if n < N and not np.isnan(d[n]):
                    ck[l][j] = _bits(d[n])

        def local_best(l):
            k, i = car_k[l], car_i[l]
            for j in range(4):
                n = t0 + l + 32 * j
                if ck[l][j] < k or (ck[l][j] == k and n < i):
                    k, i = ck[l][j], n
            return k, i

        best = [local_best(l) for l in range(32)]
        nk, ni = [TAKEN] * 32, [TAKEN] * 32
        for r in range(K):
            mk = min(b[0] for b in best)                                   # REDUX.MIN
            wi = min(b[1] if b[0] == mk else TAKEN for b in best)           # REDUX.MIN
            assert mk != TAKEN
            nk[32 - K + r], ni[32 - K + r] = mk, wi
            for l in range(32):
                if best[l] == (mk, wi):
                    if (car_k[l], car_i[l]) == (mk, wi):
                        car_k[l] = TAKEN
                    for j in range(4):
                        if ck[l][j] == mk and t0 + l + 32 * j == wi:
                            ck[l][j] = TAKEN
                    best[l] = local_best(l)
        car_k, car_i = nk, ni
    out = []
    for l in range(32 - K, 32):
        i = car_i[l]
        if i >= N:
            p = TAKEN - i
            i = p if 0 <= p < N else 0
        out.append((car_k[l], i))
    return out


def _replace_max(d, K):
    """WarpSet: unsorted set, a candidate below the largest member (dist, index) replaces it; sorted at the end."""
    N = len(d)
    S = [(INF_BITS, TAKEN - p) for p in range(K)]
    for n, v in enumerate(d):
        if np.isnan(v):
            continue                                                        # cand_less is false for NaN
        c = (_bits(v), n)
        m = max(S)
        if c < m:
            S[S.index(m)] = c
    out = []
    for kb, i in sorted(S):
        if i >= N:
            p = TAKEN - i
            i = p if 0 <= p < N else 0
        out.append((kb, i))
    return out


@pytest.mark.parametrize("N,K,kind", [(256, 8, "rand"), (256, 8, "ties"), (300, 16, "ties"), (512, 16, "inf"), (130, 5, "nan"),
                                      (16, 16, "inf"), (40, 8, "fewvalid"), (129, 1, "ties"), (512, 3, "rand")])
def test_extract_min_equals_replace_max_equals_contract(N, K, kind):
    rng = np.random.default_rng(N * 31 + K)
    for _ in range(6):
        d = rng.random(N).astype(np.float32) * 4
        if kind == "ties":
            d = np.round(d * 2) / 2                                         # 9 distinct values: order decided by the index
        if kind == "inf":
            d[rng.random(N) < 0.6] = np.inf
        if kind == "nan":
            d[rng.random(N) < 0.3] = np.nan
        if kind == "fewvalid":
            d[:] = np.nan
            d[rng.choice(N, K // 2, replace=False)] = rng.random(K // 2).astype(np.float32)
        want = _contract(d, K)
        assert _extract_min(d, K) == want
        assert _replace_max(d, K) == want
