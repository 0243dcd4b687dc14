"""CPU emulation of the C-ABI kernels (TEST INFRASTRUCTURE): each `hrn_*` entry point restated with torch on
the CPU from its documented contract (include/hregnet_b200.h).  Patched into pcd_reg_hregnet_b200.engine so
that the HOST-SIDE orchestration (segment order, weight-column permutation, index plumbing, BN folding) can be
checked against the reference graph in the build container, which has no GPU.  It is never used by the product.
"""
import contextlib
import ctypes

import torch

from oracle import native
from pcd_reg_hregnet_b200 import _lib, engine


def _act(x, act):
    if act == _lib.ACT_RELU:
        return torch.relu(x)
    if act == _lib.ACT_SOFTPLUS_EPS:
        return torch.nn.functional.softplus(x) + 0.001
    if act == _lib.ACT_SIGMOID:
        return torch.sigmoid(x)
    return x


def _rows_matrix(view):
    cols = []
    r = torch.arange(view.rows)
    for mat, mode, ch, col0, scale in view.segs:
        if mode == _lib.SEG_DIRECT:
            src = r
        elif mode == _lib.SEG_BROADCAST:
            src = r // view.c.group
        else:
            b = r // view.c.rows_per_batch
            src = b * view.c.src_rows_per_batch + view.gather_idx.reshape(-1).long()
        x = mat[src, col0:col0 + ch]
        if scale is not None:
            x = x * scale.reshape(-1, 1)
        cols.append(x)
    return torch.cat(cols, dim=1)


def _launch_layer(view, W, b, act, out):
    out.copy_(_act(_rows_matrix(view) @ W.t() + (b if b is not None else 0), act))


def _fps(xyz, w, temp, idx, B, N, M, st):
    idx.copy_(native.fps(xyz, M, w, temp))


def _knn(p1, q_idx, p2, B, M, N, D, K, dists, idx64, idx32, nn, q_out, st):
    if q_idx is not None:
        p1 = p2[torch.arange(B)[:, None], q_idx.long()].contiguous()
        if q_out is not None:
            q_out.copy_(p1)
    d, i, n = native.knn_points(p1, p2, K=K, return_nn=nn is not None)
    if dists is not None: dists.copy_(d)
    if idx64 is not None: idx64.copy_(i)
    if idx32 is not None: idx32.copy_(i.int())
    if nn is not None: nn.copy_(n)


def _knn3_sorted(p1, q_idx, p2, B, M, N, K, pts, boxes, dists, idx64, idx32, nn, q_out, st):
    _knn(p1, q_idx, p2, B, M, N, 3, K, dists, idx64, idx32, nn, q_out, st)       # same contract, same results


def _knn3_sort(p2, B, N, pts, boxes, st):
    pass                                                                          # the emulated search needs no scratch


def _knn3_search(p1, q_idx, p2, B, M, N, K, pts, boxes, dists, idx64, idx32, nn, q_out, st):
    _knn(p1, q_idx, p2, B, M, N, 3, K, dists, idx64, idx32, nn, q_out, st)


def _gather_rows(x, idx, out, B, N, M, U, st):
    out.view(B, M, U).copy_(x.view(B, N, U)[torch.arange(B)[:, None], idx.view(B, M).long()])


def _transpose(x, out, B, R, C, st):
    out.view(B, C, R).copy_(x.view(B, R, C).transpose(1, 2))


def _group_geometry(q, p, idx, wq, wp, B, M, k, N, out, ldo, nn, st):
    i = idx.view(B, M, k).long()
    b = torch.arange(B)[:, None, None]
    pn = p.view(B, N, 3)[b, i]
    qq = q.view(B, M, 1, 3).expand(-1, -1, k, -1)
    rel = pn - qq
    o = out.view(B, M, k, ldo)
    o[..., 0:3] = rel
    o[..., 3] = torch.sqrt((rel * rel).sum(-1))
    if wq is not None:
        o[..., 4:7] = qq
        o[..., 7:10] = pn
        o[..., 10] = wq.view(B, M, 1).expand(-1, -1, k)
        o[..., 11] = wp.view(B, N)[b, i]
    if nn is not None:
        nn.view(B, M, k, 3).copy_(pn)


def _group_attention(E, ldE, C, groups, k, a, st):
    a.copy_(torch.softmax(E.view(groups, k, -1)[..., :C].max(dim=-1)[0], dim=-1).reshape(-1))


def _group_weighted_sum(a, V, ldV, C, groups, k, idx, gpb, N, out, ldo, st):
    if idx is None:
        v = V.view(groups, k, -1)[..., :C]
    else:
        r = torch.arange(groups * k)
        b = (r // k) // gpb
        v = V[b * N + idx.reshape(-1).long(), :C].view(groups, k, C)
    out.copy_((a.view(groups, k, 1) * v).sum(dim=1))


def _group_attend(E, ldE, C, groups, k, a, af, ldaf, xyz, idx, gpb, N, cor, st):
    Ek = E.view(groups, k, -1)[..., :C]
    w = torch.softmax(Ek.max(dim=-1)[0], dim=-1)
    if a is not None:
        a.copy_(w.reshape(-1))
    af.copy_((w.unsqueeze(-1) * Ek).sum(dim=1))
    if cor is not None:
        b = (torch.arange(groups * k) // k) // gpb
        v = xyz.reshape(-1, 3)[b * N + idx.reshape(-1).long()].view(groups, k, 3)
        cor.copy_((w.unsqueeze(-1) * v).sum(dim=1))


def _group_max(X, ldX, C, groups, k, out, ldo, st):
    out.copy_(X.view(groups, k, -1)[..., :C].max(dim=1)[0])


def _sigma_to_weights(sig, w, B, M, st):
    x = 1.0 / (sig + 1e-5)
    w.copy_(x / x.mean(dim=1, keepdim=True))


def _transform_points(x, R, t, out, B, N, st):
    out.copy_(torch.einsum("bij,bnj->bni", R.view(B, 3, 3), x) + t.view(B, 1, 3))


def _cosine_matrix(S, D, B, N1, N2, C, nS, nD, cosm, rowmax, colmax, st):
    nS.copy_(S.norm(dim=-1)); nD.copy_(D.norm(dim=-1))
    c = torch.einsum("bnc,bmc->bnm", D, S) / (nD[:, :, None] * nS[:, None, :] + 1e-6)
    cosm.copy_(c)
    rowmax.copy_(c.max(dim=2)[0]); colmax.copy_(c.max(dim=1)[0])


def _cosine_pick(cosm, rowmax, colmax, idx, B, N1, N2, k, out, ldo, c_sd, c_ds, st):
    i = idx.view(B, N1, k).long()
    b = torch.arange(B)[:, None, None]
    n1 = torch.arange(N1)[None, :, None]
    c = cosm[b, i, n1]
    o = out.view(B, N1, k, ldo)
    o[..., c_sd] = c / (colmax[b, n1] + 1e-6)
    o[..., c_ds] = c / (rowmax[b, i] + 1e-6)


def _weighted_kabsch(src, cor, w, B, N, Rp, tp, R, t, Rc, tc, p12, st):
    """fp64 sums + the library's own host closed form (hrn_pose_from_covariance_host)."""
    L = _lib.lib()
    for b in range(B):
        wn = (w[b] / (w[b].double().sum().float() + 1e-4)).double()
        den = float((wn.sum().float() + 1e-4))
        xb = ((wn[:, None] * src[b].double()).sum(0) / den).float()
        yb = ((wn[:, None] * cor[b].double()).sum(0) / den).float()
        xc = (src[b] - xb).double()
        yc = (cor[b] - yb).double() * wn[:, None]
        H = (xc.t() @ yc).contiguous()
        xbd, ybd = xb.double().contiguous(), yb.double().contiguous()
        R9 = torch.empty(9, dtype=torch.float64); t3 = torch.empty(3, dtype=torch.float64)
        rc = L.hrn_pose_from_covariance_host(H.data_ptr(), xbd.data_ptr(), ybd.data_ptr(), R9.data_ptr(), t3.data_ptr())
        assert rc == 0
        R[b] = R9.view(3, 3).float(); t[b] = t3.float()
        if Rc is not None:
            Rc[b] = R[b] @ Rp[b]
            tc[b] = R[b] @ tp[b] + t[b]
        if p12 is not None:
            p12[b, :9] = (Rc if Rc is not None else R)[b].reshape(9)
            p12[b, 9:] = (tc if tc is not None else t)[b]


_TABLE = {
    "hrn_fps": _fps, "hrn_knn": _knn, "hrn_knn3_sorted": _knn3_sorted, "hrn_knn3_sort": _knn3_sort, "hrn_knn3_search": _knn3_search, "hrn_gather_rows": _gather_rows, "hrn_transpose": _transpose,
    "hrn_group_geometry": _group_geometry, "hrn_group_attention": _group_attention,
    "hrn_group_weighted_sum": _group_weighted_sum, "hrn_group_max": _group_max, "hrn_group_attend": _group_attend,
    "hrn_sigma_to_weights": _sigma_to_weights, "hrn_transform_points": _transform_points,
    "hrn_cosine_matrix": _cosine_matrix, "hrn_cosine_pick": _cosine_pick, "hrn_weighted_kabsch": _weighted_kabsch,
}


@contextlib.contextmanager
def emulated_kernels():
    saved = (engine.ptr, engine.stream, engine.call, engine._launch_layer_fp32)
    prec = engine.get_precision()
    engine.set_precision("fp32")          # the emulation covers the exact-fp32 orchestration path
    engine.ptr = lambda t: t
    engine.stream = lambda: 0
    engine.call = lambda name, *a: _TABLE[name](*a)
    engine._launch_layer_fp32 = _launch_layer
    try:
        yield
    finally:
        engine.ptr, engine.stream, engine.call, engine._launch_layer_fp32 = saved
        engine.set_precision(prec)
