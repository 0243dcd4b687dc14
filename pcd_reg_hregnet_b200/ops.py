"""Drop-in point ops: same names, argument meaning, dtypes and error behaviour as the reference's
`models/utils.py:14-89` (FurthestPointSampling / WeightedFurthestPointSampling / GatherOperation) and as the
two pytorch3d functions the reference imports (`knn_points`, `knn_gather`, models/HRegNet/layers.py:7).

All of them run hand-written sm_100a kernels through the C ABI (include/hregnet_b200.h); CPU tensors raise.
"""
from collections import namedtuple

import torch
from torch.autograd import Function

from . import _lib
from ._lib import call, ptr, stream


class FurthestPointSampling(Function):
    """xyz [B,N,3] float32 contiguous, npoint -> int32 [B,npoint]   (models/utils.py:14-34)."""

    @staticmethod
    def forward(ctx, xyz: torch.Tensor, npoint: int) -> torch.Tensor:
        assert xyz.is_contiguous()
        B, N, _ = xyz.size()
        output = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
        # temp=NULL: min-distances start at 1e10 (utils.py:25) and never leave the SM; large clouds need scratch
        temp = None if N <= 16384 else torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
        call("hrn_fps", ptr(xyz), None, ptr(temp), ptr(output), B, N, npoint, stream())
        ctx.mark_non_differentiable(output)
        return output

    @staticmethod
    def backward(ctx, a=None):
        return None, None


furthest_point_sample = FurthestPointSampling.apply


class WeightedFurthestPointSampling(Function):
    """xyz [B,N,3], weights [B,N] -> int32 [B,npoint]   (models/utils.py:36-58)."""

    @staticmethod
    def forward(ctx, xyz: torch.Tensor, weights: torch.Tensor, npoint: int) -> torch.Tensor:
        assert xyz.is_contiguous()
        assert weights.is_contiguous()
        B, N, _ = xyz.size()
        output = torch.empty(B, npoint, dtype=torch.int32, device=xyz.device)
        temp = None if N <= 16384 else torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
        call("hrn_fps", ptr(xyz), ptr(weights), ptr(temp), ptr(output), B, N, npoint, stream())
        ctx.mark_non_differentiable(output)
        return output

    @staticmethod
    def backward(ctx, a=None):
        return None, None, None


weighted_furthest_point_sample = WeightedFurthestPointSampling.apply


class GatherOperation(Function):
    """features [B,C,N], idx [B,npoint] int32 -> [B,C,npoint]; differentiable w.r.t. features (utils.py:60-89)."""

    @staticmethod
    def forward(ctx, features: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        assert features.is_contiguous()
        assert idx.is_contiguous()
        B, npoint = idx.size()
        _, C, N = features.size()
        output = torch.empty(B, C, npoint, dtype=torch.float32, device=features.device)
        call("hrn_gather_points", ptr(features), ptr(idx), ptr(output), B, C, N, npoint, stream())
        ctx.for_backwards = (idx, C, N)
        return output

    @staticmethod
    def backward(ctx, grad_out):
        idx, C, N = ctx.for_backwards
        B, npoint = idx.size()
        grad_features = torch.zeros(B, C, N, dtype=torch.float32, device=grad_out.device)
        grad_out_data = grad_out.data.contiguous()
        call("hrn_gather_points_grad", ptr(grad_out_data), ptr(idx), ptr(grad_features), B, C, N, npoint, stream())
        return grad_features, None


gather_operation = GatherOperation.apply

_KNN = namedtuple("KNN", "dists idx knn")


def _knn_forward(p1, p2, K):
    B, M, D = p1.shape
    N = p2.shape[1]
    dists = torch.empty(B, M, K, dtype=torch.float32, device=p1.device)
    idx = torch.empty(B, M, K, dtype=torch.int64, device=p1.device)
    if D == 3 and 256 <= N <= 32768:       # spatially culled exact search (same results, see csrc/knn_sorted.cu)
        from .engine import knn_scratch
        pts, boxes = knn_scratch(B, N, p1.device)
        call("hrn_knn3_sorted", ptr(p1), None, ptr(p2), B, M, N, K, ptr(pts), ptr(boxes), ptr(dists), ptr(idx), None, None,
             None, stream())
    else:
        call("hrn_knn", ptr(p1), None, ptr(p2), B, M, N, D, K, ptr(dists), ptr(idx), None, None, None, stream())
    return dists, idx


class _KnnPoints(Function):
    """dists / idx of the exact search; differentiable w.r.t. both clouds through the squared distances
    (d dists / d p1 = 2 (p1 - p2[idx]), scattered with the opposite sign into p2), like pytorch3d's _knn_points."""

    @staticmethod
    def forward(ctx, p1, p2, K):
        dists, idx = _knn_forward(p1, p2, K)
        ctx.save_for_backward(p1, p2, idx)
        ctx.mark_non_differentiable(idx)
        return dists, idx

    @staticmethod
    def backward(ctx, grad_dists, grad_idx):
        p1, p2, idx = ctx.saved_tensors
        B, M, D = p1.shape
        N, K = p2.shape[1], idx.shape[2]
        g = grad_dists.contiguous().float()
        gp1 = torch.empty_like(p1) if ctx.needs_input_grad[0] else None
        gp2 = torch.zeros_like(p2) if ctx.needs_input_grad[1] else None
        if gp1 is not None or gp2 is not None:
            call("hrn_knn_dists_grad", ptr(p1), ptr(p2), ptr(idx), ptr(g), ptr(gp1), ptr(gp2), B, M, N, D, K, stream())
        return gp1, gp2, None


class _KnnGather(Function):
    """x [B,N,U], idx [B,M,K] int64 -> [B,M,K,U]; differentiable w.r.t. x (scatter-add)."""

    @staticmethod
    def forward(ctx, x, idx):
        B, N, U = x.shape
        _, M, K = idx.shape
        out = torch.empty(B, M, K, U, dtype=torch.float32, device=x.device)
        call("hrn_knn_gather", ptr(x), ptr(idx), ptr(out), B, N, M, K, U, stream())
        ctx.save_for_backward(idx)
        ctx.dims = (B, N, M, K, U)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        idx, = ctx.saved_tensors
        B, N, M, K, U = ctx.dims
        g = grad_out.contiguous().float()
        gx = torch.zeros(B, N, U, dtype=torch.float32, device=g.device)
        call("hrn_knn_gather_grad", ptr(g), ptr(idx), ptr(gx), B, N, M, K, U, stream())
        return gx, None


def knn_points(p1, p2, lengths1=None, lengths2=None, norm: int = 2, K: int = 1, version: int = -1,
               return_nn: bool = False, return_sorted: bool = True):
    """pytorch3d.ops.knn_points stand-in: (dists [B,M,K] squared, ascending; idx int64; nn [B,M,K,D] | None).

    Deterministic order (dist asc, index asc).  Differentiable like pytorch3d's: `dists` w.r.t. p1 and p2, `knn` w.r.t.
    p2 (it is knn_gather(p2, idx)).  `lengths*` (ragged batches) and norm != 2 are not used by the reference
    (layers.py:20,278,316,322,434) and are rejected."""
    if lengths1 is not None or lengths2 is not None or norm != 2:
        raise NotImplementedError("ragged batches / L1 norm are not part of the HRegNet path")
    p1 = p1.contiguous()
    p2 = p2.contiguous()
    if p1.dtype != torch.float32 or p2.dtype != torch.float32:
        raise TypeError("knn_points expects float32")
    dists, idx = _KnnPoints.apply(p1, p2, K)
    nn = _KnnGather.apply(p2, idx) if return_nn else None
    return _KNN(dists, idx, nn)


def knn_gather(x, idx, lengths=None):
    """pytorch3d.ops.knn_gather stand-in: x [B,N,U], idx [B,M,K] int64 -> [B,M,K,U]; differentiable w.r.t. x."""
    if lengths is not None:
        raise NotImplementedError("ragged batches are not part of the HRegNet path")
    return _KnnGather.apply(x.contiguous(), idx.contiguous())


class _PointUtilsShim:
    """`point_utils_cuda`-compatible module object: the four positional entry points of the reference's pybind
    module (models/PointUtils/src/point_utils_api.cpp:6-13) on top of the new kernels, so the UNMODIFIED
    reference `models/utils.py` runs on them:  sys.modules['point_utils_cuda'] = ops.point_utils_cuda"""

    __name__ = "point_utils_cuda"

    @staticmethod
    def furthest_point_sampling_wrapper(b, n, m, points, temp, idx):
        call("hrn_fps", ptr(points), None, ptr(temp), ptr(idx), b, n, m, stream())
        return 1

    @staticmethod
    def weighted_furthest_point_sampling_wrapper(b, n, m, points, weights, temp, idx):
        call("hrn_fps", ptr(points), ptr(weights), ptr(temp), ptr(idx), b, n, m, stream())
        return 1

    @staticmethod
    def gather_points_wrapper(b, c, n, npoints, points, idx, out):
        call("hrn_gather_points", ptr(points), ptr(idx), ptr(out), b, c, n, npoints, stream())
        return 1

    @staticmethod
    def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
        call("hrn_gather_points_grad", ptr(grad_out), ptr(idx), ptr(grad_points), b, c, n, npoints, stream())
        return 1


point_utils_cuda = _PointUtilsShim()
