"""ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/native_ops.c header).

ctypes front-end of oracle/_build/liboracle_native.so plus the torch-CPU versions of the two
third-party ops the reference takes from pytorch3d (knn_points / knn_gather, models/HRegNet/layers.py:7).
Everything here runs on the CPU and takes / returns CPU torch tensors.

Nothing under pcd_reg_hregnet_b200/ may import this module.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle_native.so")
_lib = None


def build() -> str:
    """Compile the C oracle (gcc, seconds).  Returns the .so path."""
    src = os.path.join(_HERE, "native_ops.c")
    if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        fp, ip, lp = ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p
        L.oracle_opt_n_threads.argtypes = [ctypes.c_int]
        L.oracle_opt_n_threads.restype = ctypes.c_int
        L.oracle_fps.argtypes = [fp, fp, fp, ip] + [ctypes.c_int] * 3
        L.oracle_gather_points.argtypes = [fp, ip, fp] + [ctypes.c_int] * 4
        L.oracle_gather_points_grad.argtypes = [fp, ip, fp] + [ctypes.c_int] * 4
        L.oracle_knn.argtypes = [fp, fp, fp, lp, fp] + [ctypes.c_int] * 5
        for f in (L.oracle_fps, L.oracle_gather_points, L.oracle_gather_points_grad, L.oracle_knn):
            f.restype = None
        _lib = L
    return _lib


def _f32(t):
    assert t.dtype == torch.float32 and t.device.type == "cpu"
    return t.contiguous()


def opt_n_threads(n: int) -> int:
    return lib().oracle_opt_n_threads(int(n))


def fps(xyz, npoint, weights=None, temp=None):
    """xyz [B,N,3] -> idx [B,npoint] int32.  temp defaults to 1e10 (models/utils.py:25)."""
    xyz = _f32(xyz)
    B, N, _ = xyz.shape
    if temp is None:
        temp = torch.full((B, N), 1e10, dtype=torch.float32)
    assert temp.is_contiguous()
    idx = torch.zeros(B, npoint, dtype=torch.int32)
    w = _f32(weights) if weights is not None else None
    lib().oracle_fps(xyz.data_ptr(), w.data_ptr() if w is not None else None, temp.data_ptr(),
                     idx.data_ptr(), B, N, npoint)
    return idx


def gather_points(points, idx):
    """points [B,C,N], idx [B,M] int32 -> [B,C,M]."""
    points = _f32(points)
    idx = idx.contiguous()
    assert idx.dtype == torch.int32
    B, C, N = points.shape
    M = idx.shape[1]
    out = torch.empty(B, C, M, dtype=torch.float32)
    lib().oracle_gather_points(points.data_ptr(), idx.data_ptr(), out.data_ptr(), B, C, N, M)
    return out


def gather_points_grad(grad_out, idx, N):
    grad_out = _f32(grad_out)
    idx = idx.contiguous()
    B, C, M = grad_out.shape
    g = torch.zeros(B, C, N, dtype=torch.float32)
    lib().oracle_gather_points_grad(grad_out.data_ptr(), idx.data_ptr(), g.data_ptr(), B, C, N, M)
    return g


def knn_points(p1, p2, K=1, return_nn=False, **_):
    """Stand-in for pytorch3d.ops.knn_points: (dists [B,M,K] squared ascending, idx int64, nn|None)."""
    p1, p2 = _f32(p1), _f32(p2)
    B, M, D = p1.shape
    N = p2.shape[1]
    assert K <= N
    d = torch.empty(B, M, K, dtype=torch.float32)
    i = torch.empty(B, M, K, dtype=torch.int64)
    nn = torch.empty(B, M, K, D, dtype=torch.float32) if return_nn else None
    lib().oracle_knn(p1.data_ptr(), p2.data_ptr(), d.data_ptr(), i.data_ptr(),
                     nn.data_ptr() if return_nn else None, B, M, N, D, K)
    return d, i, nn


def knn_gather(x, idx, lengths=None):
    """Stand-in for pytorch3d.ops.knn_gather: x [B,N,U], idx [B,M,K] -> [B,M,K,U]."""
    B, N, U = x.shape
    _, M, K = idx.shape
    return x[torch.arange(B)[:, None, None], idx]


def knn_bruteforce_numpy(p1, p2, K):
    """Independent (slow, fp32-faithful) cross-check of oracle_knn for tiny cases: explicit python loops over
    the D accumulation with np.float32 fma emulated in float64 (exact for one fma: product of two floats
    is exact in double, the sum is rounded once to double then to float -- double rounding can differ
    from a true fma only in <2^-29 of cases; tests use it on small integers-grid inputs where it is exact)."""
    p1 = p1.numpy().astype(np.float32)
    p2 = p2.numpy().astype(np.float32)
    B, M, D = p1.shape
    N = p2.shape[1]
    dist = np.zeros((B, M, N), np.float32)
    for d in range(D):
        diff = (p1[:, :, None, d] - p2[:, None, :, d]).astype(np.float32)
        dist = (diff.astype(np.float64) * diff.astype(np.float64) + dist.astype(np.float64)).astype(np.float32)
    order = np.lexsort((np.broadcast_to(np.arange(N), dist.shape), dist), axis=-1)[..., :K]
    return np.take_along_axis(dist, order, -1), order.astype(np.int64)
