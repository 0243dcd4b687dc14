"""ctypes loader of libhregnet_b200.so (C ABI: include/hregnet_b200.h).

There is NO fallback: if the library is missing or a call fails, an exception is raised.  Build it with
`python -c "import __graft_entry__ as g; g.build()"` or `make -C pcd_reg_hregnet_b200/csrc`.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhregnet_b200.so")

c_int, c_ll, c_vp = ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p


class Seg(ctypes.Structure):
    _fields_ = [("ptr", c_vp), ("row_scale", c_vp), ("channels", ctypes.c_int32), ("ld", ctypes.c_int32),
                ("col0", ctypes.c_int32), ("mode", ctypes.c_int32)]


class Rows(ctypes.Structure):
    _fields_ = [("seg", Seg * 4), ("gather_idx", c_vp), ("n_seg", ctypes.c_int32), ("group", ctypes.c_int32),
                ("rows_per_batch", ctypes.c_int32), ("src_rows_per_batch", ctypes.c_int32)]


SEG_DIRECT, SEG_BROADCAST, SEG_GATHER = 0, 1, 2
ACT_NONE, ACT_RELU, ACT_SOFTPLUS_EPS, ACT_SIGMOID = 0, 1, 2, 3

# name -> argtypes; every function returns int except hrn_version
SIGNATURES = {
    "hrn_fps": [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp],
    "hrn_gather_points": [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp],
    "hrn_gather_points_grad": [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp],
    "hrn_knn": [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "hrn_knn3_sorted": [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "hrn_knn3_sort": [c_vp, c_int, c_int, c_vp, c_vp, c_vp],
    "hrn_knn3_search": [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "hrn_knn_gather": [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp],
    "hrn_knn_gather_grad": [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp],
    "hrn_knn_dists_grad": [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp],
    "hrn_gather_rows": [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp],
    "hrn_transpose": [c_vp, c_vp, c_int, c_int, c_int, c_vp],
    "hrn_layer_fp32": [ctypes.POINTER(Rows), c_vp, c_vp, c_int, c_vp, c_int, c_ll, c_int, c_vp],
    "hrn_layer_tc": [ctypes.POINTER(Rows), c_vp, c_vp, c_int, c_vp, c_int, c_ll, c_int, c_int, c_int, c_int, c_vp],
    "hrn_layer_tc_groupmax": [ctypes.POINTER(Rows), c_vp, c_vp, c_int, c_vp, c_int, c_ll, c_int, c_int, c_int, c_int, c_int, c_vp],
    "hrn_level_fused": [c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp],
    "hrn_level_ws": [c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp],
    "hrn_chain_tc": [ctypes.POINTER(Rows), c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_vp,
                     c_int, c_vp, c_vp, c_ll, c_int, c_vp, c_vp, c_int, c_vp],
    "hrn_chain_wide": [ctypes.POINTER(Rows), c_vp, c_ll, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_int, c_vp, c_vp,
                       c_ll, c_int, c_vp],
    "hrn_chain_wide_head": [ctypes.POINTER(Rows), c_vp, c_ll, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_ll, c_int, c_vp],
    "hrn_group_attention": [c_vp, c_int, c_int, c_ll, c_int, c_vp, c_vp],
    "hrn_group_weighted_sum": [c_vp, c_vp, c_int, c_int, c_ll, c_int, c_vp, c_int, c_int, c_vp, c_int, c_vp],
    "hrn_group_attend": [c_vp, c_int, c_int, c_ll, c_int, c_vp, c_vp, c_int, c_vp, c_vp, c_int, c_int, c_vp, c_vp],
    "hrn_group_max": [c_vp, c_int, c_int, c_ll, c_int, c_vp, c_int, c_vp],
    "hrn_group_geometry": [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, c_vp, c_vp],
    "hrn_sigma_to_weights": [c_vp, c_vp, c_int, c_int, c_vp],
    "hrn_transform_points": [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp],
    "hrn_cosine_matrix": [c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "hrn_cosine_features_tc": [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_int, c_int, c_int, c_vp],
    "hrn_cosine_pick": [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_int, c_int, c_int, c_vp],
    "hrn_weighted_kabsch": [c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "hrn_regression_head": [c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp],
    "hrn_range_filter": [c_vp, c_vp, c_vp, c_int, c_ll, ctypes.c_float, c_vp, c_vp, c_vp, c_vp, c_vp],
    "hrn_se3_exp": [c_vp, c_int, c_vp, c_vp],
    "hrn_resample_gather": [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp],
    "hrn_pose_errors": [c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "hrn_pose_from_covariance_host": [c_vp, c_vp, c_vp, c_vp, c_vp],
}

_lib = None


class HrnError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HrnError(f"{LIB_PATH} not built; run __graft_entry__.build() -- there is no CPU/eager fallback")
        L = ctypes.CDLL(LIB_PATH)
        for name, args in SIGNATURES.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = c_int
        L.hrn_version.restype = ctypes.c_char_p
        L.hrn_version.argtypes = []
        for name in ("hrn_level_pack_bytes", "hrn_level_bias_count", "hrn_level_ws_pack_bytes", "hrn_level_ws_bias_count"):
            getattr(L, name).restype = c_int
            getattr(L, name).argtypes = [c_int]
        _lib = L
    return _lib


def check(code: int, what: str):
    if code != 0:
        if code >= 1000:
            msg = {1001: "bad argument", 1002: "unsupported size"}.get(code, "error")
        else:
            msg = f"cudaError {code}"
        raise HrnError(f"{what} failed: {msg} ({code})")


_pending_dev = None      # device of the tensors whose pointers were taken for the next call()


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (or None).  Remembers the tensor's device for the launch: call()
    runs the kernel on THAT device and on its current stream, whatever the process-wide current device is."""
    global _pending_dev
    if t is None:
        return None
    if not t.is_cuda:
        raise HrnError("hregnet_b200 kernels need CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise HrnError("tensor must be contiguous")
    d = t.device.index
    _pending_dev = d if _pending_dev in (None, d) else -1          # -1: mixed devices, refused by call()
    return t.data_ptr()


class _CurrentStream:
    """Placeholder argument: call() replaces it by the current stream of the device the launch goes to."""


_CUR = _CurrentStream()


def stream():
    return _CUR


def call(name, *args):
    global _pending_dev
    dev, _pending_dev = _pending_dev, None
    if dev == -1:
        raise HrnError(f"{name}: the tensors of one kernel launch live on different devices")
    f = getattr(lib(), name)
    if dev is None or dev == torch.cuda.current_device():
        st = torch.cuda.current_stream().cuda_stream
        check(f(*[st if a is _CUR else a for a in args]), name)
        return
    with torch.cuda.device(dev):      # a net on cuda:1 while cuda:0 is current: launch where the data lives
        st = torch.cuda.current_stream(dev).cuda_stream
        check(f(*[st if a is _CUR else a for a in args]), name)
