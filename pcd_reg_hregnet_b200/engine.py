"""Host-side orchestration of the sm_100a kernels for the HRegNet forward path.

Everything here is plumbing: it allocates device buffers with torch, describes the virtual activation
matrices ("rows", csrc/rows.cuh) and launches kernels through the C ABI on torch's current stream.  The
arithmetic of the path lives in csrc/*.cu.  Internal layout is channels-last: a tensor of per-neighbour
features is a [rows, C] matrix with rows = (cloud, keypoint, neighbour) flattened.

Folded parameters (BatchNorm eval statistics merged into the 1x1 convolutions) come from fold.py.
"""
import ctypes
import os

import torch

from . import _lib
from ._lib import (ACT_NONE, ACT_RELU, ACT_SIGMOID, ACT_SOFTPLUS_EPS, SEG_BROADCAST, SEG_DIRECT, SEG_GATHER, Rows, call,
                   ptr, stream)

# "fp32": exact CUDA-core layers; "tc": tcgen05 kernels, bf16 hi/lo operands everywhere (bf16x3, fp32-class accuracy);
# "tcf" (opt-in): as "tc" for feature extraction, single-pass fp16 operands in the correspondence stages (CoarseReg /
# FineReg).  Measured on the B200 (tests/test_gpu_layers.py, profiles/r02b_bench_1gpu_tcf.json): the stage gates hold -- coarse
# correspondences 7e-4 of the 1e-3 gate, fine 3.5e-5, weights 3e-5 -- but the margin of the coarse stage is 1.4x and the step
# gains 2.5 % (5.64 -> 5.50 ms), so "tc" stays the default.  No 2- or 1-pass mode meets the gate for feature extraction
# (profiles/r02_precision_emulation.txt: descriptors up to 5.6e-3, sigmas 6.3e-3), which therefore always runs bf16x3.
_PRECISION = "tc"
PRECISIONS = ("fp32", "tc", "tcf")
_fast_depth = 0          # > 0 while a correspondence stage is being enqueued


def _tc():
    return _PRECISION in ("tc", "tcf")


def mma_prec():
    """MMA operand mode of the launch being enqueued: 1 = single-pass fp16, 3 = bf16 hi/lo."""
    return 1 if (_PRECISION == "tcf" and _fast_depth > 0) else 3


class _fast_stage:
    """with _fast_stage(): the shared-MLP launches inside may use the single-pass fp16 operands (mode "tcf" only)."""

    def __enter__(self):
        global _fast_depth
        _fast_depth += 1

    def __exit__(self, *exc):
        global _fast_depth
        _fast_depth -= 1

_FUSED_CHAINS = True     # tc mode: three-layer conv stacks (<= 256 wide) + their group reductions in one kernel
_SIDE_STREAM = True      # overlap independent branches of CoarseReg on a second CUDA stream
_side_streams = {}


def _side_stream(device):
    key = (device.type, device.index)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


_FUSED_HEADS = True      # tc mode: per-keypoint heads (mlp1/mlp2/mlp3, width <= 256) as one chain launch
_FUSED_LEVELS = True     # tc mode: run levels 1 and 2 (detector + descriptor) as one persistent tcgen05 kernel each
_WIDE_CHAIN = True       # tc modes: CoarseReg convs_1 + attention tail on a 2-CTA cluster per tile (csrc/chain_wide.cu)
_COSINE_TC = True        # tc modes: CoarseReg's cosine-similarity features as one tcgen05 kernel per similarity
_LEVEL_WS = (2, 3)         # tc mode: levels that run on the warp-specialised fused level kernel (csrc/level_ws.cu)


def on_side_stream(fn, like):
    """Runs fn() on a side stream forked from the current one (events: also valid under CUDA-graph capture) and returns a
    function that joins the streams and hands back fn's result.  `like`: a tensor on the device the work goes to.
    Falls back to running fn() in place when side streams are switched off."""
    if not (_SIDE_STREAM and like.is_cuda):
        res = fn()
        return lambda: res
    main = torch.cuda.current_stream(like.device)
    side = _api_stream(like.device)
    side.wait_stream(main)
    with torch.cuda.stream(side):
        res = fn()

    def join():
        torch.cuda.current_stream(like.device).wait_stream(side)
        for t in (res.values() if isinstance(res, dict) else (res if isinstance(res, (list, tuple)) else [res])):
            if torch.is_tensor(t):
                t.record_stream(torch.cuda.current_stream(like.device))
        return res
    return join


_api_streams = {}


def _api_stream(device):
    """A second side stream (the first one carries CoarseReg's descriptor branch and the Morton sorts)."""
    key = torch.device(device).index
    if key not in _api_streams:
        _api_streams[key] = torch.cuda.Stream(device=device)
    return _api_streams[key]


def set_precision(mode: str):
    """'fp32' = exact CUDA-core FFMA layers; 'tc' = tcgen05 tensor-core layers (bf16x3 split, fp32 accumulate); 'tcf' = 'tc'
    with single-pass fp16 operands in the correspondence stages."""
    global _PRECISION
    if mode not in PRECISIONS:
        raise ValueError(mode)
    _PRECISION = mode


def get_precision():
    return _PRECISION


def _addr(t):
    return None if t is None else t.data_ptr()


class RowsView:
    """Builder for hrn_rows_t; keeps the referenced tensors alive until the launch has been enqueued."""

    def __init__(self, rows, group=1, gather_idx=None, rows_per_batch=0, src_rows_per_batch=0):
        self.rows = int(rows)
        self.c = Rows()
        self.c.n_seg = 0
        self.c.group = int(group)
        self.c.gather_idx = _addr(gather_idx)
        self.c.rows_per_batch = int(rows_per_batch)
        self.c.src_rows_per_batch = int(src_rows_per_batch)
        self.keep = [gather_idx]
        self.gather_idx = gather_idx
        self.segs = []          # python-side mirror of the struct (debugging / CPU emulation in tests)
        self.K = 0

    def add(self, mat, mode=SEG_DIRECT, channels=None, col0=0, row_scale=None):
        """mat: 2-D view [src_rows, ld] (last dim contiguous)."""
        assert mat.dim() == 2 and mat.stride(1) == 1 and mat.dtype == torch.float32
        s = self.c.seg[self.c.n_seg]
        s.ptr = mat.data_ptr()
        s.row_scale = _addr(row_scale)
        s.channels = int(channels if channels is not None else mat.shape[1] - col0)
        s.ld = int(mat.stride(0))
        s.col0 = int(col0)
        s.mode = int(mode)
        self.c.n_seg += 1
        self.K += s.channels
        self.keep += [mat, row_scale]
        self.segs.append((mat, int(mode), s.channels, int(col0), row_scale))
        return self


def layer(view: RowsView, W, b, act, out=None):
    """out [rows, Cout] = act(X W^T + b)."""
    Cout, K = W.shape
    assert K == view.K, (K, view.K)
    if out is None:
        out = torch.empty(view.rows, Cout, dtype=torch.float32, device=W.device)
    if _tc():
        from . import engine_tc
        prec = mma_prec()
        if prec == 1 and not engine_tc.fast_layer_ok(view):
            prec = 3
        return engine_tc.layer_tc(view, W, b, act, out, prec)
    _launch_layer_fp32(view, W, b, act, out)
    return out


def _launch_layer_fp32(view, W, b, act, out):
    call("hrn_layer_fp32", ctypes.byref(view.c), ptr(W), ptr(b), act, ptr(out), out.stride(0), view.rows, W.shape[0],
         stream())


def stack(view: RowsView, layers, last_act=None):
    """Chain of folded layers [(W, b, act), ...] starting from a virtual rows view."""
    if _tc() and _FUSED_CHAINS and _FUSED_HEADS and len(layers) == 3 and len(view.segs) == 1:
        from . import engine_tc                      # per-keypoint heads mlp1 -> mlp2 -> mlp3 in one launch
        if engine_tc.chain_supported(view, layers, last_relu_only=False):
            return engine_tc.chain(view, layers, engine_tc.EPI_STORE, last_act=last_act)[0]
    x = None
    for li, (W, b, act) in enumerate(layers):
        if li == len(layers) - 1 and last_act is not None:
            act = last_act
        v = view if li == 0 else RowsView(x.shape[0]).add(x)
        x = layer(v, W, b, act)
    return x


def stack_group_max(view: RowsView, layers, k):
    """group_max(stack(view, layers), k) with the maximum taken in the last layer's epilogue when the tensor-core
    kernel supports it (the per-row output of the last layer is then never written)."""
    if _tc() and len(layers) >= 1:
        from . import engine_tc
        W, b, act = layers[-1]
        x = None
        for li, (Wl, bl, al) in enumerate(layers[:-1]):
            x = layer(view if li == 0 else RowsView(x.shape[0]).add(x), Wl, bl, al)
        v = view if x is None else RowsView(x.shape[0]).add(x)
        if engine_tc.layer_tc_groupmax_ok(v, W, act, k):
            return engine_tc.layer_tc_groupmax(v, W, b, act, k)
        return group_max(layer(v, W, b, act), k)
    return group_max(stack(view, layers), k)


def _chain_ok(view, layers, k):
    if not _tc() or not _FUSED_CHAINS or k not in (8, 16, 32):
        return False
    from . import engine_tc
    return engine_tc.chain_supported(view, layers)


def group_attention(E, k):
    rows, C = E.shape
    a = torch.empty(rows, dtype=torch.float32, device=E.device)
    call("hrn_group_attention", ptr(E), E.stride(0), C, rows // k, k, ptr(a), stream())
    return a


def group_weighted_sum(a, V, k, idx=None, groups_per_batch=0, N=0):
    """out[g,:] = sum_j a[g*k+j] V[row,:]; V is [rows, C] (idx None) or [B*N, C] gathered by idx."""
    groups = a.shape[0] // k
    C = V.shape[1]
    out = torch.empty(groups, C, dtype=torch.float32, device=a.device)
    call("hrn_group_weighted_sum", ptr(a), ptr(V), V.stride(0), C, groups, k, ptr(idx), groups_per_batch, N, ptr(out),
         out.stride(0), stream())
    return out


def group_max(X, k):
    rows, C = X.shape
    out = torch.empty(rows // k, C, dtype=torch.float32, device=X.device)
    call("hrn_group_max", ptr(X), X.stride(0), C, rows // k, k, ptr(out), out.stride(0), stream())
    return out


def group_geometry(q, p, idx, wq=None, wp=None, ld=None, want_nn=False):
    """q [B,M,3], p [B,N,3], idx [B,M,k] int32 -> misc [B*M*k, ld] (+ nn [B*M*k,3])."""
    B, M, k = idx.shape
    N = p.shape[1]
    ncol = 12 if wq is not None else 4
    ld = ld or ncol
    out = torch.empty(B * M * k, ld, dtype=torch.float32, device=q.device)
    nn = torch.empty(B * M * k, 3, dtype=torch.float32, device=q.device) if want_nn else None
    call("hrn_group_geometry", ptr(q), ptr(p), ptr(idx), ptr(wq), ptr(wp), B, M, k, N, ptr(out), ld, ptr(nn), stream())
    return out, nn


_SORTED_KNN = True


def knn_scratch(B, N, device):
    """Scratch of the spatially culled search: Morton-ordered points (float4) and one box per 32 points."""
    N2 = 1024
    while N2 < N:
        N2 *= 2
    return (torch.empty(B * N2, 4, dtype=torch.float32, device=device),
            torch.empty(B * (N2 // 32) * 6, dtype=torch.float32, device=device))


class KnnPresort:
    """Morton sort + boxes of the culled search on the side stream: it depends on the cloud only, so it runs on the SMs
    the sampling kernel leaves idle (one cluster of 2 SMs per cloud: 128 of 148 at 64 clouds) while the queries are still
    being chosen.  Usage: ps = KnnPresort.fork(xyz) BEFORE the sampling is launched (records the fork point),
    ps.launch() right AFTER it (so the sampling clusters are placed first), knn_idx(..., presorted=ps) joins."""

    def __init__(self, p2, pts, boxes, side):
        self.p2, self.pts, self.boxes, self.side = p2, pts, boxes, side

    @staticmethod
    def fork(p2):
        B, N, D = p2.shape
        if not (_SORTED_KNN and _SIDE_STREAM and D == 3 and 8192 <= N <= 32768):
            return None
        pts, boxes = knn_scratch(B, N, p2.device)
        side = _side_stream(p2.device)
        side.wait_stream(torch.cuda.current_stream(p2.device))
        return KnnPresort(p2, pts, boxes, side)

    def launch(self):
        B, N, _ = self.p2.shape
        with torch.cuda.stream(self.side):
            call("hrn_knn3_sort", ptr(self.p2), B, N, ptr(self.pts), ptr(self.boxes), stream())
        self.pts.record_stream(self.side)
        self.boxes.record_stream(self.side)


_SORTED_MIN = int(os.environ.get("HRN_SORTED_MIN", "256"))       # smallest cloud searched with the Morton-culled kernel (in-box A/B: 1024 -> 256 = -0.04 ms per step; brute force at 1024: +0.15 ms)


def knn_idx(p1, p2, K, q_idx=None, presorted=None):
    """int32 neighbour indices [B,M,K] (+ gathered queries when q_idx is given)."""
    B, N, D = p2.shape
    M = q_idx.shape[1] if q_idx is not None else p1.shape[1]
    idx = torch.empty(B, M, K, dtype=torch.int32, device=p2.device)
    q_out = torch.empty(B, M, 3, dtype=torch.float32, device=p2.device) if q_idx is not None else None
    if presorted is not None:
        pts, boxes = presorted.pts, presorted.boxes
        torch.cuda.current_stream(p2.device).wait_stream(presorted.side)
        call("hrn_knn3_search", ptr(p1) if q_idx is None else None, ptr(q_idx), ptr(p2), B, M, N, K, ptr(pts), ptr(boxes),
             None, None, ptr(idx), None, ptr(q_out), stream())
        return idx, q_out
    if D == 3 and _SORTED_MIN <= N <= 32768 and _SORTED_KNN:
        pts, boxes = knn_scratch(B, N, p2.device)
        call("hrn_knn3_sorted", ptr(p1) if q_idx is None else None, ptr(q_idx), ptr(p2), B, M, N, K, ptr(pts), ptr(boxes),
             None, None, ptr(idx), None, ptr(q_out), stream())
        return idx, q_out
    call("hrn_knn", ptr(p1) if q_idx is None else None, ptr(q_idx), ptr(p2), B, M, N, D, K, None, None, ptr(idx), None,
         ptr(q_out), stream())
    return idx, q_out


def fps(xyz, M, weights=None):
    B, N, _ = xyz.shape
    idx = torch.empty(B, M, dtype=torch.int32, device=xyz.device)
    temp = None if N <= 16384 else torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
    call("hrn_fps", ptr(xyz), ptr(weights), ptr(temp), ptr(idx), B, N, M, stream())
    return idx


def stack_clouds(src, dst):
    """[src; dst] as one batch of 2B clouds.  When the two already lie back to back in one allocation (the Registrar
    keeps them that way) this is a view; otherwise a copy (torch.cat)."""
    if (src.shape == dst.shape and src.dtype == dst.dtype and src.is_contiguous() and dst.is_contiguous()
            and src.untyped_storage().data_ptr() == dst.untyped_storage().data_ptr()
            and dst.storage_offset() == src.storage_offset() + src.numel()):
        return src.as_strided((2 * src.shape[0],) + tuple(src.shape[1:]), src.stride(), src.storage_offset())
    return torch.cat([src, dst], dim=0)


def transpose(x):
    """[B,R,C] -> [B,C,R] contiguous."""
    B, R, C = x.shape
    out = torch.empty(B, C, R, dtype=torch.float32, device=x.device)
    call("hrn_transpose", ptr(x), ptr(out), B, R, C, stream())
    return out


def sigma_to_weights(sigma):
    B, M = sigma.shape
    w = torch.empty_like(sigma)
    call("hrn_sigma_to_weights", ptr(sigma), ptr(w), B, M, stream())
    return w


def transform_points(x, R, t):
    B, N, _ = x.shape
    out = torch.empty_like(x)
    call("hrn_transform_points", ptr(x), ptr(R), ptr(t), ptr(out), B, N, stream())
    return out


def weighted_kabsch(src, cor, w, prev=None, packed=False):
    """-> (R [B,3,3], t [B,3]) and, with prev=(R_prev, t_prev), also the composed (R_c, t_c).
    packed=True: the final pose is also written as rows [R | t] of a [B,12] tensor attached to the returned final
    rotation as `.hrn_pose12` -- what dist.gather_poses sends, so no packing kernel runs in front of the collective."""
    B, N, _ = src.shape
    R = torch.empty(B, 3, 3, dtype=torch.float32, device=src.device)
    t = torch.empty(B, 3, dtype=torch.float32, device=src.device)
    p12 = torch.empty(B, 12, dtype=torch.float32, device=src.device) if packed else None
    if prev is None:
        call("hrn_weighted_kabsch", ptr(src), ptr(cor), ptr(w), B, N, None, None, ptr(R), ptr(t), None, None, ptr(p12), stream())
        if packed:
            R.hrn_pose12 = p12
        return R, t
    Rc, tc = torch.empty_like(R), torch.empty_like(t)
    call("hrn_weighted_kabsch", ptr(src), ptr(cor), ptr(w), B, N, ptr(prev[0]), ptr(prev[1]), ptr(R), ptr(t), ptr(Rc),
         ptr(tc), ptr(p12), stream())
    if packed:
        Rc.hrn_pose12 = p12
    return R, t, Rc, tc


# ------------------------------------------------------------------------------------------------------------------
# fused stages
# ------------------------------------------------------------------------------------------------------------------

def random_sample_idx(n_points, n_sample, batch_sizes, device):
    """The `fps=False` branch of KeypointDetector (reference layers.py:144-147): `torch.randperm(N)[:nsample]` drawn from
    the HOST generator, ONE draw per detector call, shared by every cloud of that call.  batch_sizes = clouds per
    reference call (the model path stacks the source and the target call into one batch: two draws, in call order).
    Returns int32 [sum(batch_sizes), n_sample] on `device`."""
    if torch.cuda.is_current_stream_capturing():
        raise RuntimeError("use_fps=False draws the sample on the host in every forward (layers.py:146): it cannot be "
                           "captured in a CUDA graph -- use Registrar(..., use_cuda_graph=False) or call the net eagerly")
    rows = [torch.randperm(n_points)[:n_sample].to(torch.int32)[None].expand(b, n_sample) for b in batch_sizes]
    return torch.cat(rows, 0).contiguous().to(device)


def detector_descriptor_level(xyz, feat_cl, weights, det, desc, M, k, want_maps=False, sample_idx=None):
    """One hierarchy level: KeypointDetector + DescExtractor (reference layers.py:134-165 + 200-209) without
    materialising grouped_features / attentive_feature_map.

    xyz [B,N,3]; feat_cl [B,N,C] channels-last or None; weights [B,N] or None (-> weighted FPS).
    det / desc: folded parameter dicts (fold.py).  Returns dict(xyz [B,M,3], sigmas [B,M], af [B,M,C_o] ,
    desc [B,M,desc_dim], and with want_maps also the rows-layout G, E, a, idx for the layer-level API).
    sample_idx int32 [B,M]: use these samples instead of (weighted) FPS (the reference's fps=False branch)."""
    B, N, _ = xyz.shape
    presorted = KnnPresort.fork(xyz)
    fidx = fps(xyz, M, weights) if sample_idx is None else sample_idx
    if presorted is not None:
        presorted.launch()
    idx, q = knn_idx(None, xyz, k, q_idx=fidx, presorted=presorted)
    if _tc() and _FUSED_LEVELS and not want_maps and (B * M * k) % 128 == 0:
        from . import engine_tc
        cin = 0 if feat_cl is None else feat_cl.shape[2]
        lw = engine_tc.which_level_ws(k, cin, det, desc) if feat_cl is not None else None
        if lw is not None and lw in _LEVEL_WS:
            keypoints, af, d = engine_tc.level_ws(lw, q, xyz, feat_cl, idx, det, desc)
            sig = stack(RowsView(B * M).add(af), det["mlp"], last_act=ACT_SOFTPLUS_EPS)
            return dict(xyz=keypoints.view(B, M, 3), sigmas=sig.view(B, M), af=af.view(B, M, -1), desc=d.view(B, M, -1))
        lv = engine_tc.which_level(k, cin, det, desc)
        if lv is not None:
            keypoints, af, d = engine_tc.level_fused(lv, q, xyz, feat_cl, idx, det, desc)
            sig = stack(RowsView(B * M).add(af), det["mlp"], last_act=ACT_SOFTPLUS_EPS)
            return dict(xyz=keypoints.view(B, M, 3), sigmas=sig.view(B, M), af=af.view(B, M, -1), desc=d.view(B, M, -1))
    geom, nn = group_geometry(q, xyz, idx, want_nn=True)
    rows = B * M * k

    def grouped():
        v = RowsView(rows, group=k, gather_idx=idx, rows_per_batch=M * k, src_rows_per_batch=N).add(geom)
        if feat_cl is not None:
            v.add(feat_cl.view(B * N, -1), SEG_GATHER)
        return v

    if _chain_ok(grouped(), det["convs"], k) and not want_maps:
        from . import engine_tc
        # detector stack + attention in one kernel: rows E*a (the attentive feature map), attentive feature, weights
        Ea, af, a = engine_tc.chain3(grouped(), det["convs"], engine_tc.EPI_ATTN, k)
        keypoints = group_weighted_sum(a, nn, k)
        sig = stack(RowsView(B * M).add(af), det["mlp"], last_act=ACT_SOFTPLUS_EPS)
        X1, X1max, _ = engine_tc.chain3(grouped(), desc["convs"], engine_tc.EPI_GROUPMAX, k)
        v = RowsView(rows, group=k).add(X1max, SEG_BROADCAST).add(X1).add(Ea)
        # (mlp1+mlp2 through the chain kernel measured 1.2 ms vs 0.69 ms for two per-layer launches + group_max:
        #  K0 = 768 needs three operand passes per tile, each exposing the gather latency)
        d = stack_group_max(v, desc["mlp"], k)
        return dict(xyz=keypoints.view(B, M, 3), sigmas=sig.view(B, M), af=af.view(B, M, -1), desc=d.view(B, M, -1))
    E = stack(grouped(), det["convs"])
    a = group_attention(E, k)
    keypoints = group_weighted_sum(a, nn, k)
    af = group_weighted_sum(a, E, k)
    sig = stack(RowsView(B * M).add(af), det["mlp"], last_act=ACT_SOFTPLUS_EPS)

    X1 = stack(grouped(), desc["convs"])
    X1max = group_max(X1, k)
    v = RowsView(rows, group=k).add(X1max, SEG_BROADCAST).add(X1).add(E, row_scale=a)
    H = stack(v, desc["mlp"])
    d = group_max(H, k)
    out = dict(xyz=keypoints.view(B, M, 3), sigmas=sig.view(B, M), af=af.view(B, M, -1), desc=d.view(B, M, -1))
    if want_maps:
        out.update(geom=geom, E=E, a=a, idx=idx)
    return out


def _tail(F, a_k, idx, dxyz, B, N1, N2, head):
    """softmax attention over the k candidates, correspondence + confidence (layers.py:385-394, 447-452)."""
    rows, C = F.shape
    if C % 4 == 0 and F.stride(0) % 4 == 0 and F.data_ptr() % 16 == 0 and a_k * C * 4 <= 96 * 1024:
        # attention, attentive feature and correspondence in one pass over F
        af = torch.empty(rows // a_k, C, dtype=torch.float32, device=F.device)
        cor = torch.empty(rows // a_k, 3, dtype=torch.float32, device=F.device)
        call("hrn_group_attend", ptr(F), F.stride(0), C, rows // a_k, a_k, None, ptr(af), af.stride(0),
             ptr(dxyz), ptr(idx), N1, N2, ptr(cor), stream())
    else:
        a = group_attention(F, a_k)
        cor = group_weighted_sum(a, dxyz.view(B * N2, 3), a_k, idx=idx, groups_per_batch=N1, N=N2)
        af = group_weighted_sum(a, F, a_k)
    w = stack(RowsView(B * N1).add(af), head, last_act=ACT_SIGMOID)
    return cor.view(B, N1, 3), w.view(B, N1), af


def fine_reg(sxyz, sfeat_cl, dxyz, dfeat_cl, ssig, dsig, P, k=8, want_af=False):
    """FineReg.forward (reference layers.py:433-454) on channels-last features [B,N,C]."""
    with _fast_stage():
        return _fine_reg(sxyz, sfeat_cl, dxyz, dfeat_cl, ssig, dsig, P, k, want_af)


_PREFETCH_Z = os.environ.get("HRN_PREFETCH_Z", "1") == "1"


def _prefetch_stage(sfeat_cl, dfeat_cl, W1, misc_channels):
    """Starts the per-point parts of a correspondence stage's first conv layer (column layout of its input:
    [misc | source features | target features]) beside the stage's neighbour search."""
    if not (_PREFETCH_Z and _tc() and _FUSED_CHAINS and _SIDE_STREAM and sfeat_cl.is_cuda):
        return
    from . import engine_tc
    C = sfeat_cl.shape[-1]
    if W1.shape[1] != misc_channels + 2 * C:
        return
    engine_tc.prefetch_point_layers(sfeat_cl, W1, [(sfeat_cl.reshape(-1, C), C, 0, misc_channels),
                                                   (dfeat_cl.reshape(-1, C), C, 0, misc_channels + C)])


def _fine_reg(sxyz, sfeat_cl, dxyz, dfeat_cl, ssig, dsig, P, k, want_af):
    B, N1, _ = sxyz.shape
    N2 = dxyz.shape[1]
    _prefetch_stage(sfeat_cl, dfeat_cl, P["convs_1"][0][0], 12)
    idx, _ = knn_idx(sxyz, dxyz, k)
    misc, _ = group_geometry(sxyz, dxyz, idx, ssig, dsig)
    v = RowsView(B * N1 * k, group=k, gather_idx=idx, rows_per_batch=N1 * k, src_rows_per_batch=N2)
    v.add(misc).add(sfeat_cl.view(B * N1, -1), SEG_BROADCAST).add(dfeat_cl.view(B * N2, -1), SEG_GATHER)
    if _chain_ok(v, P["convs_1"], k):
        from . import engine_tc
        _, af, a = engine_tc.chain3(v, P["convs_1"], engine_tc.EPI_ATTN, k, want_rows=False)
        # the correspondence (a small gather) runs beside the confidence head: both only need the attention's results
        cor = on_side_stream(lambda: group_weighted_sum(a, dxyz.view(B * N2, 3), k, idx=idx, groups_per_batch=N1, N=N2), a)
        w = stack(RowsView(B * N1).add(af), P["mlp"], last_act=ACT_SIGMOID).view(B, N1)
        cor = cor().view(B, N1, 3)
    else:
        F = stack(v, P["convs_1"])
        cor, w, af = _tail(F, k, idx, dxyz, B, N1, N2, P["mlp"])
    return (cor, w, af) if want_af else (cor, w)


def _neighbour_aware(xyz, desc_cl, P, k):
    """Neighbourhood-attentive descriptors (reference layers.py:316-337)."""
    B, N, C = desc_cl.shape
    nidx, _ = knn_idx(xyz, xyz, k)
    geom, _ = group_geometry(xyz, xyz, nidx)
    v = RowsView(B * N * k, group=k, gather_idx=nidx, rows_per_batch=N * k, src_rows_per_batch=N)
    v.add(desc_cl.view(B * N, C), SEG_GATHER).add(geom)
    if _chain_ok(v, P["convs_2"], k):
        from . import engine_tc
        _, _, a = engine_tc.chain3(v, P["convs_2"], engine_tc.EPI_ATTN, k, want_rows=False, want_groups=False)
    else:
        a = group_attention(stack(v, P["convs_2"]), k)
    return group_weighted_sum(a, desc_cl.view(B * N, C), k, idx=nidx, groups_per_batch=N, N=N).view(B, N, C)


def _cosine_features(S, D, idx, misc, col_sd, col_ds):
    B, N1, C = S.shape
    N2 = D.shape[1]
    k = idx.shape[2]
    dev = S.device
    if (_tc() and _COSINE_TC and N1 in (128, 256) and N2 <= 256 and N2 % 16 == 0 and C % 32 == 0 and k == 8
            and S.data_ptr() % 16 == 0 and D.data_ptr() % 16 == 0):
        # one tcgen05 kernel per similarity: contraction, both families of maxima and the picks (csrc/coarse_tc.cu)
        call("hrn_cosine_features_tc", ptr(S), ptr(D), ptr(idx), B, N1, N2, C, k, ptr(misc), misc.stride(0), col_sd, col_ds,
             stream())
        return
    nS = torch.empty(B, N1, device=dev)
    nD = torch.empty(B, N2, device=dev)
    cosm = torch.empty(B, N2, N1, device=dev)
    rowmax = torch.empty(B, N2, device=dev)
    colmax = torch.empty(B, N1, device=dev)
    call("hrn_cosine_matrix", ptr(S), ptr(D), B, N1, N2, C, ptr(nS), ptr(nD), ptr(cosm), ptr(rowmax), ptr(colmax), stream())
    call("hrn_cosine_pick", ptr(cosm), ptr(rowmax), ptr(colmax), ptr(idx), B, N1, N2, k, ptr(misc), misc.stride(0),
         col_sd, col_ds, stream())


def coarse_reg(sxyz, sdesc_cl, dxyz, ddesc_cl, ssig, dsig, P, k=8, both=None, want_dists=False):
    """CoarseReg.forward with use_sim=use_neighbor=True (reference layers.py:273-396).

    both = (xyz [2B,N,3], desc_cl [2B,N,C]) with the source clouds in the first half and the target clouds in the
    second (the model path has them contiguous): the two neighbour-aware branches (layers.py:316-337) then run as ONE
    batch of 2B clouds instead of two."""
    with _fast_stage():
        return _coarse_reg(sxyz, sdesc_cl, dxyz, ddesc_cl, ssig, dsig, P, k, both, want_dists)


def _coarse_reg(sxyz, sdesc_cl, dxyz, ddesc_cl, ssig, dsig, P, k, both, want_dists):
    B, N1, C = sdesc_cl.shape
    N2 = dxyz.shape[1]
    _prefetch_stage(sdesc_cl, ddesc_cl, P["convs_1"][0][0], 16)

    def descriptor_branch():
        i, _ = knn_idx(sdesc_cl, ddesc_cl, k)                      # 256-d descriptor space (layers.py:278)
        m, _ = group_geometry(sxyz, dxyz, i, ssig, dsig, ld=16)     # cols 0..11; 12..15 = similarity features
        _cosine_features(sdesc_cl, ddesc_cl, i, m, 12, 13)
        return i, m

    def neighbour_branch():
        if both is not None and N1 == N2:
            nb = _neighbour_aware(both[0], both[1], P, k)
            return nb[:B], nb[B:]
        return _neighbour_aware(sxyz, sdesc_cl, P, k), _neighbour_aware(dxyz, ddesc_cl, P, k)

    if _SIDE_STREAM and sxyz.is_cuda:
        # the descriptor-space search + plain similarity features and the neighbour-aware branch (layers.py:316-337) are
        # independent until the second similarity pick: run the former on a side stream (fork/join with events, also
        # valid under CUDA-graph capture)
        main = torch.cuda.current_stream(sxyz.device)
        side = _side_stream(sxyz.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            idx, misc = descriptor_branch()
        s_nbr, d_nbr = neighbour_branch()
        main.wait_stream(side)
        idx.record_stream(main)
        misc.record_stream(main)
    else:
        idx, misc = descriptor_branch()
        s_nbr, d_nbr = neighbour_branch()
    _cosine_features(s_nbr, d_nbr, idx, misc, 14, 15)
    v = RowsView(B * N1 * k, group=k, gather_idx=idx, rows_per_batch=N1 * k, src_rows_per_batch=N2)
    v.add(misc).add(sdesc_cl.view(B * N1, C), SEG_BROADCAST).add(ddesc_cl.view(B * N2, C), SEG_GATHER)
    wide = False
    if _tc() and _FUSED_CHAINS and _WIDE_CHAIN:
        from . import engine_tc
        wide = engine_tc.chain_wide_supported(v, P["convs_1"], k)
    if wide:
        # conv stack + attention on a 2-CTA cluster per tile: the 512-wide activations never reach HBM (csrc/chain_wide.cu)
        af, a = engine_tc.chain_wide(v, P["convs_1"], k)
        cor = on_side_stream(lambda: group_weighted_sum(a, dxyz.view(B * N2, 3), k, idx=idx, groups_per_batch=N1, N=N2), a)
        hv = RowsView(B * N1).add(af)
        if engine_tc.chain_wide_head_supported(hv, P["mlp"]):
            w = engine_tc.chain_wide_head(hv, P["mlp"], ACT_SIGMOID).view(B, N1)      # 512 -> 512 -> 512 -> 1 in one launch
        else:
            w = stack(hv, P["mlp"], last_act=ACT_SIGMOID).view(B, N1)
        cor = cor().view(B, N1, 3)
    else:
        F = stack(v, P["convs_1"])
        cor, w, _ = _tail(F, k, idx, dxyz, B, N1, N2, P["mlp"])
    if want_dists:
        # model_v4's CoarseReg also returns two by-products of the feature assembly (model_v4/layers.py:252,282):
        # coord_dist = |candidate - source keypoint| (geometry column 3) and feats_dist = 1 - the normalised
        # dst->src cosine similarity picked at the candidates (similarity column 13)
        m3 = misc.view(B, N1, k, -1)
        return cor, w, m3[..., 3].contiguous(), 1.0 - m3[..., 13]
    return cor, w
