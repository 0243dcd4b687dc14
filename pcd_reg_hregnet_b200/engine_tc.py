"""Host side of the tensor-core shared-MLP layer (csrc/mlp_tc.cu): weight packing + launch.

pack_weights(): fp32 [Cout, K] (BatchNorm already folded) -> bf16 hi/lo split, K padded per rows-segment to a
multiple of 8 (and in total to a multiple of 32), Cout padded to a multiple of 16, tiled per pipeline stage as
[n_stage][hi|lo][K/8 within stage = 4][NP][8] so that one stage is one contiguous cp.async.bulk.
Packed weights are cached per (weight tensor, segment signature)."""
import ctypes
import weakref

import torch

from . import engine

KC = 32
_cache = {}


def _remember(cache, key, owner, value):
    """Packed copies live exactly as long as the folded weight tensor they were made from: when a module re-folds
    (load_state_dict, an eval between training epochs) the old folded tensors die and their packs go with them --
    the caches do not grow over time.  Keys hold data_ptr / version only; `owner` is referenced weakly."""
    cache[key] = value
    weakref.finalize(owner, cache.pop, key, None)
    return value


def _split_planes(Wp, prec):
    """fp32 -> the operand planes of a precision mode: 3 = bf16 hi + bf16 lo (three MMAs per MAC), 1 = one fp16 plane."""
    if prec == 1:
        return [Wp.to(torch.float16).view(torch.bfloat16)]          # same 16-bit container type downstream
    hi = Wp.to(torch.bfloat16)
    return [hi, (Wp - hi.float()).to(torch.bfloat16)]


def pack_weights(W, seg_channels, prec=3):
    key = (W.data_ptr(), W._version, tuple(seg_channels), W.device, prec)
    hit = _cache.get(key)
    if hit is not None:
        return hit
    Cout, K = W.shape
    assert sum(seg_channels) == K
    NP = max(16, (Cout + 15) // 16 * 16)
    if NP > 256:
        assert NP <= 512
        NP = 512
    kpad = sum((c + 7) // 8 * 8 for c in seg_channels)
    n_stage = (kpad + KC - 1) // KC
    Wp = torch.zeros(NP, n_stage * KC, dtype=torch.float32, device=W.device)
    src, dst = 0, 0
    for c in seg_channels:
        Wp[:Cout, dst:dst + c] = W[:, src:src + c]
        src += c
        dst += (c + 7) // 8 * 8
    planes = _split_planes(Wp, prec)
    t = torch.stack(planes, 0).view(len(planes), NP, n_stage, KC // 8, 8).permute(2, 0, 3, 1, 4).contiguous()
    return _remember(_cache, key, W, (t, NP, n_stage, None))


def fast_layer_ok(view):
    """The single-pass fp16 variant of the per-layer kernel exists for 16-byte aligned segments only."""
    return all(ch % 4 == 0 and col0 % 4 == 0 and mat.stride(0) % 4 == 0 and mat.data_ptr() % 16 == 0
               for mat, mode, ch, col0, scale in view.segs)


def layer_tc(view, W, b, act, out, prec=3):
    seg_channels = [s[2] for s in view.segs]
    Wp, NP, n_stage, _ = pack_weights(W, seg_channels, prec)
    if b is None:
        b = torch.zeros(W.shape[0], dtype=torch.float32, device=W.device)
    engine.call("hrn_layer_tc", ctypes.byref(view.c), engine.ptr(Wp), engine.ptr(b), act, engine.ptr(out), out.stride(0),
                view.rows, W.shape[0], NP, n_stage, prec, engine.stream())
    return out


def layer_tc_groupmax_ok(view, W, act, k):
    """Preconditions of hrn_layer_tc_groupmax (include/hregnet_b200.h)."""
    if k not in (8, 16, 32) or act not in (engine.ACT_NONE, engine.ACT_RELU) or view.rows % k:
        return False
    cout = W.shape[0]
    if cout % 32 or cout > 512 or (cout > 256 and cout != 512):
        return False
    for mat, mode, ch, col0, scale in view.segs:
        if ch % 4 or col0 % 4 or mat.stride(0) % 4 or mat.data_ptr() % 16:
            return False
    return True


def layer_tc_groupmax(view, W, b, act, k, prec=3):
    """G [rows / k, Cout] = max over each k consecutive rows of act(W x + b): layer + reference max(dim=3) in one launch."""
    seg_channels = [s[2] for s in view.segs]
    Wp, NP, n_stage, _ = pack_weights(W, seg_channels, prec)
    if b is None:
        b = torch.zeros(W.shape[0], dtype=torch.float32, device=W.device)
    out = torch.empty(view.rows // k, W.shape[0], dtype=torch.float32, device=W.device)
    engine.call("hrn_layer_tc_groupmax", ctypes.byref(view.c), engine.ptr(Wp), engine.ptr(b), act, engine.ptr(out),
                out.stride(0), view.rows, W.shape[0], NP, n_stage, k, prec, engine.stream())
    return out


# ------------------------------------------------------------------------------------------------------------------
# fused level kernels (csrc/level_fused.cu)
# ------------------------------------------------------------------------------------------------------------------
_level_cache = {}
_LEVEL_DIMS = {1: dict(k=64, cin=0, c1=32, c2=32, co=64, cmid=32, cd=64)}      # levels 2 / 3: level_ws below


def _umma_tiles(W, K_pad):
    """fp32 [N, K] -> bytes of a UMMA B operand block: hi plane [K_pad/8][N][8] bf16, then the lo plane."""
    N, K = W.shape
    Wp = torch.zeros(N, K_pad, dtype=torch.float32, device=W.device)
    Wp[:, :K] = W
    hi = Wp.to(torch.bfloat16)
    lo = (Wp - hi.float()).to(torch.bfloat16)
    planes = [t.view(N, K_pad // 8, 8).permute(1, 0, 2).contiguous().view(-1) for t in (hi, lo)]
    return torch.cat(planes).view(torch.uint8)


def level_supported(level_dims, det, desc):
    d = level_dims
    shapes = [tuple(W.shape) for W, _, _ in det["convs"] + desc["convs"] + desc["mlp"]]
    first = (d["c1"], d["cin"] + 4)
    return shapes == [first, (d["c2"], d["c1"]), (d["co"], d["c2"])] * 2 + [(d["cmid"], 3 * d["co"]), (d["cd"], d["cmid"])]


def which_level(k, cin, det, desc):
    for lv, d in _LEVEL_DIMS.items():
        if d["k"] == k and d["cin"] == cin and level_supported(d, det, desc):
            return lv
    return None


def pack_level(level, det, desc):
    """Folded parameter dicts of detector_1 / desc_extractor_1 -> (Wpack uint8, biases fp32) in the LevelCfg layout of
    csrc/level_fused.cu: MMA blocks in execution order [d1;x1] d2 d3 x2 x3 mlp1[X1] mlp1[E*a] mlp2, then the fp32 block
    mlp1[max_k X1] for the per-keypoint mat-vec (two half-matrices [CO/2][CMID], input channels c with (c >> 2) & 1 == h
    in half h, 16 floats of padding behind each).  The grouped input channels are re-ordered from the reference's
    [rel(3), dist(1), feat(C)] (layers.py:21-26) to [feat(C), rel(3), dist(1), 1, 0-pad]: the gathered feature row lands
    on 8-channel chunk boundaries, and the constant-1 channel multiplies the biases of d1 / x1 (stored as a weight column)."""
    key = (level,) + tuple((W.data_ptr(), W._version) for W, _, _ in det["convs"] + desc["convs"] + desc["mlp"])
    hit = _level_cache.get(key)
    if hit is not None:
        return hit
    d = _LEVEL_DIMS[level]
    (d1, bd1, _), (d2, bd2, _), (d3, bd3, _) = det["convs"]
    (x1, bx1, _), (x2, bx2, _), (x3, bx3, _) = desc["convs"]
    (m1, bm1, _), (m2, bm2, _) = desc["mlp"]
    CO, cin, CMID = d["co"], d["cin"], d["cmid"]
    KG = (cin + 5 + 15) // 16 * 16
    perm = list(range(4, 4 + cin)) + [0, 1, 2, 3]
    first = torch.cat([torch.cat([d1[:, perm], bd1[:, None]], 1), torch.cat([x1[:, perm], bx1[:, None]], 1)], 0).contiguous()
    parts = [_umma_tiles(first, KG), _umma_tiles(d2, d["c1"]), _umma_tiles(d3, d["c2"]),
             _umma_tiles(x2, d["c1"]), _umma_tiles(x3, d["c2"]),
             _umma_tiles(m1[:, CO:2 * CO].contiguous(), CO), _umma_tiles(m1[:, 2 * CO:].contiguous(), CO),
             _umma_tiles(m2, CMID)]
    Wa = m1[:, :CO]                                                       # [CMID, CO]: the max_k(X1) block (layers.py:203-205)
    c = torch.arange(CO, device=Wa.device)
    pad = torch.zeros(16, dtype=torch.float32, device=Wa.device)
    for h in (0, 1):
        parts.append(Wa[:, c[((c >> 2) & 1) == h]].t().contiguous().view(-1).view(torch.uint8))
        parts.append(pad.view(torch.uint8))
    from ._lib import lib
    if lib().hrn_level_pack_bytes(level) > sum(p.numel() for p in parts):
        # library built with the biases of d2 d3 x2 x3 m2 as MMA pieces: per layer [2 chunks][N][8] bf16 with the bias as
        # hi (K column 0) + lo (K column 1), multiplied by a resident block of ones
        for b in (bd2, bd3, bx2, bx3, bm2):
            piece = torch.zeros(2, b.numel(), 8, dtype=torch.bfloat16, device=b.device)
            hi = b.to(torch.bfloat16)
            piece[0, :, 0] = hi
            piece[0, :, 1] = (b - hi.float()).to(torch.bfloat16)
            parts.append(piece.view(-1).view(torch.uint8))
    Wpack = torch.cat(parts).contiguous()
    biases = torch.cat([bd1, bd2, bd3, bx1, bx2, bx3, bm1, bm2]).contiguous()
    assert Wpack.numel() == lib().hrn_level_pack_bytes(level) and biases.numel() == lib().hrn_level_bias_count(level)
    return _remember(_level_cache, key, d1, (Wpack, biases, None))


def level_fused(level, q, xyz, feat_cl, idx, det, desc):
    """q [B,M,3], xyz [B,N,3], feat_cl [B,N,C] | None, idx [B,M,k] int32 -> keypoints [B*M,3], af [B*M,CO], desc [B*M,CD]."""
    B, M, k = idx.shape
    N = xyz.shape[1]
    d = _LEVEL_DIMS[level]
    Wpack, biases, _ = pack_level(level, det, desc)
    dev = xyz.device
    kp = torch.empty(B * M, 3, dtype=torch.float32, device=dev)
    af = torch.empty(B * M, d["co"], dtype=torch.float32, device=dev)
    ds = torch.empty(B * M, d["cd"], dtype=torch.float32, device=dev)
    engine.call("hrn_level_fused", level, engine.ptr(q), engine.ptr(xyz), engine.ptr(feat_cl), engine.ptr(idx),
                engine.ptr(Wpack), engine.ptr(biases), engine.ptr(kp), engine.ptr(af), engine.ptr(ds), B, M, N, k,
                engine.stream())
    return kp, af, ds


# ------------------------------------------------------------------------------------------------------------------
# warp-specialised fused level kernel (csrc/level_ws.cu): levels 2 and 3
# ------------------------------------------------------------------------------------------------------------------
_level_ws_cache = {}
_LEVEL_WS_DIMS = {2: dict(k=32, cin=64, c=64), 3: dict(k=16, cin=128, c=128)}


def which_level_ws(k, cin, det, desc):
    for lv, d in _LEVEL_WS_DIMS.items():
        dims = dict(k=d["k"], cin=d["cin"], c1=d["c"], c2=d["c"], co=2 * d["c"], cmid=d["c"], cd=2 * d["c"])
        if d["k"] == k and d["cin"] == cin and level_supported(dims, det, desc):
            return lv
    return None


def pack_level_ws(level, det, desc):
    """Folded parameter dicts of detector_l / desc_extractor_l -> (Wpack uint8, WaT fp32 [2C, C], biases fp32) in the
    LwCfg layout of csrc/level_ws.cu: K=16 weight pieces of the 8 MMA layers in issue order
    [d1;x1] d2 x2 d3 x3 mlp1[E*a] mlp1[X1] mlp2 (d2 x2 d3 x3 mlp2 each preceded by their bias piece); the max_k(X1) block of mlp1 (its first 2C input channels,
    layers.py:203-205) goes to the CUDA cores as the transposed fp32 matrix WaT.  Grouped input channels re-ordered as in
    pack_level: [feat(C), rel(3), dist(1), 1, 0-pad]."""
    key = (level,) + tuple((W.data_ptr(), W._version) for W, _, _ in det["convs"] + desc["convs"] + desc["mlp"])
    hit = _level_ws_cache.get(key)
    if hit is not None:
        return hit
    d = _LEVEL_WS_DIMS[level]
    C, cin = d["c"], d["cin"]
    CO = 2 * C
    (d1, bd1, _), (d2, bd2, _), (d3, bd3, _) = det["convs"]
    (x1, bx1, _), (x2, bx2, _), (x3, bx3, _) = desc["convs"]
    (m1, bm1, _), (m2, bm2, _) = desc["mlp"]
    KG = (cin + 5 + 15) // 16 * 16
    perm = list(range(4, 4 + cin)) + [0, 1, 2, 3]
    first = torch.zeros(2 * C, KG, dtype=torch.float32, device=d1.device)
    first[:, :cin + 4] = torch.cat([d1[:, perm], x1[:, perm]], 0)
    first[:, cin + 4] = torch.cat([bd1, bx1])            # weight column of the constant-1 channel = the biases of d1 / x1

    def bias_piece(b):
        # [2 chunks][N][8] bf16 with the bias as hi (K column 0) + lo (K column 1): multiplied by a resident block of ones,
        # the layer's first MMA (csrc/level_ws.cu)
        piece = torch.zeros(2, b.numel(), 8, dtype=torch.bfloat16, device=b.device)
        hi = b.to(torch.bfloat16)
        piece[0, :, 0] = hi
        piece[0, :, 1] = (b - hi.float()).to(torch.bfloat16)
        return piece.view(-1).view(torch.uint8)

    parts = [_pieces(first, KG), bias_piece(bd2), _pieces(d2, C), bias_piece(bx2), _pieces(x2, C),
             bias_piece(bd3), _pieces(d3, C), bias_piece(bx3), _pieces(x3, C),
             _pieces(m1[:, 2 * CO:].contiguous(), CO), _pieces(m1[:, CO:2 * CO].contiguous(), CO),
             bias_piece(bm2), _pieces(m2, C)]
    from ._lib import lib
    if lib().hrn_level_ws_pack_bytes(level) < sum(p.numel() for p in parts):
        parts = [p for i, p in enumerate(parts) if i not in (1, 3, 5, 7, 11)]     # a library built with additive biases (A/B runs)
    Wpack = torch.cat(parts).contiguous()
    WaT = m1[:, :CO].t().contiguous().view(CO // 4, 4, C).permute(0, 2, 1).contiguous()     # [CO/4][C][4]
    biases = torch.cat([bd1, bd2, bd3, bx1, bx2, bx3, bm1, bm2]).contiguous()
    from ._lib import lib
    assert Wpack.numel() == lib().hrn_level_ws_pack_bytes(level) and biases.numel() == lib().hrn_level_ws_bias_count(level)
    return _remember(_level_ws_cache, key, d1, (Wpack, WaT, biases))


def level_ws(level, q, xyz, feat_cl, idx, det, desc):
    """q [B,M,3], xyz [B,N,3], feat_cl [B,N,C], idx [B,M,k] int32 -> keypoints [B*M,3], af [B*M,2C], desc [B*M,2C]."""
    B, M, k = idx.shape
    N = xyz.shape[1]
    CO = 2 * _LEVEL_WS_DIMS[level]["c"]
    Wpack, WaT, biases = pack_level_ws(level, det, desc)
    dev = xyz.device
    kp = torch.empty(B * M, 3, dtype=torch.float32, device=dev)
    af = torch.empty(B * M, CO, dtype=torch.float32, device=dev)
    ds = torch.empty(B * M, CO, dtype=torch.float32, device=dev)
    engine.call("hrn_level_ws", level, engine.ptr(q), engine.ptr(xyz), engine.ptr(feat_cl), engine.ptr(idx),
                engine.ptr(Wpack), engine.ptr(WaT), engine.ptr(biases), engine.ptr(kp), engine.ptr(af), engine.ptr(ds),
                B, M, N, k, engine.stream())
    return kp, af, ds


# ------------------------------------------------------------------------------------------------------------------
# three-layer chain kernel (csrc/chain_tc.cu)
# ------------------------------------------------------------------------------------------------------------------
EPI_STORE, EPI_GROUPMAX, EPI_ATTN = 0, 1, 2
_chain_cache = {}


def _pieces(W, K_pad, prec=3):
    """fp32 [N, K] -> packed K=16 pieces: [piece][hi|lo][2 chunks][N][8] bf16, or [piece][2 chunks][N][8] fp16 (as uint8)."""
    N, K = W.shape
    Wp = torch.zeros(N, K_pad, dtype=torch.float32, device=W.device)
    Wp[:, :K] = W
    planes = _split_planes(Wp, prec)
    t = torch.stack(planes, 0).view(len(planes), N, K_pad // 16, 2, 8).permute(2, 0, 3, 1, 4).contiguous()
    return t.view(-1).view(torch.uint8)


def chain_supported(view, layers, last_relu_only=True):
    """2 or 3 folded layers, widths <= 256 and multiples of 16 (the last one may be narrower: it is zero-padded),
    hidden activations ReLU, rows a multiple of 128, every segment 16-byte aligned."""
    from ._lib import ACT_RELU
    if len(layers) not in (2, 3) or view.rows % 128 != 0:
        return False
    for li, (W, b, act) in enumerate(layers):
        last = li == len(layers) - 1
        if W.shape[0] > 256 or (not last and (W.shape[0] % 16 or act != ACT_RELU)):
            return False
        if last and last_relu_only and act != ACT_RELU:
            return False
    for mat, mode, ch, col0, scale in view.segs:
        if ch % 4 or col0 % 4 or mat.stride(0) % 4 or mat.data_ptr() % 16:
            return False
    return True


def pack_chain(layers, seg_channels, prec=3):
    key = tuple((W.data_ptr(), W._version) for W, _, _ in layers) + (tuple(seg_channels), prec)
    hit = _chain_cache.get(key)
    if hit is not None:
        return hit
    W1 = layers[0][0]
    chunks = sum((c + 7) // 8 for c in seg_channels)
    chunks0 = (chunks + 1) // 2 * 2
    W1p = torch.zeros(W1.shape[0], chunks0 * 8, dtype=torch.float32, device=W1.device)
    src = dst = 0
    for c in seg_channels:
        W1p[:, dst:dst + c] = W1[:, src:src + c]
        src += c
        dst += (c + 7) // 8 * 8
    mats = [W1p] + [W for W, _, _ in layers[1:]]
    cout = mats[-1].shape[0]
    np_last = max(16, (cout + 15) // 16 * 16)
    if np_last != cout:                                   # zero rows: padded output columns are never stored
        pad = torch.zeros(np_last, mats[-1].shape[1], dtype=torch.float32, device=W1.device)
        pad[:cout] = mats[-1]
        mats[-1] = pad
    Wpack = torch.cat([_pieces(m, m.shape[1], prec) for m in mats]).contiguous()
    bias = torch.cat([b for _, b, _ in layers]).contiguous()
    widths = [m.shape[0] for m in mats]
    return _remember(_chain_cache, key, layers[0][0], (Wpack, bias, chunks0, widths, cout, None))


def chain(view, layers, mode, kseg=8, want_rows=True, want_groups=True, last_act=None, prec=None):
    """Runs 2-3 folded layers on the virtual rows.  Returns (Y rows | None, G groups | None, a rows | None)."""
    global SPLIT_K
    prec = engine.mma_prec() if prec is None else prec
    Zb = Zg = None
    SPLIT_K = 0
    if SPLIT_CHAINS and len(layers) >= 2 and (kseg == 8 or all(sg[1] != engine.SEG_BROADCAST for sg in view.segs)):
        split = _split_first_layer(view, layers)
        if split is not None:
            view, layers, Zb, Zg = split
    Wpack, bias, chunks0, widths, cout, _ = pack_chain(layers, [s[2] for s in view.segs], prec)
    nl = len(layers)
    act = layers[-1][2] if last_act is None else last_act
    n = widths + [16] * (3 - nl)
    dev = bias.device
    Y = torch.empty(view.rows, cout, dtype=torch.float32, device=dev) if want_rows else None
    G = torch.empty(view.rows // kseg, cout, dtype=torch.float32, device=dev) if (want_groups and mode != EPI_STORE) else None
    a = torch.empty(view.rows, dtype=torch.float32, device=dev) if mode == EPI_ATTN else None
    engine.call("hrn_chain_tc", ctypes.byref(view.c), engine.ptr(Wpack), engine.ptr(bias), nl, n[0], n[1], n[2], cout, act,
                chunks0, mode, kseg, engine.ptr(Y), cout, engine.ptr(G), engine.ptr(a), view.rows, prec,
                _view_ptr(Zb), _view_ptr(Zg), Zg.stride(0) if Zg is not None else 0, engine.stream())
    return Y, G, a


# ------------------------------------------------------------------------------------------------------------------
# wide chain on a 2-CTA cluster (csrc/chain_wide.cu): CoarseReg convs_1 + attention tail
# ------------------------------------------------------------------------------------------------------------------
_wide_cache = {}
# accounting hooks for bench.py (algorithmic work = the reference's MACs): input channels of the first layer that
# _split_first_layer applied once per point in the last chain_wide call, and a flag that is set while those per-point
# launches run (their MACs are part of that first layer, not extra algorithmic work)
SPLIT_K = 0
IN_SPLIT = False
_zprefetch = {}          # (points data_ptr, rows, channels, col0, W1 data_ptr, W1 version, W1 column) -> (Z, side stream)
SPLIT_CHAINS = True      # apply the first-layer split in the single-CTA chain kernel too (FineReg convs_1, CoarseReg convs_2)


def chain_wide_supported(view, layers, k):
    """Three folded ReLU layers, every width a multiple of 64 and <= 512, groups of 8 rows, rows a multiple of 128."""
    from ._lib import ACT_RELU
    if len(layers) != 3 or k != 8 or view.rows % 128 != 0:
        return False
    for W, b, act in layers:
        if W.shape[0] % 64 or W.shape[0] > 512 or act != ACT_RELU:
            return False
    if max(W.shape[0] for W, _, _ in layers) <= 256:
        return False                                      # the single-CTA chain kernel covers these
    for mat, mode, ch, col0, scale in view.segs:
        if ch % 4 or col0 % 4 or mat.stride(0) % 4 or mat.data_ptr() % 16:
            return False
    return True


def pack_chain_wide(layers, seg_channels, prec=3):
    """-> (Wpack uint8 [2 * rank_bytes], rank_bytes, bias fp32 [2 * sum(n/2)], chunks0, widths).  Rank r of the cluster
    computes output columns [r n/2, (r+1) n/2) of every layer.  Layer 1's K order = the virtual rows (segments padded to 8
    channels, total to 16); layers 2 / 3 consume the previous layer's 32-column blocks in the order both CTAs produce
    them -- block j of rank 0, block j of rank 1, ... -- so their weight columns are permuted accordingly."""
    key = tuple((W.data_ptr(), W._version) for W, _, _ in layers) + (tuple(seg_channels), prec, "wide")
    hit = _wide_cache.get(key)
    if hit is not None:
        return hit
    W1 = layers[0][0]
    chunks = sum((c + 7) // 8 for c in seg_channels)
    chunks0 = (chunks + 1) // 2 * 2
    W1p = torch.zeros(W1.shape[0], chunks0 * 8, dtype=torch.float32, device=W1.device)
    src = dst = 0
    for c in seg_channels:
        W1p[:, dst:dst + c] = W1[:, src:src + c]
        src += c
        dst += (c + 7) // 8 * 8
    mats = [W1p]
    for li in (1, 2):
        W = layers[li][0]
        h = layers[li - 1][0].shape[0] // 2               # previous layer's columns per rank
        order = torch.cat([torch.arange(rk * h + 32 * j, rk * h + 32 * j + 32) for j in range(h // 32) for rk in (0, 1)])
        mats.append(W[:, order.to(W.device)].contiguous())
    widths = [m.shape[0] for m in mats]
    per_rank, bias = [], []
    for rk in (0, 1):
        per_rank.append(torch.cat([_pieces(m[rk * (n // 2):(rk + 1) * (n // 2)].contiguous(), m.shape[1], prec)
                                   for m, n in zip(mats, widths)]))
        bias += [b[rk * (n // 2):(rk + 1) * (n // 2)] for (_, b, _), n in zip(layers, widths)]
    assert per_rank[0].numel() == per_rank[1].numel()
    Wpack = torch.cat(per_rank).contiguous()
    return _remember(_wide_cache, key, W1, (Wpack, per_rank[0].numel(), torch.cat(bias).contiguous(), chunks0, widths))


def _split_first_layer(view, layers):
    """The first layer is linear in front of its ReLU, so the input segments that are constant inside a group (BROADCAST) or
    depend on the gathered source row only (GATHER) can be applied ONCE PER POINT instead of once per row:
        W1 x = W_direct x_direct + (W_b x_b)[r / k] + (W_g x_g)[b * N + idx[r]].
    Returns (direct-only view, layers with the reduced first layer, Zb [groups, n1] | None, Zg [src rows, n1]) or None when
    the view does not have that shape (one GATHER segment, at most one BROADCAST segment, no row scales, >= 1 DIRECT segment)."""
    from ._lib import ACT_NONE, SEG_BROADCAST, SEG_DIRECT, SEG_GATHER
    W1, b1, act1 = layers[0]
    kinds = [sg[1] for sg in view.segs]
    if kinds.count(SEG_BROADCAST) > 1 or kinds.count(SEG_GATHER) != 1 or SEG_DIRECT not in kinds:
        return None
    if W1.shape[0] % 32:
        return None
    if any(sg[4] is not None for sg in view.segs):
        return None
    reduced = engine.RowsView(view.rows, group=view.c.group, gather_idx=view.gather_idx, rows_per_batch=view.c.rows_per_batch,
                              src_rows_per_batch=view.c.src_rows_per_batch)
    cols, c0, pts = [], 0, {}
    for mat, mode, ch, col0, _ in view.segs:
        if mode == SEG_DIRECT:
            reduced.add(mat, SEG_DIRECT, channels=ch, col0=col0)
            cols.append(W1[:, c0:c0 + ch])
        else:
            pts[mode] = (mat, ch, col0, c0)
        c0 += ch
    global SPLIT_K
    Z = {}
    SPLIT_K = sum(p[1] for p in pts.values())
    for mode, (mat, ch, col0, cs) in pts.items():
        hit = _zprefetch.pop(_zkey(mat, ch, col0, W1, cs), None)
        if hit is not None:                      # launched ahead on a side stream (prefetch_point_layers): join it
            torch.cuda.current_stream(mat.device).wait_stream(hit[1])
            hit[0].record_stream(torch.cuda.current_stream(mat.device))
            Z[mode] = hit[0]
        else:
            Z[mode] = _point_layer(mat, ch, col0, W1, cs)
    Zb, Zg = Z.get(SEG_BROADCAST), Z[SEG_GATHER]
    key = (W1.data_ptr(), W1._version, "direct")
    Wd = _wide_cache.get(key)
    if Wd is None:
        Wd = _remember(_wide_cache, key, W1, torch.cat(cols, 1).contiguous())
    return reduced, [(Wd, b1, act1)] + list(layers[1:]), Zb, Zg


def _zkey(mat, ch, col0, W1, cs):
    return (mat.data_ptr(), mat.shape[0], ch, col0, W1.data_ptr(), W1._version, cs)


def _point_layer(mat, ch, col0, W1, cs):
    """Z [points, n1] = points[:, col0:col0+ch] . W1[:, cs:cs+ch]^T  (fp32 out, no bias, no activation): the per-point part
    of a chain's first layer."""
    from ._lib import ACT_NONE, SEG_DIRECT
    global IN_SPLIT
    key = (W1.data_ptr(), W1._version, cs, ch)
    Wc = _wide_cache.get(key)
    if Wc is None:
        Wc = _remember(_wide_cache, key, W1, W1[:, cs:cs + ch].contiguous())
    src = engine.RowsView(mat.shape[0]).add(mat, SEG_DIRECT, channels=ch, col0=col0)
    IN_SPLIT = True
    try:
        return engine.layer(src, Wc, None, ACT_NONE)
    finally:
        IN_SPLIT = False


def prefetch_point_layers(like, W1, specs):
    """Launches the per-point first-layer parts `specs` = [(points [P, C], channels, col0, W1 column), ...] on the API side
    stream, forked from the current stream NOW, and parks the results for the chain launch that will ask for them
    (_split_first_layer joins the stream).  They depend on the features only: a correspondence stage starts them before its
    latency-bound neighbour search / geometry kernels."""
    if not SPLIT_CHAINS or W1.shape[0] % 32 or not like.is_cuda:
        return
    if len(_zprefetch) > 16:                     # results nobody asked for (a stage that took another path)
        _zprefetch.clear()
    side = engine._api_stream(like.device)
    side.wait_stream(torch.cuda.current_stream(like.device))
    with torch.cuda.stream(side):
        for mat, ch, col0, cs in specs:
            _zprefetch[_zkey(mat, ch, col0, W1, cs)] = (_point_layer(mat, ch, col0, W1, cs), side)


def _view_ptr(t):
    """Device pointer of a 2-D view whose rows are contiguous (engine.ptr insists on fully contiguous tensors)."""
    if t is None:
        return None
    assert t.is_cuda and t.dim() == 2 and t.stride(1) == 1
    return t.data_ptr()


def chain_wide_head_supported(view, layers):
    """Three layers, the two hidden ones ReLU and multiples of 64 up to 512 wide (at least one > 256), one output column."""
    from ._lib import ACT_RELU
    if len(layers) != 3 or view.rows % 128 != 0 or len(view.segs) != 1:
        return False
    (W1, _, a1), (W2, _, a2), (W3, _, _) = layers
    if a1 != ACT_RELU or a2 != ACT_RELU or W3.shape[0] != 1:
        return False
    if W1.shape[0] % 64 or W2.shape[0] % 64 or max(W1.shape[0], W2.shape[0]) > 512 or max(W1.shape[0], W2.shape[0]) <= 256:
        return False
    mat, mode, ch, col0, scale = view.segs[0]
    return not (ch % 4 or col0 % 4 or mat.stride(0) % 4 or mat.data_ptr() % 16 or scale is not None)


def chain_wide_head(view, layers, act, prec=None):
    """Per-keypoint head c -> c -> c -> 1 (layers.py:391-394) in ONE launch of the cluster kernel: returns act(y) [rows]."""
    prec = engine.mma_prec() if prec is None else prec
    W3, b3, a3 = layers[2]
    key = (W3.data_ptr(), W3._version, "head64")
    padded = _wide_cache.get(key)
    if padded is None:                                     # last layer zero-padded to 64 output columns
        Wp = torch.zeros(64, W3.shape[1], dtype=torch.float32, device=W3.device)
        Wp[:1] = W3
        bp = torch.zeros(64, dtype=torch.float32, device=W3.device)
        bp[:1] = b3
        padded = _remember(_wide_cache, key, W3, (Wp, bp))
    lay = [layers[0], layers[1], (padded[0], padded[1], a3)]
    Wpack, rank_bytes, bias, chunks0, widths = pack_chain_wide(lay, [s[2] for s in view.segs], prec)
    Y = torch.empty(view.rows, dtype=torch.float32, device=bias.device)
    engine.call("hrn_chain_wide_head", ctypes.byref(view.c), engine.ptr(Wpack), rank_bytes, engine.ptr(bias), widths[0],
                widths[1], widths[2], chunks0, act, engine.ptr(Y), view.rows, prec, engine.stream())
    return Y


def chain_wide(view, layers, kseg=8, prec=None, split_first=True):
    """convs + attention tail on the virtual rows: returns (AF [rows / kseg, n3], a [rows])  (layers.py:364-390)."""
    prec = engine.mma_prec() if prec is None else prec
    global SPLIT_K
    Zb = Zg = None
    SPLIT_K = 0
    split = _split_first_layer(view, layers) if split_first else None
    if split is not None:
        view, layers, Zb, Zg = split
    Wpack, rank_bytes, bias, chunks0, widths = pack_chain_wide(layers, [s[2] for s in view.segs], prec)
    dev = bias.device
    G = torch.empty(view.rows // kseg, widths[2], dtype=torch.float32, device=dev)
    a = torch.empty(view.rows, dtype=torch.float32, device=dev)
    engine.call("hrn_chain_wide", ctypes.byref(view.c), engine.ptr(Wpack), rank_bytes, engine.ptr(bias), widths[0], widths[1],
                widths[2], chunks0, kseg, _view_ptr(Zb), _view_ptr(Zg), Zg.stride(0) if Zg is not None else 0,
                engine.ptr(G), engine.ptr(a), view.rows, prec, engine.stream())
    return G, a


def chain3(view, layers, mode, kseg, want_rows=True, want_groups=True):
    return chain(view, layers, mode, kseg, want_rows, want_groups)
