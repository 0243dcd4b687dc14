"""Host side of the tensor-core shared-MLP layer (csrc/mlp_tc.cu): weight packing + launch.

pack_weights(): fp32 [Cout, K] (BatchNorm already folded) -> bf16 hi/lo split, K padded per rows-segment to a
multiple of 8 (and in total to a multiple of 32), Cout padded to a multiple of 16, tiled per pipeline stage as
[n_stage][hi|lo][K/8 within stage = 4][NP][8] so that one stage is one contiguous cp.async.bulk.
Packed weights are cached per (weight tensor, segment signature)."""
import ctypes

import torch

from . import engine

KC = 32
_cache = {}


def pack_weights(W, seg_channels):
    key = (W.data_ptr(), W._version, tuple(seg_channels), W.device)
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
    hi = Wp.to(torch.bfloat16)
    lo = (Wp - hi.float()).to(torch.bfloat16)
    t = torch.stack([hi, lo], 0).view(2, NP, n_stage, KC // 8, 8).permute(2, 0, 3, 1, 4).contiguous()
    out = (t, NP, n_stage, W)          # keep W alive so data_ptr keys stay unique
    _cache[key] = out
    return out


def layer_tc(view, W, b, act, out):
    seg_channels = [s[2] for s in view.segs]
    Wp, NP, n_stage, _ = pack_weights(W, seg_channels)
    if b is None:
        b = torch.zeros(W.shape[0], dtype=torch.float32, device=W.device)
    engine.call("hrn_layer_tc", ctypes.byref(view.c), engine.ptr(Wp), engine.ptr(b), act, engine.ptr(out), out.stride(0),
                view.rows, W.shape[0], NP, n_stage, engine.stream())
    return out


# ------------------------------------------------------------------------------------------------------------------
# fused level-1 kernel (csrc/level_fused.cu)
# ------------------------------------------------------------------------------------------------------------------
_level1_cache = {}


def _umma_tiles(W, K_pad):
    """fp32 [N, K] -> bytes of the resident UMMA B operand: hi plane [K_pad/8][N][8] bf16, then the lo plane."""
    N, K = W.shape
    Wp = torch.zeros(N, K_pad, dtype=torch.float32, device=W.device)
    Wp[:, :K] = W
    hi = Wp.to(torch.bfloat16)
    lo = (Wp - hi.float()).to(torch.bfloat16)
    planes = [t.view(N, K_pad // 8, 8).permute(1, 0, 2).contiguous().view(-1) for t in (hi, lo)]
    return torch.cat(planes).view(torch.uint8)


def pack_level1(det, desc):
    """det / desc: folded parameter dicts of detector_1 / desc_extractor_1 -> (Wpack uint8, biases fp32) in the
    layout of LevelCfg<64,32,32,64,32,64> (csrc/level_fused.cu)."""
    key = tuple((W.data_ptr(), W._version) for W, _, _ in det["convs"] + desc["convs"] + desc["mlp"])
    hit = _level1_cache.get(key)
    if hit is not None:
        return hit
    (d1, bd1, _), (d2, bd2, _), (d3, bd3, _) = det["convs"]
    (x1, bx1, _), (x2, bx2, _), (x3, bx3, _) = desc["convs"]
    (m1, bm1, _), (m2, bm2, _) = desc["mlp"]
    CO = d3.shape[0]
    assert d1.shape == (32, 4) and d2.shape == (32, 32) and d3.shape == (64, 32) and m1.shape == (32, 3 * CO) and m2.shape == (64, 32)
    parts = [_umma_tiles(d1, 16), _umma_tiles(d2, 32), _umma_tiles(d3, 32),
             _umma_tiles(x1, 16), _umma_tiles(x2, 32), _umma_tiles(x3, 32),
             _umma_tiles(m1[:, :CO].contiguous(), CO), _umma_tiles(m1[:, CO:2 * CO].contiguous(), CO),
             _umma_tiles(m1[:, 2 * CO:].contiguous(), CO), _umma_tiles(m2, 32)]
    Wpack = torch.cat(parts).contiguous()
    biases = torch.cat([bd1, bd2, bd3, bx1, bx2, bx3, bm1, bm2]).contiguous()
    from ._lib import lib
    assert Wpack.numel() == lib().hrn_level1_pack_bytes() and biases.numel() == lib().hrn_level1_bias_count()
    out = (Wpack, biases, [t[0] for t in det["convs"] + desc["convs"] + desc["mlp"]])
    _level1_cache[key] = out
    return out


def level1_fused(q, xyz, idx, det, desc):
    """q [B,M,3], xyz [B,N,3], idx [B,M,64] int32 -> keypoints [B*M,3], af [B*M,64], desc [B*M,64]."""
    B, M, k = idx.shape
    N = xyz.shape[1]
    Wpack, biases, _ = pack_level1(det, desc)
    dev = xyz.device
    kp = torch.empty(B * M, 3, dtype=torch.float32, device=dev)
    af = torch.empty(B * M, 64, dtype=torch.float32, device=dev)
    d = torch.empty(B * M, 64, dtype=torch.float32, device=dev)
    engine.call("hrn_level1_fused", engine.ptr(q), engine.ptr(xyz), engine.ptr(idx), engine.ptr(Wpack), engine.ptr(biases),
                engine.ptr(kp), engine.ptr(af), engine.ptr(d), B, M, N, k, engine.stream())
    return kp, af, d
