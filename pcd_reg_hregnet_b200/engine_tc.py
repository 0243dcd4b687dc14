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
