"""BatchNorm (eval) folding and weight-column permutation for the shared-MLP kernels.

The reference stacks are  Conv(1x1, bias=False) -> BatchNorm -> ReLU  (layers.py:118-121,186-198,249-260,420-423)
and  Conv1d(bias=True) -> BatchNorm1d -> ReLU  (layers.py:124-130,262-268,425-431).  In eval mode
    BN(x) = (x - running_mean) / sqrt(running_var + eps) * weight + bias
so each triple is  y = relu(W' x + b')  with  W' = diag(s) W,  b' = beta + s (conv_bias - mean),  s = gamma/sqrt(var+eps).

The kernels read the grouped feature channels in the order [small geometry/weight/similarity block | broadcast
block | gathered block]; `perm` re-orders the columns of the first layer's weight accordingly (free, host side).
"""
import torch
import torch.nn as nn

from ._lib import ACT_NONE, ACT_RELU


def _fold_conv_bn(conv, bn):
    W = conv.weight.detach().reshape(conv.out_channels, -1).float()
    cb = conv.bias.detach().float() if conv.bias is not None else torch.zeros(conv.out_channels, device=W.device)
    if bn is None:
        return W.contiguous(), cb.contiguous()
    s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return (W * s[:, None]).contiguous(), (bn.bias.detach().float() + s * (cb - bn.running_mean.detach().float())).contiguous()


def fold_sequential(seq, perm=None):
    """nn.Sequential of [Conv, BN, ReLU]* (or a bare Conv) -> [(W, b, act), ...]."""
    mods = list(seq)
    out = []
    i = 0
    while i < len(mods):
        conv = mods[i]
        assert isinstance(conv, (nn.Conv1d, nn.Conv2d)), type(conv)
        bn = mods[i + 1] if i + 1 < len(mods) and isinstance(mods[i + 1], (nn.BatchNorm1d, nn.BatchNorm2d)) else None
        j = i + (2 if bn is not None else 1)
        relu = j < len(mods) and isinstance(mods[j], nn.ReLU)
        W, b = _fold_conv_bn(conv, bn)
        if perm is not None and not out:
            W = W[:, perm].contiguous()
        out.append((W, b, ACT_RELU if relu else ACT_NONE))
        i = j + (1 if relu else 0)
    return out


def fold_head(mlp1, mlp2, mlp3):
    return fold_sequential(mlp1) + fold_sequential(mlp2) + fold_sequential(mlp3)


def pair_perm(C, n_sim, device):
    """Column permutation for CoarseReg / FineReg first layers.  Reference channel order (layers.py:364-380,
    444-445):  [rel(3) dist(1) src_xyz(3) nbr_xyz(3) | src_feat(C) | nbr_feat(C) | src_w dst_w | sims(n_sim)]
    kernel order:  [the 10 geometry channels, src_w, dst_w, sims | src_feat(C) | nbr_feat(C)]."""
    geo = list(range(10))
    wts = [10 + 2 * C, 11 + 2 * C]
    sims = [12 + 2 * C + i for i in range(n_sim)]
    feats = list(range(10, 10 + 2 * C))
    return torch.tensor(geo + wts + sims + feats, dtype=torch.long, device=device)
