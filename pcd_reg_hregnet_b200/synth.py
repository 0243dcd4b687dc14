"""Seeded synthetic LiDAR registration pairs: no dataset is reachable offline.

The scene mimics what the reference's loaders hand to the model (dataset/dataset_utils.py:113-125,188-223;
dataset/config.json:16,20-25): an ego-centred sweep, range < 80 m, ~60 % ground returns and ~40 % returns
from vertical structure, float32 [N,3]; the source cloud is an independent re-sampling of the same scene
moved by a uniform twist with |angle| <= 20 deg and |t| <= 0.5 m per axis
(transform/dataset_transforms.py:86-90).  Generated on the CPU from `torch.Generator(seed)` so that the
CPU oracle, the fixtures and the GPU see identical bytes.
"""
import math

import torch


def _scene_planes(gen, n_planes=24):
    c = (torch.rand(n_planes, 2, generator=gen) * 2 - 1) * 60.0          # plane centre (x,y)
    yaw = torch.rand(n_planes, generator=gen) * math.pi
    half = torch.rand(n_planes, generator=gen) * 10.0 + 2.0              # half-length
    return c, yaw, half


def _sample_cloud(gen, n, planes):
    n_ground = int(0.6 * n)
    n_struct = n - n_ground
    r = torch.rand(n_ground, generator=gen) * 78.0 + 2.0
    th = torch.rand(n_ground, generator=gen) * 2 * math.pi
    ground = torch.stack([r * torch.cos(th), r * torch.sin(th),
                          -1.8 + 0.03 * torch.randn(n_ground, generator=gen)], -1)
    c, yaw, half = planes
    pid = torch.randint(0, c.shape[0], (n_struct,), generator=gen)
    u = (torch.rand(n_struct, generator=gen) * 2 - 1) * half[pid]
    sx = c[pid, 0] + u * torch.cos(yaw[pid])
    sy = c[pid, 1] + u * torch.sin(yaw[pid])
    sz = torch.rand(n_struct, generator=gen) * 5.8 - 1.8
    struct = torch.stack([sx, sy, sz], -1)
    pts = torch.cat([ground, struct], 0) + 0.02 * torch.randn(n, 3, generator=gen)
    rng = pts.norm(dim=-1, keepdim=True).clamp_min(1e-6)
    pts = pts * torch.clamp(79.5 / rng, max=1.0)                          # keep range < 80 m
    return pts[torch.randperm(n, generator=gen)]


def _rodrigues(w):
    th = w.norm()
    if th < 1e-12:
        return torch.eye(3, dtype=torch.float64)
    k = w / th
    K = torch.tensor([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]], dtype=torch.float64)
    return torch.eye(3, dtype=torch.float64) + torch.sin(th) * K + (1 - torch.cos(th)) * (K @ K)


def make_pair(seed: int, n_points: int = 16384, max_deg: float = 20.0, max_t: float = 0.5):
    """Returns src [N,3], dst [N,3] float32 and the ground-truth (R [3,3], t [3]) with dst ~= R src + t."""
    gen = torch.Generator().manual_seed(int(seed))
    planes = _scene_planes(gen)
    dst = _sample_cloud(gen, n_points, planes)
    src_in_dst = _sample_cloud(gen, n_points, planes)
    w = (torch.rand(3, generator=gen, dtype=torch.float64) * 2 - 1) * math.radians(max_deg) / math.sqrt(3.0)
    t = (torch.rand(3, generator=gen, dtype=torch.float64) * 2 - 1) * max_t
    R = _rodrigues(w)
    src = ((src_in_dst.double() - t) @ R).float()                          # R^T (x - t)
    return src.contiguous(), dst.contiguous(), R.float(), t.float()


def make_batch(seeds, n_points: int = 16384):
    """Stacks pairs: src [B,N,3], dst [B,N,3], R [B,3,3], t [B,3]."""
    ps = [make_pair(s, n_points) for s in seeds]
    return tuple(torch.stack([p[i] for p in ps], 0).contiguous() for i in range(4))


def duplicate_padded_cloud(seed: int, n_points: int, n_unique: int):
    """A cloud padded with exact duplicate points (dataset/dataset_utils.py:203-207) -- exercises tie-breaks."""
    gen = torch.Generator().manual_seed(int(seed))
    base = _sample_cloud(gen, n_unique, _scene_planes(gen))
    pad = base[torch.randint(0, n_unique, (n_points - n_unique,), generator=gen)]
    return torch.cat([base, pad], 0).contiguous()


# ------------------------------------------------------------------------------------------------------------------
# seeded networks for benchmarks / smoke runs / tests (no checkpoint of the registration heads is reachable offline)
# ------------------------------------------------------------------------------------------------------------------
import os as _os

PRETRAINED_FEATS = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "tests", "golden",
                                 "nusc_feats_state.npz")


class Args:
    """The 4-attribute `args` contract of the reference models (models/HRegNet/models.py:11-12,18,67)."""
    use_fps = True
    use_weights = True
    freeze_detector = False
    freeze_feats = False


def pretrained_feats(path=None):
    """The reference's pretrained HierFeatureExtraction state_dict (ckpt/pretrained/nusc_feats.pth re-saved as npz)."""
    import numpy as np
    with np.load(path or PRETRAINED_FEATS) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


def randomize_bn_(module, gen):
    """Non-trivial BatchNorm running statistics / affine parameters for the seeded registration heads (default
    statistics would make BatchNorm folding a no-op).  Draw order is part of the fixture contract (tests/golden)."""
    for m in module.modules():
        if isinstance(m, (torch.nn.BatchNorm1d, torch.nn.BatchNorm2d)):
            n = m.num_features
            m.running_mean.copy_(torch.randn(n, generator=gen) * 0.1)
            m.running_var.copy_(torch.rand(n, generator=gen) * 0.5 + 0.75)
            m.weight.data.copy_(torch.rand(n, generator=gen) * 0.5 + 0.75)
            m.bias.data.copy_(torch.randn(n, generator=gen) * 0.1)


def build_net(model="hregnet", seed=7, device="cpu", args=None):
    """HRegNet / Model_V2 / Model_V4 of this package in eval mode: the reference's pretrained feature extractor +
    `torch.manual_seed(seed)` default-initialised registration heads with randomised BatchNorm statistics -- exactly
    the weights of the committed golden fixtures (tests/golden/make_golden.py builds the reference classes the same way)."""
    if model == "hregnet":
        from .models import HRegNet as cls
    elif model == "v2":
        from .model_v2 import Model_V2 as cls
    elif model == "v4":
        from .model_v4 import Model_V4 as cls
    else:
        raise ValueError(model)
    torch.manual_seed(seed)
    net = cls(args or Args())
    net.feature_extraction.load_state_dict(pretrained_feats())
    g = torch.Generator().manual_seed(seed + 1)
    for name in ("coarse_corres", "fine_corres_2", "fine_corres_1"):
        randomize_bn_(getattr(net, name), g)
    return net.eval().to(device)
