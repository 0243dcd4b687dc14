"""Shared helpers of the test-suite."""
import json
import os

import numpy as np
import torch

from oracle import ref_harness as H

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Args:
    use_fps = True
    use_weights = True
    freeze_detector = False
    freeze_feats = False


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


def pretrained_feats():
    return load_golden("nusc_feats_state")


def build_product_hregnet(seed=7, device="cpu"):
    """The product HRegNet with exactly the weights of oracle.ref_harness.build_reference_hregnet(seed): the
    reference's pretrained feature extractor (golden npz copy) + seeded default-initialised registration heads
    with randomised BatchNorm statistics.  Works without /root/reference."""
    from pcd_reg_hregnet_b200.models import HRegNet
    torch.manual_seed(seed)
    net = HRegNet(Args())
    net.feature_extraction.load_state_dict(pretrained_feats())
    g = torch.Generator().manual_seed(seed + 1)
    for name in ("coarse_corres", "fine_corres_2", "fine_corres_1"):
        H.randomize_bn_(getattr(net, name), g)
    return net.eval().to(device)


def unflatten(d, prefix):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def rel_err(a, b):
    """Per-tensor relative error  max|a-b| / max|b|  (element-wise relative error is
    meaningless on post-ReLU zeros)."""
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def build_product_model_v2(seed=7, device="cpu"):
    """Model_V2 (Adaption-1) with the pretrained feature extractor and seeded heads, as build_product_hregnet."""
    from pcd_reg_hregnet_b200.model_v2 import Model_V2
    torch.manual_seed(seed)
    net = Model_V2(Args())
    net.feature_extraction.load_state_dict(pretrained_feats())
    g = torch.Generator().manual_seed(seed + 1)
    for name in ("coarse_corres", "fine_corres_2", "fine_corres_1"):
        H.randomize_bn_(getattr(net, name), g)
    return net.eval().to(device)


def build_product_model_v4(seed=7, device="cpu"):
    """Model_V4 (coarse stage with coord_dist / feats_dist) with the pretrained feature extractor and seeded heads."""
    from pcd_reg_hregnet_b200.model_v4 import Model_V4
    torch.manual_seed(seed)
    net = Model_V4(Args())
    net.feature_extraction.load_state_dict(pretrained_feats())
    g = torch.Generator().manual_seed(seed + 1)
    for name in ("coarse_corres", "fine_corres_2", "fine_corres_1"):
        H.randomize_bn_(getattr(net, name), g)
    return net.eval().to(device)
