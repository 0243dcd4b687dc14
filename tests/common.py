"""Shared helpers of the test-suite."""
import json
import os

import numpy as np
import torch

from pcd_reg_hregnet_b200 import synth
from pcd_reg_hregnet_b200.synth import Args  # noqa: F401  (re-exported)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


def pretrained_feats():
    return load_golden("nusc_feats_state")


def build_product_hregnet(seed=7, device="cpu"):
    """The product HRegNet with exactly the weights of oracle.ref_harness.build_reference_hregnet(seed): the
    reference's pretrained feature extractor (golden npz copy) + seeded default-initialised registration heads
    with randomised BatchNorm statistics.  Works without /root/reference."""
    return synth.build_net("hregnet", seed, device)


def build_product_model_v2(seed=7, device="cpu"):
    return synth.build_net("v2", seed, device)


def build_product_model_v4(seed=7, device="cpu"):
    return synth.build_net("v4", seed, device)


def unflatten(d, prefix):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def rel_err(a, b):
    """Per-tensor relative error  max|a-b| / max|b|  (element-wise relative error is
    meaningless on post-ReLU zeros)."""
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


