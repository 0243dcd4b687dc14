"""Drop-in `Model_V4` and its `CoarseReg` -- reference models/model_v4/models.py:60-183 and
models/model_v4/layers.py:211-369 (model variants on the same ops).  Model_V4 is Model_V2 (Adaption-1) whose
coarse stage additionally returns `coord_dist` [B,N,k] (distance of every correspondence candidate to its source
keypoint, layers.py:252) and `feats_dist` [B,N,k] (1 - normalised dst->src cosine similarity of the candidate,
layers.py:282); its forward returns the loss inputs only (no `src_feats` / `dst_feats` / per-level correspondences,
models.py:160-181).  Same constructor, forward signature, returned keys / shapes and state_dict keys as the reference;
both extra tensors are by-products of the feature assembly the coarse stage already does (engine.coarse_reg)."""
import torch
import torch.nn as nn

from . import engine, layers
from .layers import WeightedSVDHead, _cl
from .model_v2 import FineReg1, FineReg2
from .models import HierFeatureExtraction


class CoarseReg(layers.CoarseReg):
    """Reference: model_v4/layers.py:211-369.  forward(...) -> (corres_xyz [B,N,3], weights [B,N], coord_dist [B,N,k],
    feats_dist [B,N,k])."""

    def forward(self, src_xyz, src_desc, dst_xyz, dst_desc, src_weights, dst_weights):
        if not (self.use_sim and self.use_neighbor):
            raise NotImplementedError("only use_sim=use_neighbor=True (the Model_V4 configuration, models.py:71) is built")
        return engine.coarse_reg(src_xyz.contiguous(), _cl(src_desc.float()), dst_xyz.contiguous(), _cl(dst_desc.float()),
                                 src_weights.contiguous(), dst_weights.contiguous(), self.folded(), self.k, want_dists=True)

    def forward_cl(self, sxyz, sdesc_cl, dxyz, ddesc_cl, ssig, dsig, both=None):
        return engine.coarse_reg(sxyz, sdesc_cl, dxyz, ddesc_cl, ssig, dsig, self.folded(), self.k, both=both,
                                 want_dists=True)


class Model_V4(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.feature_extraction = HierFeatureExtraction(args)
        if args.freeze_feats:
            for p in self.parameters():
                p.requires_grad = False
        self.coarse_corres = CoarseReg(k=8, in_channels=256, use_sim=True, use_neighbor=True)
        self.fine_corres_2 = FineReg2(k=8, in_channels=128)
        self.fine_corres_1 = FineReg1(k=8, in_channels=64)
        self.svd_head = WeightedSVDHead()

    def host_prologue(self, batch, device, slot=0):
        """See Model_V2.host_prologue."""
        self.fine_corres_2.draw_permutations(batch, device, slot)

    def bind_host_draws(self, slot):
        self.fine_corres_2.drawn_slot = slot

    def forward(self, src_points, dst_points):
        B = src_points.shape[0]
        both = self.feature_extraction.forward_cl(engine.stack_clouds(src_points, dst_points), calls=2)
        S = {k: v[:B] for k, v in both.items()}
        D = {k: v[B:] for k, v in both.items()}
        cor3, w3, coord_dist, feats_dist = self.coarse_corres.forward_cl(
            S["xyz_3"], S["desc_3"], D["xyz_3"], D["desc_3"], S["sigmas_3"], D["sigmas_3"], both=(both["xyz_3"], both["desc_3"]))
        R3, t3 = engine.weighted_kabsch(S["xyz_3"], cor3, w3)
        xyz2_t = engine.transform_points(S["xyz_2"], R3, t3)
        cor2, w2, w2_prime, f2, f2_prime = self.fine_corres_2.forward_cl(xyz2_t, S["desc_2"], D["xyz_2"], D["desc_2"],
                                                                         S["sigmas_2"], D["sigmas_2"])
        _, _, R2, t2 = engine.weighted_kabsch(xyz2_t, cor2, w2, prev=(R3, t3))
        xyz1_t = engine.transform_points(S["xyz_1"], R2, t2)
        cor1, w1 = self.fine_corres_1.forward_cl(xyz1_t, S["desc_1"], D["xyz_1"], D["desc_1"], S["sigmas_1"],
                                                 D["sigmas_1"])
        _, _, R1, t1 = engine.weighted_kabsch(xyz1_t, cor1, w1, prev=(R2, t2), packed=True)
        return {
            "rotation": [R3, R2, R1], "translation": [t3, t2, t1],
            "src_feats_desc_2": engine.transpose(S["desc_2"]), "src_feats_sigmas_2": S["sigmas_2"],
            "src_xyz_2_trans": xyz2_t, "dst_xyz_2": D["xyz_2"],
            "src_dst_feats_2": f2, "src_dst_feats_2_prime": f2_prime,
            "src_dst_weights_2": w2, "src_dst_weights_2_prime": w2_prime,
            "coord_dist": coord_dist, "feats_dist": feats_dist,
        }
