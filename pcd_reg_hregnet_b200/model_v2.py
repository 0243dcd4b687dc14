"""Drop-in `Model_V2` ("Adaption-1": mutual-information loss after the coarse stage) and its `FineReg1` / `FineReg2`
blocks -- reference models/model_v2/models.py:60-183 and models/model_v2/layers.py:368-501.  Same constructor, forward
signature, returned keys / shapes and state_dict keys as the reference.  The backbone, CoarseReg, the SVD head and the
pose cascade are the HRegNet ones (the reference duplicates them verbatim in model_v2/layers.py:55-366,503-551).

`FineReg2` differs from `FineReg` by one more head `mlpx` (Conv1d 2C->C + BN + ReLU on the attentive features) and by
two batch-shuffled copies (`*_prime`) drawn with the HOST generator `torch.randperm(B)` in the reference's order
(layers.py:491-497) -- call `torch.manual_seed(s)` before forward to reproduce them.
"""
import torch
import torch.nn as nn

from . import engine, fold
from .engine import RowsView
from .layers import CoarseReg, FineReg, WeightedSVDHead, _cl
from .models import HierFeatureExtraction


class FineReg1(FineReg):
    """Reference: model_v2/layers.py:368-424 (identical to HRegNet's FineReg)."""


class FineReg2(FineReg):
    """Reference: model_v2/layers.py:426-501.  forward(...) -> (corres_xyz [B,N,3], weights [B,N], weights_prime [B,N],
    attentive_features [B,C,N], attentive_features_prime [B,C,N])."""

    def __init__(self, k, in_channels):
        super().__init__(k, in_channels)
        self.mlpx = nn.Sequential(nn.Conv1d(in_channels * 2, in_channels, kernel_size=1), nn.BatchNorm1d(in_channels),
                                  nn.ReLU())

    def _fold(self):
        P = super()._fold()
        P["mlpx"] = fold.fold_sequential(self.mlpx)
        return P

    # The two batch shuffles are HOST draws in the reference.  For a forward captured in a CUDA graph they cannot be
    # drawn inside the forward: draw_permutations() draws them before each replay -- same generator, same order as the
    # reference (layers.py:493, :497) -- into two device index buffers which the captured forward reads.
    drawn_slot = None       # set only while a Registrar warms up / captures a forward: an eager forward draws for itself

    def _perm_buffers(self, batch, device, slot, create=False):
        # one pair of buffers per (batch, device, slot), never reallocated: captured graphs hold their addresses
        # (slot = which of a Registrar's captures of this net reads them; forwards in flight must not share a pair)
        if "_perm" not in self.__dict__:
            self.__dict__["_perm"] = {}
        key = (int(batch), torch.device(device), int(slot))
        if key not in self._perm and create:
            self._perm[key] = (torch.empty(batch, dtype=torch.int64, device=device),
                               torch.empty(batch, dtype=torch.int64, device=device))
        return self._perm.get(key)

    def draw_permutations(self, batch, device, slot=0):
        p = self._perm_buffers(batch, device, slot, create=True)
        p[0].copy_(torch.randperm(batch))
        p[1].copy_(torch.randperm(batch))

    def forward_cl(self, sxyz, sfeat_cl, dxyz, dfeat_cl, ssig, dsig):
        P = self.folded()
        B, N1, _ = sxyz.shape
        cor, w, af = engine.fine_reg(sxyz, sfeat_cl, dxyz, dfeat_cl, ssig, dsig, P, self.k, want_af=True)
        (Wx, bx, act), = P["mlpx"]
        feats = engine.transpose(engine.layer(RowsView(B * N1).add(af), Wx, bx, act).view(B, N1, -1))   # [B,C,N]
        drawn = self._perm_buffers(B, feats.device, self.drawn_slot) if self.drawn_slot is not None else None
        if drawn is not None:                                       # drawn ahead of a captured forward
            feats_prime = feats.index_select(0, drawn[0])
            w_prime = w.index_select(0, drawn[1])
        else:
            feats_prime = feats[torch.randperm(feats.size(0))]      # host RNG, reference order (layers.py:493)
            w_prime = w[torch.randperm(w.size(0))]                  # (layers.py:497)
        return cor, w, w_prime, feats, feats_prime

    def forward(self, src_xyz, src_feat, dst_xyz, dst_feat, src_weights, dst_weights):
        return self.forward_cl(src_xyz.contiguous(), _cl(src_feat), dst_xyz.contiguous(), _cl(dst_feat),
                               src_weights.contiguous(), dst_weights.contiguous())


class RegressionHead(nn.Module):
    """Reference: model_v2/layers.py:625-668 (the pose head of model_v3): the pose is regressed from the two weighted means.
    forward(src [B,N,3], src_corres [B,N,3], weights [B,N]) -> (rotation [B,3], translation [B,3]).  Same parameter names
    / shapes (fc{1,2,3}_{rot,trans}); one kernel launch (csrc/kabsch.cu: hrn_regression_head)."""
    DIMS = (128, 64, 3)

    def __init__(self):
        super().__init__()
        h1, h2, n_rot = self.DIMS
        self.fc1_rot, self.fc2_rot, self.fc3_rot = nn.Linear(6, h1), nn.Linear(h1, h2), nn.Linear(h2, n_rot)
        self.fc1_trans, self.fc2_trans, self.fc3_trans = nn.Linear(6, h1), nn.Linear(h1, h2), nn.Linear(h2, 3)

    def _regress(self, src, src_corres, weights):
        import ctypes
        h1, h2, n_rot = self.DIMS
        B, N, _ = src.shape
        ps = [p.detach() for fc in (self.fc1_rot, self.fc2_rot, self.fc3_rot, self.fc1_trans, self.fc2_trans, self.fc3_trans)
              for p in (fc.weight, fc.bias)]
        arr = (ctypes.c_void_p * 12)(*[engine.ptr(p) for p in ps])
        rot = torch.empty(B, n_rot, dtype=torch.float32, device=src.device)
        trans = torch.empty(B, 3, dtype=torch.float32, device=src.device)
        engine.call("hrn_regression_head", engine.ptr(src.contiguous()), engine.ptr(src_corres.contiguous()),
                    engine.ptr(weights.contiguous()), B, N, arr, h1, h2, n_rot, engine.ptr(rot), engine.ptr(trans), engine.stream())
        return rot, trans

    def forward(self, src, src_corres, weights):
        return self._regress(src, src_corres, weights)


class Regression_6dR_3dt_Head(RegressionHead):
    """Reference: model_v2/layers.py:555-623.  As shipped, the reference's translation branch cannot run: `fc3_trans` is
    Linear(64, 3) but receives the 32 outputs of `fc2_trans` (layers.py:566 vs :595-596), so its forward raises in
    F.linear.  The constructor (parameter names and shapes, the mismatch included, so checkpoints load) and that error
    behaviour are kept; `compute_rotation_matrix_from_6d` (layers.py:605-622) is provided on the device."""
    DIMS = (64, 32, 6)

    def __init__(self):
        super().__init__()
        self.fc3_trans = nn.Linear(64, 3)            # sic (layers.py:566)

    def forward(self, src, src_corres, weights):
        raise RuntimeError("mat1 and mat2 shapes cannot be multiplied (%dx32 and 64x3): the reference's "
                           "Regression_6dR_3dt_Head.fc3_trans expects 64 inputs (model_v2/layers.py:566,596)" % src.shape[0])

    @staticmethod
    def compute_rotation_matrix_from_6d(x):
        b = x.view(x.shape[0], 3, 2)

        def l2n(v):
            return v / (torch.sqrt(torch.sum(v ** 2, dim=1, keepdim=True)) + 1e-6)
        b1 = l2n(b[:, :, 0])
        b2 = l2n(b[:, :, 1] - b1 * torch.sum(b1 * b[:, :, 1], dim=1, keepdim=True))
        return torch.stack([b1, b2, torch.cross(b1, b2, dim=1)], dim=-1)


class Model_V2(nn.Module):
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
        """What a forward draws from the HOST generator, drawn ahead of it (runner.Registrar calls this before every
        graph replay, with the slot of the capture it is about to replay): the two batch shuffles of FineReg2."""
        self.fine_corres_2.draw_permutations(batch, device, slot)

    def bind_host_draws(self, slot):
        """slot (int): forwards read the permutations host_prologue(..., slot) drew -- set while a forward is being
        captured; None: forwards draw for themselves."""
        self.fine_corres_2.drawn_slot = slot

    def forward(self, src_points, dst_points):
        B = src_points.shape[0]
        both = self.feature_extraction.forward_cl(engine.stack_clouds(src_points, dst_points), calls=2)
        S = {k: v[:B] for k, v in both.items()}
        D = {k: v[B:] for k, v in both.items()}
        cor3, w3 = self.coarse_corres.forward_cl(S["xyz_3"], S["desc_3"], D["xyz_3"], D["desc_3"], S["sigmas_3"],
                                                 D["sigmas_3"], both=(both["xyz_3"], both["desc_3"]))
        R3, t3 = engine.weighted_kabsch(S["xyz_3"], cor3, w3)
        xyz2_t = engine.transform_points(S["xyz_2"], R3, t3)
        cor2, w2, w2_prime, f2, f2_prime = self.fine_corres_2.forward_cl(xyz2_t, S["desc_2"], D["xyz_2"], D["desc_2"],
                                                                         S["sigmas_2"], D["sigmas_2"])
        _, _, R2, t2 = engine.weighted_kabsch(xyz2_t, cor2, w2, prev=(R3, t3))
        xyz1_t = engine.transform_points(S["xyz_1"], R2, t2)
        cor1, w1 = self.fine_corres_1.forward_cl(xyz1_t, S["desc_1"], D["xyz_1"], D["desc_1"], S["sigmas_1"],
                                                 D["sigmas_1"])
        _, _, R1, t1 = engine.weighted_kabsch(xyz1_t, cor1, w1, prev=(R2, t2), packed=True)

        def api(d):
            return {k: (engine.transpose(v) if k.startswith("desc_") else v) for k, v in d.items()}

        src_feats, dst_feats = api(S), api(D)
        return {
            "src_xyz_corres_3": cor3, "src_xyz_corres_2": cor2, "src_xyz_corres_1": cor1,
            "rotation": [R3, R2, R1], "translation": [t3, t2, t1],
            "src_feats_desc_2": src_feats["desc_2"], "src_feats_sigmas_2": src_feats["sigmas_2"],
            "src_xyz_2_trans": xyz2_t, "dst_xyz_2": dst_feats["xyz_2"],
            "src_dst_feats_2": f2, "src_dst_feats_2_prime": f2_prime,
            "src_dst_weights_2": w2, "src_dst_weights_2_prime": w2_prime,
            "src_feats": src_feats, "dst_feats": dst_feats,
        }
