"""Drop-in modules for the reference's models/HRegNet/layers.py: same class names, constructor signatures,
forward argument order / return tuples and state_dict keys (so `ckpt/pretrained/nusc_feats.pth` and users'
checkpoints load unchanged) -- but `forward` launches the sm_100a kernels of libhregnet_b200.so instead of
building the graph out of ATen ops.

Inference (eval mode, no gradients recorded): the fused kernels; BatchNorm uses its running statistics, folded into the
1x1 convolutions (fold.py).  Training mode, or gradients being recorded through the module: the differentiable forward of
train_path.py (this package's index kernels and differentiable gather / kNN ops, the module's own Conv + BatchNorm
containers under ATen's autograd, the Kabsch kernel with an analytic-by-autograd backward) -- same values, trainable.
"""
import torch
import torch.nn as nn

from . import engine, fold, train_path
from .engine import SEG_BROADCAST, SEG_GATHER, RowsView
from .ops import knn_points


def _pointwise_stack(dims, two_d=True):
    conv, bn = (nn.Conv2d, nn.BatchNorm2d) if two_d else (nn.Conv1d, nn.BatchNorm1d)
    mods = []
    for cin, cout in zip(dims[:-1], dims[1:]):
        mods += [conv(cin, cout, kernel_size=1, bias=False), bn(cout), nn.ReLU()]
    return nn.Sequential(*mods)


def _head(c):
    """Conv1d+BN+ReLU, Conv1d+BN+ReLU, Conv1d -> 1   (parameter containers named mlp1 / mlp2 / mlp3)."""
    m1 = nn.Sequential(nn.Conv1d(c, c, kernel_size=1), nn.BatchNorm1d(c), nn.ReLU())
    m2 = nn.Sequential(nn.Conv1d(c, c, kernel_size=1), nn.BatchNorm1d(c), nn.ReLU())
    m3 = nn.Sequential(nn.Conv1d(c, 1, kernel_size=1))
    return m1, m2, m3


class _Folded(nn.Module):
    """Caches BN-folded weights; re-folds when any parameter / buffer changed (version counters)."""

    def _fold(self):
        raise NotImplementedError

    def folded(self):
        if self.training:
            raise RuntimeError(f"{type(self).__name__}: the B200 path implements the inference forward only; call .eval()")
        sig = tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))
        if getattr(self, "_fold_sig", None) != sig:
            with torch.no_grad():
                self._fold_cache = self._fold()
            self._fold_sig = sig
        return self._fold_cache


def _cl(x):
    """[B,C,N] -> channels-last [B,N,C] through the transpose kernel."""
    return engine.transpose(x.contiguous())


class KeypointDetector(_Folded):
    """Reference: layers.py:89-165.  forward(xyz [B,N,3], features [B,C,N] | None, weights [B,N] | None) ->
    (keypoints [B,M,3], sigmas [B,M], attentive_feature [B,C_o,M], grouped_features [B,C+4,M,k],
    attentive_feature_map [B,C_o,M,k])."""

    def __init__(self, nsample, k, in_channels, out_channels, fps=True):
        super().__init__()
        self.nsample, self.k, self.fps = nsample, k, fps
        self.convs = _pointwise_stack([in_channels + 4, *out_channels])
        self.C_o1 = out_channels[-1]
        self.mlp1, self.mlp2, self.mlp3 = _head(self.C_o1)
        self.softplus = nn.Softplus()

    def _fold(self):
        return dict(convs=fold.fold_sequential(self.convs), mlp=fold.fold_head(self.mlp1, self.mlp2, self.mlp3))

    def forward(self, xyz, features, weights=None):
        if train_path.needs_autograd(self, xyz, features):
            return train_path.keypoint_detector(self, xyz, features, weights)
        B, N, _ = xyz.shape
        M, k = self.nsample, self.k
        feat_cl = _cl(features) if features is not None else None
        P = self.folded()
        # fps=False: one host permutation per call, shared by the clouds of the batch (layers.py:144-147)
        sample = None if self.fps else engine.random_sample_idx(N, M, [B], xyz.device)
        # detector-only use of the fused level: run the detector half of the stage
        lv = _detector_only(xyz.contiguous(), feat_cl, weights, P, M, k, sample)
        E, a, geom, idx = lv["E"], lv["a"], lv["geom"], lv["idx"]
        rows = B * M * k
        if feat_cl is not None:
            g = torch.empty(rows, feat_cl.shape[2], device=xyz.device)
            engine.call("hrn_gather_rows", engine.ptr(feat_cl), engine.ptr(idx.view(B, M * k)), engine.ptr(g), B, N, M * k,
                        feat_cl.shape[2], engine.stream())
            G = torch.cat([geom, g], dim=1)
        else:
            G = geom
        grouped = engine.transpose(G.view(B, M * k, -1)).view(B, -1, M, k)
        afm = engine.transpose((E * a[:, None]).view(B, M * k, -1)).view(B, -1, M, k)
        return lv["xyz"], lv["sigmas"], engine.transpose(lv["af"]), grouped, afm


def _detector_only(xyz, feat_cl, weights, P, M, k, sample_idx=None):
    B, N, _ = xyz.shape
    fidx = engine.fps(xyz, M, weights) if sample_idx is None else sample_idx
    idx, q = engine.knn_idx(None, xyz, k, q_idx=fidx)
    geom, nn_xyz = engine.group_geometry(q, xyz, idx, want_nn=True)
    v = RowsView(B * M * k, group=k, gather_idx=idx, rows_per_batch=M * k, src_rows_per_batch=N).add(geom)
    if feat_cl is not None:
        v.add(feat_cl.view(B * N, -1), SEG_GATHER)
    E = engine.stack(v, P["convs"])
    a = engine.group_attention(E, k)
    kp = engine.group_weighted_sum(a, nn_xyz, k)
    af = engine.group_weighted_sum(a, E, k)
    sig = engine.stack(RowsView(B * M).add(af), P["mlp"], last_act=engine.ACT_SOFTPLUS_EPS)
    return dict(xyz=kp.view(B, M, 3), sigmas=sig.view(B, M), af=af.view(B, M, -1), E=E, a=a, geom=geom, idx=idx)


class DescExtractor(_Folded):
    """Reference: layers.py:167-209.  forward(grouped_features [B,C+4,M,k], attentive_feature_map [B,C_d,M,k])
    -> desc [B,desc_dim,M]."""

    def __init__(self, in_channels, out_channels, C_detector, desc_dim):
        super().__init__()
        dims = [in_channels + 4, *out_channels]
        self.convs = _pointwise_stack(dims)
        self.C_o1 = dims[-1]
        self.mlp1 = _pointwise_stack([2 * self.C_o1 + C_detector, dims[-2]])
        self.mlp2 = _pointwise_stack([dims[-2], desc_dim])

    def _fold(self):
        return dict(convs=fold.fold_sequential(self.convs),
                    mlp=fold.fold_sequential(self.mlp1) + fold.fold_sequential(self.mlp2))

    def forward(self, grouped_features, attentive_feature_map):
        if train_path.needs_autograd(self, grouped_features, attentive_feature_map):
            return train_path.desc_extractor(self, grouped_features, attentive_feature_map)
        B, C, M, k = grouped_features.shape
        P = self.folded()
        G = engine.transpose(grouped_features.contiguous().view(B, C, M * k)).view(B * M * k, C)
        A = engine.transpose(attentive_feature_map.contiguous().view(B, -1, M * k)).view(B * M * k, -1)
        X1 = engine.stack(RowsView(B * M * k).add(G), P["convs"])
        X1max = engine.group_max(X1, k)
        v = RowsView(B * M * k, group=k).add(X1max, SEG_BROADCAST).add(X1).add(A)
        H = engine.stack(v, P["mlp"])
        d = engine.group_max(H, k)
        return engine.transpose(d.view(B, M, -1))


class CoarseReg(_Folded):
    """Reference: layers.py:211-396.  forward(src_xyz, src_desc [B,C,N], dst_xyz, dst_desc, src_weights, dst_weights)
    -> (corres_xyz [B,N,3], weights [B,N])."""

    def __init__(self, k, in_channels, use_sim=True, use_neighbor=True):
        super().__init__()
        self.k, self.use_sim, self.use_neighbor = k, use_sim, use_neighbor
        self.in_channels = in_channels
        c2 = in_channels * 2
        extra = 16 if (use_sim and use_neighbor) else (14 if (use_sim or use_neighbor) else 12)
        self.convs_1 = _pointwise_stack([c2 + extra, c2, c2, c2])
        self.convs_2 = _pointwise_stack([in_channels + 4, in_channels, in_channels, in_channels])
        self.mlp1, self.mlp2, self.mlp3 = _head(c2)

    def _fold(self):
        dev = self.convs_1[0].weight.device
        return dict(convs_1=fold.fold_sequential(self.convs_1, fold.pair_perm(self.in_channels, 4, dev)),
                    convs_2=fold.fold_sequential(self.convs_2),
                    mlp=fold.fold_head(self.mlp1, self.mlp2, self.mlp3))

    def forward(self, src_xyz, src_desc, dst_xyz, dst_desc, src_weights, dst_weights):
        if not (self.use_sim and self.use_neighbor):
            raise NotImplementedError("only use_sim=use_neighbor=True (the HRegNet configuration, models.py:71) is built")
        if train_path.needs_autograd(self, src_xyz, src_desc, dst_xyz, dst_desc, src_weights, dst_weights):
            return train_path.coarse_reg(self, src_xyz, src_desc, dst_xyz, dst_desc, src_weights, dst_weights)
        return engine.coarse_reg(src_xyz.contiguous(), _cl(src_desc), dst_xyz.contiguous(), _cl(dst_desc),
                                 src_weights.contiguous(), dst_weights.contiguous(), self.folded(), self.k)

    def forward_cl(self, sxyz, sdesc_cl, dxyz, ddesc_cl, ssig, dsig, both=None):
        return engine.coarse_reg(sxyz, sdesc_cl, dxyz, ddesc_cl, ssig, dsig, self.folded(), self.k, both=both)


class FineReg(_Folded):
    """Reference: layers.py:398-454."""

    def __init__(self, k, in_channels):
        super().__init__()
        self.k, self.in_channels = k, in_channels
        c2 = in_channels * 2
        self.convs_1 = _pointwise_stack([c2 + 12, c2, c2, c2])
        self.mlp1, self.mlp2, self.mlp3 = _head(c2)

    def _fold(self):
        dev = self.convs_1[0].weight.device
        return dict(convs_1=fold.fold_sequential(self.convs_1, fold.pair_perm(self.in_channels, 0, dev)),
                    mlp=fold.fold_head(self.mlp1, self.mlp2, self.mlp3))

    def forward(self, src_xyz, src_feat, dst_xyz, dst_feat, src_weights, dst_weights):
        if train_path.needs_autograd(self, src_xyz, src_feat, dst_xyz, dst_feat, src_weights, dst_weights):
            return train_path.fine_reg(self, src_xyz, src_feat, dst_xyz, dst_feat, src_weights, dst_weights)
        return engine.fine_reg(src_xyz.contiguous(), _cl(src_feat), dst_xyz.contiguous(), _cl(dst_feat),
                               src_weights.contiguous(), dst_weights.contiguous(), self.folded(), self.k)

    def forward_cl(self, sxyz, sfeat_cl, dxyz, dfeat_cl, ssig, dsig):
        return engine.fine_reg(sxyz, sfeat_cl, dxyz, dfeat_cl, ssig, dsig, self.folded(), self.k)


class WeightedSVDHead(nn.Module):
    """Reference: layers.py:456-504.  forward(src [B,N,3], src_corres [B,N,3], weights [B,N]) -> (r [B,3,3], t [B,3])."""

    def forward(self, src, src_corres, weights):
        if torch.is_grad_enabled() and any(t.requires_grad for t in (src, src_corres, weights)):
            return train_path.svd_head(src, src_corres, weights)
        return engine.weighted_kabsch(src.contiguous(), src_corres.contiguous(), weights.contiguous())


def knn_group(xyz1, xyz2, features2, k):
    """Reference: layers.py:9-27.  -> (grouped_features [B,4+C,M,k], knn_xyz [B,M,k,3])."""
    xyz1, xyz2 = xyz1.contiguous(), xyz2.contiguous()
    B, M, _ = xyz1.shape
    N = xyz2.shape[1]
    idx, _ = engine.knn_idx(xyz1, xyz2, k)
    geom, nn_xyz = engine.group_geometry(xyz1, xyz2, idx, want_nn=True)
    if features2 is not None:
        f_cl = _cl(features2)
        g = torch.empty(B * M * k, f_cl.shape[2], device=xyz1.device)
        engine.call("hrn_gather_rows", engine.ptr(f_cl), engine.ptr(idx.view(B, M * k)), engine.ptr(g), B, N, M * k,
                    f_cl.shape[2], engine.stream())
        geom = torch.cat([geom, g], dim=1)
    return engine.transpose(geom.view(B, M * k, -1)).view(B, -1, M, k), nn_xyz.view(B, M, k, 3)


def calc_cosine_similarity(desc1, desc2):
    """Reference: layers.py:29-41 (kept for API completeness; the model path uses hrn_cosine_matrix)."""
    inner = torch.sum(desc1 * desc2, dim=-1)
    return inner / (torch.norm(desc1, dim=-1) * torch.norm(desc2, dim=-1) + 1e-6)
