"""ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU (torch, fp32) restatement of the reference's Python layers for the
registration forward path, in functional form on a reference-keyed state_dict.  Nothing under
pcd_reg_hregnet_b200/ may import this module.

Each function cites the reference lines it follows (paths relative to /root/reference):
    knn_group               models/HRegNet/layers.py:9-27
    keypoint_detector       models/HRegNet/layers.py:134-165
    desc_extractor          models/HRegNet/layers.py:200-209
    hier_feature_extraction models/HRegNet/models.py:26-58
    coarse_reg              models/HRegNet/layers.py:273-396
    fine_reg                models/HRegNet/layers.py:433-454
    weighted_svd_head       models/HRegNet/layers.py:469-504
    hregnet_forward         models/HRegNet/models.py:77-148
    fine_reg2               models/model_v2/layers.py:464-501
    model_v2_forward        models/model_v2/models.py:77-183
    coarse_reg(want_dists)  models/model_v4/layers.py:237-369 (coord_dist :252, feats_dist :282)
    regression_head         models/model_v2/layers.py:625-668
    model_v4_forward        models/model_v4/models.py:77-183

Native ops underneath (FPS, gather, kNN) are oracle/native.py.  Pinning: tests/test_oracle_vs_reference.py runs
these functions against the UNMODIFIED reference modules (oracle/ref_harness.py) on seeded inputs in the build
container, and tests/golden/*.npz stores reference outputs for the GPU box, where /root/reference is absent.

Eval-mode semantics only (BatchNorm uses running statistics), as in the reference's test scripts
(test/test_v3.py:95).
"""
import torch
import torch.nn.functional as F

from . import native

BN_EPS = 1e-5


def _conv(x, w):
    """1x1 convolution without bias (same ATen op as the reference's nn.Conv2d / nn.Conv1d)."""
    return F.conv2d(x, w) if x.dim() == 4 else F.conv1d(x, w)


def _bn(x, sd, p):
    """Eval-mode BatchNorm through the same ATen op as the reference's nn.BatchNorm1d/2d."""
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.1, BN_EPS)


def _conv_stack(x, sd, p, n=3):
    """n x [Conv(1x1, bias=False), BN, ReLU] stored at indices 0,1 / 3,4 / 6,7 of an nn.Sequential."""
    for i in range(n):
        x = torch.relu(_bn(_conv(x, sd[f"{p}.{3 * i}.weight"]), sd, f"{p}.{3 * i + 1}"))
    return x


def _conv1d_bn_relu(x, sd, p):
    return torch.relu(_bn(F.conv1d(x, sd[p + ".0.weight"], sd[p + ".0.bias"]), sd, p + ".1"))


def _head(x, sd, p):
    """mlp1, mlp2 (Conv1d+BN+ReLU), mlp3 (Conv1d -> 1): x [B,C,N] -> [B,N]."""
    x = _conv1d_bn_relu(_conv1d_bn_relu(x, sd, p + "mlp1"), sd, p + "mlp2")
    return F.conv1d(x, sd[p + "mlp3.0.weight"], sd[p + "mlp3.0.bias"]).squeeze(1)


def knn_group(xyz1, xyz2, features2, k):
    _, idx, nn = native.knn_points(xyz1, xyz2, K=k, return_nn=True)
    rel = nn - xyz1.unsqueeze(2)
    dist = torch.norm(rel, dim=-1, keepdim=True)
    parts = [rel, dist]
    if features2 is not None:
        parts.append(native.knn_gather(features2.permute(0, 2, 1).contiguous(), idx))
    return torch.cat(parts, dim=-1).permute(0, 3, 1, 2).contiguous(), nn


def keypoint_detector(sd, p, xyz, features, weights, nsample, k, sample_idx=None):
    """-> keypoints [B,M,3], sigmas [B,M], attentive_feature [B,C,M], grouped [B,4+C,M,k], afm [B,C,M,k], fps_idx
    sample_idx [B,M]: the fps=False branch (layers.py:144-147) -- the caller draws torch.randperm(N)[:nsample]."""
    fps_idx = native.fps(xyz, nsample, weights) if sample_idx is None else sample_idx
    B = xyz.shape[0]
    sampled = xyz[torch.arange(B)[:, None], fps_idx.long()]
    grouped, nn = knn_group(sampled, xyz, features, k)
    emb = _conv_stack(grouped, sd, p + "convs")
    att = torch.softmax(emb.max(dim=1)[0], dim=-1)                      # [B,M,k]
    keypoints = (att.unsqueeze(-1) * nn).sum(dim=2)
    afm = emb * att.unsqueeze(1)
    af = afm.sum(dim=-1)
    sig = F.softplus(_head(af, sd, p)) + 0.001
    return keypoints, sig, af, grouped, afm, fps_idx


def desc_extractor(sd, p, grouped, afm):
    x1 = _conv_stack(grouped, sd, p + "convs")
    x2 = x1.max(dim=3, keepdim=True)[0].expand_as(x1)
    x = torch.cat([x2, x1, afm], dim=1)
    x = torch.relu(_bn(_conv(x, sd[p + "mlp1.0.weight"]), sd, p + "mlp1.1"))
    x = torch.relu(_bn(_conv(x, sd[p + "mlp2.0.weight"]), sd, p + "mlp2.1"))
    return x.max(dim=3)[0]


LEVELS = ((1, 1024, 64), (2, 512, 32), (3, 256, 16))


def sigma_weights(sig):
    w = 1.0 / (sig + 1e-5)
    return w / w.mean(dim=1, keepdim=True)


def hier_feature_extraction(sd, p, points, use_weights=True, levels=LEVELS, trace=None):
    out, xyz, feat, w = {}, points, None, None
    for lv, nsample, k in levels:
        kp, sig, af, grouped, afm, fidx = keypoint_detector(sd, f"{p}detector_{lv}.", xyz, feat, w, nsample, k)
        out[f"xyz_{lv}"], out[f"sigmas_{lv}"] = kp, sig
        out[f"desc_{lv}"] = desc_extractor(sd, f"{p}desc_extractor_{lv}.", grouped, afm)
        if trace is not None:
            trace[f"fps_idx_{lv}"], trace[f"af_{lv}"], trace[f"in_xyz_{lv}"] = fidx, af, xyz
            trace[f"in_w_{lv}"] = w
        xyz, feat = kp, af
        w = sigma_weights(sig) if use_weights else None
    return out


def _cos_features(S, D, idx):
    """S [B,N1,C], D [B,N2,C], idx [B,N1,k] -> (src_dst_cos, dst_src_cos) [B,N1,k]   (layers.py:29-41,292-313)"""
    inner = torch.einsum("bnc,bmc->bnm", D, S)                           # [B,N2,N1]
    cosm = inner / (D.norm(dim=-1)[:, :, None] * S.norm(dim=-1)[:, None, :] + 1e-6)
    A = cosm / (cosm.max(dim=2, keepdim=True)[0] + 1e-6)                # dst_src_cos_norm [B,N2,N1]
    Bm = cosm / (cosm.max(dim=1, keepdim=True)[0] + 1e-6)               # src_dst_cos_norm^T  [B,N2,N1]
    B_, N1, k = idx.shape
    b = torch.arange(B_)[:, None, None]
    i = torch.arange(N1)[None, :, None]
    return Bm[b, idx, i], A[b, idx, i]


def _nbr_desc(sd, p, xyz, desc, k):
    _, nidx, nxyz = native.knn_points(xyz, xyz, K=k, return_nn=True)
    nfeat = native.knn_gather(desc, nidx)
    rel = nxyz - xyz.unsqueeze(2)
    f = torch.cat([nfeat, rel, torch.norm(rel, dim=-1, keepdim=True)], dim=-1).permute(0, 3, 1, 2)
    w = torch.softmax(_conv_stack(f, sd, p + "convs_2").max(dim=1)[0], dim=-1)
    return (nfeat * w.unsqueeze(-1)).sum(dim=2)


def _pair_tail(sd, p, feats, nbr_xyz):
    f = _conv_stack(feats.permute(0, 3, 1, 2), sd, p + "convs_1")        # [B,C,N,k]
    att = torch.softmax(f.max(dim=1)[0], dim=-1)
    cor = (att.unsqueeze(-1) * nbr_xyz).sum(dim=2)
    af = (att.unsqueeze(1) * f).sum(dim=-1)
    return cor, torch.sigmoid(_head(af, sd, p)), af


def coarse_reg(sd, p, sxyz, sdesc, dxyz, ddesc, sw, dw, k=8, trace=None, want_dists=False):
    S, D = sdesc.permute(0, 2, 1).contiguous(), ddesc.permute(0, 2, 1).contiguous()
    _, idx, Dk = native.knn_points(S, D, K=k, return_nn=True)
    nbr_xyz = native.knn_gather(dxyz, idx)
    sx = sxyz.unsqueeze(2).expand(-1, -1, k, -1)
    rel = nbr_xyz - sx
    sd_cos, ds_cos = _cos_features(S, D, idx)
    s_n, d_n = _nbr_desc(sd, p, sxyz, S, k), _nbr_desc(sd, p, dxyz, D, k)
    sd_ncos, ds_ncos = _cos_features(s_n, d_n, idx)
    feats = torch.cat([rel, torch.norm(rel, dim=-1, keepdim=True), sx, nbr_xyz,
                       S.unsqueeze(2).expand(-1, -1, k, -1), Dk,
                       sw[:, :, None, None].expand(-1, -1, k, 1), native.knn_gather(dw.unsqueeze(-1), idx),
                       sd_cos.unsqueeze(-1), ds_cos.unsqueeze(-1), sd_ncos.unsqueeze(-1), ds_ncos.unsqueeze(-1)], dim=-1)
    if trace is not None:
        trace.update(coarse_idx=idx, coarse_feats=feats, src_nbr_desc=s_n, dst_nbr_desc=d_n)
    cor, w, _ = _pair_tail(sd, p, feats, nbr_xyz)
    if want_dists:                                                     # model_v4/layers.py:252,282
        return cor, w, torch.norm(rel, dim=-1).contiguous(), 1 - ds_cos
    return cor, w


def fine_reg(sd, p, sxyz, sfeat, dxyz, dfeat, sw, dw, k=8, trace=None, name=""):
    _, idx, nbr_xyz = native.knn_points(sxyz, dxyz, K=k, return_nn=True)
    Sf, Df = sfeat.permute(0, 2, 1).contiguous(), dfeat.permute(0, 2, 1).contiguous()
    sx = sxyz.unsqueeze(2).expand(-1, -1, k, -1)
    rel = nbr_xyz - sx
    feats = torch.cat([rel, torch.norm(rel, dim=-1, keepdim=True), sx, nbr_xyz,
                       Sf.unsqueeze(2).expand(-1, -1, k, -1), native.knn_gather(Df, idx),
                       sw[:, :, None, None].expand(-1, -1, k, 1), native.knn_gather(dw.unsqueeze(-1), idx)], dim=-1)
    if trace is not None:
        trace[name + "idx"] = idx
    cor, w, af = _pair_tail(sd, p, feats, nbr_xyz)
    return cor, w


def fine_reg2(sd, p, sxyz, sfeat, dxyz, dfeat, sw, dw, k=8):
    """models/model_v2/layers.py:464-501 (FineReg2): FineReg + mlpx + two host-RNG batch shuffles, in that order."""
    _, idx, nbr_xyz = native.knn_points(sxyz, dxyz, K=k, return_nn=True)
    Sf, Df = sfeat.permute(0, 2, 1).contiguous(), dfeat.permute(0, 2, 1).contiguous()
    sx = sxyz.unsqueeze(2).expand(-1, -1, k, -1)
    rel = nbr_xyz - sx
    feats = torch.cat([rel, torch.norm(rel, dim=-1, keepdim=True), sx, nbr_xyz,
                       Sf.unsqueeze(2).expand(-1, -1, k, -1), native.knn_gather(Df, idx),
                       sw[:, :, None, None].expand(-1, -1, k, 1), native.knn_gather(dw.unsqueeze(-1), idx)], dim=-1)
    cor, w, af = _pair_tail(sd, p, feats, nbr_xyz)
    fx = _conv1d_bn_relu(af, sd, p + "mlpx")
    fx_prime = fx[torch.randperm(fx.size(0))]
    w_prime = w[torch.randperm(w.size(0))]
    return cor, w, w_prime, fx, fx_prime


def model_v2_forward(sd, src, dst):
    """models/model_v2/models.py:77-183."""
    fe = "feature_extraction."
    S = hier_feature_extraction(sd, fe, src)
    D = hier_feature_extraction(sd, fe, dst)
    cor3, w3 = coarse_reg(sd, "coarse_corres.", S["xyz_3"], S["desc_3"], D["xyz_3"], D["desc_3"], S["sigmas_3"], D["sigmas_3"])
    R3, t3 = weighted_svd_head(S["xyz_3"], cor3, w3)
    x2 = _apply(R3, t3, S["xyz_2"])
    cor2, w2, w2p, f2, f2p = fine_reg2(sd, "fine_corres_2.", x2, S["desc_2"], D["xyz_2"], D["desc_2"], S["sigmas_2"], D["sigmas_2"])
    R2_, t2_ = weighted_svd_head(x2, cor2, w2)
    R2, t2 = _compose(R2_, t2_, R3, t3)
    x1 = _apply(R2, t2, S["xyz_1"])
    cor1, w1 = fine_reg(sd, "fine_corres_1.", x1, S["desc_1"], D["xyz_1"], D["desc_1"], S["sigmas_1"], D["sigmas_1"])
    R1_, t1_ = weighted_svd_head(x1, cor1, w1)
    R1, t1 = _compose(R1_, t1_, R2, t2)
    return {
        "src_xyz_corres_3": cor3, "src_xyz_corres_2": cor2, "src_xyz_corres_1": cor1,
        "rotation": [R3, R2, R1], "translation": [t3, t2, t1],
        "src_feats_desc_2": S["desc_2"], "src_feats_sigmas_2": S["sigmas_2"], "src_xyz_2_trans": x2, "dst_xyz_2": D["xyz_2"],
        "src_dst_feats_2": f2, "src_dst_feats_2_prime": f2p, "src_dst_weights_2": w2, "src_dst_weights_2_prime": w2p,
        "src_feats": S, "dst_feats": D,
    }


def model_v4_forward(sd, src, dst):
    """models/model_v4/models.py:77-183: Model_V2's cascade, the coarse stage also returning coord_dist / feats_dist."""
    fe = "feature_extraction."
    S = hier_feature_extraction(sd, fe, src)
    D = hier_feature_extraction(sd, fe, dst)
    cor3, w3, coord_dist, feats_dist = coarse_reg(sd, "coarse_corres.", S["xyz_3"], S["desc_3"], D["xyz_3"], D["desc_3"],
                                                  S["sigmas_3"], D["sigmas_3"], want_dists=True)
    R3, t3 = weighted_svd_head(S["xyz_3"], cor3, w3)
    x2 = _apply(R3, t3, S["xyz_2"])
    cor2, w2, w2p, f2, f2p = fine_reg2(sd, "fine_corres_2.", x2, S["desc_2"], D["xyz_2"], D["desc_2"], S["sigmas_2"], D["sigmas_2"])
    R2_, t2_ = weighted_svd_head(x2, cor2, w2)
    R2, t2 = _compose(R2_, t2_, R3, t3)
    x1 = _apply(R2, t2, S["xyz_1"])
    cor1, w1 = fine_reg(sd, "fine_corres_1.", x1, S["desc_1"], D["xyz_1"], D["desc_1"], S["sigmas_1"], D["sigmas_1"])
    R1_, t1_ = weighted_svd_head(x1, cor1, w1)
    R1, t1 = _compose(R1_, t1_, R2, t2)
    return {
        "rotation": [R3, R2, R1], "translation": [t3, t2, t1],
        "src_feats_desc_2": S["desc_2"], "src_feats_sigmas_2": S["sigmas_2"], "src_xyz_2_trans": x2, "dst_xyz_2": D["xyz_2"],
        "src_dst_feats_2": f2, "src_dst_feats_2_prime": f2p, "src_dst_weights_2": w2, "src_dst_weights_2_prime": w2p,
        "coord_dist": coord_dist, "feats_dist": feats_dist,
        "_stage_inputs": {"S": S, "D": D},                             # (not a reference key: teacher-forcing in the tests)
    }


def weighted_svd_head(src, cor, weights, dtype=torch.float32):
    """dtype=float64 gives the 'exact' answer used to put |new - truth| next to |reference - truth|."""
    src, cor, weights = src.to(dtype), cor.to(dtype), weights.to(dtype)
    eps = 1e-4
    w = (weights / (weights.sum(dim=1, keepdim=True) + eps)).unsqueeze(2)
    den = w.sum(dim=1).unsqueeze(1) + eps
    sm = torch.matmul(w.transpose(1, 2), src) / den
    cm = torch.matmul(w.transpose(1, 2), cor) / den
    H = torch.matmul((src - sm).transpose(1, 2), w * (cor - cm))
    U, _, Vh = torch.linalg.svd(H)
    V = Vh.transpose(1, 2)
    det = torch.det(torch.matmul(V.transpose(1, 2), U.transpose(1, 2)))
    Dm = torch.diag_embed(torch.stack([torch.ones_like(det), torch.ones_like(det), det], dim=1))
    R = torch.matmul(V, torch.matmul(Dm, U.transpose(1, 2)))
    t = cm.transpose(1, 2) - torch.matmul(R, sm.transpose(1, 2))
    return R, t.view(-1, 3)


def regression_head(sd, p, src, cor, weights):
    """RegressionHead.forward (models/model_v2/layers.py:636-668) on a state_dict with prefix p."""
    w = (weights / (weights.sum(dim=1, keepdim=True) + 1e-4)).unsqueeze(2)
    x = torch.cat([(w * src).sum(dim=1), (w * cor).sum(dim=1)], dim=1)
    out = []
    for br in ("rot", "trans"):
        h = torch.relu(F.linear(x, sd[f"{p}fc1_{br}.weight"], sd[f"{p}fc1_{br}.bias"]))
        h = torch.relu(F.linear(h, sd[f"{p}fc2_{br}.weight"], sd[f"{p}fc2_{br}.bias"]))
        out.append(F.linear(h, sd[f"{p}fc3_{br}.weight"], sd[f"{p}fc3_{br}.bias"]))
    return out[0], out[1]


def _compose(Ra, ta, Rb, tb):
    """[Ra|ta] * [Rb|tb]   (models.py:100-110)"""
    return torch.matmul(Ra, Rb), torch.matmul(Ra, tb.unsqueeze(2)).squeeze(2) + ta


def _apply(R, t, x):
    return (torch.matmul(R, x.permute(0, 2, 1)) + t.unsqueeze(2)).permute(0, 2, 1).contiguous()


def hregnet_forward(sd, src, dst, trace=None):
    fe = "feature_extraction."
    tr_s = {} if trace is not None else None
    tr_d = {} if trace is not None else None
    S = hier_feature_extraction(sd, fe, src, trace=tr_s)
    D = hier_feature_extraction(sd, fe, dst, trace=tr_d)
    cor3, w3 = coarse_reg(sd, "coarse_corres.", S["xyz_3"], S["desc_3"], D["xyz_3"], D["desc_3"], S["sigmas_3"],
                          D["sigmas_3"], trace=trace)
    R3, t3 = weighted_svd_head(S["xyz_3"], cor3, w3)
    x2 = _apply(R3, t3, S["xyz_2"])
    cor2, w2 = fine_reg(sd, "fine_corres_2.", x2, S["desc_2"], D["xyz_2"], D["desc_2"], S["sigmas_2"], D["sigmas_2"],
                        trace=trace, name="fine2_")
    R2_, t2_ = weighted_svd_head(x2, cor2, w2)
    R2, t2 = _compose(R2_, t2_, R3, t3)
    x1 = _apply(R2, t2, S["xyz_1"])
    cor1, w1 = fine_reg(sd, "fine_corres_1.", x1, S["desc_1"], D["xyz_1"], D["desc_1"], S["sigmas_1"], D["sigmas_1"],
                        trace=trace, name="fine1_")
    R1_, t1_ = weighted_svd_head(x1, cor1, w1)
    R1, t1 = _compose(R1_, t1_, R2, t2)
    if trace is not None:
        trace.update(src_trace=tr_s, dst_trace=tr_d, src_xyz_2_trans=x2, src_xyz_1_trans=x1)
    return {
        "src_xyz_corres_3": cor3, "src_xyz_corres_2": cor2, "src_xyz_corres_1": cor1,
        "src_dst_weights_3": w3, "src_dst_weights_2": w2, "src_dst_weights_1": w1,
        "rotation": [R3, R2, R1], "translation": [t3, t2, t1], "src_feats": S, "dst_feats": D,
    }


def rotation_angle_deg(Ra, Rb):
    """Geodesic angle between rotations in fp64, well conditioned near 0:
    theta = atan2(|vee(E - E^T)|/2, (tr E - 1)/2), E = Ra^T Rb.  (metrics/calibeval.py:172-196 uses acos.)"""
    E = torch.matmul(Ra.double().transpose(-1, -2), Rb.double())
    v = torch.stack([E[..., 2, 1] - E[..., 1, 2], E[..., 0, 2] - E[..., 2, 0], E[..., 1, 0] - E[..., 0, 1]], -1)
    s = 0.5 * v.norm(dim=-1)
    c = 0.5 * (E[..., 0, 0] + E[..., 1, 1] + E[..., 2, 2] - 1.0)
    return torch.rad2deg(torch.atan2(s, c))
