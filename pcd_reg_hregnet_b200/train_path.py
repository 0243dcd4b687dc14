"""Differentiable forward of the drop-in modules ("make the new layers trainable behind the same API").

The fused tcgen05 kernels of engine.py are inference kernels: BatchNorm is folded, no activation is kept.  When a module is
in training mode, or gradients are being recorded through it, its forward runs here instead: the same arithmetic as the
reference's layers (models/HRegNet/layers.py, cited per function), built from

  * this package's index kernels -- FPS / weighted FPS / kNN (bit-exact, not differentiable by nature),
  * this package's differentiable ops -- `ops.gather_operation`, `ops.knn_gather`, `ops.knn_points` (their backward passes
    are the scatter-add kernels of csrc/gather.cu: hrn_gather_points_grad, hrn_knn_gather_grad, hrn_knn_dists_grad),
  * the modules' OWN parameter containers (`self.convs`, `self.mlp1`, ... : Conv + BatchNorm + ReLU): BatchNorm therefore
    uses batch statistics and updates its running statistics in training mode exactly like the reference, and every
    parameter receives its gradient through ATen's autograd (cuDNN / cuBLAS kernels, as in the reference),
  * the pose head: the weighted Kabsch KERNEL in the forward pass, and for the backward pass the same closed form
    re-evaluated under autograd (torch.linalg.svd) -- `WeightedKabsch`.

ATen's precision switches apply to this path as they do to the reference (cuDNN convolutions default to TF32 on this GPU;
`torch.backends.cudnn.allow_tf32 = False` for fp32).  Layout: channel-first tensors in and out, like the reference's modules.  Checked against the oracle under autograd
(tests/test_gpu_train.py): outputs equal to the fused inference path, gradients equal to the oracle's.
"""
import torch
import torch.nn.functional as F

from . import engine, ops


def needs_autograd(module, *tensors):
    """True when the forward has to be differentiable: the module trains, or gradients are recorded for a parameter / input."""
    if module.training:
        return True
    if not torch.is_grad_enabled():
        return False
    return any(p.requires_grad for p in module.parameters()) or any(t is not None and t.requires_grad for t in tensors)


def knn_group(xyz1, xyz2, features2, k):
    """layers.py:9-27 -> grouped [B,4+C,M,k], knn_xyz [B,M,k,3]; differentiable w.r.t. xyz1, xyz2, features2."""
    _, idx, nn = ops.knn_points(xyz1, xyz2, K=k, return_nn=True)
    rel = nn - xyz1.unsqueeze(2)
    parts = [rel, torch.norm(rel, dim=-1, keepdim=True)]
    if features2 is not None:
        parts.append(ops.knn_gather(features2.permute(0, 2, 1).contiguous(), idx))
    return torch.cat(parts, dim=-1).permute(0, 3, 1, 2).contiguous(), nn


def keypoint_detector(m, xyz, features, weights):
    """layers.py:134-165 on the module's own containers -> (keypoints, sigmas, attentive_feature, grouped, attentive_map)."""
    B, N, _ = xyz.shape
    xyz = xyz.contiguous()
    if m.fps:
        with torch.no_grad():
            idx = (ops.furthest_point_sample(xyz, m.nsample) if weights is None
                   else ops.weighted_furthest_point_sample(xyz, weights.detach().contiguous(), m.nsample))
        sampled = ops.gather_operation(xyz.permute(0, 2, 1).contiguous(), idx).permute(0, 2, 1).contiguous()
    else:
        sampled = xyz[:, torch.randperm(N)[:m.nsample].to(xyz.device), :].contiguous()      # layers.py:145-147
    grouped, nn = knn_group(sampled, xyz, features, m.k)
    emb = m.convs(grouped)                                              # [B,C_o,M,k]
    att = torch.softmax(emb.max(dim=1)[0], dim=-1)                      # [B,M,k]
    keypoints = (att.unsqueeze(-1) * nn).sum(dim=2)
    afm = emb * att.unsqueeze(1)
    af = afm.sum(dim=-1)
    sig = F.softplus(m.mlp3(m.mlp2(m.mlp1(af)))).squeeze(1) + 0.001
    return keypoints, sig, af, grouped, afm


def desc_extractor(m, grouped, afm):
    """layers.py:200-209."""
    x1 = m.convs(grouped)
    x2 = x1.max(dim=3, keepdim=True)[0].expand_as(x1)
    return m.mlp2(m.mlp1(torch.cat([x2, x1, afm], dim=1))).max(dim=3)[0]


def hier_feature_extraction(m, points):
    """models.py:26-58."""
    out, xyz, feat, w = {}, points, None, None
    for lv in (1, 2, 3):
        det, desc = getattr(m, f"detector_{lv}"), getattr(m, f"desc_extractor_{lv}")
        kp, sig, af, grouped, afm = keypoint_detector(det, xyz, feat, w)
        out[f"xyz_{lv}"], out[f"sigmas_{lv}"], out[f"desc_{lv}"] = kp, sig, desc_extractor(desc, grouped, afm)
        xyz, feat = kp, af
        if m.use_weights:
            w = 1.0 / (sig + 1e-5)
            w = w / w.mean(dim=1, keepdim=True)
        else:
            w = None
    return out


def _cos_features(S, D, idx):
    """layers.py:29-41, 292-313: normalised cosine similarities picked at the candidates -> (src_dst, dst_src) [B,N1,k]."""
    cosm = torch.einsum("bnc,bmc->bnm", D, S) / (D.norm(dim=-1)[:, :, None] * S.norm(dim=-1)[:, None, :] + 1e-6)
    A = cosm / (cosm.max(dim=2, keepdim=True)[0] + 1e-6)
    Bm = cosm / (cosm.max(dim=1, keepdim=True)[0] + 1e-6)
    b = torch.arange(idx.shape[0], device=idx.device)[:, None, None]
    i = torch.arange(idx.shape[1], device=idx.device)[None, :, None]
    return Bm[b, idx, i], A[b, idx, i]


def _nbr_desc(m, xyz, desc_cl, k):
    """layers.py:316-337: neighbourhood-attentive descriptors of one cloud (desc_cl [B,N,C])."""
    _, nidx, nxyz = ops.knn_points(xyz, xyz, K=k, return_nn=True)
    nfeat = ops.knn_gather(desc_cl, nidx)
    rel = nxyz - xyz.unsqueeze(2)
    f = torch.cat([nfeat, rel, torch.norm(rel, dim=-1, keepdim=True)], dim=-1).permute(0, 3, 1, 2)
    w = torch.softmax(m.convs_2(f).max(dim=1)[0], dim=-1)
    return (nfeat * w.unsqueeze(-1)).sum(dim=2)


def _pair_tail(m, feats, nbr_xyz):
    f = m.convs_1(feats.permute(0, 3, 1, 2))
    att = torch.softmax(f.max(dim=1)[0], dim=-1)
    cor = (att.unsqueeze(-1) * nbr_xyz).sum(dim=2)
    af = (att.unsqueeze(1) * f).sum(dim=-1)
    return cor, torch.sigmoid(m.mlp3(m.mlp2(m.mlp1(af))).squeeze(1)), af


def _pair_features(sxyz, Sf, dxyz, Df, sw, dw, idx, nbr_feat):
    k = idx.shape[2]
    nbr_xyz = ops.knn_gather(dxyz.contiguous(), idx)
    sx = sxyz.unsqueeze(2).expand(-1, -1, k, -1)
    rel = nbr_xyz - sx
    parts = [rel, torch.norm(rel, dim=-1, keepdim=True), sx, nbr_xyz, Sf.unsqueeze(2).expand(-1, -1, k, -1), nbr_feat,
             sw[:, :, None, None].expand(-1, -1, k, 1), ops.knn_gather(dw.unsqueeze(-1).contiguous(), idx)]
    return parts, nbr_xyz, rel


def coarse_reg(m, sxyz, sdesc, dxyz, ddesc, sw, dw, want_dists=False):
    """layers.py:273-396 (use_sim = use_neighbor = True)."""
    S, D = sdesc.permute(0, 2, 1).contiguous(), ddesc.permute(0, 2, 1).contiguous()
    _, idx, Dk = ops.knn_points(S, D, K=m.k, return_nn=True)              # descriptor-space candidates
    parts, nbr_xyz, rel = _pair_features(sxyz, S, dxyz, D, sw, dw, idx, Dk)
    sd_cos, ds_cos = _cos_features(S, D, idx)
    sd_n, ds_n = _cos_features(_nbr_desc(m, sxyz.contiguous(), S, m.k), _nbr_desc(m, dxyz.contiguous(), D, m.k), idx)
    feats = torch.cat(parts + [sd_cos.unsqueeze(-1), ds_cos.unsqueeze(-1), sd_n.unsqueeze(-1), ds_n.unsqueeze(-1)], dim=-1)
    cor, w, _ = _pair_tail(m, feats, nbr_xyz)
    if want_dists:                                                       # model_v4/layers.py:252,282
        return cor, w, torch.norm(rel, dim=-1).contiguous(), 1 - ds_cos
    return cor, w


def fine_reg(m, sxyz, sfeat, dxyz, dfeat, sw, dw, want_af=False):
    """layers.py:433-454."""
    sxyz = sxyz.contiguous()
    _, idx, _ = ops.knn_points(sxyz, dxyz.contiguous(), K=m.k)
    Sf, Df = sfeat.permute(0, 2, 1).contiguous(), dfeat.permute(0, 2, 1).contiguous()
    parts, nbr_xyz, _ = _pair_features(sxyz, Sf, dxyz, Df, sw, dw, idx, ops.knn_gather(Df, idx))
    cor, w, af = _pair_tail(m, torch.cat(parts, dim=-1), nbr_xyz)
    return (cor, w, af) if want_af else (cor, w)


class WeightedKabsch(torch.autograd.Function):
    """WeightedSVDHead (layers.py:469-504): forward = the fused covariance + Jacobi kernel (csrc/kabsch.cu); backward = the
    same closed form w' = w/(sum w + 1e-4), means / (sum w' + 1e-4), H = Xc^T diag(w') Yc, R = V diag(1,1,det) U^T,
    t = ybar - R xbar re-evaluated in fp64 under autograd (torch.linalg.svd supplies the SVD derivative)."""

    @staticmethod
    def forward(ctx, src, cor, w):
        ctx.save_for_backward(src, cor, w)
        return engine.weighted_kabsch(src.contiguous(), cor.contiguous(), w.contiguous())

    @staticmethod
    def backward(ctx, gR, gt):
        src, cor, w = ctx.saved_tensors
        with torch.enable_grad():
            s = src.detach().double().requires_grad_(True)
            c = cor.detach().double().requires_grad_(True)
            ww = w.detach().double().requires_grad_(True)
            R, t = kabsch_formula(s, c, ww)
            gs, gc, gw = torch.autograd.grad([R, t], [s, c, ww], [gR.double(), gt.double()], allow_unused=True)
        cast = lambda g, ref: None if g is None else g.to(ref.dtype)
        return cast(gs, src), cast(gc, cor), cast(gw, w)


def kabsch_formula(src, cor, weights, eps=1e-4):
    wn = (weights / (weights.sum(dim=1, keepdim=True) + eps)).unsqueeze(2)
    den = wn.sum(dim=1, keepdim=True) + eps
    sm = (wn * src).sum(dim=1, keepdim=True) / den
    cm = (wn * cor).sum(dim=1, keepdim=True) / den
    H = torch.matmul((src - sm).transpose(1, 2), wn * (cor - cm))
    U, _, Vh = torch.linalg.svd(H)
    V = Vh.transpose(1, 2)
    d = torch.det(torch.matmul(V, U.transpose(1, 2)))
    Dm = torch.diag_embed(torch.stack([torch.ones_like(d), torch.ones_like(d), d], dim=1))
    R = torch.matmul(V, torch.matmul(Dm, U.transpose(1, 2)))
    t = cm.transpose(1, 2) - torch.matmul(R, sm.transpose(1, 2))
    return R, t.squeeze(2)


def svd_head(src, cor, w):
    return WeightedKabsch.apply(src, cor, w)


def _compose(Ra, ta, Rb, tb):
    return torch.matmul(Ra, Rb), torch.matmul(Ra, tb.unsqueeze(2)).squeeze(2) + ta


def _apply(R, t, x):
    return (torch.matmul(R, x.permute(0, 2, 1)) + t.unsqueeze(2)).permute(0, 2, 1).contiguous()


def hregnet_forward(m, src, dst):
    """models.py:77-148 -- the extractor is called once per cloud like the reference (BatchNorm batch statistics are per call)."""
    S = hier_feature_extraction(m.feature_extraction, src)
    D = hier_feature_extraction(m.feature_extraction, dst)
    cor3, w3 = coarse_reg(m.coarse_corres, S["xyz_3"], S["desc_3"], D["xyz_3"], D["desc_3"], S["sigmas_3"], D["sigmas_3"])
    R3, t3 = svd_head(S["xyz_3"], cor3, w3)
    x2 = _apply(R3, t3, S["xyz_2"])
    cor2, w2 = fine_reg(m.fine_corres_2, x2, S["desc_2"], D["xyz_2"], D["desc_2"], S["sigmas_2"], D["sigmas_2"])
    R2_, t2_ = svd_head(x2, cor2, w2)
    R2, t2 = _compose(R2_, t2_, R3, t3)
    x1 = _apply(R2, t2, S["xyz_1"])
    cor1, w1 = fine_reg(m.fine_corres_1, x1, S["desc_1"], D["xyz_1"], D["desc_1"], S["sigmas_1"], D["sigmas_1"])
    R1_, t1_ = svd_head(x1, cor1, w1)
    R1, t1 = _compose(R1_, t1_, R2, t2)
    return {
        "src_xyz_corres_3": cor3, "src_xyz_corres_2": cor2, "src_xyz_corres_1": cor1,
        "src_dst_weights_3": w3, "src_dst_weights_2": w2, "src_dst_weights_1": w1,
        "rotation": [R3, R2, R1], "translation": [t3, t2, t1], "src_feats": S, "dst_feats": D,
    }
