"""Drop-in `HierFeatureExtraction` / `HRegNet` (reference models/HRegNet/models.py:7-148): identical constructor
(`args.use_fps, args.use_weights, args.freeze_detector, args.freeze_feats`), forward signature, returned
dictionary keys / shapes and state_dict keys.

B200-first differences in HOW the graph runs (not in what it computes):
  * the source and target clouds go through feature extraction as ONE batch of 2B clouds (the reference calls
    the extractor twice, models.py:79-80) -- twice the CTAs per launch for FPS / kNN / the shared MLPs;
  * features stay channels-last between stages; grouped tensors are never materialised (engine.py);
  * pose composition T2 = T2_ T3, T1 = T1_ T2 (models.py:100-127) is fused into the Kabsch kernel.
"""
import torch
import torch.nn as nn

from . import engine, train_path
from .layers import CoarseReg, DescExtractor, FineReg, KeypointDetector, WeightedSVDHead


class HierFeatureExtraction(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.use_fps = args.use_fps
        self.use_weights = args.use_weights
        widths = ([32, 32, 64], [64, 64, 128], [128, 128, 256])
        samples, ks, cin = (1024, 512, 256), (64, 32, 16), (0, 64, 128)
        for lv in range(3):
            setattr(self, f"detector_{lv + 1}",
                    KeypointDetector(nsample=samples[lv], k=ks[lv], in_channels=cin[lv], out_channels=widths[lv],
                                     fps=self.use_fps))
        if args.freeze_detector:
            for p in self.parameters():
                p.requires_grad = False
        for lv in range(3):
            setattr(self, f"desc_extractor_{lv + 1}",
                    DescExtractor(in_channels=cin[lv], out_channels=widths[lv], C_detector=widths[lv][-1],
                                  desc_dim=widths[lv][-1]))

    def forward_cl(self, points, calls=1):
        """Channels-last internal result: per level l: xyz_l [B,M,3], sigmas_l [B,M], desc_l [B,M,C].
        calls: how many reference calls of the extractor this batch stands for (HRegNet stacks the source and the target
        call, models.py:79-80, into one batch: calls=2).  Only the use_fps=False branch cares: it draws one host
        permutation per detector per call, in the reference's order (call 1: levels 1,2,3; then call 2)."""
        xyz, feat, w = points.contiguous(), None, None
        out = {}
        draws = None
        if not self.use_fps:
            per = points.shape[0] // calls
            n_in = [points.shape[1], self.detector_1.nsample, self.detector_2.nsample]
            host = [[torch.randperm(n_in[l])[:getattr(self, f"detector_{l + 1}").nsample].to(torch.int32) for l in range(3)]
                    for _ in range(calls)]
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("use_fps=False draws its samples on the host in every forward (layers.py:146); "
                                   "run it eagerly (Registrar(..., use_cuda_graph=False))")
            draws = [torch.cat([host[c][l][None].expand(per, -1) for c in range(calls)], 0).contiguous().to(points.device)
                     for l in range(3)]
        for lv in (1, 2, 3):
            det, desc = getattr(self, f"detector_{lv}"), getattr(self, f"desc_extractor_{lv}")
            r = engine.detector_descriptor_level(xyz, feat, w, det.folded(), desc.folded(), det.nsample, det.k,
                                                 sample_idx=None if draws is None else draws[lv - 1])
            out[f"xyz_{lv}"], out[f"sigmas_{lv}"], out[f"desc_{lv}"] = r["xyz"], r["sigmas"], r["desc"]
            xyz, feat = r["xyz"], r["af"]
            w = engine.sigma_to_weights(r["sigmas"]) if self.use_weights and lv < 3 else None     # sampling weights of the next level
        return out

    def forward(self, points):
        if train_path.needs_autograd(self, points):
            return train_path.hier_feature_extraction(self, points)
        cl = self.forward_cl(points)
        return {k: (engine.transpose(v) if k.startswith("desc_") else v) for k, v in cl.items()}


class HRegNet(nn.Module):
    def __init__(self, args):
        super().__init__()
        self.feature_extraction = HierFeatureExtraction(args)
        if args.freeze_feats:
            for p in self.parameters():
                p.requires_grad = False
        self.coarse_corres = CoarseReg(k=8, in_channels=256, use_sim=True, use_neighbor=True)
        self.fine_corres_2 = FineReg(k=8, in_channels=128)
        self.fine_corres_1 = FineReg(k=8, in_channels=64)
        self.svd_head = WeightedSVDHead()

    def forward(self, src_points, dst_points):
        if train_path.needs_autograd(self, src_points, dst_points):
            return train_path.hregnet_forward(self, src_points, dst_points)     # training: differentiable path
        B = src_points.shape[0]
        both = self.feature_extraction.forward_cl(engine.stack_clouds(src_points, dst_points), calls=2)
        S = {k: v[:B] for k, v in both.items()}
        D = {k: v[B:] for k, v in both.items()}
        # API layout [B,C,N] of the descriptors: one transpose per level over the stacked source + target batch, on the
        # side stream beside the registration stages (they only read the channels-last tensors)
        api_desc = engine.on_side_stream(lambda: {k: engine.transpose(v) for k, v in both.items() if k.startswith("desc_")},
                                         both["desc_1"])

        cor3, w3 = self.coarse_corres.forward_cl(S["xyz_3"], S["desc_3"], D["xyz_3"], D["desc_3"], S["sigmas_3"],
                                                 D["sigmas_3"], both=(both["xyz_3"], both["desc_3"]))
        R3, t3 = engine.weighted_kabsch(S["xyz_3"], cor3, w3)

        xyz2_t = engine.transform_points(S["xyz_2"], R3, t3)
        cor2, w2 = self.fine_corres_2.forward_cl(xyz2_t, S["desc_2"], D["xyz_2"], D["desc_2"], S["sigmas_2"],
                                                 D["sigmas_2"])
        _, _, R2, t2 = engine.weighted_kabsch(xyz2_t, cor2, w2, prev=(R3, t3))

        xyz1_t = engine.transform_points(S["xyz_1"], R2, t2)
        cor1, w1 = self.fine_corres_1.forward_cl(xyz1_t, S["desc_1"], D["xyz_1"], D["desc_1"], S["sigmas_1"],
                                                 D["sigmas_1"])
        _, _, R1, t1 = engine.weighted_kabsch(xyz1_t, cor1, w1, prev=(R2, t2), packed=True)
        api_desc = api_desc()                                                  # join the side stream

        def api(d, lo, hi):
            return {k: (api_desc[k][lo:hi] if k.startswith("desc_") else v) for k, v in d.items()}

        return {
            "src_xyz_corres_3": cor3, "src_xyz_corres_2": cor2, "src_xyz_corres_1": cor1,
            "src_dst_weights_3": w3, "src_dst_weights_2": w2, "src_dst_weights_1": w1,
            "rotation": [R3, R2, R1], "translation": [t3, t2, t1],
            "src_feats": api(S, 0, B), "dst_feats": api(D, B, 2 * B),
        }
