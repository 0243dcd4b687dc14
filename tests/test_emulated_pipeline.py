"""CPU: the product's host orchestration (engine.py / layers.py / models.py: virtual-rows segments, weight-column
permutations, BN folding, index plumbing, pose cascade) run over a torch emulation of the kernels (tests/emu.py)
and compared with the oracle / the golden reference outputs.  Kernels themselves are tested on the GPU."""
import pytest
import torch

from oracle import ref_layers as RL
from pcd_reg_hregnet_b200 import engine
import emu
from common import build_product_hregnet, load_golden, rel_err, unflatten


@pytest.fixture(scope="module")
def net():
    torch.set_num_threads(8)
    return build_product_hregnet(seed=7)


def test_feature_levels_teacher_forced(net):
    gd = load_golden("hregnet_b2_n2048")
    sd = net.state_dict()
    trace = {}
    with torch.no_grad():
        want = RL.hier_feature_extraction(sd, "feature_extraction.", gd["src"], trace=trace)
        fe = net.feature_extraction
        with emu.emulated_kernels():
            for lv in (1, 2, 3):
                det, desc = getattr(fe, f"detector_{lv}"), getattr(fe, f"desc_extractor_{lv}")
                feat = trace.get(f"af_{lv - 1}")
                feat_cl = feat.permute(0, 2, 1).contiguous() if feat is not None else None
                r = engine.detector_descriptor_level(trace[f"in_xyz_{lv}"].contiguous(), feat_cl, trace[f"in_w_{lv}"],
                                                     det.folded(), desc.folded(), det.nsample, det.k)
                assert rel_err(r["xyz"], want[f"xyz_{lv}"]) < 1e-5
                assert rel_err(r["sigmas"], want[f"sigmas_{lv}"]) < 1e-4
                assert rel_err(r["desc"].permute(0, 2, 1), want[f"desc_{lv}"]) < 1e-4
                assert rel_err(r["af"].permute(0, 2, 1), trace[f"af_{lv}"]) < 1e-4


def test_registration_stages_teacher_forced(net):
    gd = load_golden("hregnet_b2_n2048")
    S, D = unflatten(gd, "src_feats."), unflatten(gd, "dst_feats.")
    sd = net.state_dict()
    with torch.no_grad(), emu.emulated_kernels():
        cor, w = net.coarse_corres(S["xyz_3"], S["desc_3"], D["xyz_3"], D["desc_3"], S["sigmas_3"], D["sigmas_3"])
        assert float((cor - gd["src_xyz_corres_3"]).abs().max()) < 2e-4
        assert float((w - gd["src_dst_weights_3"]).abs().max()) < 1e-5
        R, t = net.svd_head(S["xyz_3"], gd["src_xyz_corres_3"], gd["src_dst_weights_3"])
        assert float(RL.rotation_angle_deg(R, gd["rotation.0"]).max()) < 1e-4
        assert float((t - gd["translation.0"]).abs().max()) < 1e-5
        x2 = engine.transform_points(S["xyz_2"].contiguous(), gd["rotation.0"], gd["translation.0"])
        cor2, w2 = net.fine_corres_2(x2, S["desc_2"], D["xyz_2"], D["desc_2"], S["sigmas_2"], D["sigmas_2"])
        assert float((cor2 - gd["src_xyz_corres_2"]).abs().max()) < 2e-4
        assert float((w2 - gd["src_dst_weights_2"]).abs().max()) < 1e-5


def test_full_forward_free_running(net):
    """Free-running end-to-end: weighted FPS amplifies 1e-7 feature differences into different keypoint sets
    (SURVEY.md section 7), so only pairs whose FPS indices all agree are held to the pose gate."""
    gd = load_golden("hregnet_uniform_b1_n1500")
    with torch.no_grad(), emu.emulated_kernels():
        out = net(gd["src"], gd["dst"])
    assert set(out.keys()) == {"src_xyz_corres_3", "src_xyz_corres_2", "src_xyz_corres_1", "src_dst_weights_3",
                               "src_dst_weights_2", "src_dst_weights_1", "rotation", "translation", "src_feats", "dst_feats"}
    assert out["src_feats"]["desc_3"].shape == (1, 256, 256) and out["rotation"][2].shape == (1, 3, 3)
    same = all(rel_err(out[s][f"xyz_{lv}"], gd[f"{s}.xyz_{lv}"]) < 1e-4 for s in ("src_feats", "dst_feats") for lv in (1, 2, 3))
    if same:
        assert float(RL.rotation_angle_deg(out["rotation"][2], gd["rotation.2"]).max()) < 1e-3
        assert float((out["translation"][2] - gd["translation.2"]).abs().max()) < 1e-4


def test_layer_level_api_shapes(net):
    g = torch.Generator().manual_seed(0)
    xyz = torch.rand(1, 1100, 3, generator=g)
    with torch.no_grad(), emu.emulated_kernels():
        det, desc = net.feature_extraction.detector_1, net.feature_extraction.desc_extractor_1
        kp, sig, af, grouped, afm = det(xyz, None)
        assert kp.shape == (1, 1024, 3) and sig.shape == (1, 1024) and af.shape == (1, 64, 1024)
        assert grouped.shape == (1, 4, 1024, 64) and afm.shape == (1, 64, 1024, 64)
        d = desc(grouped, afm)
        want = RL.hier_feature_extraction(net.state_dict(), "feature_extraction.", xyz, levels=((1, 1024, 64),))
        assert rel_err(d, want["desc_1"]) < 1e-4 and rel_err(kp, want["xyz_1"]) < 1e-5


def test_model_v4_coarse_stage_extra_outputs():
    """models/model_v4: coord_dist / feats_dist of the coarse stage through the product's host orchestration."""
    from common import build_product_model_v4
    net4 = build_product_model_v4(seed=7)
    gd = load_golden("hregnet_b2_n2048")
    S, D = unflatten(gd, "src_feats."), unflatten(gd, "dst_feats.")
    with torch.no_grad():
        want = RL.coarse_reg(net4.state_dict(), "coarse_corres.", S["xyz_3"], S["desc_3"], D["xyz_3"], D["desc_3"],
                             S["sigmas_3"], D["sigmas_3"], want_dists=True)
        with emu.emulated_kernels():
            got = net4.coarse_corres(S["xyz_3"], S["desc_3"], D["xyz_3"], D["desc_3"], S["sigmas_3"], D["sigmas_3"])
    assert len(got) == 4 and got[2].shape == (2, 256, 8) and got[3].shape == (2, 256, 8)
    assert float((got[2] - want[2]).abs().max()) < 1e-5 and float((got[3] - want[3]).abs().max()) < 1e-5
    assert float((got[0] - want[0]).abs().max()) < 2e-4 and float((got[1] - want[1]).abs().max()) < 1e-5
