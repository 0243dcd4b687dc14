"""CPU, build container only: pins the oracle restatement (oracle/ref_layers.py) and the golden fixtures to the
UNMODIFIED reference Python graph (/root/reference/models/HRegNet) run through oracle/ref_harness.py."""
import numpy as np
import pytest
import torch

from oracle import ref_harness as H, ref_layers as RL
from common import Args, build_product_hregnet, load_golden
from pcd_reg_hregnet_b200 import synth

pytestmark = pytest.mark.skipif(not H.available(), reason="/root/reference not present (GPU box)")


@pytest.fixture(scope="module")
def ref_net():
    torch.set_num_threads(8)
    return H.build_reference_hregnet(seed=7)


def test_feature_extraction_bit_identical(ref_net):
    src, _, _, _ = synth.make_batch([11], 3000)
    with torch.no_grad():
        a = ref_net.feature_extraction(src)
        b = RL.hier_feature_extraction(ref_net.state_dict(), "feature_extraction.", src)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_full_forward_matches_reference(ref_net):
    src, dst, _, _ = synth.make_batch([21, 22], 2048)
    with torch.no_grad():
        a = ref_net(src, dst)
        b = RL.hregnet_forward(ref_net.state_dict(), src, dst)
    for lv in range(3):
        assert float(RL.rotation_angle_deg(a["rotation"][lv], b["rotation"][lv]).max()) < 1e-4
        assert float((a["translation"][lv] - b["translation"][lv]).abs().max()) < 1e-5
    for k in ("src_xyz_corres_3", "src_xyz_corres_2", "src_xyz_corres_1"):
        assert float((a[k] - b[k]).abs().max()) < 1e-4
    for k in ("src_dst_weights_3", "src_dst_weights_2", "src_dst_weights_1"):
        assert float((a[k] - b[k]).abs().max()) < 1e-6


def test_layers_stagewise(ref_net):
    L = H.load_reference().layers
    sd = ref_net.state_dict()
    g = torch.Generator().manual_seed(0)
    sx, dx = torch.randn(2, 256, 3, generator=g) * 20, torch.randn(2, 256, 3, generator=g) * 20
    sdsc, ddsc = torch.rand(2, 256, 256, generator=g), torch.rand(2, 256, 256, generator=g)
    sw, dw = torch.rand(2, 256, generator=g) + 0.1, torch.rand(2, 256, generator=g) + 0.1
    with torch.no_grad():
        c0, w0 = ref_net.coarse_corres(sx, sdsc, dx, ddsc, sw, dw)
        c1, w1 = RL.coarse_reg(sd, "coarse_corres.", sx, sdsc, dx, ddsc, sw, dw)
        assert float((c0 - c1).abs().max()) < 1e-4 and float((w0 - w1).abs().max()) < 1e-6
        f0 = ref_net.fine_corres_2(sx, sdsc[:, :128], dx, ddsc[:, :128], sw, dw)
        f1 = RL.fine_reg(sd, "fine_corres_2.", sx, sdsc[:, :128], dx, ddsc[:, :128], sw, dw)
        assert float((f0[0] - f1[0]).abs().max()) < 1e-4 and float((f0[1] - f1[1]).abs().max()) < 1e-6
        r0, t0 = ref_net.svd_head(sx, c0, w0)
        r1, t1 = RL.weighted_svd_head(sx, c0, w0)
        assert float(RL.rotation_angle_deg(r0, r1).max()) < 1e-4 and float((t0 - t1).abs().max()) < 1e-5


def test_product_state_dict_and_seeded_init_equal_reference(ref_net):
    prod = build_product_hregnet(seed=7)
    rs, ps = ref_net.state_dict(), prod.state_dict()
    assert list(rs.keys()) == list(ps.keys())
    for k in rs:
        assert rs[k].shape == ps[k].shape and torch.equal(rs[k], ps[k]), k
    # Model-level API: same constructor contract (args.use_fps, use_weights, freeze_*), models.py:62-75
    assert sum(p.numel() for p in prod.parameters()) == 2467846


def test_golden_fixtures_are_reference_outputs(ref_net):
    for name in ("hregnet_b2_n2048", "hregnet_uniform_b1_n1500"):
        gd = load_golden(name)
        with torch.no_grad():
            out = ref_net(gd["src"], gd["dst"])
        for i in range(3):
            assert torch.equal(out["rotation"][i], gd[f"rotation.{i}"])
            assert torch.equal(out["translation"][i], gd[f"translation.{i}"])
        assert torch.equal(out["src_feats"]["desc_3"], gd["src_feats.desc_3"])


def test_model_v2_matches_reference():
    """Adaption-1 (models/model_v2): state_dict keys, seeded init and forward (incl. the host-RNG shuffles)."""
    from common import build_product_model_v2
    ns = H.load_reference()
    assert ns.Model_V2 is not None, getattr(ns, "v2_error", None)
    torch.manual_seed(7)
    ref = ns.Model_V2(Args())
    ref.feature_extraction.load_state_dict(torch.load(H.pretrained_feats_path(), map_location="cpu"))
    g = torch.Generator().manual_seed(8)
    for name in ("coarse_corres", "fine_corres_2", "fine_corres_1"):
        H.randomize_bn_(getattr(ref, name), g)
    ref.eval()
    prod = build_product_model_v2(seed=7)
    rs, ps = ref.state_dict(), prod.state_dict()
    assert list(rs.keys()) == list(ps.keys()) and all(torch.equal(rs[k], ps[k]) for k in rs)
    assert sum(p.numel() for p in prod.parameters()) == 2500998
    src, dst, _, _ = synth.make_batch([41, 42, 43], 2048)
    with torch.no_grad():
        torch.manual_seed(0); a = ref(src, dst)
        torch.manual_seed(0); b = RL.model_v2_forward(rs, src, dst)
    assert set(a.keys()) == set(b.keys())
    for k in ("src_dst_feats_2", "src_dst_feats_2_prime", "src_dst_weights_2", "src_dst_weights_2_prime"):
        assert float((a[k] - b[k]).abs().max()) < 1e-5, k
    for lv in range(3):
        assert float(RL.rotation_angle_deg(a["rotation"][lv], b["rotation"][lv]).max()) < 1e-4
        assert float((a["translation"][lv] - b["translation"][lv]).abs().max()) < 1e-5


def test_model_v4_matches_reference():
    """models/model_v4 (coarse stage returning coord_dist / feats_dist): state_dict keys, seeded init and forward."""
    from common import build_product_model_v4
    ns = H.load_reference()
    assert ns.Model_V4 is not None, getattr(ns, "v4_error", None)
    torch.manual_seed(7)
    ref = ns.Model_V4(Args())
    ref.feature_extraction.load_state_dict(torch.load(H.pretrained_feats_path(), map_location="cpu"))
    g = torch.Generator().manual_seed(8)
    for name in ("coarse_corres", "fine_corres_2", "fine_corres_1"):
        H.randomize_bn_(getattr(ref, name), g)
    ref.eval()
    prod = build_product_model_v4(seed=7)
    rs, ps = ref.state_dict(), prod.state_dict()
    assert list(rs.keys()) == list(ps.keys()) and all(torch.equal(rs[k], ps[k]) for k in rs)
    src, dst, _, _ = synth.make_batch([44, 45], 2048)
    with torch.no_grad():
        torch.manual_seed(0); a = ref(src, dst)
        torch.manual_seed(0); b = RL.model_v4_forward(rs, src, dst)
    b.pop("_stage_inputs")
    assert set(a.keys()) == set(b.keys())
    assert a["coord_dist"].shape == (2, 256, 8) and a["feats_dist"].shape == (2, 256, 8)
    assert float((a["coord_dist"] - b["coord_dist"]).abs().max()) < 1e-5
    assert float((a["feats_dist"] - b["feats_dist"]).abs().max()) < 1e-5
    for k in ("src_dst_feats_2", "src_dst_weights_2", "src_dst_weights_2_prime"):
        assert float((a[k] - b[k]).abs().max()) < 1e-5, k
    for lv in range(3):
        assert float(RL.rotation_angle_deg(a["rotation"][lv], b["rotation"][lv]).max()) < 1e-4
        assert float((a["translation"][lv] - b["translation"][lv]).abs().max()) < 1e-5


def test_metrics_oracle_matches_reference_functions():
    """oracle/ref_metrics.py against the UNMODIFIED models/utils.py:calc_error_np and losses/losses.py:calc_rot_rre_err /
    calc_tran_rte_err (Euler conversion injected, see ref_harness.load_reference_losses) on fresh random poses."""
    import numpy as np
    from oracle import ref_metrics as RM
    sys_path_tests = __import__("os").path.join(__import__("os").path.dirname(__file__), "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("_mk", __import__("os").path.join(sys_path_tests, "make_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    U, Ls = H.load_reference().utils, H.load_reference_losses()
    pred_R, pred_t = mk.random_poses(32, 101)
    gt_R, gt_t = mk.random_poses(32, 102)
    for i in range(32):
        want = U.calc_error_np(pred_R[i].numpy(), pred_t[i].numpy(), gt_R[i].numpy(), gt_t[i].numpy())
        got = RM.calc_error_np(pred_R[i].numpy(), pred_t[i].numpy(), gt_R[i].numpy(), gt_t[i].numpy())
        assert abs(want[0] - got[0]) < 1e-3 and abs(want[1] - got[1]) < 1e-6
    R_err_deg, geo = Ls.calc_rot_rre_err(pred_R, gt_R)
    T_err_mean, eucl = Ls.calc_tran_rte_err(pred_t, gt_t)
    o_deg, o_geo, _ = RM.calc_rot_rre_err(pred_R.numpy(), gt_R.numpy())
    o_tm, o_eu, _ = RM.calc_tran_rte_err(pred_t.numpy(), gt_t.numpy())
    assert np.abs(o_geo - geo.numpy()).max() < 2e-3 and np.abs(o_deg - R_err_deg.numpy()).max() < 2e-3
    assert np.abs(o_eu - eucl.numpy()).max() < 1e-6 and np.abs(o_tm - T_err_mean.numpy()).max() < 1e-6


def test_preprocess_oracle_matches_reference_classes():
    """oracle/ref_preprocess.py against the UNMODIFIED dataset/dataset_utils.py (PointCloudFilter.remove_points_by_range,
    PointCloudResampler under a seeded numpy RNG) and transform/rodrigues.py (SE3.exp) on fresh inputs."""
    import numpy as np
    from oracle import ref_preprocess as RP
    DU = H.load_reference_file("dataset/dataset_utils.py", "_ref_dataset_utils", stubs=("open3d",))
    RO = H.load_reference_file("transform/rodrigues.py", "_ref_rodrigues")
    rng = np.random.default_rng(77)
    for n, num in ((9000, 2048), (1500, 2048), (2048, 2048)):
        pc = (rng.normal(size=(n, 3)) * 35).astype(np.float32)
        it = rng.random(n).astype(np.float32)
        a, b = DU.PointCloudFilter(max_range=50.0).remove_points_by_range(pc, it)
        c, d = RP.remove_points_by_range(pc, it, 50.0)
        assert np.array_equal(a, c) and np.array_equal(b, d)
        m = a.shape[0]
        np.random.seed(n)
        idx = np.random.choice(m, num - m, replace=True) if m <= num else np.random.choice(m, num, replace=False)
        np.random.seed(n)
        r1 = DU.PointCloudResampler(num)(a, b)
        r2 = RP.resample(a, b, num, idx)
        assert np.array_equal(r1[0], r2[0]) and np.array_equal(r1[1], r2[1])
    x = torch.tensor(rng.normal(size=(32, 6)) * 0.4, dtype=torch.float32)
    x[:4] *= 1e-3
    assert np.abs(RO.SE3().exp(x).numpy() - RP.se3_exp(x.numpy())).max() < 1e-6


def test_calib_eval_oracle_matches_reference_class():
    """oracle/ref_metrics.calib_eval_results == the unmodified reference CalibEval (metrics/calibeval.py) on fresh seeded
    transforms (not the golden ones): dictionary keys / order identical, numbers within fp32 rounding of the angles."""
    import numpy as np
    from oracle import ref_metrics as RM
    CE = H.load_reference_calibeval()
    g = torch.Generator().manual_seed(77)

    def tf(n):
        q = torch.nn.functional.normalize(torch.randn(n, 4, generator=g), dim=1)
        w, x, y, z = q.unbind(1)
        R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                         2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                         2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1).view(n, 3, 3)
        T = torch.eye(4).repeat(n, 1, 1)
        T[:, :3, :3], T[:, :3, 3] = R, torch.randn(n, 3, generator=g) * 0.3
        return T

    batches = [(tf(n), tf(n)) for n in (4, 7, 2, 9)]
    ev = CE.CalibEval(None)
    for gt, pr in batches:
        ev.add_batch(gt, pr)
    want = ev.get_results()
    got = RM.calib_eval_results([(a.numpy(), b.numpy()) for a, b in batches])
    assert list(got) == list(want)
    for k in want:
        assert np.abs(np.asarray(got[k]) - np.asarray(want[k])).max() < 2e-3, k
