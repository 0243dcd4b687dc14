"""CPU (any box): the oracle restatement reproduces the committed reference outputs (tests/golden/*.npz)."""
import pytest
import torch

from oracle import ref_layers as RL
from common import build_product_hregnet, load_golden, unflatten


@pytest.mark.parametrize("name", ["hregnet_b2_n2048", "hregnet_uniform_b1_n1500"])
def test_oracle_reproduces_golden(name):
    torch.set_num_threads(8)
    gd = load_golden(name)
    sd = build_product_hregnet(seed=7).state_dict()     # same keys / values as the reference net (see common.py)
    trace = {}
    with torch.no_grad():
        out = RL.hregnet_forward(sd, gd["src"], gd["dst"], trace=trace)
    for side in ("src", "dst"):
        for lv in (1, 2, 3):
            assert torch.equal(trace[f"{side}_trace"][f"fps_idx_{lv}"], gd[f"{side}_fps_idx_{lv}"])
        for k, v in unflatten(gd, f"{side}_feats.").items():
            assert torch.equal(out[f"{side}_feats"][k], v), k            # bit-identical feature extraction
    assert torch.equal(trace["coarse_idx"], gd["coarse_idx"])
    for i in range(3):
        assert float(RL.rotation_angle_deg(out["rotation"][i], gd[f"rotation.{i}"]).max()) < 1e-4
        assert float((out["translation"][i] - gd[f"translation.{i}"]).abs().max()) < 1e-5
    for lv in (3, 2, 1):
        assert float((out[f"src_xyz_corres_{lv}"] - gd[f"src_xyz_corres_{lv}"]).abs().max()) < 1e-4
        assert float((out[f"src_dst_weights_{lv}"] - gd[f"src_dst_weights_{lv}"]).abs().max()) < 1e-6


def test_metrics_oracle_reproduces_reference_outputs():
    """oracle/ref_metrics.py against tests/golden/pose_metrics.npz = outputs of the reference's calc_error_np,
    calc_rot_rre_err and calc_tran_rte_err (fp32 torch in the reference, fp64 numpy in the oracle)."""
    import numpy as np
    from oracle import ref_metrics as RM
    g = {k: v.numpy() for k, v in load_golden("pose_metrics").items()}
    per_pair = np.array([RM.calc_error_np(g["pred_R"][i], g["pred_t"][i], g["gt_R"][i], g["gt_t"][i]) for i in range(64)])
    d = np.abs(per_pair - g["calc_error_np"])                        # calc_error_np ran in fp32 numpy
    assert d[:, 0].max() < 2e-3 and d[:, 1].max() < 1e-6
    R_err_deg, geo, _ = RM.calc_rot_rre_err(g["pred_R"], g["gt_R"])
    T_err_mean, eucl, _ = RM.calc_tran_rte_err(g["pred_t"], g["gt_t"])
    assert np.abs(geo - g["geo"]).max() < 2e-3 and np.abs(R_err_deg - g["R_err_deg"]).max() < 2e-3   # fp32 acos / atan2
    assert np.abs(eucl - g["eucl"]).max() < 1e-6 and np.abs(T_err_mean - g["T_err_mean"]).max() < 1e-6


def _calib_inputs():
    """The seeded transforms tests/golden/make_golden.py::calib_eval_golden fed to the reference's MultiLayerCalibEval."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("_mk", os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    return {layer: [(mk.random_transforms(B, 100 + 10 * layer + bi), mk.random_transforms(B, 200 + 10 * layer + bi))
                    for bi, B in enumerate((5, 8, 3))] for layer in range(3)}


def test_calib_eval_oracle_reproduces_reference_json():
    """oracle/ref_metrics.calib_eval_results against tests/golden/calib_eval.json = the file the reference's own
    MultiLayerCalibEval.save_all_results wrote (metrics/calibeval.py:367-380): same keys, numbers within fp32 rounding."""
    import json
    import os
    import numpy as np
    from oracle import ref_metrics as RM
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "calib_eval.json")))
    assert {"dataset", "model", "translation", "rotation", "distribution", "layer_0", "layer_1", "layer_2"} == set(gold)
    for layer, batches in _calib_inputs().items():
        got = RM.calib_eval_results([(g.numpy(), p.numpy()) for g, p in batches])
        want = gold[f"layer_{layer}"]
        assert list(got) == list(want)                                       # same keys, same order
        for k in want:
            tol = 2e-3 if k in ("pred_calib", "error_calib", "mean_error", "sd", "mean_sd", "mean_sd_dRT") else 0
            assert np.abs(np.asarray(got[k]) - np.asarray(want[k])).max() < tol, (layer, k)


def test_regression_head_oracle_reproduces_reference_outputs():
    """oracle/ref_layers.regression_head against tests/golden/regression_head.npz = the reference's RegressionHead
    (models/model_v2/layers.py:625-668)."""
    from oracle import ref_layers as RL
    g = load_golden("regression_head")
    sd = {k[3:]: v for k, v in g.items() if k.startswith("sd.")}
    rot, trans = RL.regression_head(sd, "", g["src"], g["cor"], g["w"])
    assert float((rot - g["rotation"]).abs().max()) < 1e-6 and float((trans - g["translation"]).abs().max()) < 1e-6


def test_preprocess_oracle_reproduces_reference_outputs():
    """oracle/ref_preprocess.py against tests/golden/preprocess.npz (reference PointCloudFilter / PointCloudResampler /
    SE3.exp outputs): kept sets and resampled clouds bit-identical, SE3.exp within fp32 rounding."""
    import numpy as np
    from oracle import ref_preprocess as RP
    g = {k: v.numpy() for k, v in load_golden("preprocess").items()}
    for n in ("a", "b"):
        fp, fi = RP.remove_points_by_range(g[f"{n}_pc"], g[f"{n}_int"], 60.0)
        assert np.array_equal(fp, g[f"{n}_filtered"]) and np.array_equal(fi, g[f"{n}_filtered_int"])
        rp, ri = RP.resample(fp, fi, 4096, g[f"{n}_idx"])
        assert np.array_equal(rp, g[f"{n}_resampled"]) and np.array_equal(ri, g[f"{n}_resampled_int"])
    assert np.abs(RP.se3_exp(g["twist"]) - g["se3_exp"]).max() < 1e-6


@pytest.mark.parametrize("which", ["v2", "v4"])
def test_oracle_reproduces_golden_model_variants(which):
    """oracle/ref_layers.model_v2_forward / model_v4_forward against the committed outputs of the UNMODIFIED reference
    classes (tests/golden/model_v{2,4}_b2_n2048.npz, made by make_golden.py variants) -- the pin of the oracle for these
    two models that travels to boxes without /root/reference.  torch.manual_seed(0) reproduces the host-generator
    batch shuffles of FineReg2 (model_v2/layers.py:493,497)."""
    from common import build_product_model_v2, build_product_model_v4
    torch.set_num_threads(8)
    gd = load_golden(f"model_{which}_b2_n2048")
    sd = (build_product_model_v2 if which == "v2" else build_product_model_v4)(seed=7).state_dict()
    with torch.no_grad():
        torch.manual_seed(0)
        out = (RL.model_v2_forward if which == "v2" else RL.model_v4_forward)(sd, gd["src"], gd["dst"])
    for i in range(3):
        assert float(RL.rotation_angle_deg(out["rotation"][i], gd[f"rotation.{i}"]).max()) < 1e-4
        assert float((out["translation"][i] - gd[f"translation.{i}"]).abs().max()) < 1e-5
    for k in ("src_dst_feats_2", "src_dst_feats_2_prime", "src_dst_weights_2", "src_dst_weights_2_prime",
              "src_xyz_2_trans", "dst_xyz_2", "src_feats_desc_2", "src_feats_sigmas_2"):
        assert float((out[k] - gd[k]).abs().max()) < 2e-5 * max(1.0, float(gd[k].abs().max())), k
    if which == "v2":
        for lv in (3, 2, 1):
            assert float((out[f"src_xyz_corres_{lv}"] - gd[f"src_xyz_corres_{lv}"]).abs().max()) < 1e-4
    else:
        assert float((out["coord_dist"] - gd["coord_dist"]).abs().max()) < 1e-5 * max(1.0, float(gd["coord_dist"].abs().max()))
        assert float((out["feats_dist"] - gd["feats_dist"]).abs().max()) < 1e-5
