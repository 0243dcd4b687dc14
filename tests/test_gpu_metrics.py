"""GPU: pose-error metrics kernel (csrc/metrics.cu, through the C ABI / pcd_reg_hregnet_b200.metrics) against the golden
reference outputs (tests/golden/pose_metrics.npz) and the fp64 oracle (oracle/ref_metrics.py).  Tolerances: angles
2e-3 degrees (fp32 acos / atan2 on R_err), translations 1e-6 m."""
import numpy as np
import pytest
import torch

from common import load_golden
from oracle import ref_metrics as RM
from pcd_reg_hregnet_b200 import metrics as M

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_metrics_match_reference_golden():
    g = load_golden("pose_metrics")
    pR, pt, gR, gt = (g[k].to(DEV) for k in ("pred_R", "pred_t", "gt_R", "gt_t"))
    R_err_deg, geo = M.calc_rot_rre_err(pR, gR)
    T_err_mean, eucl = M.calc_tran_rte_err(pt, gt)
    rot, trans = M.calc_error(pR, pt, gR, gt)
    torch.cuda.synchronize()
    assert float((geo.cpu() - g["geo"]).abs().max()) < 2e-3
    assert float((R_err_deg.cpu() - g["R_err_deg"]).abs().max()) < 2e-3
    assert float((eucl.cpu() - g["eucl"]).abs().max()) < 1e-6
    assert float((T_err_mean.cpu() - g["T_err_mean"]).abs().max()) < 1e-6
    ce = g["calc_error_np"]
    assert float((rot.cpu().double() - ce[:, 0]).abs().max()) < 2e-3 and float((trans.cpu().double() - ce[:, 1]).abs().max()) < 1e-6


@pytest.mark.parametrize("B", [1, 127, 128, 1000])
def test_calib_convention_and_running_means(B):
    """mode 1 (CalibEval.add_batch: error = pred_tf . gt_tf), the accumulated sums over two batches, and the ragged last
    block of the reduction."""
    g = torch.Generator().manual_seed(B)

    def tf(n):
        q = torch.nn.functional.normalize(torch.randn(n, 4, generator=g), dim=1)
        w, x, y, z = q.unbind(1)
        R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                         2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                         2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], 1).view(n, 3, 3)
        T = torch.eye(4).repeat(n, 1, 1)
        T[:, :3, :3] = R
        T[:, :3, 3] = torch.randn(n, 3, generator=g)
        return T

    gt1, pr1, gt2, pr2 = tf(B), tf(B), tf(B), tf(B)
    meter = M.PoseErrorMeter(DEV)
    meter.add_batch(gt1.to(DEV), pr1.to(DEV))
    meter.add_batch(gt2.to(DEV), pr2.to(DEV))
    res = meter.result()
    o_geo = np.concatenate([RM.calib_error(gt1.numpy(), pr1.numpy())[0], RM.calib_error(gt2.numpy(), pr2.numpy())[0]])
    o_tr = np.concatenate([RM.calib_error(gt1.numpy(), pr1.numpy())[1], RM.calib_error(gt2.numpy(), pr2.numpy())[1]])
    assert res["count"] == 2 * B
    assert abs(res["geodesic_deg"] - o_geo.mean()) < 5e-3 and abs(res["translation"] - o_tr.mean()) < 1e-4
    err = pr1 @ gt1
    gd = M.geodesic_distance(err.to(DEV))
    assert abs(gd[0] - RM.calib_error(gt1.numpy(), pr1.numpy())[0].mean()) < 5e-3
    assert abs(gd[1] - RM.calib_error(gt1.numpy(), pr1.numpy())[1].mean()) < 1e-4


def test_calib_eval_classes_write_the_reference_json(tmp_path):
    """pcd_reg_hregnet_b200.metrics.MultiLayerCalibEval / CalibEval (metrics/calibeval.py:11-337,344-380) on the device:
    the same inputs the reference's own evaluator was fed (tests/golden/calib_eval.json) -> the same JSON document
    (keys, order, list lengths; numbers within fp32 angle rounding), and the single-evaluator file name of save_results."""
    import json
    import os
    import types
    from test_golden_oracle import _calib_inputs
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "calib_eval.json")))
    cfg = types.SimpleNamespace(dataset="synthetic", dataset_config=types.SimpleNamespace(
        version="_v0", model="HRegNet", max_trans_error=0.5, max_rot_error=20.0, distribution="uniform", results_path=str(tmp_path)))
    ev = M.MultiLayerCalibEval(cfg, num_layers=3)
    for layer, batches in _calib_inputs().items():
        for g, p in batches:
            ev.add_batch(layer, g.to(DEV), p.to(DEV))
    out = tmp_path / "all.json"
    ev.save_all_results(str(out))
    got = json.load(open(out))
    assert list(got) == list(gold)
    for k in ("dataset", "model", "translation", "rotation", "distribution"):
        assert got[k] == gold[k]
    for layer in range(3):
        a, b = got[f"layer_{layer}"], gold[f"layer_{layer}"]
        assert list(a) == list(b)
        for k in b:
            assert np.asarray(a[k]).shape == np.asarray(b[k]).shape
            assert np.abs(np.asarray(a[k]) - np.asarray(b[k])).max() < 5e-3, (layer, k)
    with pytest.raises(ValueError):
        ev.add_batch(3, batches[0][0].to(DEV), batches[0][1].to(DEV))
    ev.evaluators[0].save_results()
    assert (tmp_path / "results__synthetic_uniform_20.0_0.5.json").exists()
    assert ev.evaluators[0].compute_recall() == 0.0
    ev.reset()
    with pytest.raises(ValueError):
        ev.evaluators[0].get_results()
