"""TEST INFRASTRUCTURE -- CPU restatement (numpy, fp64 unless stated) of the reference's pose-error metrics.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Follows (reference file:line):
  calc_error_np            models/utils.py:132-138
  calc_rot_rre_err         losses/losses.py:138-152
  calc_tran_rte_err        losses/losses.py:154-164
  geodesic_distance        metrics/calibeval.py:172-196
  add_batch error          metrics/calibeval.py:83-84 (error = pred_tf.bmm(gt_tf))
  calib_eval_results       metrics/calibeval.py:72-164, 45-70 (CalibEval.add_batch / get_stats / getSD / get_results)

Pinned against the unmodified reference functions run on CPU by tests/test_oracle_vs_reference.py (calc_error_np directly;
calc_rot_rre_err / calc_tran_rte_err with the Euler conversion below injected for the absent pytorch3d).
matrix_to_euler_angles is pytorch3d==0.7.8 code (Dockerfile:45-46), not vendored in the reference: restated from its
published algorithm -- for the "XYZ" convention (atan2(-M12, M22), asin(M02), atan2(-M01, M00)) -- PARITY UNPINNED for
that one function (no reference test or fixture exercises it)."""
import numpy as np


def matrix_to_euler_angles_xyz(M):
    M = np.asarray(M, dtype=np.float64)
    return np.stack([np.arctan2(-M[..., 1, 2], M[..., 2, 2]), np.arcsin(np.clip(M[..., 0, 2], -1.0, 1.0)),
                     np.arctan2(-M[..., 0, 1], M[..., 0, 0])], -1)


def _geo_deg(R_err):
    c = np.clip((np.trace(R_err, axis1=-2, axis2=-1) - 1.0) / 2.0, -1.0, 1.0)
    return np.degrees(np.arccos(c))


def calc_error_np(pred_R, pred_t, gt_R, gt_t):
    pred_R, gt_R = np.asarray(pred_R, np.float64), np.asarray(gt_R, np.float64)
    return float(_geo_deg(pred_R.T @ gt_R)), float(np.linalg.norm(np.asarray(pred_t, np.float64) - np.asarray(gt_t, np.float64)))


def calc_rot_rre_err(pred_R, gt_R):
    pred_R, gt_R = np.asarray(pred_R, np.float64), np.asarray(gt_R, np.float64)
    R_err = np.swapaxes(pred_R, -1, -2) @ gt_R
    eul = np.degrees(matrix_to_euler_angles_xyz(R_err))
    return np.abs(eul).mean(0), _geo_deg(R_err), eul


def calc_tran_rte_err(pred_t, gt_t):
    e = np.asarray(pred_t, np.float64) - np.asarray(gt_t, np.float64)
    return np.abs(e).mean(0), np.linalg.norm(e, axis=1), e


def calib_error(gt_tf, pred_tf):
    """error = pred_tf @ gt_tf; -> (geodesic degrees [B], translation norm [B])."""
    err = np.asarray(pred_tf, np.float64) @ np.asarray(gt_tf, np.float64)
    return _geo_deg(err[:, :3, :3]), np.linalg.norm(err[:, :3, 3], axis=1)


def calib_eval_results(batches):
    """CalibEval over `batches` = [(gt_tf [B,4,4], pred_tf [B,4,4]), ...] -> the dictionary of CalibEval.get_results
    (metrics/calibeval.py:45-70): per-sample Euler XYZ angles (degrees) / translations of error = pred_tf . gt_tf and of the
    prediction, per-BATCH mean geodesic / Euclidean distances, their |.| means and standard deviations.  The reference
    unpacks getSD() as `sd_t, sd_r, ...` although it returns (rotation, translation, ...) (:50 vs :152-164): "sd" therefore
    lists the translation deviations first, and so does this restatement."""
    loss_r, loss_t, pred, geo = [], [], [], []
    for gt_tf, pred_tf in batches:
        gt_tf, pred_tf = np.asarray(gt_tf, np.float64), np.asarray(pred_tf, np.float64)
        err = pred_tf @ gt_tf
        loss_r.append(np.degrees(matrix_to_euler_angles_xyz(err[:, :3, :3])))
        loss_t.append(err[:, :3, 3])
        pred.append(np.concatenate([np.degrees(matrix_to_euler_angles_xyz(pred_tf[:, :3, :3])), pred_tf[:, :3, 3]], 1))
        geo.append([_geo_deg(err[:, :3, :3]).mean(), np.linalg.norm(err[:, :3, 3], axis=1).mean()])
    loss_r, loss_t, pred, geo = np.concatenate(loss_r), np.concatenate(loss_t), np.concatenate(pred), np.asarray(geo)
    r, t, g = np.abs(loss_r).mean(0), np.abs(loss_t).mean(0), geo.mean(0)
    sd_t, sd_r, sd_dR, sd_dT = np.abs(loss_r).std(0), np.abs(loss_t).std(0), np.abs(geo[:, 0]).std(0), np.abs(geo[:, 1]).std(0)
    return {"pred_calib": pred.tolist(), "error_calib": np.concatenate((loss_r, loss_t), 1).tolist(),
            "mean_error": sum([r.tolist(), t.tolist(), g.tolist()], []), "sd": sum([sd_r.tolist(), sd_t.tolist()], []),
            "mean_sd": [float(np.mean(sd_r)), float(np.mean(sd_t))], "mean_sd_dRT": [float(np.mean(sd_dR)), float(np.mean(sd_dT))]}
