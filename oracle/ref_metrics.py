"""TEST INFRASTRUCTURE -- CPU restatement (numpy, fp64 unless stated) of the reference's pose-error metrics.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Follows (reference file:line):
  calc_error_np            models/utils.py:132-138
  calc_rot_rre_err         losses/losses.py:138-152
  calc_tran_rte_err        losses/losses.py:154-164
  geodesic_distance        metrics/calibeval.py:172-196
  add_batch error          metrics/calibeval.py:83-84 (error = pred_tf.bmm(gt_tf))

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
