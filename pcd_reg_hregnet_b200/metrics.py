"""Pose-error metrics on device: the reference's evaluation functions with the same names,
argument order and return values, computed by ONE kernel launch (csrc/metrics.cu) on the tensors the pose head (or the
multi-GPU pose all-gather, dist.gather_poses) left on the device -- no .cpu() round trips, no Python lists.

Mirrors (reference file:line):
  calc_rot_rre_err, calc_tran_rte_err   losses/losses.py:138-164
  calc_error                            models/utils.py:132-138 (calc_error_np, batched)
  geodesic_distance                     metrics/calibeval.py:172-196
  PoseErrorMeter                        metrics/calibeval.py:72-106 (add_batch) + the mean_error summary (:45-70)
"""
import torch

from ._lib import HrnError, call, ptr, stream


def _rows9(R):
    if R.shape[-2:] != (3, 3):
        raise HrnError("rotation must be [B,3,3]")
    return R.reshape(-1, 9).contiguous().float()


def pose_errors(pred_R, pred_t, gt_R, gt_t, mode=0, sums=None):
    """Per-pair (geo_deg [B], eucl [B], euler_xyz_deg [B,3], t_err [B,3]) of R_err / t_err; mode 0: R_err = pred_R^T gt_R,
    t_err = pred_t - gt_t; mode 1: error = pred_tf . gt_tf.  `sums` [8] (optional) is accumulated into."""
    pR, gR = _rows9(pred_R), _rows9(gt_R)
    pt, gt_ = pred_t.reshape(-1, 3).contiguous().float(), gt_t.reshape(-1, 3).contiguous().float()
    B = pR.shape[0]
    if not (gR.shape[0] == pt.shape[0] == gt_.shape[0] == B):
        raise HrnError("pose_errors: batch sizes differ")
    dev = pR.device
    geo, eucl = torch.empty(B, device=dev), torch.empty(B, device=dev)
    euler, terr = torch.empty(B, 3, device=dev), torch.empty(B, 3, device=dev)
    call("hrn_pose_errors", ptr(pR), ptr(pt), ptr(gR), ptr(gt_), B, int(mode), ptr(geo), ptr(eucl), ptr(euler), ptr(terr),
         ptr(sums), stream())
    return geo, eucl, euler, terr


def calc_rot_rre_err(pred_R, gt_R):
    """-> (mean |Euler XYZ error| in degrees per axis [3], geodesic distance in degrees per pair [B])."""
    z = torch.zeros(pred_R.shape[0], 3, device=pred_R.device)
    geo, _, euler, _ = pose_errors(pred_R, z, gt_R, z)
    return euler.abs().mean(dim=0), geo


def calc_tran_rte_err(pred_t, gt_t):
    """-> (mean |translation error| per axis [3], Euclidean distance per pair [B])."""
    eye = torch.eye(3, device=pred_t.device).expand(pred_t.shape[0], 3, 3)
    _, eucl, _, terr = pose_errors(eye, pred_t, eye, gt_t)
    return terr.abs().mean(dim=0), eucl


def calc_error(pred_R, pred_t, gt_R, gt_t):
    """Batched calc_error_np: (rotation error in degrees [B], translation error [B])."""
    geo, eucl, _, _ = pose_errors(pred_R, pred_t, gt_R, gt_t)
    return geo, eucl


def geodesic_distance(x):
    """x [B,4,4] error transforms -> [mean geodesic angle in degrees, mean translation norm] (Python floats, like the
    reference)."""
    B = x.shape[0]
    eye = torch.eye(3, device=x.device).expand(B, 3, 3)
    z = torch.zeros(B, 3, device=x.device)
    geo, eucl, _, _ = pose_errors(eye, z, x[:, :3, :3], -x[:, :3, 3])
    return [geo.mean().item(), eucl.mean().item()]


class PoseErrorMeter:
    """Running means over batches / ranks without leaving the device: add_batch(gt_tf, pred_tf) follows CalibEval.add_batch
    (error = pred_tf . gt_tf); result() -> dict(geodesic_deg, translation, euler_abs_deg [3], t_abs [3], count)."""

    def __init__(self, device):
        self.sums = torch.zeros(8, device=device)
        self.count = 0

    def add_batch(self, gt_tf, pred_tf):
        pose_errors(pred_tf[:, :3, :3], pred_tf[:, :3, 3], gt_tf[:, :3, :3], gt_tf[:, :3, 3], mode=1, sums=self.sums)
        self.count += int(gt_tf.shape[0])

    def add_poses(self, pred_R, pred_t, gt_R, gt_t):
        pose_errors(pred_R, pred_t, gt_R, gt_t, mode=0, sums=self.sums)
        self.count += int(pred_R.shape[0])

    def result(self):
        m = (self.sums / max(self.count, 1)).tolist()
        return dict(geodesic_deg=m[0], translation=m[1], euler_abs_deg=m[2:5], t_abs=m[5:8], count=self.count)
