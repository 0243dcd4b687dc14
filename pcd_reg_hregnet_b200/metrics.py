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


# ------------------------------------------------------------------------------------------------------------------
# The reference's evaluators, same class names / methods / result dictionaries (metrics/calibeval.py:11-337, 344-380)
# ------------------------------------------------------------------------------------------------------------------
class CalibEval:
    """Reference: metrics/calibeval.py:11-337.  add_batch(gt_tf [B,4,4], pred_tf [B,4,4]) accumulates, per sample, the Euler
    XYZ angles (degrees) and translation of the error transform pred_tf . gt_tf and of the prediction, and per batch the
    mean geodesic / Euclidean distance -- everything stays on the device (one kernel launch per batch, no .cpu() sync);
    get_stats / getSD / get_results / save_results produce the reference's numbers, keys and JSON layout."""

    def __init__(self, config=None, translation_threshold=None, rotation_threshold=None):
        self.config = config
        self.translation_threshold = translation_threshold
        self.rotation_threshold = rotation_threshold
        self.reset()

    def reset(self):
        self._err_euler, self._err_trans, self._pred, self._geo = [], [], [], []
        self.success_idx = []                  # the reference never fills it (its threshold code is commented out, :103-106)
        self.results = {}

    def add_batch(self, gt_tf, pred_tf, idx=None, return_results=False):
        B = gt_tf.shape[0]
        pR, pt = pred_tf[:, :3, :3], pred_tf[:, :3, 3]
        geo, eucl, e_euler, e_t = pose_errors(pR, pt, gt_tf[:, :3, :3], gt_tf[:, :3, 3], mode=1)
        eye = torch.eye(3, device=gt_tf.device).expand(B, 3, 3)
        z = torch.zeros(B, 3, device=gt_tf.device)
        _, _, p_euler, _ = pose_errors(eye, z, pR, z)                   # R_err = I^T pred_R: Euler angles of the prediction
        self._err_euler.append(e_euler)
        self._err_trans.append(e_t)
        self._pred.append(torch.cat([p_euler, pt.reshape(B, 3).float()], dim=1))
        self._geo.append(torch.stack([geo.mean(), eucl.mean()]))       # per batch, like the reference (:99, :196)

    # -- host side: one read-back ------------------------------------------------------------------------------------
    def _host(self):
        import numpy as np
        if not self._err_euler:
            raise ValueError("no batches added")
        loss_r = torch.cat(self._err_euler).double().cpu().numpy()
        loss_t = torch.cat(self._err_trans).double().cpu().numpy()
        pred = torch.cat(self._pred).double().cpu().numpy()
        geo = torch.stack(self._geo).double().cpu().numpy()
        return np, loss_r, loss_t, pred, geo

    def get_stats(self):
        np, loss_r, loss_t, _, geo = self._host()
        return np.abs(loss_r).mean(axis=0), np.abs(loss_t).mean(axis=0), geo.mean(axis=0)

    def getSD(self):
        np, loss_r, loss_t, _, geo = self._host()
        return np.abs(loss_r).std(axis=0), np.abs(loss_t).std(axis=0), np.abs(geo[:, 0]).std(axis=0), np.abs(geo[:, 1]).std(axis=0)

    def compute_recall(self):
        n = sum(int(t.shape[0]) for t in self._err_euler)
        return len(self.success_idx) / n if n else 0.0

    def get_results(self):
        np, loss_r, loss_t, pred, _ = self._host()
        r, t, g = self.get_stats()
        sd_t, sd_r, sd_dR, sd_dT = self.getSD()        # the reference unpacks getSD() in this order (calibeval.py:50, :295)
        self.results = {
            "pred_calib": pred.tolist(),
            "error_calib": np.concatenate((loss_r, loss_t), axis=1).tolist(),
            "mean_error": sum([r.tolist(), t.tolist(), g.tolist()], []),
            "sd": sum([sd_r.tolist(), sd_t.tolist()], []),
            "mean_sd": [np.mean(sd_r).tolist(), np.mean(sd_t).tolist()],
            "mean_sd_dRT": [np.mean(sd_dR).tolist(), np.mean(sd_dT).tolist()],
        }
        return self.results

    def save_results(self):
        import json
        import os
        self.get_results()
        dc = self.config.dataset_config
        name = "results_" + "_" + self.config.dataset + "_" + dc.distribution + "_" + str(dc.max_rot_error) + "_" + \
            str(dc.max_trans_error) + ".json"                             # calibeval.py:321-327
        with open(os.path.join(dc.results_path, name), "w") as f:
            json.dump(self.results, f, indent=4)


class MultiLayerCalibEval:
    """Reference: metrics/calibeval.py:344-380 -- one CalibEval per pose level, one combined JSON file."""

    def __init__(self, config=None, num_layers=3, translation_threshold=None, rotation_threshold=None):
        self.config = config
        self.num_layers = num_layers
        self.evaluators = {layer: CalibEval(config, translation_threshold, rotation_threshold) for layer in range(num_layers)}

    def reset(self):
        for ev in self.evaluators.values():
            ev.reset()

    def add_batch(self, layer, gt_tf, pred_tf, idx=None, return_results=False):
        if layer not in self.evaluators:
            raise ValueError(f"Layer {layer} is not valid. Valid layers: 0 to {self.num_layers - 1}.")
        self.evaluators[layer].add_batch(gt_tf, pred_tf, idx, return_results)

    def save_all_results(self, output_file):
        import json
        combined = {f"layer_{layer}": ev.get_results() for layer, ev in self.evaluators.items()}
        dc = self.config.dataset_config
        combined.update({"dataset": self.config.dataset + dc.version, "model": dc.model, "translation": dc.max_trans_error,
                         "rotation": dc.max_rot_error, "distribution": dc.distribution})
        with open(output_file, "w") as f:
            json.dump(combined, f, indent=4)
