"""TEST INFRASTRUCTURE -- CPU restatement (numpy) of the reference's input-pipeline steps in front of the registration path.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Follows (reference file:line):
  remove_points_by_range   dataset/dataset_utils.py:113-125
  resample                 dataset/dataset_utils.py:188-223 (PointCloudResampler.__call__; the random index list is an
                           argument here -- the reference draws it with np.random.choice)
  se3_exp                  transform/rodrigues.py:526-550 with sinc1/2/3 of rodrigues.py:6-18,100-112,132-144
Pinned against the unmodified reference classes run on CPU (tests/test_oracle_vs_reference.py: PointCloudFilter,
PointCloudResampler with a seeded numpy RNG, SE3.exp) and tests/golden/preprocess.npz."""
import numpy as np


def remove_points_by_range(point_cloud, intensity, max_range):
    rng = np.linalg.norm(point_cloud, axis=1)
    keep = rng < max_range
    return point_cloud[keep, :], (intensity[keep] if intensity is not None else None)


def resample(point_cloud, intensity, num_points, indices):
    n = point_cloud.shape[0]
    if num_points == -1:
        return point_cloud, intensity
    if n <= num_points:
        pc = np.vstack((point_cloud, point_cloud[indices]))
        it = np.hstack((intensity, intensity[indices])) if intensity is not None else None
        return pc, it
    return point_cloud[indices], (intensity[indices] if intensity is not None else None)


def _sinc(t):
    t = np.asarray(t, np.float64)
    small = np.abs(t) < 0.01
    t2 = t * t
    ts = np.where(small, 1.0, t)
    s1 = np.where(small, 1 - t2 / 6 * (1 - t2 / 20 * (1 - t2 / 42)), np.sin(ts) / ts)
    s2 = np.where(small, 0.5 * (1 - t2 / 12 * (1 - t2 / 30 * (1 - t2 / 56))), (1 - np.cos(ts)) / (ts * ts))
    s3 = np.where(small, 1 / 6 * (1 - t2 / 20 * (1 - t2 / 42 * (1 - t2 / 72))), (ts - np.sin(ts)) / ts ** 3)
    return s1, s2, s3


def se3_exp(x):
    x = np.asarray(x, np.float64).reshape(-1, 6)
    w, v = x[:, :3], x[:, 3:]
    t = np.linalg.norm(w, axis=1)
    W = np.zeros((x.shape[0], 3, 3))
    W[:, 0, 1], W[:, 0, 2], W[:, 1, 0], W[:, 1, 2], W[:, 2, 0], W[:, 2, 1] = -w[:, 2], w[:, 1], w[:, 2], -w[:, 0], -w[:, 1], w[:, 0]
    S = W @ W
    s1, s2, s3 = (a[:, None, None] for a in _sinc(t))
    I = np.eye(3)[None]
    R = I + s1 * W + s2 * S
    V = I + s2 * W + s3 * S
    g = np.zeros((x.shape[0], 4, 4))
    g[:, :3, :3] = R
    g[:, :3, 3] = (V @ v[:, :, None])[:, :, 0]
    g[:, 3, 3] = 1.0
    return g
