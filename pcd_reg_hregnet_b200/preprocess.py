"""Input pipeline step in front of the registration path, on the device.

Mirrors (reference file:line):
  remove_points_by_range   dataset/dataset_utils.py:113-125 (PointCloudFilter.remove_points_by_range)
  PointCloudResampler      dataset/dataset_utils.py:177-223
  se3_exp / apply_transform   transform/rodrigues.py:526-550,577-590 (SE3.exp / SE3.transform) as used by
                           transform/dataset_transforms.py:128-140 (gt = exp(-x), igt = exp(x), p1 = igt . p0)
  prepare_pairs            the three steps for a batch of raw sweeps -> (src [B,n,3], dst [B,n,3], gt [B,4,4], igt [B,4,4])

The reference runs these per sample with numpy in DataLoader workers.  Random draws stay explicit: every function takes
the index list / twist it should use (pass the reference's own numpy draws for bit-identical results) and only draws on
the device (torch RNG) when none is given."""
import torch

from ._lib import HrnError, call, ptr, stream


def remove_points_by_range_batched(xyz, intensity, offsets, max_range, max_sweep_points=None):
    """xyz [total,3] fp32 (sweeps concatenated), intensity [total] | None, offsets [S+1] int64 (device);
    max_sweep_points: host upper bound of the sweep sizes (default: total).
    -> (xyz_out, intensity_out | None, count [S] int32): kept points of sweep s = rows offsets[s] .. offsets[s]+count[s]."""
    if xyz.dtype != torch.float32 or xyz.dim() != 2 or xyz.shape[1] != 3:
        raise HrnError("xyz must be [total,3] float32")
    xyz = xyz.contiguous()
    S = offsets.numel() - 1
    mx = int(xyz.shape[0] if max_sweep_points is None else max_sweep_points)
    out = torch.empty_like(xyz)
    iout = torch.empty_like(intensity) if intensity is not None else None
    count = torch.empty(S, dtype=torch.int32, device=xyz.device)
    scratch = torch.empty(max(1, S * ((mx + 4095) // 4096)), dtype=torch.int32, device=xyz.device)
    call("hrn_range_filter", ptr(xyz), ptr(intensity.contiguous() if intensity is not None else None),
         ptr(offsets.to(torch.int64).contiguous()), S, mx, float(max_range), ptr(out), ptr(iout), ptr(count), ptr(scratch),
         stream())
    return out, iout, count


def remove_points_by_range(point_cloud, intensity, max_range):
    """One sweep: (point_cloud[range < max_range], intensity[range < max_range]), order preserved."""
    n = point_cloud.shape[0]
    offs = torch.tensor([0, n], dtype=torch.int64, device=point_cloud.device)
    out, iout, count = remove_points_by_range_batched(point_cloud, intensity, offs, max_range)
    c = int(count.item())
    return out[:c], (iout[:c] if iout is not None else None)


class PointCloudResampler:
    """Fixed-size resampling: pad with randomly repeated points when the cloud has <= num_points points, random subset
    without replacement otherwise (num_points = -1: unchanged).  `indices`: the index list to use -- pad indices
    [num_points - N] or selected indices [num_points] (e.g. the reference's np.random.choice draw); None = draw on device."""

    def __init__(self, num_points=1024):
        self._num_points = num_points

    def __call__(self, point_cloud, intensity=None, indices=None, generator=None):
        n, m = point_cloud.shape[0], self._num_points
        if m == -1:
            return point_cloud, intensity
        dev = point_cloud.device
        if n <= m:
            if indices is None:
                indices = torch.randint(0, n, (m - n,), device=dev, generator=generator)
            idx = torch.cat([torch.arange(n, device=dev), indices.to(dev).long()])
        else:
            idx = torch.randperm(n, device=dev, generator=generator)[:m] if indices is None else indices.to(dev).long()
        if idx.numel() != m:
            raise HrnError("resampler: index list has the wrong length")
        idx32 = idx.to(torch.int32).contiguous().view(1, m)
        out = torch.empty(1, m, 3, dtype=torch.float32, device=dev)
        call("hrn_gather_rows", ptr(point_cloud.contiguous().view(1, n, 3)), ptr(idx32), ptr(out), 1, n, m, 3, stream())
        return out[0], (intensity[idx] if intensity is not None else None)


def se3_exp(x):
    """twist [..., 6] = (w, v) -> [..., 4, 4]."""
    x_ = x.reshape(-1, 6).contiguous().float()
    g = torch.empty(x_.shape[0], 16, device=x.device)
    call("hrn_se3_exp", ptr(x_), x_.shape[0], ptr(g), stream())
    return g.view(*x.shape[:-1], 4, 4)


def apply_transform(points, g):
    """points [B,N,3], g [B,4,4] -> R p + t."""
    B, N, _ = points.shape
    R = g[:, :3, :3].reshape(B, 9).contiguous()
    t = g[:, :3, 3].contiguous()
    out = torch.empty_like(points)
    call("hrn_transform_points", ptr(points.contiguous()), ptr(R), ptr(t), ptr(out), B, N, stream())
    return out


def prepare_pairs(sweeps, max_range, num_points, twists, indices=None, generator=None):
    """sweeps: list of B raw clouds [N_i,3] (device).  Range filter (one launch for the batch) -> fixed-size resampling ->
    SE(3) perturbation of the source.  Returns (src [B,n,3] = igt . dst, dst [B,n,3], gt [B,4,4] = exp(-x), igt = exp(x))."""
    B = len(sweeps)
    dev = sweeps[0].device
    sizes = torch.tensor([0] + [int(s.shape[0]) for s in sweeps], dtype=torch.int64)
    offs = torch.cumsum(sizes, 0).to(dev)
    xyz, _, count = remove_points_by_range_batched(torch.cat(sweeps, 0), None, offs, max_range, int(sizes.max()))
    n = int(num_points)
    idx = torch.zeros(B, n, dtype=torch.int32, device=dev)
    if indices is None:                                  # draw on the device (needs the kept counts on the host)
        for b, c in enumerate(count.tolist()):
            if c <= n:
                idx[b, c:] = torch.randint(0, max(c, 1), (n - c,), device=dev, generator=generator, dtype=torch.int32)
            else:
                idx[b] = torch.randperm(c, device=dev, generator=generator)[:n].to(torch.int32)
    else:                                                # the caller's draws: pad picks [n - count] or subset [n]
        for b, ix in enumerate(indices):
            ix = ix.to(dev).to(torch.int32)
            idx[b, n - ix.numel():] = ix
    dst = torch.empty(B, n, 3, dtype=torch.float32, device=dev)
    call("hrn_resample_gather", ptr(xyz), ptr(offs), ptr(count), ptr(idx.contiguous()), B, n, ptr(dst), stream())
    igt = se3_exp(twists)
    gt = se3_exp(-twists)
    return apply_transform(dst, igt), dst, gt, igt
