"""GPU: input-pipeline kernels (csrc/preprocess.cu through pcd_reg_hregnet_b200.preprocess) against the golden reference
outputs (tests/golden/preprocess.npz) and the oracle (oracle/ref_preprocess.py).  Range filter and resampling are
bit-exact; SE3.exp within 1e-6."""
import numpy as np
import pytest
import torch

from common import load_golden
from oracle import ref_preprocess as RP
from pcd_reg_hregnet_b200 import preprocess as P

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_reference_golden():
    g = load_golden("preprocess")
    for n in ("a", "b"):
        fp, fi = P.remove_points_by_range(g[f"{n}_pc"].to(DEV), g[f"{n}_int"].to(DEV), 60.0)
        assert torch.equal(fp.cpu(), g[f"{n}_filtered"]) and torch.equal(fi.cpu(), g[f"{n}_filtered_int"])
        rp, ri = P.PointCloudResampler(4096)(fp, fi, indices=g[f"{n}_idx"].to(DEV))
        assert torch.equal(rp.cpu(), g[f"{n}_resampled"]) and torch.equal(ri.cpu(), g[f"{n}_resampled_int"])
    ge = P.se3_exp(g["twist"].to(DEV))
    torch.cuda.synchronize()
    assert float((ge.cpu() - g["se3_exp"]).abs().max()) < 1e-6


def test_range_filter_batched_edge_cases():
    """Ragged batch in one launch: empty sweep, everything kept, nothing kept, sizes around the 1024-point chunk, points
    exactly on the range boundary (strict <), 100k-point sweep."""
    rng = np.random.default_rng(3)
    sizes = [0, 1, 1023, 1024, 1025, 5000, 100000, 7]
    sweeps = [(rng.normal(size=(n, 3)) * 30).astype(np.float32) for n in sizes]
    sweeps[2] *= 0.01                     # all inside
    sweeps[3] += 1000.0                   # all outside
    sweeps[7][:, :] = 0.0
    sweeps[7][:, 0] = np.float32(50.0)    # ||p|| == max_range exactly -> dropped
    sweeps[7][3, 0] = np.nextafter(np.float32(50.0), np.float32(0.0))
    offs = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int64, device=DEV)
    xyz = torch.from_numpy(np.concatenate(sweeps, 0)).to(DEV)
    inten = torch.arange(xyz.shape[0], dtype=torch.float32, device=DEV)
    out, iout, count = P.remove_points_by_range_batched(xyz, inten, offs, 50.0)
    torch.cuda.synchronize()
    o = offs.tolist()
    for i, sw in enumerate(sweeps):
        want, wi = RP.remove_points_by_range(sw, np.arange(o[i], o[i + 1], dtype=np.float32), 50.0)
        c = int(count[i])
        assert c == want.shape[0], (i, c, want.shape[0])
        assert np.array_equal(out[o[i]:o[i] + c].cpu().numpy(), want)
        assert np.array_equal(iout[o[i]:o[i] + c].cpu().numpy(), wi)
    assert int(count[7]) == 1


def test_prepare_pairs_matches_oracle():
    rng = np.random.default_rng(9)
    B, n = 3, 2048
    raw = [(rng.normal(size=(m, 3)) * np.array([30.0, 30.0, 2.0])).astype(np.float32) for m in (5000, 1800, 2600)]
    tw = (rng.normal(size=(B, 6)) * 0.1).astype(np.float32)
    idx = []
    for r in raw:
        m = RP.remove_points_by_range(r, None, 45.0)[0].shape[0]
        idx.append(rng.integers(0, m, n - m) if m <= n else rng.permutation(m)[:n])
    src, dst, gt, igt = P.prepare_pairs([torch.from_numpy(r).to(DEV) for r in raw], 45.0, n, torch.from_numpy(tw).to(DEV),
                                        indices=[torch.from_numpy(np.asarray(i)).to(DEV) for i in idx])
    torch.cuda.synchronize()
    G = RP.se3_exp(tw)
    for b in range(B):
        f = RP.remove_points_by_range(raw[b], None, 45.0)[0]
        d = RP.resample(f, None, n, idx[b])[0]
        assert np.array_equal(dst[b].cpu().numpy(), d)
        s = d.astype(np.float64) @ G[b, :3, :3].T + G[b, :3, 3]
        assert np.abs(src[b].cpu().numpy() - s).max() < 2e-5
    assert float((igt.cpu().double() - torch.from_numpy(G)).abs().max()) < 1e-6
    assert float((gt.cpu().double() @ igt.cpu().double() - torch.eye(4, dtype=torch.float64)).abs().max()) < 1e-5
