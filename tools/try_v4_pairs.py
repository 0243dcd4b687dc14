"""Fixture selection, third stage (GPU box): candidate Model_V4 fixtures (tests/golden/tmp_v4/*.npz, reference outputs) through
the same checks as tests/test_gpu_model.py::test_golden_end_to_end_model_variants, in every precision mode."""
import glob, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from pcd_reg_hregnet_b200 import engine
from common import build_product_model_v4

def rel(a, b):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))

for f in sorted(glob.glob(os.path.join(ROOT, "tests/golden/tmp_v4/*.npz"))):
    gd = {k: torch.from_numpy(v) for k, v in np.load(f).items()}
    row = []
    for m in engine.PRECISIONS:
        engine.set_precision(m)
        net = build_product_model_v4(seed=7, device="cuda")
        with torch.no_grad():
            torch.manual_seed(0)
            out = net(gd["src"].cuda(), gd["dst"].cuda())
        B = gd["src"].shape[0]
        ok = [b for b in range(B) if rel(out["dst_xyz_2"][b].cpu(), gd["dst_xyz_2"][b]) < 1e-4 and rel(out["src_feats_sigmas_2"][b].cpu(), gd["src_feats_sigmas_2"][b]) < 1e-3]
        full = len(ok) == B and rel(out["src_xyz_2_trans"].cpu(), gd["src_xyz_2_trans"]) < 1e-4
        # coord_dist / feats_dist per keypoint as SETS (rows sorted by coord_dist), like the test does in the tensor-core modes
        ocd, oix = out["coord_dist"].cpu().sort(-1)
        gcd, gix = gd["coord_dist"].sort(-1)
        ofd, gfd = out["feats_dist"].cpu().gather(-1, oix), gd["feats_dist"].gather(-1, gix)
        bad = int(((ocd - gcd).abs().amax(-1) > 1e-3 * gcd.abs().max()).sum())
        parts = ([rel(out[k].cpu(), gd[k]) for k in ("src_dst_feats_2", "src_dst_feats_2_prime")] + [rel(ocd, gcd)] +
                 [float((out[k].cpu() - gd[k]).abs().max()) for k in ("src_dst_weights_2", "src_dst_weights_2_prime")] +
                 [float((ofd - gfd).abs().max())]) if full else [float("nan")]
        row.append(f"{m}: ok={len(ok)}/{B} full={int(full)} [feats, feats', coord, w, w', fdist]=" + ",".join(f"{v:.1e}" for v in parts) + f" kp_sets_differ={bad}")
    engine.set_precision("tc")
    print(os.path.basename(f), " | ".join(row), flush=True)
