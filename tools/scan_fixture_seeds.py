"""Fixture selection (test infrastructure, runs on the GPU box): which seeded synthetic pairs keep their level-2 / level-3
keypoint sets between the oracle (== the reference's feature extraction, bit for bit) and the product in every exact
precision mode?  Weighted FPS consumes network outputs, so a pair either reproduces the reference's picks or diverges
(SURVEY.md section 7); free-running golden tests need pairs of the first kind.

    python tools/scan_fixture_seeds.py [first_seed] [count] [n_points]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_layers as RL  # noqa: E402
from pcd_reg_hregnet_b200 import engine, synth  # noqa: E402


def main():
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
    modes = [m for m in ("fp32", "tc", "tc2", "tc1") if m in getattr(engine, "PRECISIONS", ("fp32", "tc"))]
    cpu = synth.build_net("hregnet", 7)
    gpu = synth.build_net("hregnet", 7, "cuda")
    sd = cpu.state_dict()
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    for seed in range(first, first + count):
        src, dst, _, _ = synth.make_batch([seed], n)
        with torch.no_grad():
            want = {s: RL.hier_feature_extraction(sd, "feature_extraction.", x) for s, x in (("src", src), ("dst", dst))}
            row = {}
            for m in modes:
                engine.set_precision(m)
                out = gpu(src.cuda(), dst.cuda())
                worst = 0.0
                for s in ("src", "dst"):
                    for lv in (2, 3):
                        a, b = out[f"{s}_feats"][f"xyz_{lv}"].cpu().double(), want[s][f"xyz_{lv}"].double()
                        worst = max(worst, float((a - b).abs().max() / b.abs().max()))
                row[m] = worst
        engine.set_precision("tc")
        print(seed, " ".join(f"{m}={v:.1e}{'*' if v < 1e-4 else ' '}" for m, v in row.items()), flush=True)


if __name__ == "__main__":
    main()
