#!/usr/bin/env python
"""PointUtils micro-benchmark sweep (BASELINE configs[4]): FPS / kNN / grouping-gather at N = 4k..131k points,
CUDA-event timed through the public ops, optionally next to the recompiled reference kernels (oracle/_ref)."""
import argparse
import importlib.util
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from pcd_reg_hregnet_b200 import engine, ops  # noqa: E402


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps


def ref_ext():
    so = os.path.join(ROOT, "oracle", "_ref", "point_utils_cuda.so")
    if not os.path.exists(so):
        return None
    spec = importlib.util.spec_from_file_location("point_utils_cuda", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def around_the_path():
    """SURVEY 8(f) rows 1-2 on the device: range filter + resampling + perturbation of 32 raw 120k-point sweeps into the
    [32,16384,3] pair batch, and the pose-error metrics of a 256-pair batch.  HBM-bound kernels: GB/s of algorithmic bytes."""
    from pcd_reg_hregnet_b200 import metrics, preprocess
    B, Nraw, n = 32, 120000, 16384
    g = torch.Generator(device="cuda").manual_seed(0)
    sweeps = [torch.randn(Nraw, 3, device="cuda", generator=g) * 40 for _ in range(B)]
    cat = torch.cat(sweeps, 0)
    offs = torch.arange(0, (B + 1) * Nraw, Nraw, device="cuda", dtype=torch.int64)
    tw = torch.randn(B, 6, device="cuda", generator=g) * 0.1
    t = timeit(lambda: preprocess.remove_points_by_range_batched(cat, None, offs, 80.0, Nraw))
    _, _, cnt = preprocess.remove_points_by_range_batched(cat, None, offs, 80.0, Nraw)
    kept = int(cnt.sum())
    r = {"op": "range_filter", "sweeps": B, "points": Nraw, "kept": kept, "ms": t,
         "GBps": (B * Nraw * 12 + kept * 12) / t / 1e6}
    print(json.dumps(r), flush=True)
    t = timeit(lambda: preprocess.prepare_pairs(sweeps, 80.0, n, tw))
    print(json.dumps({"op": "prepare_pairs (filter + resample + SE3 perturbation, incl. host glue)", "pairs": B, "ms": t}), flush=True)
    R = torch.linalg.qr(torch.randn(256, 3, 3, device="cuda", generator=g))[0]
    tt = torch.randn(256, 3, device="cuda", generator=g)
    t = timeit(lambda: metrics.pose_errors(R, tt, R.flip(0), tt.flip(0)))
    print(json.dumps({"op": "pose_errors", "pairs": 256, "ms": t}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="4096,8192,16384,32768,65536,131072")
    ap.add_argument("--batches", default="1,32,64")
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--around", action="store_true", help="time the steps around the path (input pipeline, metrics) instead")
    a = ap.parse_args()
    if a.around:
        return around_the_path()
    ref = ref_ext()
    rows = []
    for B in [int(b) for b in a.batches.split(",")]:
        for N in [int(n) for n in a.sizes.split(",")]:
            if B * N > 64 * 65536:
                continue
            xyz = torch.rand(B, N, 3, device="cuda") * 100
            M = 1024
            t_fps = timeit(lambda: ops.furthest_point_sample(xyz, M))
            r = {"op": "fps", "B": B, "N": N, "M": M, "ms": t_fps, "us_per_iter": 1e3 * t_fps / (M - 1)}
            if ref is not None and N <= 65536:
                temp = torch.empty(B, N, device="cuda")
                out = torch.empty(B, M, dtype=torch.int32, device="cuda")

                def run_ref():
                    temp.fill_(1e10)
                    ref.furthest_point_sampling_wrapper(B, N, M, xyz, temp, out)
                r["ref_kernel_ms"] = timeit(run_ref, reps=2, warm=1)
            rows.append(r)
            print(json.dumps(r), flush=True)
            if a.quick:
                continue
            idx = ops.furthest_point_sample(xyz, M)
            q = xyz[torch.arange(B, device="cuda")[:, None], idx.long()]
            for K in (16, 32, 64):
                t = timeit(lambda: engine.knn_idx(None, xyz, K, q_idx=idx))
                r = {"op": "knn", "B": B, "N": N, "M": M, "K": K, "ms": t, "Gpair_per_s": B * M * N / t / 1e6}
                if B * M * N <= (1 << 30):
                    # "reference GPU path" bar: pytorch3d is not installable offline, so the library route to
                    # the same result -- torch.cdist + topk (distances not bit-identical: cdist uses the GEMM expansion)
                    r["torch_cdist_topk_ms"] = timeit(lambda: torch.cdist(q, xyz).topk(K, dim=2, largest=False), reps=2, warm=1)
                rows.append(r)
                print(json.dumps(r), flush=True)
    return rows


def knnd_bench():
    """CoarseReg descriptor-space search: 256 queries x 256 refs x 256 dims, K=8, 32 pairs; and the level-1 search."""
    p1 = torch.rand(32, 256, 256, device="cuda")
    p2 = torch.rand(32, 256, 256, device="cuda")
    t = timeit(lambda: engine.knn_idx(p1, p2, 8), reps=20)
    print(json.dumps({"op": "knn_desc", "B": 32, "M": 256, "N": 256, "D": 256, "K": 8, "ms": t}), flush=True)
    xyz = torch.rand(64, 16384, 3, device="cuda") * 100
    idx = ops.furthest_point_sample(xyz, 1024)
    for flag in (True, False):
        engine._SORTED_KNN = flag
        t = timeit(lambda: engine.knn_idx(None, xyz, 64, q_idx=idx))
        print(json.dumps({"op": "knn_l1", "sorted": flag, "ms": t}), flush=True)
    engine._SORTED_KNN = True


if __name__ == "__main__":
    if "--knnd" in sys.argv:
        knnd_bench()
    else:
        main()
