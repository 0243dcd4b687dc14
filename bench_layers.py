#!/usr/bin/env python
"""Micro-benchmark of the per-layer tensor-core shared-MLP kernel (hrn_layer_tc) on the shapes of the headline workload.

    python bench_layers.py            # one JSON line per shape: us, useful TFLOP/s, max |err| against fp64 matmul

Shapes (32 pairs x 16,384 points): the coarse-level `convs_1` stack (65,536 rows, 528 -> 512 -> 512 ->
512, first layer over a gathered 3-segment virtual row), the level-3 descriptor `mlp1`/`mlp2` (262,144 rows, 768 -> 256
-> 256) and the per-keypoint confidence head (8,192 rows).  Timing: CUDA events, L2 flushed between iterations.
"""
import argparse
import json
import sys

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="", help="substring filter on the case name")
    ap.add_argument("--chains", action="store_true", help="time the fused chains instead of single layers")
    ap.add_argument("--bars", action="store_true",
                    help="library bars (BASELINE.md plan B4): cuBLAS GEMMs in fp32 / tf32 / bf16 on the conv-stack shapes and "
                         "torch.linalg.svd on the pose solve, next to this package's kernels on the same data")
    args = ap.parse_args()
    if args.chains:
        return chains(args)
    if args.bars:
        return bars(args)
    from pcd_reg_hregnet_b200 import engine, engine_tc
    from pcd_reg_hregnet_b200.engine import RowsView, SEG_BROADCAST, SEG_GATHER, ACT_RELU

    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(3)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def rnd(*s):
        return torch.randn(*s, device=dev, generator=g)

    cases = []
    # (name, rows, view builder, Cout)
    B, N1, k, C = 32, 256, 8, 256
    idx = torch.randint(0, N1, (B, N1, k), device=dev, generator=g, dtype=torch.int64)
    misc, S, D = rnd(B * N1 * k, 16), rnd(B * N1, C), rnd(B * N1, C)

    def coarse_view():
        v = RowsView(B * N1 * k, group=k, gather_idx=idx, rows_per_batch=N1 * k, src_rows_per_batch=N1)
        return v.add(misc).add(S, SEG_BROADCAST).add(D, SEG_GATHER)

    cases.append(("coarse convs_1[0] 528->512 (gathered)", coarse_view, 512))
    X512 = rnd(B * N1 * k, 512)
    cases.append(("coarse convs_1[1] 512->512", lambda: RowsView(B * N1 * k).add(X512), 512))
    X768 = rnd(262144, 768)
    cases.append(("L3 mlp1 768->256", lambda: RowsView(262144).add(X768), 256))
    X256 = rnd(262144, 256)
    cases.append(("L3 mlp2 256->256", lambda: RowsView(262144).add(X256), 256))
    X64 = rnd(1 << 20, 64)
    cases.append(("1M rows 64->64", lambda: RowsView(1 << 20).add(X64), 64))
    Xh = rnd(8192, 512)
    cases.append(("head 512->512 (8192 rows)", lambda: RowsView(8192).add(Xh), 512))
    cases.append(("head 512->1 (8192 rows)", lambda: RowsView(8192).add(Xh), 1))

    for name, mk, cout in cases:
        if args.only and args.only not in name:
            continue
        v = mk()
        K = sum(s[2] for s in v.segs)
        W = rnd(cout, K) / K ** 0.5
        b = rnd(cout)
        out = torch.empty(v.rows, cout, device=dev)
        engine_tc.layer_tc(v, W, b, ACT_RELU, out)
        torch.cuda.synchronize()
        # accuracy on a row sample against fp64
        if len(v.segs) == 1:
            X = v.segs[0][0]
            sel = torch.arange(0, v.rows, max(1, v.rows // 4096), device=dev)
            ref = torch.relu(X[sel].double() @ W.double().t() + b.double())
            err = (out[sel].double() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
        else:
            err = None
        ts = []
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engine_tc.layer_tc(v, W, b, ACT_RELU, out)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        print(json.dumps({"layer": name, "rows": v.rows, "K": K, "N": cout, "us": round(us, 1),
                          "useful_tflops": round(2.0 * v.rows * K * cout / us / 1e6, 1), "rel_err": err}))
        sys.stdout.flush()


def chains(args):
    """Fused 3-layer chains (hrn_chain_tc) on the shapes of the headline workload."""
    from pcd_reg_hregnet_b200 import engine_tc
    from pcd_reg_hregnet_b200.engine import RowsView, SEG_BROADCAST, SEG_GATHER, ACT_RELU

    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(5)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def rnd(*s):
        return torch.randn(*s, device=dev, generator=g)

    def grouped_view(B, N, k, C, geom):
        idx = torch.randint(0, N, (B, N, k), device=dev, generator=g, dtype=torch.int64)
        misc, S, D = rnd(B * N * k, geom), rnd(B * N, C), rnd(B * N, C)
        v = RowsView(B * N * k, group=k, gather_idx=idx, rows_per_batch=N * k, src_rows_per_batch=N)
        return v.add(misc).add(S, SEG_BROADCAST).add(D, SEG_GATHER)

    cases = [
        ("fine L2 convs_1 12+128+128 -> 256x3 attn k8", lambda: grouped_view(32, 512, 8, 128, 12), [256, 256, 256], engine_tc.EPI_ATTN, 8, False),
        ("fine L1 convs_1 12+64+64 -> 128x3 attn k8", lambda: grouped_view(32, 1024, 8, 64, 12), [128, 128, 128], engine_tc.EPI_ATTN, 8, False),
        ("L3 convs 132 -> 128,128,256 groupmax k16 +rows", lambda: RowsView(262144).add(rnd(262144, 132)), [128, 128, 256], engine_tc.EPI_GROUPMAX, 16, True),
        ("direct 272 -> 256x3 attn k8", lambda: RowsView(131072).add(rnd(131072, 272)), [256, 256, 256], engine_tc.EPI_ATTN, 8, False),
    ]
    for name, mk, widths, mode, kseg, want_rows in cases:
        if args.only and args.only not in name:
            continue
        v = mk()
        K = sum(s[2] for s in v.segs)
        layers, kin = [], K
        for w in widths:
            layers.append((rnd(w, kin) / kin ** 0.5, rnd(w) * 0.1, ACT_RELU))
            kin = w
        if not engine_tc.chain_supported(v, layers):
            print(json.dumps({"chain": name, "unsupported": True}))
            continue
        engine_tc.chain3(v, layers, mode, kseg, want_rows=want_rows)
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engine_tc.chain3(v, layers, mode, kseg, want_rows=want_rows)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        us = ts[len(ts) // 2]
        flops = 2.0 * v.rows * sum(a * b for a, b in zip([K] + widths[:-1], widths))
        print(json.dumps({"chain": name, "rows": v.rows, "K": K, "widths": widths, "us": round(us, 1),
                          "useful_tflops": round(flops / us / 1e6, 1)}))
        sys.stdout.flush()


def bars(args):
    """BASELINE.md plan B4: what the vendor libraries do with the same shapes on the same B200.  GEMM bars: torch.matmul
    (cuBLAS) on a materialised [rows, K] input (the reference materialises it; ours reads the virtual rows) in exact fp32,
    TF32 and bf16, next to hrn_layer_tc (bf16x3, fp32-class accuracy) -- with the error of each against fp64, because the
    1e-3 feature gate is what decides which of them may be used.  Pose bar: the reference's WeightedSVDHead arithmetic
    (weighted covariance + torch.linalg.svd + determinant fix, layers.py:469-504) next to hrn_weighted_kabsch."""
    from pcd_reg_hregnet_b200 import engine, engine_tc
    from pcd_reg_hregnet_b200.engine import RowsView, ACT_NONE

    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(9)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        return ts[len(ts) // 2]

    shapes = [("coarse convs_1 528->512", 65536, 528, 512), ("coarse convs_1 512->512", 65536, 512, 512),
              ("L3 mlp1 768->256", 262144, 768, 256), ("L3 convs 128->256", 262144, 128, 256),
              ("L1 convs 32->64", 4194304, 32, 64)]
    for name, rows, K, N in shapes:
        if args.only and args.only not in name:
            continue
        X = torch.randn(rows, K, device=dev, generator=g)
        W = torch.randn(N, K, device=dev, generator=g) / K ** 0.5
        sel = torch.arange(0, rows, max(1, rows // 2048), device=dev)
        ref = X[sel].double() @ W.double().t()
        scale = ref.abs().max().item()
        flop = 2.0 * rows * K * N
        out = torch.empty(rows, N, device=dev)
        res = {"bar": name, "rows": rows, "K": K, "N": N}

        def rec(tag, us, y):
            res[tag] = {"us": round(us, 1), "tflops": round(flop / us / 1e6, 1),
                        "rel_err": float((y[sel].double() - ref).abs().max().item() / scale)}

        torch.backends.cuda.matmul.allow_tf32 = False
        rec("cublas_fp32", timed(lambda: torch.matmul(X, W.t(), out=out)), out)
        torch.backends.cuda.matmul.allow_tf32 = True
        rec("cublas_tf32", timed(lambda: torch.matmul(X, W.t(), out=out)), out)
        torch.backends.cuda.matmul.allow_tf32 = False
        Xb, Wb = X.bfloat16(), W.bfloat16()
        outb = torch.empty(rows, N, device=dev, dtype=torch.bfloat16)
        rec("cublas_bf16 (inputs already bf16)", timed(lambda: torch.matmul(Xb, Wb.t(), out=outb)), outb)
        if K % 4 == 0:
            bias = torch.zeros(N, device=dev)
            v = RowsView(rows).add(X)
            rec("hrn_layer_tc (bf16x3)", timed(lambda: engine_tc.layer_tc(v, W, bias, ACT_NONE, out)), out)
        print(json.dumps(res))
        sys.stdout.flush()
    # pose solve: 32 pairs x 256 / 512 / 1024 correspondences
    for n in (256, 512, 1024):
        src = torch.randn(32, n, 3, device=dev, generator=g) * 20
        cor = src + 0.05 * torch.randn(32, n, 3, device=dev, generator=g)
        w = torch.rand(32, n, device=dev, generator=g)

        def ref_head():
            ww = w / (w.sum(1, keepdim=True) + 1e-4)
            sc = (src * ww[..., None]).sum(1, keepdim=True)
            cc = (cor * ww[..., None]).sum(1, keepdim=True)
            H = ((src - sc) * ww[..., None]).transpose(1, 2) @ (cor - cc)
            U, _, Vh = torch.linalg.svd(H)
            V = Vh.transpose(1, 2)
            d = torch.det(V @ U.transpose(1, 2))
            D = torch.diag_embed(torch.stack([torch.ones_like(d), torch.ones_like(d), d], 1))
            R = V @ D @ U.transpose(1, 2)
            return R, cc.squeeze(1) - (R @ sc.transpose(1, 2)).squeeze(2)

        print(json.dumps({"bar": f"pose solve, 32 pairs x {n} correspondences",
                          "torch (weighted covariance + linalg.svd + det fix, eager)": {"us": round(timed(ref_head), 1)},
                          "hrn_weighted_kabsch": {"us": round(timed(lambda: engine.weighted_kabsch(src, cor, w)), 1)}}))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
