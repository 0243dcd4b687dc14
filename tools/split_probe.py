"""Probe: does a 32-pair step get faster as two 16-pair half-batches replayed on two CUDA streams?

The sampling kernels (FPS) are latency-bound with one small cluster per cloud and leave most SMs idle, the shared-MLP
kernels are persistent and fill the chip: two half-batch graphs in flight let one half's sampling run beside the other
half's tensor-core work.  Prints ms per 32 pairs for: one 32-pair graph; two 16-pair graphs back to back on one
stream; two 16-pair graphs on two streams (optionally the second delayed by `--skew` of a half-step).

    python tools/split_probe.py [--steps 20]
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--pairs", type=int, default=32)
    ap.add_argument("--points", type=int, default=16384)
    ap.add_argument("--parts", type=int, default=2)
    args = ap.parse_args()
    from common import build_product_hregnet
    from pcd_reg_hregnet_b200 import engine, synth
    from pcd_reg_hregnet_b200.runner import Registrar

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    engine.set_precision("tc")
    B, N, P = args.pairs, args.points, args.parts
    src_h, dst_h, _, _ = synth.make_batch(range(1000, 1000 + B), N)
    net = build_product_hregnet(seed=7, device=dev)
    whole = Registrar(net, B, N)
    whole.load(src_h, dst_h)
    whole.capture()
    h = B // P
    parts = []
    for i in range(P):
        r = Registrar(net, h, N)
        r.load(src_h[i * h:(i + 1) * h], dst_h[i * h:(i + 1) * h])
        r.capture()
        parts.append(r)
    torch.cuda.synchronize()
    # same poses from the parts as from the whole batch
    ow = whole.run_device()
    Rw = ow["rotation"][-1].clone()
    Rp = torch.cat([r.run_device()["rotation"][-1] for r in parts])
    torch.cuda.synchronize()
    print("max |R_whole - R_parts| =", float((Rw - Rp).abs().max()))

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    streams = [torch.cuda.Stream(device=dev) for _ in range(P)]
    main_s = torch.cuda.current_stream(dev)

    def run_whole():
        whole.run_device()

    def run_serial():
        for r in parts:
            r.run_device()

    def run_parallel():
        for s, r in zip(streams, parts):
            s.wait_stream(main_s)
            with torch.cuda.stream(s):
                r.run_device()
        for s in streams:
            main_s.wait_stream(s)

    def run_staggered(delay_ms):
        def fn():
            for i, (s, r) in enumerate(zip(streams, parts)):
                s.wait_stream(main_s)
                with torch.cuda.stream(s):
                    if i and delay_ms > 0:
                        torch.cuda._sleep(int(i * delay_ms * 1.965e6))       # one spinning thread: part i starts late
                    r.run_device()
            for s in streams:
                main_s.wait_stream(s)
        return fn

    stag = [("%d graphs, staggered %.1f ms" % (P, d), run_staggered(d)) for d in (0.4, 0.8, 1.0, 1.3, 1.7, 2.2)]
    for name, fn in (("one graph, %d pairs" % B, run_whole), ("%d graphs, one stream" % P, run_serial),
                     ("%d graphs, %d streams" % (P, P), run_parallel), *stag, ("one graph again", run_whole)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for s, e in ev:
            flush.fill_(1)
            s.record()
            fn()
            e.record()
        torch.cuda.synchronize()
        ms = sorted(s.elapsed_time(e) for s, e in ev)
        print(f"{name:28s}: median {ms[len(ms) // 2]:.3f} ms  min {ms[0]:.3f}  max {ms[-1]:.3f}  per {B} pairs")


def pipelined(B=32, depths=(1, 2, 3)):
    """Throughput form: whole batches alternate between two graphs on two streams (consecutive batches overlap)."""
    from common import build_product_hregnet
    from pcd_reg_hregnet_b200 import engine, synth
    from pcd_reg_hregnet_b200.runner import Registrar

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    engine.set_precision("tc")
    N = 16384
    src_h, dst_h, _, _ = synth.make_batch(range(1000, 1000 + B), N)
    net = build_product_hregnet(seed=7, device=dev)
    for depth in depths:
        regs = []
        for i in range(depth):
            r = Registrar(net, B, N)
            r.load(src_h, dst_h)
            r.capture()
            regs.append(r)
        streams = [torch.cuda.Stream(device=dev) for _ in range(depth)]
        torch.cuda.synchronize()
        for K in (24, 48):
            for rep in range(2):
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                main_s = torch.cuda.current_stream(dev)
                t0.record()
                for s in streams:
                    s.wait_stream(main_s)
                for k in range(K):
                    with torch.cuda.stream(streams[k % depth]):
                        regs[k % depth].run_device()
                for s in streams:
                    main_s.wait_stream(s)
                t1.record()
                torch.cuda.synchronize()
            print(f"depth {depth}: {K} batches of {B} pairs back to back: {t0.elapsed_time(t1) / K:.3f} ms per batch = "
                  f"{t0.elapsed_time(t1) / K * 32 / B:.3f} ms per 32 pairs")
        del regs


if __name__ == "__main__":
    if "--pipelined" in sys.argv:          # --pipelined [pairs per batch] [depths, comma-separated]
        i = sys.argv.index("--pipelined")
        rest = sys.argv[i + 1:]
        pipelined(int(rest[0]) if rest else 32, tuple(int(x) for x in rest[1].split(",")) if len(rest) > 1 else (1, 2, 3))
        sys.exit(0)
    main()
