#!/usr/bin/env python
"""bench.py -- HRegNet registration forward, pairs/sec (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--pairs-per-gpu 32] [--points 16384]

A "step" = one pass of the hot path (HRegNet.forward: FPS -> kNN -> grouping -> detector/descriptor MLPs ->
coarse + fine correspondence -> weighted-SVD poses) over one batch of synthetic 16,384-point LiDAR pairs per GPU
(BASELINE configs[1]: batch 32, keypoints 1024/512/256), followed -- for N > 1 -- by the NCCL all-gather of poses.
Weak scaling: 32 pairs per GPU (N=8 -> the 256-pair job of configs[2]).

  value     pairs/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e       pairs/s through the public API (pcd_reg_hregnet_b200.runner.Registrar.map) from pinned HOST buffers:
            per batch the H2D copy of both clouds + forward + D2H of the poses, all inside the timed region (the
            copies of neighbouring batches overlap the forward on a copy stream, and two forwards are in flight on
            two streams)
  in_flight pairs/s of the same K steps from device-resident inputs with map()'s two forwards in flight (rotating
            input batches larger than L2 instead of a flush) -- what e2e is bounded by; `value` stays one forward at a time
  roofline  dominant kernel family (shared-MLP layers): algorithmic FLOP / CUDA-event time of those launches
  cpu_baseline  the reference's CPU path on the box's host cores, bounded sample (rank 0, N=1 only): the UNMODIFIED
            reference graph (baseline/_ref, staged by __graft_entry__.build()) run batched on the CPU through
            oracle/ref_harness.py, native ops = the oracle's C restatement (kind "reference"); where the staged reference
            is absent, the oracle port of the graph (kind "port")
  gpu_reference the UNMODIFIED reference graph on the same B200 (its own CUDA extension recompiled for sm_100a in
            oracle/_ref + torch cdist/topk standing in for the uninstallable pytorch3d), CUDA-event timed
  parity    free-running index agreement per level and pose deltas of the first pairs of the bench batch vs the oracle
  --impl reference   times that CPU implementation as the reference arm (rank 0 only)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "registration pairs/sec (16k-pt HRegNet fwd)"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=32)
    ap.add_argument("--points", type=int, default=16384)
    ap.add_argument("--precision", default=os.environ.get("HRN_PRECISION", "auto"))
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--cpu-sample-pairs", type=int, default=4, help="pairs per batched CPU reference forward")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--total-pairs", type=int, default=256,
                    help="N > 1: also time the strong-scaling job of BASELINE configs[2] (this many pairs sharded over the GPUs)")
    ap.add_argument("--model", default="hregnet", choices=["hregnet", "v2", "v4"],
                    help="hregnet = BASELINE configs[1] (default); v2 = Adaption-1 / Model_V2 (configs[3], use --points 32768); "
                         "v4 = Model_V4 (Model_V2 + coord_dist / feats_dist of the coarse stage)")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_tflops_sustained=d["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 9 and r[1].isdigit())
        mx = max([int(r[2]) for r in self.rows if len(r) >= 9 and r[2].isdigit()] or [0])
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(sm)}


def workload_config(args, world):
    """`config` of the JSON line -- the same dict for both arms (the reference arm runs a bounded sample of it)."""
    name = dict(hregnet="HRegNet baseline", v2="Adaption-1 (Model_V2)", v4="Model_V4")[args.model]
    return {"workload": f"{name} forward, batch {args.pairs_per_gpu} synthetic {args.points}-pt pairs per GPU (keypoints 1024/512/256)",
            "pairs_per_gpu": args.pairs_per_gpu, "points": args.points,
            "parallelism": f"pairs sharded over {world} GPU(s), pose all-gather"}


def cpu_reference_forward(n_points, batch):
    """-> (callable running ONE batched forward of `batch` pairs on the CPU, kind, description).
    kind "reference": the unmodified reference graph (oracle/ref_harness.py; /root/reference or the staged baseline/_ref)
    with the oracle's C restatement of the CUDA-only native ops underneath; kind "port": the oracle's restatement of
    the graph too (only where the reference files are absent)."""
    from oracle import ref_harness as H
    from pcd_reg_hregnet_b200 import synth
    src, dst, _, _ = synth.make_batch(range(1000, 1000 + batch), n_points)
    if H.available():
        net = H.build_reference_hregnet(seed=7)
        return (lambda: net(src, dst)), "reference", ("unmodified reference graph (models/HRegNet/models.py:77-148) on the CPU, "
                                                      "native ops = oracle C restatement, torch-CPU for the layers")
    from oracle import ref_layers as RL
    sd = synth.build_net("hregnet", 7).state_dict()
    return (lambda: RL.hregnet_forward(sd, src, dst)), "port", "oracle port of the reference graph (torch-CPU over the C native ops)"


def run_reference(args, rank):
    if rank != 0:
        return
    steps, warm = args.steps, args.warmup
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    per_step = max(1, args.cpu_sample_pairs)                           # bounded sample: one batched forward per "step"
    fwd, kind, what = cpu_reference_forward(args.points, per_step)
    with torch.no_grad():
        for _ in range(max(1, min(warm, 2))):
            fwd()
        t0 = time.perf_counter()
        for _ in range(steps):
            fwd()
        dt = time.perf_counter() - t0
    v = steps * per_step / dt
    sample = f"{per_step} pairs x {args.points} pts per step as ONE batched forward (B={per_step}), {steps} steps, {dt:.1f} s"
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "implementation": what + f"; all {cores} host cores",
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def cpu_baseline_subprocess(args):
    """The CPU leg in its own process (the CPU harness redirects torch's .cuda() entry points: it must not share a
    process with the GPU arm).  Bounded: 1 warm-up + 3 batched forwards."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "3", "--warmup", "1",
           "--points", str(args.points), "--cpu-sample-pairs", str(args.cpu_sample_pairs)]
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
        return json.loads(line)["cpu_baseline"]
    except Exception as e:  # the baseline is a reported number, never a reason to lose the bench line
        return {"value": None, "unit": UNIT, "cores": len(os.sched_getaffinity(0)), "kind": "unavailable", "sample": repr(e)[:200]}


def gpu_reference(args, dev):
    """The UNMODIFIED reference graph on this GPU: its own CUDA extension recompiled for sm_100a
    (oracle/_ref/point_utils_cuda.so) + torch.cdist/topk for the uninstallable pytorch3d, same batch, CUDA events."""
    try:
        from oracle import ref_harness as H
        from pcd_reg_hregnet_b200 import synth
        if not H.available():
            return {"unavailable": "reference files not staged (baseline/_ref)"}
        net = H.build_reference_hregnet(seed=7, device="cuda")
        B, N = args.pairs_per_gpu, args.points
        src, dst, _, _ = synth.make_batch(range(1000, 1000 + B), N)
        src, dst = src.to(dev), dst.to(dev)
        with torch.no_grad():
            net(src, dst)
            torch.cuda.synchronize()
            ts = []
            for _ in range(3):
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record(); net(src, dst); e.record()
                torch.cuda.synchronize()
                ts.append(s.elapsed_time(e))
        ms = sorted(ts)[1]
        del net
        torch.cuda.empty_cache()
        return {"value": B / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "batch": B,
                "what": "unmodified reference HRegNet.forward on the same B200: reference PointUtils kernels recompiled for sm_100a "
                        "(oracle/_ref), torch.cdist+topk in place of pytorch3d.knn_points, eager PyTorch 2.11 (cuDNN/cuBLAS/cuSOLVER)"}
    except Exception as e:
        return {"unavailable": repr(e)[:300]}


def parity_report(out, src_h, dst_h, n_pairs=2):
    """Free-running parity of the bench batch's first pairs against the oracle (CPU, seconds): index-agreement rate of
    the keypoint sets per level and the pose deltas.  Weighted FPS makes levels 2/3 chaotic: the
    stage-wise gates are the teacher-forced tests; this is the reported free-running picture."""
    from oracle import ref_layers as RL
    from pcd_reg_hregnet_b200 import synth
    sd = synth.build_net("hregnet", 7).state_dict()
    torch.set_num_threads(len(os.sched_getaffinity(0)))
    rep = {"pairs": n_pairs, "keypoints_equal": {}, "pose": []}
    with torch.no_grad():
        want = RL.hregnet_forward(sd, src_h[:n_pairs].clone(), dst_h[:n_pairs].clone())
    for lv in (1, 2, 3):
        same = []
        for s in ("src", "dst"):
            a = out[f"{s}_feats"][f"xyz_{lv}"][:n_pairs].cpu().double()
            b = want[f"{s}_feats"][f"xyz_{lv}"].double()
            same.append(float(((a - b).abs().amax(dim=2) < 1e-3).double().mean()))
        rep["keypoints_equal"][f"level_{lv}"] = sum(same) / len(same)
    d1 = out["src_feats"]["desc_1"][:n_pairs].cpu().double()
    w1 = want["src_feats"]["desc_1"].double()
    rep["desc_1_rel_err"] = float((d1 - w1).abs().max() / w1.abs().max())
    for lv in range(3):
        ang = RL.rotation_angle_deg(out["rotation"][lv][:n_pairs].cpu(), want["rotation"][lv])
        dt = (out["translation"][lv][:n_pairs].cpu() - want["translation"][lv]).abs().amax(dim=1)
        rep["pose"].append({"level": 3 - lv, "max_deg": float(ang.max()), "max_m": float(dt.max())})
    return rep


_REAL_STDOUT = None


def emit(line: str):
    """The ONE JSON line goes to the real stdout; everything else (NCCL banners, library chatter) was re-routed to
    stderr at the file-descriptor level by main()."""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (line + "\n").encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from pcd_reg_hregnet_b200 import _lib, dist as hdist, engine, synth
    from pcd_reg_hregnet_b200.runner import Registrar

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    precision = args.precision
    if precision == "auto":
        # bf16x3 everywhere.  "tcf" (single-pass fp16 correspondence stages) also meets every stage gate, but with a 1.4x
        # margin on the coarse correspondences and 2.5 % of step time to gain (profiles/r02b_bench_1gpu_tcf.json): opt-in
        precision = "tc"
    engine.set_precision(precision)

    B, N = args.pairs_per_gpu, args.points
    lo = rank * B
    src_h, dst_h, _, _ = synth.make_batch(range(1000 + lo, 1000 + lo + B), N)
    src_h, dst_h = src_h.pin_memory(), dst_h.pin_memory()
    net = synth.build_net(args.model, seed=7, device=dev)
    reg = Registrar(net, B, N, use_cuda_graph=not args.no_graph, in_flight=int(os.environ.get("HRN_IN_FLIGHT", "2")))
    reg.load(src_h, dst_h)
    reg.capture()

    # kernel launch accounting (claims): count C-ABI calls of one eager forward
    launches = _count_launches(reg)

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def step():
        out = reg.run_device()
        if world > 1:
            hdist.gather_poses(out["rotation"][-1], out["translation"][-1])
        return out

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.synchronize()
    for s, e in ev:
        flush.fill_(1)                                                  # L2 flush between timed iterations (untimed)
        s.record()
        step()
        e.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_ms = sum(s.elapsed_time(e) for s, e in ev)

    # ---- end to end through the public API, host buffers -------------------------------------------------------
    post = (lambda out: hdist.gather_poses(out["rotation"][-1], out["translation"][-1])) if world > 1 else None
    for _ in range(2):
        reg(src_h, dst_h)
    for _ in reg.map(((src_h, dst_h) for _ in range(4)), post=post):      # untimed: captures the second lane, warms the pipeline
        pass
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # throughput form of the public API: every batch is copied H2D from pinned memory and its poses D2H inside the
    # timed region; Registrar.map only takes the copies of batch i+1 / i-1 off the critical path of batch i
    t0 = time.perf_counter()
    n_done = 0
    for R_h, t_h in reg.map(((src_h, dst_h) for _ in range(args.steps)), post=post):
        n_done += 1
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    assert n_done == args.steps

    # ---- device-resident throughput with the forwards of map() in flight (explains e2e > value) -----------------
    # `value` above times one forward at a time with an L2 flush in front of each; map() keeps two in flight.  Here
    # the same K steps run back to back on map()'s lanes from 16 rotating device-resident input batches (201 MB at the
    # default size: larger than the 126 MB L2, so no flush), each step copying its batch into the lane's input buffers.
    lanes = reg._pipe["lanes"]
    n_rot = 16
    rot = [(reg.src.clone(), reg.dst.clone()) for _ in range(n_rot)]
    main_s = torch.cuda.current_stream(dev)

    def pipelined(K):
        for ln in lanes:
            if ln["stream"] is not None:
                ln["stream"].wait_stream(main_s)
        for k in range(K):
            ln = lanes[k % len(lanes)]
            with torch.cuda.stream(ln["stream"] if ln["stream"] is not None else main_s):
                r = ln["reg"] if ln["reg"] is not None else reg
                r.src.copy_(rot[k % n_rot][0], non_blocking=True)
                r.dst.copy_(rot[k % n_rot][1], non_blocking=True)
                out = r.run_device()
                if post is not None:
                    post(out)
        for ln in lanes:
            if ln["stream"] is not None:
                main_s.wait_stream(ln["stream"])

    pipelined(4)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    pipelined(args.steps)
    p1.record()
    torch.cuda.synchronize()
    pipe_ms = p0.elapsed_time(p1)
    del rot
    clocks = sampler.stop()

    tt = torch.tensor([t_ms, e2e_ms, pipe_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_ms, e2e_ms, pipe_ms = tt.tolist()

    # ---- per-kernel-family breakdown + roofline of the dominant family (rank 0) ---------------------------------
    prof = _profile_families(reg, steps=min(args.steps, 3)) if rank == 0 else None
    parity = None
    if rank == 0 and world == 1 and args.model == "hregnet" and not args.no_parity:
        try:
            parity = parity_report(reg.run_device(), src_h, dst_h)
        except Exception as e:  # reported picture, never a reason to lose the bench line
            parity = {"unavailable": repr(e)[:200]}

    # ---- strong scaling (BASELINE configs[2] as worded: a fixed job of --total-pairs pairs sharded over the GPUs) ----
    strong = None
    if world > 1 and args.total_pairs >= world and args.model == "hregnet":
        Bs = args.total_pairs // world
        k_s = max(2, min(args.steps, 5))
        s_src, s_dst, _, _ = synth.make_batch([1000 + (rank * Bs + i) % 64 for i in range(Bs)], N)   # 64 distinct scenes
        if Bs == B:
            reg_s = reg
        else:
            reg_s = Registrar(net, Bs, N, use_cuda_graph=not args.no_graph, in_flight=1)
            reg_s.load(s_src.pin_memory(), s_dst.pin_memory())
            reg_s.capture()

        def step_s():
            o = reg_s.run_device()
            hdist.gather_poses(o["rotation"][-1], o["translation"][-1])

        for _ in range(3):
            step_s()
        torch.cuda.synchronize()
        dist.barrier()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k_s)]
        for s0, e0 in evs:
            flush.fill_(1)
            s0.record()
            step_s()
            e0.record()
        torch.cuda.synchronize()
        dist.barrier()
        ts = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], device=dev, dtype=torch.float64)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        strong = {"total_pairs": Bs * world, "pairs_per_gpu": Bs, "steps": k_s, "ms_per_step": float(ts) / k_s,
                  "value": Bs * world * k_s / (float(ts) / 1e3), "unit": UNIT, "scaling": "strong",
                  "note": "BASELINE configs[2]: a fixed job of 256 pairs sharded contiguously over the GPUs, poses all-gathered; "
                          "same timing rules as `value`"}
        if reg_s is not reg:
            del reg_s

    if rank == 0:
        pk = peaks()
        total_pairs = B * world * args.steps
        value = total_pairs / (t_ms / 1e3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": t_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "fp32", "tc": "bf16x3 (tcgen05, fp32 accumulate) + fp32",
                      "tcf": "bf16x3 (feature extraction) + fp16 single pass (correspondence stages), tcgen05, fp32 accumulate; fp32 elsewhere"}[precision],
            "data": "synthetic",
            "config": workload_config(args, world),
            "run": {"l2": "L2 flushed (256 MiB write) between timed iterations", "cuda_graph": not args.no_graph,
                    "precision": precision},
            "clocks": clocks,
            "e2e": {"value": total_pairs / (e2e_ms / 1e3), "unit": UNIT,
                    "h2d_bytes_per_step": 2 * B * N * 3 * 4, "d2h_bytes_per_step": B * 12 * 4,
                    "api": f"Registrar.map, {len(reg._pipe['lanes'])} forward(s) in flight"},
            "in_flight": {"value": total_pairs / (pipe_ms / 1e3), "unit": UNIT, "ms_per_step": pipe_ms / args.steps,
                          "forwards_in_flight": len(lanes),
                          "note": "same K steps, device-resident inputs, back to back on Registrar.map's lanes; "
                                  f"{n_rot} rotating input batches ({n_rot * 2 * B * N * 12 / 1e6:.0f} MB), no L2 flush; "
                                  "`value` is one forward at a time with an L2 flush in front of each"},
            "gpu_launches": launches * args.steps,
            "kernel_breakdown_ms_per_step": prof["families"],
            "roofline": prof["roofline"](pk),
            "kernel_rooflines": prof["kernel_rooflines"](pk, clocks.get("sm_mhz") or 1965),
        }
        if parity is not None:
            line["parity"] = parity
        if strong is not None:
            line["strong_scaling"] = strong
        if world == 1 and args.model == "hregnet" and not args.no_gpu_reference:
            del reg
            torch.cuda.empty_cache()
            line["gpu_reference"] = gpu_reference(args, dev)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_subprocess(args)
        emit(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# kernels launched per C-ABI call (for the gpu_launches claim)
_KERNELS_PER_CALL = {"hrn_cosine_matrix": 4}


def _count_launches(reg):
    from pcd_reg_hregnet_b200 import engine
    n = [0]
    orig = engine.call

    def counting(name, *a):
        n[0] += _KERNELS_PER_CALL.get(name, 1)
        return orig(name, *a)

    engine.call = counting
    try:
        reg._forward()
        torch.cuda.synchronize()
    finally:
        engine.call = orig
    return n[0]


def _profile_families(reg, steps=3):
    """Eager (non-graph) instrumented passes: CUDA events around every C-ABI call on the launching stream,
    aggregated per kernel family; algorithmic FLOP counted for the shared-MLP layer launches."""
    from pcd_reg_hregnet_b200 import engine
    rec = []
    other = {}                      # algorithmic work per step of the non-tensor kernel families
    orig = engine.call

    def timed(name, *a):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        r = orig(name, *a)
        e.record()
        fl, passes = 0.0, 3
        if name.startswith("hrn_layer"):
            passes = a[10] if name == "hrn_layer_tc" else (a[11] if name == "hrn_layer_tc_groupmax" else 0)
            view_rows, cout = a[6], a[7]
            rows_t = a[0]._obj
            K = sum(rows_t.seg[i].channels for i in range(rows_t.n_seg))
            fl = 2.0 * view_rows * K * cout
            from pcd_reg_hregnet_b200 import engine_tc as _et
            if _et.IN_SPLIT:
                # per-point part of a chain's first layer (engine_tc._split_first_layer): its algorithmic MACs are counted
                # with the chain launch below, at the reference's rate of once per ROW; only the issued MACs count here
                rec.append((name, s, e, 0.0, fl * passes))
                return r
        elif name == "hrn_chain_tc":
            # (..., rows a[16], prec a[17], Zb a[18], Zg a[19], ldz a[20], stream)
            passes = a[17]
            rows_t, nl, n1, n2, cout, view_rows = a[0]._obj, a[3], a[4], a[5], a[7], a[16]
            K = sum(rows_t.seg[i].channels for i in range(rows_t.n_seg))
            from pcd_reg_hregnet_b200 import engine_tc as _et
            tail = (n1 * n2 + n2 * cout if nl == 3 else n1 * cout)
            # algorithmic = the reference's first layer over all its input channels for every row
            fl = 2.0 * view_rows * ((K + (_et.SPLIT_K if a[19] is not None else 0)) * n1 + tail)
            rec.append((name, s, e, fl, 2.0 * view_rows * (K * n1 + tail) * passes))
            return r
        elif name == "hrn_chain_wide":
            # (in, W, rank_bytes, bias, n1, n2, n3, chunks0, kseg, Zb, Zg, ldz, G, a, rows, prec, stream)
            passes = a[15]
            rows_t, n1, n2, n3, view_rows = a[0]._obj, a[4], a[5], a[6], a[14]
            K = sum(rows_t.seg[i].channels for i in range(rows_t.n_seg))
            from pcd_reg_hregnet_b200 import engine_tc as _et
            # algorithmic = the reference's first layer over all its input channels for every row (layers.py:364-375)
            fl = 2.0 * view_rows * ((K + _et.SPLIT_K) * n1 + n1 * n2 + n2 * n3)
            rec.append((name, s, e, fl, 2.0 * view_rows * (K * n1 + n1 * n2 + n2 * n3) * passes))
            return r
        elif name == "hrn_chain_wide_head":
            # (in, W, rank_bytes, bias, n1, n2, n3 (padded), chunks0, act, Y, rows, prec, stream): c -> c -> c -> 1
            passes = a[11]
            rows_t, n1, n2, view_rows = a[0]._obj, a[4], a[5], a[10]
            K = sum(rows_t.seg[i].channels for i in range(rows_t.n_seg))
            fl = 2.0 * view_rows * (K * n1 + n1 * n2 + n2 * 1)
            rec.append((name, s, e, fl, 2.0 * view_rows * (K * n1 + n1 * n2 + n2 * a[6]) * passes))
            return r
        elif name == "hrn_level_fused":
            # algorithmic MACs per neighbour row of a fused level: detector + descriptor conv stacks + mlp1 + mlp2
            lv, Bc, Mc, kc = a[0], a[10], a[11], a[13]
            cin, c1, c2, co, cm, cd = {1: (0, 32, 32, 64, 32, 64), 2: (64, 64, 64, 128, 64, 128)}[lv]
            macs = 2 * ((cin + 4) * c1 + c1 * c2 + c2 * co) + 3 * co * cm + cm * cd
            fl = 2.0 * Bc * Mc * kc * macs
        elif name == "hrn_level_ws":
            # the reference's algorithmic MACs per neighbour row of the level (incl. the repeated max_k(X1) block of mlp1,
            # layers.py:203-205, which this kernel evaluates once per keypoint instead of once per neighbour)
            lv, Bc, Mc, kc = a[0], a[11], a[12], a[14]
            cin, c = {2: (64, 64), 3: (128, 128)}[lv]
            macs = 2 * ((cin + 4) * c + c * c + c * 2 * c) + 3 * 2 * c * c + c * 2 * c
            fl = 2.0 * Bc * Mc * kc * macs
        elif name == "hrn_fps":
            # (M-1)*N distance updates x 9 lane-instructions (3 FADD, FMUL, 2 FFMA, FMNMX, FSETP, SEL)
            Bc, Nc, Mc = a[4], a[5], a[6]
            other["hrn_fps"] = other.get("hrn_fps", 0.0) + 9.0 * Bc * max(Mc - 1, 0) * Nc / steps
        elif name in ("hrn_knn", "hrn_knn3_sorted", "hrn_knn3_search"):
            # brute-force-equivalent work of the exact search: M*N pair distances x (3D - 1) flop (8 at D = 3)
            Bc, Mc, Nc = a[3], a[4], a[5]
            Dc = a[6] if name == "hrn_knn" else 3
            other["knn"] = other.get("knn", 0.0) + float(Bc) * Mc * Nc * (3 * Dc - 1) / steps
        rec.append((name, s, e, fl, fl * passes))
        return r

    engine.call = timed
    try:
        for _ in range(steps):
            reg._forward()
        torch.cuda.synchronize()
    finally:
        engine.call = orig
    fam, flops, issued = {}, {}, {}
    for name, s, e, fl, fli in rec:
        fam[name] = fam.get(name, 0.0) + s.elapsed_time(e) / steps
        flops[name] = flops.get(name, 0.0) + fl / steps
        issued[name] = issued.get(name, 0.0) + fli / steps
    fam = dict(sorted(fam.items(), key=lambda kv: -kv[1]))
    layer_names = [n for n in fam if n.startswith("hrn_layer") or n in ("hrn_level_fused", "hrn_level_ws", "hrn_chain_tc", "hrn_chain_wide", "hrn_chain_wide_head")]
    layer_ms = sum(fam[n] for n in layer_names)
    layer_fl = sum(flops[n] for n in layer_names)
    layer_issued = sum(issued[n] for n in layer_names)
    n_layer = sum(1 for r in rec if r[0] in layer_names) / steps

    traffic, traffic_src = None, None
    import glob
    tps = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if tps:     # DRAM bytes of the same kernel family over one step: not measurable without a profiler, so it is read from the
                # newest committed ncu launch list of this same command (profiles/step_breakdown.py), named in traffic_source
        tj = json.load(open(tps[-1]))
        traffic = tj["tensor_family_dram_bytes_per_step"]
        traffic_src = f"profiles/{os.path.basename(tps[-1])} (committed capture, not measured in this run): " + tj["source"]

    def roofline(pk):
        ach = layer_fl / (layer_ms / 1e3) / 1e12 if layer_ms > 0 else 0.0
        return {"kernel": "+".join(layer_names) + " (shared-MLP tensor-core kernels)", "bound": "tensor", "achieved": ach,
                "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops_sustained"],
                "traffic": traffic, "traffic_unit": "bytes per step over the family's launches (ncu dram__bytes_read+write)",
                "traffic_source": traffic_src, "peak_source": pk["source"] + ", sustained bf16 (kernel timed inside a long step)",
                "launches_per_step": n_layer, "ms_per_step": layer_ms, "share_of_step": layer_ms / sum(fam.values()),
                "algorithmic_gflop_per_step": layer_fl / 1e9,
                # precision == "tc": every useful MAC is three bf16 tensor-core MACs (hi*hi + lo*hi + hi*lo), so the
                # tensor pipe itself runs at 3x `frac`; 1/3 is the ceiling of `frac` for this arithmetic
                # tensor-core MACs issued per algorithmic MAC, FLOP-weighted over the family: 3 = bf16 hi/lo operands
                # (hi*hi + lo*hi + hi*lo), 1 = single-pass fp16 (mode "tcf": the correspondence stages)
                "tensor_macs_per_useful_mac": (layer_issued / layer_fl) if layer_fl > 0 else None,
                "frac_of_tensor_pipe_issued": (layer_issued / layer_fl * ach / pk["bf16_tflops_sustained"]) if layer_fl > 0 else None}

    def kernel_rooflines(pk, sm_mhz):
        """The sampling and neighbour-search families against the FP32 issue rate of the chip at the measured clock
        (they are on-chip bound: their HBM traffic is a few MB per step).  kNN counts the brute-force-equivalent work of
        the exact search, so the culled kernels can exceed what a brute-force search could reach."""
        lanes = 148 * 128 * sm_mhz * 1e6                    # lane-instructions per second
        out = {}
        if "hrn_fps" in other and fam.get("hrn_fps"):
            ach = other["hrn_fps"] / (fam["hrn_fps"] / 1e3)
            out["hrn_fps"] = {"bound": "fp32 issue (on-chip)", "achieved": ach / 1e12, "peak": lanes / 1e12,
                              "unit": "T lane-instr/s", "frac": ach / lanes, "ms_per_step": fam["hrn_fps"]}
        knn_ms = sum(fam.get(n, 0.0) for n in ("hrn_knn", "hrn_knn3_sorted", "hrn_knn3_search"))
        if "knn" in other and knn_ms > 0:
            ach = other["knn"] / (knn_ms / 1e3)
            out["knn (search kernels)"] = {"bound": "fp32 (on-chip), brute-force-equivalent flop", "achieved": ach / 1e12,
                                           "peak": 2 * lanes / 1e12, "unit": "TFLOP/s", "frac": ach / (2 * lanes),
                                           "ms_per_step": knn_ms}
        return out

    return {"families": {k: round(v, 4) for k, v in fam.items()}, "roofline": roofline, "kernel_rooflines": kernel_rooflines}


if __name__ == "__main__":
    main()
