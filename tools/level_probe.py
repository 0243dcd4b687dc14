"""Times one hierarchy level (detector + descriptor stage) at the BASELINE configs[1] shape on its own: 64 clouds, the
level's real inputs taken from a forward of the previous levels.  python tools/level_probe.py <level> [reps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pcd_reg_hregnet_b200 import engine, synth  # noqa: E402


def main():
    level = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    net = synth.build_net("hregnet", 7, "cuda")
    src, dst, _, _ = synth.make_batch(range(1000, 1032), 16384)
    pts = torch.cat([src, dst]).cuda()
    fe = net.feature_extraction
    with torch.no_grad():
        xyz, feat, w = pts, None, None
        for lv in (1, 2, 3):
            det, desc = getattr(fe, f"detector_{lv}"), getattr(fe, f"desc_extractor_{lv}")
            if lv == level:
                break
            r = engine.detector_descriptor_level(xyz, feat, w, det.folded(), desc.folded(), det.nsample, det.k)
            xyz, feat, w = r["xyz"], r["af"], engine.sigma_to_weights(r["sigmas"])
        B, N, _ = xyz.shape
        fidx = engine.fps(xyz, det.nsample, w)
        idx, q = engine.knn_idx(None, xyz, det.k, q_idx=fidx)
        from pcd_reg_hregnet_b200 import engine_tc

        def run():
            if level in engine._LEVEL_WS:
                return engine_tc.level_ws(level, q, xyz, feat, idx, det.folded(), desc.folded())
            return engine_tc.level_fused(level, q, xyz, feat, idx, det.folded(), desc.folded())

        for _ in range(3):
            run()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            run()
        e.record()
        torch.cuda.synchronize()
        print(f"level {level}: {s.elapsed_time(e) / reps * 1e3:.1f} us per launch ({B} clouds, {det.nsample} x {det.k} rows each)")
        from pcd_reg_hregnet_b200 import _lib
        L = _lib.lib()
        if hasattr(L, "hrn_level_ws_prof") and level in engine._LEVEL_WS:      # library built with -DLW_PROF
            import ctypes
            buf = (ctypes.c_ulonglong * 64)()
            L.hrn_level_ws_prof(buf, 1)
            run()
            torch.cuda.synchronize()
            L.hrn_level_ws_prof(buf, 0)
            names = ["gather", "wait L0", "drain C1d+C1x", "wait d2", "C2d|wait x2|C2x", "wait d3", "attention", "E*a drain+x3", "wait m1c",
                     "-", "-", "-", "-", "X1 drain", "bar", "mat-vec", "bar+wait m1b", "M1 drain",
                     "wait m2", "desc epilogue"]
            tot = sum(buf[i] for i in range(20))
            print("epilogue warp 0 of CTA 0, cycles per phase (share):")
            for i, nme in enumerate(names):
                print(f"  {nme:14s} {buf[i]:10d}  {100.0 * buf[i] / max(tot, 1):5.1f} %")
            mt = sum(buf[i] for i in range(32, 36))
            for i, nme in zip(range(32, 36), ["wait gather", "wait operand", "wait weights", "issue"]):
                print(f"  MMA {nme:12s} {buf[i]:10d}  {100.0 * buf[i] / max(mt, 1):5.1f} %")


if __name__ == "__main__":
    main()
