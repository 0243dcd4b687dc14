"""Times the level-1 neighbour search (64 clouds x 16384 points, 1024 sampled queries, K = 64): Morton sort and culled
search separately.  HRN_KNN_SORT=b selects the bitonic sort.  python tools/knn_time.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pcd_reg_hregnet_b200 import engine, synth
from pcd_reg_hregnet_b200.engine import call, ptr, stream

B, N, M, K = 64, 16384, 1024, 64
x = torch.stack([synth.make_pair(100 + b, N)[1] for b in range(8)]).repeat(8, 1, 1).cuda().contiguous()
fidx = engine.fps(x, M)
pts, boxes = engine.knn_scratch(B, N, x.device)
idx = torch.empty(B, M, K, dtype=torch.int32, device=x.device)
q = torch.empty(B, M, 3, device=x.device)


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / reps * 1e3


sort = lambda: call("hrn_knn3_sort", ptr(x), B, N, ptr(pts), ptr(boxes), stream())
search = lambda: call("hrn_knn3_search", None, ptr(fidx), ptr(x), B, M, N, K, ptr(pts), ptr(boxes), None, None, ptr(idx), None, ptr(q), stream())
print(f"sort {t(sort):.1f} us, search {t(search):.1f} us  (HRN_KNN_SORT={os.environ.get('HRN_KNN_SORT', 'radix')})")
