"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_<tag>.txt"""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
    v = float(row["Metric Value"].replace(",", ""))
    v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print(f"# {sys.argv[1]}: {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.3f} ms total (cold-cache, serialised: compare SHARES)")
print(f"{'us':>12} {'share':>7} {'n':>5}  kernel")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:12.1f} {100 * v[1] / tot:6.1f}% {v[0]:5d}  {k[:100]}")
