"""Per-kernel breakdown of ONE 32-pair step from an ncu launch list that carries gpu__time_duration.sum and the two
dram__bytes counters (step boundaries = successive level-1 FPS launches).  Writes the table to stdout and the
DRAM traffic of the tensor-core kernel family to profiles/<tag>_traffic.json (the newest one is read by bench.py for
roofline.traffic).
usage: python profiles/step_breakdown.py gpurun_out/launches_final.csv [tag=r02] > profiles/r02_launches_final.txt"""
import collections
import csv
import json
import os
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
by = {}
for r in csv.DictReader(lines):
    d = by.setdefault(int(r["ID"]), {"name": re.sub(r"\(.*", "", r["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")})
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    if r["Metric Name"] == "gpu__time_duration.sum":
        d["us"] = {"ns": v / 1e3, "us": v, "ms": v * 1e3}[u]
    elif r["Metric Name"].startswith("dram__bytes"):
        d[r["Metric Name"]] = v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
ids = sorted(by)
# a step starts with the level-1 furthest-point sampling: the only fps launch over the full clouds
marks = [n for n, i in enumerate(ids) if "fps_cluster_kernel<512" in by[i]["name"] or "fps_cull_kernel" in by[i]["name"]]
a = marks[0]
if len(marks) > 1:
    b = marks[1]
else:                                                  # window holds one step start: the step ends at the last pose kernel
    b = max(n for n, i in enumerate(ids) if "kabsch_kernel" in by[i]["name"] and n > a) + 1
fam = collections.defaultdict(lambda: [0, 0.0, 0.0])
tot = 0.0
for n in range(a, b):
    d = by[ids[n]]
    f = fam[d["name"][:40]]
    f[0] += 1
    f[1] += d["us"]
    f[2] += d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
    tot += d["us"]
print(f"# one 32-pair step (ncu launch list, cold-cache serialised: compare SHARES): {tot / 1e3:.3f} ms, {b - a} launches")
print(f"{'us':>10} {'share':>6} {'n':>4} {'dram MB':>10}  kernel")
for k, v in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} {100 * v[1] / tot:5.1f}% {v[0]:4d} {v[2] / 1e6:10.1f}  {k}")
tc = [k for k in fam if any(s in k for s in ("level_fused", "level_ws", "chain3", "chain_ws", "chain_wide", "layer_tc", "layer_ws"))]
out = {"tensor_family_dram_bytes_per_step": sum(fam[k][2] for k in tc),
       "tensor_family_us_per_step_ncu": sum(fam[k][1] for k in tc), "n_launches": sum(fam[k][0] for k in tc),
       "share_of_step_ncu": sum(fam[k][1] for k in tc) / tot,
       "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none, one 32-pair step"}
tag = sys.argv[2] if len(sys.argv) > 2 else "r02"
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), tag + "_traffic.json"), "w"), indent=1)
