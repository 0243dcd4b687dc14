"""Hottest SASS lines of an `ncu --set full --import-source on` report (source page), with their stall breakdown.
usage: python profiles/top_stalls.py report.ncu-rep [n]"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader([l for l in out.splitlines() if not l.startswith("==")]))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]]) for r in body)
inst = sum(int(r[ix["Instructions Executed"]]) for r in body)
print(f"# {rows[0][1][:100]}\n# samples {tot}, warp instructions {inst}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]]) for r in body) for s in stalls}
print("# stall totals:", ", ".join(f"{k[6:]} {v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]]))[:n]
for i in sorted(order):
    r = body[i]
    top = sorted(((int(r[ix[s]]), s[6:]) for s in stalls), reverse=True)[:2]
    print(f"{i:5d} {int(r[ix['# Samples']]):6d} {int(r[ix['Instructions Executed']]):9d}  {r[ix['Source']].strip()[:70]:70s} {top}")
