#!/usr/bin/env python
"""Top stalled SASS instructions of one launch in an .ncu-rep (source page, SASS view).
usage: ncu_stalls.py report.ncu-rep <launch index> [top N]"""
import csv
import subprocess
import sys

rep, idx = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout.splitlines()
print(out[0][:160])
rows = list(csv.reader(out[1:]))
h = rows[0]
ci = {n: h.index(n) for n in h}
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
data = []
seen = set()
for i, r in enumerate(rows[1:]):
    if r and r[0] in seen:
        continue
    seen.add(r[0] if r else None)
    if len(r) < len(h) or not r[ci["# Samples"]].isdigit():
        continue
    s = int(r[ci["# Samples"]] or 0)
    data.append((s, i, r))
total = sum(d[0] for d in data)
print(f"total samples {total}")
agg = {}
for s_, i, r in data:
    for c in stall_cols:
        agg[c[6:]] = agg.get(c[6:], 0) + int(r[ci[c]] or 0)
print("by reason:", " ".join(f"{k}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v))
for s, i, r in sorted(data, reverse=True)[:top]:
    reasons = sorted(((int(r[ci[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{s:7d} {100.0 * s / total:5.1f}%  #{i:5d} {r[ci['Source']].strip()[:70]:70s} " + " ".join(f"{n}={v}" for v, n in reasons if v))
