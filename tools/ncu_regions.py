#!/usr/bin/env python
"""Per-instruction stall listing of one kernel from an .ncu-rep (source page, SASS):
   python tools/ncu_regions.py report.ncu-rep [min_samples]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
thr = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
h = {n: i for i, n in enumerate(rows[hi])}
data = rows[hi + 1:]
st = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(float(r[h["# Samples"]] or 0) for r in data)
print("total samples", int(tot), "instructions", len(data))
for i, r in enumerate(data):
    s = float(r[h["# Samples"]] or 0)
    if s < thr:
        continue
    reasons = sorted(((float(r[h[k]] or 0), k[6:]) for k in st), reverse=True)
    rs = " ".join(f"{n}={int(v)}" for v, n in reasons[:3] if v > 0)
    print(f"{i:4d} {int(s):4d} {r[h['Instructions Executed']]:>9s} {r[h['Source']][:64]:64s} {rs}")
