#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source sass` output: per kernel
launch, executed warp instructions by opcode and the hottest SASS lines."""
import csv
import sys
from collections import Counter

path = sys.argv[1]
sel = int(sys.argv[2]) if len(sys.argv) > 2 else None  # launch index to detail
top = int(sys.argv[3]) if len(sys.argv) > 3 else 0
launches = []
cur = None
for row in csv.reader(open(path)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": [], "hdr": None}
        launches.append(cur)
    elif row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] is not None:
        cur["rows"].append(row)
for li, L in enumerate(launches):
    h = {n: i for i, n in enumerate(L["hdr"])}
    ops = Counter()
    total = 0
    for r in L["rows"]:
        try:
            n = int(float(r[h["Instructions Executed"]]))
        except ValueError:
            continue
        src = r[h["Source"]].strip()
        if src.startswith("@"):
            src = src.split(None, 1)[1]
        op = src.split()[0].split(".")[0] if src else "?"
        ops[op] += n
        total += n
    print(f"[{li}] {L['name'][:70]}  total warp-inst={total}")
    if sel is None or sel == li:
        print("   " + ", ".join(f"{o}:{c} ({100*c/total:.1f}%)" for o, c in ops.most_common(22)))
        if top:
            rows = sorted(L["rows"], key=lambda r: -float(r[h["# Samples"]] or 0))[:top]
            for r in rows:
                print(f"     samples={r[h['# Samples']]:>6} exec={r[h['Instructions Executed']]:>9} {r[h['Source']][:90]}")
