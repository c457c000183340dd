#!/usr/bin/env python
"""Key per-launch metrics from an .ncu-rep (needs `ncu` on PATH):
   python tools/ncu_kernel_summary.py report.ncu-rep [launch_index]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
sel = int(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__cycles_active.avg", "smsp__cycles_active.avg"]
for i, r in enumerate(rows[2:]):
    if sel is not None and i != sel:
        continue
    d = dict(zip(hdr, r))
    units = dict(zip(hdr, rows[1]))
    print(f"[{i}] {d['Kernel Name'][:90]}")
    for k in KEYS:
        if k in d:
            print(f"    {k:72s} {d[k]:>16s} {units[k]}")
    stalls = []
    for k in hdr:
        if "warp_issue_stalled" in k and k.endswith("per_warp_active.pct") and "not_issued" not in k:
            try:
                stalls.append((float(d[k]), k.split("warp_issue_stalled_")[1].split("_per_warp")[0]))
            except ValueError:
                pass
    stalls.sort(reverse=True)
    print("    stalls(% of warp-active samples): " + ", ".join(f"{n}={v:.1f}" for v, n in stalls[:8]))
