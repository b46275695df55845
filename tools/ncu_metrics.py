#!/usr/bin/env python
"""Print selected metrics of an .ncu-rep (raw page) -- helper for profiles/ summaries.
usage: tools/ncu_metrics.py report.ncu-rep [substring ...]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
subs = sys.argv[2:] or ["gpu__time_duration.sum", "launch__registers_per_thread",
                        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
                        "smsp__issue_active.avg.pct", "sm__warps_active.avg.pct_of_peak_sustained_active",
                        "smsp__thread_inst_executed_per_inst_executed.ratio",
                        "dram__bytes_read.sum", "dram__bytes_write.sum", "local",
                        "smsp__average_warp", "issue_stalled"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    print("== kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
    for h, u, v in zip(hdr, units, r):
        if any(s in h for s in subs):
            print("  %-95s %s %s" % (h, v, u))
