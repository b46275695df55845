#!/usr/bin/env python
"""Text summary of an .ncu-rep for profiles/: headline metrics, stall reasons, executed
instruction mix with SIMT efficiency.  usage: tools/ncu_summary.py report.ncu-rep [B]"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
print("kernel:", m.get("Kernel Name", ("?",))[0])
for k in keys:
    if k in m:
        print("  %-72s %s %s" % (k, m[k][0], m[k][1]))
print("  stall reasons (warps per issue-active cycle):")
for h in hdr:
    if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
        v = float(m[h][0])
        if v >= 0.02:
            print("    %-30s %.3f" % (h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h2 = rows[1]
iS, iI, iT = h2.index("Source"), h2.index("Instructions Executed"), h2.index("Thread Instructions Executed")
tot = tott = 0
by, byt = collections.Counter(), collections.Counter()
for r in rows[2:]:
    try:
        n, t = int(r[iI]), int(r[iT])
    except (ValueError, IndexError):
        continue
    toks = r[iS].split()
    op = (toks[1] if toks[0].startswith("@") else toks[0]).split(".")[0]
    by[op] += n
    byt[op] += t
    tot += n
    tott += t
print("  executed warp instructions: %d, average active threads %.2f / 32" % (tot, tott / tot))
if B:
    print("  warp instructions per 32 trajectories: %.4g" % (tot * 32.0 / B))
fp = sum(by[o] for o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print("  FP64-pipe share of executed instructions: %.1f%%" % (100.0 * fp / tot))
for op, n in by.most_common(16):
    print("    %-8s %6.2f%%  avg active threads %.1f" % (op, 100.0 * n / tot, byt[op] / n))
