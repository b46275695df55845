#!/usr/bin/env python
"""Per-source-line view of an .ncu-rep for an LTO-linked solver kernel: joins the SASS
page of the report with nvdisasm's line table of the same cubin (re-linked here; the
link is deterministic).  usage: tools/ncu_by_line.py report.ncu-rep solver rhs neq [top]
Prints, per source line: executed warp instructions, share, average active threads,
and stall samples."""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rep, solver, rhs, neq = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
import numbalsoda_b200 as nb  # noqa: E402
from numbalsoda_b200 import _lib  # noqa: E402

h = nb.get_handle(nb.builtin(rhs), _lib.LSODA if solver.startswith("lsoda") else _lib.DOP853, neq)
open("/tmp/_k.cubin", "wb").write(h.cubin())
kname = os.environ.get("NCU_KERNEL", "b200ode_%s_kernel" % solver)
dis = subprocess.run(["nvdisasm", "-gi", "/tmp/_k.cubin"], capture_output=True, text=True).stdout
line_of = []  # per instruction, in order; innermost inlined location
cur = ("?", 0)
fresh = True  # the first annotation after an instruction is the innermost frame
on = False
for ln in dis.splitlines():
    if ln.startswith("\t.section"):
        on = (".text.%s," % kname) in ln
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        if fresh:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            fresh = False
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", ln):
        line_of.append(cur)
        fresh = True
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr = rows[1]
iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
body = rows[2:]
if len(body) != len(line_of):
    print("warning: %d SASS rows in the report vs %d in the re-linked cubin" % (len(body), len(line_of)))
agg = collections.defaultdict(lambda: [0, 0, 0, 0])
tot = 0
for k, r in enumerate(body):
    if k >= len(line_of):
        break
    n, t, sm = int(r[iI]), int(r[iT]), int(r[iS])
    a = agg[line_of[k]]
    a[0] += n
    a[1] += t
    a[2] += sm
    a[3] += 1
    tot += n
tsm = sum(a[2] for a in agg.values())
print("%-28s %6s %7s %7s %8s %6s" % ("file:line", "SASS", "inst %", "active", "samples%", "cum %"))
cum = 0.0
for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    cum += 100.0 * a[2] / tsm
    print("%-28s %6d %7.2f %7.1f %8.2f %6.1f" % ("%s:%d" % (f, l), a[3], 100.0 * a[0] / tot,
                                                   a[1] / max(a[0], 1), 100.0 * a[2] / tsm, cum))
