#!/usr/bin/env python
"""Walk the SASS of a profiled solver kernel in address order and print, at every
reconvergence / branch boundary where the active-lane count changes, how often the
region ran and with how many lanes -- the view that shows which source blocks a warp
executes in separate lane groups.  usage: tools/ncu_trace.py report.ncu-rep solver rhs neq"""
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
rep, solver, rhs, neq = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4])
thr = float(sys.argv[5]) if len(sys.argv) > 5 else 3.0
import numbalsoda_b200 as nb  # noqa: E402
from numbalsoda_b200 import _lib  # noqa: E402

h = nb.get_handle(nb.builtin(rhs), _lib.LSODA if solver.startswith("lsoda") else _lib.DOP853, neq)
open("/tmp/_k.cubin", "wb").write(h.cubin())
kname = "b200ode_%s_kernel" % solver
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hdr, body = rows[1], rows[2:]
iS, iI, iA = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
dis = subprocess.run(["nvdisasm", "-gi", "/tmp/_k.cubin"], capture_output=True, text=True).stdout
line_of, cur, fresh, on = [], ("?", 0), True, False
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
prev = None
start = 0
acc_n = acc_t = 0
for k, r in enumerate(body):
    n, a = int(r[iI]), float(r[iA])
    if n == 0:
        continue
    if prev is not None and abs(a - prev) > thr:
        f, l = line_of[start] if start < len(line_of) else ("?", 0)
        f2, l2 = line_of[k - 1] if k - 1 < len(line_of) else ("?", 0)
        print("sass %5d-%5d  warp-inst %12d  lanes %5.1f   %s:%d .. %s:%d" % (start, k - 1, acc_n, acc_t / max(acc_n, 1), f, l, f2, l2))
        start, acc_n, acc_t = k, 0, 0
    acc_n += n
    acc_t += n * a
    prev = a
