#!/usr/bin/env python
"""Small runs of every kernel for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _problems as pr  # noqa: E402
import numbalsoda_b200 as nb  # noqa: E402

which = sys.argv[1:] or ["dop853", "lsoda", "lsodaw", "ivp"]
u0 = np.array([1.0, 0.0, 0.0])
if "dop853" in which:
    us, ok = nb.dop853(nb.builtin("lorenz"), u0, np.linspace(0, 2, 21), pr.lorenz_sweep(200), 1e-8, 1e-8)
    print("dop853", ok.all())
if "lsoda" in which:
    us, ok = nb.lsoda(nb.builtin("robertson"), u0, np.linspace(0, 1e3, 20), pr.robertson_sweep(200), 1e-8, 1e-8)
    print("lsoda", ok.all())
if "lsodaw" in which:
    for strict in (False, True):
        us, ok = nb.lsoda(nb.builtin("brusselator1d"), pr.brusselator_u0(), np.linspace(0, 1.0, 6),
                          pr.brusselator_sweep(6), 1e-6, 1e-8, strict=strict)
        print("lsodaw strict=%s" % strict, ok.all())
    us, ok = nb.lsoda(pr.upper_rhs, np.ones(12), np.linspace(0, 0.5, 5),
                      np.array([[1000.0, 5000.0, 12.0]] * 3), 1e-6, 1e-9)
    print("lsodaw pivoting", ok.all())
if "ivp" in which:
    D = np.column_stack([pr.lorenz_sweep(300), np.full(300, 45.0)])
    for te in (np.linspace(0, 3, 31), None):
        sol = nb.solve_ivp(nb.builtin("lorenz"), np.array([0.0, 3.0]), u0, te, n_events=1,
                           event_fcn=nb.builtin("lorenz_z"), data=D, rtol=1e-8, atol=1e-8)
        print("ivp", "t_eval" if te is not None else "ragged", sol.success.all(), sol.event_found.mean())
