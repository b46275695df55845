#!/usr/bin/env python
"""Generate the committed golden vectors (run in the build container, where
/root/reference exists; the GPU box only reads the .npz files).

  lsoda_reference.npz : outputs and private counters of the UNMODIFIED reference
                        LSODA (oracle/_ref/libref_probe.so, compiled from
                        /root/reference/src by oracle/Makefile) with numba-cfunc
                        right-hand sides -- the same machine code path a
                        numbalsoda user exercises.
  dop853_oracle.npz   : outputs and counters of oracle/dop853_oracle.c (the
                        reference's DOP853 is Fortran and cannot be built here).

Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _oracle as orc  # noqa: E402
import _problems as pr  # noqa: E402

NAMES = ["lv_test", "lv_readme", "robertson", "lorenz", "lorenz_short"]


def main():
    assert orc.HAVE_REF, "reference sources not available: run where /root/reference exists"
    ls, dp = {}, {}
    for name in NAMES:
        p = pr.problem(name)
        f = pr.cfunc_address(p["rhs"])
        us, ok, st = orc.ref_lsoda(f, p["u0"], p["t_eval"], p["data"], p["rtol"], p["atol"])
        ls[name + "/usol"] = us
        ls[name + "/success"] = np.array(ok)
        ls[name + "/counters"] = np.array(st, dtype=np.int64)  # nst nfe nje nswitch nqu istate rows meth
        us, ok, st = orc.dop853(f, p["u0"], p["t_eval"], p["data"], p["rtol"], p["atol"])
        dp[name + "/usol"] = us
        dp[name + "/success"] = np.array(ok)
        dp[name + "/counters"] = np.array([st.nstep, st.nfcn, st.naccpt, st.nrejct, st.nevent,
                                           st.idid, st.rows], dtype=np.int64)
    # tests/test_exit_on_warning.py: only the flags are defined (u0 is an int64 array there)
    f = pr.cfunc_address(pr.growth_rhs)
    te = np.linspace(100, 101, 11)
    u0 = np.array([1]).view(np.float64)  # the reference test passes an int64 array (:29,:44)
    for eow in (False, True):
        us, ok, st = orc.ref_lsoda(f, u0, te, np.array([1e15]), 1e-9, 1e-9,
                                   exit_on_warning=eow)
        ls["growth_eow%d/usol" % eow] = us
        ls["growth_eow%d/success" % eow] = np.array(ok)
        ls["growth_eow%d/counters" % eow] = np.array(st, dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "lsoda_reference.npz"), **ls)
    np.savez_compressed(os.path.join(HERE, "dop853_oracle.npz"), **dp)
    for k in sorted(ls):
        if k.endswith("counters"):
            print("lsoda ", k, ls[k])
    for k in sorted(dp):
        if k.endswith("counters"):
            print("dop853", k, dp[k])


if __name__ == "__main__":
    main()
