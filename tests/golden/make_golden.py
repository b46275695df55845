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
  solve_ivp_oracle.npz: every ODEResult field and the counters of
                        oracle/solve_ivp_oracle.c on the cases of the reference's
                        solve_ivp.ipynb (cells 11, 26, 27) and tests
                        (tests/test_numbalsoda.py:33-56), next to the values the
                        reference itself printed there (`*/reference_t_event`).

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

# solve_ivp.ipynb: printed by the reference (gfortran build on its author's machine)
NOTEBOOK_T_EVENT_CELL_11 = 0.7611343988149695
NOTEBOOK_T_EVENT_CELL_26 = 0.7258968839312535
NOTEBOOK_T_EVENT_CELL_27 = [0.7463063968797129, 0.7679450731640909, 0.7909998733859201,
                            0.8157222725496384, 0.8424776519856828, 0.8718277396448789,
                            0.9047034490300899, 0.9428490823677572, 0.9901493996386634,
                            1.0587894770207835]


def ivp_cases():
    """name -> (builtin rhs, t_span, y0, t_eval, n_events, builtin events, data, kwargs, printed)"""
    u0 = np.array([5.0, 0.8])
    te = np.linspace(0.0, 50.0, 1000)
    span = np.array([0.0, 50.0])
    c = {"cell11": ("lotka_volterra2", span, u0, te, 1, "lv_prey", np.array([1.0, 1.0]),
                    dict(rtol=1e-6), NOTEBOOK_T_EVENT_CELL_11),
         "cell26": ("lotka_volterra2", span, u0, te, 1, "lv_prey", np.array([0.6, 0.15]),
                    dict(rtol=1e-6), NOTEBOOK_T_EVENT_CELL_26)}
    for i, t in enumerate(NOTEBOOK_T_EVENT_CELL_27):
        S = np.array([np.sum(np.full(6, 0.1 - i * 0.01)), np.sum(np.full(3, 0.1 + i * 0.01))])
        c["cell27_%d" % i] = ("lotka_volterra2", span, u0, te, 1, "lv_prey", S, dict(rtol=1e-6), t)
    te11 = np.linspace(0.0, 10.0, 11)
    c["test_t_eval"] = ("lotka_volterra2", np.array([0.0, 10.0]), u0, te11, 0, None, np.array([1.0, 1.0]),
                        dict(rtol=1e-8, atol=1e-8), None)
    c["test_steps"] = ("lotka_volterra2", np.array([0.0, 10.0]), u0, None, 0, None, np.array([1.0, 1.0]),
                       dict(rtol=1e-8, atol=1e-8), None)
    c["backward"] = ("lotka_volterra2", np.array([0.0, -50.0]), u0, -te, 0, None, np.array([1.0, 1.0]),
                     dict(rtol=1e-9, atol=1e-30), None)
    return c


def make_ivp():
    out = {}
    for name, (rhs, span, y0, te, ne, ev, data, kw, printed) in ivp_cases().items():
        r = orc.solve_ivp(orc.builtin_rhs(rhs), span, y0, te, ne,
                          orc.builtin_rhs("events_" + ev) if ev else 0, data, cap=4000, **kw)
        out[name + "/t"], out[name + "/y"] = r["t"], r["y"]
        out[name + "/y_event"] = r["y_event"]
        out[name + "/t_event"] = np.array(r["t_event"])
        out[name + "/ints"] = np.array([r["success"], orc.IVP_MESSAGES.index(r["message"]), r["nfev"],
                                        len(r["t"]), r["rows_needed"], r["event_found"], r["ind_event"],
                                        r["idid"], r["nstep"], r["naccpt"], r["nrejct"]], dtype=np.int64)
        if printed is not None:
            out[name + "/reference_t_event"] = np.array(printed)
            assert abs(r["t_event"] - printed) < 1e-13, (name, r["t_event"], printed)
    np.savez_compressed(os.path.join(HERE, "solve_ivp_oracle.npz"), **out)
    for k in sorted(out):
        if k.endswith("ints"):
            print("ivp   ", k, out[k])


def main():
    make_ivp()
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
