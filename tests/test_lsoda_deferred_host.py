"""The deferred (block-served) organisation of the thread-per-trajectory LSODA kernel, emulated on
the host with virtual threads in lockstep (tests/lsoda_deferred_host.cpp), against the CPU oracle:
states and counters bit for bit on sweeps whose trajectories need Jacobians, retractions, rescales
and output rows at different passes -- the requests of several trajectories interleave in the
service exactly as on the GPU.  tests/test_gpu_lsoda.py then runs the kernel itself."""
import ctypes as ct
import os
import subprocess

import numpy as np
import pytest

import _oracle as orc
import _problems as pr

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "lsoda_deferred_host.cpp")
HDRS = [os.path.join(ROOT, "numbalsoda_b200", "csrc", f)
        for f in ("lsoda_core.cuh", "lsoda_vec_thread.cuh", "lsoda_tables.cuh", "b200ode_pow.cuh",
                  "b200ode_fastmath.cuh")]
_libs = {}


def lib(n, threads=8):
    key = (n, threads)
    if key not in _libs:
        os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
        so = os.path.join(HERE, "_build", "liblsoda_deferred_n%d_t%d.so" % (n, threads))
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in [SRC] + HDRS):
            subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-mfma", "-fPIC", "-shared",
                                   "-Wno-unknown-pragmas", "-DB200_NEQ=%d" % n, "-DLS_T=%d" % threads,
                                   "-o", so, SRC])
        _libs[key] = ct.CDLL(so)
    return _libs[key]


def run(fptr, u0, te, data, rtol, atol, mxstep=10000, exit_on_warning=False, threads=8):
    u0 = np.ascontiguousarray(u0, dtype=np.float64)
    te = np.ascontiguousarray(te, dtype=np.float64)
    data = np.ascontiguousarray(data, dtype=np.float64)
    B = u0.shape[0] if u0.ndim == 2 else data.shape[0]
    n = u0.shape[-1]
    us = np.full((B, len(te), n), np.nan)
    ok = np.full(B, 999, dtype=np.int32)
    cnt = np.zeros((B, 8), dtype=np.int32)
    p = lambda a: a.ctypes.data_as(ct.c_void_p)  # noqa: E731
    lib(n, threads).lsoda_deferred_host(
        ct.c_void_p(fptr), ct.c_long(B), n, p(u0), ct.c_long(n if u0.ndim == 2 else 0), p(data),
        ct.c_long(data.shape[-1] * 8 if data.ndim == 2 else 0), len(te), p(te), p(us), ct.c_double(rtol),
        ct.c_double(atol), int(mxstep), p(ok), int(exit_on_warning), p(cnt))
    return us, ok == 1, cnt


@pytest.mark.parametrize("threads", [1, 8, 32])
def test_robertson_sweep_bitwise(threads):
    B = 100
    P = pr.robertson_sweep(B, 7)
    u0, te = np.array([1.0, 0.0, 0.0]), np.linspace(0, 1e5, 100)
    f = orc.builtin_rhs("robertson")
    us, ok, cnt = run(f, u0, te, P, 1e-8, 1e-8, threads=threads)
    ref, okr, cr = orc.solve_batch("lsoda", f, u0, te, P, 1e-8, 1e-8)
    assert ok.all() and okr.all() and np.array_equal(cnt, cr) and np.array_equal(us, ref)
    assert (cnt[:, 2] > 20).all()  # every trajectory asked for Jacobians


@pytest.mark.parametrize("name", ["lv_test", "lv_readme", "robertson", "lorenz_short"])
def test_reference_problems_bitwise(name):
    p = pr.problem(name)
    f = orc.builtin_rhs(p["builtin"])
    us, ok, cnt = run(f, p["u0"], p["t_eval"], p["data"][None, :], p["rtol"], p["atol"])
    ref, okr, st = orc.lsoda(f, p["u0"], p["t_eval"], p["data"], p["rtol"], p["atol"])
    assert ok[0] == okr
    assert list(cnt[0]) == [st.nst, st.nfe, st.nje, st.nswitch, st.nqu, st.istate, st.rows, st.meth]
    assert np.array_equal(us[0], ref)


def test_many_rows_per_step_and_failures():
    """Several output rows inside one step (one request, many rows), duplicate output times, the
    three-failures recovery (right-hand side refreshed after the deferred retraction), step
    budgets that end trajectories with requests pending, tolerances that fail at once."""
    f = orc.builtin_rhs("lotka_volterra")
    u0 = np.array([5.0, 0.8])
    P = np.random.default_rng(3).uniform(0.5, 3.0, (24, 1))
    cases = [dict(te=np.linspace(0, 50, 4000), rtol=1e-3, atol=1e-6),
             dict(te=np.array([0.0, 1.0, 1.0, 2.0, 2.0, 2.0, 30.0]), rtol=1e-8, atol=1e-8),
             dict(te=np.linspace(0, 10, 11), rtol=1e-8, atol=1e-8, mxstep=25),
             dict(te=np.linspace(0, 10, 11), rtol=1e-17, atol=1e-20),
             dict(te=np.array([0.0, 1e-9, 2e-9, 5.0, 5.0 + 1e-9]), rtol=1e-6, atol=1e-9),
             dict(te=np.array([0.0, 1.0, 0.5, 2.0]), rtol=1e-8, atol=1e-8)]
    for c in cases:
        kw = dict(mxstep=c.get("mxstep", 10000))
        us, ok, cnt = run(f, u0, c["te"], P, c["rtol"], c["atol"], **kw)
        ref, okr, cr = orc.solve_batch("lsoda", f, u0, c["te"], P, c["rtol"], c["atol"], **kw)
        assert list(ok) == list(okr) and np.array_equal(cnt, cr), c
        for b in range(len(P)):
            assert np.array_equal(us[b, :cr[b, 6]], ref[b, :cr[b, 6]]), c
    # stiff problems that fail error tests repeatedly (order reset path)
    fv = pr.cfunc_address(_vdp)
    Pv = np.array([[1000.0], [300.0], [3000.0], [50.0]])
    te = np.linspace(0, 3000, 61)
    us, ok, cnt = run(fv, np.array([2.0, 0.0]), te, Pv, 1e-6, 1e-8)
    ref, okr, cr = orc.solve_batch("lsoda", fv, np.array([2.0, 0.0]), te, Pv, 1e-6, 1e-8)
    assert list(ok) == list(okr) and np.array_equal(cnt, cr) and np.array_equal(us, ref)


def _vdp(t, u, du, p):
    du[0] = u[1]
    du[1] = p[0] * ((1.0 - u[0] * u[0]) * u[1]) - u[0]


@pytest.mark.parametrize("n", [1, 2, 4, 5, 8])
def test_other_system_sizes(n):
    fc = pr.cfunc_address(pr.chain_rhs if n > 2 else (_decay1 if n == 1 else pr.lv_rhs))
    if n == 1:
        u0, P, te = np.array([1.0]), np.array([[k] for k in (1.0, 50.0, 2000.0)]), np.linspace(0, 5, 21)
    elif n == 2:
        u0, P, te = np.array([5.0, 0.8]), np.array([[1.0], [2.5], [0.7]]), np.linspace(0, 20, 41)
    else:
        u0, te = 1.0 + 0.1 * np.arange(n), np.linspace(0, 5, 11)
        P = np.array([[k, float(n)] for k in (2.0, 200.0, 5000.0, 40.0)])
    us, ok, cnt = run(fc, u0, te, P, 1e-7, 1e-9)
    ref, okr, cr = orc.solve_batch("lsoda", fc, u0, te, P, 1e-7, 1e-9)
    assert list(ok) == list(okr) and np.array_equal(cnt, cr) and np.array_equal(us, ref)


def _decay1(t, u, du, p):
    du[0] = -p[0] * u[0] + 1.0
