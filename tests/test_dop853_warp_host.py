"""The warp-per-trajectory DOP853 core (numbalsoda_b200/csrc/dop853_warp.cuh) compiled as
host C++ (one emulated lane) against the CPU oracle: states and counters bit for bit."""
import ctypes as ct
import os
import subprocess

import numpy as np
import pytest

import _oracle as orc
import _problems as pr

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "dop853_warp_host.cpp")
HDRS = [os.path.join(ROOT, "numbalsoda_b200", "csrc", f)
        for f in ("dop853_warp.cuh", "b200ode_warp.cuh", "dop853_tableau.cuh", "b200ode_pow.cuh")]


def build_lib():
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    so = os.path.join(HERE, "_build", "libdop853_warp_host.so")
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in [SRC] + HDRS):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-mfma", "-fPIC", "-shared",
                               "-Wno-unknown-pragmas", "-o", so, SRC])
    return ct.CDLL(so)


@pytest.fixture(scope="module")
def lib():
    return build_lib()


def run(lib, fptr, u0, te, data, rtol, atol, mxstep=10000):
    u0 = np.ascontiguousarray(u0, dtype=np.float64)
    te = np.ascontiguousarray(te, dtype=np.float64)
    data = np.ascontiguousarray(data, dtype=np.float64)
    us = np.full((len(te), len(u0)), np.nan)
    ok = ct.c_int(999)
    cnt = (ct.c_int * 8)()
    p = lambda a: a.ctypes.data_as(ct.c_void_p)  # noqa: E731
    lib.dop853_warp_host(ct.c_void_p(fptr), len(u0), p(u0), p(data), len(te), p(te), p(us),
                         ct.c_double(rtol), ct.c_double(atol), int(mxstep), ct.byref(ok), cnt)
    return us, ok.value == 1, list(cnt)


def ocnt(st):
    return [st.nstep, st.nfcn, st.naccpt, st.nrejct, st.nevent, st.idid, st.rows, 0]


@pytest.mark.parametrize("name", ["lv_test", "lv_readme", "lorenz_short", "lorenz", "robertson"])
def test_reference_problems_bitwise(lib, name):
    p = pr.problem(name)
    if name == "robertson":
        p["t_eval"] = np.linspace(0, 40, 5)
    f = orc.builtin_rhs(p["builtin"])
    us, ok, cnt = run(lib, f, p["u0"], p["t_eval"], p["data"], p["rtol"], p["atol"], mxstep=100000)
    ref, okr, st = orc.dop853(f, p["u0"], p["t_eval"], p["data"], p["rtol"], p["atol"], mxstep=100000)
    assert ok == okr and cnt == ocnt(st)
    assert np.array_equal(us[:st.rows], ref[:st.rows], equal_nan=True)


def test_brusselator_n64_and_edge_cases(lib):
    f = orc.builtin_rhs("brusselator1d")
    u0 = pr.brusselator_u0()
    te = np.linspace(0.0, 2.0, 21)
    for d in ([1.0, 3.0, 0.5], [1.1, 2.7, 5.0]):
        us, ok, cnt = run(lib, f, u0, te, d, 1e-7, 1e-9)
        ref, okr, st = orc.dop853(f, u0, te, d, 1e-7, 1e-9)
        assert ok == okr and cnt == ocnt(st) and np.array_equal(us, ref)
        assert cnt[0] > 5
    for kw in (dict(mxstep=0), dict(mxstep=15), dict(te=np.array([0.0])), dict(te=np.linspace(2, 0, 5)),
               dict(te=np.array([0.0, 1e-4, 1e-4, 2e-4, 0.5]))):
        te2 = kw.pop("te", te)
        us, ok, cnt = run(lib, f, u0, te2, [1.0, 3.0, 0.5], 1e-7, 1e-9, **kw)
        ref, okr, st = orc.dop853(f, u0, te2, [1.0, 3.0, 0.5], 1e-7, 1e-9, **kw)
        assert ok == okr and cnt[5] == st.idid and cnt[6] == st.rows, (kw, cnt, ocnt(st))
        if st.idid != -1:
            assert cnt == ocnt(st)
        assert np.array_equal(us[:st.rows], ref[:st.rows])
