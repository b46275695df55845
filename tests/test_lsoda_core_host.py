"""The LSODA device core (numbalsoda_b200/csrc/lsoda_core.cuh), compiled as host C++
by tests/lsoda_core_host.cpp, against the CPU oracle: states and the step / RHS /
Jacobian / switch counters must be identical bit for bit.  This validates the restated
control flow and storage scheme here, without a GPU; tests/test_gpu_lsoda.py then runs
the same comparisons through the CUDA kernel."""
import ctypes as ct
import math
import os
import subprocess

import numpy as np
import pytest

import _oracle as orc
import _problems as pr

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "lsoda_core_host.cpp")
HDRS = [os.path.join(ROOT, "numbalsoda_b200", "csrc", f)
        for f in ("lsoda_core.cuh", "lsoda_vec_thread.cuh", "lsoda_vec_warp.cuh", "lsoda_tables.cuh",
                  "b200ode_pow.cuh", "b200ode_fastmath.cuh") if os.path.exists(os.path.join(ROOT, "numbalsoda_b200", "csrc", f))]
_libs = {}


def core(n, fast=False):
    """n = system size for the thread backend, "warp" for the warp backend (any size).
    fast: the throughput link mode's arithmetic (b200ode_fastmath.cuh) with libm stand-ins."""
    key = (n, fast)
    if key not in _libs:
        os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
        so = os.path.join(HERE, "_build", "liblsoda_core_%s%s.so" % ("n%d" % n if n != "warp" else n,
                                                                       "_fast" if fast else ""))
        deps = [SRC] + HDRS
        if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
            flag = "-DB200_NEQ=%d" % n if n != "warp" else "-DB200_WARP_BACKEND"
            subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-mfma", "-fPIC", "-shared",
                                   "-Wno-unknown-pragmas", flag] + (["-DB200_HOST_FAST"] if fast else []) +
                                  ["-o", so, SRC])
        _libs[key] = ct.CDLL(so)
    return _libs[key]


BACKEND = {"thread": lambda n: core(n), "warp": lambda n: core("warp")}
_backend = "thread"


@pytest.fixture(params=["thread", "warp"], autouse=True)
def backend(request):
    """Every test runs on both vector backends of the LSODA core."""
    global _backend
    _backend = request.param
    yield request.param
    _backend = "thread"


def run_core(fptr, u0, te, data, rtol, atol, mxstep=10000, exit_on_warning=False, fast=False):
    u0 = np.ascontiguousarray(u0, dtype=np.float64)
    te = np.ascontiguousarray(te, dtype=np.float64)
    data = np.ascontiguousarray(data, dtype=np.float64)
    us = np.full((len(te), len(u0)), np.nan)
    ok = ct.c_int(999)
    cnt = (ct.c_int * 8)()
    p = lambda a: a.ctypes.data_as(ct.c_void_p)  # noqa: E731
    lib = core(len(u0) if _backend == "thread" else "warp", fast)
    lib.lsoda_core_host(ct.c_void_p(fptr), len(u0), p(u0), p(data), len(te), p(te),
                                  p(us), ct.c_double(rtol), ct.c_double(atol), int(mxstep),
                                  ct.byref(ok), int(exit_on_warning), cnt)
    return us, ok.value == 1, list(cnt)


def ocnt(st):
    return [st.nst, st.nfe, st.nje, st.nswitch, st.nqu, st.istate, st.rows, st.meth]


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


def _late_stiffness(t, u, du, p):
    # a fast oscillation followed closely (non-stiff), then, from t = p[0] on, pulled to it hard
    k = 1.0 if t < p[0] else 1.0e6
    du[0] = -k * (u[0] - math.cos(20.0 * t)) - 20.0 * math.sin(20.0 * t)
    du[1] = u[0] - u[1]


def test_method_switch_is_still_considered_after_40k_steps():
    """The reference decrements `icount` on every step (LSODA.cpp:1203) and consults methodswitch
    while it is negative; the core keeps it in a 16-bit field that must stop at -1, not wrap.
    More than 2^15 Adams steps, then the problem turns stiff: the switch to BDF has to come at
    the reference's step."""
    import numba
    from numba import types
    sig = types.void(types.double, types.CPointer(types.double), types.CPointer(types.double),
                     types.CPointer(types.double))
    cf = numba.cfunc(sig)(_late_stiffness)
    u0, data = np.array([1.0, 0.0]), np.array([150.0])
    te = np.array([0.0, 100.0, 149.0, 151.0, 160.0])
    ref, okr, st = orc.lsoda(cf.address, u0, te, data, rtol=1e-10, atol=1e-12, mxstep=200000)
    assert okr and st.nswitch >= 1 and st.meth == 2 and st.nst > 40000
    us, ok, cnt = run_core(cf.address, u0, te, data, 1e-10, 1e-12, mxstep=200000)
    assert ok and cnt == ocnt(st) and same(us, ref)


@pytest.mark.parametrize("name", ["lv_test", "lv_readme", "robertson", "lorenz", "lorenz_short"])
def test_reference_problems_bitwise(name):
    p = pr.problem(name)
    f = orc.builtin_rhs(p["builtin"])
    us, ok, cnt = run_core(f, p["u0"], p["t_eval"], p["data"], p["rtol"], p["atol"])
    ref, okr, st = orc.lsoda(f, p["u0"], p["t_eval"], p["data"], p["rtol"], p["atol"])
    assert ok == okr and cnt == ocnt(st)
    assert same(us, ref)


def test_tables_equal_cfode_of_the_oracle():
    """lsoda_tables.cuh (generated) against what the reference's cfode arithmetic gives:
    if a coefficient were off by one ulp the parity tests above would fail, but this
    pins the tables directly -- Adams elco[nq][1] and tesco, BDF likewise -- through the
    generator's independent transcription."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen", os.path.join(ROOT, "tools", "gen_lsoda_tables.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    e1, t1, e2, t2, cm1, cm2 = gen.tables()

    class T(ct.Structure):
        _fields_ = [("elco", ct.c_double * 14 * 13 * 2), ("tesco", ct.c_double * 4 * 13 * 2),
                    ("cm1", ct.c_double * 13), ("cm2", ct.c_double * 6), ("sm1", ct.c_double * 13),
                    ("conit", ct.c_double * 13), ("rcp", ct.c_double * 16),
                    ("cm12", ct.c_double * 13), ("cm21", ct.c_double * 13), ("pad", ct.c_double * 1)]
    lib = BACKEND[_backend](2)
    lib.lsoda_core_tables.restype = ct.POINTER(T)
    t = lib.lsoda_core_tables().contents
    assert np.array_equal(np.array(t.elco[0]), np.array(e1))
    assert np.array_equal(np.array(t.elco[1]), np.array(e2))
    assert np.array_equal(np.array(t.tesco[0]), np.array(t1))
    assert np.array_equal(np.array(t.tesco[1]), np.array(t2))
    assert list(t.cm1) == cm1 and list(t.cm2) == cm2
    assert list(t.conit)[1:] == [0.5 / (q + 2) for q in range(1, 13)]
    assert list(t.rcp)[1:] == [1.0 / k for k in range(1, 16)]
    assert list(t.cm12)[1:6] == [cm1[q] / cm2[q] for q in range(1, 6)]
    assert list(t.cm21)[1:6] == [cm2[q] / cm1[q] for q in range(1, 6)]
    # spot values every LSODA implementation shares
    assert t.elco[1][1][1] == 1.0 and t.elco[1][2][1] == 2.0 / 3.0 and t.elco[0][2][1] == 0.5


@pytest.mark.parametrize("problem,B", [("robertson", 300), ("lorenz", 40), ("lv", 200)])
def test_parameter_sweeps_bitwise(problem, B):
    rng = np.random.default_rng(11)
    if problem == "robertson":
        P, u0, te, name, tol = pr.robertson_sweep(B), [1.0, 0, 0], np.linspace(0, 1e5, 100), "robertson", 1e-8
    elif problem == "lorenz":
        P, u0, te, name, tol = pr.lorenz_sweep(B), [1.0, 0, 0], np.linspace(0, 20, 101), "lorenz", 1e-8
    else:
        P, u0, te, name, tol = rng.uniform(0.5, 3.0, (B, 1)), [5.0, 0.8], np.linspace(0, 30, 61), "lotka_volterra", 1e-6
    f = orc.builtin_rhs(name)
    for b in range(B):
        rt = tol * 10.0 ** rng.integers(-2, 3)
        us, ok, cnt = run_core(f, u0, te, P[b], rt, tol)
        ref, okr, st = orc.lsoda(f, u0, te, P[b], rt, tol)
        assert ok == okr and cnt == ocnt(st), (b, cnt, ocnt(st))
        assert same(us[:st.rows], ref[:st.rows]), b


def test_failures_and_edge_cases_match_oracle():
    f = orc.builtin_rhs("lotka_volterra")
    u0, d = [5.0, 0.8], [1.0]
    cases = [
        dict(te=np.linspace(0, 10, 11), rtol=1e-8, atol=1e-8, mxstep=5),       # istate -1
        dict(te=np.linspace(0, 10, 11), rtol=1e-8, atol=1e-8, mxstep=0),
        dict(te=np.linspace(0, 10, 11), rtol=1e-8, atol=1e-8, mxstep=-1),
        dict(te=np.array([0.0]), rtol=1e-8, atol=1e-8),                        # nt = 1
        dict(te=np.array([0.0, 0.0, 1.0]), rtol=1e-8, atol=1e-8),              # tout too close
        dict(te=np.array([0.0, 1.0, 1.0, 2.0, 2.0]), rtol=1e-8, atol=1e-8),    # duplicates
        dict(te=np.array([0.0, 1.0, 0.5, 2.0]), rtol=1e-8, atol=1e-8),         # not monotone
        dict(te=np.array([1.0, 0.0]), rtol=1e-8, atol=1e-8),                   # decreasing
        dict(te=np.linspace(0, 10, 11), rtol=-1.0, atol=1e-8),                 # illegal tolerance
        dict(te=np.linspace(0, 10, 11), rtol=0.0, atol=0.0),                   # ewt <= 0 quirk
        dict(te=np.linspace(0, 10, 11), rtol=0.0, atol=1e-10),
        dict(te=np.linspace(0, 10, 11), rtol=1e-17, atol=1e-20),               # too much accuracy
        dict(te=np.array([0.0, 1e-9, 2e-9, 5.0, 5.0 + 1e-9]), rtol=1e-6, atol=1e-9),
    ]
    for c in cases:
        kw = dict(rtol=c["rtol"], atol=c["atol"], mxstep=c.get("mxstep", 10000))
        us, ok, cnt = run_core(f, u0, c["te"], d, **kw)
        ref, okr, st = orc.lsoda(f, u0, c["te"], d, **kw)
        assert ok == okr and cnt == ocnt(st), (c, cnt, ocnt(st))
        assert same(us[:st.rows], ref[:st.rows]), c


def test_exit_on_warning_like_the_reference_test():
    # /root/reference/tests/test_exit_on_warning.py: du = p0*u*sin(t), p0 = 1e15
    import numba  # noqa: F401
    fptr = pr.cfunc_address(pr.growth_rhs)
    u0 = np.array([1]).view(np.float64) if False else np.array([4.9e-324])  # int64 1 reinterpreted
    te = np.linspace(100, 101, 11)
    d = np.array([1e15])
    for eow in (True, False):
        us, ok, cnt = run_core(fptr, u0, te, d, 1e-9, 1e-9, exit_on_warning=eow)
        ref, okr, st = orc.lsoda(fptr, u0, te, d, 1e-9, 1e-9, exit_on_warning=eow)
        assert ok == okr == (not eow)
        assert cnt == ocnt(st) and same(us[:st.rows], ref[:st.rows])


def test_stiff_problems_with_pivoting_and_failures():
    """Van der Pol (mu = 1000) and a stiff linear system with large off-diagonal
    Jacobian entries (exercises the integer-truncating pivot search with column swaps),
    plus Robertson with a tolerance so loose that error-test failures pile up."""
    import numba as nb
    from numba import types
    sig = types.void(types.double, types.CPointer(types.double), types.CPointer(types.double),
                     types.CPointer(types.double))

    @nb.cfunc(sig)
    def vdp(t, u, du, p):
        du[0] = u[1]
        du[1] = p[0] * ((1.0 - u[0] * u[0]) * u[1]) - u[0]

    @nb.cfunc(sig)
    def lin3(t, u, du, p):
        du[0] = -p[0] * u[0] + p[1] * u[1]
        du[1] = p[0] * u[0] - p[1] * u[1] - p[2] * u[1] + 3.0 * u[2]
        du[2] = p[2] * u[1] - 3.0 * u[2]

    runs = [(vdp.address, [2.0, 0.0], [1000.0], np.linspace(0, 3000, 61), 1e-6, 1e-8),
            (vdp.address, [2.0, 0.0], [50.0], np.linspace(0, 200, 41), 1e-4, 1e-6),
            (lin3.address, [1.0, 0.0, 0.0], [1e4, 2e6, 5e3], np.linspace(0, 10, 21), 1e-7, 1e-10),
            (lin3.address, [1.0, 2.0, 3.0], [1e-3, 7e5, 1e7], np.linspace(0, 100, 21), 1e-5, 1e-9),
            (orc.builtin_rhs("robertson"), [1.0, 0, 0], [0.04, 3e7, 1e4], np.linspace(0, 1e11, 12), 1e-2, 1e-4),
            (orc.builtin_rhs("robertson"), [1.0, 0, 0], [0.04, 3e7, 1e4], np.linspace(0, 1e11, 12), 1e-10, 1e-14)]
    for f, u0, d, te, rtol, atol in runs:
        us, ok, cnt = run_core(f, u0, te, d, rtol, atol)
        ref, okr, st = orc.lsoda(f, u0, te, d, rtol, atol)
        assert ok == okr and cnt == ocnt(st), (cnt, ocnt(st))
        assert same(us[:st.rows], ref[:st.rows])


def test_brusselator_n64_bitwise():
    """BASELINE config 4's problem (n = 64 method of lines, stiff): warp backend only."""
    if _backend != "warp":
        pytest.skip("n = 64 is the warp backend's range")
    f = orc.builtin_rhs("brusselator1d")
    rng = np.random.default_rng(4)
    te = np.linspace(0.0, 10.0, 21)
    for k in range(6):
        A, B, D = 1.0 + 0.2 * rng.uniform(-1, 1), 3.0 + 0.5 * rng.uniform(-1, 1), 10.0 ** rng.uniform(1.5, 3.5)
        x = (np.arange(32) + 1) / 33.0
        u0 = np.concatenate([A + 0.5 * np.sin(np.pi * x), B / A + 0.5 * np.sin(2 * np.pi * x)])
        us, ok, cnt = run_core(f, u0, te, [A, B, D], 1e-6, 1e-8)
        ref, okr, st = orc.lsoda(f, u0, te, [A, B, D], 1e-6, 1e-8)
        assert ok == okr and cnt == ocnt(st), (k, cnt, ocnt(st))
        assert same(us[:st.rows], ref[:st.rows]), k
        assert cnt[2] > 0 and cnt[7] == 2  # Jacobians were needed: it is stiff


def test_warp_backend_fast_solves_close_to_reference_order():
    """The fast mode's triangular solves (full column interchanges, swap-free column
    sweep, reciprocal diagonal) are the same mathematics in another summation order:
    on stiff problems whose LU pivots, results stay within a few error-weight units of
    the reference-order run and the step counts stay close."""
    if _backend != "warp":
        pytest.skip("warp backend only")
    f = orc.builtin_rhs("brusselator1d")
    te = np.linspace(0.0, 10.0, 21)
    P = pr.brusselator_sweep(6, seed=9)
    u0 = np.ascontiguousarray(pr.brusselator_u0())
    p = lambda a: a.ctypes.data_as(ct.c_void_p)  # noqa: E731
    for b in range(6):
        d = np.ascontiguousarray(P[b])
        out = []
        for flag in (0, 2):
            us = np.full((len(te), 64), np.nan)
            ok = ct.c_int(999)
            cnt = (ct.c_int * 8)()
            core("warp").lsoda_core_host(ct.c_void_p(f), 64, p(u0), p(d), len(te), p(te), p(us),
                                         ct.c_double(1e-6), ct.c_double(1e-8), 10000, ct.byref(ok),
                                         flag, cnt)
            out.append((us, ok.value, list(cnt)))
        (us0, ok0, c0), (us1, ok1, c1) = out
        assert ok0 == ok1 == 1
        err = np.abs(us1 - us0) / (1e-6 * np.abs(us0) + 1e-8)
        assert err.max() < 5.0, err.max()
        assert abs(c1[0] - c0[0]) <= 0.1 * c0[0] and c1[2] > 0


@pytest.mark.parametrize("n", [3, 12, 40])
def test_pivoting_problem_bitwise_and_fast_solves(n):
    """A system whose LU really interchanges columns: bit parity of both backends with
    the oracle, and the warp backend's fast solves (interchanges applied to the
    multipliers + one final permutation) against the reference-order result."""
    if _backend == "thread" and n > 3:
        pytest.skip("thread backend: n <= 8")
    fptr = pr.cfunc_address(pr.upper_rhs)
    u0 = np.ones(n)
    d = np.array([1000.0, 5000.0, float(n)])
    te = np.linspace(0.0, 2.0, 9)
    us, ok, cnt = run_core(fptr, u0, te, d, 1e-6, 1e-9)
    ref, okr, st = orc.lsoda(fptr, u0, te, d, 1e-6, 1e-9)
    assert ok == okr and cnt == ocnt(st) and same(us[:st.rows], ref[:st.rows])
    assert cnt[2] > 0 and cnt[7] == 2
    if _backend == "warp":
        p = lambda a: a.ctypes.data_as(ct.c_void_p)  # noqa: E731
        us1 = np.full((len(te), n), np.nan)
        okc, c1 = ct.c_int(999), (ct.c_int * 8)()
        core("warp").lsoda_core_host(ct.c_void_p(fptr), n, p(u0), p(d), len(te), p(te), p(us1),
                                     ct.c_double(1e-6), ct.c_double(1e-9), 10000, ct.byref(okc), 2, c1)
        assert okc.value == 1
        err = np.abs(us1 - ref) / (1e-6 * np.abs(ref) + 1e-9)
        assert err.max() < 20.0, err.max()


@pytest.mark.parametrize("n", [4, 5, 6, 7, 8])
def test_every_thread_kernel_size_bitwise(n):
    """The thread backend is compiled per system size (n = 1..8): sizes 4..8 on a stiff
    chain and on the pivoting problem, both backends."""
    for rhs, d, te, tol in ((pr.chain_rhs, [80.0, float(n)], np.linspace(0, 5, 11), (1e-7, 1e-9)),
                            (pr.upper_rhs, [1000.0, 5000.0, float(n)], np.linspace(0, 2, 9), (1e-6, 1e-9))):
        fptr = pr.cfunc_address(rhs)
        u0 = 1.0 + 0.1 * np.arange(n)
        us, ok, cnt = run_core(fptr, u0, te, d, *tol)
        ref, okr, st = orc.lsoda(fptr, u0, te, d, *tol)
        assert ok == okr and cnt == ocnt(st), (n, cnt, ocnt(st))
        assert same(us[:st.rows], ref[:st.rows])


def test_fast_mode_robertson_close_to_oracle():
    """The throughput link mode (divisions by reciprocal, single-precision step / order /
    method ratios; libm stand-ins on the host) against the oracle on config C3's sweep: the
    same bound as the GPU fast-mode test (tests/test_gpu_lsoda.py), no bias in the step count."""
    B = 64
    P = pr.robertson_sweep(B)
    f = orc.builtin_rhs("robertson")
    u0 = np.array([1.0, 0.0, 0.0])
    te = np.linspace(0, 1e5, 100)
    worst, dn = 0.0, []
    for b in range(B):
        ref, okr, st = orc.lsoda(f, u0, te, P[b], 1e-8, 1e-8)
        us, ok, cnt = run_core(f, u0, te, P[b], 1e-8, 1e-8, fast=True)
        assert ok and okr
        worst = max(worst, (np.abs(us - ref) / (1e-8 * np.abs(ref) + 1e-8)).max())
        dn.append(cnt[0] - st.nst)
    assert worst <= 100.0
    assert abs(np.mean(dn)) < 10 and np.abs(dn).mean() < 40


def test_fast_mode_failure_paths_match():
    """Failure semantics do not depend on the arithmetic mode."""
    f = orc.builtin_rhs("lotka_volterra")
    u0, d = np.array([5.0, 0.8]), np.array([1.0])
    for kw in (dict(te=np.linspace(0, 10, 11), rtol=1e-8, atol=1e-8, mxstep=5),
               dict(te=np.array([0.0, 0.0, 1.0]), rtol=1e-8, atol=1e-8),
               dict(te=np.array([0.0, 1.0, 0.5, 2.0]), rtol=1e-8, atol=1e-8),
               dict(te=np.linspace(0, 10, 11), rtol=0.0, atol=0.0),
               dict(te=np.linspace(0, 10, 11), rtol=1e-17, atol=1e-20)):
        te = kw.pop("te")
        ref, okr, st = orc.lsoda(f, u0, te, d, **kw)
        us, ok, cnt = run_core(f, u0, te, d, fast=True, **kw)
        assert ok == okr and cnt[5] == st.istate and cnt[6] == st.rows
