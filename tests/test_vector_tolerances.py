"""Per-component tolerances: rtol / atol of shape [n] or [B,n].

The reference's drivers pass scalars (numbalsoda/driver.py:29), but its solvers implement one
value per component: LSODA's itol_ = 3 / 4 (LSODA.cpp:500-520, :624-640, :1368-1390) and
dop853's itol = 1 (dop853_module.f90:603-612, :779-803).  Pins, without a GPU:
  * the LSODA oracle against the COMPILED REFERENCE driven with its private tolerance vectors
    filled per component and itol_ = 4 (oracle/ref_shim.cpp): states and counters bit for bit;
  * the DOP853 oracle against scipy's dop853 (Hairer's code, vector tolerances there too);
  * the device cores compiled as host C++ against the oracle, bit for bit.
tests/test_gpu_vector_tolerances.py runs the kernels against the oracle."""
import ctypes as ct

import numpy as np
import pytest

import _oracle as orc
import _problems as pr


def _tols(n, seed, lo=(-9, -5), hi=(-12, -7)):
    rng = np.random.default_rng(seed)
    return 10.0 ** rng.uniform(*lo, n), 10.0 ** rng.uniform(*hi, n)


@pytest.mark.skipif(not orc.HAVE_REF, reason="compiled reference not available")
@pytest.mark.parametrize("name", ["robertson", "lv_test", "lorenz_short"])
def test_lsoda_oracle_equals_the_compiled_reference_with_itol_4(name):
    p = pr.problem(name)
    f = orc.builtin_rhs(p["builtin"])
    n = len(p["u0"])
    for seed in range(6):
        rv, av = _tols(n, seed)
        ref, okr, st = orc.ref_lsoda_vtol(f, p["u0"], p["t_eval"], p["data"], rv, av)
        us, ok, cnt = orc.solve_batch("lsoda", f, p["u0"], p["t_eval"], p["data"][None, :], rv, av)
        assert ok[0] == okr and list(cnt[0]) == st
        assert np.array_equal(us[0], ref, equal_nan=True)
    # zero / negative entries take the reference's paths too (ewt <= 0 quirk, illegal input)
    for rv, av in ((np.zeros(n), np.zeros(n)), (np.full(n, 1e-6), np.r_[-1e-9, np.full(n - 1, 1e-9)]),
                   (np.zeros(n), np.full(n, 1e-9))):
        ref, okr, st = orc.ref_lsoda_vtol(f, p["u0"], p["t_eval"], p["data"], rv, av)
        us, ok, cnt = orc.solve_batch("lsoda", f, p["u0"], p["t_eval"], p["data"][None, :], rv, av)
        # (the reference's counters are uninitialised memory when it fails before its first step,
        # and it leaves istate = 1 behind on its ewt <= 0 path: compare outcome and rows)
        assert ok[0] == okr and cnt[0][6] == st[6]
        assert np.array_equal(us[0, :st[6]], ref[:st[6]], equal_nan=True)


def test_uniform_vectors_equal_the_scalars_bitwise():
    for solver, name in (("lsoda", "robertson"), ("dop853", "lorenz_short")):
        p = pr.problem(name)
        f = orc.builtin_rhs(p["builtin"])
        n = len(p["u0"])
        a = orc.solve_batch(solver, f, p["u0"], p["t_eval"], p["data"][None, :], p["rtol"], p["atol"])
        b = orc.solve_batch(solver, f, p["u0"], p["t_eval"], p["data"][None, :],
                            np.full(n, p["rtol"]), np.full(n, p["atol"]))
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[2], b[2])


def _scaled_lorenz(t, u, du, p):
    # Lorenz in the variables u_i / s_i, s = p[3:6]
    x = u[0] * p[3]
    y = u[1] * p[4]
    z = u[2] * p[5]
    du[0] = (p[0] * (y - x)) / p[3]
    du[1] = (x * (p[1] - z) - y) / p[4]
    du[2] = (x * y - p[2] * z) / p[5]


@pytest.mark.parametrize("solver", ["dop853", "lsoda"])
def test_per_component_atol_equals_a_power_of_two_rescaling(solver):
    """Scale invariance pins the vector path through the (pinned) scalar path.  With s_i powers of
    two, the system written in u_i / s_i and solved with the scalar atol performs the same
    floating-point operations, exactly scaled, as the original system solved with
    atol_i = atol * s_i: every weight atol_i + rtol |y_i| is s_i times the scaled system's, so
    error norms, step sizes and counters are identical and the states agree bit for bit after
    scaling back.  (scipy 1.18's dop853 wrapper takes scalars only, so it cannot serve here.)"""
    p = pr.problem("lorenz_short")
    sc = np.array([4.0, 0.25, 64.0])
    f = orc.builtin_rhs("lorenz")
    fs = pr.cfunc_address(_scaled_lorenz)
    rtol, atol = 1e-7, 1e-9
    a = orc.solve_batch(solver, f, p["u0"], p["t_eval"], p["data"][None, :], rtol, atol * sc)
    b = orc.solve_batch(solver, fs, p["u0"] / sc, p["t_eval"], np.r_[p["data"], sc][None, :], rtol, atol)
    assert a[1][0] and b[1][0]
    assert np.array_equal(a[2], b[2])  # counters
    assert np.array_equal(a[0], b[0] * sc)
    # and the vector matters: a scalar atol gives another step sequence
    c = orc.solve_batch(solver, f, p["u0"], p["t_eval"], p["data"][None, :], rtol, atol)
    assert not np.array_equal(a[2], c[2])


@pytest.mark.parametrize("backend", ["thread", "warp"])
def test_lsoda_device_core_on_the_host(backend):
    import test_lsoda_core_host as host
    f = orc.builtin_rhs("robertson")
    u0, te = np.array([1.0, 0.0, 0.0]), np.linspace(0, 1e5, 100)
    lib = host.core(3 if backend == "thread" else "warp")
    p = lambda a: a.ctypes.data_as(ct.c_void_p)  # noqa: E731
    for seed in range(5):
        P = pr.robertson_sweep(1, seed)[0]
        rv, av = _tols(3, seed)
        ref, okr, cr = orc.solve_batch("lsoda", f, u0, te, P[None, :], rv, av)
        for rvec, avec, rs, as_ in ((rv, av, 0.0, 0.0), (rv, None, 0.0, 1e-9), (None, av, 1e-7, 0.0)):
            if rvec is None or avec is None:
                ref, okr, cr = orc.solve_batch("lsoda", f, u0, te, P[None, :],
                                               rv if rvec is not None else rs, av if avec is not None else as_)
            us = np.full((len(te), 3), np.nan)
            ok = ct.c_int(999)
            cnt = (ct.c_int * 8)()
            lib.lsoda_core_host_vtol(ct.c_void_p(f), 3, p(u0), p(P), len(te), p(te), p(us), ct.c_double(rs),
                                     ct.c_double(as_), p(rvec) if rvec is not None else None,
                                     p(avec) if avec is not None else None, 10000, ct.byref(ok), 0, cnt)
            assert (ok.value == 1) == bool(okr[0]) and list(cnt) == list(cr[0])
            assert np.array_equal(us, ref[0])


def test_dop853_warp_core_on_the_host():
    import test_dop853_warp_host as host
    lib = host.build_lib()
    f = orc.builtin_rhs("brusselator1d")
    n = 64
    u0, te, d = pr.brusselator_u0(), np.linspace(0, 2, 21), np.array([1.0, 3.0, 0.5])
    rv, av = _tols(n, 3, lo=(-8, -5), hi=(-10, -7))
    ref, okr, cr = orc.solve_batch("dop853", f, u0, te, d[None, :], rv, av)
    p = lambda a: a.ctypes.data_as(ct.c_void_p)  # noqa: E731
    us = np.full((len(te), n), np.nan)
    ok = ct.c_int(999)
    cnt = (ct.c_int * 8)()
    lib.dop853_warp_host_vtol(ct.c_void_p(f), n, p(u0), p(d), len(te), p(te), p(us), ct.c_double(0.0),
                              ct.c_double(0.0), p(rv), p(av), 10000, ct.byref(ok), cnt)
    assert ok.value == 1 and okr[0] and list(cnt)[:7] == list(cr[0][:7])
    assert np.array_equal(us, ref[0])


def test_tolerance_layout_of_the_python_driver():
    from numbalsoda_b200._solve import tolerance_layout
    from numbalsoda_b200 import _lib
    assert tolerance_layout(1e-6, 5, 3, "rtol") == (1e-6, None, _lib.TOL_PER_TRAJECTORY)
    s, a, d = tolerance_layout(np.full(5, 1e-6), 5, 3, "rtol")
    assert d == _lib.TOL_PER_TRAJECTORY and a.shape == (5,)
    s, a, d = tolerance_layout(np.full(3, 1e-6), 5, 3, "rtol")
    assert d == _lib.TOL_PER_COMPONENT and a.shape == (3,)
    s, a, d = tolerance_layout(np.full((1, 3), 1e-6), 3, 3, "rtol")
    assert d == _lib.TOL_PER_COMPONENT and a.shape == (3,)
    s, a, d = tolerance_layout(np.full((5, 3), 1e-6), 5, 3, "rtol")
    assert d == _lib.TOL_FULL
    s, a, d = tolerance_layout(np.full(3, 1e-6), 3, 3, "rtol")  # ambiguous: per trajectory
    assert d == _lib.TOL_PER_TRAJECTORY
    s, a, d = tolerance_layout(np.full(3, 1e-6), 1, 3, "rtol", batched=False)  # one trajectory: per component
    assert d == _lib.TOL_PER_COMPONENT
    with pytest.raises(ValueError):
        tolerance_layout(np.full(4, 1e-6), 5, 3, "rtol")
