"""Pin the DOP853 restatement (oracle/dop853_oracle.c).  The reference's DOP853
is Fortran and cannot be compiled in this image, so the pins are:
 - the reference's known-answer test (tests/test_numbalsoda.py:22-31),
 - an independent implementation: scipy.integrate.ode('dop853') (a translation of
   Hairer's original code) configured with the reference's controller constants
   (dop853_module.f90:55-62) -- identical accepted-step sequences and states,
 - the committed golden vectors (tests/golden/dop853_oracle.npz)."""
import ctypes as ct
import os

import numpy as np
import pytest

import _oracle as orc
import _problems as pr

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "dop853_oracle.npz"))
CASES = ["lv_test", "lv_readme", "lorenz", "lorenz_short", "robertson"]


def _counters(st):
    return [st.nstep, st.nfcn, st.naccpt, st.nrejct, st.nevent, st.idid, st.rows]


def test_reference_known_answer_lv():
    # /root/reference/tests/test_numbalsoda.py:22-31
    p = pr.problem("lv_test")
    us, ok, _ = orc.dop853(pr.cfunc_address(p["rhs"]), p["u0"], p["t_eval"], p["data"], 1e-8, 1e-8)
    assert ok
    assert np.all(np.isclose(us[-1], np.array([0.38583246250193476, 4.602012234037773])))
    # and it is in fact much closer to the true solution than that test requires
    from scipy.integrate import solve_ivp
    s = solve_ivp(lambda t, u: [u[0] - u[0] * u[1], u[0] * u[1] - u[1]], [0, 10], [5.0, 0.8],
                  method="DOP853", rtol=1e-13, atol=1e-14)
    assert np.max(np.abs(us[-1] / s.y[:, -1] - 1)) < 1e-7
    assert np.array_equal(us[0], p["u0"])  # row 0 = contd8 at s = 0 = y0 exactly


@pytest.mark.parametrize("name", CASES)
def test_matches_golden_bitwise(name):
    p = pr.problem(name)
    for f in (pr.cfunc_address(p["rhs"]), orc.builtin_rhs(p["builtin"])):
        us, ok, st = orc.dop853(f, p["u0"], p["t_eval"], p["data"], p["rtol"], p["atol"])
        assert ok == bool(GOLD[name + "/success"])
        assert _counters(st) == list(GOLD[name + "/counters"])
        rows = st.rows
        assert np.array_equal(us[:rows], GOLD[name + "/usol"][:rows])


def _scipy_accepted(pyf, u0, t1, rtol, atol):
    from scipy.integrate import ode
    r = ode(pyf).set_integrator("dop853", rtol=rtol, atol=atol, nsteps=100000, dfactor=0.333,
                                ifactor=6.0, safety=0.9, beta=0.0)
    acc = []
    r.set_solout(lambda t, y: acc.append((t, y.copy())))
    r.set_initial_value(u0, 0.0)
    r.integrate(t1)
    return np.array([a[0] for a in acc]), np.array([a[1] for a in acc])


@pytest.mark.parametrize("name,t1,rtol,atol,tol", [
    ("lv_test", 10.0, 1e-8, 1e-8, 1e-13),
    ("lv_readme", 50.0, 1e-3, 1e-6, 1e-12),
    ("lorenz", 15.0, 1e-8, 1e-8, 1e-11),
    ("lorenz", 100.0, 1e-8, 1e-8, 1e-9),   # chaotic: 3200 steps amplify the last-bit differences
])
def test_against_scipy_hairer_dop853(name, t1, rtol, atol, tol):
    """With scipy's rejection rule emulated (test-only knob, see dop853_oracle.c)
    the oracle takes exactly scipy's accepted steps and lands on its states; up to
    the first rejection it does so without the knob as well."""
    p = pr.problem(name)
    data = p["data"]

    def pyf(t, u):
        du = np.zeros(len(u))
        p["rhs"](t, u, du, data)
        return du

    ts, ys = _scipy_accepted(pyf, p["u0"], t1, rtol, atol)
    f = orc.builtin_rhs(p["builtin"])
    orc.lib.oracle_dop853_set_scipy_reject_rule(1)
    try:
        us, ok, st = orc.dop853(f, p["u0"], ts, data, rtol, atol, mxstep=100000)
    finally:
        orc.lib.oracle_dop853_set_scipy_reject_rule(0)
    assert ok and st.naccpt == len(ts) - 1 and st.rows == len(ts)
    assert np.max(np.abs(us - ys) / (1 + np.abs(ys))) < tol
    # reference rule: identical prefix until the first rejected step
    us2, ok2, st2 = orc.dop853(f, p["u0"], ts, data, rtol, atol, mxstep=100000)
    same = np.all(np.abs(us2 - ys) <= 1e-13 * (1 + np.abs(ys)), axis=1)
    assert same[:2].all()  # t0 and the first accepted step (first rejection may follow)


def test_wrapper_semantics():
    f = orc.builtin_rhs("lotka_volterra")
    u0 = np.array([5.0, 0.8])
    d = np.array([1.0])
    te = np.linspace(0, 10, 11)
    # nmax <= 0 is rejected by set_parameters (dop853_module.f90:243-246)
    _, ok, st = orc.dop853(f, u0, te, d, 1e-8, 1e-8, mxstep=0)
    assert not ok and st.idid == -1
    # nmax counts attempted steps over the whole span (dop853_module.f90:539)
    _, ok, st = orc.dop853(f, u0, te, d, 1e-8, 1e-8, mxstep=20)
    assert not ok and st.idid == -2 and st.nstep == 21
    _, ok, st = orc.dop853(f, u0, te, d, 1e-8, 1e-8, mxstep=51)
    assert ok and st.nstep == 51
    # t_eval[0] == t_eval[-1]: h = 0 -> "step size too small" (dop853_module.f90:546)
    _, ok, st = orc.dop853(f, u0, np.array([0.0]), d, 1e-8, 1e-8)
    assert not ok and st.idid == -3
    # decreasing t_eval: integrates backwards, solout never fires (c_interface :60)
    us, ok, st = orc.dop853(f, u0, np.linspace(10, 0, 5), d, 1e-8, 1e-8)
    assert ok and st.rows == 0
    # dense rows inside one step + duplicates
    us, ok, st = orc.dop853(f, u0, np.array([0.0, 1e-4, 1e-4, 2e-4, 1.0]), d, 1e-8, 1e-8)
    assert ok and st.rows == 5 and np.array_equal(us[1], us[2])


def test_overflowing_trial_step_recovers():
    """err = NaN must be rejected with the maximal shrink (gfortran MIN/MAX ignore
    NaN operands); Robertson over [0, 40] blows up on its first trial steps."""
    f = orc.builtin_rhs("robertson")
    us, ok, st = orc.dop853(f, np.array([1.0, 0, 0]), np.linspace(0, 40, 5),
                            np.array([0.04, 3e7, 1e4]), 1e-8, 1e-8, mxstep=100000)
    assert ok and st.rows == 5 and np.isfinite(us).all()
    assert abs(us[-1].sum() - 1.0) < 1e-8  # mass conservation


def test_batch_driver():
    f = orc.builtin_rhs("lorenz")
    P = pr.lorenz_sweep(8)
    te = np.linspace(0, 5, 51)
    us, ok, cnt = orc.solve_batch("dop853", f, np.array([1.0, 0, 0]), te, P, 1e-8, 1e-8, nthreads=2)
    assert ok.all()
    a, _, st = orc.dop853(f, np.array([1.0, 0, 0]), te, P[5], 1e-8, 1e-8)
    assert np.array_equal(a, us[5]) and cnt[5, 0] == st.nstep


@pytest.mark.parametrize("rtol,atol", [(1e-3, 1e-6), (1e-6, 1e-8), (1e-8, 1e-8), (1e-10, 1e-12)])
def test_throughput_mode_controller_stays_within_tolerance(rtol, atol):
    """The product's default link mode takes the step-size controller through reciprocals and a
    single-precision power.  Emulated in the oracle (test-only knob): the
    solution moves by a small fraction of rtol and the accept / reject sequence is unchanged on
    the reference's Lotka-Volterra problem; a Lorenz sweep (short horizon) stays within 1e-10."""
    f = orc.builtin_rhs("lotka_volterra")
    te = np.linspace(0, 50, 1000)
    ref, ok, st = orc.dop853(f, [5.0, 0.8], te, [1.0], rtol, atol)
    orc.lib.oracle_dop853_set_fast_ctrl(1)
    try:
        us, ok2, st2 = orc.dop853(f, [5.0, 0.8], te, [1.0], rtol, atol)
        lz = [orc.dop853(orc.builtin_rhs("lorenz"), [1.0, 0.0, 0.0], np.linspace(0, 5, 51), P, 1e-8, 1e-8)
              for P in pr.lorenz_sweep(20)]
    finally:
        orc.lib.oracle_dop853_set_fast_ctrl(0)
    assert ok and ok2
    assert (st.nstep, st.naccpt, st.nrejct) == (st2.nstep, st2.naccpt, st2.nrejct)
    assert (np.abs(us - ref) / np.abs(ref)).max() <= 0.1 * rtol
    for (a, oka, _), P in zip(lz, pr.lorenz_sweep(20)):
        b, okb, _ = orc.dop853(orc.builtin_rhs("lorenz"), [1.0, 0.0, 0.0], np.linspace(0, 5, 51), P, 1e-8, 1e-8)
        assert oka and okb and (np.abs(a - b) / (np.abs(b) + 1.0)).max() <= 1e-10
