"""Pin the LSODA restatement (oracle/lsoda_oracle.c) before trusting it:
 - against the reference's own known-answer tests,
 - bit-for-bit (states AND private counters) against the compiled, unmodified
   reference (oracle/_ref, built from /root/reference/src when present),
 - against the committed golden vectors produced by that compiled reference
   (tests/golden/lsoda_reference.npz), which is what runs on the GPU box."""
import os

import numpy as np
import pytest

import _oracle as orc
import _problems as pr

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "lsoda_reference.npz"))
CASES = ["lv_test", "lv_readme", "robertson", "lorenz", "lorenz_short"]


def _counters(st):
    return [st.nst, st.nfe, st.nje, st.nswitch, st.nqu, st.istate, st.rows, st.meth]


def test_reference_known_answer_lv():
    # /root/reference/tests/test_numbalsoda.py:11-20
    p = pr.problem("lv_test")
    us, ok, _ = orc.lsoda(pr.cfunc_address(p["rhs"]), p["u0"], p["t_eval"], p["data"], 1e-8, 1e-8)
    assert ok
    assert np.all(np.isclose(us[-1], np.array([0.38583246250193476, 4.602012234037773])))
    # the literals in the reference test are the lsoda output itself: bit-for-bit
    assert us[-1, 0] == 0.38583246250193476 and us[-1, 1] == 4.602012234037773


def test_reference_readme_example():
    # /root/reference/comparison2scipy.ipynb:175 shows usol[-1] of the README example
    p = pr.problem("lv_readme")
    us, ok, _ = orc.lsoda(pr.cfunc_address(p["rhs"]), p["u0"], p["t_eval"], p["data"])
    assert ok
    assert us[-1, 0] == 0.15204821320141523 and us[-1, 1] == 0.10552706409364211


@pytest.mark.parametrize("name", CASES)
def test_matches_golden_bitwise(name):
    p = pr.problem(name)
    us, ok, st = orc.lsoda(pr.cfunc_address(p["rhs"]), p["u0"], p["t_eval"], p["data"],
                           p["rtol"], p["atol"])
    assert ok == bool(GOLD[name + "/success"])
    assert _counters(st) == list(GOLD[name + "/counters"])
    assert np.array_equal(us, GOLD[name + "/usol"])


@pytest.mark.parametrize("name", CASES)
def test_builtin_c_rhs_equals_numba_cfunc(name):
    """oracle/rhs_builtin.c performs the same IEEE operations as numba's cfunc."""
    p = pr.problem(name)
    us, ok, st = orc.lsoda(orc.builtin_rhs(p["builtin"]), p["u0"], p["t_eval"], p["data"],
                           p["rtol"], p["atol"])
    assert _counters(st) == list(GOLD[name + "/counters"])
    assert np.array_equal(us, GOLD[name + "/usol"])


@pytest.mark.parametrize("eow", [False, True])
def test_exit_on_warning_flags(eow):
    # /root/reference/tests/test_exit_on_warning.py:22-50 (u0 is an int64 array there)
    f = pr.cfunc_address(pr.growth_rhs)
    u0 = np.array([1]).view(np.float64)
    us, ok, st = orc.lsoda(f, u0, np.linspace(100, 101, 11), np.array([1e15]), 1e-9, 1e-9,
                           exit_on_warning=eow)
    assert ok == (not eow)
    assert ok == bool(GOLD["growth_eow%d/success" % eow])
    assert _counters(st) == list(GOLD["growth_eow%d/counters" % eow])
    if ok:
        assert np.array_equal(us, GOLD["growth_eow%d/usol" % eow])


needs_ref = pytest.mark.skipif(not orc.HAVE_REF, reason="compiled reference (oracle/_ref) absent")


@needs_ref
@pytest.mark.parametrize("name", CASES)
def test_bitwise_against_compiled_reference(name):
    p = pr.problem(name)
    f = pr.cfunc_address(p["rhs"])
    a, oka, st = orc.lsoda(f, p["u0"], p["t_eval"], p["data"], p["rtol"], p["atol"])
    b, okb, cb = orc.ref_lsoda(f, p["u0"], p["t_eval"], p["data"], p["rtol"], p["atol"])
    c, okc = orc.ref_lsoda_wrapper(f, p["u0"], p["t_eval"], p["data"], p["rtol"], p["atol"])
    assert oka == okb == okc
    assert _counters(st) == cb
    assert np.array_equal(a, b) and np.array_equal(a, c)


@needs_ref
def test_randomised_sweeps_against_compiled_reference():
    """Parameter sweeps of the benchmark shapes, tolerances and
    horizons drawn at random: states bit-identical, counters identical."""
    rng = np.random.default_rng(7)
    fl = orc.builtin_rhs("lorenz")
    fr = orc.builtin_rhs("robertson")
    fv = orc.builtin_rhs("lotka_volterra")
    lp = pr.lorenz_sweep(12, seed=3)
    rp = pr.robertson_sweep(12, seed=3)
    for i in range(12):
        rtol = 10.0 ** rng.uniform(-10, -3)
        atol = 10.0 ** rng.uniform(-12, -5)
        for f, u0, data, te in (
                (fl, [1.0, 0.0, 0.0], lp[i], np.linspace(0, rng.uniform(2, 20), 57)),
                (fr, [1.0, 0.0, 0.0], rp[i], np.linspace(0, 10 ** rng.uniform(0, 5), 33)),
                (fv, [5.0, 0.8], np.array([rng.uniform(0.5, 2)]), np.linspace(0, 30, 200))):
            u0 = np.array(u0)
            a, oka, st = orc.lsoda(f, u0, te, data, rtol, atol)
            b, okb, cb = orc.ref_lsoda(f, u0, te, data, rtol, atol)
            assert oka == okb
            ca = _counters(st)
            # the shim sees method switches only at output granularity: a pair of
            # switches inside one output interval is invisible to it
            assert ca[3] >= cb[3] and (ca[3] - cb[3]) % 2 == 0
            ca[3] = cb[3]
            assert ca == cb
            assert np.array_equal(a, b)


@needs_ref
def test_edge_cases_against_compiled_reference():
    f = orc.builtin_rhs("lotka_volterra")
    u0 = np.array([5.0, 0.8])
    d = np.array([1.0])
    cases = [
        np.array([0.0]),                               # nt = 1: only row 0
        np.array([0.0, 0.0, 1.0]),                     # t_eval[1] == t_eval[0]: "tout too close"
        np.array([0.0, 1.0, 1.0, 1.0, 2.0]),           # later duplicates are fine
        np.array([0.0, 1.0, 0.5, 2.0]),                # non-monotone: fails at row 2
        np.array([0.0, 1e-3, 2e-3, 5.0]),              # several rows inside one step
        np.array([3.0, 4.0, 9.0]),                     # t0 != 0
    ]
    for te in cases:
        a, oka, st = orc.lsoda(f, u0, te, d, 1e-6, 1e-9)
        b, okb, cb = orc.ref_lsoda(f, u0, te, d, 1e-6, 1e-9)
        assert oka == okb, te
        rows = cb[6]
        assert st.rows == rows, te
        assert np.array_equal(a[:rows], b[:rows], equal_nan=True), te
    # mxstep is a per-output-interval budget (LSODA.cpp:778 with :671)
    te = np.linspace(0, 50, 6)
    for mx in (1, 5, 20, 60, 500):
        a, oka, st = orc.lsoda(f, u0, te, d, 1e-8, 1e-8, mxstep=mx)
        b, okb, cb = orc.ref_lsoda(f, u0, te, d, 1e-8, 1e-8, mxstep=mx)
        assert oka == okb and st.rows == cb[6] and st.nst == cb[0], mx
        assert np.array_equal(a[:cb[6]], b[:cb[6]])


def test_batch_driver_matches_single_calls():
    f = orc.builtin_rhs("robertson")
    P = pr.robertson_sweep(16, seed=1)
    te = np.linspace(0, 1e5, 100)
    us, ok, cnt = orc.solve_batch("lsoda", f, np.array([1.0, 0, 0]), te, P, 1e-8, 1e-8, nthreads=3)
    assert ok.all()
    for b in (0, 7, 15):
        a, oka, st = orc.lsoda(f, np.array([1.0, 0, 0]), te, P[b], 1e-8, 1e-8)
        assert np.array_equal(a, us[b]) and cnt[b, 0] == st.nst and cnt[b, 2] == st.nje
