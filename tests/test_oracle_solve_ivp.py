"""oracle/solve_ivp_oracle.c (CPU restatement of the reference's solve_ivp: DOP853 with dense
output after every step, terminal events by Brent's method) against what the reference itself
printed and asserts:

  * /root/reference/solve_ivp.ipynb cell 11: t_event = 0.7611343988149695, y_event =
    [1., 5.02890877], message, ind_event; cell 26: t_event = 0.7258968839312535; cell 27: ten
    event times of a parameter sweep -- all printed with 16 digits by the reference;
  * tests/test_numbalsoda.py:33-56 (end point with and without t_eval);
  * solve_ivp.ipynb cell 6: forward and backward integration against scipy's DOP853.

The event times are Brent iterates with tol = 1e-8: two runs that took a different iteration
path would differ at the 1e-9 level.  The oracle reproduces all twelve to a few ulp (the
notebook was produced by a gfortran build on the author's machine -- FMA contraction and libm
differ from this image's -ffp-contract=off build, so the last digits are not expected to
match); the tests ask for 1e-13.
"""
import numpy as np
import pytest

import _oracle as orc
import _problems as pr

nb = pytest.importorskip("numba")
from numba import cfunc, types  # noqa: E402

SIG = types.void(types.double, types.CPointer(types.double), types.CPointer(types.double),
                 types.CPointer(types.double))


@cfunc(SIG)
def lv(t, u, du, p):
    du[0] = u[0] - u[0] * u[1]
    du[1] = u[0] * u[1] - u[1]


@cfunc(SIG)
def lv_sums(t, u, du, p):  # solve_ivp.ipynb cell 21 with np.sum(p.arr1), np.sum(p.arr2) in p[0], p[1]
    du[0] = u[0] - u[0] * u[1] * p[0]
    du[1] = u[0] * u[1] - u[1] * p[1]


@cfunc(SIG)
def ev_u0_is_1(t, u, out, p):
    out[0] = u[0] - 1


U0 = np.array([5.0, 0.8])
TE = np.linspace(0.0, 50.0, 1000)
SPAN = np.array([0.0, 50.0])

NOTEBOOK_CELL_27 = [0.7463063968797129, 0.7679450731640909, 0.7909998733859201, 0.8157222725496384,
                    0.8424776519856828, 0.8718277396448789, 0.9047034490300899, 0.9428490823677572,
                    0.9901493996386634, 1.0587894770207835]


def test_notebook_simple_event():
    r = orc.solve_ivp(lv.address, SPAN, U0, TE, 1, ev_u0_is_1.address, rtol=1e-6)
    assert r["message"] == "A termination event occurred." and r["success"]
    assert r["event_found"] and r["ind_event"] == 0
    assert abs(r["t_event"] - 0.7611343988149695) < 1e-13
    assert np.allclose(r["y_event"], [1.0, 5.02890877], rtol=0, atol=5e-9)
    # rows: every t_eval point up to the event
    assert len(r["t"]) == np.searchsorted(TE, r["t_event"], side="right")
    assert r["y"].shape == (len(r["t"]), 2) and np.array_equal(r["y"][0], U0)


def test_notebook_struct_data_event():
    S = np.array([np.sum(0.1 * np.ones(6)), np.sum(0.05 * np.ones(3))])
    r = orc.solve_ivp(lv_sums.address, SPAN, U0, TE, 1, ev_u0_is_1.address, data=S, rtol=1e-6)
    assert abs(r["t_event"] - 0.7258968839312535) < 1e-13


def test_notebook_event_sweep():
    got = []
    for i in range(10):
        a1 = 0.1 * np.ones(6)
        a2 = 0.05 * np.ones(3)
        a1[:] = 0.1 - i * 0.01
        a2[:] = 0.1 + i * 0.01
        S = np.array([np.sum(a1), np.sum(a2)])
        got.append(orc.solve_ivp(lv_sums.address, SPAN, U0, TE, 1, ev_u0_is_1.address, data=S,
                                 rtol=1e-6)["t_event"])
    assert np.abs(np.array(got) - NOTEBOOK_CELL_27).max() < 1e-13, got


def test_reference_known_answers():
    te = np.linspace(0.0, 10.0, 11)
    span = np.array([0.0, 10.0])
    for kw in (dict(t_eval=te), dict()):
        r = orc.solve_ivp(lv.address, span, U0, rtol=1e-8, atol=1e-8, **kw)
        assert r["success"] and r["message"].startswith("The solver successfully")
        assert np.isclose(r["y"][-1, 0], 0.38583246250193476)
        assert np.isclose(r["y"][-1, 1], 4.602012234037773)
    # without t_eval every accepted step is a row, the first one the initial state
    assert r["t"][0] == 0.0 and r["t"][-1] == 10.0 and len(r["t"]) == r["naccpt"] + 1
    assert r["nfev"] == 2 + 11 * r["nstep"] + 4 * r["naccpt"]


@pytest.mark.parametrize("sign", [1.0, -1.0])
def test_forward_and_backward_against_scipy(sign):
    """solve_ivp.ipynb cell 6."""
    from scipy import integrate
    te = np.linspace(0.0, sign * 50.0, 1000)
    span = np.array([0.0, sign * 50.0])
    r = orc.solve_ivp(lv.address, span, U0, te, rtol=1e-9, atol=1e-30)
    sp = integrate.solve_ivp(lambda t, u: [u[0] - u[0] * u[1], u[0] * u[1] - u[1]], span, U0,
                             t_eval=te, rtol=1e-9, atol=1e-30, method="DOP853")
    assert r["success"] and np.all(np.isclose(sp.y.T, r["y"]))


def test_dop853_rows_agree_with_the_pinned_dop853_oracle():
    """Same step sequence as oracle_dop853 (iout = 3) when there are no events: the dense
    rows are then the same polynomials evaluated at the same points."""
    p = pr.problem("lv_test")
    f = orc.builtin_rhs("lotka_volterra")
    ref, ok, st = orc.dop853(f, p["u0"], p["t_eval"], p["data"], 1e-8, 1e-8)
    r = orc.solve_ivp(f, [p["t_eval"][0], p["t_eval"][-1]], p["u0"], p["t_eval"], data=p["data"],
                      rtol=1e-8, atol=1e-8)
    assert ok and r["success"] and (r["nstep"], r["naccpt"]) == (st.nstep, st.naccpt)
    assert np.array_equal(r["y"], ref)


def test_edge_cases():
    # t_eval not monotone (dop853_solve_ivp.f90:278-303)
    r = orc.solve_ivp(lv.address, SPAN, U0, np.array([0.0, 2.0, 1.0]))
    assert not r["success"] and "does not always increase" in r["message"] and len(r["t"]) == 0
    r = orc.solve_ivp(lv.address, [0.0, -5.0], U0, np.array([0.0, -2.0, -1.0]))
    assert not r["success"]
    # first_step and max_step
    a = orc.solve_ivp(lv.address, [0.0, 5.0], U0, rtol=1e-6, max_step=0.01)
    assert a["success"] and np.diff(a["t"]).max() <= 0.01 * (1 + 1e-9) and len(a["t"]) >= 500
    b = orc.solve_ivp(lv.address, [0.0, 5.0], U0, rtol=1e-6, first_step=1e-4)
    assert b["success"] and b["t"][1] == 1e-4
    # no sign change: the end of the interval is reached
    r = orc.solve_ivp(lv.address, [0.0, 0.5], U0, None, 1, ev_u0_is_1.address, rtol=1e-6)
    assert r["success"] and not r["event_found"] and r["t"][-1] == 0.5
    # capacity (ours): more accepted steps than rows
    r = orc.solve_ivp(lv.address, SPAN, U0, rtol=1e-8, atol=1e-8, cap=10)
    assert not r["success"] and r["message"] == "output capacity exceeded" and r["rows_needed"] > 10


def _sweep_event_times():
    got, rej = [], []
    for i in range(10):
        a1 = 0.1 * np.ones(6)
        a2 = 0.05 * np.ones(3)
        a1[:] = 0.1 - i * 0.01
        a2[:] = 0.1 + i * 0.01
        S = np.array([np.sum(a1), np.sum(a2)])
        r = orc.solve_ivp(lv_sums.address, SPAN, U0, TE, 1, ev_u0_is_1.address, data=S, rtol=1e-6)
        got.append(r["t_event"])
        rej.append(r["nstep"] - r["naccpt"])
    return np.array(got), np.array(rej)


def test_notebook_outputs_pin_the_step_controller_and_its_rejection_rule():
    """Every one of the notebook's runs rejects steps (1-4 of its 6-10 attempts), so the event
    times the reference printed depend on dp86co's controller *including* the one line scipy's
    dop853 cannot cover: hnew = h/min(facc1, fac11/safe) after a rejection (module :725).  With
    Hairer's original rule (h/facc1) the oracle misses the printed values by five orders of
    magnitude more than with the reference's."""
    got, rej = _sweep_event_times()
    assert (rej >= 1).all()
    assert np.abs(got - NOTEBOOK_CELL_27).max() < 1e-13
    orc.lib.oracle_ivp_set_scipy_reject_rule(1)
    try:
        alt, _ = _sweep_event_times()
    finally:
        orc.lib.oracle_ivp_set_scipy_reject_rule(0)
    assert np.abs(alt - NOTEBOOK_CELL_27).max() > 1e-9
    assert (np.abs(alt - NOTEBOOK_CELL_27) > 1e-11).sum() >= 8


def test_the_dop853_oracle_inherits_the_pin():
    """dop853_oracle.c and solve_ivp_oracle.c restate dp86co independently (iout = 3 / iout = 2).
    On the notebook's problem -- rejected steps included -- they take the same steps and produce
    the same rows bit for bit, so what pins the one pins the other."""
    for data, rtol in ((np.array([0.6, 0.3]), 1e-6), (np.array([0.1, 1.0]), 1e-6), (np.array([1.0, 1.0]), 1e-9)):
        ref, ok, st = orc.dop853(lv_sums.address, U0, TE, data, rtol, 1e-6)
        r = orc.solve_ivp(lv_sums.address, SPAN, U0, TE, data=data, rtol=rtol, atol=1e-6)
        assert ok and r["success"] and st.nstep - st.naccpt >= 1
        assert (r["nstep"], r["naccpt"], r["nrejct"]) == (st.nstep, st.naccpt, st.nrejct)
        assert np.array_equal(r["y"], ref)


def test_golden_vectors():
    """The committed fixtures (tests/golden/solve_ivp_oracle.npz) are what the oracle produces --
    and carry the event times the reference printed."""
    import _ivp_golden as gold
    assert len(gold.CASES) == 15
    for name, (rhs, span, y0, te, ne, ev, data, kw, printed) in gold.CASES.items():
        r = orc.solve_ivp(orc.builtin_rhs(rhs), span, y0, te, ne,
                          orc.builtin_rhs("events_" + ev) if ev else 0, data, cap=4000, **kw)
        f = dict(success=r["success"], message=orc.IVP_MESSAGES.index(r["message"]), nfev=r["nfev"],
                 nt_out=len(r["t"]), rows_needed=r["rows_needed"], event_found=r["event_found"],
                 ind_event=r["ind_event"], idid=r["idid"], nstep=r["nstep"], naccpt=r["naccpt"],
                 nrejct=r["nrejct"])
        gold.check(name, f, r["t"], r["y"], r["t_event"], r["y_event"])
