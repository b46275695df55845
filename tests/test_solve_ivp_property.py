"""Property tests (hypothesis) of the solve_ivp restatements on random problems: the oracle, the
host-compiled thread core and the host-compiled warp core agree bit for bit on every ODEResult
field, and the results have the properties the domain offers -- rows are the requested output
times (or strictly monotone solver steps ending at the end of the span or at the event), an event
state lies on its level to the Brent tolerance, and integrating back from the end point returns
to the start."""

import numpy as np
import pytest

import _oracle as orc
from test_solve_ivp_core_host import _Core, _compile, run

nb = pytest.importorskip("numba")
hyp = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402
from numba import cfunc, types  # noqa: E402

SIG = types.void(types.double, types.CPointer(types.double), types.CPointer(types.double),
                 types.CPointer(types.double))


@cfunc(SIG)
def osc(t, u, du, p):
    # n coupled nonlinear oscillators (n even), p = [n, w, eps, level]
    n = int(p[0])
    for i in range(0, n, 2):
        j = (i + 2) % n
        du[i] = u[i + 1]
        du[i + 1] = -p[1] * p[1] * u[i] - p[2] * u[i] * u[i] * u[i] + 0.1 * (u[j] - u[i])


@cfunc(SIG)
def ev_level(t, u, out, p):
    out[0] = u[0] - p[3]


@pytest.fixture(scope="module")
def cores():
    return (_Core(_compile("solve_ivp_core_host").solve_ivp_core_host),
            _Core(_compile("solve_ivp_warp_host").solve_ivp_warp_host))


FIELDS = ("message", "nfev", "success", "event_found", "ind_event", "idid", "nstep", "naccpt", "nrejct",
          "rows_needed", "t_event")


@settings(max_examples=40, deadline=None)
@given(n=st.sampled_from([2, 4, 6, 8]), w=st.floats(0.5, 3.0), eps=st.floats(0.0, 1.0),
       level=st.floats(-0.9, 0.9), tf=st.floats(0.5, 8.0), back=st.booleans(), with_te=st.booleans(),
       rexp=st.integers(4, 9), seed=st.integers(0, 10 ** 6))
def test_three_restatements_agree_and_results_are_consistent(cores, n, w, eps, level, tf, back, with_te,
                                                             rexp, seed):
    rng = np.random.default_rng(seed)
    y0 = rng.uniform(-1.0, 1.0, n)
    y0[0] = 1.0  # starts above every level, so a crossing is a real sign change
    data = np.array([float(n), w, eps, level])
    span = np.array([0.0, -tf if back else tf])
    te = np.linspace(span[0], span[1], 17) if with_te else None
    rtol = 10.0 ** -rexp
    kw = dict(rtol=rtol, atol=1e-3 * rtol)
    ref = orc.solve_ivp(osc.address, span, y0, te, 1, ev_level.address, data, cap=20000, **kw)
    for core in cores:
        got = run(core, osc.address, span, y0, te, 1, ev_level.address, data, cap=20000, **kw)[0]
        for k in FIELDS:
            assert got[k] == ref[k], k
        # (equal_nan: when the last step's x + h rounds one ulp below t_eval[-1] the reference's
        # solout never writes the final row -- SURVEY.md appendix A, DOP853 item 4 -- and all
        # three restatements leave it unwritten alike)
        assert np.array_equal(got["t"], ref["t"]) and np.array_equal(got["y"], ref["y"], equal_nan=True)
        assert np.array_equal(got["y_event"], ref["y_event"])
    assert ref["success"]
    t, y = ref["t"], ref["y"]
    sgn = -1.0 if back else 1.0
    if ref["event_found"]:
        # the event state is on the level (Brent tol 1e-8 on t; |du0/dt| is O(1))
        assert abs(ref["y_event"][0] - level) < 1e-6 and 0.0 < sgn * ref["t_event"] <= tf
        assert np.all(sgn * t <= sgn * ref["t_event"])
    if with_te:
        assert np.array_equal(t, te[:len(t)]) and (ref["event_found"] or len(t) == len(te))
    else:
        assert t[0] == 0.0 and np.all(np.diff(sgn * t) > 0)
        # the last step is h = xend - x, and x + h may round one ulp off xend (module :554-557)
        assert ref["event_found"] or abs(t[-1] - span[1]) <= 2.0 * np.spacing(abs(span[1]))
        assert len(t) == (ref["naccpt"] if ref["event_found"] else ref["naccpt"] + 1)
    if not ref["event_found"] and not with_te and rexp >= 7:
        # round trip: integrate back from the end point
        r2 = orc.solve_ivp(osc.address, span[::-1].copy(), y[-1], None, 0, 0, data, cap=20000, **kw)
        assert r2["success"] and np.abs(r2["y"][-1] - y0).max() < 2e4 * rtol * max(1.0, tf)
