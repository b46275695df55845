"""Host logic of numbalsoda_b200.solve_ivp (argument normalisation, batching by array rank, the
two-pass ragged output, ODEResult / ODEResultBatch) without a GPU: the one call into the C ABI,
b200ode_solve_ivp, is replaced by a test shim that runs the host-compiled solver core
(tests/solve_ivp_core_host.cpp) on the same b200ode_ivp_problem structure.  The product has no such
path (tests/test_abi.py: solving without a device fails loudly); GPU parity is
tests/test_gpu_solve_ivp.py."""
import ctypes as ct

import numpy as np
import pytest

import _oracle as orc
from test_solve_ivp_core_host import IvpLaunch, lib  # noqa: F401  (fixture)

nb = pytest.importorskip("numba")

U0 = np.array([5.0, 0.8])
ONE = np.array([1.0])
TE = np.linspace(0.0, 50.0, 1000)
SPAN = np.array([0.0, 50.0])


@pytest.fixture()
def drv(lib, monkeypatch):  # noqa: F811
    import numbalsoda_b200 as b2
    from numbalsoda_b200 import _lib, rhs as _rhs, driver_solve_ivp as drvmod

    class FakeHandle:
        def __init__(self, f, ev, n, ne):
            self.f, self.ev, self.n, self.ne, self.ptr = f, ev, n, ne, self

    def get_ivp_handle(rhs, neq, n_events=0, events=None, strict=False, exact_ctrl=False):
        if n_events > 0 and (events is None or (isinstance(events, int) and events == 0)):
            raise ValueError("n_events = %d but no event function was given" % n_events)
        return FakeHandle(rhs, events if n_events else 0, neq, n_events)

    real = _lib.lib

    class FakeLib:
        def __getattr__(self, k):
            return getattr(real, k)

        @staticmethod
        def b200ode_solve_ivp(h, pbref, dev):
            pb = pbref._obj
            assert pb.size == ct.sizeof(_lib.IvpProblem) and pb.neq == h.n and pb.n_events == h.ne
            L = IvpLaunch(B=pb.B, t_span=pb.t_span, t_span_stride=pb.t_span_stride, y0=pb.y0,
                          y0_stride=pb.y0_stride, data=pb.data,
                          data_stride=max(pb.data_stride, 0), np=-1, nt=pb.nt, t_eval=pb.teval,
                          n_events=pb.n_events, cap=pb.nt if pb.nt else pb.cap,
                          row_offset=None if pb.nt else pb.row_offset, first_step=pb.first_step,
                          max_step=pb.max_step, rtol=pb.rtol, atol=pb.atol, rtol_b=pb.rtol_b,
                          atol_b=pb.atol_b, t_out=pb.t_out, y_out=pb.y_out, result=pb.result,
                          t_event=pb.t_event, y_event=pb.y_event, next=None)
            return lib.solve_ivp_core_host(ct.c_void_p(h.f), ct.c_void_p(h.ev), pb.neq, ct.byref(L))

    monkeypatch.setattr(_rhs, "get_ivp_handle", get_ivp_handle)
    monkeypatch.setattr(drvmod._lib, "lib", FakeLib())
    yield b2
    monkeypatch.undo()


def same(r, ref, reached=None):
    """reached: rows the integration got to (t_eval points beyond t_span stay uninitialised
    memory, in the reference as here)"""
    if reached is not None:
        r.y, ref = r.y[:reached], dict(ref, y=ref["y"][:reached])
        r.t, ref["t"] = r.t[:reached], ref["t"][:reached]
    assert r.message == ref["message"]
    assert (r.nfev, r.success, r.event_found, r.ind_event) == (ref["nfev"], ref["success"],
                                                               ref["event_found"], ref["ind_event"])
    assert r.t_event == ref["t_event"] and np.array_equal(r.y_event, ref["y_event"])
    assert np.array_equal(r.t, ref["t"]) and np.array_equal(r.y, ref["y"], equal_nan=True)


def test_single_results_have_the_reference_fields(drv):
    f, ev = orc.builtin_rhs("lotka_volterra"), orc.builtin_rhs("events_lv_prey")
    r = drv.solve_ivp(f, SPAN, U0, TE, n_events=1, event_fcn=ev, data=ONE, rtol=1e-6)
    same(r, orc.solve_ivp(f, SPAN, U0, TE, 1, ev, ONE, rtol=1e-6))
    assert abs(r.t_event - 0.7611343988149695) < 1e-13  # solve_ivp.ipynb cell 11
    for kw in (dict(), dict(max_rows=5000), dict(first_step=1e-3, max_step=0.5)):
        r = drv.solve_ivp(f, SPAN, U0, data=ONE, rtol=1e-8, atol=1e-8, **kw)
        okw = {k: v for k, v in kw.items() if k != "max_rows"}
        same(r, orc.solve_ivp(f, SPAN, U0, data=ONE, rtol=1e-8, atol=1e-8, **okw))
    r = drv.solve_ivp(f, SPAN, U0, data=ONE, rtol=1e-8, atol=1e-8, max_rows=10)
    assert not r.success and r.message == "output capacity exceeded" and len(r.t) == 10
    assert r.counters["rows_needed"] > 10
    r = drv.solve_ivp(f, SPAN, U0, np.array([0.0, 2.0, 1.0]), data=ONE)
    assert not r.success and "does not always increase" in r.message and r.y.shape == (0, 2)


@pytest.mark.parametrize("mode", ["t_eval", "two_pass", "max_rows"])
def test_ensembles(drv, mode):
    rng = np.random.default_rng(3)
    B = 40
    data = np.column_stack([rng.uniform(9, 11, B), rng.uniform(20, 36, B), rng.uniform(2, 3.33, B),
                            rng.uniform(30, 45, B)])
    y0 = np.column_stack([rng.uniform(0.5, 1.5, B), rng.uniform(-0.5, 0.5, B), rng.uniform(0, 1, B)])
    ts = np.column_stack([np.zeros(B), rng.uniform(0.3, 6.0, B)])
    f, ev = orc.builtin_rhs("lorenz"), orc.builtin_rhs("events_lorenz_z")
    te = np.linspace(0.0, 6.0, 61) if mode == "t_eval" else None
    cap = 60 if mode == "max_rows" else 4000
    got = drv.solve_ivp(f, ts, y0, te, n_events=1, event_fcn=ev, data=data, rtol=1e-7, atol=1e-9,
                        max_rows=cap if mode == "max_rows" else None)
    assert len(got) == B and 0 < got.event_found.sum() < B
    for b, r in enumerate(got):
        same(r, orc.solve_ivp(f, ts[b], y0[b], te, 1, ev, data[b], rtol=1e-7, atol=1e-9, cap=cap),
             reached=None if te is None else np.searchsorted(te, ts[b, 1], side="right"))
    if mode == "two_pass":
        assert got.row_offset[-1] == got.rows_needed.sum() == len(got.t)
    if mode == "max_rows":
        assert (~got.success).any() and got.success.any()
    # shared y0 / t_span with batched data, and per-trajectory tolerances
    rt = rng.choice([1e-5, 1e-7], B)
    got = drv.solve_ivp(f, np.array([0.0, 2.0]), y0[0], te, data=data, rtol=rt, atol=1e-9)
    for b in (0, 7, B - 1):
        same(got[b], orc.solve_ivp(f, [0.0, 2.0], y0[0], te, data=data[b], rtol=rt[b], atol=1e-9),
             reached=None if te is None else np.searchsorted(te, 2.0, side="right"))


def test_argument_errors(drv):
    f = orc.builtin_rhs("lotka_volterra")
    with pytest.raises(ValueError, match="length 2"):
        drv.solve_ivp(f, np.array([0.0, 1.0, 2.0]), U0, data=ONE)
    with pytest.raises(ValueError, match="method"):
        drv.solve_ivp(f, SPAN, U0, method="RK45", data=ONE)
    with pytest.raises(ValueError, match="no event function"):
        drv.solve_ivp(f, SPAN, U0, n_events=1, data=ONE)
    with pytest.raises(TypeError):
        drv.solve_ivp(f, SPAN, U0, data=12345)
    with pytest.raises(ValueError, match="disagree"):
        drv.solve_ivp(f, np.zeros((3, 2)), np.zeros((4, 2)), data=ONE)
    with pytest.raises(TypeError, match="float64"):
        drv.solve_ivp(f, SPAN, np.array([5, 1]), data=ONE)


@pytest.mark.parametrize("mode", ["t_eval", "two_pass", "max_rows"])
def test_sharding_over_devices(drv, mode):
    """device=[...] shards the ensemble by trajectory, one host thread per device, no exchange
    (SURVEY.md 8e); the ragged rows of a shard are addressed through its slice of row_offset."""
    rng = np.random.default_rng(4)
    B = 23
    data = np.column_stack([rng.uniform(9, 11, B), rng.uniform(20, 36, B), rng.uniform(2, 3.33, B),
                            rng.uniform(30, 45, B)])
    ts = np.column_stack([np.zeros(B), rng.uniform(0.3, 3.0, B)])
    y0 = np.array([1.0, 0.0, 0.0])
    f, ev = orc.builtin_rhs("lorenz"), orc.builtin_rhs("events_lorenz_z")
    kw = dict(n_events=1, event_fcn=ev, data=data, rtol=1e-7, atol=1e-9)
    if mode == "t_eval":
        kw["t_eval"] = np.linspace(0.0, 3.0, 16)
    if mode == "max_rows":
        kw["max_rows"] = 300
    one = drv.solve_ivp(f, ts, y0, **kw)
    three = drv.solve_ivp(f, ts, y0, device=[0, 0, 0], **kw)
    assert np.array_equal(one._r, three._r) and np.array_equal(one.t_event, three.t_event)
    assert np.array_equal(one.y_event, three.y_event)
    for k, (a, b) in enumerate(zip(one, three)):
        # (t_eval rows beyond a trajectory's span are uninitialised memory)
        m = np.searchsorted(kw["t_eval"], ts[k, 1], side="right") if mode == "t_eval" else len(a.t)
        assert np.array_equal(a.t, b.t) and np.array_equal(a.y[:m], b.y[:m])
