"""Parity of the CUDA solve_ivp path (numbalsoda_b200.solve_ivp -> b200ode_solve_ivp, the C ABI)
with the CPU oracle (oracle/solve_ivp_oracle.c, pinned by the event times the reference printed
in solve_ivp.ipynb) on the same inputs.  strict=True links the kernels with FMA contraction
disabled: every field of the reference's ODEResult is then compared bit for bit."""
import ctypes as ct

import numpy as np
import pytest

import _oracle as orc
import _problems as pr

pytestmark = pytest.mark.gpu

U0 = np.array([5.0, 0.8])
ONE = np.array([1.0])
TE = np.linspace(0.0, 50.0, 1000)
SPAN = np.array([0.0, 50.0])
LV = "lotka_volterra"


def _nb():
    import numbalsoda_b200 as nb
    return nb


def same(r, ref):
    """r: numbalsoda_b200.ODEResult, ref: dict from _oracle.solve_ivp"""
    assert r.message == ref["message"]
    assert (r.nfev, r.success, r.event_found, r.ind_event) == (ref["nfev"], ref["success"],
                                                               ref["event_found"], ref["ind_event"])
    for k in ("idid", "nstep", "naccpt", "nrejct", "rows_needed"):
        assert r.counters[k] == ref[k], k
    assert r.t_event == ref["t_event"] and np.array_equal(r.y_event, ref["y_event"])
    assert np.array_equal(r.t, ref["t"]) and np.array_equal(r.y, ref["y"], equal_nan=True)


def ev_two(t, u, out, p):  # the first event never fires, the second is the notebook's
    out[0] = u[1] - 30.0
    out[1] = u[0] - 1.0


def test_notebook_event_strict_bitwise():
    """solve_ivp.ipynb cells 9-11: t_event = 0.7611343988149695, y_event = [1., 5.02890877]."""
    nb = _nb()
    r = nb.solve_ivp(nb.builtin(LV), SPAN, U0, TE, n_events=1, event_fcn=nb.builtin("lv_prey"),
                     data=ONE, rtol=1e-6, strict=True)
    ref = orc.solve_ivp(orc.builtin_rhs(LV), SPAN, U0, TE, 1, orc.builtin_rhs("events_lv_prey"), ONE,
                        rtol=1e-6)
    same(r, ref)
    assert r.message == "A termination event occurred." and r.success and r.ind_event == 0
    assert abs(r.t_event - 0.7611343988149695) < 1e-13
    assert np.allclose(r.y_event, [1.0, 5.02890877], rtol=0, atol=5e-9)
    assert len(r.t) == np.searchsorted(TE, r.t_event, side="right") and np.array_equal(r.y[0], U0)


def test_reference_usage_with_numba_cfuncs():
    """The reference's calling convention: cfunc addresses for the right-hand side and the
    events function (tests/test_numbalsoda.py:33-56, solve_ivp.ipynb cell 9)."""
    import numba
    nb = _nb()
    rhs = numba.cfunc(nb.lsoda_sig)(pr.lv_rhs)
    ev = numba.cfunc(nb.lsoda_sig)(ev_two)
    te = np.linspace(0.0, 10.0, 11)
    for kw in (dict(t_eval=te), dict()):
        sol = nb.solve_ivp(rhs.address, np.array([0.0, 10.0]), U0, data=ONE, rtol=1e-8, atol=1e-8, **kw)
        assert sol.success and sol.message.startswith("The solver successfully")
        assert np.isclose(sol.y[-1, 0], 0.38583246250193476)
        assert np.isclose(sol.y[-1, 1], 4.602012234037773)
    assert sol.t[0] == 0.0 and sol.t[-1] == 10.0 and len(sol.t) == sol.counters["naccpt"] + 1
    r = nb.solve_ivp(rhs.address, SPAN, U0, TE, n_events=2, event_fcn=ev.address, data=ONE, rtol=1e-6,
                     strict=True)
    ref = orc.solve_ivp(rhs.address, SPAN, U0, TE, 2, ev.address, ONE, rtol=1e-6)
    same(r, ref)
    assert r.event_found and r.ind_event == 1


@pytest.mark.parametrize("kw", [
    dict(t_eval=TE, rtol=1e-8, atol=1e-8),
    dict(t_eval=None, rtol=1e-8, atol=1e-8),
    dict(t_eval=None, n_events=1, rtol=1e-6),
    dict(t_eval=None, rtol=1e-6, max_step=0.01, span=[0.0, 5.0]),
    dict(t_eval=None, rtol=1e-6, first_step=1e-4, span=[0.0, 5.0]),
    dict(t_eval=-TE, span=[0.0, -50.0], rtol=1e-9, atol=1e-30),
    dict(t_eval=None, span=[0.0, -3.0], n_events=1),
    dict(t_eval=np.array([0.0, 2.0, 1.0])),
    dict(t_eval=-TE),
    dict(t_eval=TE[:1]),
    dict(t_eval=np.array([0.0, 0.0, 1.0, 1.0, 50.0, 50.0])),
    dict(t_eval=np.linspace(0.0, 80.0, 9)),  # t_eval beyond t_span: rows the solver never reaches
])
def test_single_trajectory_cases_strict_bitwise(kw):
    nb = _nb()
    kw = dict(kw)
    span = np.array(kw.pop("span", SPAN), dtype=np.float64)
    ne = kw.pop("n_events", 0)
    te = kw.pop("t_eval")
    r = nb.solve_ivp(nb.builtin(LV), span, U0, te, n_events=ne,
                     event_fcn=nb.builtin("lv_prey") if ne else 0, data=ONE, strict=True, **kw)
    ref = orc.solve_ivp(orc.builtin_rhs(LV), span, U0, te, ne,
                        orc.builtin_rhs("events_lv_prey") if ne else 0, ONE, **kw)
    if te is not None and ref["success"] and te[-1] > max(span):
        # unreached rows are uninitialised memory in the reference; compare the written ones
        k = np.searchsorted(te, max(span), side="right")
        ref["y"], r.y = ref["y"][:k], r.y[:k]
    same(r, ref)


def _lorenz_ensemble(B, seed=3):
    rng = np.random.default_rng(seed)
    data = np.column_stack([rng.uniform(9, 11, B), rng.uniform(20, 36, B), rng.uniform(2, 3.33, B),
                            rng.uniform(30, 45, B)])
    y0 = np.column_stack([rng.uniform(0.5, 1.5, B), rng.uniform(-0.5, 0.5, B), rng.uniform(0, 1, B)])
    ts = np.column_stack([np.zeros(B), rng.uniform(0.3, 6.0, B)])
    return ts, y0, data


@pytest.mark.parametrize("mode", ["t_eval", "two_pass", "max_rows"])
def test_lorenz_event_sweep_strict_bitwise(mode):
    """Per-trajectory t_span, y0 and data; trajectories that do and do not hit z = level."""
    nb = _nb()
    B = 1500
    ts, y0, data = _lorenz_ensemble(B)
    te = np.linspace(0.0, 6.0, 61) if mode == "t_eval" else None
    cap = 40 if mode == "max_rows" else 4000
    got = nb.solve_ivp(nb.builtin("lorenz"), ts, y0, te, n_events=1, event_fcn=nb.builtin("lorenz_z"),
                       data=data, rtol=1e-7, atol=1e-9, strict=True,
                       max_rows=cap if mode == "max_rows" else None)
    ref = orc.solve_ivp_batch(orc.builtin_rhs("lorenz"), ts, y0, te, 1,
                              orc.builtin_rhs("events_lorenz_z"), data, rtol=1e-7, atol=1e-9, cap=cap)
    rr = ref["result"]
    assert len(got) == B and 0.05 < rr["event_found"].mean() < 0.95
    assert np.array_equal(got.success, rr["success"] == 1)
    for k, f in (("message_code", "message"), ("nfev", "nfev"), ("nt_out", "nt_out"),
                 ("ind_event", "ind_event"), ("idid", "idid"), ("nstep", "nstep"), ("naccpt", "naccpt"),
                 ("nrejct", "nrejct"), ("rows_needed", "rows_needed")):
        assert np.array_equal(getattr(got, k), rr[f]), k
    assert np.array_equal(got.event_found, rr["event_found"] == 1)
    assert np.array_equal(got.t_event, rr["t_event"]) and np.array_equal(got.y_event, ref["y_event"])
    if mode == "max_rows":
        assert (~got.success).any() and got.success.any()  # some trajectories need more rows
        assert got[int(np.argmin(got.success))].message == "output capacity exceeded"
    for b in range(B):
        k = rr["nt_out"][b]
        r = got[b]
        assert len(r.t) == k
        if te is not None:  # rows beyond t_span are never written (uninitialised in the reference)
            k = min(k, np.searchsorted(te, ts[b, 1], side="right"))
        assert np.array_equal(r.y[:k], ref["y"][b, :k]), b
        if te is None:
            assert np.array_equal(r.t, ref["t"][b, :k]), b
    if mode == "two_pass":
        assert got.row_offset[-1] == rr["rows_needed"].sum() == got.y.shape[0]


def test_fast_mode_within_tolerance():
    """Default link mode (FMA contraction on): same event decisions, event times to ~1e-8 (the
    Brent tolerance), states within 10 rtol (north_star's bound)."""
    nb = _nb()
    B = 500
    ts, y0, data = _lorenz_ensemble(B, seed=11)
    ts[:, 1] = np.minimum(ts[:, 1], 2.0)  # short horizon: Lorenz is chaotic
    te = np.linspace(0.0, 2.0, 21)
    rtol = 1e-8
    got = nb.solve_ivp(nb.builtin("lorenz"), ts, y0, te, n_events=1, event_fcn=nb.builtin("lorenz_z"),
                       data=data, rtol=rtol, atol=rtol)
    ref = orc.solve_ivp_batch(orc.builtin_rhs("lorenz"), ts, y0, te, 1,
                              orc.builtin_rhs("events_lorenz_z"), data, rtol=rtol, atol=rtol)
    rr = ref["result"]
    assert got.success.all() and np.array_equal(got.event_found, rr["event_found"] == 1)
    assert np.array_equal(got.nt_out, rr["nt_out"])
    ev = got.event_found
    assert np.abs(got.t_event[ev] - rr["t_event"][ev]).max() < 5e-8
    for b in range(B):
        k = min(rr["nt_out"][b], np.searchsorted(te, ts[b, 1], side="right"))
        a, r = got.y[b, :k], ref["y"][b, :k]
        assert (np.abs(a - r) / (np.abs(r) + 1.0)).max() <= 10 * rtol


def test_failures_are_reported_like_the_reference():
    nb = _nb()

    def stiff(t, u, du, p):
        du[0] = -p[0] * (u[0] - np.cos(t))
    import numba
    cf = numba.cfunc(nb.lsoda_sig)(stiff)
    lam = np.array([1e4])
    r = nb.solve_ivp(cf.address, np.array([0.0, 100.0]), np.array([0.0]), data=lam, strict=True,
                     max_rows=8)
    ref = orc.solve_ivp(cf.address, [0.0, 100.0], [0.0], data=lam, cap=8)
    # idid = -4: the stiffness detector stops the integration (iprint = 0, module :642-650)
    assert ref["idid"] == -4 and not r.success and r.nfev == 0 and r.counters["idid"] == -4
    assert (r.counters["nstep"], r.counters["naccpt"]) == (ref["nstep"], ref["naccpt"])
    # a mixed ensemble: one increasing and one decreasing span against an increasing t_eval
    ts = np.array([[0.0, 5.0], [0.0, -5.0]])
    got = nb.solve_ivp(nb.builtin(LV), ts, U0, np.linspace(0, 5, 6), data=ONE)
    assert got.success.tolist() == [True, False]
    assert got[1].message == '"t_eval" does not always increase or decrease' and len(got[1].t) == 0


def test_argument_errors():
    nb = _nb()
    with pytest.raises(ValueError, match="length 2"):
        nb.solve_ivp(nb.builtin(LV), np.array([0.0, 1.0, 2.0]), U0, data=ONE)
    with pytest.raises(ValueError, match="method"):
        nb.solve_ivp(nb.builtin(LV), SPAN, U0, method="RK45", data=ONE)
    with pytest.raises(ValueError, match="no event function"):
        nb.solve_ivp(nb.builtin(LV), SPAN, U0, n_events=1, data=ONE)
    with pytest.raises(TypeError):
        nb.solve_ivp(nb.builtin(LV), SPAN, U0, data=12345)


def test_chunked_host_path_and_device_api(monkeypatch):
    """Many chunks through the host-buffer path (B200ODE_IVP_CHUNK forces small ones), ragged
    and t_eval layouts, against one big launch on device-resident buffers through
    b200ode_solve_ivp_device (torch tensors: plumbing only)."""
    import torch
    from numbalsoda_b200 import _lib, rhs as _rhs
    nb = _nb()
    B = 5000
    ts, y0, data = _lorenz_ensemble(B, seed=5)
    ts[:, 1] = np.minimum(ts[:, 1], 1.5)
    kw = dict(n_events=1, event_fcn=nb.builtin("lorenz_z"), data=data, rtol=1e-6, atol=1e-8, strict=True)
    whole = nb.solve_ivp(nb.builtin("lorenz"), ts, y0, **kw)
    monkeypatch.setenv("B200ODE_IVP_CHUNK", "333")
    parts = nb.solve_ivp(nb.builtin("lorenz"), ts, y0, **kw)
    assert np.array_equal(whole._r, parts._r) and np.array_equal(whole.row_offset, parts.row_offset)
    assert np.array_equal(whole.t, parts.t) and np.array_equal(whole.y, parts.y)
    assert np.array_equal(whole.t_event, parts.t_event) and np.array_equal(whole.y_event, parts.y_event)
    te = np.linspace(0.0, 1.5, 16)
    whole_te = nb.solve_ivp(nb.builtin("lorenz"), ts, y0, te, **kw)
    monkeypatch.delenv("B200ODE_IVP_CHUNK")

    # device-resident call, ragged rows
    h = _rhs.get_ivp_handle(nb.builtin("lorenz"), 3, 1, nb.builtin("lorenz_z"), strict=True)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)  # noqa: E731
    d_ts, d_y0, d_data, d_off = t(ts), t(y0), t(data), t(whole.row_offset)
    R = int(whole.row_offset[-1])
    d_t = torch.empty(R, dtype=torch.float64, device=dev)
    d_y = torch.empty((R, 3), dtype=torch.float64, device=dev)
    d_res = torch.zeros((B, _lib.IVP_NRESULT), dtype=torch.int32, device=dev)
    d_tev = torch.empty(B, dtype=torch.float64, device=dev)
    d_yev = torch.empty((B, 3), dtype=torch.float64, device=dev)
    pb = _lib.IvpProblem()
    pb.size = ct.sizeof(_lib.IvpProblem)
    pb.B, pb.neq, pb.nt, pb.np, pb.n_events = B, 3, 0, 4, 1
    pb.t_span, pb.t_span_stride, pb.y0, pb.y0_stride = d_ts.data_ptr(), 2, d_y0.data_ptr(), 3
    pb.data, pb.data_stride = d_data.data_ptr(), 32
    pb.rtol, pb.atol = 1e-6, 1e-8
    pb.row_offset, pb.t_out, pb.y_out = d_off.data_ptr(), d_t.data_ptr(), d_y.data_ptr()
    pb.result, pb.t_event, pb.y_event = d_res.data_ptr(), d_tev.data_ptr(), d_yev.data_ptr()
    stream = torch.cuda.current_stream().cuda_stream
    _lib.check(_lib.lib.b200ode_solve_ivp_device(h.ptr, ct.byref(pb), ct.c_void_p(stream)))
    torch.cuda.synchronize()
    assert np.array_equal(d_res.cpu().numpy(), whole._r)
    assert np.array_equal(d_t.cpu().numpy(), whole.t) and np.array_equal(d_y.cpu().numpy(), whole.y)
    assert np.array_equal(d_tev.cpu().numpy(), whole.t_event)
    # ... and with t_eval
    d_te = t(te)
    d_y2 = torch.full((B, len(te), 3), float("nan"), dtype=torch.float64, device=dev)
    pb.nt, pb.teval, pb.row_offset, pb.t_out, pb.y_out = len(te), d_te.data_ptr(), None, None, d_y2.data_ptr()
    _lib.check(_lib.lib.b200ode_solve_ivp_device(h.ptr, ct.byref(pb), ct.c_void_p(stream)))
    torch.cuda.synchronize()
    assert np.array_equal(d_res.cpu().numpy(), whole_te._r)
    y2 = d_y2.cpu().numpy()
    for b in range(0, B, 7):
        k = whole_te.nt_out[b]
        k = min(k, np.searchsorted(te, ts[b, 1], side="right"))
        assert np.array_equal(y2[b, :k], whole_te.y[b, :k])


def test_full_size_properties():
    """2^18 trajectories of the Lorenz sweep with a terminal event, device resident through the
    host path: determinism across calls and a seeded subsample bit-identical to the oracle."""
    nb = _nb()
    B = 1 << 18
    ts, y0, data = _lorenz_ensemble(B, seed=9)
    te = np.linspace(0.0, 6.0, 13)
    kw = dict(n_events=1, event_fcn=nb.builtin("lorenz_z"), data=data, rtol=1e-6, atol=1e-8, strict=True)
    a = nb.solve_ivp(nb.builtin("lorenz"), ts, y0, te, **kw)
    b = nb.solve_ivp(nb.builtin("lorenz"), ts, y0, te, **kw)
    assert np.array_equal(a._r, b._r) and np.array_equal(a.t_event, b.t_event)
    assert a.success.all() and 0.05 < a.event_found.mean() < 0.95
    # an event time lies inside the span, and the event state is on the level
    ev = a.event_found
    assert (a.t_event[ev] > 0).all() and (a.t_event[ev] <= ts[ev, 1]).all()
    assert np.abs(a.y_event[ev, 2] - data[ev, 3]).max() < 1e-5
    idx = np.random.default_rng(1).choice(B, 64, replace=False)
    ref = orc.solve_ivp_batch(orc.builtin_rhs("lorenz"), ts[idx], y0[idx], te, 1,
                              orc.builtin_rhs("events_lorenz_z"), data[idx], rtol=1e-6, atol=1e-8)
    rr = ref["result"]
    assert np.array_equal(a.t_event[idx], rr["t_event"]) and np.array_equal(a.nstep[idx], rr["nstep"])
    for j, bb in enumerate(idx):
        k = min(rr["nt_out"][j], np.searchsorted(te, ts[bb, 1], side="right"))
        assert np.array_equal(a.y[bb, :k], ref["y"][j, :k])


def ev_brus(t, u, out, p):
    out[0] = u[5] - 1.0
    out[1] = u[40] - 3.0


@pytest.mark.parametrize("mode", ["t_eval", "two_pass"])
def test_warp_kernel_n64_strict_bitwise(mode):
    """Systems of more than 8 equations run one trajectory per warp (solve_ivp_warp.cuh): the
    n = 64 Brusselator with two events, numba-compiled right-hand side and events function."""
    import numba
    nb = _nb()
    B = 40
    P = pr.brusselator_sweep(B)
    y0 = pr.brusselator_u0()
    rhs = numba.cfunc(nb.lsoda_sig)(pr.brusselator_rhs)
    ev = numba.cfunc(nb.lsoda_sig)(ev_brus)
    te = np.linspace(0.0, 2.0, 21) if mode == "t_eval" else None
    got = nb.solve_ivp(rhs.address, np.array([0.0, 2.0]), y0, te, n_events=2, event_fcn=ev.address,
                       data=P, rtol=1e-6, atol=1e-8, strict=True)
    ref = orc.solve_ivp_batch(rhs.address, [0.0, 2.0], y0, te, 2, ev.address, P, rtol=1e-6, atol=1e-8,
                              cap=6000)
    rr = ref["result"]
    assert set(rr["ind_event"][rr["event_found"] == 1]) == {0, 1}
    for k, f in (("message_code", "message"), ("nfev", "nfev"), ("nt_out", "nt_out"),
                 ("ind_event", "ind_event"), ("idid", "idid"), ("nstep", "nstep"), ("naccpt", "naccpt"),
                 ("nrejct", "nrejct"), ("rows_needed", "rows_needed")):
        assert np.array_equal(getattr(got, k), rr[f]), k
    assert np.array_equal(got.t_event, rr["t_event"]) and np.array_equal(got.y_event, ref["y_event"])
    for b in range(B):
        k = rr["nt_out"][b]
        r = got[b]
        assert len(r.t) == k
        if te is not None and rr["idid"][b] < 0:
            continue  # a failed trajectory's rows beyond the failure are uninitialised
        assert np.array_equal(r.y, ref["y"][b, :k]), b
        if te is None:
            assert np.array_equal(r.t, ref["t"][b, :k]), b


def test_warp_kernel_n12_fast_mode_and_failures():
    nb = _nb()
    B = 64
    rng = np.random.default_rng(2)
    P = np.column_stack([rng.uniform(0.5, 2.0, B), np.full(B, 12.0)])
    y0 = np.ones(12)
    te = np.linspace(0.0, 1.0, 11)
    import numba
    cf = numba.cfunc(nb.lsoda_sig)(pr.chain_rhs)
    got = nb.solve_ivp(cf.address, np.array([0.0, 1.0]), y0, te, data=P, rtol=1e-7, atol=1e-9)
    ref = orc.solve_ivp_batch(cf.address, [0.0, 1.0], y0, te, data=P, rtol=1e-7, atol=1e-9)
    assert got.success.all() and (ref["result"]["success"] == 1).all()
    assert (np.abs(got.y - ref["y"]) / (np.abs(ref["y"]) + 1.0)).max() <= 10 * 1e-7
    # decreasing t_eval against an increasing span: the reference's message
    bad = nb.solve_ivp(cf.address, np.array([0.0, 1.0]), y0, te[::-1].copy(), data=P[:3])
    assert not bad.success.any() and "does not always increase" in bad[0].message
    with pytest.raises(nb.B200OdeError, match="exceeds 128"):
        nb.solve_ivp(cf.address, np.array([0.0, 1.0]), np.ones(129), te, data=P[:3])


def test_golden_vectors_strict_bitwise():
    """The committed fixtures (tests/golden/solve_ivp_oracle.npz: the cases of the reference's
    solve_ivp.ipynb and tests, with the event times the reference printed) through the C ABI."""
    import _ivp_golden as gold
    nb = _nb()
    for name, (rhs, span, y0, te, ne, ev, data, kw, printed) in gold.CASES.items():
        r = nb.solve_ivp(nb.builtin(rhs), span, y0, te, n_events=ne,
                         event_fcn=nb.builtin(ev) if ev else 0, data=data, strict=True, **kw)
        f = dict(success=r.success, message=orc.IVP_MESSAGES.index(r.message), nfev=r.nfev,
                 nt_out=len(r.t), event_found=r.event_found, ind_event=r.ind_event, **r.counters)
        gold.check(name, f, r.t, r.y, r.t_event, r.y_event)


def ev_chain(t, u, out, p):
    out[0] = u[0] - 0.4
    out[1] = u[1] - 2.0  # never


def stiff1(t, u, du, p):
    du[0] = -p[0] * (u[0] - 1.0) + p[1]


def ev1(t, u, out, p):
    out[0] = u[0] - 0.5


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 7, 8, 9, 33])
def test_every_system_size_strict_bitwise(n):
    """One thread-kernel image per n <= 8, the warp kernel above: each against the oracle."""
    import numba
    nb = _nb()
    B = 37
    rng = np.random.default_rng(n)
    if n == 1:
        rhs, ev, ne = numba.cfunc(nb.lsoda_sig)(stiff1), numba.cfunc(nb.lsoda_sig)(ev1), 1
        data = np.column_stack([rng.uniform(0.5, 3.0, B), rng.uniform(-0.2, 0.2, B)])
        y0 = np.zeros(1)
    else:
        rhs, ev, ne = numba.cfunc(nb.lsoda_sig)(pr.chain_rhs), numba.cfunc(nb.lsoda_sig)(ev_chain), 2
        data = np.column_stack([rng.uniform(0.8, 2.5, B), np.full(B, float(n))])
        y0 = np.linspace(1.0, 1.5, n)
    for te in (None, np.linspace(0.0, 3.0, 13)):
        got = nb.solve_ivp(rhs.address, np.array([0.0, 3.0]), y0, te, n_events=ne, event_fcn=ev.address,
                           data=data, rtol=1e-7, atol=1e-10, strict=True)
        ref = orc.solve_ivp_batch(rhs.address, [0.0, 3.0], y0, te, ne, ev.address, data, rtol=1e-7,
                                  atol=1e-10, cap=3000)
        rr = ref["result"]
        assert rr["event_found"].any() and (rr["success"] == 1).all()
        for k, f in (("nfev", "nfev"), ("nt_out", "nt_out"), ("ind_event", "ind_event"), ("idid", "idid"),
                     ("nstep", "nstep"), ("naccpt", "naccpt"), ("nrejct", "nrejct")):
            assert np.array_equal(getattr(got, k), rr[f]), (n, k)
        assert np.array_equal(got.t_event, rr["t_event"]) and np.array_equal(got.y_event, ref["y_event"])
        for b in range(B):
            k = rr["nt_out"][b]
            r = got[b]
            assert np.array_equal(r.y, ref["y"][b, :k]), (n, b)
            if te is None:
                assert np.array_equal(r.t, ref["t"][b, :k]), (n, b)
