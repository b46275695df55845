"""The solve_ivp cores -- thread per trajectory (numbalsoda_b200/csrc/solve_ivp_core.cuh, n <= 8) and
warp per trajectory (solve_ivp_warp.cuh, one emulated lane): DOP853 with dense output after every
step, terminal events by Brent's method -- compiled as host C++ against the CPU oracle (oracle/solve_ivp_oracle.c, pinned by the reference's notebook outputs):
every field of the reference's ODEResult bit for bit."""
import ctypes as ct
import os
import subprocess

import numpy as np
import pytest

import _oracle as orc

nb = pytest.importorskip("numba")
from numba import cfunc, types  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
HDRS = [os.path.join(ROOT, "numbalsoda_b200", "csrc", f)
        for f in ("solve_ivp_core.cuh", "solve_ivp_warp.cuh", "b200ode_warp.cuh", "b200ode_launch.h",
                  "dop853_tableau.cuh", "b200ode_pow.cuh", "b200ode_fastmath.cuh")]
NRESULT = 12


class IvpLaunch(ct.Structure):  # B200IvpLaunch, numbalsoda_b200/csrc/b200ode_launch.h
    _fields_ = [("B", ct.c_longlong), ("t_span", ct.c_void_p), ("t_span_stride", ct.c_longlong),
                ("y0", ct.c_void_p), ("y0_stride", ct.c_longlong), ("data", ct.c_void_p),
                ("data_stride", ct.c_longlong), ("np", ct.c_int), ("nt", ct.c_int),
                ("t_eval", ct.c_void_p), ("n_events", ct.c_int),
                ("cap", ct.c_longlong), ("row_offset", ct.c_void_p), ("first_step", ct.c_double), ("max_step", ct.c_double),
                ("rtol", ct.c_double), ("atol", ct.c_double), ("rtol_b", ct.c_void_p),
                ("atol_b", ct.c_void_p), ("t_out", ct.c_void_p), ("y_out", ct.c_void_p),
                ("result", ct.c_void_p), ("t_event", ct.c_void_p), ("y_event", ct.c_void_p),
                ("next", ct.c_void_p)]


def _compile(stem, fast=False):
    """fast: the throughput link mode's arithmetic (b200ode_fastmath.cuh) with libm stand-ins."""
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    src = os.path.join(HERE, stem + ".cpp")
    so = os.path.join(HERE, "_build", "lib%s%s.so" % (stem, "_fast" if fast else ""))
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in [src] + HDRS):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-mfma", "-fPIC", "-shared",
                               "-Wno-unknown-pragmas"] + (["-DIVP_HOST_FAST=1"] if fast else []) +
                              ["-o", so, src])
    return ct.CDLL(so)


class _Core:
    def __init__(self, fn):
        self.solve_ivp_core_host = fn


@pytest.fixture(scope="module", params=["thread", "warp"])
def lib(request):
    """Both cores behind one entry-point name (same B200IvpLaunch contract)."""
    if request.param == "thread":
        return _Core(_compile("solve_ivp_core_host").solve_ivp_core_host)
    return _Core(_compile("solve_ivp_warp_host").solve_ivp_warp_host)


def run(lib, f, t_span, y0, t_eval=None, n_events=0, events=0, data=None, first_step=0.0,
        max_step=0.0, rtol=1e-3, atol=1e-6, cap=20000):
    """Batched call of the host-compiled core -> list of dicts shaped like _oracle.solve_ivp's."""
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    ts = np.ascontiguousarray(t_span, dtype=np.float64)
    te = np.ascontiguousarray([] if t_eval is None else t_eval, dtype=np.float64)
    data = np.ascontiguousarray([0.0] if data is None else data, dtype=np.float64)
    B = max(y0.shape[0] if y0.ndim == 2 else 1, ts.shape[0] if ts.ndim == 2 else 1,
            data.shape[0] if data.ndim == 2 else 1)
    n, nt = y0.shape[-1], len(te)
    rows = nt if nt else cap
    t_out = np.full((B, rows), np.nan)
    y_out = np.full((B, rows, n), np.nan)
    res = np.full((B, NRESULT), -77, np.int32)
    t_event = np.full(B, np.nan)
    y_event = np.full((B, n), np.nan)
    p = lambda a: a.ctypes.data  # noqa: E731
    L = IvpLaunch(B=B, t_span=p(ts), t_span_stride=2 if ts.ndim == 2 else 0, y0=p(y0),
                  y0_stride=n if y0.ndim == 2 else 0, data=p(data),
                  data_stride=data.shape[1] * 8 if data.ndim == 2 else 0, np=-1, nt=nt,
                  t_eval=p(te), n_events=n_events, cap=rows, row_offset=None, first_step=first_step, max_step=max_step, rtol=rtol,
                  atol=atol, rtol_b=None, atol_b=None, t_out=p(t_out), y_out=p(y_out),
                  result=p(res), t_event=p(t_event), y_event=p(y_event), next=None)
    assert lib.solve_ivp_core_host(ct.c_void_p(f), ct.c_void_p(events), n, ct.byref(L)) == 0
    out = []
    for b in range(B):
        r = res[b]
        k = r[3]
        out.append(dict(message=orc.IVP_MESSAGES[r[1]], nfev=int(r[2]), success=bool(r[0]),
                        t=(te[:k].copy() if nt else t_out[b, :k]), y=y_out[b, :k],
                        event_found=bool(r[5]), ind_event=int(r[6]), t_event=float(t_event[b]),
                        y_event=y_event[b], idid=int(r[7]), nstep=int(r[8]), naccpt=int(r[9]),
                        nrejct=int(r[10]), rows_needed=int(r[4])))
    return out


def same(a, b):
    for k in ("message", "nfev", "success", "event_found", "ind_event", "idid", "nstep", "naccpt",
              "nrejct", "rows_needed"):
        assert a[k] == b[k], (k, a[k], b[k])
    assert a["t_event"] == b["t_event"]
    assert np.array_equal(a["y_event"], b["y_event"])
    assert np.array_equal(a["t"], b["t"]) and np.array_equal(a["y"], b["y"], equal_nan=True)


SIG = types.void(types.double, types.CPointer(types.double), types.CPointer(types.double),
                 types.CPointer(types.double))


@cfunc(SIG)
def lv(t, u, du, p):
    du[0] = u[0] - u[0] * u[1] * p[0]
    du[1] = u[0] * u[1] - u[1] * p[1]


@cfunc(SIG)
def ev_two(t, u, out, p):  # two events: u1 == 3 (never first) and u0 == 1
    out[0] = u[1] - 30.0
    out[1] = u[0] - 1.0


@cfunc(SIG)
def ev_exact_zero(t, u, out, p):
    out[0] = t - 0.5


@cfunc(SIG)
def lorenz(t, u, du, p):
    du[0] = p[0] * (u[1] - u[0])
    du[1] = u[0] * (p[1] - u[2]) - u[1]
    du[2] = u[0] * u[1] - p[2] * u[2]


@cfunc(SIG)
def ev_lorenz(t, u, out, p):
    out[0] = u[2] - 40.0


@cfunc(SIG)
def stiff(t, u, du, p):
    du[0] = -p[0] * (u[0] - np.cos(t))


U0 = np.array([5.0, 0.8])
ONES = np.array([1.0, 1.0])
TE = np.linspace(0.0, 50.0, 1000)
SPAN = np.array([0.0, 50.0])


@pytest.mark.parametrize("kw", [
    dict(t_eval=TE, n_events=2, events="ev_two", rtol=1e-6),
    dict(t_eval=None, n_events=2, events="ev_two", rtol=1e-6),
    dict(t_eval=TE, rtol=1e-8, atol=1e-8),
    dict(t_eval=None, rtol=1e-8, atol=1e-8),
    dict(t_eval=None, rtol=1e-6, max_step=0.01, span=[0.0, 5.0]),
    dict(t_eval=None, rtol=1e-6, first_step=1e-4, span=[0.0, 5.0]),
    dict(t_eval=-TE, span=[0.0, -50.0], rtol=1e-9, atol=1e-30),
    dict(t_eval=None, span=[0.0, -3.0], n_events=2, events="ev_two"),
    dict(t_eval=np.array([0.0, 2.0, 1.0])),
    dict(t_eval=-TE),                      # increasing span, decreasing t_eval
    dict(t_eval=None, rtol=1e-8, atol=1e-8, cap=10),
    dict(t_eval=TE[:1]),                   # a single output time
    dict(t_eval=np.array([0.0, 0.0, 1.0, 1.0, 50.0, 50.0])),
    dict(t_eval=None, n_events=1, events="ev_exact_zero", span=[0.0, 1.0], first_step=0.5, rtol=1e-2),
])
def test_lotka_volterra_cases(lib, kw):
    kw = dict(kw)
    span = np.array(kw.pop("span", SPAN), dtype=np.float64)
    ev = {"ev_two": ev_two, "ev_exact_zero": ev_exact_zero}[kw.pop("events")].address if "events" in kw else 0
    te = kw.pop("t_eval")
    ref = orc.solve_ivp(lv.address, span, U0, te, events=ev, data=ONES, **kw)
    got = run(lib, lv.address, span, U0, te, events=ev, data=ONES, **kw)[0]
    same(got, ref)


def test_batched_sweep_with_events(lib):
    """Per-trajectory t_span, y0 and data; Lorenz trajectories that do and do not hit z = 40."""
    rng = np.random.default_rng(5)
    B = 24
    data = np.column_stack([rng.uniform(9, 11, B), rng.uniform(20, 36, B), rng.uniform(2, 3.33, B)])
    y0 = np.column_stack([rng.uniform(0.5, 1.5, B), rng.uniform(-0.5, 0.5, B), rng.uniform(0, 1, B)])
    ts = np.column_stack([np.zeros(B), rng.uniform(0.3, 6.0, B)])
    for te in (None, np.linspace(0.0, 6.0, 61)):
        got = run(lib, lorenz.address, ts, y0, te, 1, ev_lorenz.address, data, rtol=1e-7, atol=1e-9,
                  cap=2000)
        hits = 0
        for b in range(B):
            ref = orc.solve_ivp(lorenz.address, ts[b], y0[b], te, 1, ev_lorenz.address, data[b],
                                rtol=1e-7, atol=1e-9, cap=2000)
            same(got[b], ref)
            hits += ref["event_found"]
        assert 0 < hits < B


def test_stiffness_abort_and_step_limit(lib):
    """idid = -4 (iprint = 0: the stiffness detector stops the integration, module :642-650)."""
    y0 = np.array([0.0])
    for lam, span in ((1e4, [0.0, 100.0]),):
        ref = orc.solve_ivp(stiff.address, span, y0, data=np.array([lam]), rtol=1e-3, cap=200000)
        got = run(lib, stiff.address, span, y0, data=np.array([lam]), rtol=1e-3, cap=200000)[0]
        assert ref["idid"] == -4 and not ref["success"] and ref["nfev"] == 0
        same(got, ref)


def test_ragged_rows(lib):
    """row_offset[B+1]: exactly sized rows after a counting pass with cap = 0."""
    rng = np.random.default_rng(8)
    B = 12
    data = np.column_stack([rng.uniform(9, 11, B), rng.uniform(20, 36, B), rng.uniform(2, 3.33, B)])
    ts = np.column_stack([np.zeros(B), rng.uniform(0.3, 3.0, B)])
    y0 = np.array([1.0, 0.0, 0.0])
    need = np.array([orc.solve_ivp(lorenz.address, ts[b], y0, data=data[b], rtol=1e-7)["rows_needed"]
                     for b in range(B)])
    off = np.zeros(B + 1, np.int64)
    np.cumsum(need, out=off[1:])
    R, n = int(off[-1]), 3
    t_out, y_out = np.full(R, np.nan), np.full((R, n), np.nan)
    res = np.zeros((B, NRESULT), np.int32)
    t_event, y_event = np.zeros(B), np.zeros((B, n))
    p = lambda a: a.ctypes.data  # noqa: E731
    L = IvpLaunch(B=B, t_span=p(ts), t_span_stride=2, y0=p(y0), y0_stride=0, data=p(data), data_stride=24,
                  np=-1, nt=0, t_eval=None, n_events=0, cap=0, row_offset=p(off), first_step=0.0,
                  max_step=0.0, rtol=1e-7, atol=1e-6, rtol_b=None, atol_b=None, t_out=p(t_out),
                  y_out=p(y_out), result=p(res), t_event=p(t_event), y_event=p(y_event), next=None)
    assert lib.solve_ivp_core_host(ct.c_void_p(lorenz.address), None, n, ct.byref(L)) == 0
    assert np.array_equal(res[:, 3], need) and res[:, 0].all() and not np.isnan(y_out).any()
    for b in range(B):
        ref = orc.solve_ivp(lorenz.address, ts[b], y0, data=data[b], rtol=1e-7)
        assert np.array_equal(t_out[off[b]:off[b + 1]], ref["t"])
        assert np.array_equal(y_out[off[b]:off[b + 1]], ref["y"])


def test_brusselator_n64_with_event():
    """The warp core at a size the thread core does not cover."""
    warp = _compile("solve_ivp_warp_host").solve_ivp_warp_host
    import _problems as pr
    f = orc.builtin_rhs("brusselator1d")
    y0 = pr.brusselator_u0()
    P = pr.brusselator_sweep(3)

    @cfunc(SIG)
    def ev(t, u, out, p):
        out[0] = u[5] - 1.0    # crossed downwards by the third parameter set
        out[1] = u[40] - 3.0   # crossed downwards by the first

    seen = set()
    for te in (None, np.linspace(0.0, 2.0, 21)):
        for b in range(3):
            ref = orc.solve_ivp(f, [0.0, 2.0], y0, te, 2, ev.address, P[b], rtol=1e-6, atol=1e-8, cap=5000)
            got = run(_Core(warp), f, [0.0, 2.0], y0, te, 2, ev.address, P[b], rtol=1e-6, atol=1e-8,
                      cap=5000)[0]
            same(got, ref)
            seen.add((ref["idid"], ref["ind_event"] if ref["event_found"] else -1))
    assert seen == {(2, 0), (2, 1)}, seen  # both events were exercised
    # ... and the stiffness abort (idid = -4: iprint = 0, module :642-650) without events
    ref = orc.solve_ivp(f, [0.0, 2.0], y0, None, data=P[1], rtol=1e-6, atol=1e-8, cap=5000)
    got = run(_Core(warp), f, [0.0, 2.0], y0, None, data=P[1], rtol=1e-6, atol=1e-8, cap=5000)[0]
    same(got, ref)
    assert ref["idid"] == -4 and not ref["success"] and ref["nstep"] > 1000


def test_throughput_mode_emulation():
    """The default link mode's controller (reciprocals, single-precision err**(-1/16)) emulated on
    the host: same event decisions as the oracle, event times to the Brent tolerance, states
    within 10 rtol (north_star's bound), step counts within a few."""
    fast = _Core(_compile("solve_ivp_core_host", fast=True).solve_ivp_core_host)
    rng = np.random.default_rng(11)
    B = 200
    data = np.column_stack([rng.uniform(9, 11, B), rng.uniform(20, 36, B), rng.uniform(2, 3.33, B)])
    y0 = np.column_stack([rng.uniform(0.5, 1.5, B), rng.uniform(-0.5, 0.5, B), rng.uniform(0, 1, B)])
    ts = np.column_stack([np.zeros(B), rng.uniform(0.3, 2.0, B)])
    te = np.linspace(0.0, 2.0, 21)
    rtol = 1e-8
    got = run(fast, lorenz.address, ts, y0, te, 1, ev_lorenz.address, data, rtol=rtol, atol=rtol)
    worst = 0.0
    for b in range(B):
        ref = orc.solve_ivp(lorenz.address, ts[b], y0[b], te, 1, ev_lorenz.address, data[b], rtol=rtol,
                            atol=rtol)
        g = got[b]
        assert (g["success"], g["event_found"], g["message"]) == (ref["success"], ref["event_found"],
                                                                  ref["message"])
        assert len(g["t"]) == len(ref["t"]) and abs(g["nstep"] - ref["nstep"]) <= 3
        if ref["event_found"]:
            assert abs(g["t_event"] - ref["t_event"]) < 5e-8
        k = min(len(ref["t"]), np.searchsorted(te, ts[b, 1], side="right"))
        if k:
            worst = max(worst, (np.abs(g["y"][:k] - ref["y"][:k]) / (np.abs(ref["y"][:k]) + 1.0)).max())
    assert 0.0 < worst <= 10 * rtol, worst


def test_golden_vectors(lib):
    """Both host-compiled cores against the committed fixtures."""
    import _ivp_golden as gold
    for name, (rhs, span, y0, te, ne, ev, data, kw, printed) in gold.CASES.items():
        r = run(lib, orc.builtin_rhs(rhs), span, y0, te, ne, orc.builtin_rhs("events_" + ev) if ev else 0,
                data, cap=4000, **kw)[0]
        f = dict(r, message=orc.IVP_MESSAGES.index(r["message"]), nt_out=len(r["t"]))
        gold.check(name, f, r["t"], r["y"], r["t_event"], r["y_event"])


@cfunc(SIG)
def chain(t, u, du, p):  # _problems.chain_rhs: a nonlinear reaction chain of any size n = int(p[1])
    n = int(p[1])
    k = p[0]
    du[0] = -k * u[0] + 0.5 * u[1] * u[1]
    for i in range(1, n - 1):
        du[i] = k * (u[i - 1] - u[i]) / (1.0 + i) + 0.1 * u[i + 1] - 0.1 * u[i] * u[i]
    du[n - 1] = k * (u[n - 2] - u[n - 1]) / n


@cfunc(SIG)
def ev_chain(t, u, out, p):
    out[0] = u[0] - 0.4
    out[1] = u[1] - 2.0  # never


@pytest.mark.parametrize("n", [2, 3, 4, 5, 6, 7, 8])
def test_every_thread_kernel_size(lib, n):
    """One image per system size n <= 8 on the device: every instantiation of the core."""
    y0 = np.linspace(1.0, 1.5, n)
    data = np.array([1.7, float(n)])
    for te in (None, np.linspace(0.0, 3.0, 13)):
        ref = orc.solve_ivp(chain.address, [0.0, 3.0], y0, te, 2, ev_chain.address, data, rtol=1e-7,
                            atol=1e-10)
        got = run(lib, chain.address, [0.0, 3.0], y0, te, 2, ev_chain.address, data, rtol=1e-7,
                  atol=1e-10)[0]
        same(got, ref)
        assert ref["event_found"] and ref["ind_event"] == 0 and ref["naccpt"] > 3
