"""ctypes bindings of the CPU oracle (oracle/liboracle.so) and, when it was built,
of the compiled reference LSODA (oracle/_ref/*.so).  Test infrastructure only."""
import ctypes as ct
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")
REFSRC = "/root/reference/src/LSODA.cpp"


class LsodaStats(ct.Structure):
    _fields_ = [(k, ct.c_int) for k in
                "nst nfe nje nswitch nqu meth istate rows nattempt nhnil".split()] + \
               [("tn", ct.c_double), ("hu", ct.c_double)]


class Dop853Stats(ct.Structure):
    _fields_ = [(k, ct.c_int) for k in
                "nfcn nstep naccpt nrejct nevent idid rows pad_".split()] + \
               [("hlast", ct.c_double), ("xlast", ct.c_double)]


def _build():
    so = os.path.join(ODIR, "liboracle.so")
    srcs = [os.path.join(ODIR, f) for f in os.listdir(ODIR) if f.endswith((".c", ".h"))]
    if not os.path.exists(so) or any(os.path.getmtime(f) > os.path.getmtime(so) for f in srcs):
        # (make compares the same time stamps: a library older than its sources is rebuilt)
        subprocess.check_call(["make", "-C", ODIR, ])
    if os.path.exists(REFSRC) and not os.path.exists(os.path.join(ODIR, "_ref", "libref_probe.so")):
        subprocess.check_call(["make", "-C", ODIR, "ref", ])


_build()
lib = ct.CDLL(os.path.join(ODIR, "liboracle.so"))
lib.oracle_builtin_rhs.restype = ct.c_void_p
lib.oracle_builtin_rhs.argtypes = [ct.c_char_p]
_probe_path = os.path.join(ODIR, "_ref", "libref_probe.so")
_ref_path = os.path.join(ODIR, "_ref", "liblsoda.so")
ref_probe = ct.CDLL(_probe_path) if os.path.exists(_probe_path) else None
ref_lib = ct.CDLL(_ref_path) if os.path.exists(_ref_path) else None
HAVE_REF = ref_probe is not None


def _p(a):
    return a.ctypes.data_as(ct.c_void_p)


def builtin_rhs(name):
    ptr = lib.oracle_builtin_rhs(name.encode())
    assert ptr, name
    return ptr


def _prep(u0, t_eval, data):
    u0 = np.ascontiguousarray(u0, dtype=np.float64)
    te = np.ascontiguousarray(t_eval, dtype=np.float64)
    data = np.ascontiguousarray(data)
    us = np.full((len(te), len(u0)), np.nan)
    return u0, te, data, us


def lsoda(fptr, u0, t_eval, data, rtol=1e-3, atol=1e-6, mxstep=10000, exit_on_warning=False):
    u0, te, data, us = _prep(u0, t_eval, data)
    ok = ct.c_int(999)
    st = LsodaStats()
    lib.oracle_lsoda(ct.c_void_p(fptr), len(u0), _p(u0), _p(data), len(te), _p(te), _p(us),
                     ct.c_double(rtol), ct.c_double(atol), int(mxstep), ct.byref(ok),
                     int(exit_on_warning), ct.byref(st))
    return us, ok.value == 1, st


def dop853(fptr, u0, t_eval, data, rtol=1e-3, atol=1e-6, mxstep=10000):
    u0, te, data, us = _prep(u0, t_eval, data)
    ok = ct.c_int(999)
    st = Dop853Stats()
    lib.oracle_dop853(ct.c_void_p(fptr), len(u0), _p(u0), _p(data), len(te), _p(te), _p(us),
                      ct.c_double(rtol), ct.c_double(atol), int(mxstep), ct.byref(ok),
                      ct.byref(st))
    return us, ok.value == 1, st


def ref_lsoda(fptr, u0, t_eval, data, rtol=1e-3, atol=1e-6, mxstep=10000,
              exit_on_warning=False):
    """The compiled reference, driven like src/wrapper.cpp drives it, plus counters:
    [nst, nfe, nje, nswitch, nqu, istate, rows, meth]."""
    u0, te, data, us = _prep(u0, t_eval, data)
    ok = ct.c_int(999)
    st = (ct.c_int * 8)()
    ref_probe.ref_lsoda_stats(ct.c_void_p(fptr), len(u0), _p(u0), _p(data), len(te), _p(te),
                              _p(us), ct.c_double(rtol), ct.c_double(atol), int(mxstep),
                              ct.byref(ok), int(exit_on_warning), st)
    return us, ok.value == 1, list(st)


def ref_lsoda_wrapper(fptr, u0, t_eval, data, rtol=1e-3, atol=1e-6, mxstep=10000,
                      exit_on_warning=False):
    """The reference's own extern "C" entry point (src/wrapper.cpp:10)."""
    u0, te, data, us = _prep(u0, t_eval, data)
    ok = ct.c_int(999)
    ref_lib.lsoda_wrapper(ct.c_void_p(fptr), len(u0), _p(u0), _p(data), len(te), _p(te), _p(us),
                          ct.c_double(rtol), ct.c_double(atol), int(mxstep), ct.byref(ok),
                          int(exit_on_warning))
    return us, ok.value == 1


def ref_lsoda_vtol(fptr, u0, t_eval, data, rtol_v, atol_v, mxstep=10000, exit_on_warning=False):
    """The compiled reference with per-component tolerances: its private rtol_ / atol_ vectors
    filled per component, itol_ = 4, driven through the public LSODA::lsoda with lsoda_update's
    fixed options (oracle/ref_shim.cpp)."""
    u0, te, data, us = _prep(u0, t_eval, data)
    rv = np.ascontiguousarray(rtol_v, dtype=np.float64)
    av = np.ascontiguousarray(atol_v, dtype=np.float64)
    assert rv.shape == av.shape == u0.shape
    ok = ct.c_int(999)
    st = (ct.c_int * 8)()
    ref_probe.ref_lsoda_stats_vtol(ct.c_void_p(fptr), len(u0), _p(u0), _p(data), len(te), _p(te),
                                   _p(us), _p(rv), _p(av), int(mxstep), ct.byref(ok),
                                   int(exit_on_warning), st)
    return us, ok.value == 1, list(st)


def _tol_arg(x, B, n):
    """-> (scalar, vector or None, stride between trajectories) of a tolerance given as a
    scalar, [B] (per trajectory), [n] / [1,n] (per component) or [B,n]"""
    a = np.asarray(x, dtype=np.float64)
    if a.ndim == 0:
        return float(a), None, 0
    if a.ndim == 1 and a.shape[0] == B and not (B == 1 and n != 1):
        return 0.0, np.ascontiguousarray(np.repeat(a[:, None], n, axis=1)), n
    if a.ndim == 1 and a.shape[0] == n:
        return 0.0, np.ascontiguousarray(a), 0
    if a.ndim == 2 and a.shape == (1, n):
        return 0.0, np.ascontiguousarray(a[0]), 0
    if a.ndim == 2 and a.shape == (B, n):
        return 0.0, np.ascontiguousarray(a), n
    raise ValueError("tolerance of shape %r for B = %d, n = %d" % (a.shape, B, n))


def solve_batch(solver, fptr, u0, t_eval, data, rtol, atol, mxstep=10000,
                exit_on_warning=False, nthreads=0, band=None):
    """solver: 'dop853' | 'lsoda'.  u0 [B,n] or [n]; data [B,p] or [p]; rtol / atol scalars,
    [B] (per trajectory), [n] or [1,n] (per component) or [B,n]."""
    u0 = np.ascontiguousarray(u0, dtype=np.float64)
    data = np.ascontiguousarray(data, dtype=np.float64)
    te = np.ascontiguousarray(t_eval, dtype=np.float64)
    B = u0.shape[0] if u0.ndim == 2 else (data.shape[0] if data.ndim == 2 else 1)
    n = u0.shape[-1]
    us = np.full((B, len(te), n), np.nan)
    ok = np.zeros(B, dtype=np.int32)
    cnt = np.zeros((B, 8), dtype=np.int32)
    rs, rv, rvs = _tol_arg(rtol, B, n)
    as_, av, avs = _tol_arg(atol, B, n)
    lib.oracle_batch_set_band(*(band if band is not None else (-1, -1)))  # lsoda: jt = 5 with (ml, mu)
    lib.oracle_solve_batch_v.restype = ct.c_int
    lib.oracle_solve_batch_v(0 if solver == "dop853" else 1, ct.c_void_p(fptr), ct.c_long(B), n,
                             _p(u0), ct.c_long(n if u0.ndim == 2 else 0), _p(data),
                             ct.c_long(data.shape[-1] * 8 if data.ndim == 2 else 0), len(te),
                             _p(te), _p(us), ct.c_double(rs), ct.c_double(as_),
                             _p(rv) if rv is not None else None, ct.c_long(rvs),
                             _p(av) if av is not None else None, ct.c_long(avs), int(mxstep),
                             int(exit_on_warning), _p(ok), _p(cnt), int(nthreads))
    lib.oracle_batch_set_band(-1, -1)
    return us, ok == 1, cnt


class IvpResult(ct.Structure):
    _fields_ = [(k, ct.c_int) for k in
                "success message nfev nt_out event_found ind_event idid nstep naccpt nrejct "
                "rows_needed pad_".split()] + [("t_event", ct.c_double)]


IVP_MESSAGES = ["The solver successfully reached the end of the integration interval.",
                "A termination event occurred.", "Failed to complete integration.",
                '"t_eval" does not always increase or decrease', "Integrator failed to initialize.",
                "brent failed", "output capacity exceeded"]


def solve_ivp(fptr, t_span, y0, t_eval=None, n_events=0, events=0, data=None, first_step=0.0,
              max_step=0.0, rtol=1e-3, atol=1e-6, cap=20000):
    """oracle_solve_ivp: returns a dict with the fields of the reference's ODEResult
    (numbalsoda/driver_solve_ivp.py:59-69) plus the private counters."""
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    ts = np.ascontiguousarray(t_span, dtype=np.float64)
    te = np.ascontiguousarray([] if t_eval is None else t_eval, dtype=np.float64)
    data = np.ascontiguousarray([0.0] if data is None else data)
    n, nt = len(y0), len(te)
    rows = nt if nt else cap
    t_out = np.full(rows, np.nan)
    y_out = np.full((rows, n), np.nan)
    y_event = np.zeros(n)
    res = IvpResult()
    lib.oracle_solve_ivp(ct.c_void_p(fptr), _p(ts), n, _p(y0), nt, _p(te), int(n_events),
                         ct.c_void_p(events), _p(data), ct.c_double(first_step),
                         ct.c_double(max_step), ct.c_double(rtol), ct.c_double(atol), int(cap),
                         _p(t_out), _p(y_out), _p(y_event), ct.byref(res))
    k = res.nt_out
    return dict(message=IVP_MESSAGES[res.message], nfev=res.nfev, success=bool(res.success),
                t=(te[:k].copy() if nt else t_out[:k]), y=y_out[:k], event_found=bool(res.event_found),
                ind_event=res.ind_event, t_event=res.t_event, y_event=y_event, idid=res.idid,
                nstep=res.nstep, naccpt=res.naccpt, nrejct=res.nrejct, rows_needed=res.rows_needed)


def solve_ivp_batch(fptr, t_span, y0, t_eval=None, n_events=0, events=0, data=None, first_step=0.0,
                    max_step=0.0, rtol=1e-3, atol=1e-6, cap=0, nthreads=0):
    """oracle_solve_ivp_batch over host threads.  t_span [2] or [B,2], y0 [n] or [B,n], data [p] or
    [B,p] -> dict of arrays: result[B] (IvpResult records), y[B,rows,n], t[B,rows] (no t_eval),
    y_event[B,n]."""
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    ts = np.ascontiguousarray(t_span, dtype=np.float64)
    te = np.ascontiguousarray([] if t_eval is None else t_eval, dtype=np.float64)
    data = np.ascontiguousarray([0.0] if data is None else data, dtype=np.float64)
    B = max([a.shape[0] for a in (y0, ts, data) if a.ndim == 2] or [1])
    n, nt = y0.shape[-1], len(te)
    rows = nt if nt else cap
    t_out = np.full((B, rows), np.nan)
    y_out = np.full((B, rows, n), np.nan)
    y_event = np.zeros((B, n))
    res = (IvpResult * B)()
    lib.oracle_solve_ivp_batch.restype = ct.c_int
    lib.oracle_solve_ivp_batch(ct.c_void_p(fptr), ct.c_void_p(events), ct.c_long(B), n, _p(ts),
                               ct.c_long(2 if ts.ndim == 2 else 0), _p(y0),
                               ct.c_long(n if y0.ndim == 2 else 0), _p(data),
                               ct.c_long(data.shape[-1] * 8 if data.ndim == 2 else 0), nt, _p(te),
                               int(n_events), ct.c_double(first_step), ct.c_double(max_step),
                               ct.c_double(rtol), ct.c_double(atol), int(cap), _p(t_out), _p(y_out),
                               _p(y_event), res, int(nthreads))
    rec = np.frombuffer(res, dtype=np.dtype([(k, np.int32) for k in
                                             "success message nfev nt_out event_found ind_event idid "
                                             "nstep naccpt nrejct rows_needed pad_".split()] +
                                            [("t_event", np.float64)])).copy()
    return dict(result=rec, y=y_out, t=t_out, y_event=y_event)
