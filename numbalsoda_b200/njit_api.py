"""lsoda / dop853 for ensembles, callable from inside @njit code.

The reference's drivers are @njit functions that call the native wrapper through a ctypes
function object, which numba lowers to a direct C call (numbalsoda/driver.py:28-43,
driver_dop853.py:23-35; README.md:60-67: "lsoda can be called within a jit-compiled numba
function").  The same works against libb200ode: `funcptr` -- the address of host machine code in
the reference -- becomes the address of a handle linked ahead of time in Python:

    h = numbalsoda_b200.njit_api.link(rhs, "lsoda", n)        # int, outside @njit

    @njit
    def many_sweeps(h, u0, t_eval, P):
        usol, success = lsoda_batched(h, u0, t_eval, P)       # usol[B,nt,n], success[B]
        ...

u0 is [B,n] (per trajectory) and data [B,p] float64, p <= 16.  Errors of the library (no device,
bad arguments) cannot raise inside nopython code with the library's message; they come back as
success[:] = False, and `last_status()` outside @njit gives the message of the last failure on
the thread.
"""
import ctypes as ct

import numpy as np

from . import _lib, rhs as _rhs

try:
    import numba as nb
except Exception as e:  # pragma: no cover
    raise ImportError("numbalsoda_b200.njit_api needs numba") from e

_vp, _ll, _i, _d = ct.c_void_p, ct.c_longlong, ct.c_int, ct.c_double
# separate function objects with integer-friendly prototypes (numba passes .ctypes.data ints)
_lsoda_c = _lib.lib.b200ode_lsoda_batched
_lsoda_c.argtypes = [_vp, _ll, _i, _vp, _ll, _vp, _ll, _i, _i, _vp, _vp, _d, _d, _i, _vp, _vp, _i, _i]
_lsoda_c.restype = _i
_dop853_c = _lib.lib.b200ode_dop853_batched
_dop853_c.argtypes = [_vp, _ll, _i, _vp, _ll, _vp, _ll, _i, _i, _vp, _vp, _d, _d, _i, _vp, _vp, _i]
_dop853_c.restype = _i

_ivp_c = _lib.lib.b200ode_solve_ivp_args
_ivp_c.argtypes = [_vp, _ll, _i, _vp, _ll, _vp, _ll, _vp, _ll, _i, _i, _vp, _i, _d, _d, _d, _d, _ll, _vp,
                   _vp, _vp, _vp, _vp, _vp, _i]
_ivp_c.restype = _i

_keep = []  # handles linked here live as long as the process (their address is all @njit code has)


def link(rhs, solver, neq, strict=False, exact_ctrl=False):
    """Link a right-hand side (cfunc address, cfunc, Python function or builtin(name)) with the
    `solver` ("lsoda" | "dop853") kernels for systems of `neq` equations; returns the handle's
    address as an int, the stand-in for the reference's `funcptr` inside @njit code."""
    sid = {"lsoda": _lib.LSODA, "dop853": _lib.DOP853}[solver]
    h = _rhs.get_handle(rhs, sid, neq, strict=strict, exact_ctrl=exact_ctrl)
    _keep.append(h)
    return int(h.ptr.value)


def link_ivp(rhs, neq, n_events=0, event_fcn=None, strict=False, exact_ctrl=False):
    """solve_ivp: link the right-hand side and the events function; returns the handle's address."""
    h = _rhs.get_ivp_handle(rhs, neq, n_events, event_fcn, strict=strict, exact_ctrl=exact_ctrl)
    _keep.append(h)
    return int(h.ptr.value)


def last_status():
    return _lib.lib.b200ode_last_error().decode()


@nb.njit
def lsoda_batched(handle, u0, t_eval, data, rtol=1.0e-3, atol=1.0e-6, mxstep=10000,
                  exit_on_warning=False, device=0):
    """Batched numbalsoda.lsoda (driver.py:28-43) from nopython code -> (usol[B,nt,n], success[B])."""
    B, neq = u0.shape
    nt = len(t_eval)
    usol = np.empty((B, nt, neq), dtype=np.float64)
    success = np.full(B, 999, np.int32)
    rc = _lsoda_c(handle, B, neq, u0.ctypes.data, neq, data.ctypes.data, data.shape[1] * 8,
                  data.shape[1], nt, t_eval.ctypes.data, usol.ctypes.data, rtol, atol, mxstep,
                  success.ctypes.data, 0, exit_on_warning, device)
    if rc != 0:
        success[:] = 0
    return usol, success == 1


@nb.njit
def dop853_batched(handle, u0, t_eval, data, rtol=1.0e-3, atol=1.0e-6, mxstep=10000, device=0):
    """Batched numbalsoda.dop853 (driver_dop853.py:23-35) from nopython code."""
    B, neq = u0.shape
    nt = len(t_eval)
    usol = np.empty((B, nt, neq), dtype=np.float64)
    success = np.full(B, 999, np.int32)
    rc = _dop853_c(handle, B, neq, u0.ctypes.data, neq, data.ctypes.data, data.shape[1] * 8,
                   data.shape[1], nt, t_eval.ctypes.data, usol.ctypes.data, rtol, atol, mxstep,
                   success.ctypes.data, 0, device)
    if rc != 0:
        success[:] = 0
    return usol, success == 1


@nb.njit
def solve_ivp_batched(handle, t_span, y0, t_eval, data, n_events=0, first_step=0.0, max_step=0.0,
                      rtol=1.0e-3, atol=1.0e-6, device=0):
    """Batched numbalsoda.solve_ivp (driver_solve_ivp.py:86-218) from nopython code.
    t_span [B,2], y0 [B,n], data [B,p] float64; t_eval [nt] or empty (every accepted step is a
    row: two passes, the first counts).  Returns (t[R], y[R,n], row_offset[B+1], result[B,12],
    t_event[B], y_event[B,n]): the rows of trajectory b are row_offset[b] : row_offset[b] +
    result[b, 3]; result columns as in include/b200ode.h (success, message, nfev, nt_out, ...).
    A failure of the library itself (no device, bad arguments) leaves result[:, 0] = 0."""
    B, neq = y0.shape
    nt = len(t_eval)
    result = np.zeros((B, 12), np.int32)
    t_event = np.zeros(B)
    y_event = np.zeros((B, neq))
    row_offset = np.zeros(B + 1, np.int64)
    p = data.shape[1]
    if nt > 0:
        for b in range(B):
            row_offset[b + 1] = (b + 1) * nt
        t = np.empty(B * nt)
        for b in range(B):
            t[b * nt:(b + 1) * nt] = t_eval
        y = np.empty((B * nt, neq))
        rc = _ivp_c(handle, B, neq, t_span.ctypes.data, 2, y0.ctypes.data, neq, data.ctypes.data, p * 8, p,
                    nt, t_eval.ctypes.data, n_events, first_step, max_step, rtol, atol, 0, 0, 0,
                    y.ctypes.data, result.ctypes.data, t_event.ctypes.data, y_event.ctypes.data, device)
    else:
        rc = _ivp_c(handle, B, neq, t_span.ctypes.data, 2, y0.ctypes.data, neq, data.ctypes.data, p * 8, p,
                    0, 0, n_events, first_step, max_step, rtol, atol, 0, 0, 0, 0, result.ctypes.data,
                    t_event.ctypes.data, y_event.ctypes.data, device)
        for b in range(B):
            row_offset[b + 1] = row_offset[b] + result[b, 4]
        R = row_offset[B]
        t = np.empty(R)
        y = np.empty((R, neq))
        if rc == 0:
            rc = _ivp_c(handle, B, neq, t_span.ctypes.data, 2, y0.ctypes.data, neq, data.ctypes.data, p * 8,
                        p, 0, 0, n_events, first_step, max_step, rtol, atol, 0, row_offset.ctypes.data,
                        t.ctypes.data, y.ctypes.data, result.ctypes.data, t_event.ctypes.data,
                        y_event.ctypes.data, device)
    if rc != 0:
        result[:, 0] = 0
    return t, y, row_offset, result, t_event, y_event
