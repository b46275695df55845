"""solve_ivp(): GPU drop-in for numbalsoda.solve_ivp (reference
numbalsoda/driver_solve_ivp.py:86-218; native side src/dop853_solve_ivp.f90:211-434).

Same positional arguments and keyword defaults.  With `y0[n]`, `t_span[2]` and `data[p]`
it returns one ODEResult with the reference's fields (driver_solve_ivp.py:59-83); a 2-D
`y0[B,n]`, `t_span[B,2]` and/or `data[B,p]` selects the ensemble path and returns an
ODEResultBatch: the same fields as arrays over the ensemble, `batch[b]` being the ODEResult
of trajectory b.

Without `t_eval` every accepted step is an output row, so trajectories have different row
counts.  By default the ensemble is integrated twice -- a first pass that only counts rows, a
second one that writes them into exactly sized (ragged) storage; `max_rows=R` integrates
once into `R` rows per trajectory instead (a trajectory that needs more ends with
success = False, message 'output capacity exceeded').
"""
import ctypes as ct

import numpy as np

from . import _lib, rhs as _rhs
from ._solve import _as_f64, _data_layout, _device_list, shard_bounds


class ODEResult:
    """Fields of the reference's jitclass (numbalsoda/driver_solve_ivp.py:59-83)."""
    __slots__ = ("message", "nfev", "success", "t", "y", "event_found", "ind_event", "t_event",
                 "y_event", "counters")

    def __init__(self, message, nfev, success, t, y, event_found, ind_event, t_event, y_event,
                 counters=None):
        self.message, self.nfev, self.success, self.t, self.y = message, nfev, success, t, y
        self.event_found, self.ind_event, self.t_event, self.y_event = (event_found, ind_event,
                                                                       t_event, y_event)
        self.counters = counters  # dict(idid, nstep, naccpt, nrejct): the solver's private counters

    def __repr__(self):
        return ("ODEResult(message=%r, nfev=%d, success=%r, t=<%d>, y=<%s>, event_found=%r, "
                "ind_event=%d, t_event=%r)" % (self.message, self.nfev, self.success, len(self.t),
                                               "x".join(map(str, self.y.shape)), self.event_found,
                                               self.ind_event, self.t_event))


class ODEResultBatch:
    """Results of an ensemble.  Arrays over trajectories: success[B], nfev[B], message_code[B],
    nt_out[B] (valid rows), event_found[B], ind_event[B], t_event[B], y_event[B,n], and the
    counters idid / nstep / naccpt / nrejct / rows_needed [B].  Rows: with t_eval, y[B,nt,n]
    (rows >= nt_out[b] are not meaningful, as the reference trims them); without,
    `row_offset[B+1]` delimits trajectory b's rows in t[R], y[R,n]."""

    def __init__(self, result, t_eval, t, y, row_offset, t_event, y_event):
        self._r = result
        self.t_eval, self.t, self.y, self.row_offset = t_eval, t, y, row_offset
        self.success = result[:, 0] == 1
        self.message_code = result[:, 1]
        self.nfev = result[:, 2].astype(np.int64)
        self.nt_out = result[:, 3]
        self.rows_needed = result[:, 4]
        self.event_found = result[:, 5] == 1
        self.ind_event = result[:, 6].astype(np.int64)
        self.idid, self.nstep, self.naccpt, self.nrejct = (result[:, k] for k in (7, 8, 9, 10))
        self.t_event, self.y_event = t_event, y_event

    def __len__(self):
        return self._r.shape[0]

    def message(self, b):
        return _lib.lib.b200ode_ivp_message(int(self.message_code[b])).decode()

    def __getitem__(self, b):
        b = int(b)
        if b < 0:
            b += len(self)
        k = int(self.nt_out[b])
        if self.t_eval is not None:
            t, y = self.t_eval[:k].copy(), self.y[b, :k]
        else:
            r0 = int(self.row_offset[b])
            t, y = self.t[r0:r0 + k], self.y[r0:r0 + k]
        cnt = dict(idid=int(self.idid[b]), nstep=int(self.nstep[b]), naccpt=int(self.naccpt[b]),
                   nrejct=int(self.nrejct[b]), rows_needed=int(self.rows_needed[b]))
        return ODEResult(self.message(b), int(self.nfev[b]), bool(self.success[b]), t, y,
                         bool(self.event_found[b]), int(self.ind_event[b]), float(self.t_event[b]),
                         self.y_event[b], cnt)

    def __iter__(self):
        return (self[b] for b in range(len(self)))


def _run(handle, devices, B, n, ts, y0, dbuf, batched_data, rec, nps, te, n_events, first_step,
         max_step, rtol, atol, rtol_b, atol_b, cap, row_offset, t_out, y_out, result, t_event,
         y_event):
    """One pass over the ensemble (sharded by trajectory over `devices`, no collective)."""
    nt = 0 if te is None else te.shape[0]

    def rows_before(b):
        # (ragged rows are addressed through row_offset itself: the arrays are not shifted)
        return b * nt if nt else (0 if row_offset is not None else b * cap)

    def run(dev, b0, b1):
        if b1 <= b0:
            return
        pb = _lib.IvpProblem()
        pb.size = ct.sizeof(_lib.IvpProblem)
        pb.B, pb.neq, pb.nt, pb.np, pb.n_events = b1 - b0, n, nt, nps, n_events
        pb.t_span = (ts[b0:b1] if ts.ndim == 2 else ts).ctypes.data
        pb.t_span_stride = 2 if ts.ndim == 2 else 0
        pb.y0 = (y0[b0:b1] if y0.ndim == 2 else y0).ctypes.data
        pb.y0_stride = n if y0.ndim == 2 else 0
        pb.data = (dbuf[b0:b1] if batched_data else dbuf).ctypes.data
        pb.data_stride = rec if batched_data else -rec
        pb.teval = te.ctypes.data if nt else None
        pb.first_step, pb.max_step, pb.rtol, pb.atol = first_step, max_step, rtol, atol
        pb.rtol_b = rtol_b[b0:b1].ctypes.data if rtol_b is not None else None
        pb.atol_b = atol_b[b0:b1].ctypes.data if atol_b is not None else None
        pb.cap = cap
        pb.row_offset = row_offset[b0:].ctypes.data if row_offset is not None else None
        r0 = rows_before(b0)
        if y_out is not None and y_out.size:
            pb.y_out = y_out.ctypes.data + r0 * n * 8
        if t_out is not None and t_out.size:
            pb.t_out = t_out.ctypes.data + r0 * 8
        pb.result = result[b0:b1].ctypes.data
        pb.t_event = t_event[b0:b1].ctypes.data
        pb.y_event = y_event[b0:b1].ctypes.data
        _lib.check(_lib.lib.b200ode_solve_ivp(handle.ptr, ct.byref(pb), dev))

    if len(devices) == 1 or B < 2 * len(devices):
        run(devices[0], 0, B)
        return
    import threading
    bounds = shard_bounds(B, len(devices))
    errs = []

    def work(i):
        try:
            run(devices[i], bounds[i], bounds[i + 1])
        except Exception as e:  # noqa: BLE001
            errs.append(e)
    th = [threading.Thread(target=work, args=(i,)) for i in range(len(devices))]
    [t.start() for t in th]
    [t.join() for t in th]
    if errs:
        raise errs[0]


def solve_ivp(funcptr, t_span, y0, t_eval=np.array([], np.double), method="DOP853", n_events=0,
              event_fcn=0, data=0, first_step=0.0, max_step=0.0, rtol=1.0e-3, atol=1.0e-6, *,
              strict=False, exact_ctrl=False, device=None, max_rows=None, out=None):
    """Solve initial value problems with the DOP853 solve_ivp kernel on the GPU.

    Arguments as the reference's solve_ivp (driver_solve_ivp.py:86-127).  funcptr / event_fcn:
    address of a numba @cfunc(lsoda_sig), the cfunc, the Python function `f(t, y, out, p)` or
    numbalsoda_b200.builtin(name).  data: a float64 array [p] or [B,p] or a record array (the
    reference takes a raw address, which cannot be followed from the device), 0 / None = none.
    Keyword-only: strict (FMA contraction off: bit-for-bit the CPU arithmetic), exact_ctrl (FMA
    contraction on, but the reference's controller arithmetic -- IEEE divisions, libm-exact pow --
    instead of the default mode's reciprocals and single-precision err**(1/8)), device (index,
    list or "all"), max_rows (see the module docstring), out (with t_eval: a preallocated
    C-contiguous float64 buffer of B*nt*n elements for the rows, e.g. from
    numbalsoda_b200.pinned_empty, to reuse across calls); rtol / atol may have shape [B]."""
    t_span = _as_f64(t_span, "t_span")
    y0 = _as_f64(y0, "y0")
    if t_span.shape[-1:] != (2,) or t_span.ndim not in (1, 2):
        raise ValueError('"t_span" must be length 2')  # driver_solve_ivp.py:154-155
    if method not in ["DOP853"]:
        raise ValueError('Invalid "method".')  # :157-158
    if y0.ndim not in (1, 2):
        raise ValueError("y0 must be [n] or [B,n]")
    te = None
    if t_eval is not None and np.size(t_eval) > 0:
        te = _as_f64(t_eval, "t_eval")
        if te.ndim != 1:
            raise ValueError("t_eval must be [nt]")
    if data is None or (np.ndim(data) == 0 and not isinstance(data, np.ndarray) and data == 0):
        data = np.array([0.0], np.float64)
    elif isinstance(data, (int, np.integer)):
        raise TypeError("data must be an array: a raw host address (the reference's uintp) cannot "
                        "be dereferenced on the device")
    n = y0.shape[-1]
    data_arr = np.asarray(data)
    sizes = [a.shape[0] for a in (y0, t_span) if a.ndim == 2]
    if data_arr.ndim == 2:
        sizes.append(data_arr.shape[0])
    B_hint = sizes[0] if sizes else None
    dbuf, batched_data, rec, nps = _data_layout(data, B_hint)
    if batched_data and not sizes:
        sizes.append(dbuf.shape[0])
    if len(set(sizes)) > 1:
        raise ValueError("y0, t_span and data disagree on the number of trajectories: %r" % sizes)
    batched = bool(sizes)
    B = sizes[0] if sizes else 1
    n_events = int(n_events)
    handle = _rhs.get_ivp_handle(funcptr, n, n_events, event_fcn, strict=strict, exact_ctrl=exact_ctrl)
    rtol_b = atol_b = None
    if np.ndim(rtol) > 0:
        rtol_b, rtol = np.ascontiguousarray(rtol, dtype=np.float64), 0.0
    if np.ndim(atol) > 0:
        atol_b, atol = np.ascontiguousarray(atol, dtype=np.float64), 0.0
    for name, arr in (("rtol", rtol_b), ("atol", atol_b)):
        if arr is not None and arr.shape != (B,):
            raise ValueError("%s must be a scalar or have shape (%d,)" % (name, B))
    devices = _device_list(device)

    result = np.zeros((B, _lib.IVP_NRESULT), np.int32)
    t_event = np.zeros(B)
    y_event = np.zeros((B, n))
    common = (handle, devices, B, n, t_span, y0, dbuf, batched_data, rec, nps, te, n_events,
              float(first_step), float(max_step), float(rtol), float(atol), rtol_b, atol_b)
    row_offset = t_out = None
    if te is not None:
        if out is not None:
            if out.dtype != np.float64 or not out.flags.c_contiguous or out.size != B * te.shape[0] * n:
                raise ValueError("out must be a C-contiguous float64 array of %d elements"
                                 % (B * te.shape[0] * n))
            y_out = out.reshape(B, te.shape[0], n)
        else:
            y_out = np.empty((B, te.shape[0], n))
        _run(*common, 0, None, None, y_out, result, t_event, y_event)
    elif max_rows is not None:
        cap = int(max_rows)
        t_out, y_out = np.empty(B * cap), np.empty((B * cap, n))
        row_offset = np.arange(B + 1, dtype=np.int64) * cap
        _run(*common, cap, None, t_out, y_out, result, t_event, y_event)
    else:
        # pass 1 counts the rows of every trajectory (nothing is stored), pass 2 writes them
        _run(*common, 0, None, None, None, result, t_event, y_event)
        row_offset = np.zeros(B + 1, np.int64)
        np.cumsum(result[:, 4], out=row_offset[1:])
        R = int(row_offset[-1])
        t_out, y_out = np.empty(R), np.empty((R, n))
        _run(*common, 0, row_offset, t_out, y_out, result, t_event, y_event)
    batch = ODEResultBatch(result, te, t_out, y_out, row_offset, t_event, y_event)
    return batch if batched else batch[0]
