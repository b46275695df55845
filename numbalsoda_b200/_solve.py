"""Shared host logic of lsoda() / dop853(): argument normalisation, batching by
array rank, the call into libb200ode, sharding over the GPUs of one box."""
import ctypes as ct
import threading

import numpy as np

from . import _lib, rhs as _rhs


def _as_f64(a, name):
    a = np.asarray(a)
    if a.dtype != np.float64:
        # the reference only passes the buffer address (driver.py:37): a non-float64
        # array is silently reinterpreted there; here it is an error
        raise TypeError("%s must be float64 (got %s)" % (name, a.dtype))
    return np.ascontiguousarray(a)


def _data_layout(data, B_hint):
    """-> (buffer, batched, record_bytes, np_staged)"""
    scalar_record = np.ndim(data) == 0  # (np.ascontiguousarray returns at least 1-D)
    data = np.ascontiguousarray(data)
    if data.dtype == np.float64:
        if data.ndim == 2:
            p = data.shape[1]
            return data, True, p * 8, (p if p <= 16 else -1)
        if data.ndim == 1:
            p = data.shape[0]
            return data, False, p * 8, (p if p <= 16 else -1)
        raise ValueError("data must be [p] or [B,p]")
    # record arrays (passing_data_to_rhs_function.ipynb): opaque bytes, passed by address
    if scalar_record or data.ndim == 0:
        return data.reshape(1), False, data.dtype.itemsize, -1
    if data.ndim == 1 and data.dtype.fields is not None and (B_hint is None or data.shape[0] == B_hint):
        return data, True, data.dtype.itemsize, -1  # one record per trajectory
    if data.ndim == 1 and B_hint is not None and data.shape[0] == B_hint:
        return data, True, data.dtype.itemsize, -1
    return data, False, data.nbytes, -1


def solve(solver, funcptr, u0, t_eval, data, rtol, atol, mxstep, exit_on_warning=False,
          strict=False, device=None, counters=False, out=None, exact_ctrl=False, ml=None, mu=None,
          pointers=None, rhs_lanes=None):
    u0 = _as_f64(u0, "u0")
    t_eval = _as_f64(t_eval, "t_eval")
    if data is None:
        data = np.array([0.0], np.float64)
    if u0.ndim not in (1, 2) or t_eval.ndim != 1:
        raise ValueError("u0 must be [n] or [B,n]; t_eval must be [nt]")
    batched_u0 = u0.ndim == 2
    data_arr = np.asarray(data)
    B_hint = u0.shape[0] if batched_u0 else (data_arr.shape[0] if data_arr.ndim == 2 else None)
    if not batched_u0 and data_arr.ndim == 1 and data_arr.dtype.fields is not None and np.ndim(data) == 1:
        B_hint = data_arr.shape[0]  # a 1-D array of records: one per trajectory
    dbuf, batched_data, rec, nps = _data_layout(data, B_hint)
    if batched_u0 and batched_data and dbuf.shape[0] != u0.shape[0]:
        raise ValueError("u0 has %d trajectories but data has %d" % (u0.shape[0], dbuf.shape[0]))
    batched = batched_u0 or batched_data
    B = u0.shape[0] if batched_u0 else (dbuf.shape[0] if batched_data else 1)
    n = u0.shape[-1]
    nt = t_eval.shape[0]
    # a record `data` (structured dtype): the right-hand side receives the record itself
    record = dbuf.dtype.fields is not None
    handle = _rhs.get_handle(funcptr, solver, n, strict=strict, exact_ctrl=exact_ctrl,
                             data_dtype=dbuf.dtype if record else None, rhs_lanes=rhs_lanes)
    if pointers and not record:
        raise ValueError("pointers= needs a record (structured dtype) `data`")
    if out is not None:
        # caller-owned result buffer, ideally page-locked (numbalsoda_b200.pinned_empty) and
        # reused across calls: allocating and first touching a fresh multi-gigabyte array
        # costs more than integrating the ensemble
        if out.dtype != np.float64 or not out.flags.c_contiguous or out.size != B * nt * n:
            raise ValueError("out must be a C-contiguous float64 array of %d elements" % (B * nt * n))
        usol = out.reshape(B, nt, n)
    else:
        usol = np.empty((B, nt, n), dtype=np.float64)  # as the reference, driver.py:35
    ok = np.full(B, 999, dtype=np.int32)
    cnt = np.zeros((B, _lib.NCOUNTERS), dtype=np.int32) if counters else None
    # rtol / atol: the reference's scalars, or arrays (tolerance_layout): one value per
    # trajectory [B], per component [n] / [1,n], or both [B,n]
    rtol, rtol_b, rtol_dims = tolerance_layout(rtol, B, n, "rtol", batched)
    atol, atol_b, atol_dims = tolerance_layout(atol, B, n, "atol", batched)

    # lsoda: banded finite-difference Jacobian (jt = 5) when ml and mu are given
    if (ml is None) != (mu is None):
        raise ValueError("ml and mu must be given together")
    if ml is not None:
        if solver != _lib.LSODA:
            raise ValueError("ml / mu are lsoda options")
        ml, mu = int(ml), int(mu)
        if not (0 <= ml < n and 0 <= mu < n):  # the reference's checker, LSODA.cpp:376-391
            raise ValueError("ml = %d, mu = %d not between 0 and neq - 1 = %d" % (ml, mu, n - 1))

    devices = _device_list(device)

    def device_records(dev):
        """`data` with the fields named in pointers= holding the device addresses of copies of
        their arrays on GPU `dev` (the reference stores host addresses there:
        passing_data_to_rhs_function.ipynb); the copies live until the call returns."""
        if not pointers:
            return dbuf, []
        patched = dbuf.copy()
        keep = []
        for field, arr in pointers.items():
            buf = _lib.DeviceBuffer(arr, dev)
            keep.append(buf)
            patched[field] = buf.ptr
        return patched, keep

    def run(dev, b0, b1):
        nb = b1 - b0
        if nb <= 0:
            return
        u0p = u0[b0:b1] if batched_u0 else u0
        dsrc, keep = device_records(dev)
        dp = dsrc[b0:b1] if batched_data else dsrc
        pb = _lib.Problem()
        pb.size = ct.sizeof(_lib.Problem)
        pb.B, pb.neq, pb.nt, pb.np = nb, n, nt, nps
        pb.u0, pb.u0_stride = u0p.ctypes.data, (n if batched_u0 else 0)
        pb.data, pb.data_stride = dp.ctypes.data, (rec if batched_data else -rec)
        pb.teval, pb.usol = t_eval.ctypes.data, usol[b0:b1].ctypes.data
        pb.rtol, pb.atol = float(rtol), float(atol)
        pb.rtol_b, pb.rtol_dims = _tol_slice(rtol_b, rtol_dims, b0, b1)
        pb.atol_b, pb.atol_dims = _tol_slice(atol_b, atol_dims, b0, b1)
        pb.mxstep, pb.exit_on_warning = int(mxstep), int(bool(exit_on_warning))
        if ml is not None:
            pb.jt, pb.ml, pb.mu = 5, ml, mu
        pb.success = ok[b0:b1].ctypes.data
        pb.counters = cnt[b0:b1].ctypes.data if counters else None
        _lib.check(_lib.lib.b200ode_solve(handle.ptr, ct.byref(pb), dev))

    if len(devices) == 1 or B < 2 * len(devices):
        run(devices[0], 0, B)
    else:
        # trajectories are independent: contiguous shards, one host thread per GPU,
        # no inter-GPU communication
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
    success = ok == 1
    if not batched:
        out = (usol[0], bool(success[0]))
        return out + ((cnt[0],) if counters else ())
    return (usol, success) + ((cnt,) if counters else ())


def tolerance_layout(x, B, n, name, batched=True):
    """rtol / atol as the drivers accept them -> (scalar, array or None, B200ODE_TOL_* shape code).

    scalar          the reference's argument (numbalsoda/driver.py:29)
    [B]             one value per trajectory
    [n] or [1, n]   one value per component, shared by the ensemble: the reference's solvers
                    support it (LSODA itol_ = 3 / 4, LSODA.cpp:1368-1390; dop853 itol = 1,
                    dop853_module.f90:603-612) although its drivers only pass scalars
    [B, n]          both
    A 1-D array whose length fits both readings (B == n) is taken per trajectory when the call
    is batched; write [1, n] for per component."""
    if np.ndim(x) == 0:
        return float(x), None, _lib.TOL_PER_TRAJECTORY
    a = np.ascontiguousarray(x, dtype=np.float64)
    if a.ndim == 1 and a.shape[0] == B and (batched or n == 1):
        return 0.0, a, _lib.TOL_PER_TRAJECTORY
    if a.ndim == 1 and a.shape[0] == n:
        return 0.0, a, _lib.TOL_PER_COMPONENT
    if a.ndim == 2 and a.shape == (1, n):
        return 0.0, np.ascontiguousarray(a[0]), _lib.TOL_PER_COMPONENT
    if a.ndim == 2 and a.shape == (B, n):
        return 0.0, a, _lib.TOL_FULL
    raise ValueError("%s must be a scalar or have shape (%d,) [per trajectory], (%d,) / (1, %d) "
                     "[per component] or (%d, %d); got %r" % (name, B, n, n, B, n, a.shape))


def _tol_slice(arr, dims, b0, b1):
    """the part of a tolerance array that belongs to trajectories [b0, b1)"""
    if arr is None:
        return None, 0
    if dims == _lib.TOL_PER_COMPONENT:
        return arr.ctypes.data, dims
    return arr[b0:b1].ctypes.data, dims


def shard_bounds(B, parts):
    """Contiguous, balanced shard boundaries: part i owns [bounds[i], bounds[i+1])."""
    base, extra = divmod(B, parts)
    bounds = [0]
    for i in range(parts):
        bounds.append(bounds[-1] + base + (1 if i < extra else 0))
    return bounds


def _device_list(device):
    if device is None:
        return [0]
    if device == "all":
        n = _lib.lib.b200ode_device_count()
        if n < 1:
            raise _lib.B200OdeError("no CUDA device: numbalsoda_b200 has no CPU fallback")
        return list(range(n))
    if isinstance(device, (list, tuple)):
        return [int(d) for d in device]
    return [int(device)]
