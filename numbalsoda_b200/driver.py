"""lsoda(): GPU drop-in for numbalsoda.lsoda (reference numbalsoda/driver.py:28-43).

Same positional arguments and keyword defaults.  With `u0[n]` and `data[p]` it
returns `(usol[nt,n], success)` like the reference; a 2-D `u0[B,n]` and/or
`data[B,p]` selects the ensemble path and returns `(usol[B,nt,n], success[B])`.
"""
import numpy as np

from . import _lib
from ._solve import solve

try:
    from numba import types as _types
    # reference numbalsoda/driver.py:8-11
    lsoda_sig = _types.void(_types.double, _types.CPointer(_types.double),
                            _types.CPointer(_types.double), _types.CPointer(_types.double))
except Exception:  # numba absent: built-in right-hand sides still work
    lsoda_sig = None


def lsoda(funcptr, u0, t_eval, data=np.array([0.0], np.float64), rtol=1.0e-3, atol=1.0e-6,
          mxstep=10000, exit_on_warning=False, *, strict=False, device=None, counters=False, out=None,
          exact_ctrl=False, ml=None, mu=None):
    """Integrate with the LSODA kernel on the GPU.

    funcptr: address of a numba @cfunc(lsoda_sig) (as in the reference), the cfunc, the
    Python function, or numbalsoda_b200.builtin(name).
    Extra keyword-only arguments: strict (FMA contraction off: bit-for-bit the CPU
    arithmetic), device (index, list of indices or "all"), counters (also return
    [nst, nfe, nje, nswitch, nqu, istate, rows, meth] per trajectory), out (a preallocated
    C-contiguous float64 result buffer of B*nt*n elements -- e.g. from
    numbalsoda_b200.pinned_empty -- to reuse across calls instead of a fresh np.empty), ml / mu
    (lower / upper bandwidth of the Jacobian: the iteration matrix is then formed from
    ml + mu + 1 right-hand-side evaluations instead of n and factored as a band matrix -- the
    jt = 5 option the reference's solver checks for, LSODA.cpp:369-392, but does not implement;
    systems of up to 8 equations keep the dense matrix).  rtol / atol may also be arrays: [B] per
    trajectory, [n] per component, [B,n] both."""
    return solve(_lib.LSODA, funcptr, u0, t_eval, data, rtol, atol, mxstep,
                 exit_on_warning=exit_on_warning, strict=strict, device=device,
                 counters=counters, out=out, exact_ctrl=exact_ctrl, ml=ml, mu=mu)


def address_as_void_pointer(addr):
    """The reference offers this numba intrinsic (driver.py:45-53) to hand a struct
    with embedded host pointers to the RHS.  Host pointers cannot be dereferenced on
    the device; pass flat float64 data or a plain record array instead."""
    raise NotImplementedError(
        "address_as_void_pointer: data records with embedded host pointers are not "
        "supported on the GPU; pass float64[p] / float64[B,p] or a numpy record array")
