"""`lsoda` / `dop853` with the reference's signature, callable from inside @njit code.

The reference's drivers are themselves @njit functions (numbalsoda/driver.py:28,
driver_dop853.py:23), so user code calls them from nopython code (README.md:60-67):

    @njit
    def test():
        usol, success = lsoda(funcptr, u0, t_eval, data=data)
        return usol

Here `lsoda` is a Python function (it has to find the Python source behind `funcptr`, compile it
for the device and link it once), so numba needs to be told what a call to it means in nopython
code: an `@overload` whose implementation steps into object mode for the call.  The first call
links the right-hand side (cached per function, solver, size and mode: numbalsoda_b200/rhs.py);
later calls cost one interpreter round trip on top of the GPU work.  A 1-D `u0` with 1-D `data`
returns `(usol[nt,n], success)` exactly like the reference; 2-D `u0[B,n]` and / or `data[B,p]`
return the ensemble `(usol[B,nt,n], success[B])`.

`njit_api.lsoda_batched(handle, ...)` remains the lower-overhead path for nopython code that
can link ahead (no interpreter involved: the C ABI through a ctypes function object)."""
import numpy as np
from numba import objmode, types
from numba.extending import overload

from .driver import lsoda
from .driver_dop853 import dop853

_DEFAULT_DATA = np.array([0.0], np.float64)


def _ndim(t):
    return t.ndim if isinstance(t, types.Array) else 0


def _typecheck(funcptr, u0, t_eval, data):
    """-> None if the call is not the reference's, else (ensemble call?, data omitted?)"""
    if not isinstance(u0, types.Array) or not isinstance(t_eval, types.Array):
        return None
    if not isinstance(funcptr, (types.Integer, types.IntegerLiteral)):
        return None  # the reference's contract: funcptr = cfunc.address
    no_data = data is None or isinstance(data, (types.Omitted, types.NoneType))
    return u0.ndim == 2 or (not no_data and _ndim(data) == 2), no_data


def _py_lsoda(funcptr, u0, t_eval, data, rtol, atol, mxstep, exit_on_warning):
    usol, ok = lsoda(funcptr, u0, t_eval, _DEFAULT_DATA if data is None else data, rtol, atol, mxstep,
                     exit_on_warning)
    return np.ascontiguousarray(usol), ok


def _py_dop853(funcptr, u0, t_eval, data, rtol, atol, mxstep):
    usol, ok = dop853(funcptr, u0, t_eval, _DEFAULT_DATA if data is None else data, rtol, atol, mxstep)
    return np.ascontiguousarray(usol), ok


@overload(lsoda)
def _ol_lsoda(funcptr, u0, t_eval, data=None, rtol=1.0e-3, atol=1.0e-6, mxstep=10000,
              exit_on_warning=False):
    tc = _typecheck(funcptr, u0, t_eval, data)
    if tc is None:
        return None
    if tc[0]:
        def impl(funcptr, u0, t_eval, data=None, rtol=1.0e-3, atol=1.0e-6, mxstep=10000,
                 exit_on_warning=False):
            with objmode(usol="float64[:,:,::1]", ok="boolean[::1]"):
                usol, ok = _py_lsoda(funcptr, u0, t_eval, data, rtol, atol, mxstep, exit_on_warning)
            return usol, ok
    else:
        def impl(funcptr, u0, t_eval, data=None, rtol=1.0e-3, atol=1.0e-6, mxstep=10000,
                 exit_on_warning=False):
            with objmode(usol="float64[:,::1]", ok="boolean"):
                usol, ok = _py_lsoda(funcptr, u0, t_eval, data, rtol, atol, mxstep, exit_on_warning)
            return usol, ok
    return impl


@overload(dop853)
def _ol_dop853(funcptr, u0, t_eval, data=None, rtol=1.0e-3, atol=1.0e-6, mxstep=10000):
    tc = _typecheck(funcptr, u0, t_eval, data)
    if tc is None:
        return None
    if tc[0]:
        def impl(funcptr, u0, t_eval, data=None, rtol=1.0e-3, atol=1.0e-6, mxstep=10000):
            with objmode(usol="float64[:,:,::1]", ok="boolean[::1]"):
                usol, ok = _py_dop853(funcptr, u0, t_eval, data, rtol, atol, mxstep)
            return usol, ok
    else:
        def impl(funcptr, u0, t_eval, data=None, rtol=1.0e-3, atol=1.0e-6, mxstep=10000):
            with objmode(usol="float64[:,::1]", ok="boolean"):
                usol, ok = _py_dop853(funcptr, u0, t_eval, data, rtol, atol, mxstep)
            return usol, ok
    return impl
