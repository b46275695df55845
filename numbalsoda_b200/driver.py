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
          exact_ctrl=False, ml=None, mu=None, pointers=None, rhs_lanes=None):
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
    trajectory, [n] per component, [B,n] both.  pointers: {field: array} for a record `data` whose
    integer fields hold the addresses of further arrays (the reference's address_as_void_pointer
    pattern): each array is copied to the device and its device address stored in that field."""
    return solve(_lib.LSODA, funcptr, u0, t_eval, data, rtol, atol, mxstep,
                 exit_on_warning=exit_on_warning, strict=strict, device=device,
                 counters=counters, out=out, exact_ctrl=exact_ctrl, ml=ml, mu=mu, pointers=pointers, rhs_lanes=rhs_lanes)


try:
    from numba.extending import intrinsic as _intrinsic

    @_intrinsic(target="generic")
    def address_as_void_pointer(typingctx, src):
        """returns a void pointer from a given memory address -- the reference's intrinsic
        (numbalsoda/driver.py:45-53), registered for the CPU and the CUDA target alike"""
        from numba.core import cgutils
        sig = _types.voidptr(src)

        def codegen(cgctx, builder, sig, args):
            return builder.inttoptr(args[0], cgutils.voidptr_t)
        return sig, codegen

    @_intrinsic(target="generic")
    def address_as_pointer(typingctx, src, dtype):
        """address_as_pointer(addr, numba.float64) -> a typed pointer that can be indexed, on the
        CPU and on the device: what `nb.carray(address_as_void_pointer(addr), n, dtype=...)` is
        used for in the reference's notebook (numba.cuda has no carray)."""
        if isinstance(dtype, _types.NumberClass):
            ty = dtype.instance_type
        elif isinstance(dtype, _types.DType):
            ty = dtype.dtype
        else:
            return None
        sig = _types.CPointer(ty)(src, dtype)

        def codegen(cgctx, builder, sig, args):
            return builder.inttoptr(args[0], cgctx.get_value_type(sig.return_type))
        return sig, codegen
except Exception:  # numba absent
    def address_as_void_pointer(addr):
        raise NotImplementedError("address_as_void_pointer needs numba")

    def address_as_pointer(addr, dtype):
        raise NotImplementedError("address_as_pointer needs numba")
