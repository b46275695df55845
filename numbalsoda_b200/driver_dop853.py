"""dop853(): GPU drop-in for numbalsoda.dop853 (reference numbalsoda/driver_dop853.py:23-35)."""
import numpy as np

from . import _lib
from ._solve import solve


def dop853(funcptr, u0, t_eval, data=np.array([0.0], np.float64), rtol=1.0e-3, atol=1.0e-6,
           mxstep=10000, *, strict=False, device=None, counters=False, out=None,
           exact_ctrl=False, pointers=None, rhs_lanes=None):
    """Integrate with the DOP853 kernel on the GPU; arguments as lsoda().  counters:
    [nstep, nfcn, naccpt, nrejct, nevent, idid, rows, 0] per trajectory.  exact_ctrl: keep the
    reference's controller arithmetic (IEEE divisions, libm-exact pow) in the default FMA
    mode (include/b200ode.h: B200ODE_EXACT_CTRL); implied by strict."""
    return solve(_lib.DOP853, funcptr, u0, t_eval, data, rtol, atol, mxstep, strict=strict,
                 device=device, counters=counters, out=out, exact_ctrl=exact_ctrl, pointers=pointers, rhs_lanes=rhs_lanes)
