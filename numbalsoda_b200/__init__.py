"""numbalsoda_b200: B200-native batched lsoda / dop853 / solve_ivp behind numbalsoda's API
(reference numbalsoda/__init__.py:1-3)."""
from .driver import lsoda_sig, lsoda, address_as_void_pointer, address_as_pointer
from .driver_dop853 import dop853
from .driver_solve_ivp import solve_ivp, ODEResult, ODEResultBatch
from .rhs import builtin, register, get_handle
from ._lib import B200OdeError, pinned_empty, DeviceBuffer

try:  # numba present: make lsoda / dop853 resolvable inside @njit code (reference driver.py:28)
    from . import _njit_overloads  # noqa: F401
except ImportError:  # pragma: no cover
    pass

__all__ = ["lsoda_sig", "lsoda", "dop853", "solve_ivp", "ODEResult", "ODEResultBatch", "address_as_void_pointer", "address_as_pointer", "DeviceBuffer", "builtin", "register",
           "get_handle", "B200OdeError", "pinned_empty"]
