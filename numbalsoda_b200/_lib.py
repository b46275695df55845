"""ctypes binding of libb200ode.so (include/b200ode.h).  There is no fallback: if
the library is missing it is built once with nvcc; if that fails, importing fails."""
import ctypes as ct
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# B200ODE_LIB: load another build of the library (tools/sweep_ivp.sh times kernel variants
# built ahead with different launch bounds)
LIBPATH = os.environ.get("B200ODE_LIB") or os.path.join(HERE, "libb200ode.so")

DOP853, LSODA, SOLVE_IVP = 0, 1, 2
IMG_LTOIR, IMG_PTX, IMG_FATBIN, IMG_BUILTIN = 0, 1, 2, 3
STRICT_FP = 1
EXACT_CTRL = 2
NCOUNTERS = 8
TOL_PER_TRAJECTORY, TOL_PER_COMPONENT, TOL_FULL = 0, 1, 2  # shape of rtol_b / atol_b
IVP_NRESULT = 12
IVP_NEVENTS_MAX = 8


class B200OdeError(RuntimeError):
    pass


class Problem(ct.Structure):
    """b200ode_problem (include/b200ode.h)."""
    _fields_ = [("size", ct.c_size_t), ("B", ct.c_longlong), ("neq", ct.c_int),
                ("u0", ct.c_void_p), ("u0_stride", ct.c_longlong), ("data", ct.c_void_p),
                ("data_stride", ct.c_longlong), ("np", ct.c_int), ("nt", ct.c_int),
                ("teval", ct.c_void_p), ("usol", ct.c_void_p), ("rtol", ct.c_double),
                ("atol", ct.c_double), ("rtol_b", ct.c_void_p), ("atol_b", ct.c_void_p),
                ("mxstep", ct.c_int), ("exit_on_warning", ct.c_int), ("success", ct.c_void_p),
                ("counters", ct.c_void_p), ("rtol_dims", ct.c_int), ("atol_dims", ct.c_int),
                ("jt", ct.c_int), ("ml", ct.c_int), ("mu", ct.c_int)]


class IvpProblem(ct.Structure):
    """b200ode_ivp_problem (include/b200ode.h)."""
    _fields_ = [("size", ct.c_size_t), ("B", ct.c_longlong), ("neq", ct.c_int),
                ("t_span", ct.c_void_p), ("t_span_stride", ct.c_longlong), ("y0", ct.c_void_p),
                ("y0_stride", ct.c_longlong), ("data", ct.c_void_p), ("data_stride", ct.c_longlong),
                ("np", ct.c_int), ("nt", ct.c_int), ("teval", ct.c_void_p), ("n_events", ct.c_int),
                ("first_step", ct.c_double), ("max_step", ct.c_double), ("rtol", ct.c_double),
                ("atol", ct.c_double), ("rtol_b", ct.c_void_p), ("atol_b", ct.c_void_p),
                ("cap", ct.c_longlong), ("row_offset", ct.c_void_p), ("t_out", ct.c_void_p),
                ("y_out", ct.c_void_p), ("result", ct.c_void_p), ("t_event", ct.c_void_p),
                ("y_event", ct.c_void_p)]


def _load(rebuilt=False):
    if not os.path.exists(LIBPATH):
        from . import build as _build
        _build.build()
        rebuilt = True
    lib = ct.CDLL(LIBPATH)
    if not all(hasattr(lib, f) for f in ("b200ode_solve_ivp_args", "b200ode_link_rhs_lanes",
                                         "b200ode_device_alloc")) and not rebuilt:
        # a library built from older sources: rebuild it once
        from . import build as _build
        _build.build(force=True)
        return _load(rebuilt=True)
    vp, ll, i, d, sz = ct.c_void_p, ct.c_longlong, ct.c_int, ct.c_double, ct.c_size_t
    lib.b200ode_last_error.restype = ct.c_char_p
    lib.b200ode_link_rhs.argtypes = [i, i, vp, sz, i, ct.c_uint, ct.POINTER(vp)]
    lib.b200ode_link_rhs_lanes.argtypes = [i, i, vp, sz, i, vp, sz, i, ct.c_uint, ct.POINTER(vp)]
    lib.b200ode_rhs_free.argtypes = [vp]
    lib.b200ode_rhs_free.restype = None
    lib.b200ode_rhs_cubin.argtypes = [vp, ct.POINTER(vp), ct.POINTER(sz)]
    common = [vp, ll, i, vp, ll, vp, ll, i, i, vp, vp, d, d, i, vp, vp]
    lib.b200ode_dop853_batched.argtypes = common + [i]
    lib.b200ode_lsoda_batched.argtypes = common + [i, i]
    lib.b200ode_dop853_batched_device.argtypes = common + [vp]
    lib.b200ode_lsoda_batched_device.argtypes = common + [i, vp]
    lib.b200ode_kernel_info.argtypes = [vp, ct.POINTER(i * 8)]
    lib.b200ode_solve.argtypes = [vp, ct.POINTER(Problem), i]
    lib.b200ode_solve_device.argtypes = [vp, ct.POINTER(Problem), vp]
    lib.b200ode_link_ivp.argtypes = [i, vp, sz, i, i, vp, sz, i, ct.c_uint, ct.POINTER(vp)]
    lib.b200ode_solve_ivp.argtypes = [vp, ct.POINTER(IvpProblem), i]
    lib.b200ode_solve_ivp_device.argtypes = [vp, ct.POINTER(IvpProblem), vp]
    lib.b200ode_ivp_message.argtypes = [i]
    lib.b200ode_ivp_message.restype = ct.c_char_p
    lib.b200ode_host_alloc.argtypes = [sz, ct.POINTER(vp)]
    lib.b200ode_host_free.argtypes = [vp]
    lib.b200ode_device_alloc.argtypes = [i, sz, ct.POINTER(vp)]
    lib.b200ode_device_upload.argtypes = [i, vp, vp, sz]
    lib.b200ode_device_free.argtypes = [i, vp]
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        raise B200OdeError("libb200ode error %d: %s" % (rc, lib.b200ode_last_error().decode()))


class DeviceBuffer:
    """A copy of a host array on one GPU: `.ptr` is its device address (an int), to be stored in
    a `data` record in the place of the host address the reference stores there
    (passing_data_to_rhs_function.ipynb: `arr.ctypes.data`).  Freed with the object."""

    def __init__(self, array, device=0):
        import numpy as np
        a = np.ascontiguousarray(array)
        self.device, self.nbytes, self.shape, self.dtype = int(device), a.nbytes, a.shape, a.dtype
        p = ct.c_void_p()
        check(lib.b200ode_device_alloc(self.device, a.nbytes, ct.byref(p)))
        self._p = p
        check(lib.b200ode_device_upload(self.device, p, a.ctypes.data, a.nbytes))

    @property
    def ptr(self):
        return int(self._p.value)

    def __del__(self):
        try:
            if self._p:
                lib.b200ode_device_free(self.device, self._p)
                self._p = None
        except Exception:
            pass


class _PinnedOwner:
    def __init__(self, ptr):
        self.ptr = ptr

    def __del__(self):
        try:
            lib.b200ode_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """np.empty in page-locked memory (falls back to pageable memory without a driver)."""
    import numpy as np
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = ct.c_void_p()
    if n == 0 or lib.b200ode_host_alloc(n, ct.byref(p)) != 0 or not p.value:
        return np.empty(shape, dtype)
    buf = (ct.c_char * n).from_address(p.value)
    buf._owner = _PinnedOwner(p)  # freed when the last array viewing the buffer goes away
    return np.frombuffer(buf, dtype=dtype).reshape(shape)
