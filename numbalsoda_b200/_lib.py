"""ctypes binding of libb200ode.so (include/b200ode.h).  There is no fallback: if
the library is missing it is built once with nvcc; if that fails, importing fails."""
import ctypes as ct
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, "libb200ode.so")

DOP853, LSODA = 0, 1
IMG_LTOIR, IMG_PTX, IMG_FATBIN, IMG_BUILTIN = 0, 1, 2, 3
STRICT_FP = 1
NCOUNTERS = 8


class B200OdeError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIBPATH):
        from . import build as _build
        _build.build()
    lib = ct.CDLL(LIBPATH)
    vp, ll, i, d, sz = ct.c_void_p, ct.c_longlong, ct.c_int, ct.c_double, ct.c_size_t
    lib.b200ode_last_error.restype = ct.c_char_p
    lib.b200ode_link_rhs.argtypes = [i, i, vp, sz, i, ct.c_uint, ct.POINTER(vp)]
    lib.b200ode_rhs_free.argtypes = [vp]
    lib.b200ode_rhs_free.restype = None
    lib.b200ode_rhs_cubin.argtypes = [vp, ct.POINTER(vp), ct.POINTER(sz)]
    common = [vp, ll, i, vp, ll, vp, ll, i, i, vp, vp, d, d, i, vp, vp]
    lib.b200ode_dop853_batched.argtypes = common + [i]
    lib.b200ode_lsoda_batched.argtypes = common + [i, i]
    lib.b200ode_dop853_batched_device.argtypes = common + [vp]
    lib.b200ode_lsoda_batched_device.argtypes = common + [i, vp]
    lib.b200ode_kernel_info.argtypes = [vp, ct.POINTER(i * 8)]
    return lib


lib = _load()


def check(rc):
    if rc != 0:
        raise B200OdeError("libb200ode error %d: %s" % (rc, lib.b200ode_last_error().decode()))
