"""Device pow == host libm pow, bit for bit, evaluated on the GPU."""
import ctypes as ct

import numpy as np
import pytest

from test_pow import pow_inputs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("k", range(2, 14))
def test_device_pow_bitwise_equal_to_host_libm(k):
    import torch
    from numbalsoda_b200 import _lib
    rng = np.random.default_rng(100 + k)
    x = np.ascontiguousarray(pow_inputs(rng, 1_000_000))
    xd = torch.tensor(x, device="cuda")
    od = torch.empty_like(xd)
    _lib.lib.b200ode_selftest_pow.argtypes = [ct.c_void_p, ct.c_double, ct.c_void_p,
                                              ct.c_longlong, ct.c_void_p]
    _lib.check(_lib.lib.b200ode_selftest_pow(xd.data_ptr(), 1.0 / k, od.data_ptr(), len(x), None))
    a = od.cpu().numpy()
    b = np.power(x, 1.0 / k)  # numpy calls the host libm pow for float64 scalars loops
    libm = ct.CDLL("libm.so.6")
    libm.pow.restype = ct.c_double
    libm.pow.argtypes = [ct.c_double, ct.c_double]
    # numpy may use its own SIMD pow: check a subsample against libm itself too
    idx = rng.integers(0, len(x), 20000)
    bl = np.array([libm.pow(float(v), 1.0 / k) for v in x[idx]])
    same = (a[idx].view(np.uint64) == bl.view(np.uint64)) | (np.isnan(a[idx]) & np.isnan(bl))
    assert same.all()
    close = np.isclose(a, b, rtol=1e-15, atol=0, equal_nan=True)
    assert close.all()
