"""The device pow (numbalsoda_b200/csrc/b200ode_pow.cuh) compiled as host code must be
the host libm's pow bit-for-bit on the exponents and ranges the solvers use."""
import ctypes as ct
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib():
    os.makedirs(os.path.join(HERE, "_build"), exist_ok=True)
    so = os.path.join(HERE, "_build", "libpow_host.so")
    subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-fPIC", "-shared", "-o", so,
                           os.path.join(HERE, "pow_host.cpp"), "-lm"])
    return ct.CDLL(so)


def pow_inputs(rng, n):
    special = np.array([0.0, 1.0, np.inf, np.nan, 5e-324, 1e-310, 2.2250738585072014e-308,
                        1.7976931348623157e308, 1 + 2 ** -52, 1 - 2 ** -53, 0.5, 2.0, 4.0, 256.0])
    return np.concatenate([10.0 ** rng.uniform(-300, 300, n), rng.uniform(0, 2, n),
                           10.0 ** rng.uniform(-12, 4, n), special])


@pytest.mark.parametrize("k", range(2, 14))
def test_bitwise_equal_to_libm(lib, k):
    rng = np.random.default_rng(k)
    x = np.ascontiguousarray(pow_inputs(rng, 300_000))
    a, b = np.empty_like(x), np.empty_like(x)
    p = lambda v: v.ctypes.data_as(ct.c_void_p)  # noqa: E731
    lib.b200pow_host(p(x), ct.c_double(1.0 / k), p(a), ct.c_long(len(x)))
    lib.libm_pow(p(x), ct.c_double(1.0 / k), p(b), ct.c_long(len(x)))
    same = (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))
    assert same.all(), (x[~same][:3], a[~same][:3], b[~same][:3])
