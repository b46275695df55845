"""numbalsoda_b200.njit_api: the batched solvers called from inside @njit code through ctypes
function objects, the way the reference's own drivers call lsoda_wrapper / dop853_wrapper
(numbalsoda/driver.py:28-43, driver_dop853.py:23-35; README.md:60-67)."""
import numpy as np
import pytest

import _oracle as orc
import _problems as pr

nb = pytest.importorskip("numba")


@nb.njit
def two_sweeps(h_dop, h_lsoda, u0, te_d, P_d, te_l, P_l, rtol):
    # several integrations in a row without leaving nopython code (the reference's use case)
    a, ok_a = _dop(h_dop, u0, te_d, P_d, rtol, rtol)
    b, ok_b = _lsoda(h_lsoda, u0, te_l, P_l, rtol, rtol, 10000, False)
    return a, ok_a, b, ok_b


def _api():
    from numbalsoda_b200 import njit_api
    global _dop, _lsoda
    _dop, _lsoda = njit_api.dop853_batched, njit_api.lsoda_batched
    return njit_api


def _inputs(B):
    u0 = np.tile(np.array([1.0, 0.0, 0.0]), (B, 1))
    return (u0, np.linspace(0.0, 2.0, 21), pr.lorenz_sweep(B), np.linspace(0.0, 1.0e3, 20),
            pr.robertson_sweep(B))


def test_compiles_and_fails_loudly_without_a_device():
    import torch
    import numbalsoda_b200 as b2
    api = _api()
    hd = api.link(b2.builtin("lorenz"), "dop853", 3, strict=True)
    hl = api.link(b2.builtin("robertson"), "lsoda", 3, strict=True)
    assert isinstance(hd, int) and hd != hl
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    a, ok_a, b, ok_b = two_sweeps(hd, hl, *_inputs(4), 1e-8)
    assert a.shape == (4, 21, 3) and b.shape == (4, 20, 3)
    assert not ok_a.any() and not ok_b.any()  # no CPU fallback: every trajectory reports failure
    assert "no CPU fallback" in api.last_status()


@pytest.mark.gpu
def test_njit_calls_match_the_oracle():
    import numbalsoda_b200 as b2
    api = _api()
    hd = api.link(b2.builtin("lorenz"), "dop853", 3, strict=True)
    hl = api.link(b2.builtin("robertson"), "lsoda", 3, strict=True)
    B = 300
    u0, te_d, P_d, te_l, P_l = _inputs(B)
    a, ok_a, b, ok_b = two_sweeps(hd, hl, u0, te_d, P_d, te_l, P_l, 1e-8)
    ra, oka, _ = orc.solve_batch("dop853", orc.builtin_rhs("lorenz"), u0, te_d, P_d, 1e-8, 1e-8)
    rb, okb, _ = orc.solve_batch("lsoda", orc.builtin_rhs("robertson"), u0, te_l, P_l, 1e-8, 1e-8)
    assert ok_a.all() and ok_b.all() and oka.all() and okb.all()
    assert np.array_equal(a, ra) and np.array_equal(b, rb)
