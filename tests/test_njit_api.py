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


@nb.njit
def ivp_twice(h, ts, y0, te, empty, P):
    a = _ivp(h, ts, y0, te, P, 1, 0.0, 0.0, 1e-7, 1e-9)
    b = _ivp(h, ts, y0, empty, P, 1, 0.0, 0.0, 1e-7, 1e-9)
    return a, b


def _ivp_inputs(B):
    rng = np.random.default_rng(3)
    P = np.column_stack([rng.uniform(9, 11, B), rng.uniform(20, 36, B), rng.uniform(2, 3.33, B),
                         rng.uniform(30, 45, B)])
    y0 = np.column_stack([rng.uniform(0.5, 1.5, B), rng.uniform(-0.5, 0.5, B), rng.uniform(0, 1, B)])
    ts = np.column_stack([np.zeros(B), rng.uniform(0.3, 3.0, B)])
    return ts, y0, np.linspace(0.0, 3.0, 31), np.empty(0), P


def test_solve_ivp_compiles_and_fails_loudly_without_a_device():
    import torch
    import numbalsoda_b200 as b2
    from numbalsoda_b200 import njit_api
    global _ivp
    _ivp = njit_api.solve_ivp_batched
    h = njit_api.link_ivp(b2.builtin("lorenz"), 3, 1, b2.builtin("lorenz_z"), strict=True)
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    (t, y, off, res, tev, yev), (t2, y2, off2, res2, _, _) = ivp_twice(h, *_ivp_inputs(4))
    assert y.shape == (4 * 31, 3) and off.tolist() == [0, 31, 62, 93, 124] and not res[:, 0].any()
    assert y2.shape == (0, 3) and not res2[:, 0].any()
    assert "no CPU fallback" in njit_api.last_status()


@pytest.mark.gpu
def test_solve_ivp_njit_matches_the_oracle():
    import numbalsoda_b200 as b2
    from numbalsoda_b200 import njit_api
    global _ivp
    _ivp = njit_api.solve_ivp_batched
    h = njit_api.link_ivp(b2.builtin("lorenz"), 3, 1, b2.builtin("lorenz_z"), strict=True)
    B = 200
    ts, y0, te, empty, P = _ivp_inputs(B)
    (t, y, off, res, tev, yev), (t2, y2, off2, res2, tev2, yev2) = ivp_twice(h, ts, y0, te, empty, P)
    f, ev = orc.builtin_rhs("lorenz"), orc.builtin_rhs("events_lorenz_z")
    ra = orc.solve_ivp_batch(f, ts, y0, te, 1, ev, P, rtol=1e-7, atol=1e-9)
    rb = orc.solve_ivp_batch(f, ts, y0, None, 1, ev, P, rtol=1e-7, atol=1e-9, cap=2000)
    for r, rr, tv, yv in ((res, ra, tev, yev), (res2, rb, tev2, yev2)):
        g = rr["result"]
        assert np.array_equal(r[:, 0], g["success"]) and np.array_equal(r[:, 3], g["nt_out"])
        assert np.array_equal(r[:, 8], g["nstep"]) and np.array_equal(tv, g["t_event"])
        assert np.array_equal(yv, rr["y_event"])
    assert 0 < res[:, 5].sum() < B
    for b in range(B):
        k = min(res[b, 3], np.searchsorted(te, ts[b, 1], side="right"))
        assert np.array_equal(y[off[b]:off[b] + k], ra["y"][b, :k])
        k2 = res2[b, 3]
        assert np.array_equal(t2[off2[b]:off2[b] + k2], rb["t"][b, :k2])
        assert np.array_equal(y2[off2[b]:off2[b] + k2], rb["y"][b, :k2])


# ---- the reference's own call, verbatim, from inside @njit (README.md:60-67, driver.py:28)
def _reference_style():
    import numbalsoda_b200 as b2
    from numbalsoda_b200 import lsoda, dop853
    cf = nb.cfunc(b2.lsoda_sig)(pr.lv_rhs)
    funcptr = cf.address
    u0, t_eval, data = np.array([5.0, 0.8]), np.linspace(0.0, 10.0, 11), np.array([1.0])

    @nb.njit
    def test():
        usol, success = lsoda(funcptr, u0, t_eval, data=data, rtol=1e-8, atol=1e-8)
        return usol, success

    @nb.njit
    def test_dop(fp, P):
        usol, success = dop853(fp, u0, t_eval, P, 1e-8, 1e-8)
        return usol, success

    return cf, test, test_dop


def test_reference_signature_resolves_in_nopython_code_and_fails_loudly_without_a_device():
    import torch
    import numbalsoda_b200 as b2
    cf, test, test_dop = _reference_style()
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    for call in (lambda: test(), lambda: test_dop(cf.address, np.ones((4, 1)))):
        with pytest.raises(b2.B200OdeError, match="no CPU fallback"):
            call()


@pytest.mark.gpu
def test_reference_signature_inside_njit_matches_the_reference_known_answer():
    cf, test, test_dop = _reference_style()
    usol, success = test()
    # /root/reference/tests/test_numbalsoda.py:19-20 (default mode: FMA contraction, so isclose)
    assert success and usol.shape == (11, 2)
    assert np.all(np.isclose(usol[-1], [0.38583246250193476, 4.602012234037773]))
    P = np.random.default_rng(1).uniform(0.5, 2.0, (6, 1))
    us, ok = test_dop(cf.address, P)
    ref, okr, _ = orc.solve_batch("dop853", cf.address, np.array([5.0, 0.8]), np.linspace(0.0, 10.0, 11), P,
                                  1e-8, 1e-8)
    assert us.shape == (6, 11, 2) and ok.all() and (np.abs(us - ref) / (1e-8 * np.abs(ref) + 1e-8)).max() < 10.0
