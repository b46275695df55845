"""Banded finite-difference Jacobian for LSODA (jt = 5 with ml, mu).

The reference's option checker accepts jt = 4 / 5 with ml, mu (LSODA.cpp:369-392) but its prja
and solsy implement miter = 2 only (:1656-1660, :1936-1942), so there is no reference output to
be identical to.  What is checked instead, without a GPU:
  * the oracle's band routines (restating ODEPACK's DPRJA / DSOLSY miter = 5 and LINPACK's
    dgbfa / dgbsl) against the DENSE reference path on the same problems: same accepted steps,
    states within a small fraction of the tolerance, min(ml+mu+1, n) RHS calls per Jacobian;
  * scipy's odeint, i.e. ODEPACK's LSODA itself with jt = 5, as an independent banded
    implementation: states agree within the tolerance;
  * the device core (warp backend) compiled as host C++ against the oracle, bit for bit, for
    narrow, asymmetric, one-sided and full bandwidths.
tests/test_gpu_banded.py runs the kernel."""
import ctypes as ct

import numpy as np
import pytest

import _oracle as orc
import _problems as pr
import test_lsoda_core_host as host


def _p(a):
    return a.ctypes.data_as(ct.c_void_p)


def test_banded_oracle_matches_the_dense_reference_path_on_the_brusselator():
    B = 24
    P = pr.brusselator_sweep(B, 1)
    u0, te = pr.brusselator_u0(), np.linspace(0, 10, 100)
    dense, okd, cd = orc.solve_batch("lsoda", orc.builtin_rhs("brusselator1d"), u0, te, P, 1e-6, 1e-8)
    band, okb, cb = orc.solve_batch("lsoda", orc.builtin_rhs("brusselator1d_il"), pr.interleave(u0), te, P,
                                    1e-6, 1e-8, band=(2, 2))
    assert okd.all() and okb.all()
    d = pr.interleave(dense)
    e = (np.abs(band - d) / (1e-6 * np.abs(d) + 1e-8)).reshape(B, -1).max(axis=1)
    assert e.max() < 10.0 and np.median(e) < 1e-3  # north_star's bound is 10 rtol
    # The grouped differences give the same Jacobian entries up to rounding, so nearly every
    # trajectory takes exactly the dense path's steps, Jacobians and switches (an occasional one
    # decides a step differently); there a Jacobian costs 5 right-hand sides instead of 64.
    same = (cb[:, [0, 2, 3]] == cd[:, [0, 2, 3]]).all(axis=1)
    assert same.mean() >= 0.9
    saved = (cd[:, 1] - cb[:, 1])[same]
    assert (np.abs(saved - (cd[:, 2] * (64 - 5))[same]) <= 3).all()  # (a corrector iteration more or less)


def test_banded_oracle_against_odepack_lsoda_in_scipy():
    """scipy.integrate.odeint = ODEPACK's LSODA; with ml, mu and no Dfun it runs jt = 5."""
    from scipy.integrate import odeint
    P = pr.brusselator_sweep(3, 2)
    u0, te = pr.interleave(pr.brusselator_u0()), np.linspace(0, 10, 21)
    band, ok, cnt = orc.solve_batch("lsoda", orc.builtin_rhs("brusselator1d_il"), u0, te, P, 1e-8, 1e-10,
                                    band=(2, 2))
    for b in range(3):
        def f(y, t):
            du = np.zeros(64)
            pr.brusselator_il_rhs(t, y, du, P[b])
            return du
        ys, info = odeint(f, u0, te, ml=2, mu=2, rtol=1e-8, atol=1e-10, full_output=True, mxstep=10000)
        assert (np.abs(band[b] - ys) / (1e-8 * np.abs(ys) + 1e-10)).max() < 200.0  # two codes, same method
        assert abs(info["nst"][-1] / cnt[b, 0] - 1.0) < 0.25


@pytest.mark.parametrize("n,ml,mu", [(20, 1, 1), (33, 1, 1), (12, 3, 2), (40, 5, 7), (9, 8, 8), (10, 0, 2),
                                     (10, 2, 0), (64, 2, 2)])
def test_device_core_band_path_bitwise_against_the_oracle(n, ml, mu):
    lib = host.core("warp")
    if n == 64:
        f, u0 = orc.builtin_rhs("brusselator1d_il"), pr.interleave(pr.brusselator_u0())
        P, te, tol = pr.brusselator_sweep(3, 5), np.linspace(0, 10, 100), (1e-6, 1e-8)
    else:
        f, u0 = pr.cfunc_address(pr.chain_rhs), 1.0 + 0.1 * np.arange(n)
        P, te, tol = np.array([[200.0 * (1 + b), float(n)] for b in range(3)]), np.linspace(0, 5, 11), (1e-7, 1e-9)
    ref, okr, cr = orc.solve_batch("lsoda", f, u0, te, P, *tol, band=(ml, mu))
    dense, okd, cd = orc.solve_batch("lsoda", f, u0, te, P, *tol)
    assert okr.all() and okd.all()
    # (chain_rhs couples u[i] to i-1, i+1: any band >= (1, 1) holds the whole Jacobian; one-sided
    # bands drop entries -- a poorer Newton matrix, the same converged corrector equation)
    assert (np.abs(ref - dense) / (tol[0] * np.abs(dense) + tol[1])).max() < 10.0
    for b in range(len(P)):
        us = np.full((len(te), n), np.nan)
        ok = ct.c_int(9)
        cnt = (ct.c_int * 8)()
        lib.lsoda_core_host_band(ct.c_void_p(f), n, _p(u0), _p(P[b]), len(te), _p(te), _p(us),
                                 ct.c_double(tol[0]), ct.c_double(tol[1]), ml, mu, 10000, ct.byref(ok), 0, cnt)
        assert ok.value == 1 and list(cnt) == list(cr[b])
        assert np.array_equal(us, ref[b])


def test_band_wider_than_the_system_is_refused_like_the_reference_checker():
    # LSODA.cpp:376-391: ml >= n or mu >= n is illegal input
    f = orc.builtin_rhs("lotka_volterra")
    us, ok, cnt = orc.solve_batch("lsoda", f, np.array([5.0, 0.8]), np.linspace(0, 1, 3), np.array([[1.0]]),
                                  1e-6, 1e-9, band=(2, 0))
    assert not ok[0] and cnt[0, 5] == -3
    import numbalsoda_b200 as nb
    with pytest.raises(ValueError):
        nb.lsoda(nb.builtin("lotka_volterra"), np.array([5.0, 0.8]), np.linspace(0, 1, 3), np.array([1.0]),
                 ml=2, mu=0)
    with pytest.raises(ValueError):
        nb.lsoda(nb.builtin("lotka_volterra"), np.array([5.0, 0.8]), np.linspace(0, 1, 3), np.array([1.0]), ml=1)
