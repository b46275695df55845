"""CPU twin of tests/test_gpu_truth.py: the throughput mode's arithmetic, emulated on the host
(LSODA: the device core compiled as host C++ with -DB200_HOST_FAST, tests/lsoda_core_host.cpp;
DOP853: the oracle's test-only fast-controller knob), is as accurate against a tight-tolerance
truth as the reference arithmetic.  The GPU test makes the same statement about the real kernels
(FMA contraction and the MUFU-based functions included)."""
import numpy as np

import _oracle as orc
import _problems as pr
import test_lsoda_core_host as host


def werr(us, truth, rtol, atol):
    return (np.abs(us - truth) / (rtol * np.abs(truth) + atol)).reshape(len(us), -1).max(axis=1)


def test_lsoda_throughput_arithmetic_as_accurate_as_the_reference():
    B = 160
    P = pr.robertson_sweep(B)
    f = orc.builtin_rhs("robertson")
    u0, te = np.array([1.0, 0.0, 0.0]), np.linspace(0, 1e5, 100)
    truth, okt, _ = orc.solve_batch("lsoda", f, u0, te, P, 1e-13, 1e-13, 500000)
    ref, okr, cr = orc.solve_batch("lsoda", f, u0, te, P, 1e-8, 1e-8)
    assert okt.all() and okr.all()
    host._backend = "thread"
    fast = np.empty_like(ref)
    nst = []
    for b in range(B):
        fast[b], ok, cnt = host.run_core(f, u0, te, P[b], 1e-8, 1e-8, fast=True)
        assert ok
        nst.append(cnt[0])
    e_ref, e_fast = werr(ref, truth, 1e-8, 1e-8), werr(fast, truth, 1e-8, 1e-8)
    # the reference's own global error on this sweep is ~11 units of rtol*|y|+atol
    assert 1.0 < e_ref.max() < 30.0
    assert e_fast.max() <= 1.5 * e_ref.max() and e_fast.mean() <= 1.25 * e_ref.mean()
    assert abs(np.mean(nst) / cr[:, 0].mean() - 1.0) < 0.02


def test_dop853_throughput_controller_as_accurate_as_the_reference():
    B = 100
    f = orc.builtin_rhs("lorenz")
    u0, te = np.array([1.0, 0.0, 0.0]), np.linspace(0, 10, 101)
    P = pr.lorenz_sweep(B)
    truth, okt, _ = orc.solve_batch("dop853", f, u0, te, P, 1e-13, 1e-13, 500000)
    ref, okr, cr = orc.solve_batch("dop853", f, u0, te, P, 1e-8, 1e-8)
    orc.lib.oracle_dop853_set_fast_ctrl(1)
    try:
        fast, okf, cf = orc.solve_batch("dop853", f, u0, te, P, 1e-8, 1e-8)
    finally:
        orc.lib.oracle_dop853_set_fast_ctrl(0)
    assert okt.all() and okr.all() and okf.all()
    e_ref, e_fast = werr(ref, truth, 1e-8, 1e-8), werr(fast, truth, 1e-8, 1e-8)
    assert e_fast.max() <= 1.5 * e_ref.max() and e_fast.mean() <= 1.25 * e_ref.mean()
    assert np.abs(cf[:, 0].astype(float) - cr[:, 0]).max() <= 3
