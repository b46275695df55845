"""The banded finite-difference Jacobian (lsoda(..., ml=, mu=)) on the
GPU: strict mode bit for bit against the oracle's band path (pinned in tests/test_banded.py against
the dense reference path and ODEPACK's own jt = 5), and against the dense run within tolerance."""
import numpy as np
import pytest

import _oracle as orc
import _problems as pr

pytestmark = pytest.mark.gpu


def _nb():
    import numbalsoda_b200 as nb
    return nb


def test_brusselator_n64_band_2_2_strict_bitwise_and_close_to_dense():
    nb = _nb()
    B = 200
    P = pr.brusselator_sweep(B, 3)
    u0, te = pr.interleave(pr.brusselator_u0()), np.linspace(0, 10, 100)
    f = orc.builtin_rhs("brusselator1d_il")
    us, ok, cnt = nb.lsoda(nb.builtin("brusselator1d_il"), u0, te, P, 1e-6, 1e-8, strict=True, counters=True,
                           ml=2, mu=2)
    ref, okr, cr = orc.solve_batch("lsoda", f, u0, te, P, 1e-6, 1e-8, band=(2, 2))
    assert ok.all() and okr.all() and np.array_equal(cnt, cr) and np.array_equal(us, ref)
    # the dense kernel on the same (interleaved) system, and the dense oracle
    ud, okd, cd = nb.lsoda(nb.builtin("brusselator1d_il"), u0, te, P, 1e-6, 1e-8, strict=True, counters=True)
    dref, _, cdr = orc.solve_batch("lsoda", f, u0, te, P, 1e-6, 1e-8)
    assert okd.all() and np.array_equal(ud, dref) and np.array_equal(cd, cdr)
    e = (np.abs(us - ud) / (1e-6 * np.abs(ud) + 1e-8)).reshape(B, -1).max(axis=1)
    print("band vs dense, strict: max %.3g median %.3g units of rtol*|y|+atol" % (e.max(), np.median(e)))
    # (nearly all trajectories take the dense path's steps exactly and agree to ~1e-5 units; one in
    # a hundred decides a step differently and then differs by the two runs' global errors)
    assert np.quantile(e, 0.95) < 10.0 and e.max() < 100.0 and np.median(e) < 1e-2  # north_star: 10 rtol
    # nearly every trajectory takes the dense path's steps and Jacobians exactly; 5 RHS calls each
    assert (cnt[:, [0, 2]] == cd[:, [0, 2]]).all(axis=1).mean() >= 0.9 and (cnt[:, 1] < 0.7 * cd[:, 1]).all()
    # default (throughput) mode: another rounding, another -- equally valid -- step sequence; the
    # two runs agree to their common global error (tests/test_gpu_truth.py explains the bound)
    uf, okf = nb.lsoda(nb.builtin("brusselator1d_il"), u0, te, P, 1e-6, 1e-8, ml=2, mu=2)
    ef = (np.abs(uf - ud) / (1e-6 * np.abs(ud) + 1e-8)).reshape(B, -1).max(axis=1)
    print("band fast vs dense strict: max %.3g median %.3g" % (ef.max(), np.median(ef)))
    assert okf.all() and ef.max() < 100.0 and np.median(ef) < 5.0


@pytest.mark.parametrize("n,ml,mu", [(20, 1, 1), (33, 1, 1), (12, 3, 2), (40, 5, 7), (9, 8, 8), (10, 0, 2),
                                     (10, 2, 0), (128, 1, 1)])
def test_bandwidths_with_a_numba_rhs_strict_bitwise(n, ml, mu):
    nb = _nb()
    u0 = 1.0 + 0.1 * np.arange(n)
    P = np.array([[200.0 * (1 + b), float(n)] for b in range(7)])
    te = np.linspace(0, 5, 11)
    us, ok, cnt = nb.lsoda(pr.chain_rhs, u0, te, P, 1e-7, 1e-9, strict=True, counters=True, ml=ml, mu=mu)
    # (a band that needs as much room as the full matrix -- here 9, 8, 8 -- runs the dense path)
    full = 2 * ml + mu + 1 >= n
    ref, okr, cr = orc.solve_batch("lsoda", pr.cfunc_address(pr.chain_rhs), u0, te, P, 1e-7, 1e-9,
                                   band=None if full else (ml, mu))
    assert list(ok) == list(okr) and np.array_equal(cnt, cr) and np.array_equal(us, ref)


def test_small_systems_keep_the_dense_matrix():
    """n <= 8 runs one trajectory per thread with the dense matrix in shared memory: ml / mu are
    accepted and the result is the dense one (a banded matrix is a dense matrix with zeros)."""
    nb = _nb()
    p = pr.problem("robertson")
    a = nb.lsoda(nb.builtin("robertson"), p["u0"], p["t_eval"], p["data"], 1e-8, 1e-8, strict=True, ml=2, mu=2)
    b = nb.lsoda(nb.builtin("robertson"), p["u0"], p["t_eval"], p["data"], 1e-8, 1e-8, strict=True)
    assert a[1] and np.array_equal(a[0], b[0])


def test_lane_parallel_rhs_strict_bitwise():
    """lsoda / dop853(..., rhs_lanes=f): all 32 lanes of a trajectory's warp share the evaluation of
    the right-hand side (a method-of-lines loop split over the lanes) instead of lane 0 doing it
    alone -- the same operations per component, so the result is the scalar one bit for bit:
    numba-compiled pair against the oracle, dense and banded, and the built-in pair."""
    nb = _nb()
    B = 60
    P = pr.brusselator_sweep(B, 4)
    u0, te = pr.interleave(pr.brusselator_u0()), np.linspace(0, 10, 100)
    fo = pr.cfunc_address(pr.brusselator_il_rhs)
    for kw, band in ((dict(ml=2, mu=2), (2, 2)), (dict(), None)):
        us, ok, cnt = nb.lsoda(pr.brusselator_il_rhs, u0, te, P, 1e-6, 1e-8, strict=True, counters=True,
                               rhs_lanes=pr.brusselator_il_lanes, **kw)
        ref, okr, cr = orc.solve_batch("lsoda", fo, u0, te, P, 1e-6, 1e-8, band=band)
        assert ok.all() and np.array_equal(cnt, cr) and np.array_equal(us, ref)
        # without the lanes function (lane 0 alone): the same bits
        us2, ok2 = nb.lsoda(pr.brusselator_il_rhs, u0, te, P, 1e-6, 1e-8, strict=True, **kw)
        assert np.array_equal(us2, us)
    # dop853, warp per trajectory (mild diffusion: non-stiff)
    Pd = np.stack([P[:, 0], P[:, 1], 10.0 ** np.random.default_rng(1).uniform(-1, 0.5, B)], axis=1)
    te2 = np.linspace(0, 2, 21)
    us, ok, cnt = nb.dop853(pr.brusselator_il_rhs, u0, te2, Pd, 1e-7, 1e-9, strict=True, counters=True,
                            rhs_lanes=pr.brusselator_il_lanes)
    ref, okr, cr = orc.solve_batch("dop853", fo, u0, te2, Pd, 1e-7, 1e-9)
    assert ok.all() and np.array_equal(cnt[:, :7], cr[:, :7]) and np.array_equal(us, ref)
