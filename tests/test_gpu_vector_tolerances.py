"""Per-component tolerances on the GPU: rtol / atol of shape [n],
[1,n] or [B,n] through the drop-in drivers and the C ABI, strict mode bit for bit against the
CPU oracle -- which is pinned to the compiled reference run with itol_ = 4 (LSODA) and by scale
invariance (both solvers): tests/test_vector_tolerances.py."""
import ctypes as ct

import numpy as np
import pytest

import _oracle as orc
import _problems as pr

pytestmark = pytest.mark.gpu


def _nb():
    import numbalsoda_b200 as nb
    return nb


def _tols(shape, seed, lo=(-9, -5), hi=(-12, -7)):
    rng = np.random.default_rng(seed)
    return 10.0 ** rng.uniform(*lo, shape), 10.0 ** rng.uniform(*hi, shape)


@pytest.mark.parametrize("solver,rhs,u0,te,sweep", [
    ("lsoda", "robertson", [1.0, 0.0, 0.0], np.linspace(0, 1e5, 100), pr.robertson_sweep),
    ("dop853", "lorenz", [1.0, 0.0, 0.0], np.linspace(0, 5, 51), pr.lorenz_sweep),
    ("lsoda", "lorenz", [1.0, 0.0, 0.0], np.linspace(0, 5, 51), pr.lorenz_sweep),
])
def test_thread_kernels_every_shape_strict_bitwise(solver, rhs, u0, te, sweep):
    nb = _nb()
    api = nb.lsoda if solver == "lsoda" else nb.dop853
    B, n = 700, 3
    P = sweep(B, 4)
    u0 = np.array(u0)
    f = orc.builtin_rhs(rhs)
    rv, av = _tols(n, 1)
    rB, aB = _tols(B, 2)
    rBn, aBn = _tols((B, n), 3)
    cases = [(rv, av), (rv[None, :], 1e-10), (1e-7, av), (rBn, aBn), (rB, av), (rBn, 1e-9), (rv, aB)]
    for rt, at in cases:
        us, ok, cnt = api(nb.builtin(rhs), u0, te, P, rt, at, strict=True, counters=True)
        ref, okr, cr = orc.solve_batch(solver, f, u0, te, P, rt, at)
        assert ok.all() and okr.all()
        assert np.array_equal(cnt[:, :7], cr[:, :7]) and np.array_equal(us, ref)
    # the default (throughput) mode takes the same arrays
    us, ok = api(nb.builtin(rhs), u0, te, P, rBn, aBn)
    ref, okr, cr = orc.solve_batch(solver, f, u0, te, P, rBn, aBn)
    # (tolerances down to 1e-12 on a chaotic system: only closeness at the loosest level is asked;
    # accuracy of the throughput mode is the subject of tests/test_gpu_truth.py)
    assert ok.all() and (np.abs(us - ref) / (1.0 + np.abs(ref))).max() < 1e3 * rBn.max()


def test_single_trajectory_and_numba_rhs():
    nb = _nb()
    p = pr.problem("robertson")
    rv, av = np.array([1e-8, 1e-5, 1e-7]), np.array([1e-10, 1e-14, 1e-8])
    us, ok, cnt = nb.lsoda(p["rhs"], p["u0"], p["t_eval"], p["data"], rv, av, strict=True, counters=True)
    ref, okr, cr = orc.solve_batch("lsoda", pr.cfunc_address(p["rhs"]), p["u0"], p["t_eval"],
                                   p["data"][None, :], rv, av)
    assert ok and us.shape == (100, 3) and np.array_equal(us, ref[0]) and list(cnt) == list(cr[0])
    # the compiled reference itself (itol_ = 4)
    if orc.HAVE_REF:
        r2, ok2, st = orc.ref_lsoda_vtol(pr.cfunc_address(p["rhs"]), p["u0"], p["t_eval"], p["data"], rv, av)
        assert np.array_equal(us, r2) and list(cnt) == st
    # failures keep the reference's semantics
    for rt, at in ((np.array([1e-6, -1e-6, 1e-6]), 1e-9), (np.zeros(3), np.zeros(3))):
        us, ok, cnt = nb.lsoda(p["rhs"], p["u0"], p["t_eval"], p["data"], rt, at, strict=True, counters=True)
        ref, okr, cr = orc.solve_batch("lsoda", pr.cfunc_address(p["rhs"]), p["u0"], p["t_eval"],
                                       p["data"][None, :], rt, at)
        assert ok == okr[0] and list(cnt) == list(cr[0])


@pytest.mark.parametrize("solver", ["lsoda", "dop853"])
def test_warp_kernels_n64(solver):
    nb = _nb()
    api = nb.lsoda if solver == "lsoda" else nb.dop853
    B, n = 24, 64
    rng = np.random.default_rng(3)
    D = 10.0 ** (rng.uniform(1.0, 2.5, B) if solver == "lsoda" else rng.uniform(-1, 0.5, B))
    P = np.stack([rng.uniform(0.8, 1.2, B), rng.uniform(2.5, 3.5, B), D], axis=1)
    u0 = np.stack([pr.brusselator_u0(P[b, 0], P[b, 1]) for b in range(B)])
    te = np.linspace(0.0, 2.0, 21)
    f = orc.builtin_rhs("brusselator1d")
    rv, av = _tols(n, 5, lo=(-7, -5), hi=(-10, -7))
    rBn, aBn = _tols((B, n), 6, lo=(-7, -5), hi=(-10, -7))
    for rt, at in ((rv, av), (rBn, aBn), (1e-6, av)):
        us, ok, cnt = api(nb.builtin("brusselator1d"), u0, te, P, rt, at, strict=True, counters=True)
        ref, okr, cr = orc.solve_batch(solver, f, u0, te, P, rt, at)
        assert ok.all() and okr.all()
        assert np.array_equal(cnt[:, :7], cr[:, :7]) and np.array_equal(us, ref)


def test_device_entry_point_and_chunked_host_path():
    """b200ode_solve_device with the three array shapes, and a host-path call large enough to be
    cut into chunks (the per-trajectory offsets into [B,n] arrays)."""
    import torch
    nb = _nb()
    from numbalsoda_b200 import _lib
    B, n, nt = 3000, 3, 21
    P = pr.lorenz_sweep(B, 9)
    teh = np.linspace(0, 2, nt)
    u0h = np.array([1.0, 0.0, 0.0])
    rBn, aBn = _tols((B, n), 7)
    rv, av = _tols(n, 8)
    f = orc.builtin_rhs("lorenz")
    h = nb.get_handle(nb.builtin("lorenz"), _lib.DOP853, n, strict=True)
    Pd, te, u0 = torch.tensor(P, device="cuda"), torch.tensor(teh, device="cuda"), torch.tensor(u0h, device="cuda")
    for rt, at, rd, ad in ((rBn, aBn, _lib.TOL_FULL, _lib.TOL_FULL), (rv, aBn[:, 0].copy(), _lib.TOL_PER_COMPONENT,
                                                                       _lib.TOL_PER_TRAJECTORY),
                           (rBn, av, _lib.TOL_FULL, _lib.TOL_PER_COMPONENT)):
        usol = torch.full((B, nt, n), float("nan"), dtype=torch.float64, device="cuda")
        ok = torch.zeros(B, dtype=torch.int32, device="cuda")
        rtd, atd = torch.tensor(rt, device="cuda"), torch.tensor(at, device="cuda")
        pb = _lib.Problem()
        pb.size = ct.sizeof(_lib.Problem)
        pb.B, pb.neq, pb.nt, pb.np = B, n, nt, 3
        pb.u0, pb.u0_stride, pb.data, pb.data_stride = u0.data_ptr(), 0, Pd.data_ptr(), 24
        pb.teval, pb.usol, pb.mxstep, pb.success = te.data_ptr(), usol.data_ptr(), 10000, ok.data_ptr()
        pb.rtol_b, pb.atol_b, pb.rtol_dims, pb.atol_dims = rtd.data_ptr(), atd.data_ptr(), rd, ad
        _lib.check(_lib.lib.b200ode_solve_device(h.ptr, ct.byref(pb), None))
        torch.cuda.synchronize()
        ref, okr, _ = orc.solve_batch("dop853", f, u0h, teh, P, rt, at)
        assert bool((ok == 1).all()) and np.array_equal(usol.cpu().numpy(), ref)
    # host path in many chunks: 300 000 trajectories x 5 rows
    B2 = 300000
    P2 = pr.lorenz_sweep(B2, 10)
    te2 = np.linspace(0, 0.5, 5)
    r2, a2 = _tols((B2, n), 11)
    us, ok = nb.dop853(nb.builtin("lorenz"), u0h, te2, P2, r2, a2, strict=True)
    idx = np.random.default_rng(0).choice(B2, 400, replace=False)
    ref, okr, _ = orc.solve_batch("dop853", f, u0h, te2, P2[idx], r2[idx], a2[idx])
    assert ok.all() and np.array_equal(us[idx], ref)
