"""Parity of the CUDA LSODA path (through the C ABI / numbalsoda-compatible driver)
with the CPU oracle -- itself pinned bit-for-bit to the compiled reference
(tests/test_oracle_lsoda.py) -- and with the golden vectors generated from the
reference's own C++ (tests/golden/lsoda_reference.npz).  strict=True links the
kernels with FMA contraction disabled: states are then compared bit-for-bit and the
step / RHS / Jacobian / method-switch counters must be identical."""
import os

import numpy as np
import pytest

import _oracle as orc
import _problems as pr

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "lsoda_reference.npz"))
CASES = ["lv_test", "lv_readme", "robertson", "lorenz_short", "lorenz"]


def _nb():
    import numbalsoda_b200 as nb
    return nb


def _cnt(st):
    return [st.nst, st.nfe, st.nje, st.nswitch, st.nqu, st.istate, st.rows, st.meth]


def same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("kind", ["builtin", "numba"])
def test_single_trajectory_strict_bitwise(name, kind):
    nb = _nb()
    p = pr.problem(name)
    rhs = nb.builtin(p["builtin"]) if kind == "builtin" else p["rhs"]
    us, ok, cnt = nb.lsoda(rhs, p["u0"], p["t_eval"], p["data"], p["rtol"], p["atol"],
                           strict=True, counters=True)
    ref, okr, st = orc.lsoda(orc.builtin_rhs(p["builtin"]), p["u0"], p["t_eval"], p["data"],
                             p["rtol"], p["atol"])
    assert ok == okr
    assert list(cnt) == _cnt(st)
    assert same(us, ref)
    # the reference's own C++, run when the fixtures were made
    assert same(us, GOLD[name + "/usol"])
    assert list(cnt) == list(GOLD[name + "/counters"])


def test_reference_known_answers():
    # /root/reference/tests/test_numbalsoda.py:11-20 and README.md / comparison2scipy.ipynb:175,
    # through the drop-in driver with a cfunc address, exactly as a numbalsoda user calls it
    import numba
    nb = _nb()
    p = pr.problem("lv_test")
    cf = numba.cfunc(nb.lsoda_sig)(p["rhs"])
    usol, success = nb.lsoda(cf.address, p["u0"], p["t_eval"], p["data"], rtol=1e-8, atol=1e-8,
                             strict=True)
    assert success and usol.shape == (11, 2)
    assert list(usol[-1]) == [0.38583246250193476, 4.602012234037773]  # bit-for-bit
    usol, success = nb.lsoda(cf.address, p["u0"], p["t_eval"], p["data"], rtol=1e-8, atol=1e-8)
    assert success and np.all(np.isclose(usol[-1], [0.38583246250193476, 4.602012234037773]))
    p = pr.problem("lv_readme")
    usol, success = nb.lsoda(cf.address, p["u0"], p["t_eval"], p["data"], strict=True)
    assert success and list(usol[-1]) == [0.15204821320141523, 0.10552706409364211]


def test_exit_on_warning_like_the_reference():
    # /root/reference/tests/test_exit_on_warning.py (u0 = int64 1 reinterpreted as a double)
    nb = _nb()
    u0 = np.array([1]).view(np.float64)
    te = np.linspace(100, 101, 11)
    d = np.array([1e15])
    _, ok = nb.lsoda(pr.growth_rhs, u0, te, d, rtol=1e-9, atol=1e-9, exit_on_warning=True)
    assert not ok
    _, ok = nb.lsoda(pr.growth_rhs, u0, te, d, rtol=1e-9, atol=1e-9, exit_on_warning=False)
    assert ok


@pytest.mark.parametrize("strict", [True, False])
def test_robertson_sweep_against_oracle(strict):
    """Config C3's parameter sweep (full benchmark_robber.py shape) at B = 2000."""
    nb = _nb()
    B = 2000
    P = pr.robertson_sweep(B)
    u0 = np.array([1.0, 0.0, 0.0])
    te = np.linspace(0, 1e5, 100)
    us, ok, cnt = nb.lsoda(nb.builtin("robertson"), u0, te, P, 1e-8, 1e-8, strict=strict,
                           counters=True)
    ref, okr, cr = orc.solve_batch("lsoda", orc.builtin_rhs("robertson"), u0, te, P, 1e-8, 1e-8)
    assert ok.all() and okr.all() and us.shape == (B, 100, 3)
    if strict:
        assert np.array_equal(cnt, cr)
        assert np.array_equal(us, ref)
        assert (cnt[:, 3] >= 1).all() and (cnt[:, 7] == 2).all()  # every one switched to BDF
    else:
        # Fused multiply-adds perturb the last bits and a different -- equally valid -- step
        # sequence follows.  Two such runs agree to their common global error, which for the
        # reference itself is ~11 units of rtol*|y| + atol on this sweep: the accuracy
        # statement proper (each mode against a tight-tolerance truth, beside the reference's
        # own error) is tests/test_gpu_truth.py.  Here only the triangle inequality: the
        # distance between the two runs cannot exceed the sum of their global errors.
        err = np.abs(us - ref) / (1e-8 * np.abs(ref) + 1e-8)
        print("fast-mode max weighted distance to the oracle: %.3g" % err.max())
        assert err.max() <= 2 * 1.5 * 11.2
        assert np.abs(cnt[:, 0].astype(float) - cr[:, 0]).mean() < 40


def test_numba_robertson_sweep_counts_match_reference_counts():
    """north_star: integer step / Jacobian-evaluation / method-switch counts match on
    Lotka-Volterra and Robertson -- with the RHS compiled by numba.cuda."""
    nb = _nb()
    B = 256
    P = pr.robertson_sweep(B, seed=3)
    u0 = np.array([1.0, 0.0, 0.0])
    te = np.linspace(0, 1e5, 100)
    us, ok, cnt = nb.lsoda(pr.robertson_rhs, u0, te, P, 1e-8, 1e-8, strict=True, counters=True)
    fptr = pr.cfunc_address(pr.robertson_rhs)
    ref, okr, cr = orc.solve_batch("lsoda", fptr, u0, te, P, 1e-8, 1e-8)
    assert ok.all() and np.array_equal(cnt, cr) and np.array_equal(us, ref)
    Pl = np.random.default_rng(2).uniform(0.5, 3.0, (B, 1))
    te = np.linspace(0, 50, 200)
    us, ok, cnt = nb.lsoda(pr.lv_rhs, np.array([5.0, 0.8]), te, Pl, 1e-6, 1e-9, strict=True,
                           counters=True)
    ref, okr, cr = orc.solve_batch("lsoda", pr.cfunc_address(pr.lv_rhs), np.array([5.0, 0.8]), te,
                                   Pl, 1e-6, 1e-9)
    assert ok.all() and np.array_equal(cnt, cr) and np.array_equal(us, ref)


def test_batched_u0_and_shared_data():
    nb = _nb()
    rng = np.random.default_rng(5)
    B = 257
    u0 = np.array([5.0, 0.8]) * rng.uniform(0.5, 1.5, (B, 2))
    te = np.linspace(0, 10, 31)
    d = np.array([1.3])
    us, ok, cnt = nb.lsoda(nb.builtin("lotka_volterra"), u0, te, d, 1e-7, 1e-9, strict=True,
                           counters=True)
    ref, okr, cr = orc.solve_batch("lsoda", orc.builtin_rhs("lotka_volterra"), u0, te, d, 1e-7, 1e-9)
    assert ok.all() and np.array_equal(us, ref) and np.array_equal(cnt, cr)


def test_failures_and_edge_cases():
    nb = _nb()
    f = nb.builtin("lotka_volterra")
    fo = orc.builtin_rhs("lotka_volterra")
    u0, d = np.array([5.0, 0.8]), np.array([1.0])
    cases = [
        dict(te=np.linspace(0, 10, 11), rtol=1e-8, atol=1e-8, mxstep=5),       # istate -1
        dict(te=np.linspace(0, 10, 11), rtol=1e-8, atol=1e-8, mxstep=0),
        dict(te=np.linspace(0, 10, 11), rtol=1e-8, atol=1e-8, mxstep=-1),
        dict(te=np.array([0.0]), rtol=1e-8, atol=1e-8),                        # nt = 1
        dict(te=np.array([0.0, 0.0, 1.0]), rtol=1e-8, atol=1e-8),              # tout too close
        dict(te=np.array([0.0, 1.0, 1.0, 2.0, 2.0]), rtol=1e-8, atol=1e-8),    # duplicates
        dict(te=np.array([0.0, 1.0, 0.5, 2.0]), rtol=1e-8, atol=1e-8),         # not monotone
        dict(te=np.array([1.0, 0.0]), rtol=1e-8, atol=1e-8),                   # decreasing
        dict(te=np.linspace(0, 10, 11), rtol=-1.0, atol=1e-8),                 # illegal tolerance
        dict(te=np.linspace(0, 10, 11), rtol=0.0, atol=0.0),                   # ewt <= 0 quirk
        dict(te=np.linspace(0, 10, 11), rtol=0.0, atol=1e-10),
        dict(te=np.linspace(0, 10, 11), rtol=1e-17, atol=1e-20),               # too much accuracy
        dict(te=np.array([0.0, 1e-9, 2e-9, 5.0, 5.0 + 1e-9]), rtol=1e-6, atol=1e-9),
    ]
    for c in cases:
        kw = dict(rtol=c["rtol"], atol=c["atol"], mxstep=c.get("mxstep", 10000))
        us, ok, cnt = nb.lsoda(f, u0, c["te"], d, strict=True, counters=True, **kw)
        ref, okr, st = orc.lsoda(fo, u0, c["te"], d, **kw)
        assert ok == okr and list(cnt) == _cnt(st), (c, list(cnt), _cnt(st))
        assert same(us[:st.rows], ref[:st.rows]), c
    # a batch in which only some trajectories fail (per-interval step budget)
    U = np.array([[5.0, 0.8], [0.5, 0.08], [5.0, 0.8]])
    P = np.array([[1.0], [1.0], [1.0]])
    te = np.linspace(0, 10, 3)
    us, ok, cnt = nb.lsoda(f, U, te, P, 1e-8, 1e-8, mxstep=120, strict=True, counters=True)
    ref, okr, cr = orc.solve_batch("lsoda", fo, U, te, P, 1e-8, 1e-8, mxstep=120)
    assert list(ok) == list(okr) and np.array_equal(cnt, cr)
    for b in range(3):
        assert same(us[b, :cr[b, 6]], ref[b, :cr[b, 6]])


def test_stiff_problems_with_pivoting():
    nb = _nb()
    import numba
    from numba import types
    sig = types.void(types.double, types.CPointer(types.double), types.CPointer(types.double),
                     types.CPointer(types.double))

    def vdp(t, u, du, p):
        du[0] = u[1]
        du[1] = p[0] * ((1.0 - u[0] * u[0]) * u[1]) - u[0]

    def lin3(t, u, du, p):
        du[0] = -p[0] * u[0] + p[1] * u[1]
        du[1] = p[0] * u[0] - p[1] * u[1] - p[2] * u[1] + 3.0 * u[2]
        du[2] = p[2] * u[1] - 3.0 * u[2]

    runs = [(vdp, [2.0, 0.0], [1000.0], np.linspace(0, 3000, 61), 1e-6, 1e-8),
            (lin3, [1.0, 0.0, 0.0], [1e4, 2e6, 5e3], np.linspace(0, 10, 21), 1e-7, 1e-10),
            (lin3, [1.0, 2.0, 3.0], [1e-3, 7e5, 1e7], np.linspace(0, 100, 21), 1e-5, 1e-9),
            (pr.robertson_rhs, [1.0, 0, 0], [0.04, 3e7, 1e4], np.linspace(0, 1e11, 12), 1e-2, 1e-4),
            (pr.robertson_rhs, [1.0, 0, 0], [0.04, 3e7, 1e4], np.linspace(0, 1e11, 12), 1e-10, 1e-14)]
    for f, u0, d, te, rtol, atol in runs:
        cf = numba.cfunc(sig)(f)
        u0, d = np.array(u0), np.array(d)
        us, ok, cnt = nb.lsoda(f, u0, te, d, rtol, atol, strict=True, counters=True)
        ref, okr, st = orc.lsoda(cf.address, u0, te, d, rtol, atol)
        assert ok == okr and list(cnt) == _cnt(st), (list(cnt), _cnt(st))
        assert same(us[:st.rows], ref[:st.rows])


def test_full_size_properties():
    """BASELINE config C3 at reduced B: every trajectory succeeds, row 0 is u0 exactly,
    mass is conserved (Robertson: u0+u1+u2 = 1) within the tolerance, and a permuted
    ensemble gives the permuted result (lane / launch-order independence)."""
    nb = _nb()
    B = 20000
    P = pr.robertson_sweep(B)
    u0 = np.array([1.0, 0.0, 0.0])
    te = np.linspace(0, 1e5, 100)
    us, ok = nb.lsoda(nb.builtin("robertson"), u0, te, P, 1e-8, 1e-8)
    assert ok.all() and np.isfinite(us).all()
    assert np.array_equal(us[:, 0, :], np.broadcast_to(u0, (B, 3)))
    assert np.abs(us.sum(axis=2) - 1.0).max() < 1e-5
    perm = np.random.default_rng(0).permutation(B)
    us2, ok2 = nb.lsoda(nb.builtin("robertson"), u0, te, P[perm], 1e-8, 1e-8)
    assert np.array_equal(us2, us[perm])


def test_device_resident_entry_point():
    """b200ode_lsoda_batched_device on torch-owned device memory and stream."""
    import ctypes as ct
    import torch
    nb = _nb()
    from numbalsoda_b200 import _lib
    B, nt = 512, 100
    P = torch.tensor(pr.robertson_sweep(B), device="cuda")
    u0 = torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64, device="cuda")
    te = torch.linspace(0, 1e5, nt, dtype=torch.float64, device="cuda")
    usol = torch.empty((B, nt, 3), dtype=torch.float64, device="cuda")
    ok = torch.zeros(B, dtype=torch.int32, device="cuda")
    h = nb.get_handle(nb.builtin("robertson"), _lib.LSODA, 3, strict=True)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        rc = _lib.lib.b200ode_lsoda_batched_device(
            h.ptr, B, 3, u0.data_ptr(), 0, P.data_ptr(), 24, 3, nt, te.data_ptr(),
            usol.data_ptr(), 1e-8, 1e-8, 10000, ok.data_ptr(), None, 0,
            ct.c_void_p(s.cuda_stream))
    _lib.check(rc)
    s.synchronize()
    ref, okr, _ = orc.solve_batch("lsoda", orc.builtin_rhs("robertson"), np.array([1.0, 0, 0]),
                                  te.cpu().numpy(), P.cpu().numpy(), 1e-8, 1e-8)
    assert bool((ok == 1).all()) and np.array_equal(usol.cpu().numpy(), ref)
    info = h.kernel_info()
    # (strict mode at 128 registers: up to three rarely used words on the stack, tests/test_abi.py)
    assert info["local_bytes"] <= 32 and info["launches"] >= 1


# ---------------------------------------------------------------- warp-per-trajectory kernel
def test_warp_kernel_brusselator_n64_strict_bitwise():
    """BASELINE config 4: n = 64 stiff method-of-lines sweep, one warp per trajectory
    (shared-memory Jacobian / LU): states and counters bit-for-bit vs the oracle."""
    nb = _nb()
    B = 96
    P = pr.brusselator_sweep(B)
    u0 = np.stack([pr.brusselator_u0(P[b, 0], P[b, 1]) for b in range(B)])
    te = np.linspace(0.0, 10.0, 100)
    us, ok, cnt = nb.lsoda(nb.builtin("brusselator1d"), u0, te, P, 1e-6, 1e-8, strict=True,
                           counters=True)
    ref, okr, cr = orc.solve_batch("lsoda", orc.builtin_rhs("brusselator1d"), u0, te, P, 1e-6, 1e-8)
    assert ok.all() and okr.all() and us.shape == (B, 100, 64)
    assert np.array_equal(cnt, cr)
    assert np.array_equal(us, ref)
    assert (cnt[:, 2] >= 5).all() and (cnt[:, 7] == 2).all()  # stiff: Jacobians, ends in BDF
    # fast mode (FMA): close to the strict result
    us2, ok2 = nb.lsoda(nb.builtin("brusselator1d"), u0, te, P, 1e-6, 1e-8)
    assert ok2.all()
    err = np.abs(us2 - ref) / (1e-6 * np.abs(ref) + 1e-8)
    print("fast-mode max weighted error (n=64): %.3g" % err.max())
    assert err.max() <= 100.0


def test_warp_kernel_numba_rhs_and_other_sizes():
    """numba-compiled RHS with a loop (u / du are shared-memory pointers in this kernel),
    and sizes that are not multiples of the warp: a stiff linear chain of n = 9, 20, 33."""
    import numba
    nb = _nb()
    P = pr.brusselator_sweep(8, seed=7)
    u0 = pr.brusselator_u0()
    te = np.linspace(0.0, 10.0, 30)
    us, ok, cnt = nb.lsoda(pr.brusselator_rhs, u0, te, P, 1e-6, 1e-8, strict=True, counters=True)
    ref, okr, cr = orc.solve_batch("lsoda", pr.cfunc_address(pr.brusselator_rhs), u0, te, P, 1e-6, 1e-8)
    assert ok.all() and np.array_equal(cnt, cr) and np.array_equal(us, ref)

    def chain(t, u, du, p):
        n = int(p[1])
        k = p[0]
        du[0] = -k * u[0] + 0.5 * u[1] * u[1]
        for i in range(1, n - 1):
            du[i] = k * (u[i - 1] - u[i]) / (1.0 + i) + 0.1 * u[i + 1] - 0.1 * u[i] * u[i]
        du[n - 1] = k * (u[n - 2] - u[n - 1]) / n

    cf = numba.cfunc(nb.lsoda_sig)(chain)
    for n in (9, 20, 33):
        u0n = 1.0 + 0.1 * np.arange(n)
        Pn = np.array([[50.0 * (1 + b), float(n)] for b in range(5)])
        ten = np.linspace(0.0, 5.0, 11)
        us, ok, cnt = nb.lsoda(chain, u0n, ten, Pn, 1e-7, 1e-9, strict=True, counters=True)
        ref, okr, cr = orc.solve_batch("lsoda", cf.address, u0n, ten, Pn, 1e-7, 1e-9)
        assert list(ok) == list(okr) and np.array_equal(cnt, cr), n
        for b in range(5):
            assert same(us[b, :cr[b, 6]], ref[b, :cr[b, 6]]), (n, b)


def test_warp_kernel_failures_and_edge_cases():
    nb = _nb()
    f, fo = nb.builtin("brusselator1d"), orc.builtin_rhs("brusselator1d")
    u0, d = pr.brusselator_u0(), np.array([1.0, 3.0, 300.0])
    cases = [dict(te=np.linspace(0, 10, 11), rtol=1e-6, atol=1e-8, mxstep=5),
             dict(te=np.array([0.0]), rtol=1e-6, atol=1e-8),
             dict(te=np.array([0.0, 0.0, 1.0]), rtol=1e-6, atol=1e-8),
             dict(te=np.array([0.0, 1.0, 0.5]), rtol=1e-6, atol=1e-8),
             dict(te=np.linspace(0, 10, 11), rtol=-1.0, atol=1e-8),
             dict(te=np.linspace(0, 10, 11), rtol=0.0, atol=0.0),
             dict(te=np.linspace(0, 1, 5), rtol=1e-17, atol=1e-20)]
    for c in cases:
        kw = dict(rtol=c["rtol"], atol=c["atol"], mxstep=c.get("mxstep", 10000))
        us, ok, cnt = nb.lsoda(f, u0, c["te"], d, strict=True, counters=True, **kw)
        ref, okr, st = orc.lsoda(fo, u0, c["te"], d, **kw)
        assert ok == okr and list(cnt) == _cnt(st), (c, list(cnt), _cnt(st))
        assert same(us[:st.rows], ref[:st.rows]), c


@pytest.mark.parametrize("n", [3, 12, 40])
def test_pivoting_lu_strict_bitwise_and_fast_close(n):
    """A stiff system whose iteration matrix makes the reference's row-wise pivot search
    interchange columns (thread kernel n = 3, warp kernel n = 12 / 40): strict mode
    bit-for-bit; fast mode (warp kernel: interchanges applied to the multipliers,
    swap-free column sweep, final permutation) close to it."""
    nb = _nb()
    B = 7
    u0 = np.ones(n)
    P = np.array([[1000.0 * (1 + 0.1 * b), 5000.0, float(n)] for b in range(B)])
    te = np.linspace(0.0, 2.0, 9)
    us, ok, cnt = nb.lsoda(pr.upper_rhs, u0, te, P, 1e-6, 1e-9, strict=True, counters=True)
    ref, okr, cr = orc.solve_batch("lsoda", pr.cfunc_address(pr.upper_rhs), u0, te, P, 1e-6, 1e-9)
    assert ok.all() and okr.all() and np.array_equal(cnt, cr) and np.array_equal(us, ref)
    assert (cnt[:, 2] > 0).all()
    us2, ok2 = nb.lsoda(pr.upper_rhs, u0, te, P, 1e-6, 1e-9)
    assert ok2.all()
    err = np.abs(us2 - ref) / (1e-6 * np.abs(ref) + 1e-9)
    assert err.max() < 50.0, err.max()


@pytest.mark.parametrize("n", [47, 56, 64])
def test_tensor_memory_matrix_strict_bitwise_and_same_as_shared_memory(n, monkeypatch):
    """32 < n <= 64 with fewer than eight trajectories per SM in shared memory: the dense
    iteration matrix lives in tensor memory (lsoda_vec_warp.cuh, dgefa_tm / dgesl_tm / prja_tm;
    blocks of eight warps).  Storage only: strict mode is bit-for-bit the oracle -- on a system
    whose pivot search interchanges columns and on the nonlinear chain, sizes that are and are not
    multiples of 8 and of 32 -- and gives the bits of the shared-memory kernel
    (B200ODE_LSODAW_TMEM=0); the throughput mode agrees with it to rounding."""
    nb = _nb()
    from numbalsoda_b200 import _lib
    B = 19  # (more warps than one block holds, not a multiple of 8)
    te = np.linspace(0.0, 2.0, 9)
    for rhs, u0, P, tol in (
            (pr.upper_rhs, np.ones(n), np.array([[1000.0 * (1 + 0.1 * b), 5000.0, float(n)] for b in range(B)]),
             (1e-6, 1e-9)),
            (pr.chain_rhs, 1.0 + 0.1 * np.arange(n), np.array([[50.0 * (1 + b), float(n)] for b in range(B)]),
             (1e-7, 1e-9))):
        monkeypatch.delenv("B200ODE_LSODAW_TMEM", raising=False)
        us, ok, cnt = nb.lsoda(rhs, u0, te, P, *tol, strict=True, counters=True)
        h = nb.get_handle(rhs, _lib.LSODA, n, strict=True)
        assert h.kernel_info()["last_grid"] == (B + 7) // 8  # blocks of eight warps: the tensor-memory kernel
        ref, okr, cr = orc.solve_batch("lsoda", pr.cfunc_address(rhs), u0, te, P, *tol)
        assert ok.all() and okr.all() and np.array_equal(cnt, cr) and np.array_equal(us, ref), rhs.__name__
        assert (cnt[:, 2] > 0).sum() >= B // 2  # Jacobians, factorisations and solves took place
        usf, okf, cntf = nb.lsoda(rhs, u0, te, P, *tol, counters=True)
        monkeypatch.setenv("B200ODE_LSODAW_TMEM", "0")
        us0, ok0, cnt0 = nb.lsoda(rhs, u0, te, P, *tol, strict=True, counters=True)
        assert h.kernel_info()["last_grid"] == B  # one warp per block: the shared-memory kernel
        usf0, okf0, cntf0 = nb.lsoda(rhs, u0, te, P, *tol, counters=True)
        assert np.array_equal(us0, us) and np.array_equal(cnt0, cnt)
        # throughput mode: the tensor-memory kernel's triangular sweeps are recurrences on the
        # candidates (one multiply-add between broadcasts) -- the same solution up to rounding
        assert okf.all() and okf0.all()
        for u in (usf, usf0):
            err = np.abs(u - ref) / (tol[0] * np.abs(ref) + tol[1])
            assert err.max() < 50.0, (rhs.__name__, err.max())
        assert abs(cntf[:, 0].mean() - cntf0[:, 0].mean()) <= 0.02 * cntf0[:, 0].mean() + 1.0
    monkeypatch.delenv("B200ODE_LSODAW_TMEM", raising=False)


@pytest.mark.parametrize("n", [65, 100, 128])
def test_warp_kernel_largest_dense_systems_strict_bitwise(n):
    """Above the tensor-memory kernel's 64 equations the dense matrix is back in shared memory
    (n = 128: 132 KB, one trajectory per SM): the nonlinear chain and the column-interchanging
    system, bit for bit the oracle."""
    nb = _nb()
    B = 5
    te = np.linspace(0.0, 2.0, 6)
    for rhs, u0, P, tol in (
            (pr.upper_rhs, np.ones(n), np.array([[1000.0 * (1 + 0.1 * b), 5000.0, float(n)] for b in range(B)]),
             (1e-6, 1e-9)),
            (pr.chain_rhs, 1.0 + 0.1 * np.arange(n), np.array([[50.0 * (1 + b), float(n)] for b in range(B)]),
             (1e-7, 1e-9))):
        us, ok, cnt = nb.lsoda(rhs, u0, te, P, *tol, strict=True, counters=True)
        ref, okr, cr = orc.solve_batch("lsoda", pr.cfunc_address(rhs), u0, te, P, *tol)
        assert list(ok) == list(okr) and np.array_equal(cnt, cr) and np.array_equal(us, ref), (n, rhs.__name__)
        us2, ok2 = nb.lsoda(rhs, u0, te, P, *tol)
        err = np.abs(us2 - ref) / (tol[0] * np.abs(ref) + tol[1])
        assert list(ok2) == list(okr) and err.max() < 50.0, (n, rhs.__name__, err.max())


def test_data_record_passed_by_address():
    """Records of more than 16 doubles are not staged into registers: the RHS gets the
    record's device address (the kernels' `_ptr` entry points), as it would a C struct."""
    nb = _nb()

    def rhs(t, u, du, p):
        du[0] = u[0] - u[0] * u[1] * p[19]
        du[1] = u[0] * u[1] - u[1] * p[0] + p[7]

    import numba
    cf = numba.cfunc(nb.lsoda_sig)(rhs)
    rng = np.random.default_rng(1)
    B = 70
    P = np.zeros((B, 20))
    P[:, 0] = rng.uniform(0.5, 2.0, B)
    P[:, 7] = rng.uniform(0.0, 0.1, B)
    P[:, 19] = rng.uniform(0.8, 1.2, B)
    u0 = np.array([5.0, 0.8])
    te = np.linspace(0.0, 20.0, 41)
    for solver, name in ((nb.lsoda, "lsoda"), (nb.dop853, "dop853")):
        us, ok, cnt = solver(rhs, u0, te, P, 1e-7, 1e-9, strict=True, counters=True)
        ref, okr, cr = orc.solve_batch(name, cf.address, u0, te, P, 1e-7, 1e-9)
        assert ok.all() and okr.all()
        assert np.array_equal(cnt[:, :7], cr[:, :7]) and np.array_equal(us, ref), name
    # one shared 20-double record
    us, ok = nb.lsoda(rhs, u0, te, P[3], 1e-7, 1e-9, strict=True)
    ref, okr, _ = orc.lsoda(cf.address, u0, te, P[3], 1e-7, 1e-9)
    assert ok and np.array_equal(us, ref)


def test_out_buffer_and_pinned_memory():
    nb = _nb()
    B = 300
    P = pr.robertson_sweep(B)
    u0 = np.array([1.0, 0.0, 0.0])
    te = np.linspace(0, 1e5, 100)
    ref, ok0 = nb.lsoda(nb.builtin("robertson"), u0, te, P, 1e-8, 1e-8)
    buf = nb.pinned_empty((B, 100, 3), np.float64)
    for _ in range(2):
        buf[:] = -1.0
        us, ok = nb.lsoda(nb.builtin("robertson"), u0, te, P, 1e-8, 1e-8, out=buf)
        assert ok.all() and np.shares_memory(us, buf) and np.array_equal(us, ref)
    us, ok = nb.dop853(nb.builtin("robertson"), u0, np.linspace(0, 40, 100), P, 1e-8, 1e-8,
                       mxstep=100000, out=buf)
    assert np.shares_memory(us, buf) and np.isfinite(us[ok]).all()
    with pytest.raises(ValueError):
        nb.lsoda(nb.builtin("robertson"), u0, te, P, 1e-8, 1e-8, out=np.empty((B, 99, 3)))


def test_empty_and_tiny_ensembles():
    nb = _nb()
    u0 = np.array([1.0, 0.0, 0.0])
    te = np.linspace(0, 1e5, 10)
    for solver in (nb.lsoda, nb.dop853):
        us, ok = solver(nb.builtin("robertson"), u0, te, np.empty((0, 3)), 1e-6, 1e-8)
        assert us.shape == (0, 10, 3) and ok.shape == (0,)
    us, ok = nb.lsoda(nb.builtin("robertson"), u0, te, np.array([[0.04, 3e7, 1e4]]), 1e-6, 1e-8, strict=True)
    ref, okr, _ = orc.lsoda(orc.builtin_rhs("robertson"), u0, te, [0.04, 3e7, 1e4], 1e-6, 1e-8)
    assert us.shape == (1, 10, 3) and ok[0] == okr and np.array_equal(us[0], ref)


@pytest.mark.parametrize("solver", ["lsoda", "dop853"])
def test_per_trajectory_tolerances(solver):
    """rtol / atol as arrays of shape [B]: each trajectory must equal
    the reference-order solve with its own scalar tolerances, bit for bit -- thread kernel
    (n = 2) and warp kernel (n = 64)."""
    nb = _nb()
    api = nb.lsoda if solver == "lsoda" else nb.dop853
    rng = np.random.default_rng(12)
    B = 40
    rt = 10.0 ** rng.uniform(-9, -4, B)
    at = 10.0 ** rng.uniform(-11, -6, B)
    u0 = np.array([5.0, 0.8])
    te = np.linspace(0.0, 20.0, 21)
    P = rng.uniform(0.5, 2.0, (B, 1))
    us, ok, cnt = api(nb.builtin("lotka_volterra"), u0, te, P, rt, at, strict=True, counters=True)
    f = orc.builtin_rhs("lotka_volterra")
    ofun = orc.lsoda if solver == "lsoda" else orc.dop853
    for b in range(B):
        ref, okr, st = ofun(f, u0, te, P[b], rt[b], at[b])
        assert ok[b] == okr and np.array_equal(us[b], ref), b
        assert cnt[b, 0] == (st.nst if solver == "lsoda" else st.nstep)
    # mixed: array rtol, scalar atol; warp kernel
    Bw = 6
    Pw = pr.brusselator_sweep(Bw, seed=2)
    if solver == "dop853":
        Pw[:, 2] = 0.5
    u0w = pr.brusselator_u0()
    tew = np.linspace(0.0, 2.0, 9)
    rtw = 10.0 ** rng.uniform(-7, -4, Bw)
    us, ok = api(nb.builtin("brusselator1d"), u0w, tew, Pw, rtw, 1e-9, strict=True)
    fw = orc.builtin_rhs("brusselator1d")
    for b in range(Bw):
        ref, okr, st = ofun(fw, u0w, tew, Pw[b], rtw[b], 1e-9)
        assert ok[b] == okr and np.array_equal(us[b], ref), b
    with pytest.raises(ValueError):
        api(nb.builtin("lotka_volterra"), u0, te, P, rt[:5], 1e-8)


def test_concurrent_calls_from_host_threads():
    """The reference's wrappers are reentrant (safe from prange); so are the
    library's entry points: several host threads solving different ensembles on the same
    handle and device at once get the results of serial calls."""
    import threading
    nb = _nb()
    u0 = np.array([1.0, 0.0, 0.0])
    te = np.linspace(0, 1e5, 50)
    Ps = [pr.robertson_sweep(3000, seed=s) for s in range(4)]
    serial = [nb.lsoda(nb.builtin("robertson"), u0, te, P, 1e-8, 1e-8, strict=True)[0] for P in Ps]
    out = [None] * 4

    def work(k):
        out[k] = nb.lsoda(nb.builtin("robertson"), u0, te, Ps[k], 1e-8, 1e-8, strict=True)[0]

    th = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    for k in range(4):
        assert np.array_equal(out[k], serial[k]), k
    # the warp kernel keeps nothing per launch outside its shared memory either
    Pw = pr.brusselator_sweep(40, seed=5)
    u0w, tew = pr.brusselator_u0(), np.linspace(0, 2.0, 5)
    ser = nb.lsoda(nb.builtin("brusselator1d"), u0w, tew, Pw, 1e-6, 1e-8)[0]
    outw = [None] * 3

    def workw(k):
        outw[k] = nb.lsoda(nb.builtin("brusselator1d"), u0w, tew, Pw, 1e-6, 1e-8)[0]

    th = [threading.Thread(target=workw, args=(k,)) for k in range(3)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert all(np.array_equal(o, ser) for o in outw)


@pytest.mark.parametrize("n", [1, 4, 5, 6, 7, 8])
def test_every_thread_kernel_size(n):
    """One kernel image per system size n = 1..8 (both solvers): each is run on a stiff
    nonlinear chain (numba RHS with a loop) and, for LSODA, on the pivoting problem."""
    import numba
    nb = _nb()
    B = 9
    if n == 1:
        def rhs(t, u, du, p):
            # (polynomial: a transcendental in the RHS is libdevice on the GPU and glibc on
            # the CPU, which are not bit-identical -- SURVEY.md appendix A)
            du[0] = -p[0] * u[0] + 0.1 * t * t - u[0] * u[0] * u[0]
        P = np.array([[0.5 + 50.0 * b] for b in range(B)])
        cases = [(rhs, P, np.array([1.0]), np.linspace(0, 5, 11))]
    else:
        P = np.array([[20.0 * (1 + b), float(n)] for b in range(B)])
        Pu = np.array([[1000.0 * (1 + 0.1 * b), 5000.0, float(n)] for b in range(B)])
        cases = [(pr.chain_rhs, P, 1.0 + 0.1 * np.arange(n), np.linspace(0, 5, 11)),
                 (pr.upper_rhs, Pu, np.ones(n), np.linspace(0, 2, 9))]
    for rhs, Pk, u0, te in cases:
        cf = numba.cfunc(nb.lsoda_sig)(rhs)
        us, ok, cnt = nb.lsoda(rhs, u0, te, Pk, 1e-7, 1e-9, strict=True, counters=True)
        ref, okr, cr = orc.solve_batch("lsoda", cf.address, u0, te, Pk, 1e-7, 1e-9)
        assert list(ok) == list(okr) and np.array_equal(cnt, cr), (n, rhs.__name__)
        assert np.array_equal(us, ref)
        if rhs is not pr.upper_rhs:
            us, ok, cnt = nb.dop853(rhs, u0, te, Pk, 1e-7, 1e-9, mxstep=200000, strict=True, counters=True)
            ref, okr, cr = orc.solve_batch("dop853", cf.address, u0, te, Pk, 1e-7, 1e-9, mxstep=200000)
            assert list(ok) == list(okr) and np.array_equal(cnt[:, :7], cr[:, :7]), (n, "dop853")
            assert np.array_equal(us, ref)
