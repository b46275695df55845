"""Parity of the CUDA DOP853 path (through the C ABI / numbalsoda-compatible
driver) with the CPU oracle on the same inputs.  strict=True links the kernels
with FMA contraction disabled: the arithmetic is then the reference's, so states
are compared bit-for-bit and the step counters must be identical."""
import os

import numpy as np
import pytest

import _oracle as orc
import _problems as pr

pytestmark = pytest.mark.gpu

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "dop853_oracle.npz"))
CASES = ["lv_test", "lv_readme", "lorenz_short", "lorenz"]


def _nb():
    import numbalsoda_b200 as nb
    return nb


def _cnt(st):
    return [st.nstep, st.nfcn, st.naccpt, st.nrejct, st.nevent, st.idid, st.rows, 0]


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("kind", ["builtin", "numba"])
def test_single_trajectory_strict_bitwise(name, kind):
    nb = _nb()
    p = pr.problem(name)
    rhs = nb.builtin(p["builtin"]) if kind == "builtin" else p["rhs"]
    us, ok, cnt = nb.dop853(rhs, p["u0"], p["t_eval"], p["data"], p["rtol"], p["atol"],
                            strict=True, counters=True)
    ref, okr, st = orc.dop853(orc.builtin_rhs(p["builtin"]), p["u0"], p["t_eval"], p["data"],
                              p["rtol"], p["atol"])
    assert ok == okr
    assert list(cnt) == _cnt(st)
    assert np.array_equal(us, ref)
    assert np.array_equal(us, GOLD[name + "/usol"])
    assert list(cnt[:7]) == list(GOLD[name + "/counters"])


def test_reference_known_answer():
    # /root/reference/tests/test_numbalsoda.py:22-31, through the drop-in driver
    nb = _nb()
    p = pr.problem("lv_test")
    import numba
    cf = numba.cfunc(nb.lsoda_sig)(p["rhs"])
    usol, success = nb.dop853(cf.address, p["u0"], p["t_eval"], p["data"], rtol=1e-8, atol=1e-8)
    assert success
    assert usol.shape == (11, 2)
    assert np.all(np.isclose(usol[-1], np.array([0.38583246250193476, 4.602012234037773])))


@pytest.mark.parametrize("strict", [True, False])
def test_lorenz_sweep_against_oracle(strict):
    """Config C2's parameter sweep (short horizon: Lorenz is chaotic)."""
    nb = _nb()
    B = 1000
    P = pr.lorenz_sweep(B)
    u0 = np.array([1.0, 0.0, 0.0])
    te = np.linspace(0, 5, 51)
    us, ok, cnt = nb.dop853(nb.builtin("lorenz"), u0, te, P, 1e-8, 1e-8, strict=strict,
                            counters=True)
    ref, okr, cr = orc.solve_batch("dop853", orc.builtin_rhs("lorenz"), u0, te, P, 1e-8, 1e-8)
    assert ok.all() and okr.all() and us.shape == (B, 51, 3)
    if strict:
        assert np.array_equal(cnt[:, :7], cr[:, :7])
        assert np.array_equal(us, ref)
    else:
        # fused multiply-adds and the throughput controller: north_star's bound is 10*rtol, taken
        # in the solver's own weights rtol*|y| + atol (no padding of the denominator);
        # accuracy against a tight-tolerance truth: tests/test_gpu_truth.py
        err = np.abs(us - ref) / (1e-8 * np.abs(ref) + 1e-8)
        print("fast-mode max weighted distance to the oracle: %.3g" % err.max())
        assert err.max() <= 10.0
        assert np.abs(cnt[:, 0].astype(float) - cr[:, 0]).max() <= 3


def test_batched_u0_and_shared_data():
    nb = _nb()
    rng = np.random.default_rng(5)
    B = 257
    u0 = np.array([5.0, 0.8]) * rng.uniform(0.5, 1.5, (B, 2))
    te = np.linspace(0, 10, 31)
    d = np.array([1.3])
    us, ok, cnt = nb.dop853(nb.builtin("lotka_volterra"), u0, te, d, 1e-7, 1e-9, strict=True,
                            counters=True)
    ref, okr, cr = orc.solve_batch("dop853", orc.builtin_rhs("lotka_volterra"), u0, te, d,
                                   1e-7, 1e-9)
    assert ok.all() and np.array_equal(us, ref) and np.array_equal(cnt[:, :7], cr[:, :7])


@pytest.mark.parametrize("nt", [2, 1001, 1100, 40000])
def test_output_grids_in_shared_and_global_memory(nt):
    """The kernel keeps a per-block copy of t_eval in shared memory when that costs no
    resident block (nt = 1001 of config C2 just fits) and reads global memory otherwise:
    both paths, all modes, many rows per step (nt = 40000: ~100 rows per step)."""
    nb = _nb()
    B = 96
    P = pr.lorenz_sweep(B)
    u0 = np.array([1.0, 0.0, 0.0])
    te = np.linspace(0, 3, nt)
    ref, okr, cr = orc.solve_batch("dop853", orc.builtin_rhs("lorenz"), u0, te, P, 1e-8, 1e-8)
    us, ok, cnt = nb.dop853(nb.builtin("lorenz"), u0, te, P, 1e-8, 1e-8, strict=True, counters=True)
    assert ok.all() and np.array_equal(us, ref) and np.array_equal(cnt[:, :7], cr[:, :7])
    for kw in (dict(), dict(exact_ctrl=True)):
        us, ok = nb.dop853(nb.builtin("lorenz"), u0, te, P, 1e-8, 1e-8, **kw)
        assert ok.all()
        assert (np.abs(us - ref) / (1e-8 * np.abs(ref) + 1e-8)).max() <= 10.0


def test_failures_and_edge_cases():
    nb = _nb()
    f = nb.builtin("lotka_volterra")
    fo = orc.builtin_rhs("lotka_volterra")
    u0 = np.array([5.0, 0.8])
    d = np.array([1.0])
    te = np.linspace(0, 10, 11)
    # nmax <= 0 (dop853_module.f90:243-246) and nmax exhausted (:539)
    for mx in (0, -3, 20, 51):
        us, ok, cnt = nb.dop853(f, u0, te, d, 1e-8, 1e-8, mxstep=mx, strict=True, counters=True)
        ref, okr, st = orc.dop853(fo, u0, te, d, 1e-8, 1e-8, mxstep=mx)
        assert ok == okr, mx
        assert cnt[5] == st.idid and cnt[6] == st.rows
        assert np.array_equal(us[:st.rows], ref[:st.rows])
    # a batch in which only some trajectories fail
    P = np.array([[1.0], [1.0], [1.0]])
    U = np.array([[5.0, 0.8], [50.0, 8.0], [5.0, 0.8]])
    # (needs 51 / 47 / 51 attempted steps: a budget of 48 lets only the middle one finish)
    us, ok, cnt = nb.dop853(f, U, te, P, 1e-8, 1e-8, mxstep=48, strict=True, counters=True)
    ref, okr, cr = orc.solve_batch("dop853", fo, U, te, P, 1e-8, 1e-8, mxstep=48)
    assert list(ok) == list(okr) and not ok.all() and ok.any()
    for b in range(3):
        r = cr[b, 6]
        assert np.array_equal(us[b, :r], ref[b, :r])
    # nt = 1 -> h = 0 -> idid = -3; decreasing t_eval -> success without rows
    us, ok, cnt = nb.dop853(f, u0, np.array([0.0]), d, 1e-8, 1e-8, counters=True)
    assert not ok and cnt[5] == -3
    us, ok, cnt = nb.dop853(f, u0, np.linspace(10, 0, 5), d, 1e-8, 1e-8, counters=True)
    assert ok and cnt[6] == 0
    # rows inside one step and duplicates
    te2 = np.array([0.0, 1e-4, 1e-4, 2e-4, 1.0])
    us, ok = nb.dop853(f, u0, te2, d, 1e-8, 1e-8, strict=True)
    ref, _, _ = orc.dop853(fo, u0, te2, d, 1e-8, 1e-8)
    assert ok and np.array_equal(us, ref)


def test_overflowing_trial_step_recovers():
    nb = _nb()
    te = np.linspace(0, 40, 5)
    d = np.array([0.04, 3e7, 1e4])
    us, ok, cnt = nb.dop853(nb.builtin("robertson"), np.array([1.0, 0, 0]), te, d, 1e-8, 1e-8,
                            mxstep=100000, strict=True, counters=True)
    ref, okr, st = orc.dop853(orc.builtin_rhs("robertson"), np.array([1.0, 0, 0]), te, d, 1e-8,
                              1e-8, mxstep=100000)
    assert ok and okr and list(cnt) == _cnt(st) and np.array_equal(us, ref)


def test_full_size_properties():
    """BASELINE config C2 at reduced B but full horizon: every trajectory succeeds,
    row 0 is u0 exactly, rows are finite, and an ensemble solved in one launch equals
    the same trajectories solved in two launches (order / lane independence)."""
    nb = _nb()
    B = 4096
    P = pr.lorenz_sweep(B)
    u0 = np.array([1.0, 0.0, 0.0])
    te = np.linspace(0, 100, 1001)
    us, ok = nb.dop853(nb.builtin("lorenz"), u0, te, P, 1e-8, 1e-8)
    assert ok.all() and np.isfinite(us).all()
    assert np.array_equal(us[:, 0, :], np.broadcast_to(u0, (B, 3)))
    perm = np.random.default_rng(0).permutation(B)
    us2, ok2 = nb.dop853(nb.builtin("lorenz"), u0, te, P[perm], 1e-8, 1e-8)
    assert np.array_equal(us2, us[perm])
    ref, okr, st = orc.dop853(orc.builtin_rhs("lorenz"), u0, te, P[17], 1e-8, 1e-8)
    # fast mode on a chaotic system: compare the early part only
    assert np.abs(us[17, :100] - ref[:100]).max() < 1e-6


def test_device_resident_entry_point():
    """b200ode_dop853_batched_device on torch-owned device memory and stream."""
    import ctypes as ct
    import torch
    nb = _nb()
    from numbalsoda_b200 import _lib
    B, nt = 512, 41
    P = torch.tensor(pr.lorenz_sweep(B), device="cuda")
    u0 = torch.tensor([1.0, 0.0, 0.0], dtype=torch.float64, device="cuda")
    te = torch.linspace(0, 4, nt, dtype=torch.float64, device="cuda")
    usol = torch.empty((B, nt, 3), dtype=torch.float64, device="cuda")
    ok = torch.zeros(B, dtype=torch.int32, device="cuda")
    h = nb.get_handle(nb.builtin("lorenz"), _lib.DOP853, 3, strict=True)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        rc = _lib.lib.b200ode_dop853_batched_device(
            h.ptr, B, 3, u0.data_ptr(), 0, P.data_ptr(), 24, 3, nt, te.data_ptr(),
            usol.data_ptr(), 1e-8, 1e-8, 10000, ok.data_ptr(), None, ct.c_void_p(s.cuda_stream))
    _lib.check(rc)
    s.synchronize()
    ref, okr, _ = orc.solve_batch("dop853", orc.builtin_rhs("lorenz"), np.array([1.0, 0, 0]),
                                  te.cpu().numpy(), P.cpu().numpy(), 1e-8, 1e-8)
    assert bool((ok == 1).all()) and np.array_equal(usol.cpu().numpy(), ref)
    info = h.kernel_info()
    assert info["local_bytes"] == 0 and info["launches"] >= 1


def test_shared_record_and_strides_on_both_paths():
    """include/b200ode.h: data_stride < 0 = one record of -data_stride bytes shared by all (what
    the Python driver passes for a 1-D `data`), data_stride = 0 the same with np doubles; u0 rows
    may be wider than neq.  Device-resident and host entry points alike."""
    import ctypes as ct
    import torch
    nb = _nb()
    from numbalsoda_b200 import _lib
    B, nt, n = 300, 21, 3
    rng = np.random.default_rng(11)
    U = np.zeros((B, 5))  # rows padded to 5 doubles; the padding must never be read as state
    U[:, :3] = [1.0, 0.0, 0.0] + 0.1 * rng.standard_normal((B, 3))
    U[:, 3:] = np.nan
    p = np.array([10.0, 28.0, 8.0 / 3.0])
    teh = np.linspace(0, 2, nt)
    ref, okr, _ = orc.solve_batch("dop853", orc.builtin_rhs("lorenz"), np.ascontiguousarray(U[:, :3]),
                                  teh, p, 1e-8, 1e-8)
    h = nb.get_handle(nb.builtin("lorenz"), _lib.DOP853, n, strict=True)
    # host path
    for stride in (-24, 0):
        us = np.full((B, nt, n), np.nan)
        ok = np.zeros(B, np.int32)
        _lib.check(_lib.lib.b200ode_dop853_batched(
            h.ptr, B, n, U.ctypes.data, 5, p.ctypes.data, stride, 3, nt, teh.ctypes.data,
            us.ctypes.data, 1e-8, 1e-8, 10000, ok.ctypes.data, None, 0))
        assert (ok == 1).all() and np.array_equal(us, ref), stride
    # device path
    Ud, pd = torch.tensor(U, device="cuda"), torch.tensor(p, device="cuda")
    te = torch.tensor(teh, device="cuda")
    for stride in (-24, 0):
        usol = torch.full((B, nt, n), float("nan"), dtype=torch.float64, device="cuda")
        ok = torch.zeros(B, dtype=torch.int32, device="cuda")
        _lib.check(_lib.lib.b200ode_dop853_batched_device(
            h.ptr, B, n, Ud.data_ptr(), 5, pd.data_ptr(), stride, 3, nt, te.data_ptr(),
            usol.data_ptr(), 1e-8, 1e-8, 10000, ok.data_ptr(), None, None))
        torch.cuda.synchronize()
        assert bool((ok == 1).all()) and np.array_equal(usol.cpu().numpy(), ref), stride
    # a stride narrower than a row is refused
    rc = _lib.lib.b200ode_dop853_batched_device(
        h.ptr, B, n, Ud.data_ptr(), 2, pd.data_ptr(), 0, 3, nt, te.data_ptr(),
        usol.data_ptr(), 1e-8, 1e-8, 10000, ok.data_ptr(), None, None)
    assert rc == -1  # B200ODE_EINVAL


def test_host_path_leaves_the_callers_device_current():
    """lsoda(..., device=k) must not change the CUDA runtime's (torch's) current device."""
    import torch
    nb = _nb()
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    torch.cuda.set_device(0)
    torch.zeros(1, device="cuda")
    us, ok = nb.dop853(nb.builtin("lorenz"), np.array([1.0, 0, 0]), np.linspace(0, 1, 5),
                       pr.lorenz_sweep(64), 1e-8, 1e-8, device=1)
    assert ok.all() and torch.cuda.current_device() == 0
    x = torch.ones(4, device="cuda")
    assert x.device.index == 0 and float(x.sum()) == 4.0


def test_reference_benchmark_rhs_verbatim_with_carray():
    """benchmark/benchmark_lorenz.py:8-16 of the reference, verbatim (it views the raw
    pointers with nb.carray): must compile for the device and reproduce the CPU cfunc."""
    import numba as nb_
    from numba import types
    nb = _nb()

    def f_(t, u_, du_, p_):
        u = nb_.carray(u_, (3,))
        p = nb_.carray(p_, (3,))
        sigma, rho, beta = p
        x, y, z = u
        du_[0] = sigma * (y - x)
        du_[1] = x * (rho - z) - y
        du_[2] = x * y - beta * z

    sig = types.void(types.double, types.CPointer(types.double), types.CPointer(types.double),
                     types.CPointer(types.double))
    cf = nb_.cfunc(sig)(f_)
    u0 = np.array([1.0, 0.0, 0.0])
    te = np.linspace(0.0, 10.0, 101)
    P = pr.lorenz_sweep(33)
    us, ok = nb.dop853(cf.address, u0, te, P, 1e-8, 1e-8, strict=True)
    ref, okr, _ = orc.solve_batch("dop853", cf.address, u0, te, P, 1e-8, 1e-8)
    assert ok.all() and okr.all() and np.array_equal(us, ref)
    us, ok = nb.lsoda(cf.address, u0, te, P, 1e-8, 1e-8, strict=True)
    ref, okr, _ = orc.solve_batch("lsoda", cf.address, u0, te, P, 1e-8, 1e-8)
    assert ok.all() and okr.all() and np.array_equal(us, ref)


# ---------------------------------------------------------------- warp-per-trajectory kernel
def test_warp_kernel_n64_strict_bitwise_and_fast():
    """Systems of more than 8 equations run one trajectory per warp (dop853_warp_kernel.cu):
    n = 64 Brusselator with mild diffusion (non-stiff), states and counters bit for bit."""
    nb = _nb()
    B = 80
    rng = np.random.default_rng(3)
    P = np.stack([rng.uniform(0.8, 1.2, B), rng.uniform(2.5, 3.5, B), 10.0 ** rng.uniform(-1, 0.7, B)], axis=1)
    u0 = np.stack([pr.brusselator_u0(P[b, 0], P[b, 1]) for b in range(B)])
    te = np.linspace(0.0, 5.0, 51)
    us, ok, cnt = nb.dop853(nb.builtin("brusselator1d"), u0, te, P, 1e-7, 1e-9, strict=True, counters=True)
    ref, okr, cr = orc.solve_batch("dop853", orc.builtin_rhs("brusselator1d"), u0, te, P, 1e-7, 1e-9)
    assert ok.all() and okr.all() and us.shape == (B, 51, 64)
    assert np.array_equal(cnt[:, :7], cr[:, :7]) and np.array_equal(us, ref)
    us2, ok2 = nb.dop853(nb.builtin("brusselator1d"), u0, te, P, 1e-7, 1e-9)
    assert ok2.all() and np.abs(us2 - ref).max() <= 1e-6 * np.abs(ref).max()


def test_warp_kernel_numba_rhs_sizes_and_edge_cases():
    import numba
    nb = _nb()

    def chain(t, u, du, p):
        n = int(p[1])
        k = p[0]
        du[0] = -k * u[0] + 0.5 * u[1] * u[1]
        for i in range(1, n - 1):
            du[i] = k * (u[i - 1] - u[i]) / (1.0 + i) + 0.1 * u[i + 1] - 0.1 * u[i] * u[i]
        du[n - 1] = k * (u[n - 2] - u[n - 1]) / n

    cf = numba.cfunc(nb.lsoda_sig)(chain)
    for n in (9, 20, 33, 128):  # ... up to the largest system the warp kernels take
        u0n = 1.0 + 0.1 * np.arange(n)
        Pn = np.array([[2.0 * (1 + b), float(n)] for b in range(5)])
        ten = np.linspace(0.0, 5.0, 11)
        us, ok, cnt = nb.dop853(chain, u0n, ten, Pn, 1e-7, 1e-9, strict=True, counters=True)
        ref, okr, cr = orc.solve_batch("dop853", cf.address, u0n, ten, Pn, 1e-7, 1e-9)
        assert list(ok) == list(okr) and np.array_equal(cnt[:, :7], cr[:, :7]) and np.array_equal(us, ref), n
    f, fo = nb.builtin("brusselator1d"), orc.builtin_rhs("brusselator1d")
    u0, d, te = pr.brusselator_u0(), np.array([1.0, 3.0, 0.5]), np.linspace(0, 2, 21)
    for kw in (dict(mxstep=0), dict(mxstep=7), dict(te=np.array([0.0])), dict(te=np.linspace(2, 0, 5)),
               dict(te=np.array([0.0, 1e-4, 1e-4, 2e-4, 0.5]))):
        te2 = kw.pop("te", te)
        us, ok, cnt = nb.dop853(f, u0, te2, d, 1e-7, 1e-9, strict=True, counters=True, **kw)
        ref, okr, st = orc.dop853(fo, u0, te2, d, 1e-7, 1e-9, **kw)
        assert ok == okr and cnt[5] == st.idid and cnt[6] == st.rows, (kw, list(cnt))
        assert np.array_equal(us[:st.rows], ref[:st.rows])
