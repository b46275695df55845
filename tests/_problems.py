"""Problem definitions shared by the tests: the reference's own examples.

Each problem has the Python right-hand side exactly as the reference writes it
(so numba produces the same IEEE operations for the CPU cfunc and, through
numba.cuda, for the device function), the initial state, parameters and output
grid.  Sources: README.md:26-37, tests/test_numbalsoda.py:5-20,
benchmark/benchmark_lorenz.py:9-31, benchmark/benchmark_robber.py:8-32,
tests/test_exit_on_warning.py:12-30 (all under /root/reference)."""
import numpy as np


def lv_rhs(t, u, du, p):
    du[0] = u[0] - u[0] * u[1]
    du[1] = u[0] * u[1] - u[1] * p[0]


def lorenz_rhs(t, u, du, p):
    sigma = p[0]
    rho = p[1]
    beta = p[2]
    x = u[0]
    y = u[1]
    z = u[2]
    du[0] = sigma * (y - x)
    du[1] = x * (rho - z) - y
    du[2] = x * y - beta * z


def robertson_rhs(t, u, du, p):
    k1 = p[0]
    k2 = p[1]
    k3 = p[2]
    du[0] = -k1 * u[0] + k3 * u[1] * u[2]
    du[1] = k1 * u[0] - k2 * u[1] ** 2 - k3 * u[1] * u[2]
    du[2] = k2 * u[1] ** 2


def brusselator_rhs(t, u, du, p):
    # this repository's n = 64 method-of-lines problem (oracle/rhs_builtin.c, same operation order)
    A = p[0]
    B = p[1]
    D = p[2]
    m = 32
    for i in range(m):
        uu = u[i]
        v = u[m + i]
        ul = A if i == 0 else u[i - 1]
        ur = A if i == m - 1 else u[i + 1]
        vl = B if i == 0 else u[m + i - 1]
        vr = B if i == m - 1 else u[m + i + 1]
        uuv = uu * uu * v
        du[i] = A + uuv - (B + 1.0) * uu + D * (ul - 2.0 * uu + ur)
        du[m + i] = B * uu - uuv + D * (vl - 2.0 * v + vr)


def brusselator_il_rhs(t, u, du, p):
    # the same problem with the species interleaved (u_0, v_0, u_1, v_1, ...): banded, ml = mu = 2
    A = p[0]
    B = p[1]
    D = p[2]
    m = 32
    for i in range(m):
        uu = u[2 * i]
        v = u[2 * i + 1]
        ul = A if i == 0 else u[2 * i - 2]
        ur = A if i == m - 1 else u[2 * i + 2]
        vl = B if i == 0 else u[2 * i - 1]
        vr = B if i == m - 1 else u[2 * i + 3]
        uuv = uu * uu * v
        du[2 * i] = A + uuv - (B + 1.0) * uu + D * (ul - 2.0 * uu + ur)
        du[2 * i + 1] = B * uu - uuv + D * (vl - 2.0 * v + vr)


def brusselator_il_lanes(t, u, du, p, lane, nlanes):
    # brusselator_il_rhs in the lane-parallel form (lsoda(..., rhs_lanes=)): the cells lane,
    # lane + nlanes, ... -- the same operations per component
    A = p[0]
    B = p[1]
    D = p[2]
    m = 32
    for i in range(lane, m, nlanes):
        uu = u[2 * i]
        v = u[2 * i + 1]
        ul = A if i == 0 else u[2 * i - 2]
        ur = A if i == m - 1 else u[2 * i + 2]
        vl = B if i == 0 else u[2 * i - 1]
        vr = B if i == m - 1 else u[2 * i + 3]
        uuv = uu * uu * v
        du[2 * i] = A + uuv - (B + 1.0) * uu + D * (ul - 2.0 * uu + ur)
        du[2 * i + 1] = B * uu - uuv + D * (vl - 2.0 * v + vr)


def brusselator_lanes(t, u, du, p, lane, nlanes):
    # brusselator_rhs, lane-parallel
    A = p[0]
    B = p[1]
    D = p[2]
    m = 32
    for i in range(lane, m, nlanes):
        uu = u[i]
        v = u[m + i]
        ul = A if i == 0 else u[i - 1]
        ur = A if i == m - 1 else u[i + 1]
        vl = B if i == 0 else u[m + i - 1]
        vr = B if i == m - 1 else u[m + i + 1]
        uuv = uu * uu * v
        du[i] = A + uuv - (B + 1.0) * uu + D * (ul - 2.0 * uu + ur)
        du[m + i] = B * uu - uuv + D * (vl - 2.0 * v + vr)


def interleave(y):
    """[u_0..u_31, v_0..v_31] (last axis) -> [u_0, v_0, u_1, v_1, ...]"""
    y = np.asarray(y)
    m = y.shape[-1] // 2
    out = np.empty_like(y)
    out[..., 0::2] = y[..., :m]
    out[..., 1::2] = y[..., m:]
    return out


def growth_rhs(t, u, du, p):
    # tests/test_exit_on_warning.py:13-14
    du[0] = p[0] * u[0] * np.sin(t)


PROBLEMS = {
    # name: (rhs, builtin name in oracle/rhs_builtin.c, u0, data, t_eval, rtol, atol)
    "lv_test": (lv_rhs, "lotka_volterra", [5.0, 0.8], [1.0], np.linspace(0.0, 10.0, 11), 1e-8, 1e-8),
    "lv_readme": (lv_rhs, "lotka_volterra", [5.0, 0.8], [1.0], np.linspace(0.0, 50.0, 1000), 1e-3, 1e-6),
    "robertson": (robertson_rhs, "robertson", [1.0, 0.0, 0.0], [0.04, 3e7, 1e4],
                  np.linspace(0.0, 1e5, 100), 1e-8, 1e-8),
    "lorenz": (lorenz_rhs, "lorenz", [1.0, 0.0, 0.0], [10.0, 28.0, 8.0 / 3.0],
               np.linspace(0.0, 100.0, 1001), 1e-8, 1e-8),
    "lorenz_short": (lorenz_rhs, "lorenz", [1.0, 0.0, 0.0], [10.0, 28.0, 8.0 / 3.0],
                     np.linspace(0.0, 10.0, 101), 1e-8, 1e-8),
}


def problem(name):
    rhs, builtin, u0, data, te, rtol, atol = PROBLEMS[name]
    return dict(rhs=rhs, builtin=builtin, u0=np.array(u0, dtype=np.float64),
                data=np.array(data, dtype=np.float64), t_eval=np.asarray(te, dtype=np.float64),
                rtol=rtol, atol=atol)


_cfuncs = {}


def cfunc_address(rhs):
    """numba @cfunc(lsoda_sig) address of a Python RHS (kept alive in a cache)."""
    import numba as nb
    from numba import types
    if rhs not in _cfuncs:
        sig = types.void(types.double, types.CPointer(types.double),
                         types.CPointer(types.double), types.CPointer(types.double))
        _cfuncs[rhs] = nb.cfunc(sig)(rhs)
    return _cfuncs[rhs].address


def lorenz_sweep(B, seed=0):
    """sigma~U(9,11), rho~U(20,36), beta~U(2,3.33)."""
    rng = np.random.default_rng(seed)
    p = np.empty((B, 3))
    p[:, 0] = rng.uniform(9.0, 11.0, B)
    p[:, 1] = rng.uniform(20.0, 36.0, B)
    p[:, 2] = rng.uniform(2.0, 3.33, B)
    return p


def robertson_sweep(B, seed=0):
    """[0.04, 3e7, 1e4] * 10^U(-0.5, 0.5)."""
    rng = np.random.default_rng(seed)
    return np.array([0.04, 3e7, 1e4]) * 10.0 ** rng.uniform(-0.5, 0.5, (B, 3))


def brusselator_sweep(B, seed=0):
    """Config C4 (no reference shape exists; defined in DESIGN.md): 1-D Brusselator on 32
    cells, n = 64, data = [A, B, D] with A~U(0.8,1.2), B~U(2.5,3.5), D = 10^U(1.5,3.5) (stiff:
    LSODA switches to BDF, 15-25 Jacobians of 64 RHS evaluations each)."""
    rng = np.random.default_rng(seed)
    p = np.empty((B, 3))
    p[:, 0] = rng.uniform(0.8, 1.2, B)
    p[:, 1] = rng.uniform(2.5, 3.5, B)
    p[:, 2] = 10.0 ** rng.uniform(1.5, 3.5, B)
    return p


def brusselator_u0(A=1.0, Bp=3.0, m=32):
    x = (np.arange(m) + 1) / (m + 1.0)
    return np.concatenate([A + 0.5 * np.sin(np.pi * x), Bp / A + 0.5 * np.sin(2 * np.pi * x)])


def upper_rhs(t, u, du, p):
    """Stiff upper-bidiagonal system whose iteration matrix has super-diagonal entries
    larger than the diagonal: the reference's row-wise pivot search (LSODA.cpp:59-77,
    :188-224) then really interchanges columns.  p = [a, K, n]."""
    n = int(p[2])
    for i in range(n - 1):
        du[i] = -p[0] * u[i] + p[1] * u[i + 1]
    du[n - 1] = -p[0] * u[n - 1] + 1.0


def chain_rhs(t, u, du, p):
    """A nonlinear reaction chain of any size n = int(p[1]) (stiff for large k = p[0]):
    exercises every system size the kernels are built for."""
    n = int(p[1])
    k = p[0]
    du[0] = -k * u[0] + 0.5 * u[1] * u[1]
    for i in range(1, n - 1):
        du[i] = k * (u[i - 1] - u[i]) / (1.0 + i) + 0.1 * u[i + 1] - 0.1 * u[i] * u[i]
    du[n - 1] = k * (u[n - 2] - u[n - 1]) / n
