"""device=[...] / device="all": the ensemble is sharded by trajectory, one host thread per device
entry, no exchange between shards.  On a one-GPU box the list names the same
device several times -- the shards then run concurrently on it -- and with more devices present
"all" is exercised too; either way the result must equal the unsharded one bit for bit."""
import numpy as np
import pytest

import _problems as pr

pytestmark = pytest.mark.gpu


def _device_sets():
    from numbalsoda_b200 import _lib
    sets = [[0, 0, 0]]
    if _lib.lib.b200ode_device_count() > 1:
        sets.append("all")
    return sets


def test_lsoda_and_dop853_shards_equal_the_whole():
    import numbalsoda_b200 as nb
    B = 1001
    u0 = np.array([1.0, 0.0, 0.0])
    for api, rhs, P, te in ((nb.dop853, "lorenz", pr.lorenz_sweep(B), np.linspace(0, 3, 31)),
                            (nb.lsoda, "robertson", pr.robertson_sweep(B), np.linspace(0, 1e3, 20))):
        whole = api(nb.builtin(rhs), u0, te, P, 1e-8, 1e-8, strict=True, counters=True)
        for dev in _device_sets():
            part = api(nb.builtin(rhs), u0, te, P, 1e-8, 1e-8, strict=True, counters=True, device=dev)
            for a, b in zip(whole, part):
                assert np.array_equal(a, b)


def test_warp_kernels_shards_equal_the_whole():
    """n = 64: the dense path keeps its matrices in tensor memory (one block of eight warps per
    SM allocates all 512 columns) and the banded path in shared memory; shards of one ensemble
    run concurrently, on one device or several."""
    import numbalsoda_b200 as nb
    B = 37
    P = pr.brusselator_sweep(B)
    te = np.linspace(0.0, 10.0, 25)
    for rhs, u0, kw in (("brusselator1d", pr.brusselator_u0(), {}),
                        ("brusselator1d_il", pr.interleave(pr.brusselator_u0()), dict(ml=2, mu=2))):
        whole = nb.lsoda(nb.builtin(rhs), u0, te, P, 1e-6, 1e-8, strict=True, counters=True, **kw)
        assert whole[1].all()
        for dev in _device_sets():
            part = nb.lsoda(nb.builtin(rhs), u0, te, P, 1e-6, 1e-8, strict=True, counters=True, device=dev, **kw)
            for a, b in zip(whole, part):
                assert np.array_equal(a, b)


@pytest.mark.parametrize("mode", ["t_eval", "two_pass", "max_rows"])
def test_solve_ivp_shards_equal_the_whole(mode):
    import numbalsoda_b200 as nb
    rng = np.random.default_rng(6)
    B = 1001
    data = np.column_stack([pr.lorenz_sweep(B), rng.uniform(30, 45, B)])
    ts = np.column_stack([np.zeros(B), rng.uniform(0.3, 3.0, B)])
    y0 = np.array([1.0, 0.0, 0.0])
    kw = dict(n_events=1, event_fcn=nb.builtin("lorenz_z"), data=data, rtol=1e-7, atol=1e-9, strict=True)
    if mode == "t_eval":
        kw["t_eval"] = np.linspace(0.0, 3.0, 16)
    if mode == "max_rows":
        kw["max_rows"] = 60
    whole = nb.solve_ivp(nb.builtin("lorenz"), ts, y0, **kw)
    for dev in _device_sets():
        part = nb.solve_ivp(nb.builtin("lorenz"), ts, y0, device=dev, **kw)
        assert np.array_equal(whole._r, part._r) and np.array_equal(whole.t_event, part.t_event)
        assert np.array_equal(whole.y_event, part.y_event)
        if mode == "t_eval":
            for b in range(B):  # rows beyond a trajectory's span are uninitialised memory
                k = min(whole.nt_out[b], np.searchsorted(kw["t_eval"], ts[b, 1], side="right"))
                assert np.array_equal(whole.y[b, :k], part.y[b, :k])
        else:
            assert np.array_equal(whole.row_offset, part.row_offset)
            for b in range(0, B, 13):
                assert np.array_equal(whole[b].t, part[b].t) and np.array_equal(whole[b].y, part[b].y)
