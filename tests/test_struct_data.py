"""`data` records with embedded pointers: the reference's
address_as_void_pointer pattern (numbalsoda/driver.py:45-53, passing_data_to_rhs_function.ipynb),
in which the right-hand side reaches a further array through an address stored in its data
structure.  The notebook's wrapper -- a cfunc over CPointer(record) that views the record and the
array with nb.carray and calls an nb.njit function -- is accepted as it is written there (only
np.sum over the array becomes a loop: numba.cuda has no array reductions); for the device the
views are rewritten (numbalsoda_b200/rhs.py: without_carray), the njit helper becomes a device
function, the record is passed by reference, and `pointers={field: array}` stores the device
address of a copy of the array where the reference stores the host address."""
import numpy as np
import pytest

import _oracle as orc

nb = pytest.importorskip("numba")
from numba import types  # noqa: E402

import numbalsoda_b200 as b2  # noqa: E402
from numbalsoda_b200 import address_as_void_pointer, address_as_pointer, _lib  # noqa: E402

args_dtype = types.Record.make_c_struct([("p", types.float64), ("arr_p", types.int64), ("len_arr", types.int64)])
NPDT = np.dtype([("p", "f8"), ("arr_p", "i8"), ("len_arr", "i8")], align=True)


@nb.njit
def rhs(t, u, du, p, arr, n):
    s = 0.0
    for k in range(n):
        s += arr[k]
    du[0] = u[0] - u[0] * u[1] * s
    du[1] = u[0] * u[1] - u[1] * p


@nb.cfunc(types.void(types.double, types.CPointer(types.double), types.CPointer(types.double),
                     types.CPointer(args_dtype)))
def wrapped(t, u, du, user_data_p):
    # passing_data_to_rhs_function.ipynb, cell 4
    user_data = nb.carray(user_data_p, 1)
    p = user_data[0].p
    arr = nb.carray(address_as_void_pointer(user_data[0].arr_p), (user_data[0].len_arr), dtype=np.float64)
    rhs(t, u, du, p, arr, user_data[0].len_arr)


def device_dialect(t, u, du, ud):
    # the same right-hand side written directly for a record argument, without carray
    arr = address_as_pointer(ud.arr_p, nb.float64)
    s = 0.0
    for k in range(ud.len_arr):
        s += arr[k]
    du[0] = u[0] - u[0] * u[1] * s
    du[1] = u[0] * u[1] - u[1] * ud.p


ARR = 0.1 * np.ones(6)
U0, TE = np.array([5.0, 0.8]), np.linspace(0.0, 50.0, 1000)


def _args(p=2.0, arr=ARR):
    return np.array((p, arr.ctypes.data, arr.shape[0]), dtype=NPDT)


def test_links_for_the_device_without_a_gpu_and_runs_on_the_cpu():
    # the intrinsics are the reference's on the CPU (the cfunc above compiled and runs):
    ref, ok, st = orc.lsoda(wrapped.address, U0, TE, _args().reshape(1).view(np.uint8), 1e-3, 1e-6)
    assert ok and st.nst > 100 and np.isfinite(ref).all()
    for f in (wrapped.address, wrapped, device_dialect):
        h = b2.get_handle(f, _lib.LSODA, 2, data_dtype=NPDT)
        assert len(h.cubin()) > 1000
    # pointers= without a record is refused before anything touches a device
    with pytest.raises(ValueError):
        b2.lsoda(b2.builtin("lotka_volterra"), U0, TE, np.array([1.0]), pointers={"arr_p": ARR})


@pytest.mark.gpu
def test_notebook_pattern_on_the_gpu_bitwise():
    args = _args()
    ref, ok, st = orc.lsoda(wrapped.address, U0, TE, args.reshape(1).view(np.uint8), 1e-3, 1e-6)
    for f in (wrapped.address, device_dialect):
        us, okg, cnt = b2.lsoda(f, U0, TE, args, strict=True, counters=True, pointers={"arr_p": ARR})
        assert okg and ok and list(cnt)[:4] == [st.nst, st.nfe, st.nje, st.nswitch]
        assert np.array_equal(us, ref)
    # dop853, and an ensemble of records that share the array (one device copy)
    B = 50
    recs = np.zeros(B, dtype=NPDT)
    recs["p"] = np.linspace(0.5, 3.0, B)
    recs["arr_p"] = ARR.ctypes.data
    recs["len_arr"] = 6
    us, okg = b2.dop853(wrapped.address, U0, TE[:200], recs, 1e-8, 1e-8, strict=True, pointers={"arr_p": ARR})
    assert us.shape == (B, 200, 2) and okg.all()
    for b in (0, 17, 49):
        r, ok, _ = orc.dop853(wrapped.address, U0, TE[:200], recs[b:b + 1].view(np.uint8), 1e-8, 1e-8)
        assert ok and np.array_equal(us[b], r)
    # per-trajectory arrays: the caller places device copies itself
    arrs = [np.full(4, 0.1 + 0.01 * b) for b in range(B)]
    bufs = [b2.DeviceBuffer(a, 0) for a in arrs]
    recs2 = recs.copy()
    recs2["arr_p"] = [d.ptr for d in bufs]
    recs2["len_arr"] = 4
    us, okg = b2.lsoda(device_dialect, U0, TE[:100], recs2, 1e-8, 1e-8, strict=True)
    host = recs2.copy()
    host["arr_p"] = [a.ctypes.data for a in arrs]
    for b in (0, 31):
        r, ok, _ = orc.lsoda(wrapped.address, U0, TE[:100], host[b:b + 1].view(np.uint8), 1e-8, 1e-8)
        assert ok and okg[b] and np.array_equal(us[b], r)
