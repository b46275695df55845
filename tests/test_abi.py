"""The C-ABI library loads without a GPU and exports what include/b200ode.h declares;
linking a right-hand side (nvJitLink, CPU-only work) produces an sm_100a cubin whose
solver state lives in registers; solving without a device fails loudly."""
import ctypes as ct
import os
import re
import subprocess

import numpy as np
import pytest

import _problems as pr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200ode.h")).read()
    return sorted(set(re.findall(r"\b(b200ode_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from numbalsoda_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(_lib.lib, n), n
    _lib.lib.b200ode_version.restype = ct.c_int
    assert _lib.lib.b200ode_version() == 101  # include/b200ode.h B200ODE_VERSION


def test_link_errors_are_reported():
    from numbalsoda_b200 import _lib
    out = ct.c_void_p()
    rc = _lib.lib.b200ode_link_rhs(0, 3, b"nope", 4, _lib.IMG_BUILTIN, 0, ct.byref(out))
    assert rc == -1 and b"unknown built-in" in _lib.lib.b200ode_last_error()
    rc = _lib.lib.b200ode_link_rhs(0, 4096, b"lorenz", 6, _lib.IMG_BUILTIN, 0, ct.byref(out))
    assert rc == -5
    rc = _lib.lib.b200ode_link_rhs(0, 3, b"garbage", 7, _lib.IMG_LTOIR, 0, ct.byref(out))
    assert rc == -2 and not out.value


@pytest.mark.parametrize("solver", ["dop853", "lsoda"])
@pytest.mark.parametrize("strict", [False, True])
def test_link_builtin_and_numba_rhs(solver, strict):
    import numbalsoda_b200 as nb
    from numbalsoda_b200 import _lib
    sid = _lib.DOP853 if solver == "dop853" else _lib.LSODA
    for rhs in (nb.builtin("lorenz"), pr.lorenz_rhs):
        h = nb.get_handle(rhs, sid, 3, strict=strict)
        cubin = h.cubin()
        assert cubin[:4] == b"\x7fELF" and len(cubin) > 10000
        assert nb.get_handle(rhs, sid, 3, strict=strict) is h  # cached


@pytest.mark.parametrize("solver", ["dop853", "lsoda"])
@pytest.mark.parametrize("mode", ["fast", "exact_ctrl", "strict"])
def test_linked_kernel_keeps_state_in_registers(tmp_path, solver, mode):
    """cuobjdump on the linked cubin, every link mode: no local-memory spills, no stack, the
    user RHS inlined (no CALL to b200ode_user_rhs survives); the libm-exact pow (its tables)
    is linked out of the default mode's thread kernels.

    One measured exception: LSODA n = 3 runs four blocks of 128 threads per SM (128 registers,
    build.py MINBLOCKS).  The default mode fits without a spill; the two verification modes
    (strict, exact_ctrl: IEEE division and the double-precision pow as out-of-line calls) park
    up to three rarely used words (pnorm, the packed order counters) on the stack and are still
    faster than with three blocks at 161 registers (5.42 M against 5.13 M trajectories/s,
    profiles/r2_sweep_lsoda_occupancy.log)."""
    import numbalsoda_b200 as nb
    from numbalsoda_b200 import _lib
    h = nb.get_handle(pr.lorenz_rhs, _lib.DOP853 if solver == "dop853" else _lib.LSODA, 3,
                      strict=mode == "strict", exact_ctrl=mode == "exact_ctrl")
    f = tmp_path / "k.cubin"
    f.write_bytes(h.cubin())
    res = subprocess.run(["cuobjdump", "-res-usage", str(f)], capture_output=True, text=True).stdout
    m = re.search(r"Function b200ode_%s_kernel:\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)" % solver, res)
    assert m, res
    regs, stack, shared, local = map(int, m.groups())
    allowed = 32 if (solver == "lsoda" and mode != "fast") else 0
    assert stack <= allowed and local == 0 and regs <= 255
    sass = subprocess.run(["cuobjdump", "-sass", str(f)], capture_output=True, text=True).stdout
    assert "b200ode_user_rhs" not in sass
    assert "DFMA" in sass
    assert "b200ode_link_flags" not in sass  # the mode is a link-time constant


@pytest.mark.parametrize("solver", ["dop853", "lsoda"])
@pytest.mark.parametrize("mode", ["fast", "exact_ctrl", "strict"])
def test_warp_kernels_do_not_spill(tmp_path, solver, mode):
    """The warp-per-trajectory images (n = 9..128; dense and banded LSODA entry points, DOP853):
    every kernel of the linked cubin without stack or local memory."""
    import numbalsoda_b200 as nb
    from numbalsoda_b200 import _lib
    h = nb.get_handle(nb.builtin("brusselator1d_il"), _lib.DOP853 if solver == "dop853" else _lib.LSODA, 64,
                      strict=mode == "strict", exact_ctrl=mode == "exact_ctrl")
    f = tmp_path / "k.cubin"
    f.write_bytes(h.cubin())
    res = subprocess.run(["cuobjdump", "-res-usage", str(f)], capture_output=True, text=True).stdout
    found = re.findall(r"Function (b200ode_\w+_kernel\w*):\s*\n\s*REG:(\d+) STACK:(\d+) SHARED:\d+ LOCAL:(\d+)", res)
    assert found, res
    if solver == "lsoda":
        assert {k for k, *_ in found} >= {"b200ode_lsodaw_kernel", "b200ode_lsodaw_band_kernel",
                                          "b200ode_lsodaw_tmem_kernel"}
        # the tensor-memory kernel really is one: tcgen05 allocation, loads and stores in its SASS
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", "b200ode_lsodaw_tmem_kernel", str(f)],
                              capture_output=True, text=True).stdout
        assert "LDTM" in sass and "STTM" in sass and "UTCATOMSWS" in sass
    for name, regs, stack, local in found:
        assert int(stack) == 0 and int(local) == 0 and int(regs) <= 255, (name, regs, stack, local)


def test_cfunc_address_is_resolved():
    import numba
    import numbalsoda_b200 as nb
    from numbalsoda_b200 import _lib
    cf = numba.cfunc(nb.lsoda_sig)(pr.lv_rhs)
    h = nb.get_handle(cf.address, _lib.DOP853, 2)
    assert h.origin is pr.lv_rhs
    with pytest.raises(ValueError):
        nb.get_handle(0x1234, _lib.DOP853, 2)


def test_no_cpu_fallback():
    import numbalsoda_b200 as nb
    from numbalsoda_b200 import _lib
    if _lib.lib.b200ode_device_count() > 0:
        pytest.skip("a device is present")
    with pytest.raises(nb.B200OdeError, match="no CPU fallback"):
        nb.dop853(nb.builtin("lorenz"), np.array([1.0, 0, 0]), np.linspace(0, 1, 5),
                  np.array([10.0, 28.0, 8 / 3]))


def test_shard_bounds():
    from numbalsoda_b200._solve import shard_bounds
    for B in (0, 1, 7, 8, 1000, 2 ** 20 + 3):
        for parts in (1, 2, 3, 8):
            b = shard_bounds(B, parts)
            assert b[0] == 0 and b[-1] == B and len(b) == parts + 1
            sizes = np.diff(b)
            assert sizes.min() >= 0 and sizes.max() - sizes.min() <= 1
