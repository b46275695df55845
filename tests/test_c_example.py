"""examples/robertson_ensemble.c: libb200ode used from plain C with a right-hand side that
nvcc compiled (an LTO fatbin, `B200ODE_IMG_FATBIN`) -- the path a C/C++ caller of the
reference's `lsoda_wrapper` (src/wrapper.cpp:10-57) takes; no Python on it.

CPU: the example builds against include/b200ode.h, the fatbin links (nvJitLink needs no GPU),
and the binary fails loudly without a device.  GPU: its strict-mode output equals the oracle's
bit for bit (states printed with %.17g, counters)."""
import ctypes as ct
import os
import shutil
import subprocess

import numpy as np
import pytest

import _oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "examples")
LIBDIR = os.path.join(ROOT, "numbalsoda_b200")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


@pytest.fixture(scope="module")
def built(tmp_path_factory):
    d = tmp_path_factory.mktemp("c_example")
    fatbin = str(d / "robertson_rhs.fatbin")
    exe = str(d / "robertson_ensemble")
    subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=lto_100a", "-rdc=true",
                           "-fatbin", "-o", fatbin, os.path.join(EX, "robertson_rhs.cu")])
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-o", exe, os.path.join(EX, "robertson_ensemble.c"),
                           "-L", LIBDIR, "-lb200ode", "-lm"])
    return fatbin, exe


def _run(exe, *args):
    env = dict(os.environ, LD_LIBRARY_PATH=LIBDIR + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    return subprocess.run([exe, *args], env=env, capture_output=True, text=True, timeout=600)


def test_fatbin_links_without_a_gpu(built):
    from numbalsoda_b200 import _lib
    image = open(built[0], "rb").read()
    for flags in (0, _lib.STRICT_FP):
        h = ct.c_void_p()
        rc = _lib.lib.b200ode_link_rhs(_lib.LSODA, 3, image, len(image), _lib.IMG_FATBIN, flags,
                                       ct.byref(h))
        assert rc == 0, _lib.lib.b200ode_last_error()
        assert h.value
        _lib.lib.b200ode_rhs_free(h)


def test_example_fails_loudly_without_a_device(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    r = _run(built[1], built[0], "4")
    assert r.returncode == 1 and "not available" in r.stderr


@pytest.mark.gpu
def test_example_matches_oracle(built):
    B = 257
    r = _run(built[1], built[0], str(B), "strict")
    assert r.returncode == 0, r.stderr
    rows = [ln.split() for ln in r.stdout.strip().splitlines()]
    assert len(rows) == B
    te = np.array([1.0e5 * i / 99 for i in range(100)])
    f = orc.builtin_rhs("robertson")
    for b in (0, 1, 100, 255, 256):
        s = b / (B - 1)
        data = np.array([0.04 * (0.5 + s), 3.0e7 * (1.5 - s), 1.0e4 * (0.75 + 0.5 * s)])
        us, ok, st = orc.lsoda(f, [1.0, 0.0, 0.0], te, data, rtol=1e-8, atol=1e-8)
        row = rows[b]
        assert int(row[0]) == b and int(row[1]) == int(ok) == 1
        assert [int(x) for x in row[2:5]] == [st.nst, st.nfe, st.nje]
        got = np.array([float(x) for x in row[6:9]])
        assert (got == us[-1]).all(), (b, got, us[-1])


# ---------------------------------------------------------------- solve_ivp from C
@pytest.fixture(scope="module")
def built_ivp(tmp_path_factory):
    d = tmp_path_factory.mktemp("c_example_ivp")
    src = os.path.join(EX, "lorenz_events.cu")
    rhs, ev, exe = str(d / "lorenz_rhs.fatbin"), str(d / "lorenz_events.fatbin"), str(d / "lorenz_events")
    for macro, out in (("LORENZ_RHS", rhs), ("LORENZ_EVENTS", ev)):
        subprocess.check_call([NVCC, "-gencode", "arch=compute_100a,code=lto_100a", "-rdc=true",
                               "-fatbin", "-D" + macro, "-o", out, src])
    subprocess.check_call(["gcc", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           "-o", exe, os.path.join(EX, "lorenz_events.c"), "-L", LIBDIR, "-lb200ode",
                           "-lm"])
    return rhs, ev, exe


def test_ivp_fatbins_link_without_a_gpu(built_ivp):
    from numbalsoda_b200 import _lib
    rhs, ev = (open(f, "rb").read() for f in built_ivp[:2])
    for flags in (0, _lib.STRICT_FP):
        h = ct.c_void_p()
        rc = _lib.lib.b200ode_link_ivp(3, rhs, len(rhs), _lib.IMG_FATBIN, 2, ev, len(ev), _lib.IMG_FATBIN,
                                       flags, ct.byref(h))
        assert rc == 0 and h.value, _lib.lib.b200ode_last_error()
        _lib.lib.b200ode_rhs_free(h)
    # argument errors of the link call
    h = ct.c_void_p()
    assert _lib.lib.b200ode_link_ivp(3, rhs, len(rhs), _lib.IMG_FATBIN, 9, ev, len(ev), _lib.IMG_FATBIN,
                                     0, ct.byref(h)) == -5  # more than 8 events
    assert _lib.lib.b200ode_link_ivp(3, rhs, len(rhs), _lib.IMG_FATBIN, 1, None, 0, 0, 0,
                                     ct.byref(h)) == -1
    assert _lib.lib.b200ode_link_ivp(129, rhs, len(rhs), _lib.IMG_FATBIN, 0, None, 0, 0, 0,
                                     ct.byref(h)) == -5


def test_ivp_example_fails_loudly_without_a_device(built_ivp):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    r = _run(built_ivp[2], built_ivp[0], built_ivp[1], "4")
    assert r.returncode == 1 and "not available" in r.stderr


@pytest.mark.gpu
def test_ivp_example_matches_oracle(built_ivp):
    from numba import cfunc, types
    sig = types.void(types.double, types.CPointer(types.double), types.CPointer(types.double),
                     types.CPointer(types.double))

    @cfunc(sig)
    def events(t, u, out, p):
        out[0] = u[2] - p[3]
        out[1] = u[0]

    B = 129
    r = _run(built_ivp[2], built_ivp[0], built_ivp[1], str(B), "strict")
    assert r.returncode == 0, r.stderr
    rows = [ln.split() for ln in r.stdout.strip().splitlines()]
    assert len(rows) == B
    f = orc.builtin_rhs("lorenz")
    seen = set()
    for b in (0, 1, 31, 64, 100, 127, 128):
        s = b / (B - 1)
        data = np.array([9.0 + 2.0 * s, 20.0 + 16.0 * s, 2.0 + 1.33 * s, 25.0 + 30.0 * (1.0 - s)])
        ref = orc.solve_ivp(f, [0.0, 5.0], [1.0, 0.0, 0.0], None, 2, events.address, data, rtol=1e-8,
                            atol=1e-8)
        row = rows[b]
        assert int(row[0]) == b
        assert [int(x) for x in row[1:7]] == [int(ref["success"]), orc.IVP_MESSAGES.index(ref["message"]),
                                              ref["nfev"], len(ref["t"]), int(ref["event_found"]),
                                              ref["ind_event"]]
        got = np.array([float(x) for x in row[7:]])
        want = np.concatenate([[ref["t_event"]], ref["y_event"], [ref["t"][-1]], ref["y"][-1]])
        assert (got == want).all(), (b, got, want)
        seen.add((ref["event_found"], ref["ind_event"]))
    assert len(seen) >= 2
