#!/usr/bin/env python
"""Benchmark of the batched-ODE hot path (BASELINE.json metric: trajectories/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload W] [--impl reference]

A "step" is one pass of the solver over one ensemble of synthetic inputs:
  lorenz_dop853   (default; BASELINE.json configs[1], SURVEY.md 8d "C2")
                  Lorenz n=3 parameter sweep, DOP853, t_eval = linspace(0,100,1001),
                  rtol = atol = 1e-8, 2^20 trajectories per GPU
  robertson_lsoda (configs[2], "C3") Robertson n=3, LSODA, t_eval = linspace(0,1e5,100),
                  rtol = atol = 1e-8, 2^20 trajectories per GPU
  brusselator_lsoda (configs[3], "C4") 1-D Brusselator n=64, LSODA (warp per trajectory),
                  t_eval = linspace(0,10,100), rtol = 1e-6, atol = 1e-8, 100 000 trajectories
  mixed           (configs[4], "C5") Lorenz + Robertson as homogeneous launches per GPU.
                  With --total T (BASELINE: 10 000 000) the total is FIXED and sharded over the
                  ranks (strong scaling), T/2 of each kind, in launches of <= 2^20 trajectories;
                  without it 2^20 + 2^20 per GPU (weak scaling).
Multi-GPU: one process per GPU (torchrun), the ensemble is sharded by trajectory, no
data-path collective; `value` = trajectories of all ranks / max-over-ranks time.

Prints ONE JSON line (rank 0).  The default line measures everything BASELINE's metric names:

  value / roofline / e2e     Lorenz DOP853 in the default link mode, with the right-hand side
                             compiled from the Python function by numba.cuda and linked by
                             nvJitLink (north_star's RHS path); `rhs_paths` has the nvcc-built
                             right-hand side shipped in the library beside it
  modes                      the same workload in all three link modes -- strict (bit-exact),
                             exact_ctrl, fast -- each with value, roofline fraction and a measured
                             parity statistic against the CPU oracle (checker use only)
  secondary.robertson_lsoda  the same block for Robertson LSODA
  cpu_baseline               the reference's CPU path on the host cores (N = 1 only)

`value` is measured with inputs and outputs resident in HBM (CUDA events on the launch stream);
`e2e` is the same workload through the call a user makes -- numbalsoda_b200.lsoda / dop853 with
numpy inputs and page-locked `out=` result buffers (two, rotated): H2D, kernels and D2H of every
row inside the timed region, all 2^20 trajectories; `e2e.ceiling` is a plain concurrent
device-to-host copy of the same bytes on the same ranks (what the host side can take).
`--impl reference` times the reference's CPU implementation of the path on the host cores on a
bounded sample (oracle/_ref when it compiled -- LSODA --, else the oracle port -- DOP853, whose
reference is Fortran).
"""
import argparse
import ctypes as ct
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOADS = {
    "lorenz_dop853": dict(solver="dop853", rhs="lorenz", n=3, p=3, u0=[1.0, 0.0, 0.0],
                          t_eval=(0.0, 100.0, 1001), rtol=1e-8, atol=1e-8, mxstep=10000,
                          batch=1 << 20, config="BASELINE.json configs[1] (SURVEY 8d C2)",
                          parity_rows=51),  # chaotic: a truth exists over a short horizon only (t <= 5)
    "robertson_lsoda": dict(solver="lsoda", rhs="robertson", n=3, p=3, u0=[1.0, 0.0, 0.0],
                            t_eval=(0.0, 1.0e5, 100), rtol=1e-8, atol=1e-8, mxstep=10000,
                            batch=1 << 20, config="BASELINE.json configs[2] (SURVEY 8d C3)"),
    "brusselator_lsoda": dict(solver="lsoda", rhs="brusselator1d", n=64, p=3, u0="brusselator",
                              t_eval=(0.0, 10.0, 100), rtol=1e-6, atol=1e-8, mxstep=10000,
                              batch=100000, config="BASELINE.json configs[3] (SURVEY 8d C4): "
                              "1-D Brusselator on 32 cells, warp per trajectory"),
    # the same problem with the species interleaved (ml = mu = 2) and LSODA's
    # banded finite-difference Jacobian -- 5 right-hand-side calls and a band LU per Jacobian
    # instead of 64 and a dense one.  The reference has no banded path: its CPU arm runs the
    # dense solver on the same system, which is what a numbalsoda user has.
    "brusselator_lsoda_banded": dict(solver="lsoda", rhs="brusselator1d_il", n=64, p=3, u0="brusselator_il",
                                     t_eval=(0.0, 10.0, 100), rtol=1e-6, atol=1e-8, mxstep=10000, band=(2, 2),
                                     batch=100000, config="BASELINE.json configs[3] (SURVEY 8d C4) with "
                                     "lsoda(..., ml=2, mu=2) (SURVEY 8f rank 3): interleaved 1-D Brusselator, "
                                     "banded finite-difference Jacobian, warp per trajectory"),
    # the reference's third entry point on the C2 sweep, with the terminal
    # event z = 50 (45 % of the trajectories stop in the first transient, the others run to
    # t = 100: the persistent work queue is what keeps the lanes busy)
    "lorenz_solve_ivp": dict(solver="solve_ivp", rhs="lorenz", n=3, p=4, u0=[1.0, 0.0, 0.0],
                             t_eval=(0.0, 100.0, 1001), rtol=1e-8, atol=1e-8, mxstep=100000,
                             events="lorenz_z", level=50.0, batch=1 << 20,
                             config="SURVEY 8f rank 2: solve_ivp (DOP853, dense output every step, "
                                    "terminal event by Brent) on the sweep of BASELINE.json configs[1]"),
}

MIXED = ["lorenz_dop853", "robertson_lsoda"]  # BASELINE.json configs[4]
MODES = {"strict": dict(strict=True), "exact_ctrl": dict(exact_ctrl=True), "fast": dict()}
MODE_TEXT = {"strict": "strict (no FMA contraction: bit-exact)", "exact_ctrl": "FMA, exact controller",
             "fast": "fast (FMA, fast controller)"}


def sweep(name, B, seed, level=None):
    import _problems as pr
    if level is not None:  # solve_ivp: the event level as a 4th parameter
        return np.ascontiguousarray(np.column_stack([pr.lorenz_sweep(B, seed), np.full(B, level)]))
    if name == "lorenz":
        return pr.lorenz_sweep(B, seed)
    if name in ("brusselator1d", "brusselator1d_il"):
        return pr.brusselator_sweep(B, seed)
    return pr.robertson_sweep(B, seed)


def python_rhs(name):
    """The right-hand side as the Python function a numbalsoda user writes (tests/_problems.py,
    the reference's own benchmark functions); numba.cuda compiles it for the device."""
    import _problems as pr
    return {"lorenz": pr.lorenz_rhs, "robertson": pr.robertson_rhs,
            "brusselator1d": pr.brusselator_rhs, "brusselator1d_il": pr.brusselator_il_rhs}[name]


def python_rhs_lanes(name):
    """The lane-parallel form of a method-of-lines right-hand side (lsoda(..., rhs_lanes=)), or None."""
    import _problems as pr
    return {"brusselator1d": pr.brusselator_lanes, "brusselator1d_il": pr.brusselator_il_lanes}.get(name)


def initial_state(wl):
    import _problems as pr
    if wl["u0"] == "brusselator":
        return pr.brusselator_u0()
    if wl["u0"] == "brusselator_il":
        return pr.interleave(pr.brusselator_u0())
    return np.array(wl["u0"], dtype=np.float64)


# ---------------------------------------------------------------- algorithmic FLOPs
def dop853_flops(cnt, n, f_rhs, nt):
    """per attempted step 157n + 37 + 11 F; accepted + F; event step
    + 153n + 6 + 3 F; per output row 14n + 3; start-up 2 F + 12 n.  cnt = [B,8]."""
    nstep = cnt[:, 0].astype(np.float64)
    naccpt = cnt[:, 2].astype(np.float64)
    nevent = cnt[:, 4].astype(np.float64)
    rows = cnt[:, 6].astype(np.float64)
    fl = nstep * (157 * n + 37 + 11 * f_rhs) + naccpt * f_rhs + \
        nevent * (153 * n + 6 + 3 * f_rhs) + rows * (14 * n + 3) + (2 * f_rhs + 12 * n)
    return float(fl.sum())


def lsoda_flops(cnt, n, f_rhs, nt, band=None):
    """SURVEY.md 8(d), LSODA: counted from nst / nfe / nje with the mean order taken
    as nqu (upper estimate of the per-step bookkeeping is avoided: only the terms
    that scale with the counters are included)."""
    nst = cnt[:, 0].astype(np.float64)
    nfe = cnt[:, 1].astype(np.float64)
    nje = cnt[:, 2].astype(np.float64)
    nq = np.maximum(cnt[:, 4].astype(np.float64), 1.0)
    rows = cnt[:, 6].astype(np.float64)
    per_step = 3 * n + (n + 1) + n * nq * (nq + 1) / 2 + n + 2 * n * (nq + 1) + (n + 1) + 25 + \
        (3 * n + 30) / (nq + 1)
    per_fe = f_rhs + 2 * n * n + 8 * n + 10
    per_je = (2 * n * n + 4 * n) + (2 * n * n + n) + n + (2.0 / 3.0) * n ** 3 + 0.5 * n * n
    if band is not None:
        # the banded algorithm's own counts: dgbsl 2 n (2 ml + mu + 1), matrix fill and DBNORM over
        # the band, dgbfa 2 n ml (ml + mu) + n
        ml, mu = band
        mb = ml + mu + 1
        per_fe = f_rhs + 2 * n * (2 * ml + mu + 1) + 8 * n + 10
        per_je = (2 * n * mb + 4 * n) + (2 * n * mb + n) + n + 2 * n * ml * (ml + mu) + n
    fl = nst * per_step + nfe * per_fe + nje * per_je + rows * (n * (3 * nq + 1) + 2)
    return float(fl.sum())


def ivp_flops(res, n, f_rhs, f_ev):
    """solve_ivp = dp86co with iout = 2: every accepted step prepares dense output (the "event
    step" term of SURVEY.md 8(d)) and evaluates the events function; per output row 14n + 3.
    res = [B,12] (include/b200ode.h)."""
    nstep = res[:, 8].astype(np.float64)
    naccpt = res[:, 9].astype(np.float64)
    rows = res[:, 3].astype(np.float64)
    fl = nstep * (157 * n + 37 + 11 * f_rhs) + naccpt * (f_rhs + 153 * n + 6 + 3 * f_rhs + f_ev) + \
        rows * (14 * n + 3) + (2 * f_rhs + 12 * n)
    return float(fl.sum())


# dram__bytes_read.sum + dram__bytes_write.sum per trajectory of the solver kernel, from the
# `ncu --set full` captures summarised under profiles/ (bytes, capture, trajectories in it)
NCU_TRAFFIC = {
    "lorenz_dop853": (732.222976e6 + 6.947206e9, "profiles/r1_dop853_v3_fast_controller.txt", 262144),
    "robertson_lsoda": (31.317504e6 + 605.663744e6, "profiles/r1_lsoda_v5_throughput_mode.txt", 262144),
    "brusselator_lsoda": (4.124928e6 + 150.701824e6, "profiles/r1_lsodaw_v1_fast_solves.txt", 4000),
    "lorenz_solve_ivp": (342.943232e6 + 3.808480e9, "profiles/r1_ivp_v1_fast_controller.txt", 262144),
}
_traffic_file = os.path.join(ROOT, "profiles", "ncu_traffic.json")
if os.path.exists(_traffic_file):  # newer captures override the round-1 constants
    try:
        for _k, _v in json.load(open(_traffic_file)).items():
            NCU_TRAFFIC[_k] = (float(_v["dram_bytes"]), _v["capture"], int(_v["trajectories"]))
    except Exception:
        pass

F_RHS = {"lorenz": 8, "robertson": 13, "lotka_volterra": 6, "brusselator1d": 608, "brusselator1d_il": 608}


# ---------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 8 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 8:
                for k, nm in enumerate(names):
                    if r[4 + k].lower().startswith("active"):
                        reasons.add(nm)
        pw = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------- peaks
def measured_peaks():
    out = {"hbm_gbs": 6650.0, "hbm_source": "fallback (B200_PROFILING.md)"}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            out["hbm_gbs"] = float(json.load(open(p))["hbm_gbs"])
            out["hbm_source"] = "MEASURED_PEAKS.json (of measured)"
        except Exception:
            pass
    return out


NOMINAL_FP64_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 148 SMs x 64 DFMA lanes x 2 x 1.965 GHz = 37.2


def fp64_peak():
    """DFMA-chain microbenchmark (tools/fp64_peak.cu) run now on this GPU; falls back
    to the value recorded under profiles/."""
    exe = os.path.join(ROOT, "tools", "fp64_peak")
    try:
        r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        j = json.loads(r.stdout.strip().splitlines()[-1])
        return float(j["fp64_tflops"]), "tools/fp64_peak DFMA chains, measured in this run", j
    except Exception:
        p = os.path.join(ROOT, "profiles", "fp64_peak.json")
        if os.path.exists(p):
            j = json.load(open(p))
            return float(j["fp64_tflops"]), "profiles/fp64_peak.json (earlier run on this pool)", j
        return NOMINAL_FP64_TFLOPS, "nominal 148 SM x 64 lanes x 2 x 1.965 GHz (not measured)", None


# ---------------------------------------------------------------- CPU arm
def cpu_arm(wl, budget_s, seed=0):
    """The reference's CPU path on the host cores, bounded sample.  Returns
    (traj_per_s, cores, kind, sample description, seconds)."""
    import _oracle as orc
    cores = os.cpu_count() or 1
    te = np.linspace(*wl["t_eval"][:2], wl["t_eval"][2])
    u0 = initial_state(wl)
    f = orc.builtin_rhs(wl["rhs"])
    use_ref = wl["solver"] == "lsoda" and orc.ref_probe is not None

    def run(P):
        B = len(P)
        if use_ref:
            us = np.empty((B, len(te), wl["n"]))
            ok = np.zeros(B, dtype=np.int32)
            orc.ref_probe.ref_lsoda_batch.restype = ct.c_int
            t0 = time.perf_counter()
            with open(os.devnull, "w") as dn:
                fd = os.dup(2)
                os.dup2(dn.fileno(), 2)  # the reference prints warnings to stderr
                try:
                    orc.ref_probe.ref_lsoda_batch(
                        ct.c_void_p(f), ct.c_long(B), wl["n"], u0.ctypes.data_as(ct.c_void_p),
                        ct.c_long(0), P.ctypes.data_as(ct.c_void_p), ct.c_long(P.shape[1] * 8),
                        len(te), te.ctypes.data_as(ct.c_void_p), us.ctypes.data_as(ct.c_void_p),
                        ct.c_double(wl["rtol"]), ct.c_double(wl["atol"]), wl["mxstep"], 0,
                        ok.ctypes.data_as(ct.c_void_p), 0)
                finally:
                    os.dup2(fd, 2)
                    os.close(fd)
            return time.perf_counter() - t0
        t0 = time.perf_counter()
        if wl["solver"] == "solve_ivp":
            orc.solve_ivp_batch(f, [te[0], te[-1]], u0, te, 1, orc.builtin_rhs("events_" + wl["events"]),
                                P, rtol=wl["rtol"], atol=wl["atol"])
        else:
            orc.solve_batch(wl["solver"], f, u0, te, P, wl["rtol"], wl["atol"], wl["mxstep"])
        return time.perf_counter() - t0

    probe = sweep(wl["rhs"], 4 * cores, seed, wl.get("level"))
    t = run(probe)
    per = max(t / len(probe), 1e-7)
    B = int(min(wl["batch"], max(8 * cores, budget_s / per)))
    P = sweep(wl["rhs"], B, seed, wl.get("level"))
    t = run(P)
    kind = "reference" if use_ref else "port"
    what = ("oracle/_ref (unmodified reference LSODA.cpp + wrapper.cpp, g++ -O3)" if use_ref else
            "oracle port (C restatement; the reference's DOP853 / solve_ivp are Fortran, no compiler here)")
    sample = "%d trajectories of the same sweep (seed %d), %s, %d threads" % (B, seed, what, cores)
    return B / t, cores, kind, sample, t


def workload_config(workload, wl, total=None, world=1):
    cfg = {"workload": "%s: %s n=%d sweep, %s, t_eval=linspace(%g,%g,%d), rtol=atol=%g, "
                       "%d trajectories per GPU" % (workload, wl["rhs"], wl["n"], wl["solver"],
                                                    *wl["t_eval"], wl["rtol"], wl["batch"]),
           "reference_config": wl["config"], "trajectories_per_gpu": wl["batch"],
           "sharding": "by trajectory, no collective",
           "l2": "outputs (%.1f GB per step) exceed L2; inputs re-read per step are %d MB; "
                 "a 256 MB buffer is rewritten between timed steps to flush L2"
                 % (wl["batch"] * wl["t_eval"][2] * wl["n"] * 8 / 1e9,
                    wl["batch"] * wl["p"] * 8 >> 20)}
    return cfg


# ---------------------------------------------------------------- the GPU arm
class Gpu:
    """Per-rank state of the product arm: device, streams, reusable device buffers."""

    def __init__(self, a, rank, world, local, ranks):
        import torch
        import numbalsoda_b200 as nb
        from numbalsoda_b200 import _lib, rhs as _rhs
        self.torch, self.nb, self._lib, self._rhs = torch, nb, _lib, _rhs
        self.a, self.rank, self.world, self.local, self.ranks = a, rank, world, local, ranks
        self.dev = torch.device("cuda", local)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.stream = torch.cuda.current_stream()
        self.buf = {}
        self.handles = {}
        self.launch_total = 0

    def barrier(self):
        self.ranks.barrier()
        self.torch.cuda.synchronize()

    def tensor(self, key, shape, dtype):
        t = self.buf.get(key)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            self.buf.pop(key, None)
            t = self.torch.empty(shape, dtype=dtype, device=self.dev)
            self.buf[key] = t
        return t

    def handle(self, wl, rhs_kind, mode):
        """rhs_kind: "numba" (Python function -> numba.cuda LTO-IR -> nvJitLink) | "builtin" """
        key = (wl["rhs"], wl["solver"], rhs_kind, mode)
        if key not in self.handles:
            nb, _lib = self.nb, self._lib
            f = python_rhs(wl["rhs"]) if rhs_kind == "numba" else nb.builtin(wl["rhs"])
            if wl["solver"] == "solve_ivp":
                h = self._rhs.get_ivp_handle(f, wl["n"], 1, nb.builtin(wl["events"]), **MODES[mode])
            else:
                solver = _lib.DOP853 if wl["solver"] == "dop853" else _lib.LSODA
                lanes = python_rhs_lanes(wl["rhs"]) if rhs_kind == "numba" else None
                h = nb.get_handle(f, solver, wl["n"], rhs_lanes=lanes, **MODES[mode])
            self.handles[key] = h
        return self.handles[key]

    def device_leg(self, wl, rhs_kind, mode, B, steps, warmup, seed, clocks=False, launch_cap=None):
        """`steps` timed passes over B trajectories resident in HBM (in launches of at most
        launch_cap trajectories).  Returns timings (max over ranks), the per-trajectory counters
        of this rank's last pass and the launch count."""
        torch, _lib = self.torch, self._lib
        n, nt = wl["n"], wl["t_eval"][2]
        ivp = wl["solver"] == "solve_ivp"
        h = self.handle(wl, rhs_kind, mode)
        Ph = sweep(wl["rhs"], B, seed, wl.get("level"))
        teh = np.linspace(*wl["t_eval"][:2], nt)
        P = torch.tensor(Ph, device=self.dev)
        u0 = torch.tensor(initial_state(wl), dtype=torch.float64, device=self.dev)
        te = torch.tensor(teh, device=self.dev)
        Bl = int(min(B, launch_cap or B))  # trajectories per launch
        usol = self.tensor("usol", (Bl, nt, n), torch.float64)
        ok = self.tensor(("ok", B), (B,), torch.int32)
        ok.zero_()
        cnt = self.tensor(("cnt", B), (B, 8), torch.int32)
        stream = self.stream
        if ivp:
            span = torch.tensor([teh[0], teh[-1]], dtype=torch.float64, device=self.dev)
            res = torch.zeros((B, _lib.IVP_NRESULT), dtype=torch.int32, device=self.dev)
            t_event = torch.zeros(B, dtype=torch.float64, device=self.dev)
            y_event = torch.zeros((B, n), dtype=torch.float64, device=self.dev)
            pb = _lib.IvpProblem()
            pb.size = ct.sizeof(_lib.IvpProblem)
            pb.B, pb.neq, pb.nt, pb.np, pb.n_events = B, n, nt, wl["p"], 1
            pb.t_span, pb.t_span_stride, pb.y0, pb.y0_stride = span.data_ptr(), 0, u0.data_ptr(), 0
            pb.data, pb.data_stride, pb.teval = P.data_ptr(), wl["p"] * 8, te.data_ptr()
            pb.rtol, pb.atol, pb.y_out = wl["rtol"], wl["atol"], usol.data_ptr()
            pb.result, pb.t_event, pb.y_event = res.data_ptr(), t_event.data_ptr(), y_event.data_ptr()
        solver = _lib.DOP853 if wl["solver"] == "dop853" else _lib.LSODA

        band = wl.get("band")

        def step():
            if ivp:
                _lib.check(_lib.lib.b200ode_solve_ivp_device(h.ptr, ct.byref(pb), ct.c_void_p(stream.cuda_stream)))
                return
            if band:  # options beyond the positional entry points: the structure
                q = _lib.Problem()
                q.size = ct.sizeof(_lib.Problem)
                q.B, q.neq, q.nt, q.np = B, n, nt, wl["p"]
                q.u0, q.u0_stride, q.data, q.data_stride = u0.data_ptr(), 0, P.data_ptr(), wl["p"] * 8
                q.teval, q.usol, q.rtol, q.atol = te.data_ptr(), usol.data_ptr(), wl["rtol"], wl["atol"]
                q.mxstep, q.success, q.counters = wl["mxstep"], ok.data_ptr(), cnt.data_ptr()
                q.jt, q.ml, q.mu = 5, band[0], band[1]
                _lib.check(_lib.lib.b200ode_solve_device(h.ptr, ct.byref(q), ct.c_void_p(stream.cuda_stream)))
                return
            for b0 in range(0, B, Bl):
                nb_ = min(Bl, B - b0)
                args = [h.ptr, nb_, n, u0.data_ptr(), 0, P.data_ptr() + b0 * wl["p"] * 8, wl["p"] * 8,
                        wl["p"], nt, te.data_ptr(), usol.data_ptr(), wl["rtol"], wl["atol"], wl["mxstep"],
                        ok.data_ptr() + 4 * b0, cnt.data_ptr() + 32 * b0]
                if solver == _lib.DOP853:
                    rc = _lib.lib.b200ode_dop853_batched_device(*args, ct.c_void_p(stream.cuda_stream))
                else:
                    rc = _lib.lib.b200ode_lsoda_batched_device(*args, 0, ct.c_void_p(stream.cuda_stream))
                _lib.check(rc)

        for _ in range(max(warmup, 0)):
            step()
            self.flush.fill_(1)
        launches0 = h.kernel_info()["launches"]
        sampler = ClockSampler(self.local) if clocks else None
        self.barrier()
        if sampler:
            sampler.start()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(steps)]
        t_wall0 = time.perf_counter()
        for k in range(steps):
            ev[k][0].record(stream)
            step()
            ev[k][1].record(stream)
            self.flush.fill_(k & 0xFF)
        self.barrier()
        t_wall = time.perf_counter() - t_wall0
        clk = sampler.stop() if sampler else None
        launches = h.kernel_info()["launches"] - launches0
        self.launch_total += launches
        ms = [e0.elapsed_time(e1) for e0, e1 in ev]
        dev_s = sum(ms) / 1e3
        dev_s_max = self.ranks.max(dev_s)  # every multi-GPU time is the slowest rank's
        n_ok = int(self.ranks.sum(int(((res[:, 0] if ivp else ok) == 1).sum().item())))
        out = {"value": self.world * B * steps / dev_s_max, "ms_per_step": 1e3 * dev_s_max / steps,
               "per_launch_s_rank": dev_s / steps, "dev_s_max": dev_s_max, "wall_s": t_wall, "clocks": clk,
               "launches": launches, "success_fraction": n_ok / (B * self.world),
               "kernel_info": h.kernel_info(), "B": B}
        if self.rank == 0:
            if ivp:
                res_h = res.cpu().numpy()
                out["res"] = res_h
                out["cnt"] = res_h[:, 8:9]
            else:
                out["cnt"] = cnt.cpu().numpy()
        return out

    # ---- end to end: the call a user makes -- numbalsoda_b200.lsoda / dop853 with host
    #      (numpy) inputs, returning usol / success on the host; H2D of the inputs, the
    #      kernels and D2H of every row are inside the timed region.  The result buffers are
    #      page-locked and reused (out=): allocating and first-touching a fresh 2.5-25 GB
    #      array per call costs more than the integration itself.  The whole ensemble goes
    #      through, in slices that rotate over two buffers (a slice's rows are the host's to
    #      read until the slice after next is requested).
    def e2e_leg(self, wl, rhs_kind, mode, B, seed, steps):
        nb, _lib = self.nb, self._lib
        n, nt = wl["n"], wl["t_eval"][2]
        ivp = wl["solver"] == "solve_ivp"
        u0h = initial_state(wl)
        teh = np.linspace(*wl["t_eval"][:2], nt)
        Ph = sweep(wl["rhs"], B, seed, wl.get("level"))
        row_bytes = nt * n * 8
        # two page-locked buffers per rank, each at most 6.4 GB and a fifth of this rank's share
        # of the host's available memory (8 ranks pin theirs on one host)
        per_buf = int(os.environ.get("B200ODE_BENCH_PINNED_MB", "6400")) << 20
        try:
            avail = [int(ln.split()[1]) << 10 for ln in open("/proc/meminfo") if ln.startswith("MemAvailable")][0]
            per_buf = int(min(per_buf, 0.2 * avail / self.world))
        except Exception:
            pass
        Bs = int(max(1, min(B, per_buf // row_bytes)))
        nslices = (B + Bs - 1) // Bs
        Bs = (B + nslices - 1) // nslices
        bufs = [nb.pinned_empty((Bs, nt, n), np.float64) for _ in range(min(2, nslices))]
        rhs = python_rhs(wl["rhs"]) if rhs_kind == "numba" else nb.builtin(wl["rhs"])
        api = nb.dop853 if wl["solver"] == "dop853" else nb.lsoda
        kw = dict(MODES[mode])
        if wl.get("band"):
            kw.update(ml=wl["band"][0], mu=wl["band"][1])
        if rhs_kind == "numba" and python_rhs_lanes(wl["rhs"]) is not None:
            kw.update(rhs_lanes=python_rhs_lanes(wl["rhs"]))

        def one_pass():
            good = 0
            for s in range(nslices):
                b0, b1 = s * Bs, min(B, (s + 1) * Bs)
                out = bufs[s % len(bufs)][:b1 - b0]
                if ivp:
                    sol = nb.solve_ivp(rhs, np.array([teh[0], teh[-1]]), u0h, teh, n_events=1,
                                       event_fcn=nb.builtin(wl["events"]), data=Ph[b0:b1], rtol=wl["rtol"],
                                       atol=wl["atol"], device=self.local, out=out, **kw)
                    good += int(sol.success.sum())
                else:
                    us, okh = api(rhs, u0h, teh, Ph[b0:b1], wl["rtol"], wl["atol"], wl["mxstep"],
                                  device=self.local, out=out, **kw)
                    good += int(okh.sum())  # the host reads the step's result
            return good

        one_pass()  # (two untimed passes: on a fresh box the second one is still ~10 % slow)
        one_pass()
        self.barrier()
        t0 = time.perf_counter()
        ke = max(1, min(steps, 3))
        good = 0
        for _ in range(ke):
            good = one_pass()
        self.barrier()
        te2e = self.ranks.max(time.perf_counter() - t0)
        d2h = int(B * row_bytes + B * (4 if not ivp else 48 + 8 + 8 * n))
        e2e = {"value": self.world * B * ke / te2e, "unit": "traj/s",
               "h2d_bytes_per_step": int(Ph.nbytes + nslices * (teh.nbytes + u0h.nbytes)),
               "d2h_bytes_per_step": d2h, "trajectories_per_gpu": B, "slices": nslices,
               "pinned_bytes_per_rank": int(len(bufs) * Bs * row_bytes),
               "steps": ke, "ms_per_step": 1e3 * te2e / ke, "success_fraction": good / B,
               "d2h_gbs_all_ranks": self.world * d2h * ke / te2e / 1e9,
               "api": ("numbalsoda_b200.solve_ivp(rhs, t_span, y0, t_eval, n_events=1, event_fcn, "
                       "data[B,p], rtol, atol, out=pinned) -> ODEResultBatch on the host") if ivp else
                      "numbalsoda_b200.%s(rhs, u0, t_eval, data[B,p], rtol, atol, out=pinned) "
                      "-> (usol[B,nt,n], success[B]) on the host" % wl["solver"]}
        # ---- the ceiling: the same bytes as plain device-to-host copies (one cuMemcpyAsync
        #      per slice into the same page-locked buffers, all ranks at once, no kernels):
        #      what this host can take from N GPUs.  e2e / ceiling says how much of the
        #      end-to-end time is the library's to improve.
        torch = self.torch
        src = self.tensor("usol", (Bs, nt, n), torch.float64)
        dst = [torch.from_numpy(b) for b in bufs]
        for d in dst:
            d.copy_(src, non_blocking=True)
        self.barrier()
        t0 = time.perf_counter()
        for s in range(nslices):
            b0, b1 = s * Bs, min(B, (s + 1) * Bs)
            dst[s % len(dst)][:b1 - b0].copy_(src[:b1 - b0], non_blocking=True)
        self.barrier()
        tc = self.ranks.max(time.perf_counter() - t0)
        ceil_gbs = self.world * B * row_bytes / tc / 1e9
        e2e["ceiling"] = {"d2h_gbs_all_ranks": ceil_gbs, "traj_per_s": self.world * B / tc,
                          "how": "torch copy_ (cudaMemcpyAsync) of the same rows, device -> the same "
                                 "page-locked buffers, all ranks concurrently, wall clock max over ranks"}
        e2e["frac_of_ceiling"] = e2e["value"] / e2e["ceiling"]["traj_per_s"]
        del bufs, dst
        return e2e


def roofline_block(workload, wl, leg, peak_tf, peak_src):
    """FP64 roofline of one device leg from the per-trajectory counters (rank 0's shard)."""
    n, nt, B = wl["n"], wl["t_eval"][2], leg["B"]
    ivp = wl["solver"] == "solve_ivp"
    if ivp:
        flops = ivp_flops(leg["res"], n, F_RHS[wl["rhs"]], 1)
        rows_h = float(leg["res"][:, 3].mean())
    else:
        if wl["solver"] == "dop853":
            flops = dop853_flops(leg["cnt"], n, F_RHS[wl["rhs"]], nt)
        else:
            flops = lsoda_flops(leg["cnt"], n, F_RHS[wl["rhs"]], nt, wl.get("band"))
    per_launch_s = leg["per_launch_s_rank"]
    achieved_tf = flops / per_launch_s / 1e12
    peaks = measured_peaks()
    out_bytes = B * (nt * n * 8 + (n + wl["p"]) * 8 + 4)
    if ivp:  # rows actually written (an event trims the rest), the result record, the event
        out_bytes = int(B * (rows_h * n * 8 + wl["p"] * 8 + 48 + 8 + 8 * n))
    tr = NCU_TRAFFIC.get(workload)
    return {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": achieved_tf / peak_tf if peak_tf else None,
            "frac_of_nominal": achieved_tf / NOMINAL_FP64_TFLOPS,
            "nominal_peak": NOMINAL_FP64_TFLOPS,
            "traffic": (tr[0] / tr[2] * B) if tr else None,
            "traffic_source": ("ncu dram bytes per trajectory x B, %s" % tr[1]) if tr else None,
            "kernel": "b200ode_ivp_kernel" if ivp else
                      "b200ode_%s_kernel" % (wl["solver"] + ("w" if wl["n"] > 8 else "")),
            "flops_per_launch": flops, "flops_per_trajectory": flops / B,
            "flop_count": "algorithmic, SURVEY.md 8(d) formulas x per-trajectory counters",
            "peak_source": peak_src,
            "hbm": {"achieved_gbs": out_bytes / per_launch_s / 1e9, "peak_gbs": peaks["hbm_gbs"],
                    "frac": out_bytes / per_launch_s / 1e9 / peaks["hbm_gbs"],
                    "bytes_per_launch": out_bytes, "peak_source": peaks["hbm_source"]},
            "roofline_traj_per_s_per_gpu": min(peak_tf * 1e12 / (flops / B),
                                               peaks["hbm_gbs"] * 1e9 / (out_bytes / B))
            if peak_tf else None}


def parity_block(gpu, wl, rhs_kind):
    """The measured parity statistic of every link mode, on rank 0 (the CPU oracle is the
    checker here, never the thing measured): 64 trajectories of the benchmark's own sweep
    through the public API, compared with the oracle run of the same inputs and with a
    tight-tolerance truth (oracle at rtol = atol = 1e-13).  Chaotic workloads are compared over
    the horizon on which a truth exists (`parity_rows` leading rows of t_eval)."""
    import _oracle as orc
    nb = gpu.nb
    Bp = 64
    P = sweep(wl["rhs"], Bp, 0)
    nt = wl.get("parity_rows", wl["t_eval"][2])
    te = np.linspace(*wl["t_eval"][:2], wl["t_eval"][2])[:nt]
    u0 = initial_state(wl)
    f = orc.builtin_rhs(wl["rhs"])
    rtol, atol = wl["rtol"], wl["atol"]
    ref, okr, cr = orc.solve_batch(wl["solver"], f, u0, te, P, rtol, atol, wl["mxstep"], band=wl.get("band"))
    truth, okt, _ = orc.solve_batch(wl["solver"], f, u0, te, P, 1e-13, 1e-13, 1000000)
    w = rtol * np.abs(truth) + atol
    e_ref = float((np.abs(ref - truth) / w).max())
    api = nb.dop853 if wl["solver"] == "dop853" else nb.lsoda
    rhs = python_rhs(wl["rhs"]) if rhs_kind == "numba" else nb.builtin(wl["rhs"])
    out = {}
    bkw = dict(ml=wl["band"][0], mu=wl["band"][1]) if wl.get("band") else {}
    if rhs_kind == "numba" and python_rhs_lanes(wl["rhs"]) is not None:
        bkw["rhs_lanes"] = python_rhs_lanes(wl["rhs"])
    for mode in MODES:
        us, ok, cnt = api(rhs, u0, te, P, rtol, atol, wl["mxstep"], device=gpu.local, counters=True,
                          **MODES[mode], **bkw)
        e = float((np.abs(us - truth) / w).max())
        out[mode] = {
            "bit_identical_to_oracle": bool(np.array_equal(us, ref)),
            "counters_identical_to_oracle": bool(np.array_equal(cnt[:, :4], cr[:, :4])),
            "max_err_vs_oracle": float((np.abs(us - ref) / (rtol * np.abs(ref) + atol)).max()),
            "max_err_vs_truth": e, "reference_max_err_vs_truth": e_ref,
            "err_ratio_to_reference": e / e_ref if e_ref > 0 else None,
            "mean_steps": float(cnt[:, 0].mean()), "reference_mean_steps": float(cr[:, 0].mean()),
            "all_succeeded": bool(ok.all())}
    note = ("%d trajectories of the sweep (seed 0), rows t_eval[:%d]; errors in units of rtol*|y|+atol; "
            "truth = CPU oracle at rtol = atol = 1e-13; oracle = the reference arithmetic (LSODA: "
            "bit-identical to the compiled reference)" % (Bp, nt))
    return out, note


def full_block(gpu, a, workload, peak, primary):
    """value / roofline / e2e / modes / rhs_paths (/ cpu_baseline) of one workload."""
    wl = dict(WORKLOADS[workload])
    if a.batch:
        wl["batch"] = a.batch
    B = wl["batch"]
    rank = gpu.rank
    mode = "strict" if a.strict else ("exact_ctrl" if a.exact_ctrl else "fast")
    has_py = wl["solver"] != "solve_ivp" and not a.builtin_rhs
    rhs_kind = "numba" if has_py else "builtin"
    steps, warmup = (a.steps, a.warmup) if primary else (max(1, min(a.steps, 3)), max(1, min(a.warmup, 3)))
    main = gpu.device_leg(wl, rhs_kind, mode, B, steps, warmup, seed=rank, clocks=primary)
    side = {}
    if not a.no_modes and wl["solver"] != "solve_ivp":
        # the other link modes and the other right-hand-side path, 3 timed steps each
        for m in MODES:
            if m != mode:
                side[m] = gpu.device_leg(wl, rhs_kind, m, B, 3, 1, seed=rank)
        if has_py:
            side["builtin"] = gpu.device_leg(wl, "builtin", mode, B, 3, 1, seed=rank)
    e2e = None
    if not a.no_e2e:
        e2e = gpu.e2e_leg(wl, rhs_kind, mode, B, rank, a.steps)
    if rank != 0:
        return None
    peak_tf, peak_src, _ = peak
    blk = {"value": main["value"], "unit": "traj/s", "ms_per_step": main["ms_per_step"],
           "steps": steps, "warmup": warmup, "mode": MODE_TEXT[mode],
           "rhs": ("Python function compiled by numba.cuda to LTO-IR, linked by nvJitLink" if rhs_kind == "numba"
                   else "nvcc-built right-hand side shipped in libb200ode.so"),
           "config": workload_config(workload, wl),
           "success_fraction": main["success_fraction"], "wall_s_timed_region": main["wall_s"],
           "mean_steps_per_trajectory": float(main["cnt"][:, 0].mean()),
           "roofline": roofline_block(workload, wl, main, peak_tf, peak_src),
           "kernel_info": main["kernel_info"], "gpu_launches": main["launches"]}
    if main["clocks"]:
        blk["clocks"] = main["clocks"]
    if e2e:
        blk["e2e"] = e2e
    if side:
        modes = {mode: {"value": main["value"], "roofline_frac": blk["roofline"]["frac"],
                        "ms_per_step": main["ms_per_step"]}}
        for m in MODES:
            if m in side:
                r = roofline_block(workload, wl, side[m], peak_tf, peak_src)
                modes[m] = {"value": side[m]["value"], "roofline_frac": r["frac"],
                            "ms_per_step": side[m]["ms_per_step"]}
        try:
            par, note = parity_block(gpu, wl, rhs_kind)
            for m in modes:
                modes[m]["parity"] = par[m]
            modes["parity_sample"] = note
        except Exception as e:  # noqa: BLE001  (the checker must not take the bench line down)
            modes["parity_sample"] = "parity leg failed: %r" % (e,)
        blk["modes"] = modes
        if "builtin" in side:
            blk["rhs_paths"] = {"numba": main["value"], "builtin": side["builtin"]["value"],
                                "numba_over_builtin": main["value"] / side["builtin"]["value"],
                                "note": "same kernel, same mode; numba = north_star's path (Python "
                                        "function -> numba.cuda LTO-IR -> nvJitLink), builtin = nvcc-built"}
    if wl["solver"] == "solve_ivp":
        blk["event_fraction"] = float(main["res"][:, 5].mean())
        blk["mean_rows_per_trajectory"] = float(main["res"][:, 3].mean())
    if not a.no_cpu and gpu.world == 1:
        v, cores, kind, sample, secs = cpu_arm(wl, 12.0)
        blk["cpu_baseline"] = {"value": v, "unit": "traj/s", "cores": cores, "kind": kind,
                               "sample": sample, "seconds": secs}
    return blk


def mixed_total(gpu, a, peak):
    """BASELINE.json configs[4] as stated: a FIXED total of mixed trajectories (half Lorenz
    DOP853, half Robertson LSODA), sharded over the ranks -- strong scaling -- two homogeneous
    kinds per GPU (one linked kernel each), each kind in launches of at most 2^20 trajectories
    into one reused row buffer (5 M Lorenz trajectories are 120 GB of rows)."""
    total = a.total
    world, rank = gpu.world, gpu.rank
    per_kind = total // 2
    lo, hi = (per_kind * rank) // world, (per_kind * (rank + 1)) // world
    Bk = hi - lo
    legs, e2es = {}, {}
    steps, warmup = max(1, min(a.steps, 3)), max(1, min(a.warmup, 1))
    for wname in MIXED:
        wl = dict(WORKLOADS[wname])
        wl["batch"] = Bk
        legs[wname] = (wl, gpu.device_leg(wl, "builtin" if a.builtin_rhs else "numba",
                                          "fast", Bk, steps, warmup, seed=rank, launch_cap=1 << 20,
                                          clocks=(wname == MIXED[0])))
        if not a.no_e2e:
            e2es[wname] = gpu.e2e_leg(wl, "builtin" if a.builtin_rhs else "numba", "fast", Bk, rank, 1)
    if rank != 0:
        return None
    peak_tf, peak_src, _ = peak
    t_dev = sum(l["dev_s_max"] / steps for _, l in legs.values())
    n_all = 2 * per_kind
    line = {"metric": "trajectories/s", "value": n_all / t_dev, "unit": "traj/s", "n_gpus": world,
            "steps": steps, "warmup": warmup, "ms_per_step": 1e3 * t_dev, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "mixed: %d Lorenz DOP853 + %d Robertson LSODA trajectories in total, "
                                   "sharded over %d GPU(s)" % (per_kind, per_kind, world),
                       "reference_config": "BASELINE.json configs[4] (SURVEY 8d C5): fixed total, strong scaling",
                       "total_trajectories": n_all, "trajectories_per_gpu": 2 * Bk,
                       "launch_cap": 1 << 20, "sharding": "by trajectory, no collective",
                       "l2": "outputs exceed L2; a 256 MB buffer is rewritten between timed steps"},
            "mode": MODE_TEXT["fast"],
            "parts": {w: {"value": l["value"], "ms_per_step": l["ms_per_step"],
                          "roofline": roofline_block(w, wl, l, peak_tf, peak_src),
                          "success_fraction": l["success_fraction"]} for w, (wl, l) in legs.items()},
            "roofline": roofline_block(MIXED[0], legs[MIXED[0]][0], legs[MIXED[0]][1], peak_tf, peak_src),
            "clocks": legs[MIXED[0]][1]["clocks"], "gpu_launches": gpu.launch_total}
    if e2es:
        t_e = sum(e["ms_per_step"] / 1e3 for e in e2es.values())
        line["e2e"] = {"value": n_all / t_e, "unit": "traj/s",
                       "h2d_bytes_per_step": sum(e["h2d_bytes_per_step"] for e in e2es.values()),
                       "d2h_bytes_per_step": sum(e["d2h_bytes_per_step"] for e in e2es.values()),
                       "parts": e2es}
    if not a.no_cpu and world == 1:
        # numba-prange-style CPU arm: both kinds on all host cores, bounded samples
        cb = {w: cpu_arm(dict(WORKLOADS[w]), 8.0) for w in MIXED}
        t_cpu = sum(per_kind / cb[w][0] for w in MIXED)
        line["cpu_baseline"] = {"value": n_all / t_cpu, "unit": "traj/s", "cores": cb[MIXED[0]][1],
                                "kind": "reference (LSODA) + port (DOP853)",
                                "sample": "; ".join(cb[w][3] for w in MIXED),
                                "parts": {w: cb[w][0] for w in MIXED}}
    return line


# ---------------------------------------------------------------- main
def reference_line(a, workload):
    wl = dict(WORKLOADS[workload])
    if a.batch:
        wl["batch"] = a.batch
    vals = []
    for _ in range(a.warmup):
        cpu_arm(wl, 1.0)
    info = None
    for _ in range(max(a.steps, 1)):
        info = cpu_arm(wl, 8.0)
        vals.append(info)
    B_s = sum(float(v[0]) * v[4] for v in vals)
    T_s = sum(v[4] for v in vals)
    v = B_s / T_s
    return {"impl": "reference", "metric": "trajectories/s", "value": v, "unit": "traj/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1e3 * T_s / max(a.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(workload, wl),
            "cpu_baseline": {"value": v, "unit": "traj/s", "cores": info[1], "kind": info[2],
                             "sample": info[3]},
            "e2e": {"value": v, "unit": "traj/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lorenz_dop853", choices=sorted(WORKLOADS) + ["mixed"])
    ap.add_argument("--batch", type=int, default=0, help="trajectories per GPU (default 2^20)")
    ap.add_argument("--total", type=int, default=0,
                    help="mixed: fixed total of trajectories over all GPUs (BASELINE: 10000000)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-modes", action="store_true", help="skip the other link modes / RHS path")
    ap.add_argument("--no-secondary", action="store_true",
                    help="default workload only: skip the Robertson LSODA block")
    ap.add_argument("--builtin-rhs", action="store_true",
                    help="headline on the nvcc-built right-hand side instead of the numba-compiled one")
    ap.add_argument("--strict", action="store_true", help="FMA contraction off (parity mode)")
    ap.add_argument("--exact-ctrl", action="store_true",
                    help="FMA mode with the reference's controller arithmetic (B200ODE_EXACT_CTRL)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        # the reference's CPU path: rank 0 alone runs and prints it
        if rank == 0:
            print(json.dumps(reference_line(a, MIXED[0] if a.workload == "mixed" else a.workload)))
        return
    if a.gpus != world and world == 1 and a.gpus > 1:
        # python bench.py --gpus N without torchrun: re-launch one rank per GPU
        import socket
        sk = socket.socket()
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
        sk.close()
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   "--nproc-per-node", str(a.gpus), "--master-addr", "127.0.0.1",
                                   "--master-port", str(port)] + sys.argv)
    import torch
    from numbalsoda_b200._dist import Ranks
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    ranks = Ranks(backend="nccl" if world > 1 else None)  # plumbing only: no data-path collective
    if world > 1 and os.environ.get("B200ODE_NUMA_BIND", "1") == "1":
        ranks.bind_to_gpu_numa_node()  # host buffers next to each rank's GPU
    gpu = Gpu(a, rank, world, local, ranks)
    peak = fp64_peak() if rank == 0 else (None, None, None)

    if a.workload == "mixed" and a.total:
        line = mixed_total(gpu, a, peak)
        ranks.close()
        if rank == 0:
            print(json.dumps(line))
        return

    names = MIXED if a.workload == "mixed" else [a.workload]
    blocks = [full_block(gpu, a, nm, peak, primary=(i == 0)) for i, nm in enumerate(names)]
    secondary = None
    if a.workload == "lorenz_dop853" and not a.no_secondary:
        secondary = full_block(gpu, a, "robertson_lsoda", peak, primary=False)
    ranks.close()
    if rank != 0:
        return

    def as_line(b):
        line = {"metric": "trajectories/s", "value": b["value"], "unit": "traj/s", "n_gpus": world,
                "steps": b["steps"], "warmup": b["warmup"], "ms_per_step": b["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic"}
        line.update({k: v for k, v in b.items() if k not in line})
        return line

    if len(blocks) == 1:
        line = as_line(blocks[0])
        if secondary:
            line["secondary"] = {"robertson_lsoda": secondary}
        line["gpu_launches"] = gpu.launch_total
        print(json.dumps(line))
        return
    # BASELINE.json configs[4], weak-scaling form: the mixed sweep is two homogeneous launches
    # per GPU (one linked kernel per solver / right-hand side); aggregate = all trajectories /
    # all time (per step: every kind's ensemble once; time of a kind = its trajectories / its rate)
    lines = [as_line(b) for b in blocks]
    tot_traj = sum(l["config"]["trajectories_per_gpu"] * l["n_gpus"] for l in lines)
    tot_s = sum(l["config"]["trajectories_per_gpu"] * l["n_gpus"] / l["value"] for l in lines)
    slow = min(lines, key=lambda l: l["value"])
    line = dict(slow)
    line.update({"value": tot_traj / tot_s, "ms_per_step": 1e3 * tot_s,
                 "config": {"workload": "mixed: " + " + ".join(l["config"]["workload"] for l in lines),
                            "reference_config": "BASELINE.json configs[4] (SURVEY 8d C5), weak scaling "
                                                "(--total N gives the fixed-total, strong-scaling form)",
                            "trajectories_per_gpu": sum(l["config"]["trajectories_per_gpu"] for l in lines),
                            "sharding": "by trajectory, no collective", "l2": lines[0]["config"]["l2"]},
                 "gpu_launches": gpu.launch_total,
                 "parts": [{k: l.get(k) for k in ("value", "ms_per_step", "roofline", "e2e", "cpu_baseline")}
                           for l in lines]})
    if all("e2e" in l for l in lines):
        e_s = sum(l["config"]["trajectories_per_gpu"] * l["n_gpus"] / l["e2e"]["value"] for l in lines)
        line["e2e"] = {"value": sum(l["config"]["trajectories_per_gpu"] * l["n_gpus"] for l in lines) / e_s,
                       "unit": "traj/s",
                       "h2d_bytes_per_step": sum(l["e2e"]["h2d_bytes_per_step"] for l in lines),
                       "d2h_bytes_per_step": sum(l["e2e"]["d2h_bytes_per_step"] for l in lines)}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
