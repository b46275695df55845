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
  mixed           (configs[4], "C5") the Lorenz and the Robertson ensembles as two
                  homogeneous launches per GPU; value = all trajectories / all time
Multi-GPU: one process per GPU (torchrun), the ensemble is sharded by trajectory, no
data-path collective; `value` = trajectories of all ranks / max-over-ranks time.

Prints ONE JSON line (rank 0).  `value` is measured with inputs and outputs resident
in HBM (CUDA events on the launch stream); `e2e` is the same workload through the call a
user makes -- numbalsoda_b200.lsoda / dop853 with numpy inputs and a page-locked `out=`
result buffer: H2D, kernels and D2H of every row inside the timed region.  `--impl reference` times the reference's CPU implementation of the path on
the host cores on a bounded sample (oracle/_ref when it compiled -- LSODA --, else the
oracle port -- DOP853, whose reference is Fortran).
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
                          batch=1 << 20, config="BASELINE.json configs[1] (SURVEY 8d C2)"),
    "robertson_lsoda": dict(solver="lsoda", rhs="robertson", n=3, p=3, u0=[1.0, 0.0, 0.0],
                            t_eval=(0.0, 1.0e5, 100), rtol=1e-8, atol=1e-8, mxstep=10000,
                            batch=1 << 20, config="BASELINE.json configs[2] (SURVEY 8d C3)"),
    "brusselator_lsoda": dict(solver="lsoda", rhs="brusselator1d", n=64, p=3, u0="brusselator",
                              t_eval=(0.0, 10.0, 100), rtol=1e-6, atol=1e-8, mxstep=10000,
                              batch=100000, config="BASELINE.json configs[3] (SURVEY 8d C4): "
                              "1-D Brusselator on 32 cells, warp per trajectory"),
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


def sweep(name, B, seed, level=None):
    import _problems as pr
    if level is not None:  # solve_ivp: the event level as a 4th parameter
        return np.ascontiguousarray(np.column_stack([pr.lorenz_sweep(B, seed), np.full(B, level)]))
    if name == "lorenz":
        return pr.lorenz_sweep(B, seed)
    if name == "brusselator1d":
        return pr.brusselator_sweep(B, seed)
    return pr.robertson_sweep(B, seed)


def initial_state(wl):
    import _problems as pr
    return pr.brusselator_u0() if wl["u0"] == "brusselator" else np.array(wl["u0"], dtype=np.float64)


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


def lsoda_flops(cnt, n, f_rhs, nt):
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

F_RHS = {"lorenz": 8, "robertson": 13, "lotka_volterra": 6, "brusselator1d": 608}


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
        return 37.2, "nominal 148 SM x 64 lanes x 2 x 1.965 GHz (not measured)", None


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


# ---------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lorenz_dop853", choices=sorted(WORKLOADS) + ["mixed"])
    ap.add_argument("--batch", type=int, default=0, help="trajectories per GPU (default 2^20)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--strict", action="store_true", help="FMA contraction off (parity mode)")
    ap.add_argument("--exact-ctrl", action="store_true",
                    help="FMA mode with the reference's controller arithmetic (B200ODE_EXACT_CTRL)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "ours" and a.gpus != world and world == 1 and a.gpus > 1:
        # python bench.py --gpus N without torchrun: re-launch one rank per GPU
        import socket
        sk = socket.socket()
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
        sk.close()
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                                   "--nproc-per-node", str(a.gpus), "--master-addr", "127.0.0.1",
                                   "--master-port", str(port)] + sys.argv)
    names = MIXED if a.workload == "mixed" else [a.workload]
    ranks = None
    if a.impl == "ours":
        import torch
        from numbalsoda_b200._dist import Ranks
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
        torch.cuda.set_device(local)
        ranks = Ranks(backend="nccl" if world > 1 else None)  # plumbing only: no data-path collective
        if world > 1 and os.environ.get("B200ODE_NUMA_BIND", "1") == "1":
            ranks.bind_to_gpu_numa_node()  # host buffers next to each rank's GPU
    lines = [bench_one(a, nm, ranks, rank, world, local) for nm in names]
    if ranks is not None:
        ranks.close()
    if rank != 0:
        return
    if len(lines) == 1:
        print(json.dumps(lines[0]))
        return
    # BASELINE.json configs[4]: the mixed sweep is two homogeneous launches per GPU (one
    # linked kernel per solver / right-hand side); aggregate = all trajectories / all time
    # (per step: every kind's ensemble once; time of a kind = its trajectories / its rate)
    tot_traj = sum(l["config"]["trajectories_per_gpu"] * l["n_gpus"] for l in lines)
    tot_s = sum(l["config"]["trajectories_per_gpu"] * l["n_gpus"] / l["value"] for l in lines)
    slow = min(lines, key=lambda l: l["value"])
    line = dict(slow)
    line.update({"value": tot_traj / tot_s, "ms_per_step": 1e3 * tot_s,
                 "config": {"workload": "mixed: " + " + ".join(l["config"]["workload"] for l in lines),
                            "reference_config": "BASELINE.json configs[4] (SURVEY 8d C5), weak scaling",
                            "trajectories_per_gpu": sum(l["config"]["trajectories_per_gpu"] for l in lines),
                            "sharding": "by trajectory, no collective", "l2": lines[0]["config"]["l2"]},
                 "gpu_launches": sum(l.get("gpu_launches", 0) for l in lines),
                 "parts": [{k: l.get(k) for k in ("value", "ms_per_step", "roofline", "e2e", "cpu_baseline")}
                           for l in lines]})
    if all("e2e" in l for l in lines):
        e_s = sum(l["config"]["trajectories_per_gpu"] * l["n_gpus"] / l["e2e"]["value"] for l in lines)
        line["e2e"] = {"value": sum(l["config"]["trajectories_per_gpu"] * l["n_gpus"] for l in lines) / e_s,
                       "unit": "traj/s",
                       "h2d_bytes_per_step": sum(l["e2e"]["h2d_bytes_per_step"] for l in lines),
                       "d2h_bytes_per_step": sum(l["e2e"]["d2h_bytes_per_step"] for l in lines)}
    print(json.dumps(line))


def bench_one(a, workload, ranks, rank, world, local):
    """One workload on this rank; returns the JSON line (rank 0) or None."""
    wl = dict(WORKLOADS[workload])
    if a.batch:
        wl["batch"] = a.batch
    cfg = {"workload": "%s: %s n=%d sweep, %s, t_eval=linspace(%g,%g,%d), rtol=atol=%g, "
                       "%d trajectories per GPU" % (workload, wl["rhs"], wl["n"], wl["solver"],
                                                    *wl["t_eval"], wl["rtol"], wl["batch"]),
           "reference_config": wl["config"], "trajectories_per_gpu": wl["batch"],
           "sharding": "by trajectory, no collective",
           "l2": "outputs (%.1f GB per step) exceed L2; inputs re-read per step are %d MB; "
                 "a 256 MB buffer is rewritten between timed steps to flush L2"
                 % (wl["batch"] * wl["t_eval"][2] * wl["n"] * 8 / 1e9,
                    wl["batch"] * wl["p"] * 8 >> 20)}

    if a.impl == "reference":
        if rank != 0:
            return None
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
        line = {"impl": "reference", "metric": "trajectories/s", "value": v, "unit": "traj/s",
                "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": 1e3 * T_s / max(a.steps, 1), "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg,
                "cpu_baseline": {"value": v, "unit": "traj/s", "cores": info[1], "kind": info[2],
                                 "sample": info[3]},
                "e2e": {"value": v, "unit": "traj/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}}
        return line

    import torch
    import numbalsoda_b200 as nb
    from numbalsoda_b200 import _lib

    dev = torch.device("cuda", local)

    B, n, nt = wl["batch"], wl["n"], wl["t_eval"][2]
    ivp = wl["solver"] == "solve_ivp"
    solver = _lib.SOLVE_IVP if ivp else _lib.DOP853 if wl["solver"] == "dop853" else _lib.LSODA
    if ivp:
        from numbalsoda_b200 import rhs as _rhs
        h = _rhs.get_ivp_handle(nb.builtin(wl["rhs"]), n, 1, nb.builtin(wl["events"]), strict=a.strict,
                                exact_ctrl=a.exact_ctrl)
    else:
        h = nb.get_handle(nb.builtin(wl["rhs"]), solver, n, strict=a.strict, exact_ctrl=a.exact_ctrl)
    Ph = sweep(wl["rhs"], B, rank, wl.get("level"))  # every rank integrates its own shard of the sweep
    teh = np.linspace(*wl["t_eval"][:2], nt)
    P = torch.tensor(Ph, device=dev)
    u0 = torch.tensor(initial_state(wl), dtype=torch.float64, device=dev)
    te = torch.tensor(teh, device=dev)
    usol = torch.empty((B, nt, n), dtype=torch.float64, device=dev)
    ok = torch.zeros(B, dtype=torch.int32, device=dev)
    cnt = torch.zeros((B, 8), dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    if ivp:
        span = torch.tensor([teh[0], teh[-1]], dtype=torch.float64, device=dev)
        res = torch.zeros((B, _lib.IVP_NRESULT), dtype=torch.int32, device=dev)
        t_event = torch.zeros(B, dtype=torch.float64, device=dev)
        y_event = torch.zeros((B, n), dtype=torch.float64, device=dev)
        pb = _lib.IvpProblem()
        pb.size = ct.sizeof(_lib.IvpProblem)
        pb.B, pb.neq, pb.nt, pb.np, pb.n_events = B, n, nt, wl["p"], 1
        pb.t_span, pb.t_span_stride, pb.y0, pb.y0_stride = span.data_ptr(), 0, u0.data_ptr(), 0
        pb.data, pb.data_stride, pb.teval = P.data_ptr(), wl["p"] * 8, te.data_ptr()
        pb.rtol, pb.atol, pb.y_out = wl["rtol"], wl["atol"], usol.data_ptr()
        pb.result, pb.t_event, pb.y_event = res.data_ptr(), t_event.data_ptr(), y_event.data_ptr()

    def step():
        if ivp:
            _lib.check(_lib.lib.b200ode_solve_ivp_device(h.ptr, ct.byref(pb), ct.c_void_p(stream.cuda_stream)))
            return
        args = [h.ptr, B, n, u0.data_ptr(), 0, P.data_ptr(), wl["p"] * 8, wl["p"], nt,
                te.data_ptr(), usol.data_ptr(), wl["rtol"], wl["atol"], wl["mxstep"],
                ok.data_ptr(), cnt.data_ptr()]
        if solver == _lib.DOP853:
            rc = _lib.lib.b200ode_dop853_batched_device(*args, ct.c_void_p(stream.cuda_stream))
        else:
            rc = _lib.lib.b200ode_lsoda_batched_device(*args, 0, ct.c_void_p(stream.cuda_stream))
        _lib.check(rc)

    peak_tf, peak_src, peak_raw = (None, None, None)
    if rank == 0:
        peak_tf, peak_src, peak_raw = fp64_peak()

    def barrier():
        ranks.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 0)):
        step()
        flush.fill_(1)
    launches0 = h.kernel_info()["launches"]
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(a.steps)]
    t_wall0 = time.perf_counter()
    for k in range(a.steps):
        ev[k][0].record(stream)
        step()
        ev[k][1].record(stream)
        flush.fill_(k & 0xFF)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = h.kernel_info()["launches"] - launches0
    ms = [e0.elapsed_time(e1) for e0, e1 in ev]
    dev_s = sum(ms) / 1e3
    dev_s_max = ranks.max(dev_s)  # every multi-GPU time is the slowest rank's
    n_ok = int(ranks.sum(int(((res[:, 0] if ivp else ok) == 1).sum().item())))
    value = world * B * a.steps / dev_s_max

    # ---- end to end: the call a user makes -- numbalsoda_b200.lsoda / dop853 with host
    #      (numpy) inputs, returning usol / success on the host; H2D of the inputs, the
    #      kernels and D2H of every row are inside the timed region.  The result buffer is
    #      page-locked and reused across steps (out=): allocating and first-touching a
    #      fresh 2.5-25 GB array per call costs more than the integration itself.
    e2e = None
    if not a.no_e2e:
        u0h = initial_state(wl)
        # at most 8 GB of page-locked rows per rank (8 ranks pin their buffers on one host):
        # the same sweep, truncated; the rate does not depend on the count once there are a
        # few chunks in flight
        Be = int(min(B, max(1, (8 << 30) // (nt * n * 8))))
        Pe = np.ascontiguousarray(Ph[:Be])
        out = nb.pinned_empty((Be, nt, n), np.float64)
        api = nb.dop853 if solver == _lib.DOP853 else nb.lsoda
        rhs = nb.builtin(wl["rhs"])

        def e2e_step_ivp():
            sol = nb.solve_ivp(rhs, np.array([teh[0], teh[-1]]), u0h, teh, n_events=1,
                               event_fcn=nb.builtin(wl["events"]), data=Pe, rtol=wl["rtol"],
                               atol=wl["atol"], strict=a.strict, exact_ctrl=a.exact_ctrl, device=local,
                               out=out)
            return int(sol.success.sum())

        def e2e_step():
            if ivp:
                return e2e_step_ivp()
            us, okh = api(rhs, u0h, teh, Pe, wl["rtol"], wl["atol"], wl["mxstep"],
                          strict=a.strict, exact_ctrl=a.exact_ctrl, device=local, out=out)
            return int(okh.sum())  # the host reads the step's result

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        ke = max(1, min(a.steps, 3))
        n_ok_e2e = 0
        for _ in range(ke):
            n_ok_e2e = e2e_step()
        barrier()
        te2e = ranks.max(time.perf_counter() - t0)
        e2e = {"value": world * Be * ke / te2e, "unit": "traj/s",
               "h2d_bytes_per_step": int(Pe.nbytes + teh.nbytes + u0h.nbytes),
               "d2h_bytes_per_step": int(Be * nt * n * 8 + Be * (4 if not ivp else 48 + 8 + 8 * n)),
               "trajectories_per_gpu": Be,
               "steps": ke, "ms_per_step": 1e3 * te2e / ke, "success_fraction": n_ok_e2e / Be,
               "api": ("numbalsoda_b200.solve_ivp(rhs, t_span, y0, t_eval, n_events=1, event_fcn, "
                       "data[B,p], rtol, atol, out=pinned) -> ODEResultBatch on the host") if ivp else
                      "numbalsoda_b200.%s(rhs, u0, t_eval, data[B,p], rtol, atol, out=pinned) "
                      "-> (usol[B,nt,n], success[B]) on the host" % wl["solver"]}
        del out

    if rank != 0:
        return None

    if ivp:
        res_h = res.cpu().numpy()
        flops = ivp_flops(res_h, n, F_RHS[wl["rhs"]], 1)
        cnt_h = res_h[:, 8:9]  # nstep
        rows_h = float(res_h[:, 3].mean())
    else:
        cnt_h = cnt.cpu().numpy()
        flops = (dop853_flops if wl["solver"] == "dop853" else lsoda_flops)(cnt_h, n, F_RHS[wl["rhs"]], nt)
    per_launch_s = dev_s / a.steps
    achieved_tf = flops / per_launch_s / 1e12
    peaks = measured_peaks()
    out_bytes = B * (nt * n * 8 + (n + wl["p"]) * 8 + 4)
    if ivp:  # rows actually written (an event trims the rest), the result record, the event
        out_bytes = int(B * (rows_h * n * 8 + wl["p"] * 8 + 48 + 8 + 8 * n))
    tr = NCU_TRAFFIC.get(workload)
    roof = {"bound": "fp64", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
            "frac": achieved_tf / peak_tf if peak_tf else None,
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
    line = {"metric": "trajectories/s", "value": value, "unit": "traj/s", "n_gpus": world,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * dev_s_max / a.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg,
            "mode": "strict" if a.strict else ("FMA, exact controller" if a.exact_ctrl else "fast (FMA, fast controller)"),
            "success_fraction": n_ok / (B * world), "wall_s_timed_region": t_wall,
            "mean_steps_per_trajectory": float(cnt_h[:, 0].mean()),
            "roofline": roof, "clocks": clocks, "gpu_launches": launches,
            "kernel_info": h.kernel_info()}
    if e2e:
        line["e2e"] = e2e
    if not a.no_cpu:
        v, cores, kind, sample, secs = cpu_arm(wl, 12.0)
        line["cpu_baseline"] = {"value": v, "unit": "traj/s", "cores": cores, "kind": kind,
                                "sample": sample, "seconds": secs}
    if ivp:
        line["event_fraction"] = float(res_h[:, 5].mean())
        line["mean_rows_per_trajectory"] = rows_h
    del usol, cnt, ok, flush
    torch.cuda.empty_cache()
    return line


if __name__ == "__main__":
    main()
