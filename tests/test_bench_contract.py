"""bench.py's output contract, checked on the arm that runs without a GPU
(`--impl reference`: the reference's CPU implementation of the path on the host cores)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "1", "--warmup", "0", *args], capture_output=True, text=True,
                         timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1  # ONE JSON line
    return json.loads(lines[0])


def test_reference_arm_line_robertson():
    j = _run("--workload", "robertson_lsoda")
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
              "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "impl",
              "cpu_baseline", "e2e"):
        assert k in j, k
    assert j["impl"] == "reference" and j["metric"] == "trajectories/s" and j["unit"] == "traj/s"
    assert j["higher_is_better"] is True and j["scaling"] == "weak" and j["vs_baseline"] is None
    assert j["dtype"] == "f64" and j["data"] == "synthetic" and "workload" in j["config"]
    cb = j["cpu_baseline"]
    assert cb["kind"] == "reference"  # the unmodified reference C++ (oracle/_ref) for LSODA
    assert cb["cores"] >= 1 and cb["value"] == j["value"] and "trajectories" in cb["sample"]
    assert j["e2e"] == {"value": j["value"], "unit": "traj/s", "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}
    assert j["value"] > 100.0


def test_reference_arm_dop853_is_labelled_a_port():
    j = _run("--workload", "lorenz_dop853")
    assert j["cpu_baseline"]["kind"] == "port"  # the Fortran reference cannot be built here
    assert j["value"] > 10.0


def test_product_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        return
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0 and "no CUDA device" in (out.stderr + out.stdout)
