set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_solve_ivp.py -x -q > gpurun_out/pytest_ivp2.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_ivp2.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke2.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke2.log
bash tools/sweep_ivp.sh > gpurun_out/sweep_ivp1.log 2>&1; cat gpurun_out/sweep_ivp1.log
