set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_solve_ivp.py -x -q > gpurun_out/pytest_ivp1.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_ivp1.log
timeout 300 python bench.py --workload lorenz_solve_ivp --steps 3 --warmup 3 > gpurun_out/IVP1_bench.json 2> gpurun_out/IVP1_bench.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/IVP1_bench.json; tail -5 gpurun_out/IVP1_bench.err
timeout 200 compute-sanitizer --tool memcheck python tools/sanitize_small.py ivp > gpurun_out/sanitize_ivp.log 2>&1; echo "sanitize rc=$?"; tail -6 gpurun_out/sanitize_ivp.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:b200ode_ivp_kernel -c 1 -o gpurun_out/prof_r1_ivp_v0 -f python bench.py --workload lorenz_solve_ivp --batch 262144 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_ivp.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/ncu_ivp.log
