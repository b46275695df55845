set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_solve_ivp.py -x -q > gpurun_out/pytest_ivp3.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_ivp3.log
timeout 300 python bench.py --workload lorenz_solve_ivp --steps 5 --warmup 3 > gpurun_out/IVP2_bench.json 2> gpurun_out/IVP2_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --workload lorenz_solve_ivp --steps 3 --warmup 3 --exact-ctrl --no-e2e --no-cpu > gpurun_out/IVP2_bench_exactctrl.json 2>> gpurun_out/IVP2_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --workload lorenz_solve_ivp --steps 3 --warmup 3 --strict --no-e2e --no-cpu > gpurun_out/IVP2_bench_strict.json 2>> gpurun_out/IVP2_bench.err; echo "bench rc=$?"
for f in gpurun_out/IVP2_bench*.json; do python -c "
import json,sys; j=json.load(open('$f')); print('$f', j['mode'], '%.4g'%j['value'], '%.3f'%j['roofline']['frac'], j.get('e2e',{}).get('value'), j.get('cpu_baseline',{}).get('value'))"; done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_lorenz_solve_ivp.csv python bench.py --workload lorenz_solve_ivp --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_l_ivp.log 2>&1; echo "ncu list rc=$?"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:b200ode_ivp_kernel -c 1 -o gpurun_out/prof_r1_ivp_v1 -f python bench.py --workload lorenz_solve_ivp --batch 262144 --steps 1 --warmup 0 --no-e2e --no-cpu > gpurun_out/ncu_ivp.log 2>&1; echo "ncu rc=$?"
