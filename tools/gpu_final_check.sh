set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu12.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu12.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke3.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke3.log
timeout 300 python bench.py --workload lorenz_solve_ivp --steps 5 --warmup 3 > gpurun_out/IVP3_bench.json 2> gpurun_out/IVP3_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/FINAL3_lorenz_dop853.json 2> gpurun_out/FINAL3_lorenz_dop853.err; echo "bench rc=$?"
for f in gpurun_out/IVP3_bench.json gpurun_out/FINAL3_lorenz_dop853.json; do python -c "
import json,sys; j=json.load(open('$f')); print('$f', j['mode'], '%.4g'%j['value'], '%.3f'%j['roofline']['frac'], j['roofline']['traffic'], j.get('e2e',{}).get('value'), j.get('cpu_baseline',{}).get('value'))"; done
