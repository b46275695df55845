set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu13.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_gpu13.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke4.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke4.log
timeout 300 python bench.py --workload brusselator_lsoda --steps 3 --warmup 3 > gpurun_out/FINAL4_brusselator_lsoda.json 2> gpurun_out/FINAL4_brusselator_lsoda.err; echo "bench rc=$?"
python -c "
import json,sys; j=json.load(open('gpurun_out/FINAL4_brusselator_lsoda.json')); print(j['mode'], '%.4g'%j['value'], '%.3f'%j['roofline']['frac'], j.get('e2e',{}).get('value'), j.get('cpu_baseline',{}).get('value'))"
