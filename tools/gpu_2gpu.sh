set -x
mkdir -p gpurun_out
for wl in lorenz_dop853 robertson_lsoda lorenz_solve_ivp; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --workload $wl --no-cpu > gpurun_out/G2_$wl.json 2> gpurun_out/G2_$wl.err; echo "$wl rc=$?"
  tail -1 gpurun_out/G2_$wl.json | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print(j['n_gpus'], '%.4g'%j['value'], '%.3f'%j['roofline']['frac'], j.get('e2e',{}).get('value'))"
done
