#!/bin/bash
# Build-parameter sweep on the GPU box: rebuild the kernels with different thread /
# register / queue budgets and time a workload.
# usage: tools/sweep_build.sh "THREADS:MINBLOCKS:QCAP:NOINLINE ..." [batch] [workload]
set -u
BATCH=${2:-524288}
WL=${3:-lorenz_dop853}
for cfg in $1; do
  IFS=: read th mb qc ni <<< "$cfg"
  B200ODE_THREADS=$th B200ODE_MINBLOCKS=$mb B200ODE_QCAP=${qc:-55} B200ODE_DRAIN_NOINLINE=${ni:-0} python -m numbalsoda_b200.build --force > /dev/null 2>&1
  out=$(python bench.py --workload $WL --steps 2 --warmup 1 --batch $BATCH --no-e2e --no-cpu 2>&1 | tail -1)
  echo "$WL threads=$th minblocks=$mb qcap=${qc:-55} noinline=${ni:-0} $(echo "$out" | python -c "import sys,json; j=json.loads(sys.stdin.read()); k=j['kernel_info']; print('traj/s=%.4g ms=%.2f frac=%.3f regs=%d local=%d blocks/sm=%d smem=%d ok=%.3f' % (j['value'], j['ms_per_step'], j['roofline']['frac'], k['registers'], k['local_bytes'], k['blocks_per_sm'], k['static_smem'], j['success_fraction']))" 2>&1)"
done
python -m numbalsoda_b200.build --force > /dev/null 2>&1
