#!/bin/bash
# Build-parameter sweep on the GPU box: rebuild the kernels with different thread /
# register / queue budgets and time the default workload.
# usage: tools/sweep_build.sh "THREADS:MINBLOCKS:QCAP:NOINLINE ..." [batch]
set -u
BATCH=${2:-524288}
for cfg in $1; do
  IFS=: read th mb qc ni <<< "$cfg"
  B200ODE_THREADS=$th B200ODE_MINBLOCKS=$mb B200ODE_QCAP=$qc B200ODE_DRAIN_NOINLINE=$ni python -m numbalsoda_b200.build --force > /dev/null 2>&1
  out=$(python bench.py --steps 2 --warmup 1 --batch $BATCH --no-e2e --no-cpu 2>&1 | tail -1)
  echo "threads=$th minblocks=$mb qcap=$qc noinline=$ni $(echo "$out" | python -c "import sys,json; j=json.loads(sys.stdin.read()); k=j['kernel_info']; print('traj/s=%.4g ms=%.2f frac=%.3f regs=%d local=%d blocks/sm=%d smem=%d ok=%.3f' % (j['value'], j['ms_per_step'], j['roofline']['frac'], k['registers'], k['local_bytes'], k['blocks_per_sm'], k['static_smem'], j['success_fraction']))" 2>&1)"
done
python -m numbalsoda_b200.build --force > /dev/null 2>&1
