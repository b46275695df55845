#!/bin/bash
# LSODA thread-kernel variants on the GPU box: rebuild lsoda_n3 with different block sizes /
# corrector iterations per pass / deferral on or off, time the Robertson workload.
# usage: tools/sweep_lsoda.sh "THREADS:MINBLOCKS:DEFS ..." [batch]
BATCH=${2:-1048576}
for cfg in $1; do
  IFS=: read th mb defs <<< "$cfg"
  B200ODE_REBUILD=lsoda_n3 B200ODE_THREADS=$th B200ODE_MINBLOCKS=$mb B200ODE_LSODA_DEFS="${defs//,/ }" python -m numbalsoda_b200.build > /dev/null 2>&1
  out=$(python bench.py --workload robertson_lsoda --steps 2 --warmup 1 --batch $BATCH --no-e2e --no-cpu --no-modes --builtin-rhs 2>&1 | tail -1)
  echo "threads=$th minblocks=$mb defs=$defs $(echo "$out" | python -c "import sys,json; j=json.loads(sys.stdin.read()); k=j['kernel_info']; print('traj/s=%.4g ms=%.2f regs=%d local=%d blocks/sm=%d smem=%d ok=%.3f' % (j['value'], j['ms_per_step'], k['registers'], k['local_bytes'], k['blocks_per_sm'], k['static_smem'], j['success_fraction']))" 2>&1)"
done
B200ODE_REBUILD=lsoda_n3 python -m numbalsoda_b200.build > /dev/null 2>&1
