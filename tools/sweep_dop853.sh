#!/bin/bash
# DOP853 thread-kernel variants on the GPU box: rebuild dop853_n3 with different geometry / build
# switches, time the Lorenz workload; with NCU=1 also the DRAM bytes of one launch.
# usage: tools/sweep_dop853.sh "THREADS:MINBLOCKS:QCAP:DEFS ..." [batch]
BATCH=${2:-1048576}
for cfg in $1; do
  IFS=: read th mb qc defs <<< "$cfg"
  B200ODE_REBUILD=dop853_n3 B200ODE_THREADS=$th B200ODE_MINBLOCKS=$mb B200ODE_QCAP=$qc B200ODE_DOP853_DEFS="${defs//,/ }" python -m numbalsoda_b200.build > /dev/null 2>&1
  out=$(python bench.py --steps 3 --warmup 1 --batch $BATCH --no-e2e --no-cpu --no-modes --no-secondary --builtin-rhs 2>&1 | tail -1)
  echo "threads=$th minblocks=$mb qcap=$qc defs=$defs $(echo "$out" | python -c "import sys,json; j=json.loads(sys.stdin.read()); k=j['kernel_info']; print('traj/s=%.4g ms=%.2f frac=%.3f regs=%d local=%d blocks/sm=%d ok=%.3f' % (j['value'], j['ms_per_step'], j['roofline']['frac'], k['registers'], k['local_bytes'], k['blocks_per_sm'], j['success_fraction']))" 2>&1)"
  if [ -n "$NCU" ]; then
    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:b200ode_dop853_kernel -s 1 -c 1 --csv python bench.py --steps 1 --warmup 1 --batch 262144 --no-e2e --no-cpu --no-modes --no-secondary --builtin-rhs 2>/dev/null | grep -E "dram__bytes|gpu__time" | awk -F'","' '{print "    " $(NF-2), $(NF-1), $NF}'
  fi
done
B200ODE_REBUILD=dop853_n3 python -m numbalsoda_b200.build > /dev/null 2>&1
