#!/bin/bash
# Time prebuilt variants of libb200ode.so (tools/variants/*.so, built in the CPU container) on a workload.
# usage: tools/sweep_libs.sh workload [batch]
WL=${1:-brusselator_lsoda}
BATCH=${2:-0}
for lib in numbalsoda_b200/libb200ode.so tools/variants/*.so; do
  out=$(B200ODE_LIB=$PWD/$lib python bench.py --workload $WL --steps 2 --warmup 1 --batch $BATCH --no-e2e --no-cpu 2>&1 | tail -1)
  echo "$WL $lib $(echo "$out" | python -c "import sys,json; j=json.loads(sys.stdin.read()); k=j['kernel_info']; print('traj/s=%.4g ms=%.2f frac=%.3f regs=%d smem=%d blocks/sm=%d ok=%.3f' % (j['value'], j['ms_per_step'], j['roofline']['frac'], k['registers'], k['static_smem'], k['blocks_per_sm'], j['success_fraction']))" 2>&1)"
done
