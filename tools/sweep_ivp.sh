#!/bin/bash
# Time prebuilt variants of the solve_ivp kernel (tools/variants/libb200ode_<cfg>.so, built in the
# CPU container with B200ODE_THREADS / B200ODE_MINBLOCKS) on the GPU box.
# usage: tools/sweep_ivp.sh [batch]
BATCH=${1:-1048576}
for lib in numbalsoda_b200/libb200ode.so tools/variants/*.so; do
  out=$(B200ODE_LIB=$PWD/$lib python bench.py --workload lorenz_solve_ivp --steps 3 --warmup 2 --batch $BATCH --no-e2e --no-cpu 2>&1 | tail -1)
  echo "$lib $(echo "$out" | python -c "import sys,json; j=json.loads(sys.stdin.read()); k=j['kernel_info']; print('traj/s=%.4g ms=%.2f frac=%.3f regs=%d local=%d threads=%d blocks/sm=%d ok=%.3f' % (j['value'], j['ms_per_step'], j['roofline']['frac'], k['registers'], k['local_bytes'], k['threads_per_block'], k['blocks_per_sm'], j['success_fraction']))" 2>&1)"
done
