#!/usr/bin/env bash
# ncu --set full of the three kernels of a C2 update (front, sensor, tail), after the same command ran clean.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary"
timeout 200 $CMD > $O/r2_plain3.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sensor_kernel|front_kernel|tail_grid_kernel" -s 9 -c 6 \
  -o $O/r2_c2_kernels $CMD > $O/r2_ncu3.log 2>&1
echo "ncu rc=$?"; ls -la $O/r2_c2_kernels.ncu-rep; tail -3 $O/r2_ncu3.log
