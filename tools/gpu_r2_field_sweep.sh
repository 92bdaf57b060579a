#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
: > $O/r2_field_sweep.txt
for cfg in "512 128" "1024 128" "768 128" "512 256" "1024 256" "2048 128" "640 160"; do set -- $cfg
  for i in 1 2; do echo -n "slab $1 tail $2: " | tee -a $O/r2_field_sweep.txt; QMCL_MAP_SLAB_ROWS=$1 QMCL_MAP_TAIL_ROWS=$2 python tools/time_field.py --iters 21 2>&1 | cut -c1-60 | tee -a $O/r2_field_sweep.txt; done
done
