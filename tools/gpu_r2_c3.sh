#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
for W in C3 C4; do
  timeout 300 python bench.py --workload $W --steps 20 --warmup 3 --breakdown --no-cpu-baseline --no-secondary > $O/r2_bd2_$W.json 2> $O/r2_bd2_$W.err
  echo "$W rc=$?"; grep -A20 "stage breakdown" $O/r2_bd2_$W.err
  python -c "import json;d=json.load(open('$O/r2_bd2_$W.json'));print(d['ms_per_step'], d['stage_ms'])"
done
