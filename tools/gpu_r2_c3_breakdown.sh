#!/usr/bin/env bash
# per-stage and per-kernel timings of the >= 1 M-particle configurations on one GPU
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
for W in C3 C4; do
  timeout 300 python bench.py --workload $W --steps 10 --warmup 3 --breakdown --no-cpu-baseline --no-secondary > $O/r2_bd_$W.json 2> $O/r2_bd_$W.err
  echo "$W rc=$?"; grep -A20 "stage breakdown" $O/r2_bd_$W.err
  python -c "import json;d=json.load(open('$O/r2_bd_$W.json'));print(d['ms_per_step'], d['stage_ms'])"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv --log-file $O/r2_c3_launches.csv python bench.py --workload C3 --steps 3 --warmup 2 --no-cpu-baseline --no-secondary > $O/r2_c3_ncu.log 2>&1
echo "ncu rc=$?"
python tools/launch_summary.py $O/r2_c3_launches.csv | head -40 | tee $O/r2_c3_launch_summary.txt
