#!/usr/bin/env bash
# Per-kernel DRAM / L2 / SM figures for every kernel of one update step: a
# metrics-only ncu pass (a few replays per launch, small CSV), C3 then C2.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size,launch__registers_per_thread
for w in ${NCU_WORKLOADS:-C3 C2}; do
  timeout ${NCU_TIMEOUT:-70} ncu --metrics $M --clock-control none --launch-skip 48 --launch-count 64 \
    --page raw --csv --log-file $O/step_$w.csv \
    python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline > $O/ncu_step_$w.log 2>&1
  echo "$w rc=$?"; ls -la $O/step_$w.csv
done
