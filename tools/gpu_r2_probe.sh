#!/usr/bin/env bash
# Round 2, call 1: latency probes, gather peaks, ncu --set full of the sensor kernel that bench.py times.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/r2_smi.txt 2>&1
timeout 120 tools/probe_sync > $O/r2_probe_sync.txt 2>&1; echo "probe_sync rc=$?"
timeout 120 tools/probe_gather $O/r2_probe_gather.json > $O/r2_probe_gather.txt 2>&1; echo "probe_gather rc=$?"
cat $O/r2_probe_sync.txt $O/r2_probe_gather.txt
timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2_plain.log 2>&1 &&
timeout 400 ncu --set full --clock-control none --import-source on -k regex:sensor_kernel -s 3 -c 2 \
  -o $O/r2_sensor_pal8 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/r2_ncu.log 2>&1
echo "ncu rc=$?"; ls -la $O/*.ncu-rep
