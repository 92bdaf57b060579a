#!/usr/bin/env bash
# K7 variant A/B (tools/k7_ab.py per variant library) + the sensor parity tests on the current build.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
: > $O/r2_k7_ab.jsonl
for rep in 1 2; do
for v in ${K7_VARIANTS:-prev cur floor wide floorwide}; do
  QMCL_B200_LIB=$PWD/quickmcl_b200/variants/lib$v.so timeout 120 python tools/k7_ab.py >> $O/r2_k7_ab.jsonl 2>> $O/r2_k7_ab.err
done
done
cat $O/r2_k7_ab.jsonl
timeout 600 python -m pytest tests/test_sensor_gpu.py tests/test_fullsize_gpu.py tests/test_filter_gpu.py -m gpu -x -q > $O/r2_t_sensor.log 2>&1; echo "rc=$?" >> $O/r2_t_sensor.log
tail -5 $O/r2_t_sensor.log
