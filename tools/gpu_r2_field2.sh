#!/usr/bin/env bash
# Round-2 field rebuild check: parity tests of the map build, then timings of the
# 8192^2 rebuild for several slab sizes, then the per-kernel launch list.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_field_gpu.py "tests/test_batch_gpu.py::test_field_rebuild_8192_is_local_and_matches_oracle_windows" -x -q -m gpu 2>&1 | tail -15 | tee $O/r2_field2_tests.log
: > $O/r2_field2_time.txt
for cfg in "512 128" "256 128" "1024 128" "512 64" "512 512" "384 96"; do
  set -- $cfg
  echo "slab_rows=$1 tail_rows=$2" | tee -a $O/r2_field2_time.txt
  QMCL_MAP_SLAB_ROWS=$1 QMCL_MAP_TAIL_ROWS=$2 python tools/time_field.py --iters 15 2>&1 | tee -a $O/r2_field2_time.txt
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_field2_launches.csv python tools/time_field.py --iters 2 > $O/r2_field2_ncu.log 2>&1
echo "ncu rc=$?"
python tools/launch_summary.py $O/r2_field2_launches.csv | head -14 | tee $O/r2_field2_launch_summary.txt
