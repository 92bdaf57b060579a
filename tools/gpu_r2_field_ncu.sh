#!/usr/bin/env bash
# parity of the map build, rebuild timing, then ncu --set full of the field kernels (one slab each)
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_field_gpu.py "tests/test_batch_gpu.py::test_field_rebuild_8192_is_local_and_matches_oracle_windows" -x -q -m gpu 2>&1 | tail -5 | tee $O/r2_field2_tests.log
python tools/time_field.py --iters 15 2>&1 | tee $O/r2_field2_time.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"edt_rows_u8|edt_columns_u8|free_cells_u8" -s 60 -c 3 \
  -o $O/r2_field_kernels python tools/time_field.py --iters 2 > $O/r2_field_ncu_full.log 2>&1
echo "ncu rc=$?"; ls -la $O/r2_field_kernels.ncu-rep
