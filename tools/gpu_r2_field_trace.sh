#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_field_gpu.py -x -q -m gpu 2>&1 | tail -3 | tee $O/r2_field2_tests.log
: > $O/r2_field_trace.txt
for ns in 2 3; do
echo "streams $ns" | tee -a $O/r2_field_trace.txt
QMCL_MAP_STREAMS=$ns QMCL_MAP_TRACE=1 python tools/time_field.py --iters 3 2>&1 | tail -20 | tee -a $O/r2_field_trace.txt
for i in 1 2; do QMCL_MAP_STREAMS=$ns python tools/time_field.py --iters 21 2>&1 | tee -a $O/r2_field_trace.txt; done
done
