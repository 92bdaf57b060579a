#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out; mkdir -p $O
for m in 0 2; do
  QMCL_B200_LIB=$PWD/quickmcl_b200/variants/libtrace.so timeout 200 python tools/resolve_trace.py C2 $m > $O/r2_resolve_trace_c2_mode$m.txt 2>&1 || echo "trace $m failed"
done
QMCL_B200_LIB=$PWD/quickmcl_b200/variants/libtrace.so timeout 200 python tools/resolve_trace.py C1 0 > $O/r2_resolve_trace_c1_mode0.txt 2>&1
tail -5 $O/r2_resolve_trace_c2_mode0.txt
