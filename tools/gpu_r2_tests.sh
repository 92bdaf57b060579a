#!/usr/bin/env bash
# GPU suite with a per-test timeout (a hung cooperative kernel must not eat the call), then the bench.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
timeout ${T_TESTS:-900} python -m pytest tests -m gpu -x -q --timeout 180 ${PYTEST_ARGS:-} > $O/r2_t_all.log 2>&1; echo "rc=$?" >> $O/r2_t_all.log
tail -25 $O/r2_t_all.log
if [ "${RUN_BENCH:-1}" = 1 ]; then
  timeout 300 python bench.py --no-cpu-baseline ${BENCH_ARGS:-} > $O/r2_bench_c2.json 2> $O/r2_bench_c2.err; echo "bench rc=$?"
  cat $O/r2_bench_c2.json | cut -c1-1500
  tail -5 $O/r2_bench_c2.err
fi
