#!/usr/bin/env bash
# Diagnostic: how much of the tail is cold-L2 instruction / data fetch?  Same bench with and without the L2 flush.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
rm -f $O/r2_ab_*.json
run() { # name env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 100 --warmup 10 > $O/r2_ab_$name.json 2> $O/r2_ab_$name.err || echo "$name failed"
}
run old_flush QMCL_B200_LIB=$PWD/quickmcl_b200/variants/libold.so
run old_noflush QMCL_B200_LIB=$PWD/quickmcl_b200/variants/libold.so QMCL_BENCH_NO_FLUSH=1
run new_flush X=1
run new_noflush QMCL_BENCH_NO_FLUSH=1
python tools/tail_phases.py $O/r2_ab_*.json
