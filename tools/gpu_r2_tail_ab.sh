#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
rm -f $O/r2_ab_*.json
run() { # name env...
  local name=$1; shift
  env "$@" timeout 200 python bench.py --no-cpu-baseline --steps 40 --warmup 5 ${EXTRA_ARGS:-} > $O/r2_ab_$name.json 2> $O/r2_ab_$name.err || echo "$name failed"
}
run base X=1
run inline QMCL_B200_LIB=$PWD/quickmcl_b200/variants/libinline.so
run inline2 QMCL_B200_LIB=$PWD/quickmcl_b200/variants/libinline.so QMCL_TAIL_BLOCKS_PER_SM=2
python tools/tail_phases.py $O/r2_ab_*.json
