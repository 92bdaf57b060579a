#!/usr/bin/env bash
# A/B of the running-sum phases: previous library vs current (+ literal crossing chunks) + K9 / filter tests + resolve trace.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
rm -f $O/r2_ab_*.json
timeout 900 python -m pytest tests/test_prefix_gpu.py tests/test_filter_gpu.py tests/test_batch_gpu.py tests/test_sharded_gpu.py -m gpu -x -q > $O/r2_t_prefix.log 2>&1
echo "pytest rc=$?"; tail -3 $O/r2_t_prefix.log
run() { # name env...
  local name=$1; shift
  env "$@" timeout 300 python bench.py --no-cpu-baseline --steps 100 --warmup 10 ${EXTRA_ARGS:-} > $O/r2_ab_$name.json 2> $O/r2_ab_$name.err || echo "$name failed"
}
[ -f quickmcl_b200/variants/libold.so ] && run old QMCL_B200_LIB=$PWD/quickmcl_b200/variants/libold.so
run new X=1
python tools/tail_phases.py $O/r2_ab_*.json
python - <<'P'
import json
for n in ("old","new"):
    try:
        d=json.load(open(f"gpurun_out/r2_ab_{n}.json")); s=d.get("secondary",{})
        print(n, "C1", s.get("C1",{}).get("ms_per_step"), "C1 resolve", s.get("C1",{}).get("tail_phases_us",{}).get("resolve"), "C3", s.get("C3",{}).get("ms_per_step"), "batch", s.get("C5_batch",{}).get("ms_per_batch_update_e2e"))
    except Exception as e: print(n, e)
P
for m in 0 2; do
  QMCL_B200_LIB=$PWD/quickmcl_b200/variants/libtrace.so timeout 200 python tools/resolve_trace.py C2 $m > $O/r2_resolve_trace_c2_mode$m.txt 2>&1 || echo "trace $m failed"
done
