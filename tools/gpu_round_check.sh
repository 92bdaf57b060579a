#!/usr/bin/env bash
# One gpurun call: parity tests of the changed paths first, then the bench
# lines (deferred vs per-value synchronisation), then the whole GPU suite and
# the ncu launch list.  Every step writes its own file under gpurun_out/ so a
# call that is cut short still leaves what finished.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
t0=$(date +%s)
stamp() { echo "[$(( $(date +%s) - t0 )) s] $*" | tee -a $O/check_timeline.txt; }

stamp "filter/cycle/golden tests (deferred read-backs)"
timeout 300 python -m pytest tests/test_filter_gpu.py tests/test_cycle.py tests/test_golden_fixtures.py \
  -m gpu -x -q > $O/t_filter_lazy.log 2>&1; echo "rc=$?" >> $O/t_filter_lazy.log
tail -3 $O/t_filter_lazy.log

stamp "bench C2 (default, deferred)"
timeout 240 python bench.py --breakdown > $O/bench_c2_lazy.json 2> $O/bench_c2_lazy_stages.txt
stamp "bench C2 (QMCL_EAGER_SYNC=1)"
QMCL_EAGER_SYNC=1 timeout 120 python bench.py --no-cpu-baseline --breakdown > $O/bench_c2_eager.json 2> $O/bench_c2_eager_stages.txt
stamp "bench C1 / C3 (deferred)"
timeout 120 python bench.py --workload C1 --no-cpu-baseline --breakdown > $O/bench_c1_lazy.json 2> $O/bench_c1_lazy_stages.txt
timeout 120 python bench.py --workload C3 --no-cpu-baseline --breakdown > $O/bench_c3_lazy.json 2> $O/bench_c3_lazy_stages.txt
head -c 600 $O/bench_c2_lazy.json; echo
head -c 300 $O/bench_c2_eager.json; echo

stamp "whole GPU suite"
timeout 600 python -m pytest tests -m gpu -x -q > $O/t_all.log 2>&1; echo "rc=$?" >> $O/t_all.log
tail -3 $O/t_all.log

stamp "smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log
tail -2 $O/smoke.log

stamp "ncu launch list (C2)"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
  --log-file $O/launches_c2.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_launches.log 2>&1
stamp "C5 field + batch"
timeout 200 python tools/time_c5.py > $O/c5.json 2> $O/c5.log
stamp "done"
