#!/usr/bin/env bash
# smoke(), the C++ host self-test, the trig-flip record, then the launch list + one ncu --set full of
# the kernels of one C2 step.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r2_smoke.log
timeout 120 tests/host/host_selftest > $O/r2_host_selftest.log 2>&1; echo "host_selftest rc=$?"; tail -4 $O/r2_host_selftest.log
timeout 300 python tools/trig_flip.py > $O/r2_trig_flip.json 2> $O/r2_trig_flip.err; echo "trig rc=$?"; cat $O/r2_trig_flip.json
timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $O/r2_plain2.log 2>&1 &&
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_c2.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $O/r2_ncu_launch.log 2>&1
echo "launch list rc=$?"
