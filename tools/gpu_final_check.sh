#!/usr/bin/env bash
# Round-end rehearsal in one gpurun call: GPU suite, smoke, the driver's bench
# invocations (both arms), the other workloads, then the per-kernel ncu pass.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -x -q > $O/t_all.log 2>&1; echo "rc=$?" >> $O/t_all.log
tail -3 $O/t_all.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -2 $O/smoke.log
timeout 120 python bench.py --breakdown > $O/bench_c2_final.json 2> $O/bench_c2_final_stages.txt
timeout 120 python bench.py --impl reference > $O/bench_c2_reference_arm.json 2> $O/bench_c2_reference_arm.log
for w in C1 C3; do
  timeout 90 python bench.py --workload $w --no-cpu-baseline --breakdown > $O/bench_${w}_final.json 2> $O/bench_${w}_final_stages.txt
done
python - <<'PY'
import json
for n in ["c2_final", "c2_reference_arm", "C1_final", "C3_final"]:
    try:
        d = json.load(open(f"gpurun_out/bench_{n}.json"))
        print(n, "value %.4g" % d["value"], "ms/step %.4f" % d["ms_per_step"], "e2e", d["e2e"]["value"])
    except Exception as e:
        print(n, "ERR", e)
PY
bash tools/gpu_ncu_steps.sh
