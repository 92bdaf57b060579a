#!/usr/bin/env bash
# Lean gpurun call: the whole GPU suite, then the C2/C1/C3 bench lines.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -m gpu -x -q > $O/t_all.log 2>&1; echo "rc=$?" >> $O/t_all.log
tail -4 $O/t_all.log
timeout 240 python bench.py --no-cpu-baseline --breakdown > $O/bench_c2.json 2> $O/bench_c2_stages.txt
for w in ${QUICK_WORKLOADS:-C1 C3}; do
  timeout 120 python bench.py --workload $w --no-cpu-baseline --breakdown > $O/bench_$w.json 2> $O/bench_${w}_stages.txt
done
python - <<'PY'
import json
for n in ["c2", "C1", "C3"]:
    try:
        d = json.load(open(f"gpurun_out/bench_{n}.json"))
        print(n, "ms/step %.4f" % d["ms_per_step"], "e2e %.4f" % d["e2e"]["ms_per_step"],
              "K7 %.4f ms frac %.3f" % (d["roofline"]["kernel_ms"], d["roofline"]["frac"]))
    except Exception as e:
        print(n, "ERR", e)
PY
cat $O/bench_c2_stages.txt | tail -13
