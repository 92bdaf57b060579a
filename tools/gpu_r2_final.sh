#!/usr/bin/env bash
# Round-2 closing call on one B200: the GPU suite, smoke(), the driver's two bench invocations
# (own arm with `secondary`, reference arm), then the ncu launch list of the bench command.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
timeout 420 python -m pytest tests -m gpu -x -q --timeout 180 --durations=8 > $O/r2_gpu_tests_final.log 2>&1
echo "rc=$?" >> $O/r2_gpu_tests_final.log
tail -4 $O/r2_gpu_tests_final.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1; echo "rc=$?" >> $O/r2_smoke.log
tail -2 $O/r2_smoke.log
S=$(date +%s)
timeout 300 python bench.py > $O/r2_bench_c2_final.json 2> $O/r2_bench_c2_final.err; echo "bench rc=$? in $(( $(date +%s) - S )) s"
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2_bench_c2_final.json"))
    r = d["roofline"]
    print("C2 ms/step %.4f median %.4f e2e %.4f launches/step %.1f | K7 %.4f ms frac %.3f %.3g evals/s" % (
        d["ms_per_step"], d["update_latency_median_ms"], d["e2e"]["ms_per_step"], d["gpu_launches_per_step"],
        r["kernel_ms"], r["frac"], r["kernel_evals_per_s"]))
    print("stage_ms", d["stage_ms"])
    print("tail_phases_us", d["tail_phases_us"])
    for k, v in d.get("secondary", {}).items():
        print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()
                  if a not in ("tail_phases_us", "workload")})
    print("cpu", d["cpu_baseline"]["value"], "clocks", d["clocks"])
except Exception as e:
    print("ERR", e)
PY
timeout 120 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2_bench_c2_reference_arm.json 2> $O/r2_ref.err; echo "ref rc=$?"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2_launches_c2_final.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > $O/r2_ncu_launch.log 2>&1
echo "launch list rc=$?"
python tools/launch_summary.py $O/r2_launches_c2_final.csv 2>/dev/null | head -20 | tee $O/r2_launches_c2_final_summary.txt
