#!/usr/bin/env bash
# `gpurun --gpus N -- 'N=<n> bash tools/gpu_r2_scale.sh'`: the driver's sharded bench invocation (C4 over
# N GPUs, one-GPU anchor and NCCL parity check in the same job, the KLD variant in `secondary`) with the
# per-rank stage breakdown on stderr.
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
N=${N:-2}
S=$(date +%s)
timeout ${T:-240} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps ${STEPS:-30} --warmup ${WARMUP:-5} --breakdown --no-cpu-baseline \
  > $O/r2_bench_c4_${N}gpu.json 2> $O/r2_bench_c4_${N}gpu.err
echo "rc=$? in $(( $(date +%s) - S )) s"
python - <<PY
import json
try:
    d = json.load(open("$O/r2_bench_c4_${N}gpu.json"))
    print("C4 x$N: ms/step %.3f median %.3f e2e %.3f | one GPU %.3f -> x%.2f | parity %s" % (
        d["ms_per_step"], d["update_latency_median_ms"], d["e2e"]["ms_per_step"],
        d["same_workload_on_one_gpu"]["ms_per_step"], d["same_workload_on_one_gpu"]["speedup_of_this_run"],
        d["parity_vs_one_gpu"]))
    print("stage_ms", {k: round(v, 4) for k, v in d["stage_ms"].items()})
    print("exchanges", d.get("exchanges"))
    for k, v in d.get("secondary", {}).items():
        print(k, {a: (round(b, 4) if isinstance(b, float) else b) for a, b in v.items()})
except Exception as e:
    print("ERR", e)
PY
grep -A14 "stage breakdown of rank 0" $O/r2_bench_c4_${N}gpu.err | head -16
tail -3 $O/r2_bench_c4_${N}gpu.err
