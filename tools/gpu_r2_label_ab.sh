#!/usr/bin/env bash
# A/B of the theta-run pre-linking of the labelling forest (QMCL_LINK_THETA) at C3 / C4 on one GPU
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_filter_gpu.py tests/test_sharded_gpu.py -x -q -m gpu 2>&1 | tail -3
for L in 0 1; do for W in C3 C4; do
  QMCL_LINK_THETA=$L timeout 300 python bench.py --workload $W --steps 10 --warmup 3 --breakdown --no-cpu-baseline --no-secondary > $O/r2_label_${L}_$W.json 2> $O/r2_label_${L}_$W.err
  echo "link=$L $W rc=$? $(python -c "import json;print(json.load(open('$O/r2_label_${L}_$W.json'))['ms_per_step'])") $(grep -E ' (bins|label) ' $O/r2_label_${L}_$W.err | tr -s ' ' | tr '\n' ' ')"
done; done
