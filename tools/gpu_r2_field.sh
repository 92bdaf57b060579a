#!/usr/bin/env bash
set -u
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
python tools/time_field.py --iters 9 > $O/r2_field_time.txt 2>&1; cat $O/r2_field_time.txt
python - <<'PY' 2>&1 | tee -a gpurun_out/r2_field_time.txt
import torch, time, numpy as np
a = torch.empty(8192*8192, dtype=torch.uint8).pin_memory()
d = torch.empty(8192*8192, dtype=torch.uint8, device="cuda")
for _ in range(3): d.copy_(a, non_blocking=True); torch.cuda.synchronize()
ts=[]
for _ in range(9):
    t0=time.perf_counter(); d.copy_(a, non_blocking=True); torch.cuda.synchronize(); ts.append(time.perf_counter()-t0)
print("pinned H2D of 67 MB alone: %.3f ms (%.1f GB/s)" % (1e3*np.median(ts), 67.1/np.median(ts)/1e3))
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r2_field_launches.csv python tools/time_field.py --iters 2 > $O/r2_field_ncu.log 2>&1
echo "ncu rc=$?"
python tools/launch_summary.py $O/r2_field_launches.csv | head -14
