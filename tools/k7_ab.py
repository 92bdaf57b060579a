"""A/B of sensor-kernel (K7) build variants: QMCL_B200_LIB=<variant .so> python tools/k7_ab.py
Prints the kernel time (CUDA events recorded by the library around K7 alone, warm L2) on the
C2 inputs (100 k tracking particles) and on 1 M global particles, and a digest of the weights
so that variants can be compared bit for bit across processes.  No torch import."""
import hashlib
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import quickmcl_b200 as q
from quickmcl_b200 import _abi, synthetic as syn

out = {"lib": os.path.basename(_abi.LIB_PATH)}
occ, res, origin = syn.make_map(2000)
gm = q.create_likelihood_map(device=0)
gm.set_parameters(q.LikelihoodMapParameters(**dict(syn.NODE_LIKELIHOOD, num_beams=0)))
gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
pose = syn.pick_truth_pose(occ, res, origin)
scan = syn.make_scan(occ, res, origin, pose, 1081, 270.0)
for name, (x, y, t) in (("tracking_100k", syn.gaussian_cloud(pose, 100_000)),
                        ("global_1M", syn.uniform_cloud(occ, res, origin, 1_000_000))):
    n = len(x)
    init = q.make_particles(x, y, t, np.full(n, 1.0 / n))
    f = q.create_particle_filter(gm, seed=3)
    f.set_particles(init)
    f.update_importance_from_observations(scan)       # warm-up (builds the palette once)
    w = f.get_particles()["weight"].copy()
    times = []
    for rep in range(3):
        f.set_timing(1)
        f.get_timing(reset=True)
        for _ in range(10):
            f.set_particles(init)
            f.update_importance_from_observations(scan)
        ms, cnt = f.get_timing(reset=True)["sensor"]
        f.set_timing(0)
        times.append(ms / max(cnt, 1))
    evals = n * len(scan)
    out[name] = {"k7_ms": [round(v, 5) for v in times], "evals_per_s": evals / (min(times) * 1e-3),
                 "frac_of_6548": evals * 4.02 / (min(times) * 1e-3) / 6548.2e9,
                 "weights_sha1": hashlib.sha1(w.tobytes()).hexdigest()[:16]}
print(json.dumps(out))
