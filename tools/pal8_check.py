"""PAL8 (one-byte palette map, qmcl_map_set_sensor_path(3)) against PAL32 on the
C2 inputs: identical weights bit for bit, and the sensor kernel's time for both
(CUDA events recorded by the library).  No torch import: a few seconds in all."""
import os
import sys
import json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import quickmcl_b200 as q
from quickmcl_b200 import synthetic as syn

out = {}
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
    got = {}
    for path in (4, 3, 4, 3):
        gm.set_sensor_path(path)
        f.set_particles(init)
        f.update_importance_from_observations(scan)       # warm-up (builds the palette once)
        w = f.get_particles()["weight"].copy()
        f.set_timing(1)
        f.get_timing(reset=True)
        for _ in range(10):
            f.set_particles(init)
            f.update_importance_from_observations(scan)
        ms, cnt = f.get_timing(reset=True)["sensor"]
        f.set_timing(0)
        got.setdefault(path, []).append((ms / max(cnt, 1), w))
    same = got[4][0][1].tobytes() == got[3][0][1].tobytes()
    evals = n * len(scan)
    out[name] = {"weights_identical": bool(same),
                 "pal32_ms": [round(v[0], 5) for v in got[4]],
                 "pal8_ms": [round(v[0], 5) for v in got[3]],
                 "pal32_evals_per_s": evals / (min(v[0] for v in got[4]) * 1e-3),
                 "pal8_evals_per_s": evals / (min(v[0] for v in got[3]) * 1e-3),
                 "max_abs_diff": float(np.max(np.abs(got[4][0][1] - got[3][0][1])))}
print(json.dumps(out, indent=1))
