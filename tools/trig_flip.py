"""How often the correctly rounded heading trig (documented deviation 3) picks a
different map cell than the host libm's sinf / cosf the reference calls: per-particle weights
of the CUDA path against the oracle in TRIG_LIBM mode.  Prints one JSON object."""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import quickmcl_b200 as q
from quickmcl_b200 import synthetic as syn
from oracle import binding as ob

out = {}
for width, n, cloud in ((400, 20000, "tracking"), (2000, 50000, "tracking"), (2000, 50000, "global")):
    occ, res, origin = syn.make_map(width)
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=0)
    om = ob.OracleMap(**lp)
    om.new_map(occ, res, origin, field_mode=ob.FIELD_EXACT_EDT)
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(**lp))
    gm.load_field(om.field(), om.state(), res, origin)
    pose = syn.pick_truth_pose(occ, res, origin)
    scan = syn.make_scan(occ, res, origin, pose, 1081, 270.0)
    x, y, t = syn.gaussian_cloud(pose, n) if cloud == "tracking" else syn.uniform_cloud(occ, res, origin, n)
    row = {}
    for mode, name in ((ob.TRIG_LIBM, "vs_libm_sinf_cosf"), (ob.TRIG_CR, "vs_correctly_rounded")):
        ref = ob.make_particles(x, y, t, np.ones(n))
        om.update_importance(scan, ref, mode)
        host = q.make_particles(x, y, t, np.ones(n))
        gm.update_importance(scan, host)
        err = np.abs(host["weight"] - ref["weight"]) / np.maximum(np.abs(ref["weight"]), 1e-300)
        row[name] = {"particles_beyond_1e-5": float((err > 1e-5).mean()),
                     "max_rel_err": float(err.max()),
                     "total_weight_rel_diff": float(abs(host["weight"].sum() - ref["weight"].sum()) / ref["weight"].sum())}
    out[f"{width}x{width}_{cloud}_{n}x{len(scan)}"] = row
print(json.dumps(out, indent=1))
