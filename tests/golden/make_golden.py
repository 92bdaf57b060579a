"""Generates tests/golden/*.npz from the CPU oracle (run here, in the build
container; the fixtures travel to the GPU box, the oracle build does too).

    python tests/golden/make_golden.py

Each fixture holds the inputs and the oracle's outputs of one small scenario of
the hot path, so that (a) the CUDA path can be checked against committed
numbers and (b) a change of the oracle itself shows up as a diff of this
directory.  Scenario = BASELINE.json configs[0] in small: 400 x 400 map at
5 cm, 360-beam scan, node defaults."""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402
from quickmcl_b200 import synthetic as syn  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
RES_THETA = float(np.float32(math.pi / 18))


def cycle(rtype, n, nmin, nmax, lowpass, name):
    occ, res, origin = syn.make_map(400)
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=0)
    om = ob.OracleMap(**lp)
    om.new_map(occ, res, origin, field_mode=ob.FIELD_EXACT_EDT)
    of = ob.OracleFilter(om, seed=11)
    of.set_parameters(resample_type=rtype, particle_count_min=nmin, particle_count_max=nmax,
                      resample_count=2, alpha_fast=0.1, alpha_slow=0.001, kld_epsilon=0.05,
                      kld_z=0.95, res_xy=0.5, res_theta=RES_THETA)
    of.set_trig_mode(ob.TRIG_CR)
    pose = syn.pick_truth_pose(occ, res, origin)
    scan = syn.make_scan(occ, res, origin, pose, 360, 360.0)
    x, y, t = syn.gaussian_cloud(pose, n, seed=5)
    start = ob.make_particles(x, y, t, np.full(n, 1.0 / n))
    of.set_particles(start)
    of.set_rng(11, 3)
    odom_new, odom_old = (0.25, 0.05, 0.1), (0.0, 0.0, 0.0)
    of.handle_odometry(odom_new, odom_old, syn.NODE_ALPHA)
    moved = of.get_particles()
    of.update_importance_from_observations(scan)
    weighted = of.get_particles()
    total = of.last_total_weight()
    of.set_lowpass(*lowpass)
    of.resample_and_cluster()
    resampled = of.get_particles()
    indices = of.resample_indices()
    keys, labels = of.bins()
    ok, est, cov = of.get_estimated_pose()
    np.savez_compressed(
        os.path.join(HERE, name), map_width=400, scan=scan, start=start, odom_new=odom_new,
        odom_old=odom_old, alpha=np.array(syn.NODE_ALPHA, np.float32), rng=np.array([11, 3]),
        params=np.array([rtype, nmin, nmax]), lowpass=np.array(lowpass), moved=moved,
        weighted=weighted, total=total, resampled=resampled, indices=indices, bin_keys=keys,
        bin_labels=labels, estimate_ok=ok, estimate=est, covariance=cov,
        field_crc=np.array([np.frombuffer(om.field().tobytes(), np.uint32).sum(dtype=np.uint64)]))
    print(name, len(resampled), "particles out,", len(keys), "bins, estimate", est)


if __name__ == "__main__":
    cycle(0, 2000, 2000, 2000, (0.0, 0.0), "c1_low_variance.npz")
    cycle(2, 3000, 100, 3000, (1.0, 0.9), "c1_kld_inject.npz")
    cycle(1, 2500, 1500, 1500, (0.0, 0.0), "c1_adaptive.npz")
