"""Generates tests/golden/ref_*.npz from the REFERENCE ITSELF: the reference's own
map_likelihood.cpp, odometry.cpp and particle_filter.cpp, compiled from /root/reference into
oracle/_ref/libqmcl_ref.so (oracle/Makefile, target `ref`), run in this container.

    python tests/golden/make_golden_ref.py

Each fixture holds the inputs of one scan cycle and what the reference computed from them:
its likelihood table and state map, the particles after handle_odometry, after
update_importance_from_observations, after resample_and_cluster, the low-pass state, the
cluster count and the pose estimate.  The reference's random engines are private and
unseeded; they read the draws of the shared Philox contract from a
script here (oracle/ref_stubs/random), and the fixture records (seed, step) of every draw so
that the oracle and the CUDA path can make the same ones.  The fixtures travel to the GPU
box; /root/reference and the library do not have to.
"""
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402  (particle layout and the Philox words only)
from oracle import ref_binding as rb  # noqa: E402
from quickmcl_b200 import synthetic as syn  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
RES_THETA = float(np.float32(math.pi / 18))
MAP_WIDTH, MAP_SEED = 200, 1
SEED, STEP_MOTION, STEP_RESAMPLE = 17, 4, 5
ODOM_NEW, ODOM_OLD = (0.25, 0.05, 0.1), (0.0, 0.0, 0.0)


def reference_cycle(rtype, n, nmin, nmax, lowpass):
    """One cycle of the reference's ParticleFilter; returns everything a fixture holds."""
    occ, res, origin = syn.make_map(MAP_WIDTH, seed=MAP_SEED)
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=0)
    rm = rb.RefMap(occ, res, origin, **lp)
    rf = rb.RefFilter(rm)
    rf.set_parameters(resample_type=rtype, particle_count_min=nmin, particle_count_max=nmax, resample_count=2,
                      alpha_fast=0.1, alpha_slow=0.001, kld_epsilon=0.05, kld_z=0.95, res_xy=0.5,
                      res_theta=RES_THETA)
    pose = syn.pick_truth_pose(occ, res, origin)
    scan = syn.make_scan(occ, res, origin, pose, 360, 360.0)
    x, y, t = syn.gaussian_cloud(pose, n, seed=5)
    start = ob.make_particles(x, y, t, np.full(n, 1.0 / n))
    rf.set_particles(start)
    rb.script(normals=rb.motion_script(SEED, STEP_MOTION, n))
    rf.handle_odometry(ODOM_NEW, ODOM_OLD, syn.NODE_ALPHA)
    moved = rf.get_particles()
    total = rm.update_importance(scan, moved.copy())
    rf.update_importance_from_observations(scan)
    weighted = rf.get_particles()
    lowpass_sensor = rf.lowpass()
    rf.set_lowpass(*lowpass)
    w_diff = max(0.0, 1.0 - lowpass[1] / lowpass[0]) if lowpass[0] else 0.0
    if rtype == 0:
        rb.script(words=rb.low_variance_script(SEED, STEP_RESAMPLE))
        injected = np.zeros(nmin, bool)
    else:
        words, injected = rb.draw_script(SEED, STEP_RESAMPLE, nmax if rtype == 2 else nmin, w_diff)
        rb.script(words=words)
    assert rf.resample_and_cluster(), "the reference's low-variance walk ran off the end: pick another step"
    assert not rb.script_state()[2]
    resampled = rf.get_particles()
    ok, est, cov = rf.get_estimated_pose()
    return dict(
        map_width=MAP_WIDTH, map_seed=MAP_SEED, field=rm.field(), state=rm.state(), scan=scan, start=start,
        odom_new=np.array(ODOM_NEW), odom_old=np.array(ODOM_OLD), alpha=np.array(syn.NODE_ALPHA, np.float32),
        rng=np.array([SEED, STEP_MOTION, STEP_RESAMPLE]), params=np.array([rtype, nmin, nmax]),
        lowpass=np.array(lowpass, np.float64), moved=moved, total=total, weighted=weighted,
        lowpass_sensor=lowpass_sensor, resampled=resampled, injected=injected[:len(resampled)],
        lowpass_after=rf.lowpass(), cluster_count=rf.num_clusters(), estimate_ok=int(ok), estimate=est,
        covariance=cov)


SCENARIOS = {
    "ref_low_variance.npz": (0, 1500, 1500, 1500, (0.0, 0.0)),
    "ref_kld_inject.npz": (2, 3000, 100, 3000, (1.0, 0.9)),
    "ref_kld_tracking.npz": (2, 3000, 100, 3000, (0.0, 0.0)),
    "ref_adaptive.npz": (1, 2500, 1500, 1500, (1.0, 0.8)),
}


if __name__ == "__main__":
    if not rb.available():
        sys.exit("oracle/_ref is not built and there is no reference tree to build it from")
    for name, args in SCENARIOS.items():
        fx = reference_cycle(*args)
        np.savez_compressed(os.path.join(HERE, name), **fx)
        size = os.path.getsize(os.path.join(HERE, name))
        print(f"{name}: {len(fx['resampled'])} particles out ({int(fx['injected'].sum())} injected), "
              f"{fx['cluster_count']} clusters, estimate {fx['estimate']}, {size / 1024:.0f} KiB")
