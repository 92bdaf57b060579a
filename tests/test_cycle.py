"""The ROS-free scan cycle (quickmcl_b200/cycle.py) against the call sequence of
laser_handler.cpp:179-288, on the CPU oracle (here) and on the CUDA filter (GPU)."""
import math

import numpy as np
import pytest

from oracle import binding as ob
from quickmcl_b200 import synthetic as syn
from quickmcl_b200.cycle import MotionModelParameters, ScanCycle

RES_THETA = float(np.float32(math.pi / 18))


class OracleAsFilter:
    """The oracle behind the IParticleFilter method names the cycle uses."""

    def __init__(self, of):
        self.of = of

    def __getattr__(self, name):
        return getattr(self.of, name)

    def get_estimated_pose(self):
        ok, pose, cov = self.of.get_estimated_pose()
        return ok == 1, pose, cov


def scripted_log():
    """(odometry, expected calls) for a drive with pauses: the reference's cadence."""
    M, S, R, C = "handle_odometry", "update_importance_from_observations", \
        "resample_and_cluster", "cluster"
    E = "get_estimated_pose"
    return [
        ((0.00, 0.0, 0.0), [C, E]),                 # first scan: no motion yet -> cluster only
        ((0.10, 0.0, 0.0), [C, E]),                 # below min_trans
        ((0.25, 0.0, 0.0), [M, S, C, E]),           # counter 0 % 2 == 0 -> NOT resampled
        ((0.30, 0.0, 0.0), [C, E]),                 # 5 cm since the last filter run
        ((0.50, 0.0, 0.0), [M, S, R, E]),           # counter 1 -> resampled
        ((0.50, 0.0, 0.6), [M, S, C, E]),           # rotation alone (> pi/6) triggers; counter 2
        ((0.75, 0.3, 0.6), [M, S, R, E]),           # counter 3
    ]


def run_script(cycle, scan):
    results = []
    for odom, want in scripted_log():
        before = len(cycle.calls)
        results.append(cycle.on_scan(odom, scan))
        assert cycle.calls[before:] == want, (odom, cycle.calls[before:])
    return results


def make_oracle(n=1500):
    occ, res, origin = syn.make_map(400)
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=30)
    om = ob.OracleMap(**lp)
    om.new_map(occ, res, origin, field_mode=ob.FIELD_EXACT_EDT)
    of = ob.OracleFilter(om, seed=21)
    of.set_parameters(resample_type=0, particle_count_min=n, particle_count_max=n,
                      res_theta=RES_THETA)
    of.set_trig_mode(ob.TRIG_CR)
    pose = syn.pick_truth_pose(occ, res, origin)
    scan = syn.make_scan(occ, res, origin, pose, 360, 360.0)
    factor = np.diag(np.sqrt([0.25, 0.25, 0.01])).astype(np.float32)
    of.set_rng(21, 0)
    of.initialise_factor(pose.astype(np.float32), factor)
    return occ, res, origin, lp, of, pose, scan, factor


def test_cadence_and_gate_on_the_oracle():
    occ, res, origin, lp, of, pose, scan, _ = make_oracle()
    cyc = ScanCycle(OracleAsFilter(of), MotionModelParameters(), resample_count=2)
    res_ = run_script(cyc, scan)
    assert [r.processed for r in res_] == [False, False, True, False, True, True, True]
    assert [r.resampled for r in res_] == [False, False, False, False, True, False, True]
    assert res_[0].recompute_transform and not res_[1].recompute_transform
    assert all(r.pose_ok for r in res_)
    # resample_count = 1 never resamples (x % 1 == 0), as in the reference
    cyc1 = ScanCycle(OracleAsFilter(of), resample_count=1)
    for k in range(4):
        r = cyc1.on_scan((1.0 + 0.3 * k, 0.0, 0.0), scan)
        assert r.processed == (k > 0) and not r.resampled


def test_empty_set_triggers_global_localisation_and_pose_reset_forces_a_run():
    occ, res, origin, lp, of, pose, scan, _ = make_oracle()
    of.set_particles(ob.particles_array(0))
    of.set_parameters(resample_type=0, particle_count_min=500, particle_count_max=800,
                      res_theta=RES_THETA)
    cyc = ScanCycle(OracleAsFilter(of))
    cyc.force_pose_reset()
    r = cyc.on_scan((0.0, 0.0, 0.0), scan)       # no motion, but the reset forces processing
    assert r.processed and r.global_localization and r.recompute_transform
    assert cyc.calls[:3] == ["global_localization", "handle_odometry",
                             "update_importance_from_observations"]
    assert of.num_particles() == 800


@pytest.mark.gpu
def test_cycle_on_cuda_filter_tracks_the_oracle():
    import quickmcl_b200 as q
    occ, res, origin, lp, of, pose, scan, factor = make_oracle()
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(**lp))
    gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
    gf = q.create_particle_filter(gm, seed=21)
    gf.set_parameters(q.ParticleFilterParameters(resample_type=q.ResampleType.low_variance,
                                                 particle_count_min=1500,
                                                 particle_count_max=1500,
                                                 space_partitioning_resolution_theta=RES_THETA))
    gf.set_rng(21, 0)
    gf.initialise_factor(pose.astype(np.float32), factor)
    g_res = run_script(ScanCycle(gf), scan)
    o_res = run_script(ScanCycle(OracleAsFilter(of)), scan)
    for g, o in zip(g_res, o_res):
        assert (g.processed, g.resampled, g.pose_ok) == (o.processed, o.resampled, o.pose_ok)
        assert np.max(np.abs(g.pose[:2] - o.pose[:2])) <= 1e-4
        assert abs(math.remainder(g.pose[2] - o.pose[2], 2 * math.pi)) <= 1e-4
