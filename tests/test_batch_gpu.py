"""BASELINE.json configs[4]: the field rebuild on a large map and the multi-robot
filter batch.

The oracle cannot brute-force an 8192 x 8192 map in test time, so the rebuild is
checked through a size-independent property: the clipped distance transform is
local (radius max_obstacle_distance), hence the field of any window, away from
the window's rim by that radius, must equal the field of the same window cut
out and rebuilt alone — by the GPU and by the oracle."""
import math

import numpy as np
import pytest

import quickmcl_b200 as q
from quickmcl_b200 import batch, synthetic as syn
from oracle import binding as ob

pytestmark = pytest.mark.gpu


def test_field_rebuild_8192_is_local_and_matches_oracle_windows():
    W = 8192
    occ, res, origin = syn.make_map(W)
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=0)
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(**lp))
    gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
    field, state, dist = gm.field(), gm.state(), gm.distance()
    assert field.shape == (W, W)
    R = int(math.ceil(lp["max_obstacle_distance"] / res))  # 40 cells
    rng = np.random.default_rng(4)
    for k in range(6):
        n = 500
        x0 = int(rng.integers(0, W - n)) if k else 0          # one window on the map's corner
        y0 = int(rng.integers(0, W - n)) if k else 0
        crop = np.ascontiguousarray(occ[y0:y0 + n, x0:x0 + n])
        om = ob.OracleMap(**lp)
        om.new_map(crop, res, (0.0, 0.0), field_mode=ob.FIELD_EXACT_EDT)
        # interior of the window: every obstacle within reach lies inside the crop
        ia = slice(R if y0 else 0, n - R)
        ib = slice(R if x0 else 0, n - R)
        assert np.array_equal(field[y0:y0 + n, x0:x0 + n][ia, ib], om.field()[ia, ib])
        assert np.array_equal(dist[y0:y0 + n, x0:x0 + n][ia, ib], om.distance()[ia, ib])
        assert np.array_equal(state[y0:y0 + n, x0:x0 + n], om.state())
    # free-cell list: row-major, exactly the free cells
    fc = gm.free_cells()
    assert len(fc) == int((state == 0).sum())
    assert np.all(np.diff(fc[:, 1].astype(np.int64) * W + fc[:, 0]) > 0)
    assert np.all(state[fc[::997, 1], fc[::997, 0]] == 0)


def test_filter_batch_matches_independent_oracles():
    n_filters, n_particles = 24, 2000
    occ, res, origin = syn.make_map(400)
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=30)
    pf = q.ParticleFilterParameters(resample_type=q.ResampleType.low_variance,
                                    particle_count_min=n_particles,
                                    particle_count_max=n_particles)
    b = batch.FilterBatch(n_filters, q.OccupancyGrid(occ, res, origin[0], origin[1]),
                          q.LikelihoodMapParameters(**lp), pf, lanes=6, seed=100)
    truth = syn.pick_truth_pose(occ, res, origin)
    poses = np.array([truth + np.array([0.05 * (i % 5), -0.04 * (i % 3), 0.02 * (i % 7)])
                      for i in range(n_filters)])
    scans = [syn.make_scan(occ, res, origin, poses[i], 360, 360.0, seed=10 + i)
             for i in range(n_filters)]
    factor = np.diag(np.sqrt([0.25, 0.25, 0.01])).astype(np.float32)
    b.for_each(lambda i, f: (f.set_rng(100 + i, 0),
                             f.initialise_factor(poses[i].astype(np.float32), factor)))
    odom_new = np.tile([0.02, 0.01, 0.01], (n_filters, 1))
    odom_old = np.zeros((n_filters, 3))
    got = b.update_native(odom_new, odom_old, syn.NODE_ALPHA, scans)
    om = ob.OracleMap(**lp)
    om.new_map(occ, res, origin, field_mode=ob.FIELD_EXACT_EDT)
    for i in range(n_filters):
        of = ob.OracleFilter(om, seed=100 + i)
        of.set_parameters(resample_type=0, particle_count_min=n_particles,
                          particle_count_max=n_particles)
        of.set_trig_mode(ob.TRIG_CR)
        of.set_rng(100 + i, 0)
        of.initialise_factor(poses[i].astype(np.float32), factor)
        of.handle_odometry(odom_new[i], odom_old[i], syn.NODE_ALPHA)
        of.update_importance_from_observations(scans[i])
        of.resample_and_cluster()
        ok, pose, cov = of.get_estimated_pose()
        g_ok, g_pose, g_cov = got[i]
        assert g_ok and ok == 1
        assert np.max(np.abs(g_pose[:2] - pose[:2])) <= 1e-4, i
        assert abs(math.remainder(g_pose[2] - pose[2], 2 * math.pi)) <= 1e-4, i
    b.close()
