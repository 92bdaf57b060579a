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


def _oracle_cycle(om, i, n_particles, pose0, factor, odom_new, odom_old, scan, rtype, seed0, nmax=None):
    of = ob.OracleFilter(om, seed=seed0 + i)
    of.set_parameters(resample_type=rtype, particle_count_min=n_particles if rtype != 2 else 100,
                      particle_count_max=nmax or n_particles, kld_epsilon=0.05, kld_z=0.95)
    of.set_trig_mode(ob.TRIG_CR)
    of.set_rng(seed0 + i, 0)
    of.initialise_factor(pose0.astype(np.float32), factor)
    of.handle_odometry(odom_new, odom_old, syn.NODE_ALPHA)
    of.update_importance_from_observations(scan)
    of.resample_and_cluster()
    return of.get_estimated_pose(), of.get_particles()


@pytest.mark.parametrize("n_filters,rtype", [(256, 0), (32, 2), (16, 1)])
def test_fused_batch_matches_independent_oracles(n_filters, rtype):
    """configs[4]: 256 robots x 2000 particles, one qmcl_batch (four launches per cycle for
    the whole batch) against 256 independent oracle filters: estimates within the bar, the
    resampled sets slot for slot; also KLD and adaptive batches."""
    n_particles = 2000
    occ, res, origin = syn.make_map(400)
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=30)
    pf = q.ParticleFilterParameters(resample_type=q.ResampleType(rtype),
                                    particle_count_min=n_particles if rtype != 2 else 100,
                                    particle_count_max=n_particles, kld_epsilon=0.05, kld_z=0.95)
    b = batch.FusedBatch(n_filters, q.OccupancyGrid(occ, res, origin[0], origin[1]),
                         q.LikelihoodMapParameters(**lp), pf, seed=100)
    truth = syn.pick_truth_pose(occ, res, origin)
    poses = np.array([truth + np.array([0.05 * (i % 5), -0.04 * (i % 3), 0.02 * (i % 7)])
                      for i in range(n_filters)])
    scans = [syn.make_scan(occ, res, origin, poses[i], 360, 360.0, seed=10 + i)
             for i in range(n_filters)]
    factor = np.diag(np.sqrt([0.25, 0.25, 0.01])).astype(np.float32)
    for i, f in enumerate(b.filters):
        f.set_rng(100 + i, 0)
        f.initialise_factor(poses[i].astype(np.float32), factor)
    odom_new = np.tile([0.02, 0.01, 0.01], (n_filters, 1))
    odom_old = np.zeros((n_filters, 3))
    launches0 = q.kernel_launch_count()
    got = b.cycle(odom_new, odom_old, syn.NODE_ALPHA, scans)
    launches = q.kernel_launch_count() - launches0
    assert b.stats() == (1, 0), b.stats()
    assert launches <= 5, launches           # front, sensor, tail (+ the map's table build, once)
    om = ob.OracleMap(**lp)
    om.new_map(occ, res, origin, field_mode=ob.FIELD_EXACT_EDT)
    check = range(n_filters) if n_filters <= 32 else list(range(0, n_filters, 9)) + [n_filters - 1]
    for i in check:
        (ok, pose, cov), op = _oracle_cycle(om, i, n_particles, poses[i], factor, odom_new[i],
                                            odom_old[i], scans[i], rtype, 100,
                                            nmax=n_particles)
        g_ok, g_pose, g_cov = got[i]
        assert g_ok and ok == 1
        assert np.max(np.abs(g_pose[:2] - pose[:2])) <= 1e-4, i
        assert abs(math.remainder(g_pose[2] - pose[2], 2 * math.pi)) <= 1e-4, i
        gp = b.filters[i].get_particles()
        assert len(gp) == len(op), (i, len(gp), len(op))
        for k in ("x", "y", "theta"):
            assert np.max(np.abs(gp[k].astype(np.float64) - op[k])) <= 1e-5, (i, k)
    # a second cycle without resampling (cluster + estimate), still fused
    launches0 = q.kernel_launch_count()
    got2 = b.cycle(odom_new * 2, odom_new, syn.NODE_ALPHA, scans, resample=False)
    assert q.kernel_launch_count() - launches0 == 3   # front, sensor, tail for the whole batch
    assert b.stats() == (2, 0)
    assert all(g[0] for g in got2)
    b.close()


def test_fused_batch_equals_single_filters_bit_for_bit():
    """A batched cycle and the same cycle through the per-filter calls leave identical
    particle sets and estimates (the batch only changes how the kernels are launched)."""
    n_filters, n_particles = 12, 1500
    occ, res, origin = syn.make_map(400)
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=0)
    pf = q.ParticleFilterParameters(resample_type=q.ResampleType.kld, particle_count_min=100,
                                    particle_count_max=n_particles, kld_epsilon=0.05, kld_z=0.95)
    grid = q.OccupancyGrid(occ, res, origin[0], origin[1])
    b = batch.FusedBatch(n_filters, grid, q.LikelihoodMapParameters(**lp), pf, seed=7)
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(**lp))
    gm.new_map(grid)
    truth = syn.pick_truth_pose(occ, res, origin)
    factor = np.diag(np.sqrt([0.25, 0.25, 0.01])).astype(np.float32)
    scans = [syn.make_scan(occ, res, origin, truth, 360, 360.0, seed=20 + i) for i in range(n_filters)]
    singles = []
    for i in range(n_filters):
        f = q.create_particle_filter(gm, seed=7 + i)
        f.set_parameters(pf)
        for g in (f, b.filters[i]):
            g.set_rng(7 + i, 0)
            g.initialise_factor(truth.astype(np.float32), factor)
        singles.append(f)
    odom_new = np.tile([0.03, -0.01, 0.02], (n_filters, 1))
    odom_old = np.zeros((n_filters, 3))
    for cyc in range(3):
        got = b.cycle(odom_new * (cyc + 1), odom_new * cyc, syn.NODE_ALPHA, scans)
        for i, f in enumerate(singles):
            f.handle_odometry(odom_new[i] * (cyc + 1), odom_new[i] * cyc, syn.NODE_ALPHA)
            f.update_importance_from_observations(scans[i])
            f.resample_and_cluster()
            ok, pose, cov = f.get_estimated_pose()
            assert ok == got[i][0]
            assert pose.tobytes() == got[i][1].tobytes() and cov.tobytes() == got[i][2].tobytes(), (cyc, i)
            assert f.get_particles().tobytes() == b.filters[i].get_particles().tobytes(), (cyc, i)
    b.close()
