"""LaserScan -> base-frame cloud on the device
(laser_handler.cpp:108-121) against the oracle's restatement."""
import math

import numpy as np
import pytest

from oracle import binding as ob
from quickmcl_b200 import synthetic as syn


def hokuyo_scan(seed=3, n=1081):
    rng = np.random.default_rng(seed)
    angle_min = float(np.float32(-2.3561945))
    inc = float(np.float32(math.radians(0.25)))
    r = rng.uniform(0.3, 12.0, n).astype(np.float32)
    r[rng.random(n) < 0.1] = np.inf            # no return
    r[rng.random(n) < 0.05] = np.nan
    r[rng.random(n) < 0.05] = 0.01              # below range_min
    r[7] = 30.0                                 # == range_max: dropped (r < range_max)
    r[8] = np.float32(0.1)                      # == range_min: kept
    return r, angle_min, inc, float(np.float32(0.1)), 30.0


def test_oracle_projection_basics():
    r = np.array([1.0, 2.0, np.inf, 0.05, 3.0], np.float32)
    pts = ob.project_scan(r, 0.0, math.pi / 2, 0.1, 10.0)
    assert len(pts) == 3
    assert np.allclose(pts[0], [1.0, 0.0]) and np.allclose(pts[1], [0.0, 2.0], atol=1e-6)
    assert np.allclose(pts[2], [3.0, 0.0], atol=1e-5)          # 4 * pi/2: full turn
    moved = ob.project_scan(r, 0.0, math.pi / 2, 0.1, 10.0, mount=(0.5, -0.25, math.pi / 2))
    assert np.allclose(moved[0], [0.5, 0.75], atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("mount", [(0.0, 0.0, 0.0), (0.21, -0.03, 0.5), (-0.1, 0.0, math.pi)])
def test_device_projection_and_update_match_oracle(mount):
    import quickmcl_b200 as q
    ranges, amin, inc, rmin, rmax = hokuyo_scan()
    want = ob.project_scan(ranges, amin, inc, rmin, rmax, mount)
    occ, res, origin = syn.make_map(400)
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=30)
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(**lp))
    gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
    pose = syn.pick_truth_pose(occ, res, origin)
    n = 3000
    x, y, t = syn.gaussian_cloud(pose, n)
    start = q.make_particles(x, y, t, np.full(n, 1.0 / n))
    f = q.create_particle_filter(gm, seed=1)
    f.set_particles(start)
    k = f.update_importance_from_scan(ranges, amin, inc, rmin, rmax, mount)
    got = f.last_cloud()
    assert k == len(want) == len(got)
    # double sincos on the device vs glibc: identical up to the float rounding of the result
    assert np.max(np.abs(got - want)) <= 2e-6
    assert np.mean(got.view(np.uint32) == want.view(np.uint32)) > 0.99
    # the update itself equals the host-cloud update on the same points
    w_scan = f.get_particles()["weight"].copy()
    f.set_particles(start)
    f.update_importance_from_observations(got)
    assert np.array_equal(w_scan, f.get_particles()["weight"])
