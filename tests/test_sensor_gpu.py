"""K7 parity: the CUDA sensor model against the CPU oracle on identical inputs
(identical field, particles and beams), through the C ABI.

Tolerance (BASELINE.json north_star): per-particle weights within 1e-5
relative.  With the index path bit-exact the observed difference is ~1e-15
(summation order only), so the tests assert 1e-12 and report the maximum.
"""
import math

import numpy as np
import pytest

import quickmcl_b200 as q
from quickmcl_b200 import synthetic as syn
from oracle import binding as ob

pytestmark = pytest.mark.gpu

REL = 1e-5      # the stated tolerance
TIGHT = 1e-12   # what bit-exact indexing actually delivers


def build_pair(width, num_beams=0, field_mode=ob.FIELD_EXACT_EDT, seed=1):
    occ, res, origin = syn.make_map(width, seed=seed)
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=num_beams)
    om = ob.OracleMap(**lp)
    om.new_map(occ, res, origin, field_mode=field_mode)
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(**lp))
    # identical field on both sides (SURVEY §8c parity protocol)
    gm.load_field(om.field(), om.state(), res, origin)
    return occ, res, origin, om, gm


def rel_err(a, b):
    denom = np.maximum(np.abs(b), 1e-300)
    return np.where(b == 0, np.abs(a), np.abs(a - b) / denom)


@pytest.mark.parametrize("width,n,beams,fov,num_beams", [
    (400, 2000, 360, 360.0, 30),     # C1, reference default num_beams
    (400, 2000, 360, 360.0, 0),      # C1, all beams
    (400, 777, 360, 360.0, 7),       # ragged: step 51 -> 8 beams
    (2000, 20000, 1081, 270.0, 0),   # C2 shape at a size the oracle finishes in seconds
])
def test_weights_match_oracle_tracking(width, n, beams, fov, num_beams):
    occ, res, origin, om, gm = build_pair(width, num_beams)
    pose = syn.pick_truth_pose(occ, res, origin)
    scan = syn.make_scan(occ, res, origin, pose, beams, fov)
    x, y, t = syn.gaussian_cloud(pose, n)
    ref = ob.make_particles(x, y, t, np.full(n, 1.0 / n))
    total_ref = om.update_importance(scan, ref, ob.TRIG_CR)
    f = q.create_particle_filter(gm)
    f.set_particles(q.make_particles(x, y, t, np.full(n, 1.0 / n)))
    f.update_importance_from_observations(scan)
    got = f.get_particles()
    total = f.last_total_weight()
    assert abs(total - total_ref) <= TIGHT * total_ref
    # update_importance_from_observations normalises by the pre-penalty total
    want = ref["weight"] / total_ref
    err = rel_err(got["weight"], want)
    assert err.max() <= TIGHT, err.max()
    assert err.max() <= REL
    assert np.array_equal(got["x"], x) and np.array_equal(got["theta"], t)


def test_weights_match_oracle_global_with_penalty():
    """C3 shape: particles all over the map, incl. outside it, in walls and in unknown space."""
    occ, res, origin, om, gm = build_pair(400, 0)
    pose = syn.pick_truth_pose(occ, res, origin)
    scan = syn.make_scan(occ, res, origin, pose, 1081, 270.0)
    n = 30000
    x, y, t = syn.uniform_cloud(occ, res, origin, n)
    ref = ob.make_particles(x, y, t, np.ones(n))
    total_ref = om.update_importance(scan, ref, ob.TRIG_CR)
    host = q.make_particles(x, y, t, np.ones(n))
    total = gm.update_importance(scan, host)   # the Map::update_importance host overload
    assert abs(total - total_ref) <= TIGHT * total_ref
    err = rel_err(host["weight"], ref["weight"])
    assert err.max() <= TIGHT, err.max()
    st = np.array([om.state_at(float(a), float(b)) for a, b in zip(x[:2000], y[:2000])])
    assert (st == 2).any() and (st == 1).any() and (st == 0).any()
    assert np.all(host["weight"][:2000][st == 2] == 0)


def test_brushfire_field_gives_identical_weights():
    """Weights on the reference's own (brushfire) field: the field is an input here."""
    occ, res, origin, om, gm = build_pair(400, 0, field_mode=ob.FIELD_BRUSHFIRE)
    pose = syn.pick_truth_pose(occ, res, origin)
    scan = syn.make_scan(occ, res, origin, pose, 360, 360.0)
    x, y, t = syn.gaussian_cloud(pose, 3000)
    ref = ob.make_particles(x, y, t, np.ones(3000))
    om.update_importance(scan, ref, ob.TRIG_CR)
    host = q.make_particles(x, y, t, np.ones(3000))
    gm.update_importance(scan, host)
    assert rel_err(host["weight"], ref["weight"]).max() <= TIGHT


def test_libm_trig_mode_reports_flip_fraction():
    """Against the host libm's sinf/cosf (what the reference calls) the only
    differences are cell flips caused by libm's 1-ulp errors; they must be rare."""
    occ, res, origin, om, gm = build_pair(400, 0)
    pose = syn.pick_truth_pose(occ, res, origin)
    scan = syn.make_scan(occ, res, origin, pose, 1081, 270.0)
    n = 20000
    x, y, t = syn.gaussian_cloud(pose, n)
    ref = ob.make_particles(x, y, t, np.ones(n))
    om.update_importance(scan, ref, ob.TRIG_LIBM)
    host = q.make_particles(x, y, t, np.ones(n))
    gm.update_importance(scan, host)
    err = rel_err(host["weight"], ref["weight"])
    frac = float((err > REL).mean())
    print(f"libm-trig: {frac:.5f} of particles beyond 1e-5, max rel {err.max():.3e}")
    assert frac < 0.01
    assert abs(host["weight"].sum() - ref["weight"].sum()) <= 1e-6 * ref["weight"].sum()


def test_ragged_particle_counts_and_small_map_edges():
    """Particle counts that do not fill the last warp group, on a map small enough
    that beams leave it: the interior fast path (no in-map test) must not be taken
    for padding particles or for particles within reach of an edge."""
    occ, res, origin, om, gm = build_pair(400)
    # the palette path (codes) is the default one when the map is built here
    gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
    pose = syn.pick_truth_pose(occ, res, origin)
    scan = syn.make_scan(occ, res, origin, pose, 360, 360.0)
    for n in (1, 3, 5, 33, 3001, 2999):
        x, y, t = syn.uniform_cloud(occ, res, origin, n, seed=n)
        ref = ob.make_particles(x, y, t, np.ones(n))
        total_ref = om.update_importance(scan, ref, ob.TRIG_CR)
        got = q.make_particles(x, y, t, np.ones(n))
        total = gm.update_importance(scan, got)
        assert abs(total - total_ref) <= 1e-6 * total_ref
        nz = ref["weight"] > 0
        assert rel_err(got["weight"][nz], ref["weight"][nz]).max() <= 1e-5


def test_zero_total_equalises():
    """particle_filter.cpp:113-120: zero total weight -> every weight 1/N."""
    occ, res, origin = syn.make_map(400)
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(z_hit=1.0, z_rand=0.0, num_beams=0,
                                                sigma_hit=0.1))
    field = np.zeros(occ.shape, np.float32)
    gm.load_field(field, np.zeros(occ.shape, np.int8), res, origin)
    f = q.create_particle_filter(gm)
    n = 100
    f.set_particles(q.make_particles(np.zeros(n), np.zeros(n), np.zeros(n), np.ones(n)))
    f.update_importance_from_observations(np.array([[1.0, 0.0], [0.0, 1.0]], np.float32))
    got = f.get_particles()
    assert f.last_total_weight() == 0.0
    assert np.all(got["weight"] == 1.0 / n)


def test_empty_inputs():
    occ, res, origin, om, gm = build_pair(400, 0)
    f = q.create_particle_filter(gm)
    f.set_particles(q.particles_array(0))
    f.update_importance_from_observations(np.zeros((5, 2), np.float32))
    assert f.num_particles() == 0
    # empty scan: every particle gets weight 0 -> total 0 -> equalise
    f.set_particles(q.make_particles(np.zeros(3), np.zeros(3), np.zeros(3), np.ones(3)))
    f.update_importance_from_observations(np.zeros((0, 2), np.float32))
    assert np.all(f.get_particles()["weight"] == 1.0 / 3)


@pytest.mark.parametrize("width,n,beams,fov,num_beams,cloud", [
    (400, 2000, 360, 360.0, 30, "tracking"),
    (400, 5000, 360, 360.0, 0, "global"),
    (2000, 20000, 1081, 270.0, 0, "tracking"),
])
def test_palette_path_matches_oracle_and_field_path(width, n, beams, fov, num_beams, cloud):
    """Map built on the GPU (exact EDT): the uint16-code + pw^3-table path must give
    the oracle's weights, and exactly the float-field path's bits."""
    occ, res, origin = syn.make_map(width)
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=num_beams)
    om = ob.OracleMap(**lp)
    om.new_map(occ, res, origin, field_mode=ob.FIELD_EXACT_EDT)
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(**lp))
    gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
    pose = syn.pick_truth_pose(occ, res, origin)
    scan = syn.make_scan(occ, res, origin, pose, beams, fov)
    if cloud == "tracking":
        x, y, t = syn.gaussian_cloud(pose, n)
    else:
        x, y, t = syn.uniform_cloud(occ, res, origin, n)
    ref = ob.make_particles(x, y, t, np.ones(n))
    total_ref = om.update_importance(scan, ref, ob.TRIG_CR)
    gm.set_sensor_path(2)                      # exact palette (double table)
    pal = q.make_particles(x, y, t, np.ones(n))
    total_pal = gm.update_importance(scan, pal)
    gm.set_sensor_path(1)                      # float field gather
    fld = q.make_particles(x, y, t, np.ones(n))
    total_fld = gm.update_importance(scan, fld)
    assert np.array_equal(pal["weight"], fld["weight"])
    assert total_pal == total_fld
    assert rel_err(pal["weight"], ref["weight"]).max() <= TIGHT
    assert abs(total_pal - total_ref) <= TIGHT * total_ref
    gm.set_sensor_path(0)                      # default: float table, <= 8-term float sums
    fast = q.make_particles(x, y, t, np.ones(n))
    total_fast = gm.update_importance(scan, fast)
    err = rel_err(fast["weight"], ref["weight"])
    print(f"PAL32 max rel err {err.max():.3e} (bar 1e-5)")
    assert err.max() <= 1e-6
    assert abs(total_fast - total_ref) <= 1e-6 * total_ref
    # parameters that only enter the table: changing them must refresh it
    gm.set_sensor_path(2)
    lp2 = dict(lp, z_hit=0.6, z_rand=0.4, max_laser_distance=20.0)
    gm.set_parameters(q.LikelihoodMapParameters(**lp2))
    om.set_parameters(**lp2)
    ref2 = ob.make_particles(x, y, t, np.ones(n))
    om.update_importance(scan, ref2, ob.TRIG_CR)
    pal2 = q.make_particles(x, y, t, np.ones(n))
    gm.update_importance(scan, pal2)
    assert rel_err(pal2["weight"], ref2["weight"]).max() <= TIGHT


@pytest.mark.parametrize("res,origin_shift", [(0.05, 0.0), (0.1, 0.013), (0.025, -3.7), (0.03, 1e-3),
                                              (0.0625, 0.0)])
def test_division_forms_pick_identical_cells(res, origin_shift):
    """The cell index (m - o) / res (map_base.cpp:66-77) may be computed with IEEE
    division, the three-op Markstein form or the two-op split reciprocal; whatever
    the load-time self-check verified must give the very same weights as IEEE
    division, on the interior fast path (tracking cloud) and on the bounds-checked
    path (global cloud reaching over the map edge)."""
    occ, _, origin = syn.make_map(600)
    origin = (origin[0] * res / 0.05 + origin_shift, origin[1] * res / 0.05 - origin_shift)
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=0)
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(**lp))
    gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
    verified = gm.division_mode()
    print(f"res {res}: verified division form {verified}")
    assert verified in (0, 1, 2)
    pose = syn.pick_truth_pose(occ, res, origin)
    scan = syn.make_scan(occ, res, origin, pose, 725, 270.0)
    n = 6000
    for cloud in (syn.gaussian_cloud(pose, n), syn.uniform_cloud(occ, res, origin, n)):
        x, y, t = cloud
        got = {}
        for mode in (0, 1, 2, -1):
            gm.set_division_mode(mode)
            for path in (0, 2):
                gm.set_sensor_path(path)
                ps = q.make_particles(x, y, t, np.ones(n))
                total = gm.update_importance(scan, ps)
                got[(mode, path)] = (ps["weight"].copy(), total)
        for path in (0, 2):
            w0, t0 = got[(0, path)]
            assert w0.max() > 0
            for mode in (1, 2, -1):
                w, tt = got[(mode, path)]
                assert np.array_equal(w.view(np.uint64), w0.view(np.uint64)), (mode, path)
                assert tt == t0
    gm.set_division_mode(-1)
    gm.set_sensor_path(0)


@pytest.mark.parametrize("sigma_hit,res", [(0.1, 0.05), (0.2, 0.05), (0.3, 0.05), (0.1, 0.1)])
def test_one_byte_palette_identical_to_float_table(sigma_hit, res):
    """Path 3 (one palette index per cell) and path 4 (uint16 squared-distance code per
    cell) look the same float pw^3 values up, so the weights must agree bit for bit —
    on the interior fast path, on the bounds-checked path reaching over the map edge,
    after a parameter change that rebuilds the table, and when the distinct values do
    not fit a byte (sigma_hit = 0.3: path 3 then runs path 4)."""
    occ, _, origin = syn.make_map(600)
    origin = (origin[0] * res / 0.05, origin[1] * res / 0.05)
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=0, sigma_hit=sigma_hit)
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(**lp))
    gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
    pose = syn.pick_truth_pose(occ, res, origin)
    scan = syn.make_scan(occ, res, origin, pose, 725, 270.0)
    n = 5003   # ragged: not a multiple of the particles per block
    for params in (lp, dict(lp, z_hit=0.6, z_rand=0.4, max_laser_distance=20.0)):
        gm.set_parameters(q.LikelihoodMapParameters(**params))
        for x, y, t in (syn.gaussian_cloud(pose, n), syn.uniform_cloud(occ, res, origin, n)):
            got = {}
            for path in (4, 3, 0):
                gm.set_sensor_path(path)
                ps = q.make_particles(x, y, t, np.ones(n))
                total = gm.update_importance(scan, ps)
                got[path] = (ps["weight"].copy(), total)
            assert got[4][0].max() > 0
            for path in (3, 0):
                assert np.array_equal(got[path][0].view(np.uint64), got[4][0].view(np.uint64)), path
                assert got[path][1] == got[4][1]
    gm.set_sensor_path(0)
