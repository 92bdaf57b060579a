"""Holds the CPU oracle against the REFERENCE'S OWN map_likelihood.cpp, map_base.cpp,
odometry.cpp, statistics.cpp and particle_filter.cpp.

oracle/_ref/libqmcl_ref.so is compiled from those translation units where they lie under
/root/reference (oracle/Makefile, target `ref`) against the stand-in headers of
oracle/ref_stubs.  The reference's random engines are private and unseeded; the stand-in
<random> makes them read a script, and these tests write into it the words the shared
Philox contract hands the oracle for the same draws (oracle/ref_binding.py: *_script).
With that, the reference's sensor model, motion model, three resamplers, KLD stop rule,
clustering and pose estimate are compared with the oracle's on the same inputs:

* bit for bit wherever the two draw the same numbers (weights, totals, moved poses,
  resampled slots, particle counts, cluster counts, low-pass state);
* injected random poses (adaptive / KLD with w_diff > 0, global localisation): cell
  exact, heading within 1e-6 - the contract maps a word to a heading in double, the
  reference's uniform_real_distribution<float> does it in float;
* the estimate within 1e-12: the reference adds the cluster statistics in the order of
  its unordered_map, the oracle in canonical order.

CPU only.  Without a prebuilt library and without a reference tree they are skipped.
"""
import math

import numpy as np
import pytest

from oracle import binding as ob
from oracle import ref_binding as rb
from quickmcl_b200 import synthetic as syn

pytestmark = pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built and no reference tree")

RES_THETA = float(np.float32(math.pi / 18))


def ulp_distance(a, b):
    ia = np.ascontiguousarray(a, np.float32).view(np.int32).astype(np.int64)
    ib = np.ascontiguousarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(ia - ib)


def same_bits(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.dtype == b.dtype and a.shape == b.shape and a.tobytes() == b.tobytes()


def make_pair(width=200, seed=1, origin=None, **lp):
    """The same occupancy grid in the reference's MapLikelihood and in the oracle (brushfire
    field, like the reference); the oracle's table is then loaded into the reference's map
    so that both read one identical field."""
    occ, res, org = syn.make_map(width, seed=seed)
    if origin is not None:
        org = origin
    params = dict(syn.NODE_LIKELIHOOD)
    params.update(lp)
    rm = rb.RefMap(occ, res, org, **params)
    om = ob.OracleMap(**params)
    om.new_map(occ, res, org, field_mode=ob.FIELD_BRUSHFIRE)
    return occ, res, org, rm, om


# ------------------------------------------------------------------ map_base.cpp, map_likelihood.cpp:50-61
@pytest.mark.parametrize("width,seed,origin,sigma,max_d", [
    (120, 1, None, 0.1, 2.0), (200, 2, (-3.25, 1.5), 0.05, 1.0), (90, 3, (0.0, 0.0), 0.2, 0.5),
])
def test_map_tables_match_reference(width, seed, origin, sigma, max_d):
    occ, res, org, rm, om = make_pair(width, seed, origin, sigma_hit=sigma, max_obstacle_distance=max_d)
    assert same_bits(rm.state(), om.state())
    assert same_bits(rm.free_cells(), om.free_cells())
    # the reference's table is expf(-(d*d)/z) per cell in the stand-in (Eigen's packet exp in a real
    # build), the contract's is the correctly rounded value: never more than one ulp apart
    ref_field, orc_field = rm.field(), om.field()
    d = ulp_distance(ref_field, orc_field)
    assert d.max() <= 1
    print(f"likelihood table {width}^2: {np.mean(d > 0) * 100:.4f}% of cells one ulp from libm's expf")
    # identical tables -> identical visualisation grid and point queries (map_likelihood.cpp:132-160)
    rm.set_field(orc_field)
    assert same_bits(rm.likelihood_grid(), om.likelihood_grid())
    rng = np.random.default_rng(seed)
    h, w = occ.shape
    for _ in range(300):
        x = np.float32(org[0] + rng.uniform(-1.0, w * res + 1.0))
        y = np.float32(org[1] + rng.uniform(-1.0, h * res + 1.0))
        assert rm.state_at(x, y) == om.state_at(x, y)


# ------------------------------------------------------------------ map_likelihood.cpp:63-130
@pytest.mark.parametrize("width,n,beams,num_beams,origin,seed", [
    (200, 500, 360, 30, None, 1),           # node default: 30 of 360 beams
    (200, 300, 1081, 0, (-3.25, 1.5), 2),   # every beam, shifted origin
    (120, 400, 97, 200, (0.0, 0.0), 3),     # more beams asked for than there are -> step 1
    (120, 64, 7, 3, (2.0, -7.5), 4),        # ragged: 7 points, step 2
    (140, 1, 360, 30, None, 5),             # one particle
    (140, 200, 0, 30, None, 6),             # empty scan
])
def test_sensor_model_matches_reference_bit_for_bit(width, n, beams, num_beams, origin, seed):
    occ, res, org, rm, om = make_pair(width, seed, origin, num_beams=num_beams)
    rm.set_field(om.field())
    pose = syn.pick_truth_pose(occ, res, org)
    scan = (syn.make_scan(occ, res, org, pose, beams, 270.0) if beams else np.zeros((0, 2), np.float32))
    rng = np.random.default_rng(seed)
    h, w = occ.shape
    # tracking cloud + poses all over the grid and beyond its edges (unknown / occupied / outside)
    x, y, t = syn.gaussian_cloud(pose, n, seed=seed)
    k = n // 2
    x[:k] = org[0] + rng.uniform(-2.0, w * res + 2.0, k)
    y[:k] = org[1] + rng.uniform(-2.0, h * res + 2.0, k)
    t[:k] = rng.uniform(-math.pi, math.pi, k)
    a = ob.make_particles(x, y, t, np.full(n, 1.0 / n))
    b = a.copy()
    total_ref = rm.update_importance(scan, a)
    total_orc = om.update_importance(scan, b, trig_mode=ob.TRIG_LIBM)
    assert same_bits(a["weight"], b["weight"])
    assert np.float64(total_ref).tobytes() == np.float64(total_orc).tobytes()
    if beams:
        assert total_ref > 0
        states = np.array([om.state_at(px, py) for px, py in zip(a["x"], a["y"])])
        assert np.all(a["weight"][states == 2] == 0)  # unknown / outside: map_likelihood.cpp:118-119


def test_sensor_model_on_the_reference_own_table_differs_only_by_the_table_ulps():
    """No table injection: the reference's own expf table against the contract's."""
    occ, res, org, rm, om = make_pair(160, 7, None, num_beams=0)
    pose = syn.pick_truth_pose(occ, res, org)
    scan = syn.make_scan(occ, res, org, pose, 360, 270.0)
    x, y, t = syn.gaussian_cloud(pose, 400, seed=3)
    a = ob.make_particles(x, y, t, np.full(400, 1.0 / 400))
    b = a.copy()
    rm.update_importance(scan, a)
    om.update_importance(scan, b, trig_mode=ob.TRIG_LIBM)
    np.testing.assert_allclose(a["weight"], b["weight"], rtol=5e-7, atol=0)


# ------------------------------------------------------------------ odometry.cpp:44-102
@pytest.mark.parametrize("odom_new,odom_old,alpha", [
    ((0.25, 0.05, 0.1), (0.0, 0.0, 0.0), syn.NODE_ALPHA),
    ((1.0, 2.0, 0.5), (1.0, 2.0, -2.9), (0.2, 0.2, 0.2, 0.2)),         # rotation on the spot (< 1 cm)
    ((-0.4, 0.1, 3.1), (0.0, 0.0, -3.1), (0.05, 0.01, 0.3, 0.02)),     # backwards, across the +-pi seam
    ((5.0, 5.0, 1.0), (5.0, 5.0, 1.0), syn.NODE_ALPHA),                # no motion: zero variances
])
def test_motion_model_matches_reference_bit_for_bit(odom_new, odom_old, alpha):
    occ, res, org, rm, om = make_pair(120, 1)
    n = 3000
    x, y, t = syn.gaussian_cloud((2.0, 2.0, 3.0), n, sigma_theta=0.4, seed=9)
    t = ((t + np.pi) % (2 * np.pi) - np.pi).astype(np.float32)
    start = ob.make_particles(x, y, t, np.full(n, 1.0 / n))
    of = ob.OracleFilter(om, seed=21)
    of.set_particles(start)
    of.set_rng(21, 5)
    of.handle_odometry(odom_new, odom_old, alpha)
    moved_orc = of.get_particles()
    moved_ref = start.copy()
    rb.script(normals=rb.motion_script(21, 5, n))
    rb.sample_odometry(odom_new, odom_old, moved_ref, alpha)
    words, normals, under = rb.script_state()
    assert (words, normals, under) == (0, 3 * n, False)
    for k in ("x", "y", "theta", "weight"):
        assert same_bits(moved_ref[k], moved_orc[k]), k


# ------------------------------------------------------------------ statistics.cpp
@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (1000, 2), (20000, 3)])
def test_statistics_match_reference_bit_for_bit(n, seed):
    import ctypes as C
    rng = np.random.default_rng(seed)
    w = rng.random(n)
    p = ob.make_particles(rng.normal(3.0, 2.0, n), rng.normal(-1.0, 0.7, n), rng.uniform(-np.pi, np.pi, n),
                          w / w.sum())
    for fin in (0, 1):
        out = np.zeros(15)
        ob.lib().orc_statistics(p.ctypes.data_as(C.POINTER(ob.Particle)), n, fin,
                                out.ctypes.data_as(C.POINTER(C.c_double)))
        assert same_bits(out, rb.statistics(p, fin)), fin


# ------------------------------------------------------------------ particle_filter.cpp
def make_filters(rtype, nmin, nmax, width=160, seed=4, kld_epsilon=0.05, kld_z=0.95, **lp):
    occ, res, org, rm, om = make_pair(width, seed, None, **lp)
    rm.set_field(om.field())
    kw = dict(resample_type=rtype, particle_count_min=nmin, particle_count_max=nmax, resample_count=2,
              alpha_fast=0.1, alpha_slow=0.001, kld_epsilon=kld_epsilon, kld_z=kld_z, res_xy=0.5,
              res_theta=RES_THETA)
    of = ob.OracleFilter(om, seed=33)
    of.set_parameters(**kw)
    of.set_trig_mode(ob.TRIG_LIBM)
    rf = rb.RefFilter(rm)
    rf.set_parameters(**kw)
    return occ, res, org, rm, om, rf, of


def w_diff_of(lowpass):
    w_slow, w_fast = lowpass
    with np.errstate(divide="ignore", invalid="ignore"):
        q = np.float64(w_fast) / np.float64(w_slow)
    d = 1.0 - q
    return d if d > 0.0 else 0.0  # std::max(0.0, NaN) is 0.0 as well


def compare_sets(ref_p, orc_p, injected=None):
    assert len(ref_p) == len(orc_p)
    if injected is None:
        injected = np.zeros(len(ref_p), bool)
    injected = injected[:len(ref_p)]
    for k in ("x", "y", "weight"):
        assert same_bits(ref_p[k], orc_p[k]), k
    assert same_bits(ref_p["theta"][~injected], orc_p["theta"][~injected])
    if injected.any():
        np.testing.assert_allclose(ref_p["theta"][injected], orc_p["theta"][injected], rtol=0, atol=1e-6)


def compare_estimates(rf, of):
    ok_r, pose_r, cov_r = rf.get_estimated_pose()
    ok_o, pose_o, cov_o = of.get_estimated_pose()
    assert bool(ok_r) == bool(ok_o)
    if ok_r:
        np.testing.assert_allclose(pose_r, pose_o, rtol=0, atol=1e-12)
        np.testing.assert_allclose(cov_r, cov_o, rtol=0, atol=1e-12)
    assert rf.num_clusters() == of.num_clusters()


def low_variance_overruns(p, seed, step):
    """True when the low-variance walk needs more weight than the set holds: the last target
    r + (M - 1) * minv lies beyond the running sum of all weights (particle_filter.cpp:224-235).
    The reference then throws from .at(i) (its assert is compiled out in Release); the oracle and
    the kernels stay on the last particle (documented deviation 2)."""
    m = len(p)
    minv = np.float64(np.float32(1.0) / np.float32(m))
    r = rb.u53(*[rb.rng_words(seed, step, rb.P_LV_OFFSET, 1)[:, k] for k in (0, 1)])[0] * minv
    return bool(r + (m - 1) * minv > np.cumsum(p["weight"])[-1])


def resample_both(rf, of, rtype, seed, step, lowpass, n_slots):
    """One resample_and_cluster on both sides with the same draws; returns the inject mask."""
    of.set_lowpass(*lowpass)
    rf.set_lowpass(*lowpass)
    of.set_rng(seed, step)
    if rtype == 0:
        rb.script(words=rb.low_variance_script(seed, step))
        injected = None
    else:
        words, injected = rb.draw_script(seed, step, n_slots, w_diff_of(lowpass))
        rb.script(words=words)
    before = of.get_particles()
    of.resample_and_cluster()
    ok = rf.resample_and_cluster()
    assert not rb.script_state()[2], "the reference drew more numbers than the script holds"
    if rtype == 0 and low_variance_overruns(before, seed, step):
        # the reference gave up half way; what it wrote until then is the oracle's, and the oracle
        # finished the set on the last particle
        assert not ok
        idx = of.resample_indices()
        first_clamped = int(np.argmax(idx == len(before) - 1))
        half = rf.get_particles()[:first_clamped]
        done = before[idx[:first_clamped]]
        for k in ("x", "y", "theta"):
            assert same_bits(half[k], done[k]), k
        rf.set_particles(of.get_particles())
        rf.cluster()
    else:
        assert ok
    return injected


def weighted_cloud(occ, res, org, rm, om, n, seed):
    pose = syn.pick_truth_pose(occ, res, org)
    scan = syn.make_scan(occ, res, org, pose, 360, 270.0)
    x, y, t = syn.gaussian_cloud(pose, n, seed=seed)
    p = ob.make_particles(x, y, t, np.full(n, 1.0 / n))
    total = om.update_importance(scan, p, trig_mode=ob.TRIG_LIBM)
    p["weight"] /= total
    return pose, scan, p


@pytest.mark.parametrize("n,seed,short", [(1, 1, False), (2, 2, False), (1500, 3, False), (4096, 5, False),
                                          (4096, 4, True)])
def test_low_variance_matches_reference(n, seed, short):
    """`short`: weights as the sensor update leaves them - divided by the total taken BEFORE the
    penalties for poses in walls (map_likelihood.cpp:113-123), so they add up to less than one and
    the walk can run off the end: the reference throws there, the oracle clamps."""
    occ, res, org, rm, om, rf, of = make_filters(0, n, n)
    _, _, p = weighted_cloud(occ, res, org, rm, om, n, seed)
    if not short:
        p["weight"] /= np.cumsum(p["weight"])[-1]
        p["weight"][-1] += 1e-6  # the running sum reaches past every target
    assert low_variance_overruns(p, 33, seed) == short
    of.set_particles(p)
    rf.set_particles(p)
    resample_both(rf, of, 0, 33, seed, (0.0, 0.0), n)
    assert rb.script_state()[0] == 1
    compare_sets(rf.get_particles(), of.get_particles())
    compare_estimates(rf, of)


@pytest.mark.parametrize("n,nmin,lowpass,seed", [
    (2000, 1500, (0.0, 0.0), 1),      # 0 / 0 -> w_diff 0, nothing injected
    (2000, 1500, (1.0, 2.0), 2),      # w_fast > w_slow -> w_diff 0
    (2500, 1200, (1.0, 0.7), 3),      # 30 % random poses
    (800, 2000, (1.0, 0.0), 4),       # w_diff 1: every slot injected, more slots than particles
])
def test_adaptive_matches_reference(n, nmin, lowpass, seed):
    occ, res, org, rm, om, rf, of = make_filters(1, nmin, 5000)
    _, _, p = weighted_cloud(occ, res, org, rm, om, n, seed)
    of.set_particles(p)
    rf.set_particles(p)
    injected = resample_both(rf, of, 1, 33, seed, lowpass, nmin)
    assert rf.num_particles() == nmin
    compare_sets(rf.get_particles(), of.get_particles(), injected)
    assert int(injected.sum()) == int(np.sum(of.resample_indices() < 0))
    # the low-pass filters are reset exactly when a random pose could have been drawn
    assert same_bits(rf.lowpass(), of.lowpass())
    # estimates agree once both hold the same headings
    rf.set_particles(of.get_particles())
    rf.cluster()
    of.cluster()
    compare_estimates(rf, of)


@pytest.mark.parametrize("n,nmin,nmax,lowpass,eps,z,seed", [
    (3000, 100, 3000, (0.0, 0.0), 0.05, 0.95, 1),     # tracking: stops after a few hundred
    (3000, 100, 3000, (1.0, 0.9), 0.05, 0.95, 2),     # 10 % random poses keep opening bins
    (3000, 100, 800, (1.0, 0.5), 0.01, 0.99, 3),      # runs into particle_count_max
    (500, 400, 2000, (0.0, 0.0), 0.2, 0.5, 4),        # stops at particle_count_min
    (3000, 100, 3000, (1.0, 0.98), 0.1, 0.9, 5),      # 2 % random poses, stops early all the same
])
def test_kld_matches_reference(n, nmin, nmax, lowpass, eps, z, seed):
    occ, res, org, rm, om, rf, of = make_filters(2, nmin, nmax, kld_epsilon=eps, kld_z=z)
    _, _, p = weighted_cloud(occ, res, org, rm, om, n, seed)
    of.set_particles(p)
    rf.set_particles(p)
    injected = resample_both(rf, of, 2, 33, seed, lowpass, nmax)
    kept = rf.num_particles()
    assert kept == of.num_particles()
    print(f"KLD kept {kept} of at most {nmax} (w_diff {w_diff_of(lowpass):.2f})")
    compare_sets(rf.get_particles(), of.get_particles(), injected)
    assert same_bits(rf.lowpass(), of.lowpass())
    if not injected[:kept].any():
        compare_estimates(rf, of)


@pytest.mark.parametrize("width,n", [(400, 20000), (2000, 100000)])
def test_c2_kld_cycle_matches_reference_at_size(width, n):
    """BASELINE.json's C2 itself (100 000 particles, 2000 x 2000 map at 5 cm, every hit of a 1081-beam
    270 degree scan, KLD) and a smaller cut of it: one whole cycle of the compiled reference against the
    oracle - moved poses, weights, total, kept set and estimate.  The reference does it in about a second."""
    occ, res, org, rm, om, rf, of = make_filters(2, 100, n, width=width, seed=1, num_beams=0)
    pose = syn.pick_truth_pose(occ, res, org)
    scan = syn.make_scan(occ, res, org, pose, 1081, 270.0)
    x, y, t = syn.gaussian_cloud(pose, n)
    start = ob.make_particles(x, y, t, np.full(n, 1.0 / n))
    of.set_particles(start)
    rf.set_particles(start)
    of.set_rng(33, 0)
    of.handle_odometry((0.25, 0.05, 0.1), (0.0, 0.0, 0.0), syn.NODE_ALPHA)
    rb.script(normals=rb.motion_script(33, 0, n))
    rf.handle_odometry((0.25, 0.05, 0.1), (0.0, 0.0, 0.0), syn.NODE_ALPHA)
    compare_sets(rf.get_particles(), of.get_particles())
    of.update_importance_from_observations(scan)
    rf.update_importance_from_observations(scan)
    compare_sets(rf.get_particles(), of.get_particles())
    assert same_bits(rf.lowpass(), of.lowpass())
    resample_both(rf, of, 2, 33, 1, tuple(of.lowpass()), n)
    assert 100 < rf.num_particles() < n
    compare_sets(rf.get_particles(), of.get_particles())
    compare_estimates(rf, of)


def test_zero_weight_input_is_left_alone_like_the_reference():
    """Both resamplers with a running sum return early on an all-zero input
    (particle_filter.cpp:255-259, :307-311)."""
    for rtype in (1, 2):
        occ, res, org, rm, om, rf, of = make_filters(rtype, 50, 200)
        x, y, t = syn.gaussian_cloud((2.0, 2.0, 0.0), 80, seed=1)
        p = ob.make_particles(x, y, t, np.zeros(80))
        of.set_particles(p)
        rf.set_particles(p)
        resample_both(rf, of, rtype, 33, 1, (0.0, 0.0), 200)
        assert rb.script_state()[0] == 0
        assert rf.num_particles() == of.num_particles()


def test_global_localization_matches_reference():
    occ, res, org, rm, om, rf, of = make_filters(2, 100, 4000)
    of.set_rng(33, 2)
    of.global_localization()
    rb.script(words=rb.global_script(33, 2, 4000))
    rf.global_localization()
    assert rb.script_state() == (8000, 0, False)
    a, b = rf.get_particles(), of.get_particles()
    assert len(a) == len(b) == 4000
    for k in ("x", "y", "weight"):
        assert same_bits(a[k], b[k]), k
    np.testing.assert_allclose(a["theta"], b["theta"], rtol=0, atol=1e-6)
    assert np.all(np.abs(a["theta"]) <= np.float32(np.pi))


def test_initialise_factors_the_covariance():
    """ParticleFilter::initialise draws mean + T z with T T^T = covariance (random.h:57-72).  The
    stand-in eigen solver is not Eigen's, so only that property is held, on both sides."""
    occ, res, org, rm, om, rf, of = make_filters(2, 3, 10)
    cov = np.array([[0.25, 0.05, 0.0], [0.05, 0.16, 0.01], [0.0, 0.01, 0.04]], np.float32)
    pose = np.array([1.0, -2.0, 0.3], np.float32)
    rb.script(normals=np.eye(3).reshape(-1))
    rf.initialise(pose, cov)
    p = rf.get_particles()
    assert len(p) == 3 and rb.script_state() == (0, 9, False)
    T = np.stack([p["x"] - pose[0], p["y"] - pose[1], p["theta"] - pose[2]]).astype(np.float64)
    np.testing.assert_allclose(T @ T.T, cov, atol=2e-6)
    np.testing.assert_allclose(p["weight"], 1.0 / 3)
    import ctypes as C
    fac = np.zeros(9, np.float32)
    ob.lib().orc_cov_factor(cov.reshape(9).ctypes.data_as(C.POINTER(C.c_float)),
                            fac.ctypes.data_as(C.POINTER(C.c_float)))
    F = fac.reshape(3, 3).astype(np.float64)
    np.testing.assert_allclose(F @ F.T, cov, atol=2e-6)


@pytest.mark.parametrize("rtype,nmin,nmax", [(0, 1500, 1500), (1, 1500, 1500), (2, 100, 3000)])
def test_scan_cycles_match_reference(rtype, nmin, nmax):
    """Six whole cycles (odometry -> sensor -> low-pass -> resample + cluster -> estimate) of the
    reference's ParticleFilter against the oracle, every intermediate set compared.  The low-pass
    filters evolve on their own; when they make the resampler inject random poses, the reference is
    handed the oracle's set afterwards so that the 1e-7 heading difference of those poses does not
    leak into the next cycle."""
    occ, res, org, rm, om, rf, of = make_filters(rtype, nmin, nmax, num_beams=60)
    pose = np.array(syn.pick_truth_pose(occ, res, org), np.float64)
    n0 = nmax if rtype == 2 else nmin
    x, y, t = syn.gaussian_cloud(pose, n0, seed=8)
    start = ob.make_particles(x, y, t, np.full(n0, 1.0 / n0))
    of.set_particles(start)
    rf.set_particles(start)
    odom = np.zeros(3)
    step = 0
    for cyc in range(6):
        delta = np.array([0.06 * math.cos(pose[2]), 0.06 * math.sin(pose[2]), 0.05])
        new_odom = odom + delta
        pose = pose + delta
        n = of.num_particles()
        of.set_rng(33, step)
        of.handle_odometry(new_odom, odom, syn.NODE_ALPHA)
        rb.script(normals=rb.motion_script(33, step, n))
        rf.handle_odometry(new_odom, odom, syn.NODE_ALPHA)
        compare_sets(rf.get_particles(), of.get_particles())
        step += 1
        odom = new_odom
        scan = syn.make_scan(occ, res, org, tuple(pose), 360, 270.0, seed=cyc)
        of.update_importance_from_observations(scan)
        rf.update_importance_from_observations(scan)
        compare_sets(rf.get_particles(), of.get_particles())
        assert same_bits(rf.lowpass(), of.lowpass())
        lowpass = tuple(of.lowpass())
        if cyc == 3:  # a kidnapped-robot moment: short-term average well below the long-term one
            lowpass = (lowpass[0], lowpass[0] * 0.8)
        injected = resample_both(rf, of, rtype, 33, step, lowpass, nmax if rtype == 2 else nmin)
        step += 1
        compare_sets(rf.get_particles(), of.get_particles(), injected)
        assert same_bits(rf.lowpass(), of.lowpass())
        if injected is not None and injected[:rf.num_particles()].any():
            rf.set_particles(of.get_particles())
            rf.cluster()
            of.cluster()
        compare_estimates(rf, of)


# ------------------------------------------------------------------ pose_restore.cpp
class RecordingFilter:
    """What PoseCheckpoint needs of an IParticleFilter; remembers how it was initialised."""

    def __init__(self, estimate=None):
        self.estimate = estimate
        self.initialised = []

    def get_estimated_pose(self):
        if self.estimate is None:
            return False, np.zeros(3), np.zeros((3, 3))
        return (True,) + self.estimate

    def initialise(self, pose, cov):
        self.initialised.append((np.asarray(pose), np.asarray(cov)))


@pytest.mark.parametrize("stored", [
    None,                                                   # no ~initial_pose: pose 0, identity covariance
    [1.5, -2.25, 0.7, 0.04, 0.09, 0.01],                    # what store_pose wrote
    [1.5, -2.25],                                           # short vector: the rest falls back
    [float("nan"), 3.0, float("nan"), 0.2, float("nan"), float("nan")],
    [],                                                     # present but empty
    [0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7],                    # longer than six: the tail is ignored
    [1e-3 + 1e-12, 7.000000001, -3.14159265358979, 1e-9, 0.25, 0.01],   # values that round in float
])
def test_pose_checkpoint_restores_like_the_reference_pose_restorer(stored):
    """quickmcl_b200.cycle.PoseCheckpoint against PoseRestorer::restore_pose itself
    (pose_restore.cpp:57-97): the pose and the float covariance the filter is re-initialised
    with, and the warnings, for missing, short, NaN and complete parameter vectors."""
    from quickmcl_b200.cycle import PoseCheckpoint
    calls, pose_ref, cov_ref, warn_ref = rb.pose_restore(stored)
    rec = RecordingFilter()
    cp = PoseCheckpoint(rec)
    cp.restore(stored)
    assert calls == 1 and len(rec.initialised) == 1
    pose, cov = rec.initialised[0]
    assert same_bits(np.asarray(pose, np.float32), pose_ref)
    assert same_bits(np.asarray(cov, np.float32).reshape(3, 3), cov_ref)
    assert cp.warnings == warn_ref


@pytest.mark.parametrize("estimate", [
    None,
    (np.array([1.25, -0.5, 2.0]), np.array([[0.04, 0.01, 0.0], [0.01, 0.09, 0.0], [0.0, 0.0, 0.02]])),
])
def test_pose_checkpoint_stores_like_the_reference_pose_restorer(estimate):
    """PoseRestorer::store_pose (pose_restore.cpp:37-55): six numbers, or nothing and a warning."""
    from quickmcl_b200.cycle import PoseCheckpoint
    if estimate is None:
        stored_ref, warn_ref = rb.pose_store(False, np.zeros(3), np.zeros((3, 3)))
    else:
        stored_ref, warn_ref = rb.pose_store(True, *estimate)
    cp = PoseCheckpoint(RecordingFilter(estimate))
    assert cp.store() == stored_ref
    assert cp.warnings == warn_ref
    if estimate is not None:                      # and the round trip through both
        _, pose_ref, cov_ref, _ = rb.pose_restore(stored_ref)
        rec = RecordingFilter()
        PoseCheckpoint(rec).restore(cp.store())
        assert same_bits(np.asarray(rec.initialised[0][0], np.float32), pose_ref)
        assert same_bits(np.asarray(rec.initialised[0][1], np.float32).reshape(3, 3), cov_ref)


# ------------------------------------------------------------------ laser_handler.cpp
class LoggingFilter:
    """An IParticleFilter that only writes down how ScanCycle calls it, in the format of
    oracle/ref_shim_node.cpp's recorder."""

    def __init__(self):
        self.log = []
        self.present = False

    def num_particles(self):
        return 1 if self.present else 0

    def global_localization(self):
        self.log.append("global_localization")
        self.present = True

    def handle_odometry(self, odom_new, odom_old, alpha):
        vals = [float(v) for v in odom_new] + [float(v) for v in odom_old] + [float(np.float32(a)) for a in alpha]
        self.log.append("handle_odometry " + " ".join(v.hex() for v in vals))

    def update_importance_from_observations(self, cloud):
        self.log.append(f"update_importance_from_observations {len(cloud)}")

    def resample_and_cluster(self):
        self.log.append("resample_and_cluster")

    def cluster(self):
        self.log.append("cluster")

    def get_estimated_pose(self):
        return False, None, None


def _canonical(lines):
    """C's %a and Python's float.hex() spell the same number differently: compare values."""
    out = []
    for ln in lines:
        head, *rest = ln.split(" ")
        if head == "handle_odometry":
            out.append((head, tuple(float.fromhex(v) for v in rest)))
        else:
            out.append((ln,))
    return out


@pytest.mark.parametrize("resample_count,seed", [(2, 0), (1, 1), (3, 2), (5, 3)])
def test_scan_cycle_makes_the_calls_of_the_reference_laser_handler(resample_count, seed):
    """quickmcl_b200.cycle.ScanCycle against LaserHandler::Impl::cloud_callback itself
    (laser_handler.cpp:179-288, compiled into oracle/_ref with recorders for its collaborators):
    for a random walk with stops, turns on the spot, jumps across +-pi, emptied particle sets and
    pose resets, the same IParticleFilter calls with the same arguments, in the same order, and
    the same recompute_transform flag handed to publish_estimated_pose."""
    from quickmcl_b200.cycle import MotionModelParameters, ScanCycle
    rng = np.random.default_rng(seed)
    alpha = (0.05, 0.1, 0.02, 0.05)
    min_trans, min_rot = 0.2, math.pi / 6          # the reference keeps them as floats; so does on_scan
    ref = rb.RefScanCycle(resample_count, min_trans, min_rot, alpha)
    lf = LoggingFilter()
    cyc = ScanCycle(lf, MotionModelParameters(alpha=alpha, min_trans=min_trans, min_rot=min_rot),
                    resample_count=resample_count)
    odom = np.array([rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-math.pi, math.pi)])
    kinds = ["stay", "creep", "drive", "turn", "wrap", "threshold"]
    for step in range(120):
        kind = kinds[rng.integers(len(kinds))]
        if kind == "creep":
            odom = odom + np.array([rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), rng.uniform(-0.2, 0.2)])
        elif kind == "drive":
            odom = odom + np.array([rng.uniform(-0.6, 0.6), rng.uniform(-0.6, 0.6), rng.uniform(-0.3, 0.3)])
        elif kind == "turn":
            odom = odom + np.array([0.0, 0.0, rng.uniform(-1.5, 1.5)])
        elif kind == "wrap":                      # odometry yaw jumps by a whole turn: no motion at all
            odom = odom + np.array([0.0, 0.0, 2 * math.pi * rng.choice([-1, 1])])
        elif kind == "threshold":                 # exactly on / next to the translation threshold
            last = cyc.odom_at_last_filter_run if cyc.odom_at_last_filter_run is not None else odom
            edge = float(np.float32(min_trans))
            odom = np.array([last[0] + rng.choice([edge, np.nextafter(edge, 1.0), -edge, 0.2, -0.2]), last[1],
                             last[2]])
        if rng.random() < 0.08:
            ref.force_pose_reset()
            cyc.force_pose_reset()
        if rng.random() < 0.1:                    # the filter lost every particle
            ref.set_particles_present(False)
            lf.present = False
        n_points = int(rng.integers(0, 400))
        lf.log.clear()
        ref_lines = ref.on_scan(odom, n_points, now=float(step))
        result = cyc.on_scan(odom, np.zeros((n_points, 2), np.float32))
        ref_calls = [ln for ln in ref_lines if not ln.startswith(("publish_", "log: "))]
        assert _canonical(ref_calls) == _canonical(lf.log), (step, kind)
        publish = [ln for ln in ref_lines if ln.startswith("publish_estimated_pose")]
        assert publish == ["publish_estimated_pose recompute" if result.recompute_transform
                           else "publish_estimated_pose"], (step, kind)
        assert ("log: No particles... Trying global localisation!" in ref_lines) == result.global_localization
        assert result.processed == any(ln.startswith("handle_odometry") for ln in ref_lines)
        assert result.resampled == ("resample_and_cluster" in ref_lines)


def test_scan_cycle_reference_drops_the_scan_without_odometry():
    """laser_handler.cpp:183-187: no odometry transform, nothing happens (the ROS-free ScanCycle is
    handed the odometry by its caller, so this branch has no counterpart there)."""
    ref = rb.RefScanCycle()
    assert ref.on_scan((0.0, 0.0, 0.0), 10, odom_ok=False) == [
        "log: ERROR: Failed to get odometry pose while handling laser"]
