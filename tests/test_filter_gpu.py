"""Parity of the filter stages (K5, K6, K9-K14) against the CPU oracle through the
C ABI, on identical inputs and one shared Philox stream.

Bars (BASELINE.json north_star): resampling indices and KLD bin counts
bit-exact given identical weights; poses within 1e-6 after the motion model;
pose estimate within 1e-4 m / 1e-4 rad.
"""
import math

import numpy as np
import pytest

import quickmcl_b200 as q
from quickmcl_b200 import synthetic as syn
from oracle import binding as ob

pytestmark = pytest.mark.gpu

RES_THETA = float(np.float32(math.pi / 18))


def pf_params(**kw):
    base = dict(resample_type=2, particle_count_min=100, particle_count_max=5000,
                resample_count=2, alpha_fast=0.1, alpha_slow=0.001, kld_epsilon=0.05,
                kld_z=0.95, res_xy=0.5, res_theta=RES_THETA)
    base.update(kw)
    return base


def make_pair(width=400, seed=1, num_beams=0, rng_seed=3, **pf):
    occ, res, origin = syn.make_map(width, seed=seed)
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=num_beams)
    om = ob.OracleMap(**lp)
    om.new_map(occ, res, origin, field_mode=ob.FIELD_EXACT_EDT)
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(**lp))
    gm.load_field(om.field(), om.state(), res, origin)
    of = ob.OracleFilter(om, seed=rng_seed)
    gf = q.create_particle_filter(gm, seed=rng_seed)
    p = pf_params(**pf)
    of.set_parameters(**p)
    gf.set_parameters(q.ParticleFilterParameters(
        resample_type=q.ResampleType(p["resample_type"]),
        particle_count_min=p["particle_count_min"], particle_count_max=p["particle_count_max"],
        resample_count=p["resample_count"], alpha_fast=p["alpha_fast"],
        alpha_slow=p["alpha_slow"], kld_epsilon=p["kld_epsilon"], kld_z=p["kld_z"],
        space_partitioning_resolution_xy=p["res_xy"],
        space_partitioning_resolution_theta=p["res_theta"]))
    return occ, res, origin, om, gm, of, gf


def set_both(of, gf, x, y, t, w):
    of.set_particles(ob.make_particles(x, y, t, w))
    gf.set_particles(q.make_particles(x, y, t, w))


def assert_same_particles(gp, op, pose_tol=0.0, w_rel=1e-12):
    assert len(gp) == len(op)
    for k in ("x", "y", "theta"):
        if pose_tol == 0.0:
            assert np.array_equal(gp[k], op[k]), k
        else:
            assert np.max(np.abs(gp[k].astype(np.float64) - op[k])) <= pose_tol, k
    denom = np.maximum(np.abs(op["weight"]), 1e-300)
    err = np.where(op["weight"] == 0, np.abs(gp["weight"]), np.abs(gp["weight"] - op["weight"]) / denom)
    assert err.max() <= w_rel, err.max()


# ---------------------------------------------------------------- a-10
def test_effective_particles_reference_vectors_and_oracle():
    """effective_particles (particle.cpp:32-39): the reference's own known answers
    (test/quickmcl/test_particle.cpp:12-27) and the oracle on a sensor-weighted cloud."""
    import ctypes as C
    occ, res, origin, om, gm, of, gf = make_pair()
    z = np.zeros(3)
    gf.set_particles(q.make_particles(z, z, z, np.full(3, 1.0 / 3)))   # equalised
    assert abs(gf.effective_particles() - 3.0) <= 4 * np.spacing(3.0)
    gf.set_particles(q.make_particles(z, z, z, np.array([1.0, 0.0, 0.0])))
    assert gf.effective_particles() == 1.0
    pose = syn.pick_truth_pose(occ, res, origin)
    scan = syn.make_scan(occ, res, origin, pose, 360, 360.0)
    for n in (1, 257, 5000, 100_003):
        x, y, t = syn.gaussian_cloud(pose, n)
        set_both(of, gf, x, y, t, np.full(n, 1.0 / n))
        gf.update_importance_from_observations(scan)
        gp = gf.get_particles()
        want = ob.lib().orc_effective_particles(gp.ctypes.data_as(C.POINTER(ob.Particle)), n)
        got = gf.effective_particles()
        assert abs(got - want) <= 1e-12 * want, (n, got, want)
    gf.set_particles(q.make_particles(z[:0], z[:0], z[:0], z[:0]))
    assert math.isinf(gf.effective_particles())                         # 1 / 0, as the loop gives


# ---------------------------------------------------------------- K5 / K6
def test_initialise_matches_oracle():
    _, _, _, om, gm, of, gf = make_pair(particle_count_min=5000)
    factor = np.diag(np.sqrt([0.25, 0.25, 0.01])).astype(np.float32)
    factor[0, 1] = 0.03
    pose = (1.5, -2.25, 3.0)
    of.initialise_factor(pose, factor)
    gf.initialise_factor(pose, factor)
    gp, op = gf.get_particles(), of.get_particles()
    assert_same_particles(gp, op, pose_tol=1e-6, w_rel=0.0)
    exact = np.mean((gp["x"] == op["x"]) & (gp["y"] == op["y"]) & (gp["theta"] == op["theta"]))
    print(f"initialise: {exact:.5f} of poses bit-identical")
    assert exact > 0.999
    assert np.all(gp["weight"] == 1.0 / 5000)
    assert np.all((gp["theta"] >= -np.float32(math.pi)) & (gp["theta"] < math.pi))


def test_initialise_from_covariance_is_statistically_right():
    _, _, _, om, gm, of, gf = make_pair(particle_count_min=200000)
    cov = np.array([[0.25, 0.05, 0.0], [0.05, 0.16, 0.0], [0.0, 0.0, 0.01]], np.float32)
    gf.initialise((0.0, 0.0, 0.0), cov)
    gp = gf.get_particles()
    emp = np.cov(np.stack([gp["x"], gp["y"], gp["theta"]]).astype(np.float64))
    assert np.allclose(emp, cov, atol=5e-3)
    of.initialise((0.0, 0.0, 0.0), cov)
    op = of.get_particles()
    assert_same_particles(gp, op, pose_tol=2e-6, w_rel=0.0)


def test_global_localization_matches_oracle_exactly():
    occ, res, origin, om, gm, of, gf = make_pair(particle_count_max=20000)
    of.global_localization()
    gf.global_localization()
    gp, op = gf.get_particles(), of.get_particles()
    assert_same_particles(gp, op, pose_tol=0.0, w_rel=0.0)
    # every pose is the corner of a free cell
    st = np.array([om.state_at(float(a) + res / 2, float(b) + res / 2)
                   for a, b in zip(gp["x"][:500], gp["y"][:500])])
    assert np.all(st == 0)


@pytest.mark.parametrize("new,old", [
    ((0.25, 0.05, 0.1), (0.0, 0.0, 0.0)),        # SURVEY §8d odometry step
    ((-0.3, 0.0, 3.1), (0.0, 0.0, -3.1)),        # backwards + angle wrap
    ((0.001, 0.001, 0.5), (0.0, 0.0, 0.0)),      # turn on the spot (trans < 0.01)
])
def test_motion_matches_oracle(new, old):
    _, _, _, om, gm, of, gf = make_pair()
    n = 30000
    x, y, t = syn.gaussian_cloud((0.3, -0.7, 1.0), n, sigma_theta=2.0)
    set_both(of, gf, x, y, t, np.full(n, 1.0 / n))
    of.set_rng(77, 5)
    gf.set_rng(77, 5)
    for _ in range(2):
        of.handle_odometry(new, old, syn.NODE_ALPHA)
        gf.handle_odometry(new, old, syn.NODE_ALPHA)
    gp, op = gf.get_particles(), of.get_particles()
    assert_same_particles(gp, op, pose_tol=1e-6, w_rel=0.0)
    exact = np.mean((gp["x"] == op["x"]) & (gp["y"] == op["y"]) & (gp["theta"] == op["theta"]))
    print(f"motion: {exact:.5f} of poses bit-identical")
    assert exact > 0.999


# ----------------------------------------------------------- K9 / K10 / K12
def weights_from_sensor(of, gf, occ, res, origin, n, cloud="tracking", beams=360, fov=360.0):
    pose = syn.pick_truth_pose(occ, res, origin)
    scan = syn.make_scan(occ, res, origin, pose, beams, fov)
    if cloud == "tracking":
        x, y, t = syn.gaussian_cloud(pose, n)
    else:
        x, y, t = syn.uniform_cloud(occ, res, origin, n)
    set_both(of, gf, x, y, t, np.full(n, 1.0 / n))
    of.update_importance_from_observations(scan)
    # identical weights are the premise of the bit-exactness claim: feed the
    # oracle's normalised weights to both sides
    op = of.get_particles()
    gf.set_particles(op.view(q.api.PARTICLE_DTYPE))
    return pose, scan, op


@pytest.mark.parametrize("n,cloud", [(2000, "tracking"), (50000, "tracking"), (30000, "global")])
def test_low_variance_indices_bit_exact(n, cloud):
    occ, res, origin, om, gm, of, gf = make_pair(resample_type=0, particle_count_min=n,
                                                 particle_count_max=n)
    weights_from_sensor(of, gf, occ, res, origin, n, cloud)
    of.set_rng(5, 9)
    gf.set_rng(5, 9)
    of.resample_and_cluster()
    gf.resample_and_cluster()
    assert np.array_equal(gf.resample_indices(), of.resample_indices())
    assert_same_particles(gf.get_particles(), of.get_particles())


def test_low_variance_runs_off_the_end_clamped():
    """Sum of weights < 1 (penalised particles): the reference would throw
    (SURVEY §5.9-5); both sides clamp to the last particle."""
    _, _, _, om, gm, of, gf = make_pair(resample_type=0)
    n = 1000
    w = np.full(n, 0.5 / n)
    set_both(of, gf, np.zeros(n), np.zeros(n), np.zeros(n), w)
    for f in (of, gf):
        f.set_rng(1, 1)
        f.resample_and_cluster()
    gi, oi = gf.resample_indices(), of.resample_indices()
    assert np.array_equal(gi, oi)
    assert gi[-1] == n - 1 and (gi == n - 1).sum() > n // 3


def test_running_sum_duplicate_keys_and_zero_weights():
    """particle.cpp:77: a weight too small to move the running sum makes the next
    record share its key and be dropped; zero weights are skipped."""
    _, _, _, om, gm, of, gf = make_pair(resample_type=1, particle_count_min=4000,
                                        alpha_fast=0.0, alpha_slow=0.0)
    w = np.array([0.0, 0.3, 1e-40, 1e-30, 0.2, 0.0, 0.0, 0.25, 5e-324, 0.25, 0.0])
    n = len(w)
    x = np.arange(n, dtype=np.float32)
    set_both(of, gf, x, x, np.zeros(n), w)
    for f in (of, gf):
        f.set_rng(11, 3)
        f.resample_and_cluster()
    gi, oi = gf.resample_indices(), of.resample_indices()
    assert np.array_equal(gi, oi)
    # keys: 1->0, 2->0.3, (3 and 4 share key 0.3: dropped), 7->0.5, 8->0.75, (9 shares 0.75)
    assert set(np.unique(gi)) == {1, 2, 7, 8}
    assert_same_particles(gf.get_particles(), of.get_particles())


# ------------------------------------------------------- adaptive / KLD
@pytest.mark.parametrize("w_slow,w_fast", [(0.0, 0.0), (1.0, 0.6), (1.0, 2.0)])
def test_adaptive_matches_oracle(w_slow, w_fast):
    occ, res, origin, om, gm, of, gf = make_pair(resample_type=1, particle_count_min=20000,
                                                 particle_count_max=20000)
    weights_from_sensor(of, gf, occ, res, origin, 20000, "global")
    for f in (of, gf):
        f.set_lowpass(w_slow, w_fast)
        f.set_rng(21, 4)
        f.resample_and_cluster()
    gi, oi = gf.resample_indices(), of.resample_indices()
    assert np.array_equal(gi, oi)
    if w_fast < w_slow:
        assert (gi == -1).mean() == pytest.approx(1 - w_fast / w_slow, abs=0.02)
        assert np.all(gf.lowpass() == 0) and np.all(of.lowpass() == 0)
    else:
        assert (gi >= 0).all()
    assert_same_particles(gf.get_particles(), of.get_particles())


@pytest.mark.parametrize("n_in,cloud,nmax,inject", [
    (5000, "tracking", 5000, False),
    (50000, "tracking", 50000, False),
    (20000, "global", 20000, False),
    (20000, "global", 20000, True),
])
def test_kld_matches_oracle(n_in, cloud, nmax, inject):
    occ, res, origin, om, gm, of, gf = make_pair(resample_type=2, particle_count_min=100,
                                                 particle_count_max=nmax)
    weights_from_sensor(of, gf, occ, res, origin, n_in, cloud)
    for f in (of, gf):
        if inject:
            f.set_lowpass(1.0, 0.8)
        f.set_rng(31, 6)
        f.resample_and_cluster()
    assert gf.num_particles() == of.num_particles()
    print(f"KLD kept {gf.num_particles()} of {nmax}")
    assert np.array_equal(gf.resample_indices(), of.resample_indices())
    gk, gl = gf.bins()
    ok, ol = of.bins()
    assert np.array_equal(gk, ok)          # bin keys AND per-bin counts, bit-exact
    assert np.array_equal(gl, ol)          # same partition, canonical labels
    assert gf.num_clusters() == of.num_clusters()
    assert_same_particles(gf.get_particles(), of.get_particles())


def test_no_positive_weights_failure_paths():
    """particle_filter.cpp:255-259 / :307-311: nothing to draw from."""
    for rtype in (1, 2):
        _, _, _, om, gm, of, gf = make_pair(resample_type=rtype, particle_count_min=50,
                                            particle_count_max=500)
        n = 200
        set_both(of, gf, np.ones(n), np.ones(n), np.zeros(n), np.zeros(n))
        for f in (of, gf):
            f.resample_and_cluster()
        assert gf.num_particles() == of.num_particles() == (50 if rtype == 1 else 0)
        ok, pose, cov = gf.get_estimated_pose()
        ook, _, _ = of.get_estimated_pose()
        if rtype == 2:
            assert not ok and ook == 0


# ------------------------------------------------------------ K13 / K14
@pytest.mark.parametrize("cloud,n", [("tracking", 20000), ("global", 40000)])
def test_cluster_and_estimate_match_oracle(cloud, n):
    occ, res, origin, om, gm, of, gf = make_pair()
    weights_from_sensor(of, gf, occ, res, origin, n, cloud)
    of.cluster()
    gf.cluster()
    gk, gl = gf.bins()
    ok, ol = of.bins()
    assert np.array_equal(gk, ok)
    assert np.array_equal(gl, ol)
    assert gf.num_clusters() == of.num_clusters()
    g_ok, g_pose, g_cov = gf.get_estimated_pose()
    o_ok, o_pose, o_cov = of.get_estimated_pose()
    assert g_ok and o_ok == 1
    assert np.max(np.abs(g_pose[:2] - o_pose[:2])) <= 1e-4
    assert abs(math.remainder(g_pose[2] - o_pose[2], 2 * math.pi)) <= 1e-4
    assert np.max(np.abs(g_pose - o_pose)) <= 1e-9    # what it actually achieves
    assert np.max(np.abs(g_cov - o_cov)) <= 1e-9
    print(f"{cloud}: {gf.num_clusters()} clusters, {len(gk)} bins")


def test_estimate_at_one_million_particles_many_clusters():
    """K14 at scale: a global cloud has hundreds of thousands of bins and many
    clusters; the fixed-point cluster weights must pick the oracle's best cluster."""
    occ, res, origin, om, gm, of, gf = make_pair(width=1000)
    n = 1_000_000
    x, y, t = syn.uniform_cloud(occ, res, origin, n, seed=11)
    rng = np.random.default_rng(12)
    w = rng.random(n) ** 4
    w[rng.random(n) < 0.3] = 0.0
    w /= w.sum()
    set_both(of, gf, x, y, t, w)
    of.cluster()
    gf.cluster()
    assert gf.num_clusters() == of.num_clusters()
    assert np.array_equal(gf.bins()[1], of.bins()[1])
    g_ok, g_pose, g_cov = gf.get_estimated_pose()
    o_ok, o_pose, o_cov = of.get_estimated_pose()
    assert g_ok and o_ok == 1
    assert np.max(np.abs(g_pose - o_pose)) <= 1e-9
    assert np.max(np.abs(g_cov - o_cov)) <= 1e-9
    # deterministic: the same bits on a second call
    g_ok2, g_pose2, g_cov2 = gf.get_estimated_pose()
    assert np.array_equal(g_pose, g_pose2) and np.array_equal(g_cov, g_cov2)
    print(f"1M global: {gf.num_clusters()} clusters, {len(gf.bins()[0])} bins")


def test_theta_wrap_joins_clusters():
    """Bins at theta ~ +pi and ~ -pi are neighbours (space_partitioning.cpp:110-114)."""
    _, _, _, om, gm, of, gf = make_pair()
    x = np.array([1.0, 1.0, 5.0], np.float32)
    y = np.array([1.0, 1.0, 5.0], np.float32)
    t = np.array([3.1, -3.1, 0.0], np.float32)
    set_both(of, gf, x, y, t, np.full(3, 1 / 3))
    of.cluster()
    gf.cluster()
    assert gf.num_clusters() == of.num_clusters() == 2
    assert np.array_equal(gf.bins()[1], of.bins()[1])


def test_estimate_without_clustering_is_an_error():
    """particle_filter.cpp:375-379: the reference abort()s on an invalid cluster id."""
    _, _, _, om, gm, of, gf = make_pair()
    n = 10
    set_both(of, gf, np.arange(n), np.zeros(n), np.zeros(n), np.full(n, 0.1))
    gf.cluster()
    gf.handle_odometry((50.0, 0.0, 0.0), (0.0, 0.0, 0.0), syn.NODE_ALPHA)
    with pytest.raises(q._abi.QmclError) as e:
        gf.get_estimated_pose()
    assert e.value.status == 5


# ----------------------------------------------------------- whole cycles
@pytest.mark.parametrize("rtype,n", [(0, 2000), (2, 5000), (1, 3000)])
def test_multi_cycle_tracking_matches_oracle(rtype, n):
    """laser_handler.cpp:243-270 cadence: odometry, sensor, resample (or cluster), estimate."""
    occ, res, origin, om, gm, of, gf = make_pair(num_beams=30, resample_type=rtype,
                                                 particle_count_min=n if rtype != 2 else 100,
                                                 particle_count_max=n)
    pose = syn.pick_truth_pose(occ, res, origin)
    factor = np.diag(np.sqrt([0.25, 0.25, 0.01])).astype(np.float32)
    if rtype == 2:
        x, y, t = syn.gaussian_cloud(pose, n)
        set_both(of, gf, x, y, t, np.full(n, 1.0 / n))
    else:
        of.initialise_factor(pose.astype(np.float32), factor)
        gf.initialise_factor(pose.astype(np.float32), factor)
    cur = pose.copy()
    odom_old = np.zeros(3)
    for cycle in range(6):
        step = np.array([0.22 * math.cos(cur[2]), 0.22 * math.sin(cur[2]), 0.05])
        cur = cur + step
        odom_new = odom_old + np.array([0.22, 0.0, 0.05])
        scan = syn.make_scan(occ, res, origin, cur, 360, 360.0, seed=10 + cycle)
        for f in (of, gf):
            f.handle_odometry(odom_new, odom_old, syn.NODE_ALPHA)
            f.update_importance_from_observations(scan)
            if cycle % 2 == 1:
                f.resample_and_cluster()
            else:
                f.cluster()
        odom_old = odom_new
        assert gf.num_particles() == of.num_particles()
        g_ok, g_pose, g_cov = gf.get_estimated_pose()
        o_ok, o_pose, o_cov = of.get_estimated_pose()
        assert g_ok == bool(o_ok)
        assert np.max(np.abs(g_pose[:2] - o_pose[:2])) <= 1e-4, cycle
        assert abs(math.remainder(g_pose[2] - o_pose[2], 2 * math.pi)) <= 1e-4, cycle
        assert np.max(np.abs(g_cov - o_cov)) <= 1e-6
    assert np.allclose(gf.lowpass(), of.lowpass(), rtol=1e-9)


# ------------------------------------------------- deferred read-backs (sync mode)
def _run_cycles(gm, occ, res, origin, rtype, n, lazy, cycles=5, inject_at=None):
    """Whole cycles through the public mirror; returns everything observable."""
    gf = q.create_particle_filter(gm, seed=7)
    gf.set_sync_mode(lazy)
    gf.set_parameters(q.ParticleFilterParameters(
        resample_type=q.ResampleType(rtype), particle_count_min=n if rtype != 2 else 100,
        particle_count_max=n, resample_count=2, alpha_fast=0.1, alpha_slow=0.001,
        kld_epsilon=0.05, kld_z=0.95, space_partitioning_resolution_xy=0.5,
        space_partitioning_resolution_theta=RES_THETA))
    pose = syn.pick_truth_pose(occ, res, origin)
    x, y, t = syn.gaussian_cloud(pose, n)
    gf.set_particles(q.make_particles(x, y, t, np.full(n, 1.0 / n)))
    cur = pose.copy()
    odom_old = np.zeros(3)
    seen = []
    for cycle in range(cycles):
        cur = cur + np.array([0.22 * math.cos(cur[2]), 0.22 * math.sin(cur[2]), 0.05])
        odom_new = odom_old + np.array([0.22, 0.0, 0.05])
        scan = syn.make_scan(occ, res, origin, cur, 360, 360.0, seed=40 + cycle)
        gf.handle_odometry(odom_new, odom_old, syn.NODE_ALPHA)
        gf.update_importance_from_observations(scan)
        if cycle == 2:
            # a second sensor update before anything consumed the first total
            gf.update_importance_from_observations(scan)
        if inject_at == cycle:
            gf.set_lowpass(1.0, 0.7)   # must order after the pending total
        if cycle % 3 == 2:
            gf.cluster()
        else:
            gf.resample_and_cluster()
        odom_old = odom_new
        ok, p, c = gf.get_estimated_pose()
        keys, labels = gf.bins()
        seen.append((ok, p.copy(), c.copy(), gf.num_particles(), gf.num_bins(), gf.num_clusters(),
                     keys.copy(), labels.copy(), np.array(gf.lowpass()), gf.last_total_weight(),
                     gf.get_particles().copy()))
    return seen


@pytest.mark.parametrize("rtype,n,inject_at", [(0, 3000, None), (1, 4000, None), (1, 4000, 1),
                                               (2, 20000, None), (2, 20000, 3)])
def test_sync_modes_agree(rtype, n, inject_at):
    """qmcl_filter_set_sync_mode: deferring the 8-byte read-backs changes when the host
    waits, never a result — every observable is identical, bit for bit."""
    occ, res, origin = syn.make_map(400, seed=1)
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(**dict(syn.NODE_LIKELIHOOD, num_beams=30)))
    gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
    a = _run_cycles(gm, occ, res, origin, rtype, n, True, inject_at=inject_at)
    b = _run_cycles(gm, occ, res, origin, rtype, n, False, inject_at=inject_at)
    for cycle, (sa, sb) in enumerate(zip(a, b)):
        assert sa[0] == sb[0], cycle
        assert np.array_equal(sa[1], sb[1]) and np.array_equal(sa[2], sb[2]), cycle
        assert sa[3:6] == sb[3:6], (cycle, sa[3:6], sb[3:6])
        assert np.array_equal(sa[6], sb[6]) and np.array_equal(sa[7], sb[7]), cycle
        assert np.array_equal(sa[8], sb[8]) and sa[9] == sb[9], cycle
        assert sa[10].tobytes() == sb[10].tobytes(), cycle


@pytest.mark.parametrize("rtype,how", [(1, "no_free_cells"), (2, "no_free_cells"),
                                       (0, "capacity"), (2, "capacity")])
def test_failed_resample_keeps_the_particles(rtype, how):
    """resample_and_cluster is a transaction: when a step fails (injection without free
    cells -> QMCL_ERR_NO_MAP; a pose far outside every map -> the dense bin table's
    QMCL_ERR_CAPACITY) the filter still holds the set it was called with."""
    from quickmcl_b200._abi import QmclError
    w, h, res = 64, 64, 0.05
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(**dict(syn.NODE_LIKELIHOOD, num_beams=0)))
    state = np.ones((h, w), np.int8) if how == "no_free_cells" else np.zeros((h, w), np.int8)
    gm.load_field(np.full((h, w), 0.5, np.float32), state, res, (-1.6, -1.6))
    gf = q.create_particle_filter(gm, seed=5)
    p = pf_params(resample_type=rtype, particle_count_min=300, particle_count_max=600)
    gf.set_parameters(q.ParticleFilterParameters(
        resample_type=q.ResampleType(rtype), particle_count_min=300, particle_count_max=600,
        space_partitioning_resolution_xy=p["res_xy"], space_partitioning_resolution_theta=p["res_theta"]))
    rng = np.random.default_rng(7)
    n = 400
    x = rng.uniform(-1, 1, n).astype(np.float32)
    y = rng.uniform(-1, 1, n).astype(np.float32)
    t = rng.uniform(-3, 3, n).astype(np.float32)
    wts = rng.uniform(0.1, 1, n)
    if how == "capacity":
        x[17] = 3e9          # its bin index saturates: the bounding box exceeds the table
        wts[17] = 50.0       # ... and it is certain to be drawn
    wts /= wts.sum()
    gf.set_particles(q.make_particles(x, y, t, wts))
    gf.set_lowpass(1.0, 0.5)   # w_diff = 0.5: every second slot would be injected
    before = gf.get_particles().copy()
    with pytest.raises(QmclError) as e:
        gf.resample_and_cluster()
        gf.synchronize()   # a fused call reports at the next call that waits for the device
    assert e.value.status == (4 if how == "no_free_cells" else 6)
    assert gf.num_particles() == n
    assert gf.get_particles().tobytes() == before.tobytes()
    assert tuple(gf.lowpass()) == (1.0, 0.5)
    if how == "no_free_cells":   # the set is usable: cluster + estimate work on it
        gf.cluster()
        ok, pose, cov = gf.get_estimated_pose()
        assert ok and np.all(np.isfinite(pose))


def test_deferred_values_settle_on_read():
    """A getter right after the call that left its value in flight returns the value
    (the getter synchronises), also when no other call intervenes."""
    occ, res, origin, om, gm, of, gf = make_pair(resample_type=0, particle_count_min=2000,
                                                 particle_count_max=2000)
    pose = syn.pick_truth_pose(occ, res, origin)
    x, y, t = syn.gaussian_cloud(pose, 2000)
    set_both(of, gf, x, y, t, np.ones(2000))
    scan = syn.make_scan(occ, res, origin, pose, 360, 360.0)
    for f in (of, gf):
        f.update_importance_from_observations(scan)
    assert gf.last_total_weight() == pytest.approx(of.last_total_weight(), rel=1e-5)
    assert np.allclose(gf.lowpass(), of.lowpass(), rtol=1e-5)
    for f in (of, gf):
        f.cluster()
    assert gf.num_clusters() == of.num_clusters() >= 1
