"""The particle-sharded path of the library against the CPU oracle.

Shards are threads of this process on ONE GPU (quickmcl_b200.sharded.LocalComm),
so these tests also run on the single-GPU box the driver uses; the real
multi-process NCCL path is the same library code behind a different qmcl_comm
table (bench.py --gpus N, tests/test_sharded_cpu.py for the table itself).

Bars: poses of the resampled set identical to the oracle's slot by slot
(= resampling indices bit-exact), weights within 1e-11 relative of the oracle's
normalisation, estimate within 1e-4 m / 1e-4 rad.
"""
import math

import numpy as np
import pytest

import quickmcl_b200 as q
from quickmcl_b200 import sharded, synthetic as syn
from quickmcl_b200._abi import PARTICLE_DTYPE, Particle, check, ptr
from oracle import binding as ob

pytestmark = pytest.mark.gpu
RES_THETA = float(np.float32(math.pi / 18))


class Setup:
    def __init__(self, width=400, **pf):
        self.occ, self.res, self.origin = syn.make_map(width)
        self.lp = dict(syn.NODE_LIKELIHOOD, num_beams=0)
        self.om = ob.OracleMap(**self.lp)
        self.om.new_map(self.occ, self.res, self.origin, field_mode=ob.FIELD_EXACT_EDT)
        self.field, self.state = self.om.field(), self.om.state()
        self.pf = dict(resample_type=0, particle_count_min=100, particle_count_max=5000,
                       resample_count=2, alpha_fast=0.1, alpha_slow=0.001, kld_epsilon=0.05,
                       kld_z=0.95, res_xy=0.5, res_theta=RES_THETA)
        self.pf.update(pf)
        self.of = ob.OracleFilter(self.om, seed=3)
        self.of.set_parameters(**self.pf)
        self.pose = syn.pick_truth_pose(self.occ, self.res, self.origin)
        self.scan = syn.make_scan(self.occ, self.res, self.origin, self.pose, 360, 360.0)

    def make_map(self):
        gm = q.create_likelihood_map()
        gm.set_parameters(q.LikelihoodMapParameters(**self.lp))
        # the free-cell list is needed for global localisation: build the map, then
        # install the oracle's field so that weights are compared on identical tables
        gm.new_map(q.OccupancyGrid(self.occ, self.res, self.origin[0], self.origin[1]))
        return gm

    def gpu_params(self):
        p = self.pf
        return q.ParticleFilterParameters(
            resample_type=q.ResampleType(p["resample_type"]),
            particle_count_min=p["particle_count_min"], particle_count_max=p["particle_count_max"],
            resample_count=p["resample_count"], alpha_fast=p["alpha_fast"],
            alpha_slow=p["alpha_slow"], kld_epsilon=p["kld_epsilon"], kld_z=p["kld_z"],
            space_partitioning_resolution_xy=p["res_xy"],
            space_partitioning_resolution_theta=p["res_theta"])


def local_particles(f):
    import ctypes as C
    a, b, c = C.c_uint64(0), C.c_uint64(0), C.c_uint64(0)
    check(f._L.qmcl_filter_shard_range(f._h, C.byref(a), C.byref(b), C.byref(c)), "range")
    out = np.zeros(b.value, dtype=PARTICLE_DTYPE)
    if b.value:
        check(f._L.qmcl_filter_copy_particles(f._h, ptr(out, Particle), b.value), "copy")
    return a.value, out, c.value


def set_block(f, rank, size, particles):
    n = len(particles)
    first, count = sharded.shard_slice(n, rank, size)
    mine = np.ascontiguousarray(particles[first:first + count])
    check(f._L.qmcl_filter_set_particles_shard(f._h, ptr(mine, Particle), count, first, n), "set")


def assemble(results):
    parts = sorted(results, key=lambda r: r[0])
    at = 0
    for first, p, total in parts:
        assert first == at, (first, at)
        at += len(p)
    assert at == parts[0][2]
    return np.concatenate([p for _, p, _ in parts])


def rel_err(got, want):
    return np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300))


@pytest.mark.parametrize("G", [2, 3, 8])
def test_init_motion_sensor_match_oracle(G):
    s = Setup(particle_count_min=5001, particle_count_max=20_003)
    odo = ((0.3, 0.1, 0.2), (0.0, 0.0, 0.0))

    def body(f, rank):
        f.set_parameters(s.gpu_params())
        f.set_rng(3, 0)
        f.global_localization()
        a = local_particles(f)
        f.handle_odometry(odo[0], odo[1], syn.NODE_ALPHA)
        b = local_particles(f)
        f.update_importance_from_observations(s.scan)
        c = local_particles(f)
        return a, b, c, f.last_total_weight(), f.num_particles()

    res = sharded.run_local_shards(G, s.make_map, body, seed=3)
    of = s.of
    of.set_rng(3, 0)
    of.global_localization()
    o0 = of.get_particles()
    g0 = assemble([r[0] for r in res])
    for k in ("x", "y", "theta"):
        assert np.array_equal(g0[k], o0[k])
    assert np.array_equal(g0["weight"], o0["weight"])
    of.handle_odometry(odo[0], odo[1], syn.NODE_ALPHA)
    o1 = of.get_particles()
    g1 = assemble([r[1] for r in res])
    for k in ("x", "y", "theta"):
        assert np.max(np.abs(g1[k].astype(np.float64) - o1[k])) <= 1e-6
    # weights: same poses into both, then compare
    of.set_particles(g1)
    of.set_trig_mode(ob.TRIG_CR)
    of.update_importance_from_observations(s.scan)
    o2 = of.get_particles()
    g2 = assemble([r[2] for r in res])
    nz = o2["weight"] > 0
    assert np.array_equal(nz, g2["weight"] > 0)
    assert rel_err(g2["weight"][nz], o2["weight"][nz]) <= 1e-5
    totals = {r[3] for r in res}
    assert len(totals) == 1                      # every rank computed the same bits
    assert abs(totals.pop() - of.last_total_weight()) <= 1e-6 * of.last_total_weight()
    assert all(r[4] == 20_003 for r in res)


@pytest.mark.parametrize("G,n,cloud,rebalance", [(2, 5000, "tracking", True),
                                                 (3, 40_001, "global", True),
                                                 (8, 40_001, "global", False),
                                                 (4, 7, "tracking", True),
                                                 (8, 300_000, "tracking", True),
                                                 (4, 40_001, "global", "threshold")])
def test_low_variance_resample_cluster_estimate_match_oracle(G, n, cloud, rebalance):
    s = Setup(particle_count_min=n, particle_count_max=n)
    if cloud == "tracking":
        x, y, t = syn.gaussian_cloud(s.pose, n)
    else:
        x, y, t = syn.uniform_cloud(s.occ, s.res, s.origin, n)
    # weights from the oracle's own sensor model so both sides start identical
    of = s.of
    of.set_particles(ob.make_particles(x, y, t, np.full(n, 1.0 / n)))
    of.set_trig_mode(ob.TRIG_CR)
    of.update_importance_from_observations(s.scan)
    start = of.get_particles()
    # skew: all the mass of the first third removed -> unequal owner ranges
    start["weight"][: n // 3] *= 1e-9
    of.set_particles(start)

    def body(f, rank):
        f.set_parameters(s.gpu_params())
        check(f._L.qmcl_filter_set_rebalance(f._h, int(bool(rebalance))), "rebalance")
        if rebalance == "threshold":   # the skewed weights move every rank far from its share: the
            # all-to-all is skipped only under a threshold as loose as this one
            check(f._L.qmcl_filter_set_rebalance_threshold(f._h, 100.0), "rebalance threshold")
        set_block(f, rank, G, start)
        f.set_rng(3, 5)
        f.resample_and_cluster()
        parts = local_particles(f)
        est = f.get_estimated_pose()
        keys, labels = f.bins()
        return parts, est, keys, labels, f.num_clusters()

    res = sharded.run_local_shards(G, s.make_map, body, seed=3)
    of.set_rng(3, 5)
    of.resample_and_cluster()
    want = of.get_particles()
    got = assemble([r[0] for r in res])
    for k in ("x", "y", "theta"):
        assert np.array_equal(got[k], want[k]), k        # slot by slot the same source particle
    # the normalising total is a fixed tree sum here, a sequential sum in the reference
    assert rel_err(got["weight"], want["weight"]) <= 1e-11
    if rebalance is True:
        for rank, r in enumerate(res):
            first, count = sharded.shard_slice(n, rank, G)
            assert r[0][0] == first and len(r[0][1]) == count
    if rebalance == "threshold":   # skipped: the blocks are the owner ranges, not the equal shares
        assert any(len(r[0][1]) != sharded.shard_slice(n, rank, G)[1] for rank, r in enumerate(res))
    ok, o_pose, o_cov = of.get_estimated_pose()
    okeys, olabels = of.bins()
    for r in res:
        g_ok, g_pose, g_cov = r[1]
        assert g_ok and ok == 1
        assert np.max(np.abs(g_pose[:2] - o_pose[:2])) <= 1e-4
        assert abs(math.remainder(g_pose[2] - o_pose[2], 2 * math.pi)) <= 1e-4
        assert np.max(np.abs(g_cov - o_cov)) <= 1e-8
        assert np.array_equal(r[2], okeys) and np.array_equal(r[3], olabels)
        assert r[4] == of.num_clusters()
    # every rank reports the same estimate bits
    assert all(np.array_equal(r[1][1], res[0][1][1]) for r in res)


def test_multi_cycle_sharded_tracking_matches_oracle():
    G, n = 4, 6000
    s = Setup(particle_count_min=n, particle_count_max=n)
    # the 3x3 factor V*sqrt(Lambda) is passed explicitly: eigenvector signs are not canonical
    cov = np.diag(np.sqrt([0.25, 0.25, 0.01])).astype(np.float32)
    steps = [((0.25 * (k + 1), 0.05 * (k + 1), 0.1 * (k + 1)), (0.25 * k, 0.05 * k, 0.1 * k))
             for k in range(4)]

    def body(f, rank):
        f.set_parameters(s.gpu_params())
        f.set_rng(3, 0)
        f.initialise_factor(s.pose.astype(np.float32), cov)
        out = []
        for new, old in steps:
            f.handle_odometry(new, old, syn.NODE_ALPHA)
            f.update_importance_from_observations(s.scan)
            f.resample_and_cluster()
            out.append((local_particles(f), f.get_estimated_pose()))
        return out

    res = sharded.run_local_shards(G, s.make_map, body, seed=3)
    of = s.of
    of.set_trig_mode(ob.TRIG_CR)
    of.set_rng(3, 0)
    of.initialise_factor(s.pose.astype(np.float32), cov)
    for k, (new, old) in enumerate(steps):
        of.handle_odometry(new, old, syn.NODE_ALPHA)
        of.update_importance_from_observations(s.scan)
        of.resample_and_cluster()
        ok, o_pose, o_cov = of.get_estimated_pose()
        g_ok, g_pose, g_cov = res[0][k][1]
        assert g_ok and ok == 1
        assert np.max(np.abs(g_pose[:2] - o_pose[:2])) <= 1e-4, k
        assert abs(math.remainder(g_pose[2] - o_pose[2], 2 * math.pi)) <= 1e-4, k
        got = assemble([r[k][0] for r in res])
        assert len(got) == of.num_particles()


def _draw_resample_case(G, n_in, cloud, rtype, nmin, nmax, lowpass, tiny_boundary=False,
                        rebalance=True):
    s = Setup(resample_type=rtype, particle_count_min=nmin, particle_count_max=nmax)
    if cloud == "tracking":
        x, y, t = syn.gaussian_cloud(s.pose, n_in)
    else:
        x, y, t = syn.uniform_cloud(s.occ, s.res, s.origin, n_in)
    of = s.of
    of.set_particles(ob.make_particles(x, y, t, np.full(n_in, 1.0 / n_in)))
    of.set_trig_mode(ob.TRIG_CR)
    of.update_importance_from_observations(s.scan)
    start = of.get_particles()
    if tiny_boundary:
        # weights too small to move the running sum on both sides of every shard
        # boundary: equal keys across shards (first record wins, particle.cpp:77)
        for r in range(1, G):
            b = sharded.shard_slice(n_in, r, G)[0]
            start["weight"][b - 40:b + 40] = 1e-30
        start["weight"][:5] = 0.0
    of.set_particles(start)
    of.set_lowpass(*lowpass)

    def body(f, rank):
        f.set_parameters(s.gpu_params())
        check(f._L.qmcl_filter_set_rebalance(f._h, int(rebalance)), "rebalance")
        set_block(f, rank, G, start)
        f.set_lowpass(*lowpass)
        f.set_rng(3, 9)
        f.resample_and_cluster()
        parts = local_particles(f)
        est = f.get_estimated_pose()
        keys, labels = f.bins()
        return parts, est, keys, labels, f.num_clusters(), f.num_particles(), tuple(f.lowpass())

    res = sharded.run_local_shards(G, s.make_map, body, seed=3)
    of.set_rng(3, 9)
    of.resample_and_cluster()
    want = of.get_particles()
    got = assemble([r[0] for r in res])
    assert len(got) == len(want) == res[0][5]
    for k in ("x", "y", "theta"):
        assert np.array_equal(got[k], want[k]), k
    assert rel_err(got["weight"], want["weight"]) <= 1e-11
    okeys, olabels = of.bins()
    ok, o_pose, o_cov = of.get_estimated_pose()
    for r in res:
        assert np.array_equal(r[2], okeys) and np.array_equal(r[3], olabels)   # bin counts bit-exact
        assert r[4] == of.num_clusters()
        g_ok, g_pose, g_cov = r[1]
        assert bool(g_ok) == (ok == 1)
        if ok == 1:
            assert np.max(np.abs(g_pose[:2] - o_pose[:2])) <= 1e-4
            assert abs(math.remainder(g_pose[2] - o_pose[2], 2 * math.pi)) <= 1e-4
        assert np.allclose(r[6], of.lowpass())
    return len(want)


@pytest.mark.parametrize("G,n_in,cloud,nmin,lowpass,tiny", [
    (2, 5000, "tracking", 3000, (0.0, 0.0), False),
    (3, 40_001, "global", 50_000, (1.0, 0.7), False),     # 30 % injected slots
    (8, 40_001, "global", 9_999, (0.0, 0.0), True),
    (4, 30_000, "tracking", 30_000, (1.0, 2.0), True),
])
def test_adaptive_resample_sharded_matches_oracle(G, n_in, cloud, nmin, lowpass, tiny):
    n = _draw_resample_case(G, n_in, cloud, 1, nmin, nmin, lowpass, tiny)
    assert n == nmin


@pytest.mark.parametrize("G,n_in,cloud,nmax,lowpass,tiny,rebalance", [
    (2, 5000, "tracking", 5000, (0.0, 0.0), False, True),
    (3, 40_001, "global", 30_000, (0.0, 0.0), False, True),
    (8, 40_001, "global", 50_000, (1.0, 0.8), True, True),
    (4, 100_000, "tracking", 100_000, (0.0, 0.0), True, False),
])
def test_kld_resample_sharded_matches_oracle(G, n_in, cloud, nmax, lowpass, tiny, rebalance):
    n = _draw_resample_case(G, n_in, cloud, 2, 100, nmax, lowpass, tiny, rebalance)
    assert 100 < n <= nmax
    print(f"KLD sharded x{G}: {n} of {nmax} kept")


def test_sharded_draws_without_any_weight_fail_like_the_reference():
    """particle_filter.cpp:255-259, :307-311: error, return; KLD leaves an empty set."""
    s = Setup(resample_type=2)
    n = 1000
    x, y, t = syn.gaussian_cloud(s.pose, n)
    start = q.make_particles(x, y, t, np.zeros(n))

    def body(f, rank):
        f.set_parameters(s.gpu_params())
        set_block(f, rank, 2, start)
        f.set_rng(3, 0)
        f.resample_and_cluster()
        return f.num_particles(), f.get_estimated_pose()[0]

    assert sharded.run_local_shards(2, s.make_map, body, seed=3) == [(0, False), (0, False)]


@pytest.mark.parametrize("total,max_rounds", [(0.9, 0), (1.0, 0)])
def test_carry_chain_needs_no_rounds_for_single_crossings(total, max_rounds):
    """Eight shards of a ramped weight array: the running sum crosses one power of two
    well inside shards 1, 2 and 4 (at 21 %, 37 % and 62 % of the particles).  Those
    shards publish the crossing
    form of the closed-form summary and the first shard its exact carry-out, so the chain
    is settled by the summaries' exchange alone - and the result is still the sequential
    sum, slot for slot (checked against the oracle).  With weights that add up to 1 the
    sum may or may not reach the next power of two at the very last particles, which no
    prediction can know: the last shard then has no usable summary, but low-variance
    resampling does not need its carry-out (the grand total) and skips the round."""
    import ctypes as C
    G, n = 8, 400_000
    s = Setup(resample_type=0, particle_count_min=n, particle_count_max=n)
    of = s.of
    x, y, t = syn.uniform_cloud(s.occ, s.res, s.origin, n)
    w = np.random.default_rng(5).uniform(0.9, 1.1, n) * (1.0 + 2.0 * np.arange(n) / n)
    w *= total / w.sum()
    start = q.make_particles(x, y, t, w)
    of.set_particles(start.view(ob.PARTICLE_DTYPE))

    def body(f, rank):
        f.set_parameters(s.gpu_params())
        set_block(f, rank, G, start)
        f.set_rng(3, 5)
        f.resample_and_cluster()
        st = (C.c_uint64 * 3)()
        check(f._L.qmcl_filter_shard_stats(f._h, st), "stats")
        return local_particles(f), int(st[1])

    res = sharded.run_local_shards(G, s.make_map, body, seed=3)
    of.set_rng(3, 5)
    of.resample_and_cluster()
    want = of.get_particles()
    got = assemble([r[0] for r in res])
    for k in ("x", "y", "theta"):
        assert np.array_equal(got[k], want[k]), k
    rounds = {r[1] for r in res}
    assert len(rounds) == 1                        # every rank walked the same chain
    print("carry-chain rounds:", rounds)
    assert rounds.pop() <= max_rounds
