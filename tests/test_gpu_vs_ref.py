"""The CUDA path against the REFERENCE'S OWN CODE (oracle/_ref, see tests/test_oracle_vs_ref.py
and tests/test_oracle_vs_ref_filter.py): the state map and the brushfire distance map of
distance_calculator.cpp, SpacePartitioning (bins, KLD bound, clusters) of
space_partitioning.cpp run on the particle sets the GPU produced, the sensor model of
map_likelihood.cpp on the reference's own likelihood table, and the resamplers and the pose
estimate of particle_filter.cpp fed the same draws through the scripted engine.

The prebuilt oracle/_ref/libqmcl_ref.so travels to the GPU box; /root/reference itself is
never read there.  Skipped when the library is absent.
"""
import ctypes as C
import math

import numpy as np
import pytest

import quickmcl_b200 as q
from quickmcl_b200 import synthetic as syn
from oracle import binding as ob
from oracle import ref_binding as rb

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built and no reference tree")]

RES_THETA = float(np.float32(math.pi / 18))


def gpu_map(occ, res, origin, **lp):
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(**lp))
    gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
    return gm


@pytest.mark.parametrize("kind", ["rooms", "noise"])
def test_state_map_and_distance_field_against_reference_brushfire(kind):
    if kind == "rooms":
        occ, res, origin = syn.make_map(400, seed=1)
        max_dist = 2.0
    else:
        rng = np.random.default_rng(0)
        occ = rng.choice(np.array([-1, 0, 20, 34, 35, 50, 65, 66, 100], np.int8), size=(300, 257),
                         p=[.05, .6, .1, .05, .05, .05, .05, .025, .025])
        res, origin, max_dist = 0.1, (-3.0, 2.5), 1.3
    gm = gpu_map(occ, res, origin, z_hit=0.5, z_rand=0.5, num_beams=0, max_laser_distance=14.0,
                 max_obstacle_distance=max_dist, sigma_hit=0.1)
    assert np.array_equal(gm.state(), rb.state_map(occ))            # compute_state_map, bit for bit
    d_gpu, d_ref = gm.distance(), rb.distance_map(occ, res, max_dist)
    # the exact transform never exceeds the brushfire's answer, agrees on the obstacles and
    # on the cells out of reach, and differs from it in a few per cent of cells only
    assert np.all(d_gpu <= d_ref)
    assert np.array_equal(d_gpu == 0, d_ref == 0)
    frac = float((d_gpu != d_ref).mean())
    print(f"{kind}: exact EDT differs from the reference's brushfire in {frac:.4%} of cells, "
          f"max {float((d_ref - d_gpu).max()):.4f} m")
    assert frac < 0.06 and float((d_ref - d_gpu).max()) < 2.0 * res


def make_filter(n_max, rtype=2):
    occ, res, origin = syn.make_map(400, seed=1)
    gm = gpu_map(occ, res, origin, **dict(syn.NODE_LIKELIHOOD, num_beams=0))
    f = q.create_particle_filter(gm, seed=5)
    f.set_parameters(q.ParticleFilterParameters(
        resample_type=q.ResampleType(rtype), particle_count_min=100, particle_count_max=n_max,
        kld_epsilon=0.05, kld_z=0.95, space_partitioning_resolution_xy=0.5,
        space_partitioning_resolution_theta=RES_THETA))
    return occ, res, origin, gm, f


def same_partition(keys_per_particle, cluster_ref, gkeys, glabels, n_clusters):
    index = {tuple(k): i for i, k in enumerate(gkeys[:, :3])}
    mine = np.array([glabels[index[tuple(k)]] for k in keys_per_particle])
    pairs = set(zip(mine.tolist(), cluster_ref.tolist()))
    return len(pairs) == n_clusters == len({a for a, _ in pairs}) == len({b for _, b in pairs})


@pytest.mark.parametrize("cloud,n", [("tracking", 20000), ("global", 8000)])
def test_kld_stop_bins_and_clusters_against_reference_space_partitioning(cloud, n):
    occ, res, origin, gm, f = make_filter(n)
    pose = syn.pick_truth_pose(occ, res, origin)
    scan = syn.make_scan(occ, res, origin, pose, 360, 360.0)
    x, y, t = syn.gaussian_cloud(pose, n) if cloud == "tracking" else syn.uniform_cloud(occ, res, origin, n)
    f.set_particles(q.make_particles(x, y, t, np.full(n, 1.0 / n)))
    f.update_importance_from_observations(scan)
    f.set_rng(5, 9)
    f.resample_and_cluster()
    kept = f.get_particles()                       # slot m = the m-th draw of the reference's loop
    n_out = len(kept)
    print(f"KLD kept {n_out} of {n} ({cloud})")
    limits, keys, cluster, n_clusters = rb.space_partitioning(kept, 100, n, 0.05, 0.95, np.float32(0.5),
                                                              np.float32(RES_THETA))
    m = np.arange(n_out, dtype=np.uint64)
    # particle_filter.cpp:329-334: the loop runs on while m < limit() and stops at the first m >= limit()
    assert np.all(m[:-1] < limits[:-1])
    assert m[-1] >= limits[-1] or n_out == n
    gkeys, glabels = f.bins()
    ref_bins, ref_counts = np.unique(keys, axis=0, return_counts=True)
    assert np.array_equal(gkeys[:, :3], ref_bins)      # SpacePartitioning::add_point's keys
    assert np.array_equal(gkeys[:, 3], ref_counts)
    assert f.num_clusters() == n_clusters              # assign_clusters / do_cluster
    assert same_partition(keys, cluster, gkeys, glabels, n_clusters)


def test_cluster_of_a_wrapping_cloud_against_reference_space_partitioning():
    """cluster() on a cloud that straddles theta = +-pi and the origin (space_partitioning.cpp:101-128)."""
    occ, res, origin, gm, f = make_filter(6000, rtype=0)
    rng = np.random.default_rng(8)
    n = 6000
    x = np.concatenate([rng.normal(0.0, 0.5, n // 2), rng.normal(4.0, 0.3, n // 2)])
    y = np.concatenate([rng.normal(0.0, 0.5, n // 2), rng.normal(-3.0, 0.3, n // 2)])
    t = (np.concatenate([rng.normal(math.pi, 0.3, n // 2), rng.normal(0.0, 0.3, n // 2)]) + math.pi) % (2 * math.pi) - math.pi
    p = q.make_particles(x, y, t, np.full(n, 1.0 / n))
    f.set_particles(p)
    f.cluster()
    _, keys, cluster, n_clusters = rb.space_partitioning(f.get_particles(), 100, 6000, 0.05, 0.95, np.float32(0.5),
                                                         np.float32(RES_THETA))
    gkeys, glabels = f.bins()
    ref_bins, ref_counts = np.unique(keys, axis=0, return_counts=True)
    assert np.array_equal(gkeys[:, :3], ref_bins) and np.array_equal(gkeys[:, 3], ref_counts)
    assert f.num_clusters() == n_clusters
    assert same_partition(keys, cluster, gkeys, glabels, n_clusters)
    # and the weight helpers of particle.cpp on what the filter holds
    got = f.get_particles()
    n_eff = rb.lib().ref_effective_particles(got.ctypes.data_as(C.POINTER(ob.Particle)), len(got))
    assert abs(f.effective_particles() - n_eff) <= 1e-9 * n


# ------------------------------------------------------------------ map_likelihood.cpp, particle_filter.cpp
def ref_and_gpu_maps(width=400, seed=1, **lp):
    """The reference's MapLikelihood built from an occupancy grid, and a GPU map that holds the
    reference's own likelihood table and state map (identical field on both sides)."""
    occ, res, origin = syn.make_map(width, seed=seed)
    params = dict(syn.NODE_LIKELIHOOD, num_beams=0)
    params.update(lp)
    rm = rb.RefMap(occ, res, origin, **params)
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(**params))
    gm.load_field(rm.field(), rm.state(), res, origin)
    return occ, res, origin, rm, gm


@pytest.mark.parametrize("n,beams,num_beams", [(20000, 1081, 0), (5000, 360, 30)])
def test_sensor_model_against_reference_update_importance(n, beams, num_beams):
    """K7 against MapLikelihood::update_importance itself (map_likelihood.cpp:63-130).  The
    reference takes cos / sin of the heading from the host libm, the kernel rounds them
    correctly (documented deviation 3): a beam end may fall into the neighbouring
    cell for a handful of particles; everything else agrees to summation order."""
    occ, res, origin, rm, gm = ref_and_gpu_maps(num_beams=num_beams)
    pose = syn.pick_truth_pose(occ, res, origin)
    scan = syn.make_scan(occ, res, origin, pose, beams, 270.0)
    x, y, t = syn.gaussian_cloud(pose, n)
    x[: n // 4], y[: n // 4], t[: n // 4] = syn.uniform_cloud(occ, res, origin, n // 4)
    ref = ob.make_particles(x, y, t, np.ones(n))
    total_ref = rm.update_importance(scan, ref)
    host = q.make_particles(x, y, t, np.ones(n))
    total = gm.update_importance(scan, host)
    denom = np.maximum(np.abs(ref["weight"]), 1e-300)
    err = np.where(ref["weight"] == 0, np.abs(host["weight"]), np.abs(host["weight"] - ref["weight"]) / denom)
    frac = float((err > 1e-5).mean())
    print(f"GPU vs reference update_importance: {frac:.5f} of {n} particles beyond 1e-5, "
          f"median {np.median(err):.1e}, max {err.max():.3e}")
    assert frac < 0.01
    assert np.median(err) <= 1e-12
    assert np.array_equal(host["weight"] == 0, ref["weight"] == 0)      # unknown / outside: weight 0
    assert abs(total - total_ref) <= 1e-6 * total_ref


def ref_and_gpu_filters(rtype, nmin, nmax, **kld):
    occ, res, origin, rm, gm = ref_and_gpu_maps()
    kw = dict(resample_type=rtype, particle_count_min=nmin, particle_count_max=nmax, kld_epsilon=0.05,
              kld_z=0.95, res_xy=0.5, res_theta=RES_THETA)
    kw.update(kld)
    rf = rb.RefFilter(rm)
    rf.set_parameters(**kw)
    gf = q.create_particle_filter(gm, seed=33)
    gf.set_parameters(q.ParticleFilterParameters(
        resample_type=q.ResampleType(rtype), particle_count_min=nmin, particle_count_max=nmax,
        kld_epsilon=kw["kld_epsilon"], kld_z=kw["kld_z"], space_partitioning_resolution_xy=0.5,
        space_partitioning_resolution_theta=RES_THETA))
    return occ, res, origin, rm, gm, rf, gf


def reference_weighted_set(occ, res, origin, rf, n, cloud):
    """A particle set weighted and normalised by the reference's own sensor update."""
    pose = syn.pick_truth_pose(occ, res, origin)
    scan = syn.make_scan(occ, res, origin, pose, 360, 360.0)
    x, y, t = syn.gaussian_cloud(pose, n) if cloud == "tracking" else syn.uniform_cloud(occ, res, origin, n)
    rf.set_particles(ob.make_particles(x, y, t, np.full(n, 1.0 / n)))
    rf.update_importance_from_observations(scan)
    return rf.get_particles()


def assert_same_sets(gp, rp, injected=None):
    assert len(gp) == len(rp)
    if injected is None:
        injected = np.zeros(len(rp), bool)
    injected = injected[:len(rp)]
    for k in ("x", "y"):
        assert np.array_equal(gp[k], rp[k]), k
    assert np.array_equal(gp["theta"][~injected], rp["theta"][~injected])
    if injected.any():
        assert np.max(np.abs(gp["theta"][injected].astype(np.float64) - rp["theta"][injected])) <= 1e-6
    assert np.allclose(gp["weight"], rp["weight"], rtol=1e-12, atol=1e-300)


@pytest.mark.parametrize("rtype,n,nmin,nmax,cloud,lowpass", [
    (0, 20000, 20000, 20000, "tracking", (0.0, 0.0)),     # low variance
    (1, 20000, 12000, 20000, "global", (1.0, 0.7)),       # adaptive, 30 % random poses
    (2, 20000, 100, 20000, "tracking", (0.0, 0.0)),       # KLD, stops after a few hundred
    (2, 20000, 100, 20000, "global", (1.0, 0.9)),         # KLD on a global cloud with random poses
])
def test_resamplers_and_estimate_against_reference_particle_filter(rtype, n, nmin, nmax, cloud, lowpass):
    """ParticleFilter::resample_and_cluster and get_estimated_pose themselves
    (particle_filter.cpp:138-161, :177-212, :214-348) on the set the reference's own sensor update
    left, with the draws of the shared Philox contract scripted into the reference's engine."""
    occ, res, origin, rm, gm, rf, gf = ref_and_gpu_filters(rtype, nmin, nmax)
    p = reference_weighted_set(occ, res, origin, rf, n, cloud)
    if rtype == 0:
        p["weight"] /= np.cumsum(p["weight"])[-1]   # weights that add up to one: the reference's walk
        p["weight"][-1] += 1e-6                     # would otherwise run off the end and throw
        rf.set_particles(p)
    gf.set_particles(p.view(q.api.PARTICLE_DTYPE))
    seed, step = 33, 7
    rf.set_lowpass(*lowpass)
    gf.set_lowpass(*lowpass)
    gf.set_rng(seed, step)
    w_diff = max(0.0, 1.0 - lowpass[1] / lowpass[0]) if lowpass[0] else 0.0
    if rtype == 0:
        rb.script(words=rb.low_variance_script(seed, step))
        injected = None
    else:
        words, injected = rb.draw_script(seed, step, nmax if rtype == 2 else nmin, w_diff)
        rb.script(words=words)
    assert rf.resample_and_cluster()
    gf.resample_and_cluster()
    assert not rb.script_state()[2]
    kept = rf.num_particles()
    assert gf.num_particles() == kept
    print(f"resample type {rtype}: {kept} particles out, {0 if injected is None else int(injected[:kept].sum())} injected")
    assert_same_sets(gf.get_particles(), rf.get_particles(), injected)
    assert np.array_equal(gf.lowpass(), rf.lowpass())
    if injected is None or not injected[:kept].any():
        assert gf.num_clusters() == rf.num_clusters()
        ok_r, pose_r, cov_r = rf.get_estimated_pose()
        ok_g, pose_g, cov_g = gf.get_estimated_pose()
        assert bool(ok_g) == bool(ok_r) and ok_r
        assert np.max(np.abs(np.asarray(pose_g) - pose_r)) <= 1e-9
        assert np.max(np.abs(np.asarray(cov_g).reshape(3, 3) - cov_r)) <= 1e-9
