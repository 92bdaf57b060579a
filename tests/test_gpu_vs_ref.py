"""The CUDA path against the REFERENCE'S OWN CODE (oracle/_ref, see tests/test_oracle_vs_ref.py)
where that code can be built: the state map and the brushfire distance map of
distance_calculator.cpp, and SpacePartitioning (bins, KLD bound, clusters) of
space_partitioning.cpp run on the particle sets the GPU produced.

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
