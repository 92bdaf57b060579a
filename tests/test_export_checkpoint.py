"""the steps right after the path.

* particle export in the node's published form (publishing.cpp:45-67, 114-154)
  against the oracle's restatement of to_marker / publish_cloud;
* the pose checkpoint (pose_restore.cpp:37-100): store -> six numbers ->
  restore, with the reference's defaults for missing / NaN entries;
* parameters changed between cycles (parameter_manager.cpp:102-189 ends in
  set_parameters on the live filter): the next clustering must use the new bin
  resolutions without any reset.
"""
import math

import numpy as np
import pytest

from oracle import binding as ob
from quickmcl_b200 import synthetic as syn
from quickmcl_b200.cycle import PoseCheckpoint


def oracle_filter(n=3000, seed=3):
    occ, res, origin = syn.make_map(400)
    om = ob.OracleMap(**dict(syn.NODE_LIKELIHOOD, num_beams=0))
    om.new_map(occ, res, origin, field_mode=ob.FIELD_EXACT_EDT)
    of = ob.OracleFilter(om, seed=seed)
    of.set_parameters(resample_type=0, particle_count_min=n, particle_count_max=n,
                      kld_epsilon=0.05, kld_z=0.95)
    return om, of


# ------------------------------------------------------------------ CPU: host logic
def test_checkpoint_store_restore_host_logic():
    om, of = oracle_filter()
    ck = PoseCheckpoint(of)
    # no estimate yet (no clustering has run): nothing to store
    of.initialise((1.0, -2.0, 0.5), np.diag([0.04, 0.09, 0.01]))
    of.cluster()
    v = ck.store()
    assert v is not None and len(v) == 6
    ok, pose, cov = of.get_estimated_pose()
    assert ok and v[:3] == [pose[0], pose[1], pose[2]]
    assert v[3:] == [cov[0, 0], cov[1, 1], cov[2, 2]]
    # restore with the stored vector: initialise(pose, diag) with the covariance cast to float
    of.set_rng(3, 5)
    pose_r, cov_r = ck.restore(v)
    assert np.array_equal(pose_r, v[:3]) and np.array_equal(np.diag(cov_r), v[3:])
    assert cov_r[0, 1] == 0 and cov_r[1, 2] == 0
    got = of.get_particles()
    om2, of2 = oracle_filter()
    of2.set_rng(3, 5)
    of2.initialise(np.asarray(v[:3], np.float32), np.diag(v[3:]).astype(np.float32))
    want = of2.get_particles()
    for k in ("x", "y", "theta", "weight"):
        assert np.array_equal(got[k], want[k])


def test_checkpoint_defaults_for_missing_and_nan_entries():
    om, of = oracle_filter(n=500)
    ck = PoseCheckpoint(of)
    pose, cov = ck.restore(None)               # no parameter: pose 0, identity (pose_restore.cpp:59-60)
    assert np.array_equal(pose, [0, 0, 0]) and np.array_equal(cov, np.eye(3))
    pose, cov = ck.restore([2.0, float("nan"), 0.25])      # short vector + NaN
    assert np.array_equal(pose, [2.0, 0.0, 0.25])
    assert np.array_equal(np.diag(cov), [0.25, 0.25, 0.1 * 0.1])
    assert sum("missing" in w for w in ck.warnings) == 3
    assert sum("NaN" in w for w in ck.warnings) == 1
    pose, cov = ck.restore([1, 2, 3, float("nan"), 0.5, float("nan")])
    assert np.array_equal(np.diag(cov), [0.25, 0.5, 0.1 * 0.1])
    assert of.num_particles() == 500


def test_marker_oracle_matches_scalar_restatement():
    rng = np.random.default_rng(4)
    n = 200
    ps = ob.make_particles(rng.normal(size=n), rng.normal(size=n),
                           rng.uniform(-math.pi, math.pi, n), rng.random(n))
    out, mw = ob.export_markers(ps)
    assert mw == ps["weight"].max()
    for i in (0, 17, 199):
        half = np.float32(ps["theta"][i]) / np.float32(2.0)
        assert out[i, 2] == np.sin(half) and out[i, 3] == np.cos(half)
        assert out[i, 4] == np.float32(ps["weight"][i] / mw)
    assert out[:, 4].max() == 1.0
    empty, mw0 = ob.export_markers(ps[:0])
    assert empty.shape == (0, 5) and mw0 == 0.0


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_export_markers_matches_oracle():
    import quickmcl_b200 as q
    occ, res, origin = syn.make_map(400)
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(**dict(syn.NODE_LIKELIHOOD, num_beams=0)))
    gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
    gf = q.create_particle_filter(gm, seed=3)
    rng = np.random.default_rng(9)
    for n in (1, 33, 100_003):
        w = rng.lognormal(0.0, 2.0, n)
        w[rng.random(n) < 0.2] = 0.0
        if n > 1:
            w[0] = 0.0
        ps = q.make_particles(rng.normal(size=n) * 3, rng.normal(size=n) * 3,
                              rng.uniform(-math.pi, math.pi, n), w)
        gf.set_particles(ps)
        got, mw = gf.export_markers()
        want, mw_ref = ob.export_markers(ps)
        assert mw == mw_ref
        assert np.array_equal(got[:, :2], want[:, :2])
        # sinf/cosf: CUDA's are within 2 ulp of libm's (absolute 2.4e-7 on [-1, 1])
        assert np.max(np.abs(got[:, 2:4] - want[:, 2:4])) <= 2.5e-7
        if mw > 0:
            assert np.array_equal(got[:, 4], want[:, 4])   # double division, one cast: exact
        else:
            assert np.all(np.isnan(got[:, 4]))
        # a capped copy still normalises by the maximum of the whole set
        part, mw2 = gf.export_markers(capacity=max(1, n // 2))
        assert mw2 == mw and np.array_equal(part, got[:len(part)], equal_nan=True)


@pytest.mark.gpu
def test_checkpoint_round_trip_on_the_cuda_filter():
    import quickmcl_b200 as q
    occ, res, origin = syn.make_map(400)
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=0)
    om, of = oracle_filter(n=4000)
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(**lp))
    gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
    gf = q.create_particle_filter(gm, seed=3)
    gf.set_parameters(q.ParticleFilterParameters(
        resample_type=q.ResampleType(0), particle_count_min=4000, particle_count_max=4000,
        kld_epsilon=0.05, kld_z=0.95))
    ck_g, ck_o = PoseCheckpoint(gf), PoseCheckpoint(of)
    for f in (gf, of):
        f.set_rng(3, 1)
        f.initialise(np.float32([0.7, -1.1, 2.0]), np.diag([0.09, 0.04, 0.02]).astype(np.float32))
        f.cluster()
    vg, vo = ck_g.store(), ck_o.store()
    assert np.allclose(vg[:3], vo[:3], atol=1e-4) and np.allclose(vg[3:], vo[3:], rtol=1e-4)
    for f, ck in ((gf, ck_g), (of, ck_o)):
        f.set_rng(3, 2)
        ck.restore(vo)                      # the same stored vector into both
    gp, op = gf.get_particles(), of.get_particles()
    assert len(gp) == len(op) == 4000
    for k in ("x", "y", "theta"):
        assert np.max(np.abs(gp[k].astype(np.float64) - op[k])) <= 2e-6, k
    assert np.array_equal(gp["weight"], op["weight"])


@pytest.mark.gpu
def test_parameters_changed_between_cycles_take_effect_without_reset():
    import quickmcl_b200 as q
    occ, res, origin = syn.make_map(400)
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=0)
    om, of = oracle_filter(n=6000)
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(**lp))
    gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
    gf = q.create_particle_filter(gm, seed=3)
    pose = syn.pick_truth_pose(occ, res, origin)
    x, y, t = syn.gaussian_cloud(pose, 6000)
    # two far-apart blobs so that the bin size decides how many clusters there are
    x[3000:] += 3.0
    w = np.full(6000, 1.0 / 6000)
    base = dict(resample_type=0, particle_count_min=6000, particle_count_max=6000,
                kld_epsilon=0.05, kld_z=0.95)
    counts = []
    for res_xy, res_t in ((0.5, math.pi / 18), (0.1, math.pi / 36), (4.0, math.pi / 4), (0.5, math.pi / 18)):
        res_t = float(np.float32(res_t))
        of.set_parameters(res_xy=res_xy, res_theta=res_t, **base)
        gf.set_parameters(q.ParticleFilterParameters(
            resample_type=q.ResampleType(0), particle_count_min=6000, particle_count_max=6000,
            kld_epsilon=0.05, kld_z=0.95, space_partitioning_resolution_xy=res_xy,
            space_partitioning_resolution_theta=res_t))
        of.set_particles(ob.make_particles(x, y, t, w))
        gf.set_particles(q.make_particles(x, y, t, w))
        of.cluster()
        gf.cluster()
        assert gf.num_clusters() == of.num_clusters()
        counts.append(gf.num_clusters())
        ok_g, pg, cg = gf.get_estimated_pose()
        ok_o, po, co = of.get_estimated_pose()
        assert ok_g == ok_o
        assert np.allclose(pg, po, atol=1e-4) and np.allclose(cg, co, atol=1e-4)
    assert counts[0] == counts[3] and len(set(counts)) > 1
