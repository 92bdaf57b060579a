"""Committed golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py
from the CPU oracle): the oracle must still reproduce them bit for bit (CPU),
and the CUDA path must match them within the north-star bars (GPU)."""
import glob
import math
import os

import numpy as np
import pytest

from oracle import binding as ob
from quickmcl_b200 import synthetic as syn

HERE = os.path.dirname(os.path.abspath(__file__))
# ref_*.npz are the reference-generated fixtures of tests/test_golden_ref_fixtures.py
FIXTURES = sorted(p for p in glob.glob(os.path.join(HERE, "golden", "*.npz"))
                  if not os.path.basename(p).startswith("ref_"))
RES_THETA = float(np.float32(math.pi / 18))


def test_fixtures_exist():
    assert len(FIXTURES) >= 3


def _pf_kwargs(fx):
    rtype, nmin, nmax = (int(v) for v in fx["params"])
    return dict(resample_type=rtype, particle_count_min=nmin, particle_count_max=nmax,
                resample_count=2, alpha_fast=0.1, alpha_slow=0.001, kld_epsilon=0.05,
                kld_z=0.95, res_xy=0.5, res_theta=RES_THETA)


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_oracle_reproduces_fixture(path):
    fx = np.load(path)
    occ, res, origin = syn.make_map(int(fx["map_width"]))
    om = ob.OracleMap(**dict(syn.NODE_LIKELIHOOD, num_beams=0))
    om.new_map(occ, res, origin, field_mode=ob.FIELD_EXACT_EDT)
    crc = np.frombuffer(om.field().tobytes(), np.uint32).sum(dtype=np.uint64)
    assert crc == fx["field_crc"][0]
    of = ob.OracleFilter(om, seed=int(fx["rng"][0]))
    of.set_parameters(**_pf_kwargs(fx))
    of.set_trig_mode(ob.TRIG_CR)
    of.set_particles(fx["start"])
    of.set_rng(int(fx["rng"][0]), int(fx["rng"][1]))
    of.handle_odometry(fx["odom_new"], fx["odom_old"], fx["alpha"])
    assert of.get_particles().tobytes() == fx["moved"].tobytes()
    of.update_importance_from_observations(fx["scan"])
    assert of.get_particles().tobytes() == fx["weighted"].tobytes()
    of.set_lowpass(*fx["lowpass"])
    of.resample_and_cluster()
    assert np.array_equal(of.resample_indices(), fx["indices"])
    assert of.get_particles().tobytes() == fx["resampled"].tobytes()
    ok, est, cov = of.get_estimated_pose()
    assert ok == int(fx["estimate_ok"]) and np.array_equal(est, fx["estimate"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p) for p in FIXTURES])
def test_cuda_path_matches_fixture(path):
    import quickmcl_b200 as q
    fx = np.load(path)
    occ, res, origin = syn.make_map(int(fx["map_width"]))
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(**dict(syn.NODE_LIKELIHOOD, num_beams=0)))
    gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
    crc = np.frombuffer(gm.field().tobytes(), np.uint32).sum(dtype=np.uint64)
    assert crc == fx["field_crc"][0]                      # field bit-exact
    kw = _pf_kwargs(fx)
    gf = q.create_particle_filter(gm, seed=int(fx["rng"][0]))
    gf.set_parameters(q.ParticleFilterParameters(
        resample_type=q.ResampleType(kw["resample_type"]),
        particle_count_min=kw["particle_count_min"], particle_count_max=kw["particle_count_max"],
        kld_epsilon=kw["kld_epsilon"], kld_z=kw["kld_z"],
        space_partitioning_resolution_theta=kw["res_theta"]))
    gf.set_particles(fx["start"].view(q.api.PARTICLE_DTYPE))
    gf.set_rng(int(fx["rng"][0]), int(fx["rng"][1]))
    gf.handle_odometry(fx["odom_new"], fx["odom_old"], fx["alpha"])
    moved = gf.get_particles()
    for k in ("x", "y", "theta"):
        assert np.max(np.abs(moved[k].astype(np.float64) - fx["moved"][k])) <= 1e-6   # motion bar
    # weights: on the fixture's poses, 1e-5 relative
    gf.set_particles(fx["moved"].view(q.api.PARTICLE_DTYPE))
    gf.update_importance_from_observations(fx["scan"])
    got = gf.get_particles()["weight"]
    want = fx["weighted"]["weight"]
    nz = want > 0
    assert np.array_equal(nz, got > 0)
    assert np.max(np.abs(got[nz] - want[nz]) / want[nz]) <= 1e-5
    assert abs(gf.last_total_weight() - float(fx["total"])) <= 1e-6 * float(fx["total"])
    # resampling on identical weights: indices, set size and bin counts bit-exact
    gf.set_particles(fx["weighted"].view(q.api.PARTICLE_DTYPE))
    gf.set_lowpass(*fx["lowpass"])
    gf.resample_and_cluster()
    assert np.array_equal(gf.resample_indices(), fx["indices"])
    res_p = gf.get_particles()
    for k in ("x", "y", "theta"):
        assert np.array_equal(res_p[k], fx["resampled"][k])
    keys, labels = gf.bins()
    assert np.array_equal(keys, fx["bin_keys"]) and np.array_equal(labels, fx["bin_labels"])
    ok, est, cov = gf.get_estimated_pose()
    assert ok == bool(fx["estimate_ok"])
    assert np.max(np.abs(est[:2] - fx["estimate"][:2])) <= 1e-4
    assert abs(math.remainder(est[2] - fx["estimate"][2], 2 * math.pi)) <= 1e-4
