"""Holds the CPU oracle against the REFERENCE'S OWN CODE where that code can be built here.

oracle/_ref/libqmcl_ref.so is compiled (oracle/Makefile, target `ref`) from reference
translation units where they lie under /root/reference against stand-in headers for the
libraries they use (oracle/ref_stubs).  This file covers distance_calculator.cpp,
particle.cpp, pose_2d.cpp, space_partitioning.cpp and the header-only scaled_map.h,
low_pass_filter.h, utils.h; tests/test_oracle_vs_ref_filter.py covers the sensor model, the
motion model, the statistics and the particle filter.  These tests feed both sides inputs far beyond the reference's own
golden vectors (tests/test_oracle_golden.py) and ask for the same bits: brushfire
distance and state maps of random grids, grid truncation, weights / N_eff / running-sum
lookups, the KLD bound while bins fill up, bin keys and the cluster partition.

CPU only.  Without a prebuilt library and without a reference tree they are skipped.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import binding as ob
from oracle import ref_binding as rb

pytestmark = pytest.mark.skipif(not rb.available(), reason="oracle/_ref not built and no reference tree")

L = ob.lib()


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def random_grid(rng, h, w, p_occ):
    """Occupancy values on both sides of every threshold of distance_calculator.cpp:121-139."""
    vals = np.array([-1, 0, 1, 34, 35, 50, 65, 66, 99, 100], np.int8)
    occ = rng.choice(vals, size=(h, w)).astype(np.int8)
    free = rng.random((h, w)) > p_occ
    occ[free & (occ > 65)] = 0
    return occ


# ------------------------------------------------------------------ distance_calculator.cpp
@pytest.mark.parametrize("resolution,max_dist", [(0.05, 2.0), (0.05, 0.3), (0.1, 2.0), (0.025, 1.0), (1.0, 4.0),
                                                 (0.07, 1.3)])
def test_distance_cache_matches_reference(resolution, max_dist):
    ref = rb.distance_cache(resolution, max_dist)
    size = C.c_int(0)
    out = np.zeros(ref.size, np.float32)
    L.orc_distance_cache(resolution, max_dist, C.byref(size), _p(out, C.c_float), out.size)
    assert size.value == ref.shape[0]
    assert np.array_equal(out.reshape(ref.shape), ref)


@pytest.mark.parametrize("h,w,p_occ,resolution,max_dist,seed", [
    (1, 1, 0.5, 0.05, 2.0, 0), (1, 40, 0.1, 0.05, 2.0, 1), (33, 1, 0.1, 0.05, 2.0, 2),
    (7, 13, 0.3, 0.05, 2.0, 3), (64, 48, 0.02, 0.05, 2.0, 4), (64, 48, 0.0, 0.05, 2.0, 5),
    (120, 90, 0.01, 0.05, 2.0, 6), (200, 160, 0.003, 0.05, 2.0, 7), (90, 120, 0.05, 0.1, 1.0, 8),
    (150, 150, 0.002, 0.025, 1.0, 9), (80, 80, 0.02, 0.07, 1.3, 10),
])
def test_brushfire_and_state_map_match_reference_on_random_grids(h, w, p_occ, resolution, max_dist, seed):
    rng = np.random.default_rng(seed)
    occ = random_grid(rng, h, w, p_occ)
    om = ob.OracleMap(max_obstacle_distance=max_dist)
    om.new_map(occ, resolution, (-1.5, 2.25), field_mode=ob.FIELD_BRUSHFIRE)
    assert np.array_equal(om.state(), rb.state_map(occ))
    # the brushfire's answer depends on the order in which the priority queue releases
    # equal distances: the restatement has to reproduce that too
    assert np.array_equal(om.distance(), rb.distance_map(occ, resolution, max_dist))


def test_brushfire_matches_reference_on_the_synthetic_map():
    from quickmcl_b200 import synthetic as syn
    occ, res, origin = syn.make_map(400)
    om = ob.OracleMap(max_obstacle_distance=2.0)
    om.new_map(occ, res, origin, field_mode=ob.FIELD_BRUSHFIRE)
    assert np.array_equal(om.state(), rb.state_map(occ))
    assert np.array_equal(om.distance(), rb.distance_map(occ, res, 2.0))


# ------------------------------------------------------------------ scaled_map.h
@pytest.mark.parametrize("resolution,offset", [(0.05, (-50.0, -50.0)), (0.05, (0.0, 0.0)), (0.1, (-12.3, 4.56)),
                                               (0.025, (3.0, -7.0)), (0.07, (-1.0, -1.0))])
def test_grid_truncation_matches_reference(resolution, offset):
    rng = np.random.default_rng(11)
    pos = np.concatenate([
        rng.uniform(-60, 60, (4000, 2)),
        # cell boundaries and their float neighbours, where division and truncation decide
        (np.asarray(offset) + resolution * rng.integers(-200, 2200, (4000, 2))),
    ]).astype(np.float32)
    pos = np.concatenate([pos, np.nextafter(pos, np.float32(np.inf)), np.nextafter(pos, np.float32(-np.inf))])
    ref = rb.world_to_map2(pos, offset, resolution)
    off = np.asarray(offset, np.float32)
    res2 = np.full(2, resolution, np.float32)
    got = np.zeros_like(ref)
    for i in range(len(pos)):
        L.orc_world_to_map(_p(pos[i], C.c_float), _p(off, C.c_float), _p(res2, C.c_float), 2, _p(got[i], C.c_int))
    assert np.array_equal(got, ref)
    cells = rng.integers(-50, 5000, (2000, 2)).astype(np.intc)
    back_ref = rb.map_to_world2(cells, offset, resolution)
    back = np.zeros_like(back_ref)
    for i in range(len(cells)):
        L.orc_map_to_world(_p(cells[i], C.c_int), _p(off, C.c_float), _p(res2, C.c_float), 2, _p(back[i], C.c_float))
    assert np.array_equal(back, back_ref)
    size = np.array([400, 300], np.intc)
    for cell in [(0, 0), (-1, 0), (0, -1), (399, 299), (400, 299), (399, 300), (200, 150), (-5, -5)]:
        c = np.asarray(cell, np.intc)
        assert bool(L.orc_is_in_map(_p(c, C.c_int), _p(size, C.c_int), 2)) == rb.is_in_map2(cell, size)


# ------------------------------------------------------------------ particle.cpp
def weighted_particles(rng, n, zeros=0.0, spread=1.0):
    w = rng.random(n) ** 4 * spread
    w[rng.random(n) < zeros] = 0.0
    return ob.make_particles(rng.normal(size=n), rng.normal(size=n), rng.uniform(-3, 3, n), w)


@pytest.mark.parametrize("n,zeros,spread,seed", [(1, 0.0, 1.0, 0), (2, 0.0, 1e-300, 1), (1000, 0.0, 1.0, 2),
                                                 (1000, 0.3, 1e-12, 3), (50_000, 0.05, 1e6, 4)])
def test_weight_helpers_match_reference(n, zeros, spread, seed):
    rng = np.random.default_rng(seed)
    p = weighted_particles(rng, n, zeros, spread)
    assert L.orc_effective_particles(_p(p, ob.Particle), n) == rb.lib().ref_effective_particles(_p(p, ob.Particle), n)
    a, b = p.copy(), p.copy()
    L.orc_normalise_weights(_p(a, ob.Particle), n)
    rb.lib().ref_normalise_weights(_p(b, ob.Particle), n)
    assert np.array_equal(a["weight"], b["weight"], equal_nan=True)
    a, b = p.copy(), p.copy()
    L.orc_normalise_weights_total(_p(a, ob.Particle), n, 0.37)
    rb.lib().ref_normalise_weights_total(_p(b, ob.Particle), n, 0.37)
    assert np.array_equal(a["weight"], b["weight"])
    a, b = p.copy(), p.copy()
    L.orc_equalise_weights(_p(a, ob.Particle), n)
    rb.lib().ref_equalise_weights(_p(b, ob.Particle), n)
    assert np.array_equal(a["weight"], b["weight"])


@pytest.mark.parametrize("n,zeros,seed", [(1, 0.0, 0), (5, 0.4, 1), (3000, 0.0, 2), (3000, 0.5, 3), (40_000, 0.1, 4)])
def test_running_sum_lookup_matches_reference(n, zeros, seed):
    rng = np.random.default_rng(seed)
    p = weighted_particles(rng, n, zeros)
    if not (p["weight"] > 0).any():
        p["weight"][0] = 0.5
    total_ref, _ = rb.running_sum_lookup(p, np.zeros(1))
    keys = np.concatenate([[0.0], np.cumsum(p["weight"])[:-1]])  # near the interval starts
    q = np.concatenate([rng.uniform(0, total_ref, 5000), keys[:2000], np.nextafter(keys[:2000], np.inf),
                        np.nextafter(keys[1:2001], -np.inf), [total_ref, total_ref * 2]])
    total_ref, idx_ref = rb.running_sum_lookup(p, q)
    idx = np.zeros(len(q), np.uint64)
    total = L.orc_running_sum_lookup(_p(p, ob.Particle), n, _p(q, C.c_double), len(q), _p(idx, C.c_uint64))
    assert total == total_ref
    assert np.array_equal(idx, idx_ref)


# ------------------------------------------------------------------ low_pass_filter.h, utils.h, pose_2d.h
def test_lowpass_angles_decomposed_match_reference():
    rng = np.random.default_rng(5)
    R = rb.lib()
    for as_float in (0, 1):
        for alpha in (0.001, 0.1, 0.5, 1.0):
            xs = np.concatenate([[0.0, 0.0], rng.random(200) * 1e-3, [0.0], rng.random(50)])
            assert (L.orc_lowpass_run(alpha, _p(xs, C.c_double), len(xs), as_float)
                    == R.ref_lowpass_run(alpha, _p(xs, C.c_double), len(xs), as_float))
    for a, b in zip(np.concatenate([rng.uniform(-50, 50, 3000), [np.pi, -np.pi, 3 * np.pi, 0.0]]),
                    np.concatenate([rng.uniform(-50, 50, 3000), [-np.pi, np.pi, 0.0, np.pi]])):
        assert L.orc_normalise_angle_f64(a) == R.ref_normalise_angle_f64(a)
        assert L.orc_normalise_angle_f32(a) == R.ref_normalise_angle_f32(a)
        assert L.orc_angle_delta_f64(a, b) == R.ref_angle_delta_f64(a, b)
        assert L.orc_angle_delta_f32(a, b) == R.ref_angle_delta_f32(a, b)
    for _ in range(500):
        pose = rng.uniform(-10, 10, 3).astype(np.float32)
        d0, d1 = np.zeros(4), np.zeros(4)
        L.orc_to_decomposed(_p(pose, C.c_float), _p(d0, C.c_double))
        R.ref_to_decomposed(_p(pose, C.c_float), _p(d1, C.c_double))
        assert np.array_equal(d0, d1)


# ------------------------------------------------------------------ space_partitioning.cpp
def clouds():
    rng = np.random.default_rng(6)
    out = {}
    n = 4000
    # tracking cloud around a pose on the far side of theta's wrap
    out["tracking_wrap"] = (rng.normal(3.2, 0.3, n), rng.normal(-1.7, 0.3, n),
                            (rng.normal(np.pi, 0.25, n) + np.pi) % (2 * np.pi) - np.pi)
    # straddles the origin: truncation toward zero makes the bins at 0 twice as wide
    out["origin"] = (rng.normal(0, 0.6, n), rng.normal(0, 0.6, n), rng.normal(0, 0.3, n))
    # several separate blobs plus scattered outliers
    cx = rng.choice([-8.0, 0.5, 9.0], n)
    out["blobs"] = (np.where(rng.random(n) < 0.02, rng.uniform(-20, 20, n), cx + rng.normal(0, 0.4, n)),
                    np.where(rng.random(n) < 0.02, rng.uniform(-20, 20, n), rng.normal(2, 0.4, n)),
                    rng.uniform(-np.pi, np.pi, n))
    # sparse global cloud: mostly singleton clusters
    out["global"] = (rng.uniform(-40, 40, n), rng.uniform(-40, 40, n), rng.uniform(-np.pi, np.pi, n))
    return out


@pytest.mark.parametrize("name", ["tracking_wrap", "origin", "blobs", "global"])
@pytest.mark.parametrize("res_xy,res_t", [(0.5, np.pi / 18), (0.25, np.pi / 9)])
def test_kld_bound_bins_and_clusters_match_reference(name, res_xy, res_t):
    x, y, t = clouds()[name]
    n = len(x)
    p = ob.make_particles(x, y, t, np.full(n, 1.0 / n))
    eps, z, nmin, nmax = 0.05, 0.95, 100, 5000
    limits, keys, cluster, n_clusters = rb.space_partitioning(p, nmin, nmax, eps, z, np.float32(res_xy),
                                                              np.float32(res_t))
    # the KLD bound after every add_point (particle_filter.cpp:329-334 reads exactly this)
    seen, k_after = set(), np.zeros(n, np.int64)
    for i, key in enumerate(map(tuple, keys)):
        seen.add(key)
        k_after[i] = len(seen)
    for k in np.unique(k_after):
        want = limits[np.searchsorted(k_after, k)]
        assert L.orc_kld_limit(int(k), eps, z, nmin, nmax) == want, (k, want)
    # bins and clusters through the oracle's filter
    om = ob.OracleMap()
    om.new_map(np.zeros((8, 8), np.int8), 0.05, (0.0, 0.0), field_mode=ob.FIELD_BRUSHFIRE)
    of = ob.OracleFilter(om, seed=1)
    of.set_parameters(resample_type=2, particle_count_min=nmin, particle_count_max=nmax, kld_epsilon=eps, kld_z=z,
                      res_xy=res_xy, res_theta=float(np.float32(res_t)))
    of.set_particles(p)
    of.cluster()
    okeys, olabels = of.bins()
    ref_bins, ref_counts = np.unique(keys, axis=0, return_counts=True)   # lexicographic, like the oracle's list
    assert np.array_equal(okeys[:, :3], ref_bins)
    assert np.array_equal(okeys[:, 3], ref_counts)
    assert of.num_clusters() == n_clusters
    # same partition: the oracle's label of a particle's bin <-> the reference's cluster id
    index = {tuple(k): i for i, k in enumerate(okeys[:, :3])}
    mine = np.array([olabels[index[tuple(k)]] for k in keys])
    pairs = set(zip(mine.tolist(), cluster.tolist()))
    assert len(pairs) == n_clusters == len({a for a, _ in pairs}) == len({b for _, b in pairs})


# ------------------------------------------------------------------ the reference's own unit tests
def test_reference_own_gtests_pass():
    """The reference's test suite (/root/reference/test: 21 TESTs, the typed ones for float and double)
    compiled against the same stand-in headers and reference sources as oracle/_ref, plus a stand-in
    gtest / gmock (oracle/Makefile, target ref_gtests): the build of the reference the oracle is held
    against passes the reference's own tests."""
    import os
    import subprocess
    if not os.path.isdir(os.path.join(rb.REFERENCE_ROOT, "test", "quickmcl")):
        pytest.skip("no reference tree")
    here = os.path.dirname(rb.LIB_PATH)
    out = subprocess.run(["make", "-s", "-C", os.path.dirname(here), "ref_gtests", f"REF={rb.REFERENCE_ROOT}"],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    text = out.stdout.decode()
    assert out.returncode == 0, text[-3000:]
    assert "28 tests ran, 0 failed." in text
    assert text.count("[       OK ]") == 28 and "FAILED" not in text


def test_gtest_stand_in_reports_failures(tmp_path):
    """Negative control for the stand-in gtest: wrong expectations must fail, 4-ULP ones must pass."""
    import os
    import subprocess
    stubs = os.path.join(os.path.dirname(os.path.dirname(rb.LIB_PATH)), "ref_stubs")
    src = tmp_path / "t.cpp"
    src.write_text(r'''
#include <gmock/gmock.h>
#include <gtest/gtest.h>
#include <cmath>
#include <vector>
struct P { double w; };
TEST(Control, passes) {
  EXPECT_FLOAT_EQ(1.0f, std::nextafter(std::nextafter(1.0f, 2.0f), 2.0f));
  EXPECT_DOUBLE_EQ(0.1 + 0.2, 0.3);
  EXPECT_EQ(3, 3u) << "never shown";
  std::vector<P> v{{0.5}, {0.5}};
  EXPECT_THAT(v, testing::Each(testing::Field(&P::w, testing::DoubleEq(0.5))));
}
TEST(Control, float_off_by_more_than_4_ulp) { EXPECT_FLOAT_EQ(1.0f, 1.000001f); }
TEST(Control, nan_is_never_equal) { EXPECT_DOUBLE_EQ(std::nan(""), std::nan("")); }
TEST(Control, each_field) {
  std::vector<P> v{{0.5}, {0.25}};
  EXPECT_THAT(v, testing::Each(testing::Field(&P::w, testing::DoubleEq(0.5))));
}
TEST(Control, plain) { EXPECT_EQ(1, 2) << "shown"; EXPECT_TRUE(false); EXPECT_FALSE(true); }
int main(int argc, char **argv) { testing::InitGoogleTest(&argc, argv); return RUN_ALL_TESTS(); }
''')
    exe = tmp_path / "t"
    subprocess.check_call(["g++", "-std=c++17", "-w", "-I", stubs, "-o", str(exe), str(src)])
    out = subprocess.run([str(exe)], stdout=subprocess.PIPE)
    text = out.stdout.decode()
    assert out.returncode == 1
    assert "[       OK ] Control.passes" in text
    assert text.count("[   FAILED ]") == 4 and "5 tests ran, 4 failed." in text
    assert "shown" in text and "never shown" not in text
