"""K1-K4 parity: likelihood-field build on the GPU against the CPU oracle.

The GPU builds the exact clipped Euclidean distance transform (north_star);
it must be bit-identical to the oracle's exact EDT, and never larger than the
reference's brushfire approximation.
"""
import numpy as np
import pytest

import quickmcl_b200 as q
from quickmcl_b200 import synthetic as syn
from oracle import binding as ob

pytestmark = pytest.mark.gpu

TEST_MAP = np.array([
    -1, -1, -1, -1, 100, 100, 100, 100, 100, 100, 100, 0, 0,
    0, 0, 100, 0, 0, 0, 0, 0, 0, 0, 0, 0], np.int8).reshape(5, 5)


def gpu_map(occ, res, origin, **lp):
    gm = q.create_likelihood_map()
    gm.set_parameters(q.LikelihoodMapParameters(**lp))
    gm.new_map(q.OccupancyGrid(occ, res, origin[0], origin[1]))
    return gm


def test_reference_golden_5x5():
    """test/quickmcl/test_distance_calculator.cpp:24-78 on the GPU path."""
    gm = gpu_map(TEST_MAP, 0.5, (0.0, 0.0), max_obstacle_distance=2.0)
    s5, s125 = np.float32(np.sqrt(0.5)), np.float32(np.sqrt(1.25))
    want = np.array([0.5, 0.5, 0.5, 0.5, 0, 0, 0, 0, 0, 0, 0, 0.5, 0.5, 0.5, 0.5, 0, 0.5, 1, 1, 1,
                     0.5, s5, s125, 1.5, 1.5], np.float32).reshape(5, 5)
    d = gm.distance()
    assert np.allclose(d, want, rtol=3e-7, atol=0)
    U, O, F = 2, 1, 0
    want_state = np.array([U, U, U, U, O, O, O, O, O, O, O, F, F, F, F, O, F, F, F, F,
                           F, F, F, F, F], np.int8).reshape(5, 5)
    assert np.array_equal(gm.state(), want_state)


@pytest.mark.parametrize("width,seed", [(400, 1), (400, 5), (1000, 2), (777, 3)])
def test_field_matches_exact_edt_oracle(width, seed):
    occ, res, origin = syn.make_map(width, seed=seed)
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=0)
    om = ob.OracleMap(**lp)
    om.new_map(occ, res, origin, field_mode=ob.FIELD_EXACT_EDT)
    gm = gpu_map(occ, res, origin, **lp)
    assert np.array_equal(gm.state(), om.state())
    assert np.array_equal(gm.distance(), om.distance())     # bit-exact
    assert np.array_equal(gm.field(), om.field())           # bit-exact
    assert np.array_equal(gm.free_cells(), om.free_cells())  # same row-major order
    assert np.array_equal(gm.get_likelihood_as_gridmap(), om.likelihood_grid())


def test_field_vs_reference_brushfire():
    occ, res, origin = syn.make_map(400, seed=1)
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=0)
    om = ob.OracleMap(**lp)
    om.new_map(occ, res, origin, field_mode=ob.FIELD_BRUSHFIRE)
    gm = gpu_map(occ, res, origin, **lp)
    d_gpu, d_ref = gm.distance(), om.distance()
    assert np.all(d_gpu <= d_ref)
    frac = float((d_gpu != d_ref).mean())
    print(f"exact EDT differs from brushfire in {frac:.4%} of cells, "
          f"max {float((d_ref - d_gpu).max()):.4f} m")
    assert frac < 0.06


def test_random_noise_map_and_non_default_params():
    rng = np.random.default_rng(0)
    occ = rng.choice(np.array([-1, 0, 20, 34, 35, 50, 65, 66, 100], np.int8), size=(300, 257),
                     p=[.05, .6, .1, .05, .05, .05, .05, .025, .025])
    lp = dict(z_hit=0.7, z_rand=0.3, num_beams=0, max_laser_distance=10.0,
              max_obstacle_distance=1.3, sigma_hit=0.2)
    om = ob.OracleMap(**lp)
    om.new_map(occ, 0.1, (-3.0, 2.5), field_mode=ob.FIELD_EXACT_EDT)
    gm = gpu_map(occ, 0.1, (-3.0, 2.5), **lp)
    assert np.array_equal(gm.state(), om.state())
    assert np.array_equal(gm.distance(), om.distance())
    assert np.array_equal(gm.field(), om.field())
    assert np.array_equal(gm.free_cells(), om.free_cells())


def test_state_at_and_random_pose():
    occ, res, origin = syn.make_map(400, seed=1)
    lp = dict(syn.NODE_LIKELIHOOD)
    om = ob.OracleMap(**lp)
    om.new_map(occ, res, origin, field_mode=ob.FIELD_EXACT_EDT)
    gm = gpu_map(occ, res, origin, **lp)
    for x, y in [(0.0, 0.0), (-9.99, -9.99), (-10.04, 0.0), (9.999, 9.999), (10.0, 0.0),
                 (-9.45, 3.0), (1e9, 0.0)]:
        assert int(gm.get_map_state_at((x, y))) == om.state_at(x, y), (x, y)
    free = om.free_cells()
    for idx in range(5):
        p = gm.generate_random_pose(seed=9, step=1, index=idx)
        cx = round((p[0] - np.float32(origin[0])) / res)
        cy = round((p[1] - np.float32(origin[1])) / res)
        assert occ[cy, cx] == 0 or (0 <= occ[cy, cx] < 35)
        assert -np.pi <= p[2] < np.pi
        assert ((free[:, 0] == cx) & (free[:, 1] == cy)).any()


@pytest.mark.parametrize("shape,slab,tail", [((333, 404), 64, 32), ((1001, 520), 128, 64),
                                             ((130, 1028), 48, 16)])
def test_fast_path_slabs_chain_exactly(monkeypatch, shape, slab, tail):
    """The byte-distance fast path (width % 4 == 0) with the load cut into many small
    row slabs: column halos across slab borders, ragged last rows, and the free-cell
    list appended slab by slab must reproduce the oracle bit for bit."""
    monkeypatch.setenv("QMCL_MAP_SLAB_ROWS", str(slab))
    monkeypatch.setenv("QMCL_MAP_TAIL_ROWS", str(tail))
    rng = np.random.default_rng(shape[0])
    occ = rng.choice(np.array([-1, 0, 20, 34, 35, 65, 66, 100, 127, -128], np.int8), size=shape,
                     p=[.05, .85, .03, .02, .01, .01, .01, .01, .005, .005])
    occ[shape[0] // 3:shape[0] // 3 + 90, 40:200] = 0     # an empty region wider than the reach
    lp = dict(syn.NODE_LIKELIHOOD, num_beams=0)
    om = ob.OracleMap(**lp)
    om.new_map(occ, 0.05, (-1.0, 2.0), field_mode=ob.FIELD_EXACT_EDT)
    gm = gpu_map(occ, 0.05, (-1.0, 2.0), **lp)
    assert np.array_equal(gm.state(), om.state())
    assert np.array_equal(gm.distance(), om.distance())
    assert np.array_equal(gm.field(), om.field())
    assert np.array_equal(gm.free_cells(), om.free_cells())
    # a second load on the same handle (status words of the first one are still there)
    occ2 = np.ascontiguousarray(occ[::-1])
    om.new_map(occ2, 0.05, (-1.0, 2.0), field_mode=ob.FIELD_EXACT_EDT)
    gm.new_map(q.OccupancyGrid(occ2, 0.05, -1.0, 2.0))
    assert np.array_equal(gm.distance(), om.distance())
    assert np.array_equal(gm.free_cells(), om.free_cells())
