"""Pins the CPU oracle against every golden vector / known-answer test the
reference's own gtest suite holds for this path.

Each test cites the reference test it restates (paths relative to
/root/reference).  gtest's EXPECT_FLOAT_EQ / EXPECT_DOUBLE_EQ mean "within
4 ULPs"; `ulps32` / `ulps64` reproduce that.
"""
import ctypes as C
import math

import numpy as np
import pytest

from oracle import binding as ob

L = ob.lib()


def ulps32(a, b):
    a, b = np.float32(a), np.float32(b)
    if a == b:
        return 0

    def key(v):
        i = int(np.array(v, np.float32).view(np.int32))
        return i if i >= 0 else -(i & 0x7FFFFFFF)
    return abs(key(a) - key(b))


def ulps64(a, b):
    a, b = np.float64(a), np.float64(b)
    if a == b:
        return 0

    def key(v):
        i = int(np.array(v, np.float64).view(np.int64))
        return i if i >= 0 else -(i & 0x7FFFFFFFFFFFFFFF)
    return abs(key(a) - key(b))


def float_eq(expected, actual):
    return ulps32(expected, actual) <= 4


def double_eq(expected, actual):
    return ulps64(expected, actual) <= 4


# --------------------------------------------------------------------------
# test/quickmcl/test_distance_calculator.cpp
# --------------------------------------------------------------------------
TEST_MAP = np.array([
    -1, -1, -1, -1, 100, 100, 100, 100, 100, 100, 100, 0, 0,
    0, 0, 100, 0, 0, 0, 0, 0, 0, 0, 0, 0], np.int8).reshape(5, 5)  # :24-35


def test_create_distance_cache():  # :7-22
    n = C.c_int(0)
    out = np.zeros(64, np.float32)
    L.orc_distance_cache(1.0, 4.0, C.byref(n), out.ctypes.data_as(C.POINTER(C.c_float)), 64)
    assert n.value == 6
    cache = out[:36].reshape(6, 6)
    assert float_eq(0.0, cache[0, 0])
    assert float_eq(1.0, cache[0, 1]) and float_eq(1.0, cache[1, 0])
    assert float_eq(math.sqrt(2), cache[1, 1])
    assert float_eq(math.sqrt(5), cache[1, 2]) and float_eq(math.sqrt(5), cache[2, 1])


EXPECTED_DISTANCE = np.array([
    0.5, 0.5, 0.5, 0.5, 0, 0, 0, 0, 0, 0, 0, 0.5, 0.5, 0.5,
    0.5, 0, 0.5, 1, 1, 1, 0.5, math.sqrt(0.5 * 0.5 + 0.5 * 0.5),
    math.sqrt(0.5 * 0.5 + 1), 1.5, 1.5]).reshape(5, 5)  # :43-46


@pytest.mark.parametrize("mode", [ob.FIELD_BRUSHFIRE, ob.FIELD_EXACT_EDT])
def test_compute_distance_map(mode):  # :37-55 (brushfire and exact EDT agree on it)
    m = ob.OracleMap(max_obstacle_distance=2.0)
    m.new_map(TEST_MAP, 0.5, (0.0, 0.0), field_mode=mode)
    d = m.distance()
    for y in range(5):
        for x in range(5):
            assert float_eq(EXPECTED_DISTANCE[y, x], d[y, x]), (x, y, d[y, x])


def test_compute_state_map():  # :57-78
    U, O, F = 2, 1, 0
    expected = np.array([U, U, U, U, O, O, O, O, O, O, O, F, F, F, F, O, F, F, F, F,
                         F, F, F, F, F], np.int8).reshape(5, 5)
    m = ob.OracleMap()
    m.new_map(TEST_MAP, 0.5, (0.0, 0.0))
    assert np.array_equal(expected, m.state())


# --------------------------------------------------------------------------
# test/quickmcl/test_scaled_map.cpp
# --------------------------------------------------------------------------
def _w2m(pos, off, res):
    pos = np.asarray(pos, np.float32)
    res = np.asarray(res, np.float32)
    out = np.zeros(len(pos), np.int32)
    offp = None
    if off is not None:
        off = np.asarray(off, np.float32)
        offp = off.ctypes.data_as(C.POINTER(C.c_float))
    L.orc_world_to_map(pos.ctypes.data_as(C.POINTER(C.c_float)), offp,
                       res.ctypes.data_as(C.POINTER(C.c_float)), len(pos),
                       out.ctypes.data_as(C.POINTER(C.c_int)))
    return tuple(int(v) for v in out)


def _m2w(cell, off, res):
    cell = np.asarray(cell, np.int32)
    res = np.asarray(res, np.float32)
    out = np.zeros(len(cell), np.float32)
    offp = None
    if off is not None:
        off = np.asarray(off, np.float32)
        offp = off.ctypes.data_as(C.POINTER(C.c_float))
    L.orc_map_to_world(cell.ctypes.data_as(C.POINTER(C.c_int)), offp,
                       res.ctypes.data_as(C.POINTER(C.c_float)), len(cell),
                       out.ctypes.data_as(C.POINTER(C.c_float)))
    return tuple(float(v) for v in out)


def _in_map(cell, size):
    cell = np.asarray(cell, np.int32)
    size = np.asarray(size, np.int32)
    return bool(L.orc_is_in_map(cell.ctypes.data_as(C.POINTER(C.c_int)),
                                size.ctypes.data_as(C.POINTER(C.c_int)), len(cell)))


def test_scaled_infinite_map_transformations():  # :21-37
    assert _w2m((2, 3), None, (2.5, 0.5)) == (0, 6)
    assert _w2m((2, 3, 4), None, (2.5, 0.5, 1.5)) == (0, 6, 2)
    assert _m2w((5, 1), None, (2.5, 0.5)) == (12.5, 0.5)
    assert _m2w((5, 1, 6), None, (2.5, 0.5, 1.5)) == (12.5, 0.5, 9.0)


def test_scaled_map_transformations():  # :39-63 (truncation toward zero, negatives)
    assert _w2m((2, 3), (10, 20), (2.5, 0.5)) == (-3, -34)
    assert _w2m((2, 3, 4), (10, 20, 30), (2.5, 0.5, 1.5)) == (-3, -34, -17)
    assert _m2w((5, 1), (10, 20), (2.5, 0.5)) == (22.5, 20.5)
    assert _m2w((5, 1, 6), (10, 20, 30), (2.5, 0.5, 1.5)) == (22.5, 20.5, 39.0)


def test_scaled_map_inside_map():  # :65-95
    assert _in_map((0, 0), (4, 8)) and _in_map((0, 0, 0), (4, 8, 10))
    assert not _in_map((-1, 0), (4, 8)) and not _in_map((0, -1, 0), (4, 8, 10))
    assert _in_map((3, 7), (4, 8)) and _in_map((3, 7, 9), (4, 8, 10))
    assert not _in_map((3, 8), (4, 8)) and not _in_map((3, 8, 10), (4, 8, 10))
    assert not _in_map((4, 8), (4, 8)) and not _in_map((4, 8, 10), (4, 8, 10))
    assert not _in_map((3, 9), (4, 8)) and not _in_map((3, 7, 11), (4, 8, 10))


# --------------------------------------------------------------------------
# test/quickmcl/test_particle.cpp
# --------------------------------------------------------------------------
def _P(ws):
    n = len(ws)
    return ob.make_particles(np.zeros(n), np.zeros(n), np.zeros(n), ws)


def test_effective_particles():  # :12-27
    p = _P([1, 1, 1])
    L.orc_equalise_weights(p.ctypes.data_as(C.POINTER(ob.Particle)), 3)
    assert double_eq(3, L.orc_effective_particles(p.ctypes.data_as(C.POINTER(ob.Particle)), 3))
    p2 = _P([1, 0, 0])
    assert double_eq(1, L.orc_effective_particles(p2.ctypes.data_as(C.POINTER(ob.Particle)), 3))


def test_normalise_particle_weights():  # :29-52
    p = _P([0.2, 0.3, 0.5, 0.0, 2.0])
    L.orc_normalise_weights(p.ctypes.data_as(C.POINTER(ob.Particle)), 5)
    for got, want in zip(p["weight"], [0.2 / 3, 0.3 / 3, 0.5 / 3, 0.0, 2.0 / 3]):
        assert double_eq(want, got)


def test_equalise_particle_weights():  # :54-66
    p = _P([0.2, 0.3, 0.4, 0.0, 231.0])
    L.orc_equalise_weights(p.ctypes.data_as(C.POINTER(ob.Particle)), 5)
    assert all(double_eq(1.0 / 5.0, w) for w in p["weight"])


def test_running_sum():  # :68-94
    p = _P([0.2, 0.3, 0.4, 0.0, 0.1])
    q = np.array([0.0, 0.1, 0.199, 0.2, 0.499, 0.5, 0.899, 0.9, 1.0])
    idx = np.zeros(len(q), np.uint64)
    total = L.orc_running_sum_lookup(p.ctypes.data_as(C.POINTER(ob.Particle)), 5,
                                     q.ctypes.data_as(C.POINTER(C.c_double)), len(q),
                                     idx.ctypes.data_as(C.POINTER(C.c_uint64)))
    assert list(idx) == [0, 0, 0, 1, 1, 2, 2, 4, 4]
    assert double_eq(1.0, total)


# --------------------------------------------------------------------------
# test/quickmcl/test_statistics.cpp
# --------------------------------------------------------------------------
def _stats(p, finalise):
    out = np.zeros(15)
    L.orc_statistics(p.ctypes.data_as(C.POINTER(ob.Particle)), len(p), int(finalise),
                     out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def _merge(a, b, finalise):
    out = np.zeros(15)
    L.orc_statistics_merge(a.ctypes.data_as(C.POINTER(C.c_double)),
                           b.ctypes.data_as(C.POINTER(C.c_double)), int(finalise),
                           out.ctypes.data_as(C.POINTER(C.c_double)))
    return out


def test_statistics_functionality():  # :29-99
    pa = ob.make_particles([1, 0, 0, 2, 0], [0, 1, 0, 0, 2], [0, 0, 1, 0.5, 0.5],
                           [0.2 / 3, 0.3 / 3, 0.5 / 3, 0.0, 2.0 / 3])
    pb = ob.make_particles([0.5, 3], [2.0, 0.5], [-1.5, 0], [0.2, 0.5])
    sa = _stats(pa, True)
    assert sa[1] == 5
    both = np.concatenate([pa, pb])
    sc = _stats(both, True)                                  # stats_c: add(a); add(b)
    sd = _merge(_stats(pa, False), _stats(pb, False), True)  # stats_d: add(stats_a); add(b)
    assert sc[1] == 7 and sd[1] == 7
    for i in range(2, 5):
        assert double_eq(sc[i], sd[i])
    for i in range(6, 15):
        assert double_eq(sc[i], sd[i])
    expected_mean = [0.98039215686274517, 1.2254901960784315, 0.18971050224686223]  # :77-78
    for i in range(3):
        assert double_eq(expected_mean[i], sc[2 + i]), i
    expected_cov = np.array([[1.7545174932718179, -0.64263744713571724, 0],
                             [-0.64263744713571724, 0.66974240676662777, 0],
                             [0, 0, -0.38069068615840829]])  # :83-85
    cov = sc[6:].reshape(3, 3)
    for y in range(3):
        for x in range(3):
            assert double_eq(expected_cov[y, x], cov[y, x]), (x, y)


# --------------------------------------------------------------------------
# test/quickmcl/test_pose_2d.cpp (the Eigen transform and to_decomposed parts;
# the ROS message conversions are out of scope)
# --------------------------------------------------------------------------
@pytest.mark.parametrize("trig", [ob.TRIG_LIBM, ob.TRIG_CR])
def test_pose_transform(trig):  # :79-94
    pose = np.array([3, 1, 2], np.float32)
    pt = np.array([1, 2], np.float32)
    out = np.zeros(2, np.float32)
    L.orc_transform_point_f32(pose.ctypes.data_as(C.POINTER(C.c_float)),
                              pt.ctypes.data_as(C.POINTER(C.c_float)),
                              out.ctypes.data_as(C.POINTER(C.c_float)), trig)
    assert float_eq(0.76525831, out[0])
    assert float_eq(1.0770037, out[1])


def test_pose_to_decomposed():  # :107-117
    pose = np.array([2, 1, 2], np.float32)
    out = np.zeros(4)
    L.orc_to_decomposed(pose.ctypes.data_as(C.POINTER(C.c_float)),
                        out.ctypes.data_as(C.POINTER(C.c_double)))
    assert float_eq(2, out[0]) and float_eq(1, out[1])
    assert float_eq(math.sin(2.0), out[2]) and float_eq(math.cos(2.0), out[3])


def test_pose_normalise():  # :119-128 and "math" :56-67
    assert float_eq(-math.pi + 2, L.orc_normalise_angle_f32(np.float32(math.pi + 2)))
    assert float_eq(-math.pi + 2, L.orc_normalise_angle_f64(math.pi + 2))
    assert float_eq(-math.pi / 4.0, L.orc_normalise_angle_f32(np.float32(-math.pi / 4.0)))


# --------------------------------------------------------------------------
# test/quickmcl/test_utils.cpp
# --------------------------------------------------------------------------
def test_normalise_angle():  # :7-18
    pi = math.pi
    assert float_eq(0.0, L.orc_normalise_angle_f32(0.0))
    assert double_eq(0.0, L.orc_normalise_angle_f64(0.0))
    # gtest's 4-ULP test against 0 is an equality test; the reference passes
    # because normalise_angle(2*M_PI) is exactly 0 in double.
    assert float_eq(0, L.orc_normalise_angle_f64(2 * pi))
    assert float_eq(pi - 0.1, L.orc_normalise_angle_f64(3 * pi - 0.1))
    assert float_eq(0, L.orc_normalise_angle_f64(4 * pi))
    assert float_eq(0, L.orc_normalise_angle_f64(-2 * pi))
    assert float_eq(-pi + 0.1, L.orc_normalise_angle_f64(-3 * pi + 0.1))
    assert float_eq(0, L.orc_normalise_angle_f64(-4 * pi))
    assert float_eq(-pi + 2, L.orc_normalise_angle_f64(pi + 2))


def test_angle_delta():  # :20-42
    pi = math.pi
    d = L.orc_angle_delta_f64
    table = [(-0.1, 0.1, 0.2), (0.1, 0.2, 0.1), (-0.1, -0.2, -0.1), (0.1, -0.1, -0.2),
             (-0.2, -0.1, 0.1), (0.2, 0.1, -0.1), (0.0, 0.1, 0.1), (0.0, -0.1, -0.1),
             (-0.2, pi - 0.1, -pi + 0.1), (0.2, -pi + 0.1, pi - 0.1),
             (0.1, -pi, pi - 0.1), (-0.1, -pi, -pi + 0.1),
             (0.1, pi, pi - 0.1), (-0.1, pi, -pi + 0.1)]
    for want, a, b in table:
        assert float_eq(want, d(a, b)), (want, a, b)


# --------------------------------------------------------------------------
# test/quickmcl/test_low_pass_filter.cpp
# --------------------------------------------------------------------------
@pytest.mark.parametrize("as_float", [0, 1])
def test_low_pass_filter(as_float):  # :26-42
    seq = [0.7] + [0.9] * 15 + [0.8]
    xs = np.array(seq)
    v1 = L.orc_lowpass_run(0.1, xs.ctypes.data_as(C.POINTER(C.c_double)), len(xs), as_float)
    assert float_eq(0.85293961, v1)
    xs2 = np.array(seq + [0.2])
    v2 = L.orc_lowpass_run(0.1, xs2.ctypes.data_as(C.POINTER(C.c_double)), len(xs2), as_float)
    assert float_eq(0.78764564, v2)


# --------------------------------------------------------------------------
# Known answers that are not in the reference's tests but pin the contract.
# --------------------------------------------------------------------------
def test_philox_known_answers():
    # Random123 kat_vectors for philox4x32-10.
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
        ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
        ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
         (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
    ]
    for ctr, key, want in kat:
        c = (C.c_uint32 * 4)(*ctr)
        k = (C.c_uint32 * 2)(*key)
        o = (C.c_uint32 * 4)()
        L.orc_philox4x32_10(c, k, o)
        assert tuple(o) == want


def test_kld_limit_values():
    # values of space_partitioning.cpp:52-74 evaluated at
    # eps=0.05, z=0.95, [100, 5000].
    lim = lambda k: L.orc_kld_limit(k, 0.05, 0.95, 100, 5000)
    assert lim(0) == 5000 and lim(1) == 5000
    assert lim(50) == 397 and lim(100) == 857 and lim(500) == 4690
    assert lim(1000) == 5000
    # direct evaluation of the formula
    for k in (2, 3, 7, 33, 250):
        c = 2.0 / (9 * (k - 1))
        t = 1 - c - math.sqrt(c) * 0.95
        want = min(max(int(math.ceil(((k - 1) / 0.1) * t * t * t)), 100), 5000)
        assert lim(k) == want
