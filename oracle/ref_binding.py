"""ctypes binding of oracle/_ref/libqmcl_ref.so: the reference's OWN code for the parts
of the path that build against the stand-in headers of oracle/ref_stubs
(distance_calculator.cpp, particle.cpp, space_partitioning.cpp, scaled_map.h,
low_pass_filter.h, utils.h; see oracle/ref_shim.cpp and oracle/Makefile).

TEST INFRASTRUCTURE ONLY: imported by tests/test_oracle_vs_ref.py, which holds the
restated oracle against it.  The library is built in this container, where
/root/reference exists (`make -C oracle ref`, also run by __graft_entry__.build());
elsewhere the prebuilt file is loaded, and `available()` is False without one.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from .binding import PARTICLE_DTYPE, Particle

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libqmcl_ref.so")
REFERENCE_ROOT = os.environ.get("QMCL_REFERENCE_ROOT", "/root/reference")
_lib = None


def available():
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "quickmcl")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref", f"REF={REFERENCE_ROOT}"],
                              stdout=subprocess.DEVNULL)
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not available():
        raise RuntimeError("oracle/_ref/libqmcl_ref.so is missing and there is no reference tree to build it from")
    L = C.CDLL(LIB_PATH)
    f32, f64, i32, u32, u64, sz = C.c_float, C.c_double, C.c_int32, C.c_uint32, C.c_uint64, C.c_size_t
    pf32, pf64, pi32, pi8 = C.POINTER(f32), C.POINTER(f64), C.POINTER(i32), C.POINTER(C.c_int8)
    pint, pP = C.POINTER(C.c_int), C.POINTER(Particle)
    sig = {
        "ref_distance_cache": (C.c_int, [f32, f32, pf32, sz]),
        "ref_distance_map": (None, [pi8, u32, u32, f32, f32, pf32]),
        "ref_state_map": (None, [pi8, u32, u32, pi8]),
        "ref_world_to_map2": (None, [pf32, sz, pf32, f32, pint]),
        "ref_map_to_world2": (None, [pint, sz, pf32, f32, pf32]),
        "ref_is_in_map2": (C.c_int, [pint, pint]),
        "ref_effective_particles": (f64, [pP, sz]),
        "ref_normalise_weights": (None, [pP, sz]),
        "ref_normalise_weights_total": (None, [pP, sz, f64]),
        "ref_equalise_weights": (None, [pP, sz]),
        "ref_running_sum_lookup": (f64, [pP, sz, pf64, sz, C.POINTER(u64)]),
        "ref_lowpass_run": (f64, [f64, pf64, sz, C.c_int]),
        "ref_normalise_angle_f64": (f64, [f64]),
        "ref_normalise_angle_f32": (f32, [f32]),
        "ref_angle_delta_f64": (f64, [f64, f64]),
        "ref_angle_delta_f32": (f32, [f32, f32]),
        "ref_to_decomposed": (None, [pf32, pf64]),
        "ref_space_partitioning": (C.c_int, [pP, sz, i32, i32, f64, f64, f32, f32, C.POINTER(u64), pi32, pi32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def distance_cache(resolution, max_dist):
    n = lib().ref_distance_cache(resolution, max_dist, None, 0)
    out = np.zeros((n, n), np.float32)
    lib().ref_distance_cache(resolution, max_dist, _p(out, C.c_float), out.size)
    return out


def distance_map(occ, resolution, max_dist):
    occ = np.ascontiguousarray(occ, np.int8)
    h, w = occ.shape
    out = np.zeros((h, w), np.float32)
    lib().ref_distance_map(_p(occ, C.c_int8), w, h, resolution, max_dist, _p(out, C.c_float))
    return out


def state_map(occ):
    occ = np.ascontiguousarray(occ, np.int8)
    h, w = occ.shape
    out = np.zeros((h, w), np.int8)
    lib().ref_state_map(_p(occ, C.c_int8), w, h, _p(out, C.c_int8))
    return out


def world_to_map2(pos_xy, offset, resolution):
    pos = np.ascontiguousarray(pos_xy, np.float32)
    off = np.asarray(offset, np.float32)
    out = np.zeros(pos.shape, np.intc)
    lib().ref_world_to_map2(_p(pos, C.c_float), len(pos), _p(off, C.c_float), resolution, _p(out, C.c_int))
    return out


def map_to_world2(cell_xy, offset, resolution):
    cells = np.ascontiguousarray(cell_xy, np.intc)
    off = np.asarray(offset, np.float32)
    out = np.zeros(cells.shape, np.float32)
    lib().ref_map_to_world2(_p(cells, C.c_int), len(cells), _p(off, C.c_float), resolution, _p(out, C.c_float))
    return out


def is_in_map2(cell, size):
    c, s = np.asarray(cell, np.intc), np.asarray(size, np.intc)
    return bool(lib().ref_is_in_map2(_p(c, C.c_int), _p(s, C.c_int)))


def running_sum_lookup(particles, queries):
    q = np.ascontiguousarray(queries, np.float64)
    idx = np.zeros(len(q), np.uint64)
    total = lib().ref_running_sum_lookup(_p(particles, Particle), len(particles), _p(q, C.c_double), len(q),
                                         _p(idx, C.c_uint64))
    return total, idx


def space_partitioning(particles, particle_count_min, particle_count_max, kld_epsilon, kld_z,
                       resolution_xy, resolution_theta):
    """(limit() after each add_point, bin key per particle, cluster id per particle, cluster count)."""
    assert particles.dtype == PARTICLE_DTYPE
    n = len(particles)
    limits = np.zeros(n, np.uint64)
    keys = np.zeros((n, 3), np.int32)
    cluster = np.zeros(n, np.int32)
    count = lib().ref_space_partitioning(_p(particles, Particle), n, particle_count_min, particle_count_max,
                                         kld_epsilon, kld_z, resolution_xy, resolution_theta,
                                         _p(limits, C.c_uint64), _p(keys, C.c_int32), _p(cluster, C.c_int32))
    return limits, keys, cluster, count
