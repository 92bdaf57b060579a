"""ctypes binding of oracle/_ref/libqmcl_ref.so: the reference's OWN translation units
of the particle-filter path (distance_calculator, map_base, map_likelihood, odometry,
particle, particle_filter, pose_2d, space_partitioning, statistics .cpp and the
header-only parts), built against the stand-in headers of oracle/ref_stubs; see
oracle/ref_shim.cpp and oracle/Makefile.

TEST INFRASTRUCTURE ONLY: imported by tests/test_oracle_vs_ref.py and
tests/test_oracle_vs_ref_filter.py, which hold the restated oracle against it, by
tests/test_gpu_vs_ref.py, and by bench.py's --impl reference / cpu_baseline legs as the
timed CPU baseline.  The library is built in this container, where
/root/reference exists (`make -C oracle ref`, also run by __graft_entry__.build());
elsewhere the prebuilt file is loaded, and `available()` is False without one.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

from .binding import (PARTICLE_DTYPE, LikelihoodParams, Particle, PFParams, P_DRAW, P_GLOBAL, P_INJECT,
                      P_LV_OFFSET, P_MOTION, normal_pairs, particles_array, rng_words)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libqmcl_ref.so")
REFERENCE_ROOT = os.environ.get("QMCL_REFERENCE_ROOT", "/root/reference")
_lib = None


_available = None


def available():
    """True when the library is there AND loads with every entry point this binding names (a
    stale prebuilt file from an older build must make the tests skip, not error)."""
    global _available
    if _available is None:
        if os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "quickmcl")):
            subprocess.check_call(["make", "-s", "-C", _HERE, "ref", f"REF={REFERENCE_ROOT}"],
                                  stdout=subprocess.DEVNULL)
        _available = False
        if os.path.exists(LIB_PATH):
            try:
                _load()
                _available = True
            except (OSError, AttributeError) as e:
                print(f"oracle/_ref: {LIB_PATH} is not usable ({e})")
    return _available


def lib():
    if _lib is None and not available():
        raise RuntimeError("oracle/_ref/libqmcl_ref.so is missing or unusable and there is no reference tree to "
                           "build it from")
    return _lib


def _load():
    global _lib
    L = C.CDLL(LIB_PATH)
    f32, f64, i32, u32, u64, sz = C.c_float, C.c_double, C.c_int32, C.c_uint32, C.c_uint64, C.c_size_t
    pf32, pf64, pi32, pi8 = C.POINTER(f32), C.POINTER(f64), C.POINTER(i32), C.POINTER(C.c_int8)
    pint, pP = C.POINTER(C.c_int), C.POINTER(Particle)
    vp, i8, pu64 = C.c_void_p, C.c_int8, C.POINTER(u64)
    pL, pF = C.POINTER(LikelihoodParams), C.POINTER(PFParams)
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
        "ref_script": (None, [pu64, sz, pf64, sz]),
        "ref_script_state": (C.c_int, [pu64]),
        "ref_last_seconds": (f64, []),
        "ref_map_create": (vp, [pi8, u32, u32, f32, f64, f64, pL]),
        "ref_map_destroy": (None, [vp]),
        "ref_map_copy_field": (None, [vp, pf32]),
        "ref_map_set_field": (None, [vp, pf32]),
        "ref_map_copy_state": (None, [vp, pi8]),
        "ref_map_num_free": (sz, [vp]),
        "ref_map_copy_free_cells": (None, [vp, pi32]),
        "ref_map_state_at": (i8, [vp, f32, f32]),
        "ref_map_probability_at": (f64, [vp, f32, f32]),
        "ref_map_likelihood_grid": (None, [vp, pi8]),
        "ref_map_random_pose": (None, [vp, pf32]),
        "ref_map_update_importance": (f64, [vp, pf32, sz, pP, sz]),
        "ref_statistics": (None, [pP, sz, C.c_int, pf64]),
        "ref_sample_odometry": (None, [pf64, pf64, pP, sz, pf32]),
        "ref_pose_restore": (C.c_int, [pf64, C.c_int, pf32, pf32, C.c_char_p, sz]),
        "ref_pose_store": (C.c_int, [C.c_int, pf64, pf64, pf64, C.c_char_p, sz]),
        "ref_cycle_create": (vp, [C.c_int, f32, f32, pf32, f64]),
        "ref_cycle_destroy": (None, [vp]),
        "ref_cycle_force_pose_reset": (None, [vp]),
        "ref_cycle_set_particles_present": (None, [vp, C.c_int]),
        "ref_cycle_on_scan": (sz, [vp, pf64, C.c_int, sz, f64, C.c_char_p, sz]),
        "ref_filter_create": (vp, [vp]),
        "ref_filter_destroy": (None, [vp]),
        "ref_filter_set_params": (None, [vp, pF]),
        "ref_filter_initialise": (None, [vp, pf32, pf32]),
        "ref_filter_global_localization": (None, [vp]),
        "ref_filter_handle_odometry": (None, [vp, pf64, pf64, pf32]),
        "ref_filter_update_sensor": (None, [vp, pf32, sz]),
        "ref_filter_resample_and_cluster": (C.c_int, [vp]),
        "ref_filter_cluster": (None, [vp]),
        "ref_filter_get_estimate": (C.c_int, [vp, pf64, pf64]),
        "ref_filter_num_particles": (sz, [vp]),
        "ref_filter_copy_particles": (None, [vp, pP]),
        "ref_filter_set_particles": (None, [vp, pP, sz]),
        "ref_filter_get_lowpass": (None, [vp, pf64]),
        "ref_filter_set_lowpass": (None, [vp, f64, f64]),
        "ref_filter_num_clusters": (C.c_int, [vp]),
        "ref_filter_cluster_ids": (None, [vp, pi32]),
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


# ------------------------------------------------------------------ scripted randomness
def script(words=(), normals=()):
    """Replaces the script the reference's (stand-in) engines read: `words` feed every uniform
    draw in call order, `normals` every normal draw (oracle/ref_stubs/random)."""
    w = np.ascontiguousarray(words, np.uint64)
    z = np.ascontiguousarray(normals, np.float64)
    lib().ref_script(_p(w, C.c_uint64), len(w), _p(z, C.c_double), len(z))


def script_state():
    """(words consumed, normals consumed, a draw found the script empty)."""
    used = np.zeros(2, np.uint64)
    under = lib().ref_script_state(_p(used, C.c_uint64))
    return int(used[0]), int(used[1]), bool(under)


def last_seconds():
    return lib().ref_last_seconds()


def _word53(lo, hi):
    """64-bit word whose libstdc++ uniform_real<double> image is exactly the contract's
    u53(lo, hi) = ((hi << 32 | lo) >> 11) * 2^-53 (the low 11 bits would round, not truncate)."""
    v = (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
    return (v >> np.uint64(11)) << np.uint64(11)


def u53(lo, hi):
    v = (hi.astype(np.uint64) << np.uint64(32)) | lo.astype(np.uint64)
    return (v >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def motion_script(seed, step, n):
    """Normals of one handle_odometry over n particles: per particle (trans, rot1, rot2) =
    (z0, z1) of words 0-1 and z0 of words 2-3 (qmcl_oracle.cpp sample_odometry)."""
    r = rng_words(seed, step, P_MOTION, n)
    a = normal_pairs(r[:, 0:2])
    b = normal_pairs(r[:, 2:4])
    return np.stack([a[:, 0], a[:, 1], b[:, 0]], axis=1).reshape(-1)


def global_script(seed, step, n):
    """Words of one global_localization over n particles: per particle (cell, heading)."""
    r = rng_words(seed, step, P_GLOBAL, n)
    cell = (r[:, 1].astype(np.uint64) << np.uint64(32)) | r[:, 0].astype(np.uint64)
    head = r[:, 2].astype(np.uint64) << np.uint64(32)
    return np.stack([cell, head], axis=1).reshape(-1)


def low_variance_script(seed, step):
    r = rng_words(seed, step, P_LV_OFFSET, 1)
    return _word53(r[:, 0], r[:, 1])


def draw_script(seed, step, n_slots, w_diff):
    """Words of an adaptive / KLD resampling of up to n_slots slots, in the order the reference
    consumes them (particle_filter.cpp:261-276, :313-327): per slot the word of p, then either
    (cell, heading) of the injected pose or the word of r.  Also returns the inject mask."""
    r = rng_words(seed, step, P_DRAW, n_slots)
    p = u53(r[:, 0], r[:, 1])
    inject = p < w_diff
    q = rng_words(seed, step, P_INJECT, n_slots)
    pw = _word53(r[:, 0], r[:, 1])
    rw = _word53(r[:, 2], r[:, 3])
    cell = (q[:, 1].astype(np.uint64) << np.uint64(32)) | q[:, 0].astype(np.uint64)
    head = q[:, 2].astype(np.uint64) << np.uint64(32)
    # slot m holds 2 words (p, r) or 3 (p, cell, heading)
    start = np.zeros(n_slots, np.int64)
    if n_slots:
        np.cumsum(2 + inject[:-1].astype(np.int64), out=start[1:])
    words = np.zeros(int(2 * n_slots + inject.sum()), np.uint64)
    words[start] = pw
    keep = ~inject
    words[start[keep] + 1] = rw[keep]
    words[start[inject] + 1] = cell[inject]
    words[start[inject] + 2] = head[inject]
    return words, inject


# ------------------------------------------------------------------ MapLikelihood / ParticleFilter
class RefMap:
    """The reference's MapLikelihood (map_likelihood.cpp, map_base.cpp), built from an occupancy grid."""

    def __init__(self, occ, resolution, origin_xy, z_hit=0.5, z_rand=0.5, num_beams=30, max_laser_distance=14.0,
                 max_obstacle_distance=2.0, sigma_hit=0.01):
        self.L = lib()
        occ = np.ascontiguousarray(occ, np.int8)
        self.height, self.width = occ.shape
        self.params = LikelihoodParams(z_hit, z_rand, num_beams, max_laser_distance, max_obstacle_distance,
                                       sigma_hit)
        self.h = self.L.ref_map_create(_p(occ, C.c_int8), self.width, self.height, resolution, origin_xy[0],
                                       origin_xy[1], C.byref(self.params))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_map_destroy(self.h)
            self.h = None

    def field(self):
        out = np.zeros((self.height, self.width), np.float32)
        self.L.ref_map_copy_field(self.h, _p(out, C.c_float))
        return out

    def set_field(self, field):
        f = np.ascontiguousarray(field, np.float32)
        assert f.shape == (self.height, self.width)
        self.L.ref_map_set_field(self.h, _p(f, C.c_float))

    def state(self):
        out = np.zeros((self.height, self.width), np.int8)
        self.L.ref_map_copy_state(self.h, _p(out, C.c_int8))
        return out

    def free_cells(self):
        n = self.L.ref_map_num_free(self.h)
        out = np.zeros((n, 2), np.int32)
        if n:
            self.L.ref_map_copy_free_cells(self.h, _p(out, C.c_int32))
        return out

    def state_at(self, x, y):
        return int(self.L.ref_map_state_at(self.h, x, y))

    def probability_at(self, x, y):
        return self.L.ref_map_probability_at(self.h, x, y)

    def likelihood_grid(self):
        out = np.zeros((self.height, self.width), np.int8)
        self.L.ref_map_likelihood_grid(self.h, _p(out, C.c_int8))
        return out

    def random_pose(self):
        out = np.zeros(3, np.float32)
        self.L.ref_map_random_pose(self.h, _p(out, C.c_float))
        return out

    def update_importance(self, cloud_xy, particles):
        cloud = np.ascontiguousarray(cloud_xy, np.float32)
        assert particles.dtype == PARTICLE_DTYPE and particles.flags.c_contiguous
        return self.L.ref_map_update_importance(self.h, _p(cloud, C.c_float), len(cloud), _p(particles, Particle),
                                                len(particles))


def statistics(particles, finalise=True):
    out = np.zeros(15)
    lib().ref_statistics(_p(particles, Particle), len(particles), int(finalise), _p(out, C.c_double))
    return out


def sample_odometry(odom_new, odom_old, particles, alpha):
    a, b = np.asarray(odom_new, np.float64), np.asarray(odom_old, np.float64)
    al = np.asarray(alpha, np.float32)
    lib().ref_sample_odometry(_p(a, C.c_double), _p(b, C.c_double), _p(particles, Particle), len(particles),
                              _p(al, C.c_float))


class RefFilter:
    """The reference's ParticleFilter (particle_filter.cpp) on a RefMap; same method names as
    oracle.binding.OracleFilter.  Every call that draws must be preceded by script(...)."""

    def __init__(self, rmap):
        self.L = lib()
        self.map = rmap
        self.h = self.L.ref_filter_create(rmap.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_filter_destroy(self.h)
            self.h = None

    def set_parameters(self, resample_type=2, particle_count_min=100, particle_count_max=5000, resample_count=2,
                       alpha_fast=0.1, alpha_slow=0.001, kld_epsilon=0.01, kld_z=0.01, res_xy=0.5,
                       res_theta=float(np.float32(np.pi / 18))):
        self.params = PFParams(resample_type, particle_count_min, particle_count_max, resample_count, alpha_fast,
                               alpha_slow, kld_epsilon, kld_z, res_xy, res_theta)
        self.L.ref_filter_set_params(self.h, C.byref(self.params))

    def initialise(self, pose, cov):
        pose = np.asarray(pose, np.float32)
        cov = np.ascontiguousarray(cov, np.float32).reshape(9)
        self.L.ref_filter_initialise(self.h, _p(pose, C.c_float), _p(cov, C.c_float))

    def global_localization(self):
        self.L.ref_filter_global_localization(self.h)

    def handle_odometry(self, odom_new, odom_old, alpha):
        a, b = np.asarray(odom_new, np.float64), np.asarray(odom_old, np.float64)
        al = np.asarray(alpha, np.float32)
        self.L.ref_filter_handle_odometry(self.h, _p(a, C.c_double), _p(b, C.c_double), _p(al, C.c_float))

    def update_importance_from_observations(self, cloud_xy):
        cloud = np.ascontiguousarray(cloud_xy, np.float32)
        self.L.ref_filter_update_sensor(self.h, _p(cloud, C.c_float), len(cloud))

    def resample_and_cluster(self):
        """False when the reference threw (its low-variance walk ran off the end of the set)."""
        return self.L.ref_filter_resample_and_cluster(self.h) == 0

    def cluster(self):
        self.L.ref_filter_cluster(self.h)

    def get_estimated_pose(self):
        pose, cov = np.zeros(3), np.zeros(9)
        ok = self.L.ref_filter_get_estimate(self.h, _p(pose, C.c_double), _p(cov, C.c_double))
        return ok, pose, cov.reshape(3, 3)

    def num_particles(self):
        return self.L.ref_filter_num_particles(self.h)

    def get_particles(self):
        out = particles_array(self.num_particles())
        if len(out):
            self.L.ref_filter_copy_particles(self.h, _p(out, Particle))
        return out

    def set_particles(self, p):
        p = np.ascontiguousarray(p)
        assert p.dtype == PARTICLE_DTYPE
        self.L.ref_filter_set_particles(self.h, _p(p, Particle), len(p))

    def lowpass(self):
        out = np.zeros(2)
        self.L.ref_filter_get_lowpass(self.h, _p(out, C.c_double))
        return out

    def set_lowpass(self, w_slow, w_fast):
        self.L.ref_filter_set_lowpass(self.h, w_slow, w_fast)

    def num_clusters(self):
        return self.L.ref_filter_num_clusters(self.h)

    def cluster_ids(self):
        out = np.zeros(self.num_particles(), np.int32)
        if len(out):
            self.L.ref_filter_cluster_ids(self.h, _p(out, C.c_int32))
        return out


# ------------------------------------------------------------------ PoseRestorer (pose_restore.cpp)
def pose_restore(v=None):
    """PoseRestorer::restore_pose with ~initial_pose = v (None: parameter absent).  Returns
    (initialise calls, pose float32[3], covariance float32[3, 3], warning lines)."""
    vals = np.ascontiguousarray([] if v is None else v, np.float64)
    pose, cov = np.zeros(3, np.float32), np.zeros(9, np.float32)
    buf = C.create_string_buffer(4096)
    calls = lib().ref_pose_restore(_p(vals, C.c_double), -1 if v is None else len(vals), _p(pose, C.c_float),
                                   _p(cov, C.c_float), buf, len(buf))
    return calls, pose, cov.reshape(3, 3), [ln for ln in buf.value.decode().split("\n") if ln]


def pose_store(ok, pose, cov):
    """PoseRestorer::store_pose on a filter whose estimate is (ok, pose, cov).  Returns
    (the stored ~initial_pose list or None, warning lines)."""
    pose = np.ascontiguousarray(pose, np.float64)
    cov = np.ascontiguousarray(cov, np.float64).reshape(9)
    out = np.zeros(6)
    buf = C.create_string_buffer(4096)
    n = lib().ref_pose_store(int(bool(ok)), _p(pose, C.c_double), _p(cov, C.c_double), _p(out, C.c_double), buf,
                             len(buf))
    return (None if n < 0 else out[:n].tolist()), [ln for ln in buf.value.decode().split("\n") if ln]


# ------------------------------------------------------------------ LaserHandler::Impl::cloud_callback
class RefScanCycle:
    """The reference's LaserHandler (laser_handler.cpp) wired to recorders: on_scan() delivers one
    point cloud to its callback and returns the log of what it did - IParticleFilter calls with
    their arguments (doubles and floats as C hex floats), publish_* calls, "log: ..." lines."""

    def __init__(self, resample_count=2, min_trans=0.2, min_rot=math.pi / 6, alpha=(0.05, 0.1, 0.02, 0.05),
                 save_pose_period=0.0):
        self.L = lib()
        a = np.asarray(alpha, np.float32)
        self.h = self.L.ref_cycle_create(resample_count, min_trans, min_rot, _p(a, C.c_float), save_pose_period)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.ref_cycle_destroy(self.h)
            self.h = None

    def force_pose_reset(self):
        self.L.ref_cycle_force_pose_reset(self.h)

    def set_particles_present(self, present):
        self.L.ref_cycle_set_particles_present(self.h, int(bool(present)))

    def on_scan(self, odom, n_points, now=0.0, odom_ok=True):
        o = np.asarray(odom, np.float64)
        buf = C.create_string_buffer(1 << 14)
        self.L.ref_cycle_on_scan(self.h, _p(o, C.c_double), int(bool(odom_ok)), n_points, now, buf, len(buf))
        return [ln for ln in buf.value.decode().split("\n") if ln]
