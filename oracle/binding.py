"""ctypes binding of the CPU oracle (oracle/libqmcl_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(quickmcl_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libqmcl_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "qmcl_oracle.cpp")
    hdr = os.path.join(_HERE, "qmcl_oracle.h")
    stale = (not os.path.exists(_LIB_PATH)
             or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


class Particle(C.Structure):
    """quickmcl::WeightedParticle layout (particle.h:32-47), 24 bytes."""
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("theta", C.c_float),
                ("pad_", C.c_float), ("weight", C.c_double)]


PARTICLE_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("theta", "<f4"),
                           ("pad_", "<f4"), ("weight", "<f8")])
assert PARTICLE_DTYPE.itemsize == 24 and C.sizeof(Particle) == 24


class LikelihoodParams(C.Structure):
    _fields_ = [("z_hit", C.c_float), ("z_rand", C.c_float),
                ("num_beams", C.c_int32), ("max_laser_distance", C.c_float),
                ("max_obstacle_distance", C.c_float), ("sigma_hit", C.c_float)]


class PFParams(C.Structure):
    _fields_ = [("resample_type", C.c_int32), ("particle_count_min", C.c_int32),
                ("particle_count_max", C.c_int32), ("resample_count", C.c_int32),
                ("alpha_fast", C.c_double), ("alpha_slow", C.c_double),
                ("kld_epsilon", C.c_double), ("kld_z", C.c_double),
                ("res_xy", C.c_float), ("res_theta", C.c_float)]


_lib = None


def _ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp, f32, f64, i32, u32, u64, sz = (C.c_void_p, C.c_float, C.c_double, C.c_int32,
                                       C.c_uint32, C.c_uint64, C.c_size_t)
    pf32, pf64, pi32, pi8 = C.POINTER(f32), C.POINTER(f64), C.POINTER(i32), C.POINTER(C.c_int8)
    pP = C.POINTER(Particle)
    sig = {
        "orc_distance_cache": (None, [f32, f32, C.POINTER(C.c_int), pf32, sz]),
        "orc_world_to_map": (None, [pf32, pf32, pf32, C.c_int, C.POINTER(C.c_int)]),
        "orc_map_to_world": (None, [C.POINTER(C.c_int), pf32, pf32, C.c_int, pf32]),
        "orc_is_in_map": (C.c_int, [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_int]),
        "orc_normalise_angle_f64": (f64, [f64]),
        "orc_normalise_angle_f32": (f32, [f32]),
        "orc_angle_delta_f64": (f64, [f64, f64]),
        "orc_angle_delta_f32": (f32, [f32, f32]),
        "orc_transform_point_f32": (None, [pf32, pf32, pf32, C.c_int]),
        "orc_to_decomposed": (None, [pf32, pf64]),
        "orc_effective_particles": (f64, [pP, sz]),
        "orc_normalise_weights": (None, [pP, sz]),
        "orc_normalise_weights_total": (None, [pP, sz, f64]),
        "orc_equalise_weights": (None, [pP, sz]),
        "orc_running_sum_lookup": (f64, [pP, sz, pf64, sz, C.POINTER(u64)]),
        "orc_statistics": (None, [pP, sz, C.c_int, pf64]),
        "orc_statistics_merge": (None, [pf64, pf64, C.c_int, pf64]),
        "orc_lowpass_run": (f64, [f64, pf64, sz, C.c_int]),
        "orc_philox4x32_10": (None, [C.POINTER(u32), C.POINTER(u32), C.POINTER(u32)]),
        "orc_rng_words": (None, [u64, u32, u32, u64, sz, C.POINTER(u32)]),
        "orc_normal_pairs": (None, [C.POINTER(u32), sz, pf64]),
        "orc_kld_limit": (u64, [u64, f64, f64, i32, i32]),
        "orc_map_create": (vp, []),
        "orc_map_destroy": (None, [vp]),
        "orc_map_set_params": (None, [vp, C.POINTER(LikelihoodParams)]),
        "orc_map_load": (None, [vp, pi8, u32, u32, f32, f64, f64, C.c_int]),
        "orc_map_load_field": (None, [vp, pf32, pi8, u32, u32, f32, f64, f64]),
        "orc_map_copy_distance": (None, [vp, pf32]),
        "orc_map_copy_field": (None, [vp, pf32]),
        "orc_map_copy_state": (None, [vp, pi8]),
        "orc_map_num_free": (sz, [vp]),
        "orc_map_copy_free_cells": (None, [vp, pi32]),
        "orc_map_state_at": (C.c_int8, [vp, f32, f32]),
        "orc_map_likelihood_grid": (None, [vp, pi8]),
        "orc_map_update_importance": (f64, [vp, pf32, sz, pP, sz, C.c_int]),
        "orc_bruteforce_d2": (None, [pi8, u32, u32, C.c_int, pi32]),
        "orc_filter_create": (vp, [vp, u64]),
        "orc_filter_destroy": (None, [vp]),
        "orc_filter_set_params": (None, [vp, C.POINTER(PFParams)]),
        "orc_filter_set_rng": (None, [vp, u64, u32]),
        "orc_filter_set_trig_mode": (None, [vp, C.c_int]),
        "orc_filter_initialise": (None, [vp, pf32, pf32]),
        "orc_cov_factor": (None, [pf32, pf32]),
        "orc_filter_initialise_factor": (None, [vp, pf32, pf32]),
        "orc_filter_global_localization": (None, [vp]),
        "orc_filter_handle_odometry": (None, [vp, pf64, pf64, pf32]),
        "orc_filter_update_sensor": (None, [vp, pf32, sz]),
        "orc_filter_resample_and_cluster": (None, [vp]),
        "orc_filter_cluster": (None, [vp]),
        "orc_filter_get_estimate": (C.c_int, [vp, pf64, pf64]),
        "orc_filter_num_particles": (sz, [vp]),
        "orc_filter_copy_particles": (None, [vp, pP, sz]),
        "orc_filter_set_particles": (None, [vp, pP, sz]),
        "orc_filter_copy_resample_indices": (sz, [vp, C.POINTER(C.c_int64), sz]),
        "orc_filter_num_bins": (sz, [vp]),
        "orc_filter_copy_bins": (sz, [vp, pi32, pi32, sz]),
        "orc_filter_num_clusters": (sz, [vp]),
        "orc_filter_get_lowpass": (None, [vp, pf64]),
        "orc_filter_set_lowpass": (None, [vp, f64, f64]),
        "orc_filter_last_total_weight": (f64, [vp]),
        "orc_last_seconds": (f64, []),
        "orc_project_scan": (sz, [pf32, sz, f32, f32, f32, f32, pf64, pf32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def particles_array(n):
    return np.zeros(n, dtype=PARTICLE_DTYPE)


def make_particles(x, y, theta, w):
    x = np.asarray(x)
    p = particles_array(len(x))
    p["x"], p["y"], p["theta"], p["weight"] = x, y, theta, w
    return p


FIELD_BRUSHFIRE, FIELD_EXACT_EDT = 0, 1
TRIG_LIBM, TRIG_CR = 0, 1


class OracleMap:
    """CPU counterpart of quickmcl::MapLikelihood (map_likelihood.h:35-93)."""

    def __init__(self, **params):
        self.L = lib()
        self.h = self.L.orc_map_create()
        self.width = self.height = 0
        self.set_parameters(**params)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_map_destroy(self.h)
            self.h = None

    def set_parameters(self, z_hit=0.5, z_rand=0.5, num_beams=30, max_laser_distance=14.0,
                       max_obstacle_distance=2.0, sigma_hit=0.01):
        self.params = LikelihoodParams(z_hit, z_rand, num_beams, max_laser_distance,
                                       max_obstacle_distance, sigma_hit)
        self.L.orc_map_set_params(self.h, C.byref(self.params))

    def new_map(self, occ, resolution, origin_xy, field_mode=FIELD_BRUSHFIRE):
        occ = np.ascontiguousarray(occ, dtype=np.int8)
        self.height, self.width = occ.shape
        self.L.orc_map_load(self.h, _ptr(occ, C.c_int8), self.width, self.height,
                            resolution, origin_xy[0], origin_xy[1], field_mode)

    def load_field(self, field, state, resolution, origin_xy):
        field = np.ascontiguousarray(field, dtype=np.float32)
        state = np.ascontiguousarray(state, dtype=np.int8)
        self.height, self.width = field.shape
        self.L.orc_map_load_field(self.h, _ptr(field, C.c_float), _ptr(state, C.c_int8),
                                  self.width, self.height, resolution,
                                  origin_xy[0], origin_xy[1])

    def distance(self):
        out = np.zeros((self.height, self.width), np.float32)
        self.L.orc_map_copy_distance(self.h, _ptr(out, C.c_float))
        return out

    def field(self):
        out = np.zeros((self.height, self.width), np.float32)
        self.L.orc_map_copy_field(self.h, _ptr(out, C.c_float))
        return out

    def state(self):
        out = np.zeros((self.height, self.width), np.int8)
        self.L.orc_map_copy_state(self.h, _ptr(out, C.c_int8))
        return out

    def free_cells(self):
        n = self.L.orc_map_num_free(self.h)
        out = np.zeros((n, 2), np.int32)
        if n:
            self.L.orc_map_copy_free_cells(self.h, _ptr(out, C.c_int32))
        return out

    def state_at(self, x, y):
        return int(self.L.orc_map_state_at(self.h, x, y))

    def likelihood_grid(self):
        out = np.zeros((self.height, self.width), np.int8)
        self.L.orc_map_likelihood_grid(self.h, _ptr(out, C.c_int8))
        return out

    def update_importance(self, cloud_xy, particles, trig_mode=TRIG_CR):
        cloud_xy = np.ascontiguousarray(cloud_xy, dtype=np.float32)
        assert particles.dtype == PARTICLE_DTYPE and particles.flags.c_contiguous
        return self.L.orc_map_update_importance(
            self.h, _ptr(cloud_xy, C.c_float), cloud_xy.shape[0],
            _ptr(particles, Particle), particles.shape[0], trig_mode)


class OracleFilter:
    """CPU counterpart of quickmcl::ParticleFilter (particle_filter.h)."""

    def __init__(self, omap, seed=0):
        self.L = lib()
        self.map = omap
        self.h = self.L.orc_filter_create(omap.h, seed)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_filter_destroy(self.h)
            self.h = None

    def set_parameters(self, resample_type=2, particle_count_min=100, particle_count_max=5000,
                       resample_count=2, alpha_fast=0.1, alpha_slow=0.001, kld_epsilon=0.01,
                       kld_z=0.01, res_xy=0.5, res_theta=float(np.float32(np.pi / 18))):
        self.params = PFParams(resample_type, particle_count_min, particle_count_max,
                               resample_count, alpha_fast, alpha_slow, kld_epsilon, kld_z,
                               res_xy, res_theta)
        self.L.orc_filter_set_params(self.h, C.byref(self.params))

    def set_rng(self, seed, step):
        self.L.orc_filter_set_rng(self.h, seed, step)

    def set_trig_mode(self, mode):
        self.L.orc_filter_set_trig_mode(self.h, mode)

    def initialise(self, pose, cov):
        pose = np.asarray(pose, np.float32)
        cov = np.ascontiguousarray(cov, np.float32).reshape(9)
        self.L.orc_filter_initialise(self.h, _ptr(pose, C.c_float), _ptr(cov, C.c_float))

    def initialise_factor(self, pose, factor):
        pose = np.asarray(pose, np.float32)
        factor = np.ascontiguousarray(factor, np.float32).reshape(9)
        self.L.orc_filter_initialise_factor(self.h, _ptr(pose, C.c_float),
                                            _ptr(factor, C.c_float))

    def global_localization(self):
        self.L.orc_filter_global_localization(self.h)

    def handle_odometry(self, odom_new, odom_old, alpha):
        a = np.asarray(odom_new, np.float64)
        b = np.asarray(odom_old, np.float64)
        al = np.asarray(alpha, np.float32)
        self.L.orc_filter_handle_odometry(self.h, _ptr(a, C.c_double), _ptr(b, C.c_double),
                                          _ptr(al, C.c_float))

    def update_importance_from_observations(self, cloud_xy):
        cloud_xy = np.ascontiguousarray(cloud_xy, dtype=np.float32)
        self.L.orc_filter_update_sensor(self.h, _ptr(cloud_xy, C.c_float), cloud_xy.shape[0])

    def resample_and_cluster(self):
        self.L.orc_filter_resample_and_cluster(self.h)

    def cluster(self):
        self.L.orc_filter_cluster(self.h)

    def get_estimated_pose(self):
        pose = np.zeros(3)
        cov = np.zeros(9)
        ok = self.L.orc_filter_get_estimate(self.h, _ptr(pose, C.c_double), _ptr(cov, C.c_double))
        return ok, pose, cov.reshape(3, 3)

    def num_particles(self):
        return self.L.orc_filter_num_particles(self.h)

    def get_particles(self):
        out = particles_array(self.num_particles())
        if len(out):
            self.L.orc_filter_copy_particles(self.h, _ptr(out, Particle), len(out))
        return out

    def set_particles(self, p):
        p = np.ascontiguousarray(p)
        assert p.dtype == PARTICLE_DTYPE
        self.L.orc_filter_set_particles(self.h, _ptr(p, Particle), len(p))

    def resample_indices(self):
        n = self.L.orc_filter_copy_resample_indices(self.h, None, 0)
        out = np.zeros(n, np.int64)
        if n:
            self.L.orc_filter_copy_resample_indices(self.h, _ptr(out, C.c_int64), n)
        return out

    def bins(self):
        n = self.L.orc_filter_num_bins(self.h)
        keys = np.zeros((n, 4), np.int32)
        labels = np.zeros(n, np.int32)
        if n:
            self.L.orc_filter_copy_bins(self.h, _ptr(keys, C.c_int32), _ptr(labels, C.c_int32), n)
        return keys, labels

    def num_clusters(self):
        return self.L.orc_filter_num_clusters(self.h)

    def lowpass(self):
        out = np.zeros(2)
        self.L.orc_filter_get_lowpass(self.h, _ptr(out, C.c_double))
        return out

    def set_lowpass(self, w_slow, w_fast):
        self.L.orc_filter_set_lowpass(self.h, w_slow, w_fast)

    def last_total_weight(self):
        return self.L.orc_filter_last_total_weight(self.h)


# purposes of the shared RNG contract (qmcl_oracle.cpp: enum Purpose)
P_INIT, P_GLOBAL, P_MOTION, P_LV_OFFSET, P_DRAW, P_INJECT = 1, 2, 3, 4, 5, 6


def rng_words(seed, step, purpose, count, first_index=0):
    """The four Philox words of draws first_index .. first_index + count - 1: uint32 [count, 4]."""
    out = np.zeros((count, 4), np.uint32)
    if count:
        lib().orc_rng_words(seed, step, purpose, first_index, count, _ptr(out, C.c_uint32))
    return out


def normal_pairs(words2):
    """Box-Muller pairs as the motion model makes them of two words each: float64 [n, 2]."""
    w = np.ascontiguousarray(words2, np.uint32).reshape(-1, 2)
    out = np.zeros((len(w), 2), np.float64)
    if len(w):
        lib().orc_normal_pairs(_ptr(w, C.c_uint32), len(w), _ptr(out, C.c_double))
    return out


def project_scan(ranges, angle_min, angle_increment, range_min, range_max, mount=(0.0, 0.0, 0.0)):
    """LaserScan -> base-frame points (laser_handler.cpp:108-121); float32 [k, 2]."""
    ranges = np.ascontiguousarray(ranges, np.float32)
    mount = np.asarray(mount, np.float64)
    out = np.zeros((len(ranges), 2), np.float32)
    k = lib().orc_project_scan(_ptr(ranges, C.c_float), len(ranges), angle_min, angle_increment,
                               range_min, range_max, _ptr(mount, C.c_double), _ptr(out, C.c_float))
    return out[:k]


def last_seconds():
    return lib().orc_last_seconds()


def export_markers(particles):
    """Restates Publishing::publish_cloud / to_marker
    (src/quickmcl_node/publishing.cpp:45-67, 114-154) on a particle array:
    max_weight is the largest weight of the set (:125-131, starts at 0), and per
    particle the marker carries x, y, sin(theta / 2.0f), cos(theta / 2.0f) — float
    arithmetic, std::sin/cos(float) — and float(weight / max_weight) (:61-62).
    Returns ((n, 5) float32, max_weight)."""
    w = particles["weight"].astype(np.float64)
    max_weight = float(max(0.0, w.max())) if len(w) else 0.0
    half = particles["theta"].astype(np.float32) / np.float32(2.0)
    out = np.zeros((len(w), 5), np.float32)
    out[:, 0] = particles["x"]
    out[:, 1] = particles["y"]
    out[:, 2] = np.sin(half, dtype=np.float32)
    out[:, 3] = np.cos(half, dtype=np.float32)
    with np.errstate(all="ignore"):
        out[:, 4] = (w / max_weight).astype(np.float32)
    return out, max_weight
