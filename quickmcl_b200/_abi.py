"""ctypes view of the C ABI in include/qmcl.h (libqmcl_b200.so).

This is what a reference-side binding does: plain
pointers and sizes, no torch types.  The library is built in-tree by
`quickmcl_b200/csrc/Makefile` (or `__graft_entry__.build()`); importing fails
loudly if it is missing — there is no CPU fallback.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QMCL_B200_LIB", os.path.join(_HERE, "libqmcl_b200.so"))

STATUS_NAMES = {0: "QMCL_OK", 1: "QMCL_ERR_INVALID_ARG", 2: "QMCL_ERR_NO_DEVICE",
                3: "QMCL_ERR_CUDA", 4: "QMCL_ERR_NO_MAP", 5: "QMCL_ERR_INVALID_CLUSTER",
                6: "QMCL_ERR_CAPACITY", 7: "QMCL_ERR_COMM", 8: "QMCL_ERR_UNSUPPORTED"}


class QmclError(RuntimeError):
    def __init__(self, status, where, detail=""):
        self.status = status
        super().__init__(f"{where}: {STATUS_NAMES.get(status, status)} {detail}".strip())


class Particle(C.Structure):
    """qmcl_particle == quickmcl::WeightedParticle (particle.h:32-47), 24 bytes."""
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("theta", C.c_float),
                ("pad_", C.c_float), ("weight", C.c_double)]


PARTICLE_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("theta", "<f4"),
                           ("pad_", "<f4"), ("weight", "<f8")])


class LikelihoodParams(C.Structure):
    _fields_ = [("z_hit", C.c_float), ("z_rand", C.c_float), ("num_beams", C.c_int32),
                ("max_laser_distance", C.c_float), ("max_obstacle_distance", C.c_float),
                ("sigma_hit", C.c_float)]


class PFParams(C.Structure):
    _fields_ = [("resample_type", C.c_int32), ("particle_count_min", C.c_int32),
                ("particle_count_max", C.c_int32), ("resample_count", C.c_int32),
                ("alpha_fast", C.c_double), ("alpha_slow", C.c_double),
                ("kld_epsilon", C.c_double), ("kld_z", C.c_double),
                ("res_xy", C.c_float), ("res_theta", C.c_float)]


vp, f32, f64 = C.c_void_p, C.c_float, C.c_double
i32, u32, u64, i64, sz = C.c_int32, C.c_uint32, C.c_uint64, C.c_int64, C.c_size_t
pf32, pf64, pi32, pi8 = C.POINTER(f32), C.POINTER(f64), C.POINTER(i32), C.POINTER(C.c_int8)
pP = C.POINTER(Particle)

# name -> argtypes (every function returns qmcl_status unless listed in _OTHER)
SIGNATURES = {
    "qmcl_map_create": [C.c_int, C.POINTER(vp)],
    "qmcl_map_destroy": [vp],
    "qmcl_map_set_stream": [vp, vp],
    "qmcl_map_set_sensor_path": [vp, C.c_int],
    "qmcl_map_set_division_mode": [vp, C.c_int],
    "qmcl_map_get_division_mode": [vp, C.POINTER(C.c_int)],
    "qmcl_map_set_params": [vp, C.POINTER(LikelihoodParams)],
    "qmcl_map_load": [vp, pi8, u32, u32, f32, f64, f64],
    "qmcl_map_load_field": [vp, pf32, pi8, u32, u32, f32, f64, f64],
    "qmcl_map_size": [vp, C.POINTER(u32), C.POINTER(u32)],
    "qmcl_map_copy_field": [vp, pf32],
    "qmcl_map_copy_distance": [vp, pf32],
    "qmcl_map_copy_state": [vp, pi8],
    "qmcl_map_num_free_cells": [vp, C.POINTER(u64)],
    "qmcl_map_copy_free_cells": [vp, pi32],
    "qmcl_map_state_at": [vp, f32, f32, pi8],
    "qmcl_map_likelihood_grid": [vp, pi8],
    "qmcl_map_update_importance": [vp, pf32, sz, pP, sz, pf64],
    "qmcl_map_generate_random_pose": [vp, u64, u32, u64, pf32],
    "qmcl_filter_create": [vp, u64, C.POINTER(vp)],
    "qmcl_filter_destroy": [vp],
    "qmcl_filter_set_comm": [vp, vp],
    "qmcl_nccl_unique_id": [C.POINTER(C.c_uint8)],
    "qmcl_nccl_comm_create": [C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, vp],
    "qmcl_nccl_comm_destroy": [vp],
    "qmcl_filter_set_rebalance": [vp, C.c_int],
    "qmcl_filter_set_rebalance_threshold": [vp, f64],
    "qmcl_filter_shard_range": [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64)],
    "qmcl_filter_shard_stats": [vp, C.POINTER(u64)],
    "qmcl_filter_set_particles_shard": [vp, pP, sz, u64, u64],
    "qmcl_filter_set_params": [vp, C.POINTER(PFParams)],
    "qmcl_filter_initialise": [vp, pf32, pf32],
    "qmcl_filter_initialise_factor": [vp, pf32, pf32],
    "qmcl_filter_global_localization": [vp],
    "qmcl_filter_handle_odometry": [vp, pf64, pf64, pf32],
    "qmcl_filter_update_sensor": [vp, pf32, sz],
    "qmcl_filter_update_sensor_dev": [vp, vp, sz],
    "qmcl_filter_update_sensor_scan": [vp, pf32, sz, f32, f32, f32, f32, pf64, C.POINTER(sz)],
    "qmcl_filter_copy_cloud": [vp, pf32, sz, C.POINTER(sz)],
    "qmcl_filter_resample_and_cluster": [vp],
    "qmcl_filter_cluster": [vp],
    "qmcl_filter_get_estimate": [vp, pf64, pf64, C.POINTER(C.c_int)],
    "qmcl_filter_num_particles": [vp, C.POINTER(sz)],
    "qmcl_filter_effective_particles": [vp, pf64],
    "qmcl_filter_copy_particles": [vp, pP, sz],
    "qmcl_filter_set_particles": [vp, pP, sz],
    "qmcl_filter_set_rng": [vp, u64, u32],
    "qmcl_filter_last_total_weight": [vp, pf64],
    "qmcl_filter_copy_resample_indices": [vp, C.POINTER(i64), sz, C.POINTER(sz)],
    "qmcl_filter_num_bins": [vp, C.POINTER(sz)],
    "qmcl_filter_copy_bins": [vp, pi32, pi32, sz, C.POINTER(sz)],
    "qmcl_filter_num_clusters": [vp, C.POINTER(sz)],
    "qmcl_filter_get_lowpass": [vp, pf64],
    "qmcl_filter_set_lowpass": [vp, f64, f64],
    "qmcl_filter_synchronize": [vp],
    "qmcl_filter_set_sync_mode": [vp, C.c_int],
    "qmcl_filter_export_markers": [vp, C.POINTER(C.c_float), C.c_size_t,
                                   C.POINTER(C.c_size_t), C.POINTER(C.c_double)],
    "qmcl_batch_update": [C.POINTER(vp), sz, C.c_int, pf64, pf64, pf32, C.POINTER(pf32),
                          C.POINTER(sz), C.c_int, pf64, pf64, C.POINTER(C.c_int)],
    "qmcl_batch_create": [vp, sz, u64, C.POINTER(vp)],
    "qmcl_batch_destroy": [vp],
    "qmcl_batch_size": [vp, C.POINTER(sz)],
    "qmcl_batch_filter": [vp, sz, C.POINTER(vp)],
    "qmcl_batch_set_params": [vp, C.POINTER(PFParams)],
    "qmcl_batch_cycle": [vp, pf64, pf64, pf32, pf32, C.POINTER(sz), C.c_int, pf64, pf64,
                         C.POINTER(C.c_int)],
    "qmcl_batch_stats": [vp, C.POINTER(u64), C.POINTER(u64)],
    "qmcl_filter_set_prefix_mode": [vp, C.c_int],
    "qmcl_filter_debug_prefix": [vp, pf64, sz, f64, pf64, pf64],
    "qmcl_filter_set_timing": [vp, C.c_int],
    "qmcl_filter_get_timing": [vp, pf64, C.POINTER(u64), C.c_int],
    "qmcl_transfer_bytes": [C.POINTER(u64)],
    "qmcl_filter_tail_profile": [vp, C.POINTER(u64), C.POINTER(u64), C.POINTER(u64), C.c_int],
}
_OTHER = {
    "qmcl_version": (C.c_char_p, []),
    "qmcl_last_error": (C.c_char_p, []),
    "qmcl_kernel_launch_count": (u64, []),
    "qmcl_stage_name": (C.c_char_p, [C.c_int]),
    "qmcl_tail_phase_name": (C.c_char_p, [C.c_int]),
    "qmcl_shard_slice": (None, [u64, C.c_int, C.c_int, C.POINTER(u64), C.POINTER(u64)]),
    "qmcl_lv_owner_range": (None, [u64, f64, f64, f64, f64, C.c_int, C.c_int, C.POINTER(u64),
                                   C.POINTER(u64)]),
}
N_STAGES = 12

_lib = None


def lib():
    """Load libqmcl_b200.so; raise if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `make -C quickmcl_b200/csrc` "
            "(or __graft_entry__.build()). There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    for name, args in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = args
    for name, (res, args) in _OTHER.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(status, where):
    if status != 0:
        raise QmclError(status, where, lib().qmcl_last_error().decode())


def ptr(a, t):
    return a.ctypes.data_as(C.POINTER(t))
