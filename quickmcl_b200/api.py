"""Host-side mirror of the reference's plugin interface, over the C ABI.

Same names, argument meaning and error behaviour as
quickmcl::Map (include/quickmcl/map.h:31-62),
quickmcl::IParticleFilter (include/quickmcl/i_particle_filter.h:30-85) and the
two factories (map_factory.h:29, particle_filter_factory.h:30-31).  The C++
mirror of the same interface lives in quickmcl_b200/host/.

Everything that computes runs in libqmcl_b200.so on the GPU; nothing here
falls back to the CPU.
"""
import ctypes as C
import dataclasses
import enum
import math

import numpy as np

from . import _abi
from ._abi import PARTICLE_DTYPE, Particle, check, ptr


class ResampleType(enum.IntEnum):
    """parameters.h:28-35"""
    low_variance = 0
    adaptive = 1
    kld = 2


class MapState(enum.IntEnum):
    """map_state.h:26-33"""
    Free = 0
    Occupied = 1
    Unknown = 2


@dataclasses.dataclass
class LikelihoodMapParameters:
    """Parameters::LikelihoodMap with the reference's defaults (parameters.h:90-113)."""
    z_hit: float = 0.5
    z_rand: float = 0.5
    num_beams: int = 30
    max_laser_distance: float = 14.0
    max_obstacle_distance: float = 2.0
    sigma_hit: float = 0.01


@dataclasses.dataclass
class ParticleFilterParameters:
    """Parameters::ParticleFilter with the reference's defaults (parameters.h:55-85)."""
    resample_type: ResampleType = ResampleType.kld
    particle_count_min: int = 100
    particle_count_max: int = 5000
    resample_count: int = 2
    alpha_fast: float = 0.1
    alpha_slow: float = 0.001
    kld_epsilon: float = 0.01
    kld_z: float = 0.01
    space_partitioning_resolution_xy: float = 0.5
    space_partitioning_resolution_theta: float = float(np.float32(math.pi / 18))


@dataclasses.dataclass
class OccupancyGrid:
    """The fields of nav_msgs/OccupancyGrid the map reads (map_base.cpp:30-40)."""
    data: np.ndarray          # int8 [height, width], row-major
    resolution: float
    origin_x: float = 0.0
    origin_y: float = 0.0

    @property
    def width(self):
        return self.data.shape[1]

    @property
    def height(self):
        return self.data.shape[0]


def particles_array(n):
    return np.zeros(n, dtype=PARTICLE_DTYPE)


def make_particles(x, y, theta, weight):
    p = particles_array(len(x))
    p["x"], p["y"], p["theta"], p["weight"] = x, y, theta, weight
    return p


class MapLikelihood:
    """quickmcl::MapLikelihood behind quickmcl::Map."""

    def __init__(self, device=-1, stream=None):
        self._L = _abi.lib()
        h = C.c_void_p()
        check(self._L.qmcl_map_create(device, C.byref(h)), "qmcl_map_create")
        self._h = h
        self.parameters = LikelihoodMapParameters()
        if stream is not None:
            self.set_stream(stream)

    def close(self):
        if getattr(self, "_h", None):
            self._L.qmcl_map_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    def set_stream(self, cuda_stream):
        """cuda_stream: integer cudaStream_t (e.g. torch.cuda.Stream.cuda_stream)."""
        check(self._L.qmcl_map_set_stream(self._h, C.c_void_p(int(cuda_stream))),
              "qmcl_map_set_stream")

    def set_sensor_path(self, path):
        """Test hook: 0 = auto (3, else 4, when built here), 1 = float field gather, 2 = uint16
        codes + double table, 3 = one-byte palette, 4 = uint16 codes + float table."""
        check(self._L.qmcl_map_set_sensor_path(self._h, path), "qmcl_map_set_sensor_path")

    def set_division_mode(self, mode):
        """Test hook: cap the cell-index division form (-1 auto, 0 IEEE, 1 three-op, 2 two-op)."""
        check(self._L.qmcl_map_set_division_mode(self._h, mode), "qmcl_map_set_division_mode")

    def division_mode(self):
        """The cheapest division form the load-time self-check verified for this map."""
        v = C.c_int(-1)
        check(self._L.qmcl_map_get_division_mode(self._h, C.byref(v)), "qmcl_map_get_division_mode")
        return v.value

    # -- quickmcl::Map -------------------------------------------------------
    def set_parameters(self, parameters: LikelihoodMapParameters):
        self.parameters = parameters
        p = _abi.LikelihoodParams(parameters.z_hit, parameters.z_rand, parameters.num_beams,
                                  parameters.max_laser_distance,
                                  parameters.max_obstacle_distance, parameters.sigma_hit)
        check(self._L.qmcl_map_set_params(self._h, C.byref(p)), "qmcl_map_set_params")

    def new_map(self, msg: OccupancyGrid):
        occ = np.ascontiguousarray(msg.data, dtype=np.int8)
        check(self._L.qmcl_map_load(self._h, ptr(occ, C.c_int8), msg.width, msg.height,
                                    msg.resolution, msg.origin_x, msg.origin_y), "qmcl_map_load")

    def update_importance(self, cloud, particles):
        """Map::update_importance on a host particle array (in place). Returns the total."""
        cloud = np.ascontiguousarray(cloud, dtype=np.float32)
        assert particles.dtype == PARTICLE_DTYPE and particles.flags.c_contiguous
        total = C.c_double(0)
        check(self._L.qmcl_map_update_importance(self._h, ptr(cloud, C.c_float), cloud.shape[0],
                                                 ptr(particles, Particle), particles.shape[0],
                                                 C.byref(total)), "qmcl_map_update_importance")
        return total.value

    def generate_random_pose(self, seed=0, step=0, index=0):
        out = np.zeros(3, np.float32)
        check(self._L.qmcl_map_generate_random_pose(self._h, seed, step, index,
                                                    ptr(out, C.c_float)),
              "qmcl_map_generate_random_pose")
        return out

    def get_likelihood_as_gridmap(self):
        w, h = self.size()
        out = np.zeros((h, w), np.int8)
        check(self._L.qmcl_map_likelihood_grid(self._h, ptr(out, C.c_int8)),
              "qmcl_map_likelihood_grid")
        return out

    def get_map_state_at(self, world_coord):
        out = C.c_int8(0)
        check(self._L.qmcl_map_state_at(self._h, float(world_coord[0]), float(world_coord[1]),
                                        C.byref(out)), "qmcl_map_state_at")
        return MapState(out.value)

    # -- parity hooks --------------------------------------------------------
    def load_field(self, field, state, resolution, origin_xy):
        field = np.ascontiguousarray(field, dtype=np.float32)
        state = np.ascontiguousarray(state, dtype=np.int8)
        h, w = field.shape
        check(self._L.qmcl_map_load_field(self._h, ptr(field, C.c_float), ptr(state, C.c_int8),
                                          w, h, resolution, origin_xy[0], origin_xy[1]),
              "qmcl_map_load_field")

    def size(self):
        w, h = C.c_uint32(0), C.c_uint32(0)
        check(self._L.qmcl_map_size(self._h, C.byref(w), C.byref(h)), "qmcl_map_size")
        return w.value, h.value

    def field(self):
        w, h = self.size()
        out = np.zeros((h, w), np.float32)
        check(self._L.qmcl_map_copy_field(self._h, ptr(out, C.c_float)), "qmcl_map_copy_field")
        return out

    def distance(self):
        w, h = self.size()
        out = np.zeros((h, w), np.float32)
        check(self._L.qmcl_map_copy_distance(self._h, ptr(out, C.c_float)),
              "qmcl_map_copy_distance")
        return out

    def state(self):
        w, h = self.size()
        out = np.zeros((h, w), np.int8)
        check(self._L.qmcl_map_copy_state(self._h, ptr(out, C.c_int8)), "qmcl_map_copy_state")
        return out

    def free_cells(self):
        n = C.c_uint64(0)
        check(self._L.qmcl_map_num_free_cells(self._h, C.byref(n)), "qmcl_map_num_free_cells")
        out = np.zeros((n.value, 2), np.int32)
        if n.value:
            check(self._L.qmcl_map_copy_free_cells(self._h, ptr(out, C.c_int32)),
                  "qmcl_map_copy_free_cells")
        return out


class ParticleFilter:
    """quickmcl::ParticleFilter behind quickmcl::IParticleFilter."""

    def __init__(self, map_: MapLikelihood, seed=0):
        self._L = _abi.lib()
        self.map = map_
        h = C.c_void_p()
        check(self._L.qmcl_filter_create(map_._h, seed, C.byref(h)), "qmcl_filter_create")
        self._h = h
        self.parameters = ParticleFilterParameters()

    @classmethod
    def _borrowed(cls, map_, handle):
        """A view of a filter owned by someone else (a qmcl_batch): never destroyed from here."""
        self = cls.__new__(cls)
        self._L = _abi.lib()
        self.map = map_
        self._h = handle
        self._owned = False
        self.parameters = ParticleFilterParameters()
        return self

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self, "_owned", True):
                self._L.qmcl_filter_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()

    # -- quickmcl::IParticleFilter -------------------------------------------
    def initialise(self, starting_point, covariance):
        pose = np.asarray(starting_point, np.float32)
        cov = np.ascontiguousarray(covariance, np.float32).reshape(9)
        check(self._L.qmcl_filter_initialise(self._h, ptr(pose, C.c_float), ptr(cov, C.c_float)),
              "qmcl_filter_initialise")

    def global_localization(self):
        check(self._L.qmcl_filter_global_localization(self._h), "qmcl_filter_global_localization")

    def handle_odometry(self, odom_new, odom_old, alpha):
        a = np.asarray(odom_new, np.float64)
        b = np.asarray(odom_old, np.float64)
        al = np.asarray(alpha, np.float32)
        check(self._L.qmcl_filter_handle_odometry(self._h, ptr(a, C.c_double), ptr(b, C.c_double),
                                                  ptr(al, C.c_float)),
              "qmcl_filter_handle_odometry")

    def update_importance_from_observations(self, cloud):
        cloud = np.ascontiguousarray(cloud, dtype=np.float32)
        check(self._L.qmcl_filter_update_sensor(self._h, ptr(cloud, C.c_float), cloud.shape[0]),
              "qmcl_filter_update_sensor")

    def update_importance_from_scan(self, ranges, angle_min, angle_increment, range_min,
                                    range_max, mount=(0.0, 0.0, 0.0)):
        """The sensor update fed with a raw LaserScan: projection (laser_handler.cpp:108-121)
        and the laser->base mount transform run on the device.  Returns the number of
        valid points."""
        ranges = np.ascontiguousarray(ranges, np.float32)
        mount = np.asarray(mount, np.float64)
        n = C.c_size_t(0)
        check(self._L.qmcl_filter_update_sensor_scan(
            self._h, ptr(ranges, C.c_float), len(ranges), angle_min, angle_increment, range_min,
            range_max, ptr(mount, C.c_double), C.byref(n)), "qmcl_filter_update_sensor_scan")
        return n.value

    def last_cloud(self):
        """Test hook: the projected cloud of the last update_importance_from_scan."""
        n = C.c_size_t(0)
        check(self._L.qmcl_filter_copy_cloud(self._h, None, 0, C.byref(n)), "qmcl_filter_copy_cloud")
        out = np.zeros((n.value, 2), np.float32)
        if n.value:
            check(self._L.qmcl_filter_copy_cloud(self._h, ptr(out, C.c_float), n.value,
                                                 C.byref(n)), "qmcl_filter_copy_cloud")
        return out

    def resample_and_cluster(self):
        check(self._L.qmcl_filter_resample_and_cluster(self._h),
              "qmcl_filter_resample_and_cluster")

    def cluster(self):
        check(self._L.qmcl_filter_cluster(self._h), "qmcl_filter_cluster")

    def set_parameters(self, parameters: ParticleFilterParameters):
        self.parameters = parameters
        p = _abi.PFParams(int(parameters.resample_type), parameters.particle_count_min,
                          parameters.particle_count_max, parameters.resample_count,
                          parameters.alpha_fast, parameters.alpha_slow, parameters.kld_epsilon,
                          parameters.kld_z, parameters.space_partitioning_resolution_xy,
                          parameters.space_partitioning_resolution_theta)
        check(self._L.qmcl_filter_set_params(self._h, C.byref(p)), "qmcl_filter_set_params")

    def get_estimated_pose(self):
        """Returns (ok, pose[3] float64, covariance[3,3] float64)."""
        pose = np.zeros(3)
        cov = np.zeros(9)
        valid = C.c_int(0)
        check(self._L.qmcl_filter_get_estimate(self._h, ptr(pose, C.c_double),
                                               ptr(cov, C.c_double), C.byref(valid)),
              "qmcl_filter_get_estimate")
        return bool(valid.value), pose, cov.reshape(3, 3)

    def get_particles(self):
        n = self.num_particles()
        out = particles_array(n)
        if n:
            check(self._L.qmcl_filter_copy_particles(self._h, ptr(out, Particle), n),
                  "qmcl_filter_copy_particles")
        return out

    def export_markers(self, capacity=None):
        """The cloud as the node publishes it (publishing.cpp:45-67, 114-154): an (n, 5)
        float32 array of x, y, sin(theta/2), cos(theta/2), weight/max_weight, and max_weight."""
        n = self.num_particles() if capacity is None else min(int(capacity), self.num_particles())
        out = np.zeros((n, 5), np.float32)
        got = C.c_size_t(0)
        mw = C.c_double(0.0)
        check(self._L.qmcl_filter_export_markers(self._h, ptr(out, C.c_float), n, C.byref(got),
                                                 C.byref(mw)), "qmcl_filter_export_markers")
        return out[:got.value], mw.value

    def num_particles(self):
        n = C.c_size_t(0)
        check(self._L.qmcl_filter_num_particles(self._h, C.byref(n)), "qmcl_filter_num_particles")
        return n.value

    def effective_particles(self):
        """effective_particles(get_particles()) of particle.cpp:32-39: 1 / sum of w^2."""
        out = C.c_double(0)
        check(self._L.qmcl_filter_effective_particles(self._h, C.byref(out)),
              "qmcl_filter_effective_particles")
        return out.value

    # -- test hooks ----------------------------------------------------------
    def initialise_factor(self, starting_point, factor):
        pose = np.asarray(starting_point, np.float32)
        fac = np.ascontiguousarray(factor, np.float32).reshape(9)
        check(self._L.qmcl_filter_initialise_factor(self._h, ptr(pose, C.c_float),
                                                    ptr(fac, C.c_float)),
              "qmcl_filter_initialise_factor")

    def set_particles(self, particles):
        particles = np.ascontiguousarray(particles)
        assert particles.dtype == PARTICLE_DTYPE
        check(self._L.qmcl_filter_set_particles(self._h, ptr(particles, Particle),
                                                len(particles)), "qmcl_filter_set_particles")

    def set_rng(self, seed, step):
        check(self._L.qmcl_filter_set_rng(self._h, seed, step), "qmcl_filter_set_rng")

    def update_importance_from_device_cloud(self, dev_ptr, n_points):
        check(self._L.qmcl_filter_update_sensor_dev(self._h, C.c_void_p(int(dev_ptr)), n_points),
              "qmcl_filter_update_sensor_dev")

    def last_total_weight(self):
        out = C.c_double(0)
        check(self._L.qmcl_filter_last_total_weight(self._h, C.byref(out)),
              "qmcl_filter_last_total_weight")
        return out.value

    def resample_indices(self):
        n = C.c_size_t(0)
        check(self._L.qmcl_filter_copy_resample_indices(self._h, None, 0, C.byref(n)),
              "qmcl_filter_copy_resample_indices")
        out = np.zeros(n.value, np.int64)
        if n.value:
            check(self._L.qmcl_filter_copy_resample_indices(self._h, ptr(out, C.c_int64),
                                                            n.value, C.byref(n)),
                  "qmcl_filter_copy_resample_indices")
        return out

    def bins(self):
        n = C.c_size_t(0)
        check(self._L.qmcl_filter_num_bins(self._h, C.byref(n)), "qmcl_filter_num_bins")
        keys = np.zeros((n.value, 4), np.int32)
        labels = np.zeros(n.value, np.int32)
        if n.value:
            check(self._L.qmcl_filter_copy_bins(self._h, ptr(keys, C.c_int32),
                                                ptr(labels, C.c_int32), n.value, C.byref(n)),
                  "qmcl_filter_copy_bins")
        return keys, labels

    def num_bins(self):
        n = C.c_size_t(0)
        check(self._L.qmcl_filter_num_bins(self._h, C.byref(n)), "qmcl_filter_num_bins")
        return n.value

    def num_clusters(self):
        n = C.c_size_t(0)
        check(self._L.qmcl_filter_num_clusters(self._h, C.byref(n)), "qmcl_filter_num_clusters")
        return n.value

    def lowpass(self):
        out = np.zeros(2)
        check(self._L.qmcl_filter_get_lowpass(self._h, ptr(out, C.c_double)),
              "qmcl_filter_get_lowpass")
        return out

    def set_lowpass(self, w_slow, w_fast):
        check(self._L.qmcl_filter_set_lowpass(self._h, w_slow, w_fast), "qmcl_filter_set_lowpass")

    def set_prefix_mode(self, mode):
        """0 = parallel bit-exact running sum (default), 1 = one-warp sequential loop."""
        check(self._L.qmcl_filter_set_prefix_mode(self._h, mode), "qmcl_filter_set_prefix_mode")

    def debug_prefix(self, w, carry_in=0.0):
        """Sequentially rounded running sum of a host array; returns (sums, carry_out)."""
        w = np.ascontiguousarray(w, np.float64)
        out = np.zeros_like(w)
        co = C.c_double(0)
        check(self._L.qmcl_filter_debug_prefix(self._h, ptr(w, C.c_double), len(w), carry_in,
                                               ptr(out, C.c_double), C.byref(co)),
              "qmcl_filter_debug_prefix")
        return out, co.value

    def synchronize(self):
        check(self._L.qmcl_filter_synchronize(self._h), "qmcl_filter_synchronize")

    # -- measurement hooks (the reference's CodeTimer scopes, timer.cpp:37-55) --
    def set_sync_mode(self, lazy):
        """True (default): deferred 8-byte read-backs; False: one synchronisation per value."""
        check(self._L.qmcl_filter_set_sync_mode(self._h, 1 if lazy else 0), "qmcl_filter_set_sync_mode")

    def set_timing(self, level):
        """0 off, 1 CUDA events around the sensor kernel, 2 around every stage."""
        check(self._L.qmcl_filter_set_timing(self._h, level), "qmcl_filter_set_timing")

    def get_timing(self, reset=True):
        """{stage name: (device ms, occurrences)} accumulated since the last reset."""
        ms = np.zeros(_abi.N_STAGES)
        cnt = np.zeros(_abi.N_STAGES, np.uint64)
        check(self._L.qmcl_filter_get_timing(self._h, ptr(ms, C.c_double), ptr(cnt, C.c_uint64),
                                             int(reset)), "qmcl_filter_get_timing")
        return {self._L.qmcl_stage_name(i).decode(): (float(ms[i]), int(cnt[i]))
                for i in range(_abi.N_STAGES)}


    def tail_profile(self, reset=True):
        """Mean device time per phase of the fused tail since the last reset:
        ([(phase name, ns per call)], fused calls, calls re-run on the multi-launch path)."""
        out = np.zeros(16, np.uint64)
        calls, fb = C.c_uint64(0), C.c_uint64(0)
        check(self._L.qmcl_filter_tail_profile(self._h, ptr(out, C.c_uint64), C.byref(calls),
                                               C.byref(fb), int(reset)), "qmcl_filter_tail_profile")
        n = max(calls.value, 1)
        return ([(self._L.qmcl_tail_phase_name(k).decode(), float(out[k]) / n) for k in range(1, 15)
                 if out[k]], calls.value, fb.value)


def create_likelihood_map(device=-1, stream=None) -> MapLikelihood:
    """map_factory.h:29"""
    return MapLikelihood(device, stream)


def create_particle_filter(map_: MapLikelihood, seed=0) -> ParticleFilter:
    """particle_filter_factory.h:30-31"""
    return ParticleFilter(map_, seed)


def kernel_launch_count():
    return _abi.lib().qmcl_kernel_launch_count()


def transfer_bytes():
    """(host->device, device->host) bytes copied by the library so far."""
    out = np.zeros(2, np.uint64)
    check(_abi.lib().qmcl_transfer_bytes(ptr(out, C.c_uint64)), "qmcl_transfer_bytes")
    return int(out[0]), int(out[1])
