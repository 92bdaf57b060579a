"""Multi-robot batch (BASELINE.json configs[4]): many independent filters on one GPU.

Every robot has its own particle set, scan and odometry; the filters never
exchange data ("replicas / round-robin, no communication").  One
filter's update at QuickMCL's default size (2000 particles) is far too small to
fill a B200 — it is a chain of ~40 short kernels and a few 8-byte read-backs —
so the batch runs `lanes` of them concurrently: each lane is a host thread with
its own CUDA stream and its own map handle (the field is replicated per lane;
0.8 MB for a 400 x 400 map), and filters are dealt round-robin to the lanes.
ctypes releases the GIL during library calls, so the lanes overlap both on the
host and on the device.

The reference has no counterpart (one filter per process); the per-robot call
sequence is laser_handler.cpp:243-276.
"""
import concurrent.futures as cf
import ctypes as C

import numpy as np
import torch

from . import _abi
from ._abi import check, ptr
from .api import (LikelihoodMapParameters, OccupancyGrid, ParticleFilter, ParticleFilterParameters,
                  create_likelihood_map, create_particle_filter)


class FilterBatch:
    def __init__(self, n_filters, grid: OccupancyGrid, map_parameters: LikelihoodMapParameters,
                 filter_parameters: ParticleFilterParameters, lanes=16, seed=0, device=None):
        self.n = int(n_filters)
        self.lanes = max(1, min(int(lanes), self.n))
        self.device = torch.cuda.current_device() if device is None else device
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(self.lanes)]
        self.maps = []
        for s in self.streams:
            m = create_likelihood_map(device=self.device)
            m.set_stream(s.cuda_stream)
            m.set_parameters(map_parameters)
            m.new_map(grid)
            self.maps.append(m)
        self.filters = []
        for i in range(self.n):
            f = create_particle_filter(self.maps[i % self.lanes], seed=seed + i)
            f.set_parameters(filter_parameters)
            self.filters.append(f)
        self.pool = cf.ThreadPoolExecutor(max_workers=self.lanes)

    def close(self):
        self.pool.shutdown(wait=True)
        for f in self.filters:
            f.close()
        for m in self.maps:
            m.close()
        self.filters, self.maps = [], []

    def _lane(self, lane, fn):
        out = {}
        with torch.cuda.device(self.device):
            for i in range(lane, self.n, self.lanes):
                out[i] = fn(i, self.filters[i])
        return out

    def for_each(self, fn):
        """Run fn(index, filter) for every filter, lanes in parallel; returns results by index."""
        futs = [self.pool.submit(self._lane, lane, fn) for lane in range(self.lanes)]
        merged = {}
        for fu in futs:
            merged.update(fu.result())
        return [merged[i] for i in range(self.n)]

    def initialise(self, poses, covariance):
        poses = np.asarray(poses, np.float32)
        return self.for_each(lambda i, f: f.initialise(poses[i], covariance))

    def update_native(self, odom_new, odom_old, alpha, scans, resample=True):
        """The same cycle through qmcl_batch_update: the lanes are native threads inside the
        library (no interpreter on the per-robot path).  Filter i runs on lane i % lanes,
        which is the lane whose map handle it was created on."""
        n = self.n
        a = np.ascontiguousarray(odom_new, np.float64).reshape(n, 3)
        b = np.ascontiguousarray(odom_old, np.float64).reshape(n, 3)
        al = np.asarray(alpha, np.float32)
        clouds = [np.ascontiguousarray(s, np.float32) for s in scans]
        cptr = (C.POINTER(C.c_float) * n)(*[ptr(c, C.c_float) for c in clouds])
        npts = (C.c_size_t * n)(*[c.shape[0] for c in clouds])
        handles = (C.c_void_p * n)(*[f._h for f in self.filters])
        poses = np.zeros((n, 3))
        covs = np.zeros((n, 9))
        valid = np.zeros(n, np.int32)
        check(_abi.lib().qmcl_batch_update(handles, n, self.lanes, ptr(a, C.c_double),
                                           ptr(b, C.c_double), ptr(al, C.c_float), cptr, npts,
                                           int(resample), ptr(poses, C.c_double),
                                           ptr(covs, C.c_double), ptr(valid, C.c_int32)),
              "qmcl_batch_update")
        return [(bool(valid[i]), poses[i], covs[i].reshape(3, 3)) for i in range(n)]

    def update(self, odom_new, odom_old, alpha, scans, resample=True):
        """One scan cycle per robot (laser_handler.cpp:243-276): motion, sensor, resample (or
        cluster), estimate.  odom_*: [n, 3]; scans: list of float32 [B_i, 2].
        Returns [(ok, pose, covariance)] per robot."""
        def one(i, f):
            f.handle_odometry(odom_new[i], odom_old[i], alpha)
            f.update_importance_from_observations(scans[i])
            if resample:
                f.resample_and_cluster()
            else:
                f.cluster()
            return f.get_estimated_pose()
        return self.for_each(one)


class FusedBatch:
    """qmcl_batch: n independent filters on ONE map handle whose scan cycle
    (laser_handler.cpp:243-276 per robot) runs as three launches for the whole batch —
    deferred motion + beam selection, sensor model, and the fused tail with one block per
    filter (csrc/tail.cu) — and one host wait.  `filters[i]` is filter i for the per-filter
    calls (initialise, set_particles, getters)."""

    def __init__(self, n_filters, grid: OccupancyGrid, map_parameters: LikelihoodMapParameters,
                 filter_parameters: ParticleFilterParameters, seed=0, device=None, stream=None):
        self._L = _abi.lib()
        self.n = int(n_filters)
        self.device = torch.cuda.current_device() if device is None else device
        self.map = create_likelihood_map(device=self.device)
        if stream is not None:
            self.map.set_stream(stream)
        self.map.set_parameters(map_parameters)
        self.map.new_map(grid)
        h = C.c_void_p()
        check(self._L.qmcl_batch_create(self.map._h, self.n, seed, C.byref(h)), "qmcl_batch_create")
        self._h = h
        self.filters = []
        for i in range(self.n):
            fh = C.c_void_p()
            check(self._L.qmcl_batch_filter(self._h, i, C.byref(fh)), "qmcl_batch_filter")
            f = ParticleFilter._borrowed(self.map, fh)
            f.set_parameters(filter_parameters)
            self.filters.append(f)

    def close(self):
        if getattr(self, "_h", None):
            for f in self.filters:
                f.close()
            self._L.qmcl_batch_destroy(self._h)
            self._h = None
            self.map.close()

    def __del__(self):
        self.close()

    def cycle(self, odom_new, odom_old, alpha, scans, resample=True):
        """One scan cycle for every robot.  odom_*: [n, 3]; scans: list of float32 [B_i, 2].
        Returns [(ok, pose, covariance)] per robot."""
        n = self.n
        a = np.ascontiguousarray(odom_new, np.float64).reshape(n, 3)
        b = np.ascontiguousarray(odom_old, np.float64).reshape(n, 3)
        al = np.asarray(alpha, np.float32)
        offs = np.zeros(n + 1, np.uint64)
        offs[1:] = np.cumsum([len(s) for s in scans])
        packed = np.ascontiguousarray(np.concatenate([np.asarray(s, np.float32).reshape(-1, 2)
                                                      for s in scans]), np.float32)
        return self.cycle_packed(a, b, al, packed, offs, resample)

    def cycle_packed(self, odom_new, odom_old, alpha, packed, offsets, resample=True):
        """The same with the scans already packed back to back (packed: [sum B_i, 2] float32,
        offsets: [n + 1] uint64 in points)."""
        n = self.n
        poses = np.zeros((n, 3))
        covs = np.zeros((n, 9))
        valid = np.zeros(n, np.int32)
        check(self._L.qmcl_batch_cycle(self._h, ptr(odom_new, C.c_double), ptr(odom_old, C.c_double),
                                       ptr(alpha, C.c_float), ptr(packed, C.c_float),
                                       offsets.ctypes.data_as(C.POINTER(C.c_size_t)), int(resample),
                                       ptr(poses, C.c_double), ptr(covs, C.c_double),
                                       ptr(valid, C.c_int32)), "qmcl_batch_cycle")
        return [(bool(valid[i]), poses[i], covs[i].reshape(3, 3)) for i in range(n)]

    def stats(self):
        """(cycles that ran fused, cycles that ran as a per-filter loop)."""
        a, b = C.c_uint64(0), C.c_uint64(0)
        check(self._L.qmcl_batch_stats(self._h, C.byref(a), C.byref(b)), "qmcl_batch_stats")
        return a.value, b.value
