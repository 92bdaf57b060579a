"""The scan cycle of QuickMCL's node without ROS and the pose
checkpoint (8f-4).

`ScanCycle.on_scan(odom, cloud)` restates `LaserHandler::Impl::cloud_callback`
(src/quickmcl_node/laser_handler.cpp:179-288): the moved-enough gate, the
fall-back to global localisation when the particle set is empty, motion ->
sensor -> resample-or-cluster with the reference's cadence, and the pose
estimate that `Publishing::publish_estimated_pose` (publishing.cpp:157-227)
would turn into the map->odom transform.  It drives any object with the
IParticleFilter methods: the CUDA filter, the sharded filter, or (in the tests)
the CPU oracle.

Cadence quirk kept on purpose (laser_handler.cpp:259-260): the filter resamples
when `counter++ % resample_count` is NON-zero, i.e. with the default 2 on every
second processed scan starting with the second, and never with 1.
"""
import dataclasses
import math
from typing import Optional, Sequence

import numpy as np


@dataclasses.dataclass
class MotionModelParameters:
    """Parameters::MotionModel (parameters.h:40-50); alpha as the node's cfg (QuickMCL.cfg:25-28).
    The reference keeps the two thresholds as floats and compares doubles with them
    (laser_handler.cpp:207-209): on_scan rounds them to float32 before use, so 0.2 means
    0.2f = 0.20000000298 here as it does there."""
    alpha: Sequence[float] = (0.05, 0.1, 0.02, 0.05)
    min_trans: float = 0.2
    min_rot: float = math.pi / 6


@dataclasses.dataclass
class CycleResult:
    processed: bool              # False: "Skipping laser scan: haven't moved"
    resampled: bool
    global_localization: bool    # the set was empty and was re-seeded
    recompute_transform: bool    # what publish_estimated_pose is told
    pose_ok: bool                # get_estimated_pose's return value
    pose: Optional[np.ndarray]
    covariance: Optional[np.ndarray]


def _normalise_angle(a):
    """utils.h:28-35"""
    return a - 2 * math.pi * math.floor((a + math.pi) / (2 * math.pi))


class ScanCycle:
    def __init__(self, particle_filter, motion: MotionModelParameters = None, resample_count=2):
        self.filter = particle_filter
        self.motion = motion or MotionModelParameters()
        self.resample_count = int(resample_count)
        self.odom_at_last_filter_run = None   # laser_handler.cpp:170
        self.pose_reset = False               # set by force_pose_reset (:148-151)
        self.resample_counter = 0             # :176
        self.calls = []                       # the IParticleFilter calls made, for tests/tracing

    def force_pose_reset(self):
        """LaserHandler::force_pose_reset: the next scan is processed even without motion."""
        self.pose_reset = True

    def _call(self, name, *args):
        self.calls.append(name)
        return getattr(self.filter, name)(*args)

    def on_scan(self, odom_new, cloud) -> CycleResult:
        odom_new = np.asarray(odom_new, np.float64)
        recompute = False
        if self.odom_at_last_filter_run is None:
            self.odom_at_last_filter_run = odom_new.copy()
            recompute = True
        diff = odom_new - self.odom_at_last_filter_run
        diff[2] = _normalise_angle(diff[2])
        min_trans = float(np.float32(self.motion.min_trans))
        min_rot = float(np.float32(self.motion.min_rot))
        has_moved = abs(diff[0]) > min_trans or abs(diff[1]) > min_trans or abs(diff[2]) > min_rot
        if not self.pose_reset and not has_moved:
            self._call("cluster")
            ok, pose, cov = self._call("get_estimated_pose")
            return CycleResult(False, False, False, recompute, bool(ok), pose, cov)
        if self.pose_reset:
            recompute = True
        self.pose_reset = False
        did_global = False
        if self.filter.num_particles() == 0:      # get_particles().empty(), without the copy
            self._call("global_localization")
            recompute = True
            did_global = True
        self._call("handle_odometry", odom_new, self.odom_at_last_filter_run,
                   np.asarray(self.motion.alpha, np.float32))
        self.odom_at_last_filter_run = odom_new.copy()
        self._call("update_importance_from_observations", cloud)
        counter = self.resample_counter
        self.resample_counter = (self.resample_counter + 1) & 0xffffffff
        resampled = (counter % self.resample_count) != 0
        if resampled:
            recompute = True
            self._call("resample_and_cluster")
        else:
            self._call("cluster")
        ok, pose, cov = self._call("get_estimated_pose")
        return CycleResult(True, resampled, did_global, recompute, bool(ok), pose, cov)


class PoseCheckpoint:
    """PoseRestorer without the parameter server (src/quickmcl/pose_restore.cpp:37-100).

    `store()` is store_pose: the six numbers the node writes to `~initial_pose`
    (x, y, theta and the diagonal of the covariance), or None when the filter has
    no estimate ("Couldn't get pose to store").  `restore(v)` is restore_pose: no
    stored vector means pose 0 with the identity covariance (:59-60); a stored one
    is read entry by entry, and a missing or NaN entry falls back to 0 for the
    pose, 0.5^2 / 0.5^2 / 0.1^2 for the variances (:66-71, :80-97).  The filter is
    then re-initialised with the covariance cast to float (:75-77).
    """
    DEFAULTS = (0.0, 0.0, 0.0, 0.5 * 0.5, 0.5 * 0.5, 0.1 * 0.1)

    def __init__(self, particle_filter):
        self.filter = particle_filter
        self.warnings = []          # what ROS_WARN would have said

    def store(self):
        ok, pose, cov = self.filter.get_estimated_pose()
        if not ok:
            self.warnings.append("Couldn't get pose to store")
            return None
        cov = np.asarray(cov, np.float64).reshape(3, 3)
        return [float(pose[0]), float(pose[1]), float(pose[2]),
                float(cov[0, 0]), float(cov[1, 1]), float(cov[2, 2])]

    def _nan_safe(self, v, index):
        if index >= len(v):
            self.warnings.append(f"Ignoring initial_pose[{index}] because it is missing")
            return self.DEFAULTS[index]
        if math.isnan(v[index]):
            self.warnings.append(f"Ignoring initial_pose[{index}] because of NaN")
            return self.DEFAULTS[index]
        return float(v[index])

    def restore(self, v=None):
        pose = np.zeros(3)
        cov = np.eye(3)
        if v is not None:
            vals = [self._nan_safe(list(v), i) for i in range(6)]
            pose[:] = vals[:3]
            cov[0, 0], cov[1, 1], cov[2, 2] = vals[3:]
        self.filter.initialise(pose.astype(np.float32), cov.astype(np.float32))
        return pose, cov
