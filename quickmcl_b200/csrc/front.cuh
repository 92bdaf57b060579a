// front.cuh — the front kernel of a sensor update on the fused path (filter.cu):
// deferred motion model + beam selection + bin bounds in one launch, for one
// filter or for a batch of filters on one map.
#pragma once

#include "cluster_dev.cuh"
#include "motion_dev.cuh"

namespace qmcl {

struct FrontJob {
  const float2 *cloud;
  size_t n_points, step;
  float2 *beams;
  int n_used;
  unsigned *rmax;
  float *x, *y, *t;
  size_t n, global_first;
  int do_motion;
  MotionParams mp;
  uint64_t seed;
  uint32_t rng_step;
  BinParams bp;
  int *block_bbox;  // [gridDim.x - 1][6]
  int *bbox;        // [6]
  unsigned int *blocks_done;
  // the map the sensor kernel is about to gather from (one-byte palette): its
  // lines are requested into L2 while the motion model runs
  const uint8_t *prefetch;
  size_t prefetch_bytes;
};


// Fills f's front job for the cloud at dev_cloud_xy and does the host-side
// bookkeeping of its launch; the caller launches it (front_kernel alone, or
// launch_front_batch for many filters sharing a map).
qmcl_status front_prepare(qmcl_filter *f, const float *dev_cloud_xy, size_t n_points, FrontJob *J,
                          int *n_used_out);
qmcl_status launch_front_batch(qmcl_map *m, const FrontJob *dev_jobs, unsigned n_jobs, size_t max_particles);
// K7's job for f after front_prepare, K8 left to the fused tail (filter.cu).
struct SensorLaunch;
qmcl_status sensor_prepare_deferred(qmcl_filter *f, int n_used, size_t n_points, SensorLaunch *a);

}  // namespace qmcl
