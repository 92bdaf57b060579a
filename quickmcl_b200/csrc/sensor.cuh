// sensor.cuh — launcher interface of the likelihood-field sensor model (K7).
#pragma once

#include "state.cuh"

namespace qmcl {

struct SensorLaunch {
  const float *x, *y, *t;  // particle poses (SoA)
  double *w;               // out: weight after the free-space penalty
  size_t n;                // particles
  const float *beams;      // used beams, float2, base frame
  int n_beams;
  const float *rmax;       // device: upper bound of the longest used beam (launch_subsample)
  double *partials;        // >= sensor_grid(n) doubles
  FilterScalars *scalars;  // total_weight written by the last block
};

unsigned sensor_grid(size_t n_particles);

// Select every step-th point of the cloud (map_likelihood.cpp:67-74,92) into
// a dense float2 array; non-finite points are moved far outside the map.
qmcl_status launch_subsample(const qmcl_map *m, const float *dev_cloud_xy, size_t n_points,
                             float *dev_beams, int *n_used_out);
qmcl_status launch_sensor(qmcl_map *m, const SensorLaunch &a);
// The same for n_jobs filters sharing map `m`, one launch (dev_jobs: device memory).
qmcl_status launch_sensor_batch(qmcl_map *m, const SensorLaunch *dev_jobs, unsigned n_jobs,
                                size_t max_particles);
// After the map's field/geometry changed: verify the fast division, drop the LUT.
qmcl_status sensor_prepare_map(qmcl_map *m);
qmcl_status sensor_prepare_map_launch(qmcl_map *m);
qmcl_status sensor_prepare_map_finish(qmcl_map *m);

}  // namespace qmcl
