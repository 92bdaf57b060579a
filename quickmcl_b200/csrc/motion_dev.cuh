// motion_dev.cuh — K6 for one particle, shared by motion_kernel (motion.cu) and
// the front kernel of a sensor update (filter.cu), which applies a deferred
// motion in the same launch that selects the beams and takes the bin bounds.
#pragma once

#include "common.cuh"

namespace qmcl {

struct MotionParams {
  double delta_trans, delta_rot1, delta_rot2;
  double sd_trans, sd_rot1, sd_rot2;
};

// odometry.cpp:86-101 for one particle (global index `gi` selects its Philox block).
__device__ __forceinline__ void motion_one(float *x, float *y, float *t, const MotionParams &mp,
                                           uint64_t seed, uint32_t step, size_t gi) {
  uint32_t r[4];
  rng_draw(seed, step, P_MOTION, gi, r);
  double z0, z1, z2, z3;
  normal_pair(r[0], r[1], &z0, &z1);
  normal_pair(r[2], r[3], &z2, &z3);
  // std::normal_distribution(0, sd) returns z * sd + 0
  const double trans_hat = mp.delta_trans - z0 * mp.sd_trans;
  const double rot1_hat = angle_delta(mp.delta_rot1, z1 * mp.sd_rot1);
  const double rot2_hat = angle_delta(mp.delta_rot2, z2 * mp.sd_rot2);
  const float th = *t;
  double s, c;
  sincos(static_cast<double>(th) + rot1_hat, &s, &c);
  *x = *x + static_cast<float>(trans_hat * c);
  *y = *y + static_cast<float>(trans_hat * s);
  const float th2 = th + static_cast<float>(rot1_hat + rot2_hat);
  *t = normalise_angle_f(th2);
}

}  // namespace qmcl
