// motion.cu — K5 (particle initialisation) and K6 (odometry motion model).
//
// Replaces ParticleFilter::initialise / global_localization
// (src/quickmcl/particle_filter.cpp:41-83, include/quickmcl/random.h:54-70,
// src/quickmcl/map_base.cpp:58-64) and sample_odometry
// (src/quickmcl/odometry.cpp:44-102).  One thread per particle, one Philox
// block per particle and purpose, indexed by the *global* particle index so
// that results do not depend on sharding or launch shape.  HBM-streaming:
// 24 B read+write per particle for motion, 20 B written for init.

#include "filter.cuh"
#include "motion_dev.cuh"

#include <cmath>

namespace qmcl {

// odometry.cpp:86-101
__global__ void __launch_bounds__(256)
motion_kernel(float *__restrict__ x, float *__restrict__ y, float *__restrict__ t, size_t n,
              size_t global_first, MotionParams mp, uint64_t seed, uint32_t step) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float px = x[i], py = y[i], pt = t[i];
  motion_one(&px, &py, &pt, mp, seed, step, global_first + i);
  x[i] = px;
  y[i] = py;
  t[i] = pt;
}

struct InitParams {
  float mean[3];
  float T[9];  // V * sqrt(Lambda), row-major
};

// particle_filter.cpp:56-68 + random.h:64-70: v = mean + T z with z three
// float standard normals; 3-term products summed as a0 + (a1 + a2) (Eigen's
// unrolled redux).  weight = 1 / particle_count (double).
__global__ void __launch_bounds__(256)
init_gaussian_kernel(float *__restrict__ x, float *__restrict__ y, float *__restrict__ t,
                     double *__restrict__ w, size_t n, size_t global_first, size_t global_n,
                     InitParams ip, uint64_t seed, uint32_t step) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t r[4];
  rng_draw(seed, step, P_INIT, global_first + i, r);
  double d0, d1, d2, d3;
  normal_pair(r[0], r[1], &d0, &d1);
  normal_pair(r[2], r[3], &d2, &d3);
  const float z0 = static_cast<float>(d0), z1 = static_cast<float>(d1),
              z2 = static_cast<float>(d2);
  float v[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const float prod = __fadd_rn(
        __fmul_rn(ip.T[k * 3 + 0], z0),
        __fadd_rn(__fmul_rn(ip.T[k * 3 + 1], z1), __fmul_rn(ip.T[k * 3 + 2], z2)));
    v[k] = __fadd_rn(ip.mean[k], prod);
  }
  x[i] = v[0];
  y[i] = v[1];
  t[i] = normalise_angle_f(v[2]);
  w[i] = __ddiv_rn(1.0, static_cast<double>(global_n));
}

// particle_filter.cpp:71-83 + map_base.cpp:58-64
__global__ void __launch_bounds__(256)
init_global_kernel(float *__restrict__ x, float *__restrict__ y, float *__restrict__ t,
                   double *__restrict__ w, size_t n, size_t global_first, size_t global_n,
                   const uint32_t *__restrict__ free_xy, unsigned long long n_free, MapGeom g,
                   uint64_t seed, uint32_t step) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t r[4];
  rng_draw(seed, step, P_GLOBAL, global_first + i, r);
  float px, py, pt;
  random_free_pose(free_xy, n_free, g, r, &px, &py, &pt);
  x[i] = px;
  y[i] = py;
  t[i] = pt;
  w[i] = __ddiv_rn(1.0, static_cast<double>(global_n));
}

// Symmetric 3x3 eigen-decomposition by cyclic Jacobi in double; factor =
// V * sqrt(Lambda) with eigenvalues ascending as Eigen's SelfAdjointEigenSolver
// orders them (random.h:60-62).  Host side: nine scalars, once per
// /initialpose message.
void covariance_factor(const float cov[9], float out[9]) {
  double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      a[i][j] = 0.5 * (static_cast<double>(cov[i * 3 + j]) + static_cast<double>(cov[j * 3 + i]));
  for (int sweep = 0; sweep < 64; ++sweep) {
    const double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    if (off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        const double theta = (a[q][q] - a[p][p]) / (2 * a[p][q]);
        const double tt =
            (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1));
        const double c = 1 / std::sqrt(tt * tt + 1), s = tt * c;
        for (int k = 0; k < 3; ++k) {
          const double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          const double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          const double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int order[3] = {0, 1, 2};
  for (int i = 0; i < 3; ++i)
    for (int j = i + 1; j < 3; ++j)
      if (a[order[j]][order[j]] < a[order[i]][order[i]]) {
        const int tmp = order[i];
        order[i] = order[j];
        order[j] = tmp;
      }
  for (int col = 0; col < 3; ++col) {
    const int e = order[col];
    const double lam = a[e][e] > 0 ? a[e][e] : 0.0;
    const double sq = std::sqrt(lam);
    for (int row = 0; row < 3; ++row) out[row * 3 + col] = static_cast<float>(v[row][e] * sq);
  }
}

// Slice [first, first+count) of `global_n` particles owned by this shard:
// contiguous blocks in global index order.
// K6 of a handle_odometry call that was left to the next sensor update.
qmcl_status flush_motion(qmcl_filter *f) {
  if (!f->motion_pending) return QMCL_OK;
  f->motion_pending = false;
  ParticleSet &ps = f->cur();
  if (!ps.n) return QMCL_OK;
  QMCL_TRY(use_device(f->map));
  MotionParams mp;
  std::memcpy(&mp, f->pending_motion, sizeof(mp));
  StageScope sc(f, QMCL_STAGE_MOTION);
  motion_kernel<<<div_up(ps.n, 256), 256, 0, f->map->stream>>>(
      ps.x.ptr, ps.y.ptr, ps.t.ptr, ps.n, f->global_first, mp, f->seed, f->pending_motion_step);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

void shard_slice(const qmcl_filter *f, size_t global_n, size_t *first, size_t *count) {
  const size_t per = (global_n + f->n_shards - 1) / f->n_shards;
  size_t a = per * f->shard_rank;
  if (a > global_n) a = global_n;
  size_t b = a + per;
  if (b > global_n) b = global_n;
  *first = a;
  *count = b - a;
}

}  // namespace qmcl

using namespace qmcl;

extern "C" {

qmcl_status qmcl_filter_initialise_factor(qmcl_filter *f, const float pose[3],
                                          const float factor[9]) {
  if (!f || !pose || !factor) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(use_device(f->map));
  QMCL_TRY(settle(f));
  const size_t global_n = static_cast<size_t>(f->params.particle_count_min);
  size_t first, n;
  shard_slice(f, global_n, &first, &n);
  ParticleSet &ps = f->cur();
  QMCL_TRY(ps.ensure(n));
  ps.n = n;
  f->global_first = first;
  f->global_n = global_n;
  f->clusters_valid = false;
  f->est_cached = false;
  f->pose_version++;
  if (n) {
    InitParams ip;
    for (int k = 0; k < 3; ++k) ip.mean[k] = pose[k];
    for (int k = 0; k < 9; ++k) ip.T[k] = factor[k];
    init_gaussian_kernel<<<div_up(n, 256), 256, 0, f->map->stream>>>(
        ps.x.ptr, ps.y.ptr, ps.t.ptr, ps.w.ptr, n, first, global_n, ip, f->seed, f->step);
    QMCL_LAUNCHED();
  }
  f->step++;
  return QMCL_OK;
}

qmcl_status qmcl_filter_initialise(qmcl_filter *f, const float pose[3], const float cov[9]) {
  if (!f || !pose || !cov) return QMCL_ERR_INVALID_ARG;
  float T[9];
  covariance_factor(cov, T);
  return qmcl_filter_initialise_factor(f, pose, T);
}

qmcl_status qmcl_filter_global_localization(qmcl_filter *f) {
  if (!f) return QMCL_ERR_INVALID_ARG;
  if (!f->map->has_map || f->map->n_free == 0) return QMCL_ERR_NO_MAP;
  QMCL_TRY(use_device(f->map));
  QMCL_TRY(settle(f));
  const size_t global_n = static_cast<size_t>(f->params.particle_count_max);
  size_t first, n;
  shard_slice(f, global_n, &first, &n);
  ParticleSet &ps = f->cur();
  QMCL_TRY(ps.ensure(n));
  ps.n = n;
  f->global_first = first;
  f->global_n = global_n;
  f->clusters_valid = false;
  f->est_cached = false;
  f->pose_version++;
  if (n) {
    init_global_kernel<<<div_up(n, 256), 256, 0, f->map->stream>>>(
        ps.x.ptr, ps.y.ptr, ps.t.ptr, ps.w.ptr, n, first, global_n, f->map->free_cells.ptr,
        f->map->n_free, f->map->geom, f->seed, f->step);
    QMCL_LAUNCHED();
  }
  f->step++;
  return QMCL_OK;
}

// odometry.cpp:44-84 on the host (a handful of double scalars per scan, the
// same expressions in the same order), then K6 over all particles.
qmcl_status qmcl_filter_handle_odometry(qmcl_filter *f, const double nw[3], const double od[3],
                                        const float alpha[4]) {
  if (!f || !nw || !od || !alpha) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(use_device(f->map));
  if (f->tail_pending) QMCL_TRY(settle(f));  // the particle count is still in flight
  const double xdiff = nw[0] - od[0];
  const double ydiff = nw[1] - od[1];
  const double delta_trans = std::sqrt(xdiff * xdiff + ydiff * ydiff);
  double delta_rot1 = 0.0;
  if (delta_trans >= 0.01) delta_rot1 = angle_delta(std::atan2(ydiff, xdiff), od[2]);
  const double delta_theta_change = angle_delta(nw[2], od[2]);
  const double delta_rot2 = angle_delta(delta_theta_change, delta_rot1);
  const double r1n =
      std::fmin(std::fabs(delta_rot1), std::fabs(angle_delta(kPi, delta_rot1)));
  const double r2n =
      std::fmin(std::fabs(delta_rot2), std::fabs(angle_delta(kPi, delta_rot2)));
  const double trans_var = alpha[2] * delta_trans * delta_trans + alpha[3] * r1n * r1n +
                           alpha[3] * r2n * r2n;
  const double rot1_var = alpha[0] * r1n * r1n + alpha[1] * delta_trans * delta_trans;
  const double rot2_var = alpha[0] * r2n * r2n + alpha[1] * delta_trans * delta_trans;
  MotionParams mp;
  mp.delta_trans = delta_trans;
  mp.delta_rot1 = delta_rot1;
  mp.delta_rot2 = delta_rot2;
  mp.sd_trans = std::sqrt(trans_var);
  mp.sd_rot1 = std::sqrt(rot1_var);
  mp.sd_rot2 = std::sqrt(rot2_var);
  ParticleSet &ps = f->cur();
  QMCL_TRY(flush_motion(f));  // an earlier motion nobody consumed
  if (ps.n && tail_enabled(f)) {
    // The sensor update that follows applies it in its front kernel (filter.cu);
    // any other reader of the poses makes flush_motion launch it.
    static_assert(sizeof(MotionParams) == sizeof(f->pending_motion), "pending_motion holds a MotionParams");
    std::memcpy(f->pending_motion, &mp, sizeof(mp));
    f->pending_motion_step = f->step;
    f->motion_pending = true;
  } else if (ps.n) {
    StageScope sc(f, QMCL_STAGE_MOTION);
    motion_kernel<<<div_up(ps.n, 256), 256, 0, f->map->stream>>>(
        ps.x.ptr, ps.y.ptr, ps.t.ptr, ps.n, f->global_first, mp, f->seed, f->step);
    QMCL_LAUNCHED();
  }
  f->clusters_valid = false;
  f->est_cached = false;
  f->pose_version++;
  f->step++;
  return QMCL_OK;
}

}  // extern "C"
