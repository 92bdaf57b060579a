// filter.cu — particle filter handle, particle I/O, sensor update + normalise.
//
// Replaces quickmcl::ParticleFilter's container handling and
// update_importance_from_observations (src/quickmcl/particle_filter.cpp:95-136)
// and the weight helpers of src/quickmcl/particle.cpp:41-64.

#include "filter.cuh"
#include "cluster.cuh"
#include "cluster_dev.cuh"
#include "front.cuh"
#include "sensor.cuh"
#include "shard.cuh"
#include "scan.cuh"

#include <cstdlib>

namespace qmcl {

// ------------------------------------------------------------------ kernels

// AoS (the reference's 24-byte WeightedParticle) <-> SoA.
__global__ void aos_to_soa_kernel(const qmcl_particle *__restrict__ in, size_t n,
                                  float *__restrict__ x, float *__restrict__ y,
                                  float *__restrict__ t, double *__restrict__ w) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const qmcl_particle p = in[i];
  x[i] = p.x;
  y[i] = p.y;
  t[i] = p.theta;
  w[i] = p.weight;
}

__global__ void soa_to_aos_kernel(const float *__restrict__ x, const float *__restrict__ y,
                                  const float *__restrict__ t, const double *__restrict__ w,
                                  size_t n, qmcl_particle *__restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  qmcl_particle p;
  p.x = x[i];
  p.y = y[i];
  p.theta = t[i];
  p.pad_ = 0.f;
  p.weight = w[i];
  out[i] = p;
}

// K8: particle.cpp:50-64 as used by particle_filter.cpp:101-120 — divide by the
// pre-penalty total when it is positive, otherwise equalise to 1/N.
__global__ void normalise_by_total_kernel(double *__restrict__ w, size_t n,
                                          const FilterScalars *__restrict__ sc, double n_total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double total = sc->total_weight;
  w[i] = (total > 0) ? __ddiv_rn(w[i], total) : __ddiv_rn(1.0, n_total);
}

qmcl_status launch_aos_to_soa(cudaStream_t s, const qmcl_particle *dev_in, size_t n, float *x,
                              float *y, float *t, double *w) {
  if (!n) return QMCL_OK;
  aos_to_soa_kernel<<<div_up(n, 256), 256, 0, s>>>(dev_in, n, x, y, t, w);
  QMCL_LAUNCHED();
  return QMCL_OK;
}
qmcl_status launch_soa_to_aos(cudaStream_t s, const float *x, const float *y, const float *t,
                              const double *w, size_t n, qmcl_particle *dev_out) {
  if (!n) return QMCL_OK;
  soa_to_aos_kernel<<<div_up(n, 256), 256, 0, s>>>(x, y, t, w, n, dev_out);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

// max over non-negative doubles = max over their bit patterns
__global__ void __launch_bounds__(256)
max_weight_kernel(const double *__restrict__ w, size_t n, unsigned long long *__restrict__ out) {
  unsigned long long m = 0;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const double v = w[i];
    if (v > 0) m = max(m, static_cast<unsigned long long>(__double_as_longlong(v)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// publishing.cpp:45-67: orientation from theta / 2.0f in float, shade =
// float(weight / max_weight) with the division in double.
__global__ void __launch_bounds__(256)
export_markers_kernel(const float *__restrict__ x, const float *__restrict__ y,
                      const float *__restrict__ t, const double *__restrict__ w, size_t n,
                      const unsigned long long *__restrict__ max_bits, float *__restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double mw = __longlong_as_double(static_cast<long long>(*max_bits));
  const float half = __fdiv_rn(t[i], 2.0f);
  float *o = out + 5 * i;
  o[0] = x[i];
  o[1] = y[i];
  o[2] = sinf(half);
  o[3] = cosf(half);
  o[4] = static_cast<float>(__ddiv_rn(w[i], mw));  // 0/0 -> NaN, as on the host
}

qmcl_status launch_export_markers(cudaStream_t s, const float *x, const float *y, const float *t,
                                  const double *w, size_t n_all, size_t n_out,
                                  unsigned long long *d_max, float *d_out) {
  QMCL_CUDA(cudaMemsetAsync(d_max, 0, sizeof(unsigned long long), s));
  const unsigned blocks = static_cast<unsigned>(std::min<size_t>(div_up(n_all, 256), 4 * kSMs));
  max_weight_kernel<<<blocks, 256, 0, s>>>(w, n_all, d_max);
  QMCL_LAUNCHED();
  export_markers_kernel<<<div_up(n_out, 256), 256, 0, s>>>(x, y, t, w, n_out, d_max, d_out);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

qmcl_status filter_scratch(qmcl_filter *f, size_t bytes, void **out) {
  QMCL_TRY(f->scratch.ensure(bytes));
  *out = f->scratch.ptr;
  return QMCL_OK;
}

// Sharded variant: the total is the sum of the shards' totals (known on the
// host), N is the global particle count.
__global__ void normalise_by_value_kernel(double *__restrict__ w, size_t n, double total,
                                          double n_global) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  w[i] = (total > 0) ? __ddiv_rn(w[i], total) : __ddiv_rn(1.0, n_global);
}

// LaserScan -> base-frame points (laser_handler.cpp:108-121 via laser_geometry,
// then the rigid mount transform of tf2 in double).  Valid ranges are compacted
// in order with the device-wide scan.
struct ScanGeom {
  float angle_min, angle_increment, range_min, range_max;
  double mx, my, cy, sy;
};
struct ValidRange {
  const float *ranges;
  ScanGeom g;
  __device__ uint32_t operator()(size_t i) const {
    const double r = ranges[i];
    return (r < static_cast<double>(g.range_max) && r >= static_cast<double>(g.range_min)) ? 1u : 0u;
  }
};
struct EmitPoint {
  const float *ranges;
  ScanGeom g;
  float2 *out;
  __device__ void operator()(size_t i, uint32_t pos, uint32_t flag) const {
    if (!flag) return;
    const double r = ranges[i];
    const double a = static_cast<double>(g.angle_min) +
                     static_cast<double>(i) * static_cast<double>(g.angle_increment);
    double sa, ca;
    sincos(a, &sa, &ca);
    const float px = static_cast<float>(r * ca);
    const float py = static_cast<float>(r * sa);
    const double x = __dadd_rn(__dadd_rn(__dmul_rn(g.cy, px), __dmul_rn(-g.sy, py)), g.mx);
    const double y = __dadd_rn(__dadd_rn(__dmul_rn(g.sy, px), __dmul_rn(g.cy, py)), g.my);
    out[pos] = make_float2(static_cast<float>(x), static_cast<float>(y));
  }
};

// low_pass_filter.h:39-50
static void lowpass_update(double *value, double alpha, double x) {
  if (*value == 0) *value = x;
  else *value = *value + alpha * (x - *value);
}

void apply_sensor_total(qmcl_filter *f, double total, double n_total) {
  f->last_total = total;
  if (total > 0) {
    const double avg = total / n_total;
    lowpass_update(&f->w_slow, f->params.alpha_slow, avg);
    lowpass_update(&f->w_fast, f->params.alpha_fast, avg);
  }
}

void settle_after_sync(qmcl_filter *f) {
  if (f->total_pending) {
    f->total_pending = false;
    apply_sensor_total(f, f->h_scalars.ptr[HS_TOTAL], f->pending_n_total);
  }
  if (f->bins_pending) {
    f->bins_pending = false;
    unsigned long long nb;
    std::memcpy(&nb, &f->h_scalars.ptr[HS_NBINS], sizeof(nb));
    f->n_bins = nb;
  }
  if (f->clusters_pending) {
    f->clusters_pending = false;
    unsigned long long nc;
    std::memcpy(&nc, &f->h_scalars.ptr[HS_NCLUSTERS], sizeof(nc));
    f->n_clusters = nc;
  }
}

qmcl_status settle(qmcl_filter *f) {
  if (f->motion_pending && !f->tail_pending) QMCL_TRY(flush_motion(f));
  if (f->normalise_pending && !f->tail_pending) QMCL_TRY(flush_sensor(f));
  if (!f->total_pending && !f->clusters_pending && !f->bins_pending && !f->tail_pending) return QMCL_OK;
  QMCL_TRY(use_device(f->map));
  QMCL_CUDA(cudaStreamSynchronize(f->map->stream));
  settle_after_sync(f);
  // the result block of a fused tail; a fall-back re-runs the call here
  if (f->tail_pending) QMCL_TRY(tail_complete(f));
  return QMCL_OK;
}

// K8 and the total's read-back, when sensor_update_common left them to a fused
// tail that did not come (a getter, a multi-launch resampling, a second sensor
// update): divide by the pre-penalty total, copy the total for w_slow / w_fast.
qmcl_status flush_sensor(qmcl_filter *f) {
  if (!f->normalise_pending) return QMCL_OK;
  f->normalise_pending = false;
  QMCL_TRY(use_device(f->map));
  cudaStream_t s = f->map->stream;
  ParticleSet &ps = f->cur();
  if (ps.n) {
    StageScope sc(f, QMCL_STAGE_NORMALISE);
    normalise_by_total_kernel<<<div_up(ps.n, 256), 256, 0, s>>>(ps.w.ptr, ps.n, f->scalars.ptr,
                                                                static_cast<double>(ps.n));
    QMCL_LAUNCHED();
  }
  QMCL_CUDA(copy_d2h(f->h_scalars.ptr + HS_TOTAL, &f->scalars.ptr->total_weight, sizeof(double), s));
  if (!f->total_copied)
    QMCL_CUDA(cudaEventCreateWithFlags(&f->total_copied, cudaEventDisableTiming));
  QMCL_CUDA(cudaEventRecord(f->total_copied, s));
  f->total_pending = true;
  f->pending_n_total = static_cast<double>(f->normalise_n);
  return QMCL_OK;
}

// The pinned scan staging buffer may be rewritten once the previous H2D copy
// out of it has run (update_sensor no longer ends with a synchronisation).
static qmcl_status staging_free(qmcl_filter *f) {
  if (!f->staging_read) {
    QMCL_CUDA(cudaEventCreateWithFlags(&f->staging_read, cudaEventDisableTiming));
    return QMCL_OK;
  }
  QMCL_CUDA(cudaEventSynchronize(f->staging_read));
  return QMCL_OK;
}

// The front kernel of a sensor update on the fused path: one launch for
//   block 0      the beam selection of map_likelihood.cpp:67-74,92 (subsample_kernel's
//                job) and the bound of the longest used beam;
//   blocks 1..   the deferred motion model (K6, odometry.cpp:86-101) over the particles,
//                grid-stride, and the bin bounds of the moved poses per block; the last
//                block to finish folds them into bbox[6] for the tail.
// Replaces motion_kernel + bbox_init_kernel + bin_bbox_kernel + a memset +
// subsample_kernel (five stream operations) by one.
// Block `bx` of the 1 + nb blocks working on job J.
__device__ __forceinline__ void front_body(const FrontJob &J, const unsigned bx, const unsigned nb) {
  __shared__ float s_len[8];
  __shared__ int s_mn[8][3], s_mx[8][3];
  __shared__ bool s_last;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (bx == 0) {
    float len = 0.f;
    for (int j = threadIdx.x; j < J.n_used; j += 256) {
      float2 p = J.cloud[static_cast<size_t>(j) * J.step];
      // x86 float->int of a NaN/huge value yields INT_MIN, i.e. "outside the map"
      if (!(fabsf(p.x) < 1e30f) || !(fabsf(p.y) < 1e30f)) p = make_float2(1e30f, 1e30f);
      J.beams[j] = p;
      len = fmaxf(len, __fadd_ru(fabsf(p.x), fabsf(p.y)));  // |p|_1 >= |p|_2, rounded up
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) len = fmaxf(len, __shfl_xor_sync(0xffffffffu, len, o));
    if (lane == 0) s_len[wid] = len;
    __syncthreads();
    if (threadIdx.x == 0) {
      float m = 0.f;
      for (int k = 0; k < 8; ++k) m = fmaxf(m, s_len[k]);
      *J.rmax = __float_as_uint(m);
    }
    return;
  }
  const unsigned b = bx - 1;
  // L2 prefetch of the sensor kernel's map (one request per 128-byte line, spread over the grid)
  for (size_t line = static_cast<size_t>(b) * 256 + threadIdx.x; line * 128 < J.prefetch_bytes;
       line += static_cast<size_t>(nb) * 256)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(J.prefetch + line * 128));
  int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
  for (size_t i = static_cast<size_t>(b) * 256 + threadIdx.x; i < J.n; i += static_cast<size_t>(nb) * 256) {
    float px = J.x[i], py = J.y[i], pt = J.t[i];
    if (J.do_motion) {
      motion_one(&px, &py, &pt, J.mp, J.seed, J.rng_step, J.global_first + i);
      J.x[i] = px;
      J.y[i] = py;
      J.t[i] = pt;
    }
    int k[3];
    bin_key(px, py, pt, J.bp, &k[0], &k[1], &k[2]);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      mn[d] = min(mn[d], k[d]);
      mx[d] = max(mx[d], k[d]);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = min(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = max(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      s_mn[wid][d] = mn[d];
      s_mx[wid][d] = mx[d];
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    int lo = INT_MAX, hi = INT_MIN;
    for (int k = 0; k < 8; ++k) {
      lo = min(lo, s_mn[k][threadIdx.x]);
      hi = max(hi, s_mx[k][threadIdx.x]);
    }
    J.block_bbox[b * 6 + threadIdx.x] = lo;
    J.block_bbox[b * 6 + 3 + threadIdx.x] = hi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(J.blocks_done, 1u) == nb - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  // the last block: bounds over all blocks
  if (threadIdx.x < 6) {
    const bool is_min = threadIdx.x < 3;
    int v = is_min ? INT_MAX : INT_MIN;
    for (unsigned k = 0; k < nb; ++k) {
      const int e = __ldcg(J.block_bbox + k * 6 + threadIdx.x);
      v = is_min ? min(v, e) : max(v, e);
    }
    J.bbox[threadIdx.x] = v;
  }
  if (threadIdx.x == 0) *J.blocks_done = 0;
}

__global__ void __launch_bounds__(256) front_kernel(const __grid_constant__ FrontJob J) {
  front_body(J, blockIdx.x, gridDim.x - 1);
}
// Many filters in one launch: blockIdx.y selects the job.
__global__ void __launch_bounds__(256) front_batch_kernel(const FrontJob *__restrict__ jobs) {
  __shared__ FrontJob s_job;
  for (unsigned k = threadIdx.x; k < sizeof(FrontJob) / 8; k += blockDim.x)
    reinterpret_cast<uint64_t *>(&s_job)[k] = reinterpret_cast<const uint64_t *>(jobs + blockIdx.y)[k];
  __syncthreads();
  const unsigned nb = static_cast<unsigned>(min(static_cast<size_t>((s_job.n + 255) / 256),
                                                static_cast<size_t>(4 * kSMs)));
  if (blockIdx.x > nb) return;
  front_body(s_job, blockIdx.x, nb);
}

// Fills the front job of filter f for the cloud at dev_cloud_xy and does the
// host-side bookkeeping of its launch (the caller launches it, alone or in a batch).
qmcl_status front_prepare(qmcl_filter *f, const float *dev_cloud_xy, size_t n_points, FrontJob *Jp,
                          int *n_used_out) {
  qmcl_map *m = f->map;
  ParticleSet &ps = f->cur();
  // map_likelihood.cpp:67-74
  size_t step = 1;
  if (m->params.num_beams) step = n_points / static_cast<size_t>(m->params.num_beams);
  if (step < 1) step = 1;
  const size_t n_used = n_points ? (n_points + step - 1) / step : 0;
  *n_used_out = static_cast<int>(n_used);
  QMCL_TRY(f->bbox.ensure(8 + 6 * static_cast<size_t>(4 * kSMs)));
  FrontJob &J = *Jp;
  std::memset(&J, 0, sizeof(J));
  J.cloud = reinterpret_cast<const float2 *>(dev_cloud_xy);
  J.n_points = n_points;
  J.step = step;
  J.beams = reinterpret_cast<float2 *>(f->beams.ptr);
  J.n_used = static_cast<int>(n_used);
  J.rmax = reinterpret_cast<unsigned *>(f->beams.ptr + 2 * n_points);
  J.x = ps.x.ptr;
  J.y = ps.y.ptr;
  J.t = ps.t.ptr;
  J.n = ps.n;
  J.global_first = f->global_first;
  J.do_motion = f->motion_pending ? 1 : 0;
  std::memcpy(&J.mp, f->pending_motion, sizeof(J.mp));
  J.seed = f->seed;
  J.rng_step = f->pending_motion_step;
  J.bp.res_xy = f->params.res_xy;
  J.bp.res_t = f->params.res_theta;
  J.bp.theta_lo = f->theta_lo;
  J.bp.theta_hi = f->theta_hi;
  J.block_bbox = f->bbox.ptr + 8;
  J.bbox = f->bbox.ptr;
  J.blocks_done = &f->scalars.ptr->blocks_done;
  // the palette map when it is the sensor path in use and fits L2 comfortably
  if (m->pal8_valid && m->pal8_n > 0 && (m->sensor_path == 0 || m->sensor_path == 3)) {
    const size_t bytes = code8_map_elems(m->geom.w, m->geom.h);
    if (bytes <= (size_t(32) << 20)) {
      J.prefetch = m->code8.ptr;
      J.prefetch_bytes = bytes;
    }
  }
  f->motion_pending = false;
  f->early_bbox_version = f->pose_version;
  return QMCL_OK;
}

unsigned front_blocks(size_t n_particles) {
  return static_cast<unsigned>(std::min<size_t>(div_up(n_particles, 256), 4 * kSMs));
}

static qmcl_status launch_front(qmcl_filter *f, const float *dev_cloud_xy, size_t n_points, int *n_used_out) {
  FrontJob J;
  QMCL_TRY(front_prepare(f, dev_cloud_xy, n_points, &J, n_used_out));
  StageScope sc(f, QMCL_STAGE_MOTION, 1);
  front_kernel<<<front_blocks(J.n) + 1, 256, 0, f->map->stream>>>(J);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

qmcl_status launch_front_batch(qmcl_map *m, const FrontJob *dev_jobs, unsigned n_jobs, size_t max_particles) {
  if (!n_jobs) return QMCL_OK;
  front_batch_kernel<<<dim3(front_blocks(max_particles) + 1, n_jobs), 256, 0, m->stream>>>(dev_jobs);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

// The sensor job of filter f after front_prepare (fused path only): K7's
// arguments, and the bookkeeping of a launch whose K8 is left to the tail.
qmcl_status sensor_prepare_deferred(qmcl_filter *f, int n_used, size_t n_points, SensorLaunch *a) {
  ParticleSet &ps = f->cur();
  QMCL_TRY(f->partials.ensure(sensor_grid(ps.n) + 1));
  a->x = ps.x.ptr;
  a->y = ps.y.ptr;
  a->t = ps.t.ptr;
  a->w = ps.w.ptr;
  a->n = ps.n;
  a->beams = f->beams.ptr;
  a->n_beams = n_used;
  a->rmax = f->beams.ptr + 2 * n_points;
  a->partials = f->partials.ptr;
  a->scalars = f->scalars.ptr;
  f->clusters_valid = false;
  f->est_cached = false;
  f->normalise_pending = true;
  f->normalise_n = ps.n;
  return QMCL_OK;
}

// Runs the beam selection, K7 and K8 on the current set for the cloud at dev_cloud_xy.
static qmcl_status sensor_update_common(qmcl_filter *f, const float *dev_cloud_xy, size_t n_points) {
  qmcl_map *m = f->map;
  ParticleSet &ps = f->cur();
  cudaStream_t s = m->stream;
  // a pending total's slot is about to be reused; a deferred normalisation or a
  // fused tail still in flight belong to the previous update.  A deferred motion
  // is ours to apply (front kernel).
  if (f->tail_pending || f->normalise_pending || f->total_pending || f->bins_pending || f->clusters_pending) {
    const bool motion = f->motion_pending;
    f->motion_pending = false;
    const qmcl_status st = settle(f);
    f->motion_pending = motion;
    QMCL_TRY(st);
  }
  // With the fused tail (tail.cu) K8 and everything after it run in one launch
  // that reads the total and the bin bounds on the device.
  // The front kernel (deferred motion + beam selection + bin bounds) serves every
  // single-GPU filter in lazy mode; K8 is left to the tail only for sets it will take.
  const bool front = tail_enabled(f) && ps.n > 0;
  const bool defer = front && ps.n <= kTailMaxParticles;
  int n_used = 0;
  if (front) {
    QMCL_TRY(launch_front(f, dev_cloud_xy, n_points, &n_used));
  } else {
    QMCL_TRY(flush_motion(f));
    QMCL_TRY(launch_subsample(f->map, dev_cloud_xy, n_points, f->beams.ptr, &n_used));
  }
  if (ps.n == 0 && !sharded(f)) {
    f->last_total = 0.0;
    return QMCL_OK;
  }
  QMCL_TRY(f->partials.ensure(sensor_grid(ps.n) + 1));
  SensorLaunch a;
  a.x = ps.x.ptr;
  a.y = ps.y.ptr;
  a.t = ps.t.ptr;
  a.w = ps.w.ptr;
  a.n = ps.n;
  a.beams = f->beams.ptr;
  a.n_beams = n_used;
  a.rmax = f->beams.ptr + 2 * n_points;
  a.partials = f->partials.ptr;
  a.scalars = f->scalars.ptr;
  if (!defer && ps.n && lazy(f) && f->params.resample_type == QMCL_RESAMPLE_KLD) {
    // KLD sizes its dense bin table by the bounding box of the set it draws
    // from.  Taken here, ahead of the sensor kernel, the six numbers are on the
    // host long before the resampling asks for them.
    StageScope sc(f, QMCL_STAGE_BINS);
    if (!front) QMCL_TRY(launch_bin_bbox(f, ps.x.ptr, ps.y.ptr, ps.t.ptr, ps.n));
    QMCL_CUDA(copy_d2h(f->h_scalars.ptr + HS_BBOX_EARLY, f->bbox.ptr, 6 * sizeof(int), s));
    f->early_bbox_version = f->pose_version;
  }
  if (ps.n) {
    StageScope sc(f, QMCL_STAGE_SENSOR, 1);
    QMCL_TRY(launch_sensor(m, a));
  } else {
    QMCL_CUDA(cudaMemsetAsync(&f->scalars.ptr->total_weight, 0, sizeof(double), s));
  }
  f->clusters_valid = false;
  f->est_cached = false;
  if (defer) {
    f->normalise_pending = true;
    f->normalise_n = ps.n;
    return QMCL_OK;
  }
  const double n_total = static_cast<double>(sharded(f) ? f->global_n : ps.n);
  if (sharded(f)) {
    // Exchange 1: the pre-penalty total is the sum over the shards, taken
    // in rank order on the device; the host picks it up like a single-GPU filter does.
    QMCL_TRY(shard_sum_f64_device(f, &f->scalars.ptr->total_weight));
  }
  if (ps.n) {
    StageScope sc(f, QMCL_STAGE_NORMALISE);
    normalise_by_total_kernel<<<div_up(ps.n, 256), 256, 0, s>>>(ps.w.ptr, ps.n, f->scalars.ptr, n_total);
    QMCL_LAUNCHED();
  }
  // The host keeps the two low-pass scalars (w_slow, w_fast), so it needs the
  // total: one 8-byte read back per scan.  Nothing on the host depends on it
  // before the next resampling, so the copy is left in flight (HS_TOTAL) and
  // folded in at the next synchronisation.
  QMCL_CUDA(copy_d2h(f->h_scalars.ptr + HS_TOTAL, &f->scalars.ptr->total_weight, sizeof(double), s));
  if (!f->total_copied)
    QMCL_CUDA(cudaEventCreateWithFlags(&f->total_copied, cudaEventDisableTiming));
  QMCL_CUDA(cudaEventRecord(f->total_copied, s));
  f->total_pending = true;
  f->pending_n_total = n_total;
  if (!defer_counts(f)) QMCL_TRY(settle(f));
  return QMCL_OK;
}

}  // namespace qmcl

using namespace qmcl;

extern "C" {

qmcl_status qmcl_filter_create(qmcl_map *map, uint64_t seed, qmcl_filter **out) {
  if (!map || !out) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(use_device(map));
  *out = nullptr;
  qmcl_filter *f = new qmcl_filter();
  f->map = map;
  f->seed = seed;
  // the fallible steps come before the map reference is taken, and their
  // buffers are released again on failure
  auto init = [&]() -> qmcl_status {
    QMCL_TRY(f->scalars.ensure(1));
    QMCL_CUDA(cudaMemsetAsync(f->scalars.ptr, 0, sizeof(FilterScalars), map->stream));
    QMCL_TRY(f->h_scalars.ensure(HS_COUNT));
    return QMCL_OK;
  };
  const qmcl_status st = init();
  if (st != QMCL_OK) {
    f->scalars.release();
    f->h_scalars.release();
    delete f;
    return st;
  }
  map->refs++;
  if (const char *e = std::getenv("QMCL_EAGER_SYNC")) f->lazy_sync = !(e[0] && e[0] != '0');
  if (const char *e = std::getenv("QMCL_TAIL")) f->use_tail = !(e[0] == '0');
  if (const char *e = std::getenv("QMCL_PREFIX_MODE")) f->prefix_mode = (e[0] == '2') ? 2 : (e[0] == '1');
  qmcl_filter_set_params(f, &f->params);
  *out = f;
  return QMCL_OK;
}

qmcl_status qmcl_filter_destroy(qmcl_filter *f) {
  if (!f) return QMCL_ERR_INVALID_ARG;
  cudaSetDevice(f->map->device);
  cudaStreamSynchronize(f->map->stream);
  for (auto &s : f->sets) {
    s.x.release();
    s.y.release();
    s.t.release();
    s.w.release();
  }
  f->beams.release();
  f->cloud.release();
  f->ranges.release();
  f->partials.release();
  f->scalars.release();
  f->h_beams.release();
  f->h_scalars.release();
  f->cdf.release();
  f->pos.release();
  f->incl.release();
  f->src_index.release();
  f->scratch.release();
  f->prefix_scratch.release();
  f->prefix_carry.release();
  f->xchg.release();
  f->xchg_send.release();
  f->xchg_recv.release();
  f->bbox.release();
  f->table.release();
  f->bin_cell.release();
  f->bin_count.release();
  f->occ8.release();
  f->bin_label.release();
  f->bin_cluster.release();
  f->scan_tmp.release();
  f->scan_legacy.release();
  f->cluster_acc.release();
  f->particle_cid.release();
  f->timer.release();
  f->tail_dev.release();
  f->tail_scratch.release();
  f->tail_first.release();
  f->tail_cluster_w.release();
  if (f->h_tail) cudaFreeHost(f->h_tail);
  if (f->staging_read) cudaEventDestroy(f->staging_read);
  if (f->total_copied) cudaEventDestroy(f->total_copied);
  qmcl_map *map = f->map;
  delete f;
  return qmcl_map_destroy(map);
}

// particle_filter.cpp:177-184 + space_partitioning.cpp:23-37
qmcl_status qmcl_filter_set_params(qmcl_filter *f, const qmcl_pf_params *p) {
  if (!f || !p) return QMCL_ERR_INVALID_ARG;
  if (!(p->res_xy > 0) || !(p->res_theta > 0)) return QMCL_ERR_INVALID_ARG;
  if (f->map) QMCL_TRY(settle(f));  // a pending total uses the alphas it was measured under
  f->params = *p;
  f->early_bbox_version = ~0ull;  // bin resolutions may have changed
  // theta_range: world_to_map of (float)(-pi+0.001) and (float)(pi-0.001):
  // float division then truncation (scaled_map.h:43-48).
  f->theta_lo = static_cast<int>(static_cast<float>(-kPi + 0.001) / p->res_theta);
  f->theta_hi = static_cast<int>(static_cast<float>(kPi - 0.001) / p->res_theta);
  f->n_bins = 0;
  f->bins_pending = false;
  f->n_clusters = 0;
  f->clusters_pending = false;  // an older count still in flight is void
  f->clusters_valid = false;
  return QMCL_OK;
}

qmcl_status qmcl_filter_set_rng(qmcl_filter *f, uint64_t seed, uint32_t step) {
  if (!f) return QMCL_ERR_INVALID_ARG;
  f->seed = seed;
  f->step = step;
  return QMCL_OK;
}

qmcl_status qmcl_filter_num_particles(const qmcl_filter *f, size_t *out) {
  if (!f || !out) return QMCL_ERR_INVALID_ARG;
  if (f->tail_pending) QMCL_TRY(settle(const_cast<qmcl_filter *>(f)));  // the count is still in flight
  *out = sharded(f) ? f->global_n : f->cur().n;
  return QMCL_OK;
}

qmcl_status qmcl_filter_set_particles_shard(qmcl_filter *f, const qmcl_particle *p, size_t n,
                                            uint64_t global_first, uint64_t global_count) {
  if (!f || global_first + n > global_count) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(qmcl_filter_set_particles(f, p, n));
  f->global_first = global_first;
  f->global_n = global_count;
  return QMCL_OK;
}

qmcl_status qmcl_filter_set_particles(qmcl_filter *f, const qmcl_particle *p, size_t n) {
  if (!f || (!p && n)) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(use_device(f->map));
  QMCL_TRY(settle(f));
  cudaStream_t s = f->map->stream;
  ParticleSet &ps = f->cur();
  QMCL_TRY(ps.ensure(n));
  ps.n = n;
  f->global_first = 0;
  f->global_n = n;
  f->clusters_valid = false;
  f->est_cached = false;
  f->pose_version++;
  if (!n) return QMCL_OK;
  void *stage = nullptr;
  QMCL_TRY(filter_scratch(f, n * sizeof(qmcl_particle), &stage));
  QMCL_CUDA(copy_h2d(stage, p, n * sizeof(qmcl_particle), s));
  QMCL_TRY(launch_aos_to_soa(s, static_cast<const qmcl_particle *>(stage), n, ps.x.ptr, ps.y.ptr,
                             ps.t.ptr, ps.w.ptr));
  QMCL_CUDA(cudaStreamSynchronize(s));
  return QMCL_OK;
}

qmcl_status qmcl_filter_copy_particles(const qmcl_filter *fc, qmcl_particle *out, size_t cap) {
  if (!fc || (!out && cap)) return QMCL_ERR_INVALID_ARG;
  qmcl_filter *f = const_cast<qmcl_filter *>(fc);
  QMCL_TRY(use_device(f->map));
  QMCL_TRY(settle(f));
  cudaStream_t s = f->map->stream;
  const ParticleSet &ps = f->cur();
  const size_t n = ps.n < cap ? ps.n : cap;
  if (!n) return QMCL_OK;
  void *stage = nullptr;
  QMCL_TRY(filter_scratch(f, n * sizeof(qmcl_particle), &stage));
  QMCL_TRY(launch_soa_to_aos(s, ps.x.ptr, ps.y.ptr, ps.t.ptr, ps.w.ptr, n,
                             static_cast<qmcl_particle *>(stage)));
  QMCL_CUDA(copy_d2h(out, stage, n * sizeof(qmcl_particle), s));
  QMCL_CUDA(cudaStreamSynchronize(s));
  return QMCL_OK;
}

// Particle cloud in the form the node publishes it (publishing.cpp:45-67,
// 114-154): per particle x, y, sin(theta/2), cos(theta/2) — the yaw quaternion,
// float arithmetic as in to_marker — and weight / max_weight as a float shade.
// 20 bytes per particle leave the device instead of the 24-byte mirror plus a
// host pass for the maximum.
qmcl_status qmcl_filter_export_markers(const qmcl_filter *fc, float *out5, size_t cap,
                                       size_t *n_out, double *max_weight) {
  if (!fc || (!out5 && cap)) return QMCL_ERR_INVALID_ARG;
  qmcl_filter *f = const_cast<qmcl_filter *>(fc);
  if (sharded(f)) {
    set_error("export_markers: the maximum weight spans all shards; gather the particles instead");
    return QMCL_ERR_UNSUPPORTED;
  }
  QMCL_TRY(use_device(f->map));
  QMCL_TRY(settle(f));
  cudaStream_t s = f->map->stream;
  const ParticleSet &ps = f->cur();
  const size_t n = ps.n < cap ? ps.n : cap;
  if (n_out) *n_out = n;
  if (max_weight) *max_weight = 0.0;
  if (!n) return QMCL_OK;
  void *stage = nullptr;
  QMCL_TRY(filter_scratch(f, n * 5 * sizeof(float) + 16, &stage));
  // the maximum runs over the whole set (publishing.cpp:125-131), also when
  // fewer particles are copied out
  unsigned long long *d_max = reinterpret_cast<unsigned long long *>(stage);
  float *d_out = reinterpret_cast<float *>(static_cast<char *>(stage) + 16);
  QMCL_TRY(launch_export_markers(s, ps.x.ptr, ps.y.ptr, ps.t.ptr, ps.w.ptr, ps.n, n, d_max, d_out));
  QMCL_CUDA(copy_d2h(out5, d_out, n * 5 * sizeof(float), s));
  double mw = 0.0;
  QMCL_CUDA(copy_d2h(&mw, d_max, sizeof(double), s));
  QMCL_CUDA(cudaStreamSynchronize(s));
  if (max_weight) *max_weight = mw;
  return QMCL_OK;
}

qmcl_status qmcl_filter_update_sensor(qmcl_filter *f, const float *cloud_xy, size_t n_points) {
  if (!f || (!cloud_xy && n_points)) return QMCL_ERR_INVALID_ARG;
  if (!f->map->has_map) return QMCL_ERR_NO_MAP;
  QMCL_TRY(use_device(f->map));
  cudaStream_t s = f->map->stream;
  QMCL_TRY(staging_free(f));  // before the staging buffer may be reallocated, too
  QMCL_TRY(f->cloud.ensure(2 * n_points + 2));
  QMCL_TRY(f->beams.ensure(2 * n_points + 2));
  QMCL_TRY(f->h_beams.ensure(2 * n_points + 2));
  // stage through pinned memory so the H2D copy is a true async DMA
  std::memcpy(f->h_beams.ptr, cloud_xy, 2 * n_points * sizeof(float));
  QMCL_CUDA(copy_h2d(f->cloud.ptr, f->h_beams.ptr, 2 * n_points * sizeof(float), s));
  QMCL_CUDA(cudaEventRecord(f->staging_read, s));
  return sensor_update_common(f, f->cloud.ptr, n_points);
}

qmcl_status qmcl_filter_update_sensor_dev(qmcl_filter *f, const float *dev_cloud_xy,
                                          size_t n_points) {
  if (!f || (!dev_cloud_xy && n_points)) return QMCL_ERR_INVALID_ARG;
  if (!f->map->has_map) return QMCL_ERR_NO_MAP;
  QMCL_TRY(use_device(f->map));
  QMCL_TRY(f->beams.ensure(2 * n_points + 2));
  return sensor_update_common(f, dev_cloud_xy, n_points);
}

qmcl_status qmcl_filter_update_sensor_scan(qmcl_filter *f, const float *ranges, size_t n_ranges,
                                           float angle_min, float angle_increment,
                                           float range_min, float range_max,
                                           const double mount[3], size_t *n_points_out) {
  if (!f || (!ranges && n_ranges) || !mount) return QMCL_ERR_INVALID_ARG;
  if (!f->map->has_map) return QMCL_ERR_NO_MAP;
  QMCL_TRY(use_device(f->map));
  cudaStream_t s = f->map->stream;
  QMCL_TRY(staging_free(f));  // before the staging buffer may be reallocated, too
  QMCL_TRY(f->ranges.ensure(n_ranges + 2));
  QMCL_TRY(f->cloud.ensure(2 * n_ranges + 2));
  QMCL_TRY(f->beams.ensure(2 * n_ranges + 2));
  QMCL_TRY(f->h_beams.ensure(2 * n_ranges + 2));
  std::memcpy(f->h_beams.ptr, ranges, n_ranges * sizeof(float));
  QMCL_CUDA(copy_h2d(f->ranges.ptr, f->h_beams.ptr, n_ranges * sizeof(float), s));
  QMCL_CUDA(cudaEventRecord(f->staging_read, s));
  ScanGeom g{angle_min, angle_increment, range_min, range_max, mount[0], mount[1],
             std::cos(mount[2]), std::sin(mount[2])};
  QMCL_TRY(f->scan_tmp.ensure_zero(scan_tmp_words(n_ranges) + 4));
  unsigned long long *d_total = scan_total_ptr(f->scan_tmp.ptr, n_ranges);
  QMCL_TRY(device_scan(s, ValidRange{f->ranges.ptr, g},
                       EmitPoint{f->ranges.ptr, g, reinterpret_cast<float2 *>(f->cloud.ptr)},
                       n_ranges, f->scan_tmp.ptr, d_total));
  unsigned long long n_points = 0;  // the subsampling step depends on it (map_likelihood.cpp:67-74)
  QMCL_CUDA(copy_d2h(&n_points, d_total, sizeof(n_points), s));
  QMCL_CUDA(cudaStreamSynchronize(s));
  f->n_cloud = n_points;
  if (n_points_out) *n_points_out = n_points;
  return sensor_update_common(f, f->cloud.ptr, n_points);
}

qmcl_status qmcl_filter_copy_cloud(const qmcl_filter *fc, float *xy_out, size_t cap, size_t *n_out) {
  if (!fc) return QMCL_ERR_INVALID_ARG;
  qmcl_filter *f = const_cast<qmcl_filter *>(fc);
  if (n_out) *n_out = f->n_cloud;
  const size_t n = f->n_cloud < cap ? f->n_cloud : cap;
  if (!n || !xy_out) return QMCL_OK;
  QMCL_TRY(use_device(f->map));
  QMCL_CUDA(copy_d2h(xy_out, f->cloud.ptr, 2 * n * sizeof(float), f->map->stream));
  QMCL_CUDA(cudaStreamSynchronize(f->map->stream));
  return QMCL_OK;
}

qmcl_status qmcl_filter_last_total_weight(const qmcl_filter *f, double *out) {
  if (!f || !out) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(settle(const_cast<qmcl_filter *>(f)));
  *out = f->last_total;
  return QMCL_OK;
}

qmcl_status qmcl_filter_get_lowpass(const qmcl_filter *f, double out[2]) {
  if (!f || !out) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(settle(const_cast<qmcl_filter *>(f)));
  out[0] = f->w_slow;
  out[1] = f->w_fast;
  return QMCL_OK;
}
qmcl_status qmcl_filter_set_lowpass(qmcl_filter *f, double w_slow, double w_fast) {
  if (!f) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(settle(f));
  f->w_slow = w_slow;
  f->w_fast = w_fast;
  return QMCL_OK;
}

qmcl_status qmcl_filter_set_sync_mode(qmcl_filter *f, int lazy_mode) {
  if (!f || lazy_mode < 0 || lazy_mode > 1) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(settle(f));
  f->lazy_sync = lazy_mode != 0;
  return QMCL_OK;
}

qmcl_status qmcl_filter_set_timing(qmcl_filter *f, int level) {
  if (!f || level < 0 || level > 2) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(settle(f));  // level 2 switches the fused tail off
  f->timer.level = level;
  return QMCL_OK;
}

qmcl_status qmcl_filter_get_timing(qmcl_filter *f, double ms_out[QMCL_N_STAGES],
                                   uint64_t count_out[QMCL_N_STAGES], int reset) {
  if (!f || !ms_out || !count_out) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(use_device(f->map));
  QMCL_TRY(settle(f));
  QMCL_CUDA(cudaStreamSynchronize(f->map->stream));
  f->timer.resolve();
  for (int i = 0; i < QMCL_N_STAGES; ++i) {
    ms_out[i] = f->timer.ms[i];
    count_out[i] = f->timer.count[i];
    if (reset) {
      f->timer.ms[i] = 0.0;
      f->timer.count[i] = 0;
    }
  }
  return QMCL_OK;
}

qmcl_status qmcl_filter_synchronize(qmcl_filter *f) {
  if (!f) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(use_device(f->map));
  QMCL_TRY(settle(f));
  QMCL_CUDA(cudaStreamSynchronize(f->map->stream));
  return QMCL_OK;
}

// Map::update_importance on a host vector (map_likelihood.cpp:63-130): the
// reference's own Map interface, kept for API completeness.  H2D, K7, D2H.
qmcl_status qmcl_map_update_importance(qmcl_map *m, const float *cloud_xy, size_t n_points,
                                       qmcl_particle *particles, size_t n, double *total_out) {
  if (!m || (!cloud_xy && n_points) || (!particles && n)) return QMCL_ERR_INVALID_ARG;
  if (!m->has_map) return QMCL_ERR_NO_MAP;
  QMCL_TRY(use_device(m));
  cudaStream_t s = m->stream;
  if (total_out) *total_out = 0.0;
  if (!n) return QMCL_OK;
  QMCL_TRY(m->tmp_x.ensure(n));
  QMCL_TRY(m->tmp_y.ensure(n));
  QMCL_TRY(m->tmp_t.ensure(n));
  QMCL_TRY(m->tmp_w.ensure(n + sensor_grid(n) + 8));
  QMCL_TRY(m->tmp_aos.ensure(n * sizeof(qmcl_particle)));
  QMCL_TRY(m->tmp_cloud.ensure(4 * n_points + 4));
  QMCL_TRY(m->tmp_scalars.ensure(1));
  qmcl_particle *d_aos = reinterpret_cast<qmcl_particle *>(m->tmp_aos.ptr);
  QMCL_CUDA(cudaMemsetAsync(m->tmp_scalars.ptr, 0, sizeof(FilterScalars), s));
  QMCL_CUDA(copy_h2d(d_aos, particles, n * sizeof(qmcl_particle), s));
  QMCL_CUDA(copy_h2d(m->tmp_cloud.ptr, cloud_xy, 2 * n_points * sizeof(float), s));
  QMCL_TRY(launch_aos_to_soa(s, d_aos, n, m->tmp_x.ptr, m->tmp_y.ptr, m->tmp_t.ptr, m->tmp_w.ptr));
  float *d_beams = m->tmp_cloud.ptr + 2 * n_points + 2;
  int n_used = 0;
  QMCL_TRY(launch_subsample(m, m->tmp_cloud.ptr, n_points, d_beams, &n_used));
  SensorLaunch a;
  a.x = m->tmp_x.ptr;
  a.y = m->tmp_y.ptr;
  a.t = m->tmp_t.ptr;
  a.w = m->tmp_w.ptr;
  a.n = n;
  a.beams = d_beams;
  a.n_beams = n_used;
  a.rmax = d_beams + 2 * n_points;
  a.partials = m->tmp_w.ptr + n;
  a.scalars = m->tmp_scalars.ptr;
  QMCL_TRY(launch_sensor(m, a));
  QMCL_TRY(launch_soa_to_aos(s, a.x, a.y, a.t, a.w, n, d_aos));
  QMCL_CUDA(copy_d2h(particles, d_aos, n * sizeof(qmcl_particle), s));
  double total = 0.0;
  QMCL_CUDA(copy_d2h(&total, &m->tmp_scalars.ptr->total_weight, sizeof(double), s));
  QMCL_CUDA(cudaStreamSynchronize(s));
  if (total_out) *total_out = total;
  return QMCL_OK;
}

}  // extern "C"
