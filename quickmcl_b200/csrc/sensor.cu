// sensor.cu — K7: likelihood-field sensor model.
//
// Replaces MapLikelihood::update_importance / get_probability_at
// (src/quickmcl/map_likelihood.cpp:63-130, :148-160) and the state lookup
// MapBase::get_map_state_at (map_base.cpp:66-77).
//
// Work decomposition: one warp owns kP particles and walks the scan with its
// 32 lanes on 32 consecutive beams.  Consecutive beams of one scan end on
// neighbouring map cells, so the 32 field gathers of one warp instruction fall
// into a handful of 32-byte sectors instead of 32 (a thread-per-particle
// mapping would touch 32 unrelated lines per instruction).  Each lane keeps kP
// independent gathers in flight; per-particle sums are reduced with warp
// shuffles in a fixed order, so results do not depend on the launch shape.
//
// The index path repeats the reference's float expression operation by
// operation (no FMA contraction, IEEE division, truncation toward zero):
//   mx = x + (c*px + (-s)*py);  ix = trunc((mx - ox) / res)
// The heading's cos/sin are the correctly rounded float values (evaluated in
// double, rounded once) — see DESIGN.md "trig contract".

#include "sensor.cuh"

namespace qmcl {

constexpr int kSensorThreads = 256;
constexpr int kSensorWarps = kSensorThreads / 32;
constexpr int kP = 4;  // particles per warp
constexpr int kParticlesPerBlock = kSensorWarps * kP;

struct SensorParams {
  MapGeom g;
  float z_hit;   // parameters.z_hit (float)
  float zr;      // z_rand / max_laser_distance, float division (map_likelihood.cpp:76-77)
  double p_max;  // max_distance_probability (map_likelihood.cpp:46-48,155)
};

unsigned sensor_grid(size_t n) { return div_up(n, kParticlesPerBlock); }

__global__ void subsample_kernel(const float2 *__restrict__ cloud, size_t n_points, size_t step,
                                 float2 *__restrict__ beams, int n_used) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_used) return;
  float2 p = cloud[static_cast<size_t>(j) * step];
  // x86 float->int of a NaN/huge value yields INT_MIN, i.e. "outside the map";
  // make that explicit instead of relying on conversion corner cases.
  if (!(fabsf(p.x) < 1e30f) || !(fabsf(p.y) < 1e30f)) p = make_float2(1e30f, 1e30f);
  beams[j] = p;
}

__global__ void __launch_bounds__(kSensorThreads)
sensor_kernel(const float *__restrict__ field, const int8_t *__restrict__ state, SensorParams sp,
              SensorLaunch a) {
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const size_t first =
      (static_cast<size_t>(blockIdx.x) * kSensorWarps + warp_in_block) * kP;

  // Lane p prepares particle first+p; the values are then broadcast.
  float lx = 0.f, ly = 0.f, lc = 1.f, ls = 0.f;
  if (lane < kP && first + lane < a.n) {
    lx = a.x[first + lane];
    ly = a.y[first + lane];
    double sd, cd;
    sincos(static_cast<double>(a.t[first + lane]), &sd, &cd);
    lc = static_cast<float>(cd);
    ls = static_cast<float>(sd);
  }
  float X[kP], Y[kP], C[kP], S[kP], NS[kP];
#pragma unroll
  for (int p = 0; p < kP; ++p) {
    X[p] = __shfl_sync(0xffffffffu, lx, p);
    Y[p] = __shfl_sync(0xffffffffu, ly, p);
    C[p] = __shfl_sync(0xffffffffu, lc, p);
    S[p] = __shfl_sync(0xffffffffu, ls, p);
    NS[p] = -S[p];
  }
  double acc[kP];
#pragma unroll
  for (int p = 0; p < kP; ++p) acc[p] = 0.0;

  const float2 *__restrict__ beams = reinterpret_cast<const float2 *>(a.beams);
  const double z_hit = static_cast<double>(sp.z_hit);
  const double zr = static_cast<double>(sp.zr);
  for (int j = lane; j < a.n_beams; j += 32) {
    const float2 b = __ldg(beams + j);
#pragma unroll
    for (int p = 0; p < kP; ++p) {
      // Eigen Isometry2f * Vector2f (pose_2d.h:118-126): translation + (R * v)
      const float mx = __fadd_rn(X[p], __fadd_rn(__fmul_rn(C[p], b.x), __fmul_rn(NS[p], b.y)));
      const float my = __fadd_rn(Y[p], __fadd_rn(__fmul_rn(S[p], b.x), __fmul_rn(C[p], b.y)));
      const int ix = cell_index(mx, sp.g.ox, sp.g.res);
      const int iy = cell_index(my, sp.g.oy, sp.g.res);
      double pz = sp.p_max;
      if (in_map(ix, iy, sp.g)) pz = __ldg(field + static_cast<size_t>(iy) * sp.g.w + ix);
      // z_hit * p is exact in double (two float significands), so the fused
      // form rounds exactly like the reference's mul-then-add.
      const double pw = fma(z_hit, pz, zr);
      acc[p] = __dadd_rn(acc[p], __dmul_rn(__dmul_rn(pw, pw), pw));
    }
  }

  double mine = 0.0;
#pragma unroll
  for (int p = 0; p < kP; ++p) {
    const double s = warp_sum_xor(acc[p]);
    if (lane == p) mine = s;
  }
  const bool valid = lane < kP && first + lane < a.n;
  if (!valid) mine = 0.0;
  // total_weight += weight happens before the penalty (map_likelihood.cpp:112)
  const double warp_pre = warp_sum_xor(mine);
  if (valid) {
    const int ix = cell_index(lx, sp.g.ox, sp.g.res);
    const int iy = cell_index(ly, sp.g.oy, sp.g.res);
    const int st = in_map(ix, iy, sp.g) ? state[static_cast<size_t>(iy) * sp.g.w + ix] : 2;
    double wgt = mine;
    if (st == 2) wgt = 0.0;
    else if (st == 1) wgt = __dmul_rn(wgt, 0.01);
    a.w[first + lane] = wgt;
  }

  __shared__ double s_warp[kSensorWarps];
  __shared__ bool s_last;
  if (lane == 0) s_warp[warp_in_block] = warp_pre;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kSensorWarps; ++i) t = __dadd_rn(t, s_warp[i]);
    a.partials[blockIdx.x] = t;
    __threadfence();
    const unsigned done = atomicAdd(&a.scalars->blocks_done, 1u);
    s_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (!s_last) return;
  // Last block: deterministic reduction of the block partials.
  __threadfence();
  double t = 0.0;
  for (unsigned i = threadIdx.x; i < gridDim.x; i += kSensorThreads)
    t = __dadd_rn(t, __ldcg(a.partials + i));
  t = warp_sum_xor(t);
  __syncthreads();
  if (lane == 0) s_warp[warp_in_block] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
#pragma unroll
    for (int i = 0; i < kSensorWarps; ++i) tot = __dadd_rn(tot, s_warp[i]);
    a.scalars->total_weight = tot;
    a.scalars->blocks_done = 0;
  }
}

qmcl_status launch_subsample(const qmcl_map *m, const float *dev_cloud_xy, size_t n_points,
                             float *dev_beams, int *n_used_out) {
  // map_likelihood.cpp:67-74
  size_t step = 1;
  if (m->params.num_beams) step = n_points / static_cast<size_t>(m->params.num_beams);
  if (step < 1) step = 1;
  const size_t n_used = n_points ? (n_points + step - 1) / step : 0;
  *n_used_out = static_cast<int>(n_used);
  if (n_used == 0) return QMCL_OK;
  subsample_kernel<<<div_up(n_used, 256), 256, 0, m->stream>>>(
      reinterpret_cast<const float2 *>(dev_cloud_xy), n_points, step,
      reinterpret_cast<float2 *>(dev_beams), static_cast<int>(n_used));
  QMCL_LAUNCHED();
  return QMCL_OK;
}

qmcl_status launch_sensor(const qmcl_map *m, const SensorLaunch &a) {
  if (a.n == 0) return QMCL_OK;
  SensorParams sp;
  sp.g = m->geom;
  sp.z_hit = m->params.z_hit;
  sp.zr = m->params.z_rand / m->params.max_laser_distance;
  sp.p_max = m->max_distance_probability;
  sensor_kernel<<<sensor_grid(a.n), kSensorThreads, 0, m->stream>>>(m->field.ptr, m->state.ptr, sp,
                                                                    a);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

}  // namespace qmcl
