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
// operation (no FMA contraction, correctly rounded division, truncation):
//   mx = x + (c*px + (-s)*py);  ix = trunc((mx - ox) / res)
// The heading's cos/sin are the correctly rounded float values (evaluated in
// double, rounded once) — see DESIGN.md "trig contract".
//
// Two value paths, both exact restatements of
//   pw = z_hit * p + z_rand/max_range (double);  weight += pw*pw*pw:
//  * FIELD  : gathers the float32 table (any field, e.g. one injected through
//             qmcl_map_load_field) and does the double arithmetic per eval.
//  * PALETTE: when the map was built here, a cell's value is a function of its
//             integer squared distance k, so the map is also kept as uint16 k
//             (2 B per eval instead of 4) and pw^3 is looked up in a
//             shared-memory table of doubles computed with the reference's
//             expression — no per-eval conversion or multiply, same bits.
//
// Division: (m - o) / res is evaluated as q = a*r, q' = fma(fma(q,-res,a), r, q)
// with r = RN(1/res), which yields the correctly rounded quotient (Markstein);
// a self-check kernel at map load compares it with IEEE division around every
// cell boundary of the map and the kernel falls back to __fdiv_rn if any index
// differs.

#include "sensor.cuh"

namespace qmcl {

constexpr int kSensorThreads = 512;
constexpr int kSensorWarps = kSensorThreads / 32;
constexpr int kP = 4;  // particles per warp
constexpr int kParticlesPerBlock = kSensorWarps * kP;

struct SensorParams {
  MapGeom g;
  float inv_res;   // RN(1 / res)
  float neg_res;   // -res
  float z_hit;     // parameters.z_hit (float)
  float zr;        // z_rand / max_laser_distance, float division (map_likelihood.cpp:76-77)
  double p_max;    // max_distance_probability (map_likelihood.cpp:46-48,155)
  const float *field;
  const uint16_t *code;
  const double *pw3_lut;  // n_codes entries + 1 (out of map)
  int n_codes;
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

template <bool FASTDIV>
__device__ __forceinline__ int cell_of(float pos, float offset, const SensorParams &sp) {
  const float a = __fsub_rn(pos, offset);
  if (FASTDIV) {
    const float q = __fmul_rn(a, sp.inv_res);
    const float rem = __fmaf_rn(q, sp.neg_res, a);
    return __float2int_rz(__fmaf_rn(rem, sp.inv_res, q));
  }
  return __float2int_rz(__fdiv_rn(a, sp.g.res));
}

// pw^3 for every code (and one extra entry for "outside the map").
__global__ void pw3_lut_kernel(const float *__restrict__ lut_field, int n_codes, float z_hit,
                               float zr, double p_max, double *__restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n_codes) return;
  const double p = (k < n_codes) ? static_cast<double>(lut_field[k]) : p_max;
  const double pw = fma(static_cast<double>(z_hit), p, static_cast<double>(zr));
  out[k] = __dmul_rn(__dmul_rn(pw, pw), pw);
}

// Compares the 3-op division with IEEE division: every cell boundary of the
// map +-4 ulps on both axes, plus a pseudo-random sweep of the coordinate range.
__global__ void fastdiv_check_kernel(SensorParams sp, int *mismatch) {
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned cells = max(sp.g.w, sp.g.h) + 8;
  bool bad = false;
  if (tid < cells * 9) {
    const int cell = static_cast<int>(tid / 9) - 4;
    const int ulps = static_cast<int>(tid % 9) - 4;
    const float edge = __fmul_rn(static_cast<float>(cell), sp.g.res);
    const float a = __int_as_float(__float_as_int(edge) + (edge >= 0 ? ulps : -ulps));
    // pos - offset == a for offset 0; the subtraction is common to both paths
    bad |= cell_of<true>(a, 0.f, sp) != cell_of<false>(a, 0.f, sp);
  }
  uint32_t s = tid * 2654435761u + 12345u;
#pragma unroll 4
  for (int i = 0; i < 64; ++i) {
    s ^= s << 13; s ^= s >> 17; s ^= s << 5;
    const float a = (static_cast<float>(s >> 8) * (1.0f / 16777216.0f)) *
                        (static_cast<float>(cells) * sp.g.res + 4.f) - 2.f;
    bad |= cell_of<true>(a, 0.f, sp) != cell_of<false>(a, 0.f, sp);
  }
  if (bad) atomicExch(mismatch, 1);
}

template <bool PALETTE, bool FASTDIV>
__global__ void __launch_bounds__(kSensorThreads)
sensor_kernel(const int8_t *__restrict__ state, SensorParams sp, SensorLaunch a) {
  extern __shared__ double s_lut[];  // PALETTE: n_codes + 1 entries
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const size_t first =
      (static_cast<size_t>(blockIdx.x) * kSensorWarps + warp_in_block) * kP;

  if (PALETTE) {
    for (int k = threadIdx.x; k <= sp.n_codes; k += kSensorThreads) s_lut[k] = sp.pw3_lut[k];
  }

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
  if (PALETTE) __syncthreads();

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
      const int ix = cell_of<FASTDIV>(mx, sp.g.ox, sp);
      const int iy = cell_of<FASTDIV>(my, sp.g.oy, sp);
      const bool inside = in_map(ix, iy, sp.g);
      if (PALETTE) {
        int code = sp.n_codes;  // outside the map -> p_max entry
        if (inside) code = __ldg(sp.code + static_cast<size_t>(iy) * sp.g.w + ix);
        acc[p] = __dadd_rn(acc[p], s_lut[code]);
      } else {
        double pz = sp.p_max;
        if (inside) pz = __ldg(sp.field + static_cast<size_t>(iy) * sp.g.w + ix);
        // z_hit * p is exact in double (two float significands), so the fused
        // form rounds exactly like the reference's mul-then-add.
        const double pw = fma(z_hit, pz, zr);
        acc[p] = __dadd_rn(acc[p], __dmul_rn(__dmul_rn(pw, pw), pw));
      }
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
    const int ix = cell_of<false>(lx, sp.g.ox, sp);
    const int iy = cell_of<false>(ly, sp.g.oy, sp);
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

static SensorParams make_params(const qmcl_map *m) {
  SensorParams sp;
  sp.g = m->geom;
  sp.inv_res = 1.0f / m->geom.res;
  sp.neg_res = -m->geom.res;
  sp.z_hit = m->params.z_hit;
  sp.zr = m->params.z_rand / m->params.max_laser_distance;
  sp.p_max = m->max_distance_probability;
  sp.field = m->field.ptr;
  sp.code = reinterpret_cast<const uint16_t *>(m->code.ptr);
  sp.pw3_lut = m->pw3_lut.ptr;
  sp.n_codes = m->n_codes;
  return sp;
}

// Called at map load: decides whether the 3-op division may be used.
qmcl_status sensor_prepare_map(qmcl_map *m) {
  SensorParams sp = make_params(m);
  QMCL_TRY(m->scan_tmp.ensure(8));
  int *d_flag = reinterpret_cast<int *>(m->scan_tmp.ptr);
  QMCL_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), m->stream));
  const unsigned threads = (std::max(m->geom.w, m->geom.h) + 8) * 9;
  fastdiv_check_kernel<<<div_up(threads, 256), 256, 0, m->stream>>>(sp, d_flag);
  QMCL_LAUNCHED();
  int flag = 1;
  QMCL_CUDA(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
  QMCL_CUDA(cudaStreamSynchronize(m->stream));
  m->fastdiv_ok = (flag == 0);
  m->lut_params_valid = false;
  return QMCL_OK;
}

template <bool PALETTE, bool FASTDIV>
static qmcl_status launch_variant(const qmcl_map *m, const SensorParams &sp, const SensorLaunch &a,
                                  size_t smem) {
  auto kern = sensor_kernel<PALETTE, FASTDIV>;
  if (smem > 48 * 1024)
    QMCL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
  kern<<<sensor_grid(a.n), kSensorThreads, smem, m->stream>>>(m->state.ptr, sp, a);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

qmcl_status launch_sensor(qmcl_map *m, const SensorLaunch &a) {
  if (a.n == 0) return QMCL_OK;
  SensorParams sp = make_params(m);
  const bool palette = m->has_code && !m->force_field_path &&
                       static_cast<size_t>(m->n_codes + 1) * sizeof(double) <= 96 * 1024;
  if (palette) {
    // refresh pw^3 when z_hit / z_rand / max range / p_max changed
    if (!m->lut_params_valid || m->lut_z_hit != sp.z_hit || m->lut_zr != sp.zr ||
        m->lut_p_max != sp.p_max) {
      QMCL_TRY(m->pw3_lut.ensure(m->n_codes + 1));
      sp.pw3_lut = m->pw3_lut.ptr;
      pw3_lut_kernel<<<div_up(m->n_codes + 1, 256), 256, 0, m->stream>>>(
          m->lut.ptr + m->n_codes, m->n_codes, sp.z_hit, sp.zr, sp.p_max, m->pw3_lut.ptr);
      QMCL_LAUNCHED();
      m->lut_params_valid = true;
      m->lut_z_hit = sp.z_hit;
      m->lut_zr = sp.zr;
      m->lut_p_max = sp.p_max;
    }
    const size_t smem = static_cast<size_t>(m->n_codes + 1) * sizeof(double);
    return m->fastdiv_ok ? launch_variant<true, true>(m, sp, a, smem)
                         : launch_variant<true, false>(m, sp, a, smem);
  }
  return m->fastdiv_ok ? launch_variant<false, true>(m, sp, a, 0)
                       : launch_variant<false, false>(m, sp, a, 0);
}

}  // namespace qmcl
