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
//  * PAL64  : when the map was built here, a cell's value is a function of its
//             integer squared distance k, so the map is also kept as uint16 k
//             (2 B per eval instead of 4) and pw^3 is looked up in a
//             shared-memory table of doubles computed with the reference's
//             expression — no per-eval conversion or multiply, same bits.
//  * PAL32  : the same table rounded to float (one shared-memory wavefront per
//             lookup instead of two or more) and summed in float for at most 8
//             terms before being folded into the double accumulator.  Worst
//             case 5e-7 relative, measured ~1e-7, against the 1e-5 bar.  This
//             stands in when PAL8 does not apply; PAL64 and FIELD stay
//             selectable (qmcl_map_set_sensor_path).
//  * PAL8   : the default when the map was built here.  Far from an obstacle z_hit * p vanishes against
//             z_rand/max_range, so the float table holds few *distinct* values
//             over the squared distances that can occur (72 at the defaults,
//             230 at sigma_hit = 0.2).  When they fit a byte the map is also
//             kept as one palette index per cell, 16 x 8 cells per 128-byte
//             line: fewer lines per gather, a 1 KB table instead of 13 KB.
//             The floats are PAL32's, so are the weights, bit for bit.
//
// Division: (m - o) / res is evaluated by the cheapest of three forms that a
// self-check kernel at map load found to give the same truncated quotient as
// IEEE division around every cell boundary of the map (fastdiv_check_kernel):
// DIV_SPLIT, the usual choice, q = fma(a, r_hi, a * r_lo) with r_hi + r_lo =
// 1/res to 48 bits (two packed ops); DIV_MARKSTEIN, q = a * r, q' =
// fma(fma(q, -res, a), r, q) with r = RN(1/res), the correctly rounded quotient
// (three ops); DIV_IEEE, __fdiv_rn, when neither passed.

#include "sensor.cuh"

#include <type_traits>
#include <vector>

namespace qmcl {

#ifndef QMCL_DIAG
#define QMCL_DIAG 0
#endif
#ifndef QMCL_SENSOR_FLOOR_FMA
#define QMCL_SENSOR_FLOOR_FMA 1  // interior path: truncation on the FP32 pipe (floor2_nonneg), round 2 A/B
#endif
#ifndef QMCL_SENSOR_WIDE_ADDR
#define QMCL_SENSOR_WIDE_ADDR 1  // palette byte address as one IMAD.WIDE, round 2 A/B
#endif
#ifndef QMCL_SENSOR_THREADS
#define QMCL_SENSOR_THREADS 256
#endif
#ifndef QMCL_SENSOR_P
#define QMCL_SENSOR_P 4
#endif
#ifndef QMCL_SENSOR_MINBLOCKS
#define QMCL_SENSOR_MINBLOCKS 5  // float-table path: 48 registers, no spills; others 4
#endif
constexpr int kSensorThreads = QMCL_SENSOR_THREADS;
constexpr int kSensorWarps = kSensorThreads / 32;
constexpr int kP = QMCL_SENSOR_P;  // particles per warp
constexpr int kParticlesPerBlock = kSensorWarps * kP;

struct SensorParams {
  MapGeom g;
  float inv_res;   // RN(1 / res)
  float inv_res_lo;  // RN(1/res - inv_res): the second word of the reciprocal
  float tiny;      // 2^-149 (see floor2_nonneg)
  float neg_res;   // -res
  float neg_zero;  // -0.0f, deliberately a run-time value (see pack2 notes)
  float z_hit;     // parameters.z_hit (float)
  float zr;        // z_rand / max_laser_distance, float division (map_likelihood.cpp:76-77)
  double p_max;    // max_distance_probability (map_likelihood.cpp:46-48,155)
  const float *field;
  const uint16_t *code;
  const double *pw3_lut;  // n_codes entries + 1 (out of map)
  const float *pw3_lut_f32;
  int n_codes;
  const uint8_t *code8;   // PAL8: palette index per cell
  const float *pal8_lut;  // PAL8: pal8_n + 2 floats
  int pal8_n;
};

unsigned sensor_grid(size_t n) { return div_up(n, kParticlesPerBlock); }

// ---- packed FP32 (sm_100 f32x2): two IEEE-rounded float ops per instruction.
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 (measured on
// B200: 67 % of results change), so a product that feeds an addition is
// written as fma(a, b, nz) with nz = -0.0f passed at RUN time: a*b + (-0) is
// a*b for every a*b, and an FMA cannot be fused into the following add.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float *lo, float *hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(*lo), "=f"(*hi) : "l"(v));
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// trunc() of two non-negative floats below 2^23 without the conversion unit:
// q * 2^-149 rounded toward zero is the subnormal whose bit pattern is the
// integer trunc(q) (subnormals are multiples of 2^-149; no flush-to-zero here).
__device__ __forceinline__ void floor2_nonneg(f32x2 q, f32x2 tiny2, int *lo, int *hi) {
  f32x2 r;
  asm("mul.rz.f32x2 %0, %1, %2;" : "=l"(r) : "l"(q), "l"(tiny2));
  asm("mov.b64 {%0, %1}, %2;" : "=r"(*lo), "=r"(*hi) : "l"(r));
}

// Also leaves an upper bound of the longest used beam in *rmax (float bits,
// zeroed by the launcher): particles farther than that from every map edge can
// skip the in-map test of each evaluation.
__global__ void subsample_kernel(const float2 *__restrict__ cloud, size_t n_points, size_t step,
                                 float2 *__restrict__ beams, int n_used,
                                 unsigned *__restrict__ rmax) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  float len = 0.f;
  if (j < n_used) {
    float2 p = cloud[static_cast<size_t>(j) * step];
    // x86 float->int of a NaN/huge value yields INT_MIN, i.e. "outside the map";
    // make that explicit instead of relying on conversion corner cases.
    if (!(fabsf(p.x) < 1e30f) || !(fabsf(p.y) < 1e30f)) p = make_float2(1e30f, 1e30f);
    beams[j] = p;
    len = __fadd_ru(fabsf(p.x), fabsf(p.y));  // |p|_1 >= |p|_2, rounded up
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) len = fmaxf(len, __shfl_xor_sync(0xffffffffu, len, o));
  if ((threadIdx.x & 31) == 0 && len > 0.f) atomicMax(rmax, __float_as_uint(len));
}

enum { DIV_IEEE = 0, DIV_MARKSTEIN = 1, DIV_SPLIT = 2 };

template <int DIV>
__device__ __forceinline__ int cell_of(float pos, float offset, const SensorParams &sp) {
  const float a = __fsub_rn(pos, offset);
  if (DIV == DIV_SPLIT) {
    // a * (r_hi + r_lo) with one rounding too many only 2^-24 below the result
    return __float2int_rz(__fmaf_rn(a, sp.inv_res, __fmul_rn(a, sp.inv_res_lo)));
  }
  if (DIV == DIV_MARKSTEIN) {
    const float q = __fmul_rn(a, sp.inv_res);
    const float rem = __fmaf_rn(q, sp.neg_res, a);
    return __float2int_rz(__fmaf_rn(rem, sp.inv_res, q));
  }
  return __float2int_rz(__fdiv_rn(a, sp.g.res));
}

// pw^3 for every code (and one extra entry for "outside the map").
__global__ void pw3_lut_kernel(const float *__restrict__ lut_field, int n_codes, float z_hit,
                               float zr, double p_max, double *__restrict__ out,
                               float *__restrict__ out_f32) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > n_codes + 1) return;
  const double p = (k < n_codes) ? static_cast<double>(lut_field[k]) : p_max;
  const double pw = fma(static_cast<double>(z_hit), p, static_cast<double>(zr));
  double pw3 = __dmul_rn(__dmul_rn(pw, pw), pw);
  if (k == n_codes + 1) pw3 = 0.0;  // "nothing pending" entry of the software pipeline
  out[k] = pw3;
  out_f32[k] = static_cast<float>(pw3);
}

// Compares the multiply-based divisions with IEEE division: every cell boundary
// of the map +-4 ulps on both axes (truncated quotients can only differ where
// the true quotient is within a few ulps of an integer, and both forms are
// monotone), plus a pseudo-random sweep of the coordinate range.
// mismatch[0]: 3-op Markstein form, mismatch[1]: 2-op split-reciprocal form.
__global__ void fastdiv_check_kernel(SensorParams sp, int *mismatch) {
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned cells = max(sp.g.w, sp.g.h) + 8;
  bool bad = false, bad2 = false;
  if (tid < cells * 9) {
    const int cell = static_cast<int>(tid / 9) - 4;
    const int ulps = static_cast<int>(tid % 9) - 4;
    const float edge = __fmul_rn(static_cast<float>(cell), sp.g.res);
    const float a = __int_as_float(__float_as_int(edge) + (edge >= 0 ? ulps : -ulps));
    // pos - offset == a for offset 0; the subtraction is common to both paths
    const int want = cell_of<DIV_IEEE>(a, 0.f, sp);
    bad |= cell_of<DIV_MARKSTEIN>(a, 0.f, sp) != want;
    bad2 |= cell_of<DIV_SPLIT>(a, 0.f, sp) != want;
  }
  uint32_t s = tid * 2654435761u + 12345u;
#pragma unroll 4
  for (int i = 0; i < 64; ++i) {
    s ^= s << 13; s ^= s >> 17; s ^= s << 5;
    const float a = (static_cast<float>(s >> 8) * (1.0f / 16777216.0f)) *
                        (static_cast<float>(cells) * sp.g.res + 4.f) - 2.f;
    const int want = cell_of<DIV_IEEE>(a, 0.f, sp);
    bad |= cell_of<DIV_MARKSTEIN>(a, 0.f, sp) != want;
    bad2 |= cell_of<DIV_SPLIT>(a, 0.f, sp) != want;
  }
  if (bad) atomicExch(mismatch, 1);
  if (bad2) atomicExch(mismatch + 1, 1);
}

enum { PATH_FIELD = 0, PATH_PAL64 = 1, PATH_PAL32 = 2, PATH_PAL8 = 3 };

// Shared-memory float table read by byte offset: the code map stores 4 * code,
// so the lookup needs no address arithmetic.
__device__ __forceinline__ float lds_f32_at(const float *table, unsigned byte_off) {
  return *reinterpret_cast<const float *>(reinterpret_cast<const char *>(table) + byte_off);
}

// One palette byte, zero-extended by the load itself: the value is known to fit
// a byte, so the table index needs no mask afterwards.
__device__ __forceinline__ unsigned ldg_u8(const uint8_t *p) {
  unsigned v;
  asm("ld.global.nc.u8 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
// base + offset as one IMAD.WIDE (FP32 pipe) instead of an IADD3 pair.
__device__ __forceinline__ const uint8_t *byte_at(const uint8_t *base, unsigned off) {
#if QMCL_SENSOR_WIDE_ADDR
  unsigned long long r;
  asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(r) : "r"(off), "l"(base));
  return reinterpret_cast<const uint8_t *>(r);
#else
  return base + off;
#endif
}

// The kernel body for block `bx` of the `nbx` blocks working on job `a` (one
// filter's particle set).  sensor_kernel runs it for a single filter,
// sensor_batch_kernel for many filters in one launch (blockIdx.y = filter).
template <int PATH, int DIV>
__device__ __forceinline__ void sensor_body(const int8_t *__restrict__ state, const SensorParams &sp,
                                            const SensorLaunch &a, const unsigned bx, const unsigned nbx) {
  extern __shared__ double s_lut[];  // PAL64: n_codes + 1 doubles; PAL32: floats
  float *s_lut_f32 = reinterpret_cast<float *>(s_lut);
  constexpr bool PALETTE = PATH != PATH_FIELD;
  const int lane = threadIdx.x & 31;
  const int warp_in_block = threadIdx.x >> 5;
  const size_t first =
      (static_cast<size_t>(bx) * kSensorWarps + warp_in_block) * kP;

  // table: [0, n_codes) codes, n_codes = outside the map, n_codes+1 = 0
  if (PATH == PATH_PAL64) {
    for (int k = threadIdx.x; k <= sp.n_codes + 1; k += kSensorThreads) s_lut[k] = sp.pw3_lut[k];
  } else if (PATH == PATH_PAL32) {
    for (int k = threadIdx.x; k <= sp.n_codes + 1; k += kSensorThreads)
      s_lut_f32[k] = sp.pw3_lut_f32[k];
  } else if (PATH == PATH_PAL8) {
    for (int k = threadIdx.x; k <= sp.pal8_n + 1; k += kSensorThreads) s_lut_f32[k] = sp.pal8_lut[k];
  }

  // Lane p prepares particle first+p; the values are then broadcast.
  float lx = 0.f, ly = 0.f, lc = 1.f, ls = 0.f;
  // every beam end of this particle lies inside the map for sure; padding
  // particles past the end of the set sit at (0, 0) and are never "safe"
  bool safe = lane >= kP;
  if (lane < kP && first + lane < a.n) {
    lx = a.x[first + lane];
    ly = a.y[first + lane];
    double sd, cd;
    sincos(static_cast<double>(a.t[first + lane]), &sd, &cd);
    lc = static_cast<float>(cd);
    ls = static_cast<float>(sd);
    // |R v| <= |v|_1 <= rmax; two cells of slack cover every rounding on the way
    const float reach = __ldg(a.rmax) + 2.f * sp.g.res;
    const float dx = lx - sp.g.ox, dy = ly - sp.g.oy;
    safe = dx - reach > 0.f && dy - reach > 0.f &&
           dx + reach < static_cast<float>(sp.g.w) * sp.g.res &&
           dy + reach < static_cast<float>(sp.g.h) * sp.g.res;
  }
  const bool group_safe = __all_sync(0xffffffffu, safe);
  // per-particle packed constants: (c, s), (-s, c), (x, y)
  f32x2 CS[kP], NSC[kP], XY[kP];
#pragma unroll
  for (int p = 0; p < kP; ++p) {
    const float x = __shfl_sync(0xffffffffu, lx, p);
    const float y = __shfl_sync(0xffffffffu, ly, p);
    const float c = __shfl_sync(0xffffffffu, lc, p);
    const float sn = __shfl_sync(0xffffffffu, ls, p);
    CS[p] = pack2(c, sn);
    NSC[p] = pack2(-sn, c);
    XY[p] = pack2(x, y);
  }
  double acc[kP];
#pragma unroll
  for (int p = 0; p < kP; ++p) acc[p] = 0.0;
  if (PALETTE) __syncthreads();

  const float2 *__restrict__ beams = reinterpret_cast<const float2 *>(a.beams);

  const double z_hit = static_cast<double>(sp.z_hit);
  const double zr = static_cast<double>(sp.zr);
  const f32x2 NZ = pack2(sp.neg_zero, sp.neg_zero);
  const f32x2 NO = pack2(-sp.g.ox, -sp.g.oy);
  const f32x2 RR = pack2(sp.inv_res, sp.inv_res);
  const f32x2 RLO = pack2(sp.inv_res_lo, sp.inv_res_lo);
  const f32x2 NRES = pack2(sp.neg_res, sp.neg_res);
  const f32x2 TINY = pack2(sp.tiny, sp.tiny);
  const unsigned w = sp.g.w;
  const unsigned strip = PATH == PATH_PAL8 ? code8_strip_stride(sp.g.h) : code_strip_stride(sp.g.h);
  // Software pipeline: the beam for the next step is fetched before this
  // step's arithmetic, and the value gathered in step i is consumed in step
  // i+1, so a full loop body of independent work covers each load's latency.
  float facc[kP];
  int code_prev[kP];    // palette paths: code gathered in the previous step
  double pz_prev[kP];   // field path: probability gathered in the previous step
#pragma unroll
  for (int p = 0; p < kP; ++p) {
    facc[p] = 0.f;
    // table entry holding 0 (uint16 codes are byte offsets, PAL8 codes are indices)
    code_prev[p] = PATH == PATH_PAL8 ? sp.pal8_n + 1 : 4 * (sp.n_codes + 1);
    pz_prev[p] = -1.0;              // "nothing pending"
  }
  // One step of the walk: this lane's beam b against the kP particles.  The
  // gathers of this step are consumed in the next one (see above).
  // `masked`: the lane may be past the end of the scan (`valid` false); it then
  // re-evaluates its last beam and contributes the table's zero entry.
  auto body = [&](auto checked_tag, auto masked_tag, const float2 b, const bool valid) {
    constexpr bool CHECKED = decltype(checked_tag)::value;
    constexpr bool MASKED = decltype(masked_tag)::value;
    const f32x2 PX = pack2(b.x, b.x), PY = pack2(b.y, b.y);
    int code_cur[kP];
    double pz_cur[kP];
#pragma unroll
    for (int p = 0; p < kP; ++p) {
      // Eigen Isometry2f * Vector2f (pose_2d.h:118-126): translation + (R * v),
      // both coordinates at once:
      //   (mx, my) = (x, y) + ((c*px, s*px) + ((-s)*py, c*py))
      const f32x2 t3 = add2(fma2(CS[p], PX, NZ), fma2(NSC[p], PY, NZ));
      const f32x2 mxy = add2(XY[p], t3);
      const f32x2 axy = add2(mxy, NO);  // m - o
      int ix, iy;
      if (DIV == DIV_IEEE) {
        float ax, ay;
        unpack2(axy, &ax, &ay);
        ix = __float2int_rz(__fdiv_rn(ax, sp.g.res));
        iy = __float2int_rz(__fdiv_rn(ay, sp.g.res));
      } else {
        f32x2 q2;
        if (DIV == DIV_SPLIT) {
          q2 = fma2(axy, RR, mul2(axy, RLO));
        } else {
          const f32x2 q = mul2(axy, RR);
          q2 = fma2(fma2(q, NRES, axy), RR, q);
        }
#if QMCL_SENSOR_FLOOR_FMA
        if (!CHECKED) {
          floor2_nonneg(q2, TINY, &ix, &iy);  // interior: both quotients are >= 0
        } else
#endif
        {
          // conversion unit (XU): the FP32 pipe is the busier one here
          float qx, qy;
          unpack2(q2, &qx, &qy);
          ix = __float2int_rz(qx);
          iy = __float2int_rz(qy);
        }
      }
      const bool inside = !CHECKED || in_map(ix, iy, sp.g);
      // 32-bit cell offset (maps are far below 2^32 cells); the code map is
      // stored in 8-cell strips (common.cuh)
      const unsigned cell =
          PATH == PATH_PAL8
              ? code8_offset(static_cast<unsigned>(ix), static_cast<unsigned>(iy), strip)
              : (PALETTE ? code_offset(static_cast<unsigned>(ix), static_cast<unsigned>(iy), strip)
                         : static_cast<unsigned>(iy) * w + static_cast<unsigned>(ix));
      if (PATH == PATH_PAL8) {
        code_cur[p] = sp.pal8_n;  // outside the map -> p_max entry
#if QMCL_DIAG == 4
        // upper bound of any map-staging scheme (shared-memory / TMA window): no gather at all
        if (inside) code_cur[p] = static_cast<int>(cell & 63u);
#elif QMCL_DIAG == 5
        // every gather hits one L1-resident 4 KB window
        if (inside) code_cur[p] = static_cast<int>(ldg_u8(byte_at(sp.code8, cell & 4095u)));
#else
        if (inside) code_cur[p] = static_cast<int>(ldg_u8(byte_at(sp.code8, cell)));
#endif
        if (MASKED && !valid) code_cur[p] = sp.pal8_n + 1;
      } else if (PALETTE) {
        code_cur[p] = 4 * sp.n_codes;  // outside the map -> p_max entry
#if QMCL_DIAG == 2 || QMCL_DIAG == 3
        if (inside) code_cur[p] = __ldg(sp.code + ((cell & 1u) + 2u * lane + 64u * p + 4096u * (bx & 255)));
#elif QMCL_DIAG == 4
        if (inside) code_cur[p] = (cell & 0x3ffu) << 2;
#else
        if (inside) code_cur[p] = __ldg(sp.code + cell);
#endif
        if (MASKED && !valid) code_cur[p] = 4 * (sp.n_codes + 1);
      } else {
        pz_cur[p] = sp.p_max;
        if (inside) pz_cur[p] = __ldg(sp.field + cell);
        if (MASKED && !valid) pz_cur[p] = -1.0;
      }
    }
#pragma unroll
    for (int p = 0; p < kP; ++p) {
      if (PATH == PATH_PAL64) {
        acc[p] = __dadd_rn(acc[p], s_lut[code_prev[p] >> 2]);
      } else if (PATH == PATH_PAL32) {
#if QMCL_DIAG == 1 || QMCL_DIAG == 3
        facc[p] = __fadd_rn(facc[p], __int_as_float(code_prev[p]));
#else
        facc[p] = __fadd_rn(facc[p], lds_f32_at(s_lut_f32, code_prev[p]));
#endif
      } else if (PATH == PATH_PAL8) {
        facc[p] = __fadd_rn(facc[p], s_lut_f32[code_prev[p]]);
      } else if (pz_prev[p] >= 0.0) {
        // z_hit * p is exact in double (two float significands), so the fused
        // form rounds exactly like the reference's mul-then-add.
        const double pw = fma(z_hit, pz_prev[p], zr);
        acc[p] = __dadd_rn(acc[p], __dmul_rn(__dmul_rn(pw, pw), pw));
      }
      code_prev[p] = code_cur[p];
      pz_prev[p] = pz_cur[p];
    }
  };
  // float partial sums never hold more than 8 terms
  auto fold = [&]() {
    if (PATH == PATH_PAL32 || PATH == PATH_PAL8) {
#pragma unroll
      for (int p = 0; p < kP; ++p) {
        acc[p] = __dadd_rn(acc[p], static_cast<double>(facc[p]));
        facc[p] = 0.f;
      }
    }
  };
  // The walk over the scan, instantiated with and without the in-map test:
  // groups of 8 full steps (all 32 lanes busy, straight-line code, one fold per
  // group), then the ragged rest.
  auto walk = [&](auto checked_tag) {
    const int n_beams = a.n_beams;
    const int n_grouped = n_beams & ~255;
    float2 b = (lane < n_beams) ? __ldg(beams + lane) : make_float2(0.f, 0.f);
    for (int base = lane; base < n_grouped; base += 256) {
#pragma unroll
      for (int s8 = 0; s8 < 8; ++s8) {
        const int jn = base + 32 * s8 + 32;
        const float2 bn = (jn < n_beams) ? __ldg(beams + jn) : b;
        body(checked_tag, std::false_type{}, b, true);
        b = bn;
      }
      fold();
    }
    // The ragged rest (< 256 beams, i.e. at most 8 steps) as straight-line groups
    // of 4, 2 and 1 steps in which lanes past the end are masked: the same code
    // quality as the grouped part instead of a divergent loop.
    const int rest_steps = (n_beams - n_grouped + 31) >> 5;
    int j = n_grouped + lane;
    auto ragged = [&](auto steps_tag) {
      constexpr int STEPS = decltype(steps_tag)::value;
#pragma unroll
      for (int s = 0; s < STEPS; ++s) {
        const float2 bn = (j + 32 < n_beams) ? __ldg(beams + j + 32) : b;
        body(checked_tag, std::true_type{}, b, j < n_beams);
        b = bn;
        j += 32;
      }
    };
    if (rest_steps & 8) {
      ragged(std::integral_constant<int, 4>{});
      ragged(std::integral_constant<int, 4>{});
    }
    if (rest_steps & 4) ragged(std::integral_constant<int, 4>{});
    if (rest_steps & 2) ragged(std::integral_constant<int, 2>{});
    if (rest_steps & 1) ragged(std::integral_constant<int, 1>{});
    fold();
  };
  if (group_safe) walk(std::false_type{});
  else walk(std::true_type{});
  // drain the pipeline
#pragma unroll
  for (int p = 0; p < kP; ++p) {
    if (PATH == PATH_PAL64) {
      acc[p] = __dadd_rn(acc[p], s_lut[code_prev[p] >> 2]);
    } else if (PATH == PATH_PAL32) {
      facc[p] = __fadd_rn(facc[p], lds_f32_at(s_lut_f32, code_prev[p]));
      acc[p] = __dadd_rn(acc[p], static_cast<double>(facc[p]));
    } else if (PATH == PATH_PAL8) {
      facc[p] = __fadd_rn(facc[p], s_lut_f32[code_prev[p]]);
      acc[p] = __dadd_rn(acc[p], static_cast<double>(facc[p]));
    } else if (pz_prev[p] >= 0.0) {
      const double pw = fma(z_hit, pz_prev[p], zr);
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
    const int ix = cell_of<DIV_IEEE>(lx, sp.g.ox, sp);
    const int iy = cell_of<DIV_IEEE>(ly, sp.g.oy, sp);
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
    a.partials[bx] = t;
    __threadfence();
    const unsigned done = atomicAdd(&a.scalars->blocks_done, 1u);
    s_last = (done == nbx - 1);
  }
  __syncthreads();
  if (!s_last) return;
  // Last block: deterministic reduction of the block partials.
  __threadfence();
  double t = 0.0;
  for (unsigned i = threadIdx.x; i < nbx; i += kSensorThreads)
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

template <int PATH, int DIV>
__global__ void __launch_bounds__(kSensorThreads,
                                  (PATH == PATH_PAL32 || PATH == PATH_PAL8) ? QMCL_SENSOR_MINBLOCKS : 4)
sensor_kernel(const int8_t *__restrict__ state, const __grid_constant__ SensorParams sp,
              const __grid_constant__ SensorLaunch a) {
  sensor_body<PATH, DIV>(state, sp, a, blockIdx.x, gridDim.x);
}

// Many filters on one map in one launch: blockIdx.y selects the job, whose
// particle set may need fewer blocks than the grid is wide.
template <int PATH, int DIV>
__global__ void __launch_bounds__(kSensorThreads,
                                  (PATH == PATH_PAL32 || PATH == PATH_PAL8) ? QMCL_SENSOR_MINBLOCKS : 4)
sensor_batch_kernel(const int8_t *__restrict__ state, const __grid_constant__ SensorParams sp,
                    const SensorLaunch *__restrict__ jobs) {
  __shared__ SensorLaunch s_a;
  if (threadIdx.x < sizeof(SensorLaunch) / 8)
    reinterpret_cast<uint64_t *>(&s_a)[threadIdx.x] =
        reinterpret_cast<const uint64_t *>(jobs + blockIdx.y)[threadIdx.x];
  __syncthreads();
  const unsigned nbx = static_cast<unsigned>((s_a.n + kParticlesPerBlock - 1) / kParticlesPerBlock);
  if (blockIdx.x >= nbx) return;
  sensor_body<PATH, DIV>(state, sp, s_a, blockIdx.x, nbx);
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
  // the bound lives behind the beams (callers allocate 2 * n_points + 2 floats)
  unsigned *rmax = reinterpret_cast<unsigned *>(dev_beams + 2 * n_points);
  QMCL_CUDA(cudaMemsetAsync(rmax, 0, sizeof(unsigned), m->stream));
  subsample_kernel<<<div_up(n_used, 256), 256, 0, m->stream>>>(
      reinterpret_cast<const float2 *>(dev_cloud_xy), n_points, step,
      reinterpret_cast<float2 *>(dev_beams), static_cast<int>(n_used), rmax);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

static SensorParams make_params(const qmcl_map *m) {
  SensorParams sp;
  sp.g = m->geom;
  sp.inv_res = 1.0f / m->geom.res;
  sp.inv_res_lo = static_cast<float>(1.0 / static_cast<double>(m->geom.res) -
                                     static_cast<double>(sp.inv_res));
  sp.tiny = 1.4012984643248171e-45f;  // 0x00000001
  sp.neg_res = -m->geom.res;
  sp.neg_zero = -0.0f;
  sp.z_hit = m->params.z_hit;
  sp.zr = m->params.z_rand / m->params.max_laser_distance;
  sp.p_max = m->max_distance_probability;
  sp.field = m->field.ptr;
  sp.code = reinterpret_cast<const uint16_t *>(m->code.ptr);
  sp.pw3_lut = m->pw3_lut.ptr;
  sp.pw3_lut_f32 = m->pw3_lut_f32.ptr;
  sp.n_codes = m->n_codes;
  sp.code8 = m->code8.ptr;
  sp.pal8_lut = m->pal8_lut.ptr;
  sp.pal8_n = m->pal8_n;
  return sp;
}

// PAL8: uint16 d^2 code map -> one palette index per cell.  A code the palette
// does not cover (rank 255) raises the flag; the caller then drops the path.
__global__ void __launch_bounds__(256)
code8_build_kernel(const uint16_t *__restrict__ code16, const uint8_t *__restrict__ rank,
                   uint8_t *__restrict__ code8, uint32_t w, uint32_t h, int *__restrict__ flag) {
  const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= w) return;
  const unsigned k = code16[code_offset(x, y, code_strip_stride(h))] >> 2;
  const uint8_t r = rank[k];
  if (r == 0xff) atomicExch(flag, 1);
  code8[code8_offset(x, y, code8_strip_stride(h))] = r;
}

// Builds the one-byte palette for the current float pw^3 table (which must be
// up to date on the device).  Runs when the table changes — a map load or a
// parameter change — not per scan: it synchronises.  Leaves pal8_n = 0 when the
// distinct values do not fit a byte.
static qmcl_status build_pal8(qmcl_map *m) {
  m->pal8_valid = true;
  m->pal8_n = 0;
  if (!m->has_code || m->n_codes < 2) return QMCL_OK;
  if (m->geom.h > 65535u) return QMCL_OK;  // the build kernel takes one grid row per map row
  cudaStream_t s = m->stream;
  const int n = m->n_codes;  // codes 0 .. kmax and kmax + 1 = "nothing within range"
  std::vector<float> lut(static_cast<size_t>(n) + 2);
  QMCL_CUDA(copy_d2h(lut.data(), m->pw3_lut_f32.ptr, lut.size() * sizeof(float), s));
  QMCL_CUDA(cudaStreamSynchronize(s));
  // squared distances the exact EDT can produce: a^2 + b^2 <= kmax, and the clip code
  const int kmax = n - 2;
  std::vector<uint8_t> reachable(n, 0);
  for (long a = 0; a * a <= kmax; ++a)
    for (long b = a; a * a + b * b <= kmax; ++b) reachable[a * a + b * b] = 1;
  reachable[kmax + 1] = 1;
  std::vector<uint8_t> rank(n, 0xff);
  std::vector<float> pal;
  for (int k = 0; k < n; ++k) {
    if (!reachable[k]) continue;
    size_t j = 0;
    while (j < pal.size() && std::memcmp(&pal[j], &lut[k], sizeof(float)) != 0) ++j;
    if (j == pal.size()) {
      if (pal.size() == 254) return QMCL_OK;  // does not fit: PAL32 stays in charge
      pal.push_back(lut[k]);
    }
    rank[k] = static_cast<uint8_t>(j);
  }
  const int npal = static_cast<int>(pal.size());
  pal.push_back(lut[n]);      // outside the map (p_max)
  pal.push_back(0.0f);        // "nothing pending" entry of the software pipeline
  QMCL_TRY(m->pal8_rank.ensure(n));
  QMCL_TRY(m->pal8_lut.ensure(256));
  QMCL_TRY(m->code8.ensure(code8_map_elems(m->geom.w, m->geom.h)));
  QMCL_TRY(m->scan_tmp.ensure(8));
  int *d_flag = reinterpret_cast<int *>(m->scan_tmp.ptr);
  QMCL_CUDA(cudaMemsetAsync(d_flag, 0, sizeof(int), s));
  QMCL_CUDA(copy_h2d(m->pal8_rank.ptr, rank.data(), rank.size(), s));
  QMCL_CUDA(copy_h2d(m->pal8_lut.ptr, pal.data(), pal.size() * sizeof(float), s));
  dim3 grid(div_up(m->geom.w, 256), m->geom.h);
  code8_build_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const uint16_t *>(m->code.ptr),
                                          m->pal8_rank.ptr, m->code8.ptr, m->geom.w, m->geom.h,
                                          d_flag);
  QMCL_LAUNCHED();
  int flag = 1;
  QMCL_CUDA(copy_d2h(&flag, d_flag, sizeof(flag), s));
  QMCL_CUDA(cudaStreamSynchronize(s));  // rank / pal are host vectors: done with them now
  if (!flag) m->pal8_n = npal;
  return QMCL_OK;
}

// Called at map load: decides whether the 3-op division may be used.  The check
// depends on the geometry only, so the load launches it ahead of its own work
// (_launch) and picks the verdict up at its final synchronisation (_finish).
qmcl_status sensor_prepare_map_launch(qmcl_map *m) {
  SensorParams sp = make_params(m);
  QMCL_TRY(m->h_count.ensure(2));
  QMCL_TRY(m->free_totals.ensure(2));
  int *d_flag = reinterpret_cast<int *>(m->free_totals.ptr + (m->free_totals.cap - 1));  // last word: never a slab total
  QMCL_CUDA(cudaMemsetAsync(d_flag, 0, 2 * sizeof(int), m->stream));
  const unsigned threads = (std::max(m->geom.w, m->geom.h) + 8) * 9;
  fastdiv_check_kernel<<<div_up(threads, 256), 256, 0, m->stream>>>(sp, d_flag);
  QMCL_LAUNCHED();
  QMCL_CUDA(copy_d2h(m->h_count.ptr + 1, d_flag, 2 * sizeof(int), m->stream));
  return QMCL_OK;
}
qmcl_status sensor_prepare_map_finish(qmcl_map *m) {
  QMCL_CUDA(cudaStreamSynchronize(m->stream));
  int flag[2];
  memcpy(flag, m->h_count.ptr + 1, sizeof(flag));
  m->div_verified = flag[0] ? DIV_IEEE : (flag[1] ? DIV_MARKSTEIN : DIV_SPLIT);
  m->lut_params_valid = false;
  return QMCL_OK;
}
qmcl_status sensor_prepare_map(qmcl_map *m) {
  QMCL_TRY(sensor_prepare_map_launch(m));
  return sensor_prepare_map_finish(m);
}

// A batch launch: dev_jobs[n_jobs] (device memory), the widest job needs grid_x blocks.
struct SensorBatch {
  const SensorLaunch *dev_jobs = nullptr;
  unsigned n_jobs = 0, grid_x = 0;
};

template <int PATH, int DIV>
static qmcl_status launch_variant(const qmcl_map *m, const SensorParams &sp, const SensorLaunch &a,
                                  size_t smem, const SensorBatch &batch) {
  if (batch.dev_jobs) {
    auto kern = sensor_batch_kernel<PATH, DIV>;
    if (smem > 48 * 1024)
      QMCL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem)));
    kern<<<dim3(batch.grid_x, batch.n_jobs), kSensorThreads, smem, m->stream>>>(m->state.ptr, sp,
                                                                                batch.dev_jobs);
    QMCL_LAUNCHED();
    return QMCL_OK;
  }
  auto kern = sensor_kernel<PATH, DIV>;
  if (smem > 48 * 1024)
    QMCL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   static_cast<int>(smem)));
  kern<<<sensor_grid(a.n), kSensorThreads, smem, m->stream>>>(m->state.ptr, sp, a);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

static qmcl_status launch_sensor_impl(qmcl_map *m, const SensorLaunch &a, const SensorBatch &batch);

qmcl_status launch_sensor(qmcl_map *m, const SensorLaunch &a) {
  if (a.n == 0) return QMCL_OK;
  return launch_sensor_impl(m, a, SensorBatch{});
}

// K7 for n_jobs filters that share map `m` in one launch (dev_jobs in device
// memory; max_particles = the largest particle count among them).
qmcl_status launch_sensor_batch(qmcl_map *m, const SensorLaunch *dev_jobs, unsigned n_jobs,
                                size_t max_particles) {
  if (!n_jobs || !max_particles) return QMCL_OK;
  SensorBatch b;
  b.dev_jobs = dev_jobs;
  b.n_jobs = n_jobs;
  b.grid_x = sensor_grid(max_particles);
  return launch_sensor_impl(m, SensorLaunch{}, b);
}

static qmcl_status launch_sensor_impl(qmcl_map *m, const SensorLaunch &a, const SensorBatch &batch) {
  SensorParams sp = make_params(m);
  // sensor_path: 0 auto (PAL8, else PAL32, when the map was built here), 1 FIELD, 2 PAL64,
  // 3 PAL8 (same as auto), 4 PAL32
  const bool can_palette =
      m->has_code && static_cast<size_t>(m->n_codes + 2) * sizeof(double) <= 96 * 1024;
  // division: the cheapest form the self-check allowed, or the one forced
  // through qmcl_map_set_division_mode (never a form the check rejected)
  int div = m->div_verified;
  if (m->div_forced >= 0) div = std::min(div, m->div_forced);
  int path = PATH_FIELD;
  if (can_palette && (m->sensor_path == 0 || m->sensor_path == 3 || m->sensor_path == 4))
    path = PATH_PAL32;
  if (can_palette && m->sensor_path == 2) path = PATH_PAL64;
  // 0 (auto) and 3 take the one-byte palette when it exists for these parameters
  const bool want_pal8 = path == PATH_PAL32 && m->sensor_path != 4;
  if (path != PATH_FIELD) {
    // refresh pw^3 when z_hit / z_rand / max range / p_max changed
    if (!m->lut_params_valid || m->lut_z_hit != sp.z_hit || m->lut_zr != sp.zr ||
        m->lut_p_max != sp.p_max) {
      QMCL_TRY(m->pw3_lut.ensure(m->n_codes + 2));
      QMCL_TRY(m->pw3_lut_f32.ensure(m->n_codes + 2));
      sp.pw3_lut = m->pw3_lut.ptr;
      sp.pw3_lut_f32 = m->pw3_lut_f32.ptr;
      pw3_lut_kernel<<<div_up(m->n_codes + 2, 256), 256, 0, m->stream>>>(
          m->lut.ptr + m->n_codes, m->n_codes, sp.z_hit, sp.zr, sp.p_max, m->pw3_lut.ptr,
          m->pw3_lut_f32.ptr);
      QMCL_LAUNCHED();
      m->lut_params_valid = true;
      m->lut_z_hit = sp.z_hit;
      m->lut_zr = sp.zr;
      m->lut_p_max = sp.p_max;
      m->pal8_valid = false;
    }
    if (want_pal8) {
      if (!m->pal8_valid) QMCL_TRY(build_pal8(m));
      if (m->pal8_n > 0) {
        sp.code8 = m->code8.ptr;
        sp.pal8_lut = m->pal8_lut.ptr;
        sp.pal8_n = m->pal8_n;
        const size_t smem8 = static_cast<size_t>(m->pal8_n + 2) * sizeof(float);
        if (div == DIV_SPLIT) return launch_variant<PATH_PAL8, DIV_SPLIT>(m, sp, a, smem8, batch);
        return div == DIV_MARKSTEIN ? launch_variant<PATH_PAL8, DIV_MARKSTEIN>(m, sp, a, smem8, batch)
                                    : launch_variant<PATH_PAL8, DIV_IEEE>(m, sp, a, smem8, batch);
      }
    }
    if (path == PATH_PAL64) {
      const size_t smem = static_cast<size_t>(m->n_codes + 2) * sizeof(double);
      return div >= DIV_MARKSTEIN ? launch_variant<PATH_PAL64, DIV_MARKSTEIN>(m, sp, a, smem, batch)
                                  : launch_variant<PATH_PAL64, DIV_IEEE>(m, sp, a, smem, batch);
    }
    const size_t smem = static_cast<size_t>(m->n_codes + 2) * sizeof(float);
    if (div == DIV_SPLIT) return launch_variant<PATH_PAL32, DIV_SPLIT>(m, sp, a, smem, batch);
    return div == DIV_MARKSTEIN ? launch_variant<PATH_PAL32, DIV_MARKSTEIN>(m, sp, a, smem, batch)
                                : launch_variant<PATH_PAL32, DIV_IEEE>(m, sp, a, smem, batch);
  }
  return div >= DIV_MARKSTEIN ? launch_variant<PATH_FIELD, DIV_MARKSTEIN>(m, sp, a, 0, batch)
                              : launch_variant<PATH_FIELD, DIV_IEEE>(m, sp, a, 0, batch);
}

}  // namespace qmcl
