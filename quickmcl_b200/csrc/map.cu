// map.cu — likelihood-field map on the device.
//
// Replaces quickmcl::MapBase / MapLikelihood / distance_calculator
// (src/quickmcl/map_base.cpp, map_likelihood.cpp, distance_calculator.cpp).
// Field build = K1 tri-state map, K2 exact clipped Euclidean distance
// transform (separable, integer d^2), K3 Gaussian table through a d^2 LUT,
// K4 free-cell compaction.  All HBM-bound; see DESIGN.md for the byte counts.

#include "scan.cuh"
#include "sensor.cuh"

#include <algorithm>
#include <string>
#include <vector>

namespace qmcl {

std::atomic<uint64_t> g_launches{0};
std::atomic<uint64_t> g_h2d_bytes{0}, g_d2h_bytes{0};
std::atomic<uint32_t> g_scan_epoch{1};
static thread_local std::string g_error;

void set_error(const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
}

qmcl_status use_device(const qmcl_map *m) {
  QMCL_CUDA(cudaSetDevice(m->device));
  return QMCL_OK;
}

// ------------------------------------------------------------------ kernels

constexpr int kEdtInf = 0x3fff;  // "no obstacle within reach" marker for g (u16)
constexpr int kSegLen = 128;     // rows per column segment in pass 1

// K1 + K2 pass 1.  One thread per (column, row segment).  The clipped
// transform only needs obstacles within R rows, so a segment is exact after
// sweeping [y0-R, y1+R).  Down sweep writes g and the tri-state map
// (distance_calculator.cpp:184-209: -1 -> Unknown, <35 -> Free, >65 ->
// Occupied, else Unknown); up sweep folds in the distance from below.
// Warp lanes are adjacent columns, so every load/store is coalesced.
__global__ void __launch_bounds__(128)
edt_columns_kernel(const int8_t *__restrict__ occ, int8_t *__restrict__ state,
                   uint16_t *__restrict__ g, int w, int h, int R, int seg_first) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= w) return;
  const int y0 = (seg_first + blockIdx.y) * kSegLen;
  const int y1 = min(h, y0 + kSegLen);
  int d = kEdtInf;
  for (int y = max(0, y0 - R); y < y1; ++y) {
    const int8_t v = __ldg(occ + static_cast<size_t>(y) * w + x);
    d = (v > 65) ? 0 : min(d + 1, kEdtInf);
    if (y >= y0) {
      const size_t i = static_cast<size_t>(y) * w + x;
      g[i] = static_cast<uint16_t>(d);
      int8_t s;
      if (v == -1) s = 2;
      else if (v < 35) s = 0;
      else if (v > 65) s = 1;
      else s = 2;
      state[i] = s;
    }
  }
  d = kEdtInf;
  for (int y = min(h, y1 + R) - 1; y >= y0; --y) {
    const int8_t v = __ldg(occ + static_cast<size_t>(y) * w + x);
    d = (v > 65) ? 0 : min(d + 1, kEdtInf);
    if (y < y1) {
      const size_t i = static_cast<size_t>(y) * w + x;
      const int cur = g[i];
      if (d < cur) g[i] = static_cast<uint16_t>(d);
    }
  }
}

// Vectorised pass 1 for widths that are a multiple of 4 and R <= 253: a thread
// owns four adjacent columns (char4 loads = 128 B per warp instruction,
// ushort4 / char4 stores), the down-sweep distances of the segment stay in
// shared memory as bytes (anything above R is "out of reach" for pass 2), so g
// is written exactly once.  Traffic per cell: ~2.6 B of occupancy (the two
// sweeps re-read the R-row halo), 1 B state, 2 B g.
constexpr int kColSeg = 64;     // rows per segment
constexpr int kColThreads = 128;  // x 4 columns = 512 columns per block
__device__ __forceinline__ int8_t tri_state(int v) {
  // distance_calculator.cpp:194-206
  return (v == -1) ? 2 : (v < 35) ? 0 : (v > 65) ? 1 : 2;
}
__global__ void __launch_bounds__(kColThreads)
edt_columns4_kernel(const int8_t *__restrict__ occ, int8_t *__restrict__ state,
                    uint16_t *__restrict__ g, int w, int h, int R, int seg_first) {
  __shared__ uchar4 s_down[kColSeg][kColThreads];
  const int x = (blockIdx.x * kColThreads + threadIdx.x) * 4;
  if (x >= w) return;
  const int y0 = (seg_first + blockIdx.y) * kColSeg;
  const int y1 = min(h, y0 + kColSeg);
  int d0 = 255, d1 = 255, d2 = 255, d3 = 255;
#pragma unroll 4
  for (int y = max(0, y0 - R); y < y1; ++y) {
    const size_t i = static_cast<size_t>(y) * w + x;
    const char4 v = __ldg(reinterpret_cast<const char4 *>(occ + i));
    d0 = (v.x > 65) ? 0 : min(d0 + 1, 255);
    d1 = (v.y > 65) ? 0 : min(d1 + 1, 255);
    d2 = (v.z > 65) ? 0 : min(d2 + 1, 255);
    d3 = (v.w > 65) ? 0 : min(d3 + 1, 255);
    if (y >= y0) {
      s_down[y - y0][threadIdx.x] = make_uchar4(d0, d1, d2, d3);
      *reinterpret_cast<char4 *>(state + i) =
          make_char4(tri_state(v.x), tri_state(v.y), tri_state(v.z), tri_state(v.w));
    }
  }
  d0 = d1 = d2 = d3 = 255;
#pragma unroll 4
  for (int y = min(h, y1 + R) - 1; y >= y0; --y) {
    const size_t i = static_cast<size_t>(y) * w + x;
    const char4 v = __ldg(reinterpret_cast<const char4 *>(occ + i));
    d0 = (v.x > 65) ? 0 : min(d0 + 1, 255);
    d1 = (v.y > 65) ? 0 : min(d1 + 1, 255);
    d2 = (v.z > 65) ? 0 : min(d2 + 1, 255);
    d3 = (v.w > 65) ? 0 : min(d3 + 1, 255);
    if (y < y1) {
      const uchar4 dn = s_down[y - y0][threadIdx.x];
      ushort4 o;
      o.x = static_cast<uint16_t>(min(d0, static_cast<int>(dn.x)));
      o.y = static_cast<uint16_t>(min(d1, static_cast<int>(dn.y)));
      o.z = static_cast<uint16_t>(min(d2, static_cast<int>(dn.z)));
      o.w = static_cast<uint16_t>(min(d3, static_cast<int>(dn.w)));
      *reinterpret_cast<ushort4 *>(g + i) = o;
    }
  }
}

// Tri-state map only (used by load_field's callers that bring their own).
// d^2 -> (distance, likelihood) table.  distance_calculator.cpp:150-157 for
// the clipping, map_likelihood.cpp:60 for the Gaussian.  The last entry is the
// "nothing within range" cell.  exp is evaluated in double and rounded once
// (the correctly rounded float; Eigen's packet exp is 1-2 ulp off it).
__global__ void edt_lut_kernel(float *__restrict__ lut_dist, float *__restrict__ lut_field,
                               int n_entries, float res, float max_dist, float z_hit_pre) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_entries) return;
  float d = max_dist;
  if (k < n_entries - 1) {
    const float cand = __fmul_rn(sqrtf(static_cast<float>(k)), res);
    if (!(cand > max_dist)) d = cand;
  }
  lut_dist[k] = d;
  const float arg = __fdiv_rn(-__fmul_rn(d, d), z_hit_pre);
  lut_field[k] = static_cast<float>(exp(static_cast<double>(arg)));
}

// K2 pass 2 + K3.  One block per (row, 256-column chunk): G = g^2 of the chunk
// plus an R-cell halo is staged in shared memory; each thread takes the
// minimum of dx^2 + G[x+dx] over the window.  Chunks with no obstacle in reach
// skip the search.  The result indexes the LUT, so no per-cell sqrt/exp.
constexpr int kRowChunk = 256;
__global__ void __launch_bounds__(kRowChunk)
edt_rows_kernel(const uint16_t *__restrict__ g, const float *__restrict__ lut_dist,
                const float *__restrict__ lut_field, float *__restrict__ distance,
                float *__restrict__ field, uint16_t *__restrict__ code, int w, int h, int y_first,
                int R, int kmax) {
  extern __shared__ int sG[];  // kRowChunk + 2R entries
  const int y = y_first + blockIdx.y;
  const int x0 = blockIdx.x * kRowChunk;
  const int span = kRowChunk + 2 * R;
  const int big = 1 << 28;
  for (int i = threadIdx.x; i < span; i += kRowChunk) {
    const int xx = x0 - R + i;
    int v = big;
    if (xx >= 0 && xx < w) {
      const int gy = g[static_cast<size_t>(y) * w + xx];
      if (gy <= R) v = gy * gy;
    }
    sG[i] = v;
  }
  __syncthreads();
  // Does anything within reach of this warp's 32 cells have an obstacle above
  // or below?  Window of the warp: span entries [32*warp, 32*warp + 32 + 2R).
  const int lane = threadIdx.x & 31, wbase = threadIdx.x & ~31;
  int any = 0;
  for (int i = lane; i < 32 + 2 * R; i += 32) any |= sG[wbase + i] != big;
  any = __any_sync(0xffffffffu, any);
  const int x = x0 + threadIdx.x;
  if (x >= w) return;
  int best = big;
  if (any) {
    const int *c = sG + threadIdx.x + R;  // centre of this thread's window
    best = c[0];
    // outward: a term dx^2 + G cannot improve on `best` once dx^2 >= best
    for (int dx = 1; dx <= R; ++dx) {
      const int dd = dx * dx;
      if (dd >= best) break;
      best = min(best, min(c[-dx], c[dx]) + dd);
    }
  }
  const int k = (best > kmax) ? (kmax + 1) : best;
  const size_t i = static_cast<size_t>(y) * w + x;
  field[i] = __ldg(lut_field + k);
  if (distance) distance[i] = __ldg(lut_dist + k);
  // stored pre-multiplied by 4: a byte offset into the float pw^3 table (sensor.cu)
  if (code) code[code_offset(x, y, code_strip_stride(h))] = static_cast<uint16_t>(4 * k);
}

// K2 pass 2 + K3, two cells per thread on 16-bit lanes.  The candidates d^2 + G
// fit 16 bits (R <= 74 cells: G <= R^2, "none" = kInf16, sums below 65536), so a
// pair of adjacent cells shares every instruction: VIADDMNMX.U16x2 (DPX, native
// on sm_100a) adds the step's d^2 to both lanes and folds the minimum, and an
// odd offset's pair is one funnel shift of the two aligned words around it.  Per
// two outward steps and two cells: 2 LDS, 2 SHF, 4 VIADDMNMX — half the
// instructions per cell-candidate of edt_rows_kernel.  Same results (the minimum
// over the same window of the same integers).
constexpr int kRow2Cells = 512;  // cells per block: 256 threads x 2
constexpr unsigned kInf16 = 60000u;
constexpr int kRow2MaxR = 74;
__global__ void __launch_bounds__(256)
edt_rows2_kernel(const uint16_t *__restrict__ g, const float *__restrict__ lut_dist,
                 const float *__restrict__ lut_field, float *__restrict__ distance,
                 float *__restrict__ field, uint16_t *__restrict__ code, int w, int h, int y_first,
                 int R, int kmax) {
  extern __shared__ unsigned sW[];  // (kRow2Cells + 2 Rp) / 2 + 2 words of uint16 pairs
  uint16_t *sG16 = reinterpret_cast<uint16_t *>(sW);
  const int Rp = (R + 1) & ~1;  // even halo: every thread's pair is word-aligned
  const int y = y_first + blockIdx.y;
  const int x0 = blockIdx.x * kRow2Cells;
  const int span = kRow2Cells + 2 * Rp + 4;
  for (int i = threadIdx.x; i < span; i += 256) {
    const int xx = x0 - Rp + i;
    unsigned v = kInf16;
    if (xx >= 0 && xx < w) {
      const int gy = g[static_cast<size_t>(y) * w + xx];
      if (gy <= R) v = static_cast<unsigned>(gy * gy);
    }
    sG16[i] = static_cast<uint16_t>(v);
  }
  __syncthreads();
  // anything within reach of this warp's 64 cells?
  const int lane = threadIdx.x & 31, wbase = (threadIdx.x & ~31) * 2;
  int any = 0;
  for (int i = lane; i < (64 + 2 * Rp) / 2; i += 32) {
    const unsigned p = sW[wbase / 2 + i];
    any |= (p & 0xffffu) != kInf16 || (p >> 16) != kInf16;
  }
  any = __any_sync(0xffffffffu, any);
  const int x = x0 + 2 * threadIdx.x;
  if (x >= w) return;
  const int cw = (Rp + 2 * threadIdx.x) / 2;  // word holding this thread's two cells
  unsigned best2 = sW[cw];
  if (any) {
    unsigned Rw = best2, Lw = best2;
    for (int k = 0;; ++k) {
      const unsigned mx = max(best2 & 0xffffu, best2 >> 16);
      if (k >= 1) {
        const int d = 2 * k;
        const unsigned dd = static_cast<unsigned>(d * d);
        if (d > R || dd >= mx) break;
        const unsigned dd2 = dd | (dd << 16);
        best2 = __viaddmin_u16x2(Rw, dd2, best2);
        best2 = __viaddmin_u16x2(Lw, dd2, best2);
      }
      const int d = 2 * k + 1;
      const unsigned dd = static_cast<unsigned>(d * d);
      if (d > R || dd >= mx) break;
      const unsigned dd2 = dd | (dd << 16);
      const unsigned Rn = sW[cw + k + 1], Ln = sW[cw - k - 1];
      best2 = __viaddmin_u16x2(__funnelshift_r(Rw, Rn, 16), dd2, best2);  // cells x+d, x+1+d
      best2 = __viaddmin_u16x2(__funnelshift_r(Ln, Lw, 16), dd2, best2);  // cells x-d, x+1-d
      Rw = Rn;
      Lw = Ln;
    }
  }
  const int b0 = static_cast<int>(best2 & 0xffffu), b1 = static_cast<int>(best2 >> 16);
  const int k0 = (b0 > kmax) ? (kmax + 1) : b0;
  const int k1 = (b1 > kmax) ? (kmax + 1) : b1;
  const size_t i = static_cast<size_t>(y) * w + x;
  const bool two = x + 1 < w;
  if (two && (i & 1) == 0) {
    *reinterpret_cast<float2 *>(field + i) = make_float2(__ldg(lut_field + k0), __ldg(lut_field + k1));
    if (distance) *reinterpret_cast<float2 *>(distance + i) = make_float2(__ldg(lut_dist + k0), __ldg(lut_dist + k1));
  } else {
    field[i] = __ldg(lut_field + k0);
    if (distance) distance[i] = __ldg(lut_dist + k0);
    if (two) {
      field[i + 1] = __ldg(lut_field + k1);
      if (distance) distance[i + 1] = __ldg(lut_dist + k1);
    }
  }
  if (code) {
    // x is even, so both cells sit in one 8-cell strip, next to each other
    const uint32_t off = code_offset(x, y, code_strip_stride(h));
    if (two) *reinterpret_cast<ushort2 *>(code + off) = make_ushort2(static_cast<uint16_t>(4 * k0), static_cast<uint16_t>(4 * k1));
    else code[off] = static_cast<uint16_t>(4 * k0);
  }
}

// Distance field on demand (qmcl_map_copy_distance): code -> metres.
__global__ void __launch_bounds__(256)
distance_from_code_kernel(const uint16_t *__restrict__ code, const float *__restrict__ lut_dist,
                          uint32_t w, uint32_t h, float *__restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= static_cast<size_t>(w) * h) return;
  const uint32_t y = static_cast<uint32_t>(i / w), x = static_cast<uint32_t>(i - size_t(y) * w);
  out[i] = __ldg(lut_dist + (code[code_offset(x, y, code_strip_stride(h))] >> 2));
}

// qmcl_map_copy_free_cells: (x, y) pairs as the reference's Vector2i list.
__global__ void __launch_bounds__(256)
free_cells_xy_kernel(const uint32_t *__restrict__ idx, size_t n_free, unsigned w,
                     int32_t *__restrict__ out_xy) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_free) return;
  const unsigned c = idx[i];
  out_xy[2 * i] = static_cast<int32_t>(c % w);
  out_xy[2 * i + 1] = static_cast<int32_t>(c / w);
}

// map_likelihood.cpp:132-146: int8(100 * p + 0.5f)
__global__ void likelihood_grid_kernel(const float *__restrict__ field, size_t n,
                                       int8_t *__restrict__ out) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = static_cast<int8_t>(__fadd_rn(__fmul_rn(100.f, field[i]), 0.5f));
}

// map_base.cpp:66-77 for one query point.
__global__ void state_at_kernel(const int8_t *__restrict__ state, MapGeom g, float x, float y,
                                int8_t *out) {
  const int ix = cell_index(x, g.ox, g.res), iy = cell_index(y, g.oy, g.res);
  *out = in_map(ix, iy, g) ? state[static_cast<size_t>(iy) * g.w + ix] : int8_t(2);
}

__global__ void random_pose_kernel(const uint32_t *__restrict__ free_xy, unsigned long long n_free,
                                   MapGeom g, uint64_t seed, uint32_t step, uint64_t index,
                                   float *out) {
  uint32_t r[4];
  rng_draw(seed, step, P_GLOBAL, index, r);
  random_free_pose(free_xy, n_free, g, r, out, out + 1, out + 2);
}

// ---------------------------------------------------------------- host side

// K4: the free-cell list (map_base.cpp:46-55, row-major order) by one single-pass
// compaction (scan.cuh, decoupled look-back).  The count stays on the device
// until the caller's next synchronisation (read_free_count).
struct FreeFlag {
  const int8_t *state;
  __device__ uint32_t operator()(size_t i) const { return state[i] == 0 ? 1u : 0u; }
};
struct EmitFree {
  uint32_t *out;
  __device__ void operator()(size_t i, uint32_t pos, uint32_t flag) const {
    if (flag) out[pos] = static_cast<uint32_t>(i);
  }
};
// (One scan item per cell.  Eight cells per item — one 8-byte load, up to eight
// indices written by one thread — was tried and is slower: 0.35 -> ~0.9 ms at
// 8192^2, the stores of a warp then spread over 32 x 256 bytes.)
static qmcl_status rebuild_free_cells(qmcl_map *m) {
  const size_t n = static_cast<size_t>(m->geom.w) * m->geom.h;
  QMCL_TRY(m->free_cells.ensure(n + 2));  // worst case: every cell is free
  QMCL_TRY(m->scan_tmp.ensure_zero(scan_tmp_words(n) + 8));
  QMCL_TRY(m->h_count.ensure(2));
  unsigned long long *d_total = scan_total_ptr(m->scan_tmp.ptr, n);
  QMCL_TRY(device_scan(m->stream, FreeFlag{m->state.ptr}, EmitFree{m->free_cells.ptr}, n, m->scan_tmp.ptr,
                       d_total));
  QMCL_CUDA(copy_d2h(m->h_count.ptr, d_total, sizeof(unsigned long long), m->stream));
  m->n_free = 0;  // read_free_count fills it in after the next synchronisation
  return QMCL_OK;
}
static void read_free_count(qmcl_map *m) { m->n_free = m->h_count.ptr[0]; }

static void set_geometry(qmcl_map *m, uint32_t w, uint32_t h, float res, double ox, double oy) {
  // map_base.cpp:37-40: offset = (float)origin, res = (float)resolution
  m->geom.w = w;
  m->geom.h = h;
  m->geom.res = res;
  m->geom.ox = static_cast<float>(ox);
  m->geom.oy = static_cast<float>(oy);
}

}  // namespace qmcl

using namespace qmcl;

// ==================================================================== C ABI
extern "C" {

const char *qmcl_version(void) { return "quickmcl_b200 0.1 (sm_100a)"; }
const char *qmcl_last_error(void) { return qmcl::g_error.c_str(); }
uint64_t qmcl_kernel_launch_count(void) { return g_launches.load(); }
qmcl_status qmcl_transfer_bytes(uint64_t out[2]) {
  if (!out) return QMCL_ERR_INVALID_ARG;
  out[0] = g_h2d_bytes.load();
  out[1] = g_d2h_bytes.load();
  return QMCL_OK;
}
const char *qmcl_stage_name(int stage) {
  static const char *names[QMCL_N_STAGES] = {"motion", "sensor", "normalise", "cdf", "select",
                                             "gather", "kld", "bins", "label",
                                             "post_normalise", "moments", "tail"};
  return (stage >= 0 && stage < QMCL_N_STAGES) ? names[stage] : "?";
}

qmcl_status qmcl_map_create(int device, qmcl_map **out) {
  if (!out) return QMCL_ERR_INVALID_ARG;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    set_error("no CUDA device: this library has no CPU fallback");
    return QMCL_ERR_NO_DEVICE;
  }
  if (device < 0) QMCL_CUDA(cudaGetDevice(&device));
  if (device >= count) return QMCL_ERR_INVALID_ARG;
  qmcl_map *m = new qmcl_map();
  m->device = device;
  QMCL_CUDA(cudaSetDevice(device));
  qmcl_map_set_params(m, &m->params);
  *out = m;
  return QMCL_OK;
}

// Filters keep a reference to their map (the node owns both through
// shared_ptr, main.h:68-70), so the storage goes when the last holder lets go.
qmcl_status qmcl_map_destroy(qmcl_map *m) {
  if (!m) return QMCL_ERR_INVALID_ARG;
  if (--m->refs > 0) return QMCL_OK;
  cudaSetDevice(m->device);
  cudaStreamSynchronize(m->stream);
  m->field.release();
  m->distance.release();
  m->state.release();
  m->free_cells.release();
  m->occ.release();
  m->edt_g.release();
  m->scan_tmp.release();
  m->lut.release();
  m->code.release();
  m->pw3_lut.release();
  m->pw3_lut_f32.release();
  m->code8.release();
  m->pal8_rank.release();
  m->pal8_lut.release();
  m->tmp_x.release();
  m->tmp_y.release();
  m->tmp_t.release();
  m->tmp_w.release();
  m->tmp_aos.release();
  m->tmp_cloud.release();
  m->tmp_scalars.release();
  m->h_count.release();
  for (cudaEvent_t e : m->slab_events) cudaEventDestroy(e);
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  delete m;
  return QMCL_OK;
}

qmcl_status qmcl_map_set_sensor_path(qmcl_map *m, int path) {
  if (!m || path < 0 || path > 4) return QMCL_ERR_INVALID_ARG;
  m->sensor_path = path;
  return QMCL_OK;
}

qmcl_status qmcl_map_set_division_mode(qmcl_map *m, int mode) {
  if (!m || mode < -1 || mode > 2) return QMCL_ERR_INVALID_ARG;
  m->div_forced = mode;
  return QMCL_OK;
}

qmcl_status qmcl_map_get_division_mode(const qmcl_map *m, int *verified) {
  if (!m || !verified) return QMCL_ERR_INVALID_ARG;
  *verified = m->div_verified;
  return QMCL_OK;
}

qmcl_status qmcl_map_set_stream(qmcl_map *m, void *cuda_stream) {
  if (!m) return QMCL_ERR_INVALID_ARG;
  m->stream = static_cast<cudaStream_t>(cuda_stream);
  return QMCL_OK;
}

// map_likelihood.cpp:40-49.  All float arithmetic as in the reference; the one
// exp is the correctly rounded float (computed in double, rounded once).
qmcl_status qmcl_map_set_params(qmcl_map *m, const qmcl_likelihood_params *p) {
  if (!m || !p) return QMCL_ERR_INVALID_ARG;
  m->params = *p;
  m->z_hit_pre = 2 * p->sigma_hit * p->sigma_hit;
  const float arg = -(p->max_obstacle_distance * p->max_obstacle_distance) / m->z_hit_pre;
  m->max_distance_probability = static_cast<float>(exp(static_cast<double>(arg)));
  return QMCL_OK;
}

qmcl_status qmcl_map_load(qmcl_map *m, const int8_t *occ, uint32_t width, uint32_t height,
                          float resolution, double origin_x, double origin_y) {
  if (!m || !occ || width == 0 || height == 0 || !(resolution > 0)) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(use_device(m));
  const size_t n = static_cast<size_t>(width) * height;
  const float max_dist = m->params.max_obstacle_distance;
  // Clipping radius in cells: one more than distance_calculator.cpp:89's
  // cache size so that every cell with d <= max_dist is inside the window.
  const long Rl = static_cast<long>(max_dist / resolution) + 1;
  if (Rl < 1 || Rl > 8000) {
    set_error("max_obstacle_distance / resolution = %ld cells is out of range", Rl);
    return QMCL_ERR_INVALID_ARG;
  }
  const int R = static_cast<int>(Rl);
  const int kmax = 2 * R * R;
  set_geometry(m, width, height, resolution, origin_x, origin_y);
  QMCL_TRY(m->occ.ensure(n));
  QMCL_TRY(m->state.ensure(n));
  QMCL_TRY(m->field.ensure(n));
  QMCL_TRY(m->edt_g.ensure((n + 1) / 2));
  QMCL_TRY(m->lut.ensure(2 * static_cast<size_t>(kmax + 2)));
  // the uint16 d^2 code map (sensor model's palette path) exists when d^2 fits
  const bool with_code = 4 * (kmax + 4) <= 65535;
  if (with_code) QMCL_TRY(m->code.ensure(code_map_elems(width, height)));
  else QMCL_TRY(m->distance.ensure(n));  // with codes the distance field is derived on demand
  uint16_t *g = reinterpret_cast<uint16_t *>(m->edt_g.ptr);
  float *lut_dist = m->lut.ptr, *lut_field = m->lut.ptr + (kmax + 2);
  edt_lut_kernel<<<div_up(kmax + 2, 256), 256, 0, m->stream>>>(lut_dist, lut_field, kmax + 2,
                                                               resolution, max_dist, m->z_hit_pre);
  QMCL_LAUNCHED();
  // The occupancy grid crosses PCIe in row slabs on a second stream while the
  // slabs already here are transformed: a slab's column pass needs R rows of the
  // next one, so it waits for that slab's copy; its row pass follows it.
  const bool vec4 = width % 4 == 0 && R <= 253;
  const int seg = vec4 ? kColSeg : kSegLen;
  int n_slabs = static_cast<int>(std::min<uint32_t>(16, std::max<uint32_t>(1, height / 1024)));
  int slab_rows = static_cast<int>(div_up(div_up(height, n_slabs), seg)) * seg;
  if (slab_rows < R + seg) {  // a slab must cover the halo of its neighbour
    n_slabs = 1;
    slab_rows = static_cast<int>(div_up(height, seg)) * seg;
  }
  n_slabs = static_cast<int>(div_up(height, slab_rows));
  if (!m->copy_stream) QMCL_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  while (m->slab_events.size() < static_cast<size_t>(n_slabs) + 1) {
    cudaEvent_t e = nullptr;
    QMCL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    m->slab_events.push_back(e);
  }
  // the copies may not overtake work already queued on the map's stream that reads `occ`
  QMCL_CUDA(cudaEventRecord(m->slab_events[n_slabs], m->stream));
  QMCL_CUDA(cudaStreamWaitEvent(m->copy_stream, m->slab_events[n_slabs], 0));
  for (int sl = 0; sl < n_slabs; ++sl) {
    const size_t r0 = static_cast<size_t>(sl) * slab_rows;
    const size_t r1 = std::min<size_t>(height, r0 + slab_rows);
    QMCL_CUDA(copy_h2d(m->occ.ptr + r0 * width, occ + r0 * width, (r1 - r0) * width, m->copy_stream));
    QMCL_CUDA(cudaEventRecord(m->slab_events[sl], m->copy_stream));
  }
  const bool rows2 = R <= kRow2MaxR;
  const size_t smem2 = ((kRow2Cells + 2 * static_cast<size_t>((R + 1) & ~1)) / 2 + 4) * sizeof(unsigned);
  const size_t smem1 = (kRowChunk + 2 * static_cast<size_t>(R)) * sizeof(int);
  if (!rows2) {
    if (smem1 > 200 * 1024) return QMCL_ERR_INVALID_ARG;
    if (smem1 > 48 * 1024)
      QMCL_CUDA(cudaFuncSetAttribute(edt_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem1)));
  }
  for (int sl = 0; sl < n_slabs; ++sl) {
    const int r0 = sl * slab_rows;
    const int r1 = static_cast<int>(std::min<size_t>(height, static_cast<size_t>(r0) + slab_rows));
    QMCL_CUDA(cudaStreamWaitEvent(m->stream, m->slab_events[std::min(sl + 1, n_slabs - 1)], 0));
    const unsigned segs = div_up(r1 - r0, seg);
    if (vec4) {
      dim3 grid1(div_up(width, 4 * kColThreads), segs);
      edt_columns4_kernel<<<grid1, kColThreads, 0, m->stream>>>(
          m->occ.ptr, m->state.ptr, g, static_cast<int>(width), static_cast<int>(height), R, r0 / seg);
    } else {
      dim3 grid1(div_up(width, 128), segs);
      edt_columns_kernel<<<grid1, 128, 0, m->stream>>>(m->occ.ptr, m->state.ptr, g,
                                                       static_cast<int>(width),
                                                       static_cast<int>(height), R, r0 / seg);
    }
    QMCL_LAUNCHED();
    if (rows2) {
      dim3 grid2(div_up(width, kRow2Cells), r1 - r0);
      edt_rows2_kernel<<<grid2, 256, smem2, m->stream>>>(
          g, lut_dist, lut_field, with_code ? nullptr : m->distance.ptr, m->field.ptr,
          with_code ? m->code.ptr : nullptr, static_cast<int>(width), static_cast<int>(height), r0, R, kmax);
    } else {
      dim3 grid2(div_up(width, kRowChunk), r1 - r0);
      edt_rows_kernel<<<grid2, kRowChunk, smem1, m->stream>>>(
          g, lut_dist, lut_field, with_code ? nullptr : m->distance.ptr, m->field.ptr,
          with_code ? m->code.ptr : nullptr, static_cast<int>(width), static_cast<int>(height), r0, R, kmax);
    }
    QMCL_LAUNCHED();
  }
  QMCL_TRY(rebuild_free_cells(m));
  m->has_map = true;
  m->has_distance = true;
  m->has_code = with_code;
  m->n_codes = kmax + 2;
  QMCL_TRY(sensor_prepare_map(m));  // synchronises: the free-cell count has arrived with it
  read_free_count(m);
  return QMCL_OK;
}

qmcl_status qmcl_map_load_field(qmcl_map *m, const float *field, const int8_t *state,
                                uint32_t width, uint32_t height, float resolution,
                                double origin_x, double origin_y) {
  if (!m || !field || !state || width == 0 || height == 0 || !(resolution > 0))
    return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(use_device(m));
  const size_t n = static_cast<size_t>(width) * height;
  set_geometry(m, width, height, resolution, origin_x, origin_y);
  QMCL_TRY(m->state.ensure(n));
  QMCL_TRY(m->field.ensure(n));
  QMCL_CUDA(copy_h2d(m->field.ptr, field, n * sizeof(float), m->stream));
  QMCL_CUDA(copy_h2d(m->state.ptr, state, n, m->stream));
  QMCL_TRY(rebuild_free_cells(m));
  m->has_map = true;
  m->has_distance = false;
  m->has_code = false;
  m->n_codes = 0;
  QMCL_TRY(sensor_prepare_map(m));  // synchronises: the free-cell count has arrived with it
  read_free_count(m);
  return QMCL_OK;
}

qmcl_status qmcl_map_size(const qmcl_map *m, uint32_t *width, uint32_t *height) {
  if (!m || !width || !height) return QMCL_ERR_INVALID_ARG;
  *width = m->geom.w;
  *height = m->geom.h;
  return QMCL_OK;
}

static qmcl_status copy_out(const qmcl_map *m, const void *src, void *dst, size_t bytes) {
  if (!m || !dst) return QMCL_ERR_INVALID_ARG;
  if (!m->has_map) return QMCL_ERR_NO_MAP;
  QMCL_TRY(use_device(m));
  QMCL_CUDA(copy_d2h(dst, src, bytes, m->stream));
  QMCL_CUDA(cudaStreamSynchronize(m->stream));
  return QMCL_OK;
}

qmcl_status qmcl_map_copy_field(const qmcl_map *m, float *out) {
  if (!m) return QMCL_ERR_INVALID_ARG;
  return copy_out(m, m->field.ptr, out, sizeof(float) * m->geom.w * m->geom.h);
}
qmcl_status qmcl_map_copy_distance(const qmcl_map *m, float *out) {
  if (!m) return QMCL_ERR_INVALID_ARG;
  if (!m->has_distance) return QMCL_ERR_NO_MAP;
  const size_t n = static_cast<size_t>(m->geom.w) * m->geom.h;
  if (m->has_code) {
    qmcl_map *mm = const_cast<qmcl_map *>(m);
    QMCL_TRY(use_device(m));
    QMCL_TRY(mm->distance.ensure(n));
    distance_from_code_kernel<<<div_up(n, 256), 256, 0, m->stream>>>(
        reinterpret_cast<const uint16_t *>(m->code.ptr), m->lut.ptr, m->geom.w, m->geom.h,
        mm->distance.ptr);
    QMCL_LAUNCHED();
  }
  return copy_out(m, m->distance.ptr, out, sizeof(float) * n);
}
qmcl_status qmcl_map_copy_state(const qmcl_map *m, int8_t *out) {
  if (!m) return QMCL_ERR_INVALID_ARG;
  return copy_out(m, m->state.ptr, out, static_cast<size_t>(m->geom.w) * m->geom.h);
}
qmcl_status qmcl_map_num_free_cells(const qmcl_map *m, uint64_t *out) {
  if (!m || !out) return QMCL_ERR_INVALID_ARG;
  if (!m->has_map) return QMCL_ERR_NO_MAP;
  *out = m->n_free;
  return QMCL_OK;
}
qmcl_status qmcl_map_copy_free_cells(const qmcl_map *m, int32_t *xy_out) {
  if (!m) return QMCL_ERR_INVALID_ARG;
  if (m->n_free == 0) return m->has_map ? QMCL_OK : QMCL_ERR_NO_MAP;
  qmcl_map *mm = const_cast<qmcl_map *>(m);
  QMCL_TRY(use_device(m));
  QMCL_TRY(mm->tmp_cloud.ensure(2 * m->n_free + 2));  // staging (int32 pairs)
  int32_t *d_xy = reinterpret_cast<int32_t *>(mm->tmp_cloud.ptr);
  free_cells_xy_kernel<<<div_up(m->n_free, 256), 256, 0, m->stream>>>(m->free_cells.ptr, m->n_free,
                                                                      m->geom.w, d_xy);
  QMCL_LAUNCHED();
  return copy_out(m, d_xy, xy_out, sizeof(int32_t) * 2 * m->n_free);
}

qmcl_status qmcl_map_state_at(const qmcl_map *m, float x, float y, int8_t *out) {
  if (!m || !out) return QMCL_ERR_INVALID_ARG;
  if (!m->has_map) return QMCL_ERR_NO_MAP;
  QMCL_TRY(use_device(m));
  qmcl_map *mm = const_cast<qmcl_map *>(m);
  QMCL_TRY(mm->scan_tmp.ensure(8));
  int8_t *d_out = reinterpret_cast<int8_t *>(mm->scan_tmp.ptr);
  state_at_kernel<<<1, 1, 0, m->stream>>>(m->state.ptr, m->geom, x, y, d_out);
  QMCL_LAUNCHED();
  return copy_out(m, d_out, out, 1);
}

qmcl_status qmcl_map_likelihood_grid(const qmcl_map *m, int8_t *out) {
  if (!m || !out) return QMCL_ERR_INVALID_ARG;
  if (!m->has_map) return QMCL_ERR_NO_MAP;
  QMCL_TRY(use_device(m));
  qmcl_map *mm = const_cast<qmcl_map *>(m);
  const size_t n = static_cast<size_t>(m->geom.w) * m->geom.h;
  QMCL_TRY(mm->occ.ensure(n));
  likelihood_grid_kernel<<<div_up(n, 256), 256, 0, m->stream>>>(m->field.ptr, n, mm->occ.ptr);
  QMCL_LAUNCHED();
  return copy_out(m, mm->occ.ptr, out, n);
}

qmcl_status qmcl_map_generate_random_pose(const qmcl_map *m, uint64_t seed, uint32_t step,
                                          uint64_t index, float pose_out[3]) {
  if (!m || !pose_out) return QMCL_ERR_INVALID_ARG;
  if (!m->has_map || m->n_free == 0) return QMCL_ERR_NO_MAP;
  QMCL_TRY(use_device(m));
  qmcl_map *mm = const_cast<qmcl_map *>(m);
  QMCL_TRY(mm->scan_tmp.ensure(8));
  float *d_out = reinterpret_cast<float *>(mm->scan_tmp.ptr);
  random_pose_kernel<<<1, 1, 0, m->stream>>>(m->free_cells.ptr, m->n_free, m->geom, seed, step,
                                             index, d_out);
  QMCL_LAUNCHED();
  return copy_out(m, d_out, pose_out, 3 * sizeof(float));
}

}  // extern "C"
