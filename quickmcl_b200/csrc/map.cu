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
#include <cstdlib>
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

// ---- fast path of the field build (width % 4 == 0, R <= 74) -----------------
// The column distances fit a byte, the two sweeps of a column segment run on
// different warps of one block (half the dependent chain each), their loads are
// issued eight rows at a time, and the row pass takes eight rows per block so
// that its fixed costs (staging, barrier, stores) are paid once per 4096 cells.

constexpr int kColBSeg = 16;       // rows per column segment
constexpr int kColBThreads = 128;  // x 4 columns; the block has 2 x 128 threads (down / up sweep)
constexpr int kColBBatch = 8;      // rows loaded ahead of the dependent chain

// 0x80 in every byte of v that holds an occupied cell (v > 65 as int8,
// distance_calculator.cpp:200): low seven bits + 62 carry into bit 7 iff they are
// >= 66, and the sign bit must be clear.
__device__ __forceinline__ unsigned occupied_mask4(unsigned v) {
  return ((v & 0x7f7f7f7fu) + 0x3e3e3e3eu) & ~v & 0x80808080u;
}
// One sweep step on four columns: distance + 1, 0 on an obstacle.  "No obstacle
// yet" starts at 128 and a sweep is at most kColBSeg + R <= 90 rows long, so the
// bytes never carry and everything >= 128 means out of reach.
__device__ __forceinline__ unsigned sweep_step4(unsigned d, unsigned v) {
  const unsigned m = occupied_mask4(v);
  const unsigned keep = ~((m >> 7) * 0xffu);
  return (d + 0x01010101u) & keep;
}

__global__ void __launch_bounds__(2 * kColBThreads)
edt_columns_u8_kernel(const int8_t *__restrict__ occ, int8_t *__restrict__ state,
                      uint8_t *__restrict__ g8, int w, int R, int row_first, int row_end, int h) {
  __shared__ unsigned s_d[2][kColBSeg][kColBThreads];
  const int t = threadIdx.x & (kColBThreads - 1);
  const int up = threadIdx.x >> 7;  // warp-uniform
  const int x = (blockIdx.x * kColBThreads + t) * 4;
  const int y0 = row_first + blockIdx.y * kColBSeg;
  const int y1 = min(row_end, y0 + kColBSeg);
  if (x < w) {
    const unsigned *occ4 = reinterpret_cast<const unsigned *>(occ + x);
    const size_t w4 = static_cast<size_t>(w) >> 2;
    unsigned d = 0x80808080u;
    if (!up) {
      const int ya = max(0, y0 - R);
      for (int yb = ya; yb < y1; yb += kColBBatch) {
        unsigned v[kColBBatch];
#pragma unroll
        for (int j = 0; j < kColBBatch; ++j)
          v[j] = (yb + j < y1) ? __ldg(occ4 + static_cast<size_t>(yb + j) * w4) : 0u;
#pragma unroll
        for (int j = 0; j < kColBBatch; ++j) {
          const int y = yb + j;
          if (y < y1) {
            d = sweep_step4(d, v[j]);
            if (y >= y0) {
              s_d[0][y - y0][t] = d;
              const char4 c = *reinterpret_cast<const char4 *>(&v[j]);
              *reinterpret_cast<char4 *>(state + static_cast<size_t>(y) * w + x) =
                  make_char4(tri_state(c.x), tri_state(c.y), tri_state(c.z), tri_state(c.w));
            }
          }
        }
      }
    } else {
      const int ya = min(h, y1 + R) - 1;
      for (int yb = ya; yb >= y0; yb -= kColBBatch) {
        unsigned v[kColBBatch];
#pragma unroll
        for (int j = 0; j < kColBBatch; ++j)
          v[j] = (yb - j >= y0) ? __ldg(occ4 + static_cast<size_t>(yb - j) * w4) : 0u;
#pragma unroll
        for (int j = 0; j < kColBBatch; ++j) {
          const int y = yb - j;
          if (y >= y0) {
            d = sweep_step4(d, v[j]);
            if (y < y1) s_d[1][y - y0][t] = d;
          }
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (y1 - y0) * kColBThreads; i += 2 * kColBThreads) {
    const int r = i / kColBThreads, c = i % kColBThreads;
    const int xx = (blockIdx.x * kColBThreads + c) * 4;
    if (xx < w)
      *reinterpret_cast<unsigned *>(g8 + static_cast<size_t>(y0 + r) * w + xx) =
          __vminu4(s_d[0][r][c], s_d[1][r][c]);
  }
}

// Row pass of the fast path: edt_rows2_kernel's search on 16-bit lanes, four cells
// (two lane pairs A = x, x+1 and B = x+2, x+3) per thread and eight rows of 512
// cells per block.  g arrives as bytes and is squared while it is staged.  B's
// candidates at step d are A's at step d + 2, so one new word per side and
// iteration feeds both pairs; the exit test runs once per two outward steps (extra
// candidates inside the window never change the minimum).  A warp first marks which
// 4-cell groups of its window hold a column distance within reach at all (one ballot
// bit per group) and every thread starts its search at the nearest such group
// instead of walking through cells that cannot contribute.  The uint16 codes leave
// through shared memory so that every 128-byte line of the strip layout (8 x 8
// cells) is written whole.
constexpr int kRowBRows = 8;
constexpr int kRowBCells = 512;
constexpr int kRowBThreads = 256;  // two halves of 128 threads, four cells per thread
// 64 bits of the 128-bit value {hi, lo} starting at bit s (0 <= s < 128)
__device__ __forceinline__ unsigned long long bits_from128(unsigned long long lo, unsigned long long hi, int s) {
  if (s == 0) return lo;
  if (s < 64) return (lo >> s) | (hi << (64 - s));
  return hi >> (s - 64);
}
__device__ __forceinline__ bool pair_finite(unsigned p) {
  return (p & 0xffffu) != kInf16 || (p >> 16) != kInf16;
}
__global__ void __launch_bounds__(kRowBThreads)
edt_rows_u8_kernel(const uint8_t *__restrict__ g8, const float *__restrict__ lut_dist,
                   const float *__restrict__ lut_field, float *__restrict__ distance,
                   float *__restrict__ field, uint16_t *__restrict__ code, int w, int h,
                   int row_first, int row_end, int R, int kmax) {
  extern __shared__ unsigned sW[];  // kRowBRows x wpr words of uint16 pairs, then the code tile
  const int Rq = (R + 3) & ~3;      // halo in cells, a multiple of 4
  const int wpr = (kRowBCells + 2 * Rq) / 2;
  uint16_t *s_code = reinterpret_cast<uint16_t *>(sW + kRowBRows * wpr);  // [strip 64][row 8][8]
  const int x0 = blockIdx.x * kRowBCells;
  const int y0 = row_first + blockIdx.y * kRowBRows;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  {
    // warp r stages rows r and r + 4: one 4-cell word of g per lane and step
    const int gw = (kRowBCells + 2 * Rq) / 4;
    for (int rr = warp; rr < kRowBRows; rr += kRowBThreads / 32) {
      const int y = y0 + rr;
      for (int j = lane; j < gw; j += 32) {
        const int xx = x0 - Rq + 4 * j;
        unsigned v = 0xffffffffu;
        if (y < row_end && xx >= 0 && xx < w)
          v = __ldg(reinterpret_cast<const unsigned *>(g8 + static_cast<size_t>(y) * w + xx));
        unsigned G[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const unsigned gb = (v >> (8 * b)) & 0xffu;
          G[b] = (gb <= static_cast<unsigned>(R)) ? gb * gb : kInf16;
        }
        *reinterpret_cast<uint2 *>(sW + rr * wpr + 2 * j) =
            make_uint2(G[0] | (G[1] << 16), G[2] | (G[3] << 16));
      }
    }
  }
  __syncthreads();
  // threads 0-127 take rows 0-3 of the tile, threads 128-255 rows 4-7
  const int tx = threadIdx.x & (kRowBThreads / 2 - 1), half = threadIdx.x / (kRowBThreads / 2);
  const int x = x0 + 4 * tx;
  const int c0 = Rq / 2 + 2 * tx;  // word of pair A; pair B is the next one
  const uint32_t stride = code_strip_stride(h);
  const int k_full = R >> 1;       // iterations k <= k_full take both steps 2k-1 and 2k
  const int wlane = tx >> 5;       // this warp's 128-cell stretch of the row
#pragma unroll 1
  for (int r = half * (kRowBRows / 2); r < (half + 1) * (kRowBRows / 2); ++r) {
    const int y = y0 + r;
    if (y >= row_end) break;  // warp-uniform
    const unsigned *Wr = sW + r * wpr;
    // one bit per 4-cell group of this warp's window: bit i = words 64 * wlane + 2i, 2i + 1
    // (i < 32 + Rq / 2 <= 70)
    unsigned fm[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int i = 32 * q + lane;
      bool fin = false;
      if (i < 32 + Rq / 2) {
        const uint2 p = *reinterpret_cast<const uint2 *>(Wr + 64 * wlane + 2 * i);
        fin = pair_finite(p.x) || pair_finite(p.y);
      }
      fm[q] = __ballot_sync(0xffffffffu, fin);
    }
    unsigned bestA = Wr[c0], bestB = Wr[c0 + 1];
    const unsigned long long M01 = fm[0] | (static_cast<unsigned long long>(fm[1]) << 32);
    const unsigned long long M23 = fm[2];
    const int b = Rq / 4 + lane;  // this thread's group in the window, 1 <= b <= 50
    const unsigned long long right = bits_from128(M01, M23, b + 1);
    const unsigned long long left = M01 << (64 - b);
    const int jR = right ? __ffsll(static_cast<long long>(right)) : 1000;
    const int jL = left ? __clzll(static_cast<long long>(left)) + 1 : 1000;
    // group at offset j is first seen at step 4j - 3, i.e. in iteration 2j - 1; a
    // group that holds a distance itself starts at step 1 (its cells are each other's
    // neighbours)
    const bool own = pair_finite(bestA) || pair_finite(bestB);
    int kk = own ? 1 : 2 * min(jR, jL) - 1;
    if (2 * kk - 1 <= R) {
      const unsigned *pr = Wr + c0 + kk + 1, *pl = Wr + c0 - kk;  // next words outward
      unsigned r0 = pr[-2], r1 = pr[-1], l1 = pl[1], l2 = pl[2];
      const unsigned d0 = static_cast<unsigned>(2 * kk - 1);
      unsigned po = d0 * d0 * 0x10001u;         // d^2 on both lanes, odd step d = 2k-1
      unsigned inc = (2 * d0 + 1) * 0x10001u;   // (d+1)^2 - d^2
      // steps 2k-1 and 2k; the exit test sees the odd step's d^2 (ties walk one more
      // iteration: harmless)
#define QMCL_EDT_STEP2(R0, R1, R2, L0, L1, L2)                                   \
      {                                                                          \
        const unsigned m = __vimax3_u16x2(bestA, bestB, bestB);                          \
        if ((po & 0xffffu) >= max(m & 0xffffu, m >> 16)) break;                  \
        R2 = *pr++;                                                              \
        L0 = *pl--;                                                              \
        bestA = __viaddmin_u16x2(__funnelshift_r(R0, R1, 16), po, bestA);        \
        bestB = __viaddmin_u16x2(__funnelshift_r(R1, R2, 16), po, bestB);        \
        bestA = __viaddmin_u16x2(__funnelshift_r(L0, L1, 16), po, bestA);        \
        bestB = __viaddmin_u16x2(__funnelshift_r(L1, L2, 16), po, bestB);        \
        const unsigned pe = po + inc;                                            \
        bestA = __viaddmin_u16x2(R1, pe, bestA);                                 \
        bestB = __viaddmin_u16x2(R2, pe, bestB);                                 \
        bestA = __viaddmin_u16x2(L0, pe, bestA);                                 \
        bestB = __viaddmin_u16x2(L1, pe, bestB);                                 \
        po = pe + inc + 0x20002u;                                                \
        inc += 0x40004u;                                                         \
      }
      unsigned r2, l0;
      // three iterations per trip: the word registers rotate by renaming
      while (true) {
        if (kk > k_full) break;
        QMCL_EDT_STEP2(r0, r1, r2, l0, l1, l2)
        if (++kk > k_full) break;
        QMCL_EDT_STEP2(r1, r2, r0, l2, l0, l1)
        if (++kk > k_full) break;
        QMCL_EDT_STEP2(r2, r0, r1, l1, l2, l0)
        ++kk;
      }
#undef QMCL_EDT_STEP2
      // R odd: the last outward step stands alone
      if ((R & 1) && kk == k_full + 1) {
        const unsigned m = __vimax3_u16x2(bestA, bestB, bestB);
        if ((po & 0xffffu) < max(m & 0xffffu, m >> 16)) {
          const unsigned *qr = Wr + c0 + kk, *ql = Wr + c0 - kk;
          bestA = __viaddmin_u16x2(__funnelshift_r(qr[-1], qr[0], 16), po, bestA);
          bestB = __viaddmin_u16x2(__funnelshift_r(qr[0], qr[1], 16), po, bestB);
          bestA = __viaddmin_u16x2(__funnelshift_r(ql[0], ql[1], 16), po, bestA);
          bestB = __viaddmin_u16x2(__funnelshift_r(ql[1], ql[2], 16), po, bestB);
        }
      }
    }
    int k[4] = {static_cast<int>(bestA & 0xffffu), static_cast<int>(bestA >> 16),
                static_cast<int>(bestB & 0xffffu), static_cast<int>(bestB >> 16)};
#pragma unroll
    for (int i = 0; i < 4; ++i) k[i] = (k[i] > kmax) ? (kmax + 1) : k[i];
    if (code)  // tile in output order: strip, row, cell
      *reinterpret_cast<uint2 *>(s_code + (tx >> 1) * 64 + r * 8 + 4 * (tx & 1)) =
          make_uint2((4u * k[0]) | ((4u * k[1]) << 16), (4u * k[2]) | ((4u * k[3]) << 16));
    if (x < w) {  // w % 4 == 0: all four cells exist, 16-byte aligned
      const size_t i = static_cast<size_t>(y) * w + x;
      *reinterpret_cast<float4 *>(field + i) = make_float4(__ldg(lut_field + k[0]), __ldg(lut_field + k[1]),
                                                           __ldg(lut_field + k[2]), __ldg(lut_field + k[3]));
      if (distance)
        *reinterpret_cast<float4 *>(distance + i) = make_float4(__ldg(lut_dist + k[0]), __ldg(lut_dist + k[1]),
                                                                __ldg(lut_dist + k[2]), __ldg(lut_dist + k[3]));
    }
  }
  if (code) {
    __syncthreads();
    // 64 strips x 128 bytes; y0 is a multiple of 8, so a strip's eight rows are one line
    const int rows = min(kRowBRows, row_end - y0);
    for (int i = threadIdx.x; i < 64 * 8; i += kRowBThreads) {
      const int strip = i >> 3, part = i & 7;  // part = row inside the strip's line
      const int xs = x0 + 8 * strip;
      if (xs < w && part < rows)
        *reinterpret_cast<uint4 *>(code + static_cast<size_t>(xs >> 3) * stride +
                                   8 * static_cast<size_t>(y0 + part)) =
            *reinterpret_cast<const uint4 *>(s_code + strip * 64 + part * 8);
    }
  }
}

// K4 of the fast path: free cells of rows [first, end) appended to the list.  16
// cells per thread (one 16-byte load), zero bytes counted with bit tricks, the
// tile's indices compacted in shared memory and written out contiguously; tile
// prefixes by decoupled look-back as in scan.cuh.  Slabs chain through *base_in.
constexpr int kFreeTile = 4096;
__device__ __forceinline__ unsigned zero_bytes4(unsigned v) {  // 0x80 in every zero byte
  return ~(((v & 0x7f7f7f7fu) + 0x7f7f7f7fu) | v | 0x7f7f7f7fu);
}
__global__ void __launch_bounds__(256)
free_cells_u8_kernel(const int8_t *__restrict__ state, size_t first, size_t end,
                     uint32_t *__restrict__ out, unsigned long long *__restrict__ status,
                     uint32_t epoch, const unsigned long long *__restrict__ base_in,
                     unsigned long long *__restrict__ total_out) {
  __shared__ uint32_t s_idx[kFreeTile];
  __shared__ uint32_t s_prefix;
  const unsigned tile = blockIdx.x;
  const size_t cell0 = first + static_cast<size_t>(tile) * kFreeTile + threadIdx.x * 16;
  unsigned v[4] = {0x01010101u, 0x01010101u, 0x01010101u, 0x01010101u};
  if (cell0 + 16 <= end) {
    const uint4 q = *reinterpret_cast<const uint4 *>(state + cell0);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  } else if (cell0 < end) {
    for (int b = 0; b < 16 && cell0 + b < end; ++b)
      if (state[cell0 + b] == 0) v[b >> 2] &= ~(0xffu << (8 * (b & 3)));
  }
  unsigned z[4];
  uint32_t c = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    z[k] = zero_bytes4(v[k]);
    c += __popc(z[k]);
  }
  uint32_t total;
  uint32_t pos = block_excl_scan_u32(c, &total);
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    volatile unsigned long long *st = status;
    uint32_t excl = 0;
    if (tile == 0) {
      if (base_in) excl = static_cast<uint32_t>(*base_in);
      if (lane == 0) st[0] = scan_pack(epoch, kScanInclusive, excl + total);
    } else {
      if (lane == 0) st[tile] = scan_pack(epoch, kScanAggregate, total);
      __threadfence();
      int hi = static_cast<int>(tile) - 1;
      while (true) {
        const int t = hi - lane;
        unsigned long long wd = 0;
        bool ready;
        do {
          wd = (t >= 0) ? st[t] : scan_pack(epoch, kScanInclusive, 0u);
          ready = (wd >> 34) == epoch && (wd & (3ull << 32)) != 0;
        } while (!__all_sync(0xffffffffu, ready));
        const unsigned incl_mask = __ballot_sync(0xffffffffu, (wd & kScanInclusive) != 0);
        const int stop = incl_mask ? (__ffs(incl_mask) - 1) : 32;
        uint32_t part = (lane <= stop) ? static_cast<uint32_t>(wd) : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        excl += part;
        if (incl_mask) break;
        hi -= 32;
      }
      if (lane == 0) {
        __threadfence();
        st[tile] = scan_pack(epoch, kScanInclusive, excl + total);
      }
    }
    if (lane == 0) {
      s_prefix = excl;
      if (tile == gridDim.x - 1) *total_out = static_cast<unsigned long long>(excl) + total;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (z[k] & (0x80u << (8 * b))) s_idx[pos++] = static_cast<uint32_t>(cell0 + 4 * k + b);
  __syncthreads();
  uint32_t *dst = out + s_prefix;
  for (uint32_t i = threadIdx.x; i < total; i += 256) dst[i] = s_idx[i];
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

static int env_int(const char *name, int dflt) {
  const char *e = std::getenv(name);
  const int v = e ? std::atoi(e) : 0;
  return v > 0 ? v : dflt;
}

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
  m->free_totals.release();
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
  for (cudaEvent_t e : m->col_events) cudaEventDestroy(e);
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  if (m->col_stream) cudaStreamDestroy(m->col_stream);
  if (m->aux_stream) cudaStreamDestroy(m->aux_stream);
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
  // The occupancy grid crosses PCIe in row slabs on a second stream while the
  // slabs already here are transformed: a slab's column pass needs R rows of the
  // next one, so it waits for that slab's copy; its row pass and its share of the
  // free-cell list follow it.  The last slab is split further so that little work
  // is left once the final bytes have arrived.
  const bool vec4 = width % 4 == 0 && R <= 253;
  const bool fast = width % 4 == 0 && R <= kRow2MaxR;
  const int seg = fast ? kColBSeg : (vec4 ? kColSeg : kSegLen);
  const int slab_pref = env_int("QMCL_MAP_SLAB_ROWS", 512);
  const int tail_pref = env_int("QMCL_MAP_TAIL_ROWS", 256);
  const int min_rows = static_cast<int>(div_up(R + seg, seg)) * seg;  // a slab covers its neighbour's halo
  std::vector<int> bounds{0};
  {
    int rows = std::max(min_rows, static_cast<int>(div_up(std::max(slab_pref, seg), seg)) * seg);
    if (static_cast<int>(div_up(height, rows)) > 30) rows = static_cast<int>(div_up(div_up(height, 30), seg)) * seg;
    int y = 0;
    while (static_cast<int>(height) - y > rows + min_rows) bounds.push_back(y += rows);
    // the rest [y, height): halved down to the tail size
    const int tail_rows = std::max(min_rows, static_cast<int>(div_up(std::max(tail_pref, seg), seg)) * seg);
    int rest = static_cast<int>(height) - y;
    while (fast && rest >= 2 * tail_rows) {
      const int half = static_cast<int>(div_up(rest / 2, seg)) * seg;
      bounds.push_back(y += half);
      rest = static_cast<int>(height) - y;
    }
    bounds.push_back(static_cast<int>(height));
  }
  const int n_slabs = static_cast<int>(bounds.size()) - 1;
  if (!m->copy_stream) QMCL_CUDA(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  while (m->slab_events.size() < static_cast<size_t>(n_slabs) + 1) {
    cudaEvent_t e = nullptr;
    QMCL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    m->slab_events.push_back(e);
  }
  QMCL_TRY(m->h_count.ensure(2));
  QMCL_TRY(m->free_totals.ensure(static_cast<size_t>(n_slabs) + 2));
  // QMCL_MAP_TRACE=1: timestamps of every slab's copy and kernels on stderr (profiling aid)
  const bool trace = env_int("QMCL_MAP_TRACE", 0) != 0;
  std::vector<cudaEvent_t> tr_copy, tr_kern;
  cudaEvent_t tr_start = nullptr;
  if (trace) {
    QMCL_CUDA(cudaEventCreate(&tr_start));
    for (int i = 0; i < n_slabs; ++i) {
      cudaEvent_t a = nullptr, b = nullptr;
      QMCL_CUDA(cudaEventCreate(&a));
      QMCL_CUDA(cudaEventCreate(&b));
      tr_copy.push_back(a);
      tr_kern.push_back(b);
    }
    QMCL_CUDA(cudaEventRecord(tr_start, m->stream));
  }
  // the copies may not overtake work already queued on the map's stream that reads `occ`
  QMCL_CUDA(cudaEventRecord(m->slab_events[n_slabs], m->stream));
  QMCL_CUDA(cudaStreamWaitEvent(m->copy_stream, m->slab_events[n_slabs], 0));
  cudaStream_t cs = m->stream, as = m->stream;  // streams of the column pass and of the free-cell list
  if (fast) {
    if (!m->col_stream) QMCL_CUDA(cudaStreamCreateWithFlags(&m->col_stream, cudaStreamNonBlocking));
    if (!m->aux_stream) QMCL_CUDA(cudaStreamCreateWithFlags(&m->aux_stream, cudaStreamNonBlocking));
    while (m->col_events.size() < static_cast<size_t>(n_slabs) + 1) {
      cudaEvent_t e = nullptr;
      QMCL_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      m->col_events.push_back(e);
    }
    // QMCL_MAP_STREAMS: 2 = everything on the map's stream, 3 = column pass + free-cell
    // list on a second one, 4 = one stream each
    const int n_streams = env_int("QMCL_MAP_STREAMS", 3);
    cs = n_streams >= 3 ? m->col_stream : m->stream;
    as = n_streams >= 4 ? m->aux_stream : cs;
    if (cs != m->stream) QMCL_CUDA(cudaStreamWaitEvent(cs, m->slab_events[n_slabs], 0));
    if (as != cs) QMCL_CUDA(cudaStreamWaitEvent(as, m->slab_events[n_slabs], 0));
  }
  // the tables and the division self-check only need the parameters and the geometry:
  // they run under the first copy
  edt_lut_kernel<<<div_up(kmax + 2, 256), 256, 0, m->stream>>>(lut_dist, lut_field, kmax + 2,
                                                               resolution, max_dist, m->z_hit_pre);
  QMCL_LAUNCHED();
  QMCL_TRY(sensor_prepare_map_launch(m));
  for (int sl = 0; sl < n_slabs; ++sl) {
    const size_t r0 = bounds[sl], r1 = bounds[sl + 1];
    QMCL_CUDA(copy_h2d(m->occ.ptr + r0 * width, occ + r0 * width, (r1 - r0) * width, m->copy_stream));
    QMCL_CUDA(cudaEventRecord(m->slab_events[sl], m->copy_stream));
    if (trace) QMCL_CUDA(cudaEventRecord(tr_copy[sl], m->copy_stream));
  }
  const bool rows2 = R <= kRow2MaxR;
  const size_t smem2 = ((kRow2Cells + 2 * static_cast<size_t>((R + 1) & ~1)) / 2 + 4) * sizeof(unsigned);
  const size_t smem1 = (kRowChunk + 2 * static_cast<size_t>(R)) * sizeof(int);
  const size_t smemB = static_cast<size_t>(kRowBRows) * ((kRowBCells + 2 * ((R + 3) & ~3)) / 2) * sizeof(unsigned) +
                       static_cast<size_t>(kRowBRows) * kRowBCells * sizeof(uint16_t);
  if (!rows2) {
    if (smem1 > 200 * 1024) return QMCL_ERR_INVALID_ARG;
    if (smem1 > 48 * 1024)
      QMCL_CUDA(cudaFuncSetAttribute(edt_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     static_cast<int>(smem1)));
  }
  if (fast) {
    QMCL_TRY(m->free_cells.ensure(n + 2));  // worst case: every cell is free
    QMCL_TRY(m->scan_tmp.ensure_zero(scan_tmp_words(n) + 8 + 2 * static_cast<size_t>(n_slabs)));
  }
  size_t tile_base = 0;  // status words used by the free-cell kernels of earlier slabs
  for (int sl = 0; sl < n_slabs; ++sl) {
    const int r0 = bounds[sl], r1 = bounds[sl + 1];
    if (fast) {
      // three streams: the column pass of a slab starts as soon as the next slab's copy
      // has landed (its halo); the row pass and the slab's share of the free-cell list
      // follow the column pass side by side.  Every kernel here is at most one wave of
      // blocks, so the overlap is what fills the device.
      QMCL_CUDA(cudaStreamWaitEvent(cs, m->slab_events[std::min(sl + 1, n_slabs - 1)], 0));
      uint8_t *g8 = reinterpret_cast<uint8_t *>(m->edt_g.ptr);
      dim3 grid1(div_up(width, 4 * kColBThreads), div_up(r1 - r0, kColBSeg));
      edt_columns_u8_kernel<<<grid1, 2 * kColBThreads, 0, cs>>>(
          m->occ.ptr, m->state.ptr, g8, static_cast<int>(width), R, r0, r1, static_cast<int>(height));
      QMCL_LAUNCHED();
      if (cs != m->stream) {
        QMCL_CUDA(cudaEventRecord(m->col_events[sl], cs));
        QMCL_CUDA(cudaStreamWaitEvent(m->stream, m->col_events[sl], 0));
      }
      dim3 grid2(div_up(width, kRowBCells), div_up(r1 - r0, kRowBRows));
      edt_rows_u8_kernel<<<grid2, kRowBThreads, smemB, m->stream>>>(
          g8, lut_dist, lut_field, with_code ? nullptr : m->distance.ptr, m->field.ptr,
          with_code ? m->code.ptr : nullptr, static_cast<int>(width), static_cast<int>(height), r0, r1, R, kmax);
      QMCL_LAUNCHED();
      const size_t first = static_cast<size_t>(r0) * width, last = static_cast<size_t>(r1) * width;
      const unsigned tiles = div_up(last - first, kFreeTile);
      uint32_t epoch = g_scan_epoch.fetch_add(1, std::memory_order_relaxed) & 0x3fffffffu;
      if (epoch == 0) epoch = g_scan_epoch.fetch_add(1, std::memory_order_relaxed) & 0x3fffffffu;
      if (as != cs) QMCL_CUDA(cudaStreamWaitEvent(as, m->col_events[sl], 0));
      free_cells_u8_kernel<<<tiles, 256, 0, as>>>(
          m->state.ptr, first, last, m->free_cells.ptr,
          reinterpret_cast<unsigned long long *>(m->scan_tmp.ptr) + tile_base, epoch,
          sl ? m->free_totals.ptr + (sl - 1) : nullptr, m->free_totals.ptr + sl);
      QMCL_LAUNCHED();
      tile_base += tiles;
      if (trace) QMCL_CUDA(cudaEventRecord(tr_kern[sl], m->stream));
      continue;
    }
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
  if (fast) {
    if (as != m->stream) {
      QMCL_CUDA(cudaEventRecord(m->col_events[n_slabs], as));
      QMCL_CUDA(cudaStreamWaitEvent(m->stream, m->col_events[n_slabs], 0));
    }
    QMCL_CUDA(copy_d2h(m->h_count.ptr, m->free_totals.ptr + (n_slabs - 1), sizeof(unsigned long long), m->stream));
    m->n_free = 0;
  } else {
    QMCL_TRY(rebuild_free_cells(m));
  }
  m->has_map = true;
  m->has_distance = true;
  m->has_code = with_code;
  m->n_codes = kmax + 2;
  QMCL_TRY(sensor_prepare_map_finish(m));  // synchronises: the free-cell count has arrived with it
  read_free_count(m);
  if (trace && fast) {
    cudaStreamSynchronize(m->copy_stream);
    for (int i = 0; i < n_slabs; ++i) {
      float tc = 0.f, tk = 0.f;
      cudaEventElapsedTime(&tc, tr_start, tr_copy[i]);
      cudaEventElapsedTime(&tk, tr_start, tr_kern[i]);
      fprintf(stderr, "slab %2d rows %5d-%5d  copy done %8.1f us  kernels done %8.1f us\n", i, bounds[i],
              bounds[i + 1], 1e3 * tc, 1e3 * tk);
      cudaEventDestroy(tr_copy[i]);
      cudaEventDestroy(tr_kern[i]);
    }
    cudaEventDestroy(tr_start);
  }
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
