// common.cuh — shared internals of libqmcl_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>

#include "../../include/qmcl.h"

namespace qmcl {

// ------------------------------------------------------------ error plumbing
void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;
extern std::atomic<uint64_t> g_h2d_bytes, g_d2h_bytes;

// Host<->device copies go through these so that the bytes crossing PCIe are
// counted (qmcl_transfer_bytes; bench.py reports them per step).
inline cudaError_t copy_h2d(void *dst, const void *src, size_t bytes, cudaStream_t s) {
  g_h2d_bytes.fetch_add(bytes, std::memory_order_relaxed);
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s);
}
inline cudaError_t copy_d2h(void *dst, const void *src, size_t bytes, cudaStream_t s) {
  g_d2h_bytes.fetch_add(bytes, std::memory_order_relaxed);
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s);
}

#define QMCL_CUDA(expr)                                                        \
  do {                                                                         \
    cudaError_t err__ = (expr);                                                \
    if (err__ != cudaSuccess) {                                                \
      ::qmcl::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,          \
                        cudaGetErrorString(err__));                            \
      return QMCL_ERR_CUDA;                                                    \
    }                                                                          \
  } while (0)

#define QMCL_TRY(expr)                                                         \
  do {                                                                         \
    qmcl_status st__ = (expr);                                                 \
    if (st__ != QMCL_OK) return st__;                                          \
  } while (0)

// Counts the launch and checks the launch error.
#define QMCL_LAUNCHED()                                                        \
  do {                                                                         \
    ::qmcl::g_launches.fetch_add(1, std::memory_order_relaxed);                \
    QMCL_CUDA(cudaGetLastError());                                             \
  } while (0)

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs

// Phase bodies shared by the stand-alone kernels and the fused tail.  Inlined:
// out of line (-DQMCL_PHASE_INLINE=0) the bodies read the job's pointers through
// generic loads from the parameter space instead of the constant bank, which
// lengthens every dependent chain (measured at C2: tail 156 -> 178 us, the
// resolve warp alone 41 -> 67 us).
#ifndef QMCL_PHASE_INLINE
#define QMCL_PHASE_INLINE 1
#endif
#if QMCL_PHASE_INLINE
#define QMCL_PHASE_FN static __device__ __forceinline__
#else
#define QMCL_PHASE_FN static __device__ __noinline__
#endif

inline unsigned div_up(size_t a, size_t b) { return static_cast<unsigned>((a + b - 1) / b); }

// Grow-only device buffer.
template <typename T> struct DevBuf {
  T *ptr = nullptr;
  size_t cap = 0;
  qmcl_status ensure(size_t n) {
    if (n <= cap) return QMCL_OK;
    size_t want = n + n / 8 + 16;
    T *np = nullptr;
    QMCL_CUDA(cudaMalloc(&np, want * sizeof(T)));
    if (ptr) cudaFree(ptr);
    ptr = np;
    cap = want;
    return QMCL_OK;
  }
  // Grow; new storage starts out zeroed (the scan's status words need that).
  qmcl_status ensure_zero(size_t n) {
    if (n <= cap) return QMCL_OK;
    const size_t want = n + n / 8 + 16;
    T *np = nullptr;
    QMCL_CUDA(cudaMalloc(&np, want * sizeof(T)));
    QMCL_CUDA(cudaMemset(np, 0, want * sizeof(T)));
    if (ptr) cudaFree(ptr);
    ptr = np;
    cap = want;
    return QMCL_OK;
  }
  // Grow and keep the first `keep` elements.
  qmcl_status ensure_keep(size_t n, size_t keep, cudaStream_t s) {
    if (n <= cap) return QMCL_OK;
    size_t want = n + n / 8 + 16;
    T *np = nullptr;
    QMCL_CUDA(cudaMalloc(&np, want * sizeof(T)));
    if (ptr && keep) {
      QMCL_CUDA(cudaMemcpyAsync(np, ptr, keep * sizeof(T), cudaMemcpyDeviceToDevice, s));
      QMCL_CUDA(cudaStreamSynchronize(s));
    }
    if (ptr) cudaFree(ptr);
    ptr = np;
    cap = want;
    return QMCL_OK;
  }
  void release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
  }
};

// Grow-only pinned host staging buffer.
template <typename T> struct PinnedBuf {
  T *ptr = nullptr;
  size_t cap = 0;
  qmcl_status ensure(size_t n) {
    if (n <= cap) return QMCL_OK;
    size_t want = n + n / 8 + 16;
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    cap = 0;
    QMCL_CUDA(cudaMallocHost(&ptr, want * sizeof(T)));
    cap = want;
    return QMCL_OK;
  }
  void release() {
    if (ptr) cudaFreeHost(ptr);
    ptr = nullptr;
    cap = 0;
  }
};

// ------------------------------------------------------------------ geometry
// scaled_map.h:94-98 + :43-48 — the float expression and the truncation are
// the reference's; nothing here may be contracted into an FMA or replaced by
// a reciprocal multiply (cell-boundary flips).
struct MapGeom {
  uint32_t w, h;
  float res, ox, oy;
};

// Layout of the uint16 code map (sensor model's palette path): vertical strips
// 8 cells wide, rows contiguous inside a strip, so that a 128-byte line holds an
// 8 x 8 block of cells.  The 32 beam ends a warp gathers at once follow a wall
// contour; row-major storage puts them on ~10 lines, this layout on ~3.5
// (measured on the synthetic scans).
//   offset(x, y) = (x >> 3) * stride + 8 * y + (x & 7),  stride = 8 * round_up(h, 8)
__host__ __device__ __forceinline__ uint32_t code_strip_stride(uint32_t h) {
  return 8u * ((h + 7u) & ~7u);
}
__host__ __device__ __forceinline__ size_t code_map_elems(uint32_t w, uint32_t h) {
  return static_cast<size_t>((w + 7u) >> 3) * code_strip_stride(h);
}
__device__ __forceinline__ uint32_t code_offset(uint32_t x, uint32_t y, uint32_t stride) {
  // (x >> 3) * stride + 8 * y + (x & 7), with x & 7 = x - 8 * (x >> 3)
  return (x >> 3) * (stride - 8u) + ((y << 3) + x);
}

// The 1-byte palette map (sensor model's PAL8 path) uses the same construction
// with strips 16 cells wide: a 128-byte line holds a 16 x 8 block of cells.
//   offset(x, y) = (x >> 4) * stride + 16 * y + (x & 15),  stride = 16 * round_up(h, 8)
__host__ __device__ __forceinline__ uint32_t code8_strip_stride(uint32_t h) {
  return 16u * ((h + 7u) & ~7u);
}
__host__ __device__ __forceinline__ size_t code8_map_elems(uint32_t w, uint32_t h) {
  return static_cast<size_t>((w + 15u) >> 4) * code8_strip_stride(h);
}
__device__ __forceinline__ uint32_t code8_offset(uint32_t x, uint32_t y, uint32_t stride) {
  return (x >> 4) * (stride - 16u) + ((y << 4) + x);
}

__device__ __forceinline__ int cell_index(float pos, float offset, float res) {
  return __float2int_rz(__fdiv_rn(__fsub_rn(pos, offset), res));
}
__device__ __forceinline__ int cell_index_inf(float pos, float res) {
  return __float2int_rz(__fdiv_rn(pos, res));
}
__device__ __forceinline__ bool in_map(int ix, int iy, const MapGeom &g) {
  return static_cast<unsigned>(ix) < g.w && static_cast<unsigned>(iy) < g.h;
}

// ---------------------------------------------------------------- RNG (§RNG)
// Philox4x32-10, key = 64-bit seed, counter = (index_lo, index_hi, step,
// purpose).  Same contract as the oracle, implemented independently.
enum Purpose : uint32_t {
  P_INIT = 1, P_GLOBAL = 2, P_MOTION = 3, P_LV_OFFSET = 4, P_DRAW = 5, P_INJECT = 6
};

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                       uint32_t c3, uint32_t k0, uint32_t k1,
                                                       uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0;
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    c1 = static_cast<uint32_t>(p1);
    c3 = static_cast<uint32_t>(p0);
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__host__ __device__ __forceinline__ void rng_draw(uint64_t seed, uint32_t step, uint32_t purpose,
                                                  uint64_t index, uint32_t out[4]) {
  philox4x32_10(static_cast<uint32_t>(index), static_cast<uint32_t>(index >> 32), step, purpose,
                static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), out);
}

__host__ __device__ __forceinline__ double uniform53(uint32_t lo, uint32_t hi) {
  const uint64_t v = (static_cast<uint64_t>(hi) << 32) | lo;
  return static_cast<double>(v >> 11) * 0x1.0p-53;
}

constexpr double kPi = 3.14159265358979323846;

// Box-Muller in double from two 32-bit words (u1 in (0,1], u2 in [0,1)).
__device__ __forceinline__ void normal_pair(uint32_t a, uint32_t b, double *z0, double *z1) {
  const double u1 = (static_cast<double>(a) + 1.0) * 0x1.0p-32;
  const double u2 = static_cast<double>(b) * 0x1.0p-32;
  const double rad = sqrt(-2.0 * log(u1));
  const double ang = (2 * kPi) * u2;
  double s, c;
  sincos(ang, &s, &c);
  *z0 = rad * c;
  *z1 = rad * s;
}

// utils.h:28-35 (evaluated in double, also for float input).
__host__ __device__ __forceinline__ double normalise_angle(double a) {
  return a - 2 * kPi * floor((a + kPi) / (2 * kPi));
}
__host__ __device__ __forceinline__ float normalise_angle_f(float a) {
  return static_cast<float>(a - 2 * kPi * floor((a + kPi) / (2 * kPi)));
}
// utils.h:38-61
__host__ __device__ __forceinline__ double angle_delta(double a, double b) {
  a = normalise_angle(a);
  b = normalise_angle(b);
  const double d1 = a - b;
  double d2 = 2 * kPi - fabs(d1);
  if (0 < d1) d2 = -d2;
  return (fabs(d1) < fabs(d2)) ? d1 : d2;
}

// map_base.cpp:58-64 under the RNG contract: free cell from words 0-1, heading
// from word 2; the pose is the cell *corner* (scaled_map.h:101-105).
__device__ __forceinline__ void random_free_pose(const uint32_t *__restrict__ free_idx,
                                                 unsigned long long n_free, const MapGeom &g,
                                                 const uint32_t r[4], float *x, float *y,
                                                 float *t) {
  const unsigned long long idx =
      __umul64hi((static_cast<unsigned long long>(r[1]) << 32) | r[0], n_free);
  const unsigned cell = free_idx[idx];
  const float cx = static_cast<float>(cell % g.w);
  const float cy = static_cast<float>(cell / g.w);
  *x = __fadd_rn(__fmul_rn(cx, g.res), g.ox);
  *y = __fadd_rn(__fmul_rn(cy, g.res), g.oy);
  const double u = static_cast<double>(r[2]) * 0x1.0p-32;
  float th = static_cast<float>(-kPi + (2 * kPi) * u);
  if (th > 3.1415925f) th = 3.1415925f;
  *t = th;
}

// ------------------------------------------------------------ warp utilities
__device__ __forceinline__ double warp_sum_xor(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace qmcl
