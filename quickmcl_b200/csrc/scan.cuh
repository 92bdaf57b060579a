// scan.cuh — hand-written device-wide exclusive scan of uint32 (reduce /
// single-block scan of block sums / apply), used by the compactions and by
// the KLD bucket counting.  Templates so that each translation unit gets its
// own instantiation.
#pragma once

#include "common.cuh"

namespace qmcl {

constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;  // contiguous items per thread
constexpr size_t kScanTile = static_cast<size_t>(kScanBlock) * kScanItems;

inline unsigned scan_blocks(size_t n) { return div_up(n, kScanTile); }

// Exclusive scan inside a 256-thread block; *total = block total.
__device__ __forceinline__ uint32_t block_excl_scan_u32(uint32_t v, uint32_t *total) {
  __shared__ uint32_t warp_tot[kScanBlock / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) warp_tot[wid] = incl;
  __syncthreads();
  uint32_t offset = 0, tot = 0;
#pragma unroll
  for (int i = 0; i < kScanBlock / 32; ++i) {
    const uint32_t t = warp_tot[i];
    if (i < wid) offset += t;
    tot += t;
  }
  __syncthreads();
  *total = tot;
  return offset + incl - v;
}

// Phase A: per-block sums of f(i).  F: uint32_t operator()(size_t i) const.
template <typename F>
__global__ void __launch_bounds__(kScanBlock)
scan_reduce_kernel(F f, size_t n, uint32_t *__restrict__ block_sums) {
  const size_t base = (static_cast<size_t>(blockIdx.x) * kScanBlock + threadIdx.x) * kScanItems;
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) c += f(base + i);
  uint32_t total;
  block_excl_scan_u32(c, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// Phase B: exclusive scan of the block sums by ONE block; grand total -> *out_total.
template <int DUMMY = 0>
__global__ void __launch_bounds__(1024)
scan_block_sums_kernel(uint32_t *__restrict__ sums, size_t n, unsigned long long *out_total) {
  __shared__ unsigned long long warp_excl[32];
  __shared__ unsigned long long chunk_total;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned long long carry = 0;
  for (size_t base = 0; base < n; base += 1024) {
    const size_t i = base + threadIdx.x;
    const unsigned long long v = (i < n) ? sums[i] : 0;
    unsigned long long incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) warp_excl[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const unsigned long long t = warp_excl[lane];
      unsigned long long ti = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long up = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += up;
      }
      warp_excl[lane] = ti - t;
      if (lane == 31) chunk_total = ti;
    }
    __syncthreads();
    if (i < n) sums[i] = static_cast<uint32_t>(carry + warp_excl[wid] + incl - v);
    carry += chunk_total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && out_total) *out_total = carry;
}

// Phase C: g(i, exclusive_prefix_of_f_at_i, f(i)) for every i.
template <typename F, typename G>
__global__ void __launch_bounds__(kScanBlock)
scan_apply_kernel(F f, G g, size_t n, const uint32_t *__restrict__ block_offsets) {
  const size_t base = (static_cast<size_t>(blockIdx.x) * kScanBlock + threadIdx.x) * kScanItems;
  uint32_t v[kScanItems];
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? f(base + i) : 0u;
    c += v[i];
  }
  uint32_t total;
  uint32_t pos = block_offsets[blockIdx.x] + block_excl_scan_u32(c, &total);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) {
      g(base + i, pos, v[i]);
      pos += v[i];
    }
}

// Host driver.  tmp must hold scan_blocks(n) + 2 uint32 (8-byte aligned tail
// for the total).  Leaves the grand total in *dev_total (device).
template <typename F, typename G>
qmcl_status device_scan(cudaStream_t s, F f, G g, size_t n, uint32_t *tmp,
                        unsigned long long *dev_total) {
  if (n == 0) {
    QMCL_CUDA(cudaMemsetAsync(dev_total, 0, sizeof(unsigned long long), s));
    return QMCL_OK;
  }
  const unsigned blocks = scan_blocks(n);
  scan_reduce_kernel<<<blocks, kScanBlock, 0, s>>>(f, n, tmp);
  QMCL_LAUNCHED();
  scan_block_sums_kernel<0><<<1, 1024, 0, s>>>(tmp, blocks, dev_total);
  QMCL_LAUNCHED();
  scan_apply_kernel<<<blocks, kScanBlock, 0, s>>>(f, g, n, tmp);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

}  // namespace qmcl
