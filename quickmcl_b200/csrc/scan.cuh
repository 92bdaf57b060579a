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

// ---- single-pass variant (decoupled look-back) ------------------------------
// One launch instead of three.  Every tile publishes its aggregate in a 64-bit
// status word {epoch:30, state:2, value:32}; a tile's exclusive prefix is the sum
// of the aggregates of its predecessors up to the first one that already knows
// its inclusive prefix.  The epoch (unique per call) makes words left by earlier
// scans read as "not ready", so the status array is never cleared; it only has
// to start out zeroed (DevBuf::ensure_zero).  Tiles are taken in blockIdx order,
// which is the order the hardware dispatches them in, so a predecessor is
// always running or finished.
constexpr unsigned long long kScanAggregate = 1ull << 32, kScanInclusive = 2ull << 32;
__device__ __forceinline__ unsigned long long scan_pack(uint32_t epoch, unsigned long long state,
                                                        uint32_t value) {
  return (static_cast<unsigned long long>(epoch) << 34) | state | value;
}

template <typename F, typename G>
__global__ void __launch_bounds__(kScanBlock)
scan_onepass_kernel(F f, G g, size_t n, unsigned long long *__restrict__ status, uint32_t epoch,
                    unsigned long long *__restrict__ out_total) {
  __shared__ uint32_t s_prefix;
  const unsigned tile = blockIdx.x;
  const size_t base = (static_cast<size_t>(tile) * kScanBlock + threadIdx.x) * kScanItems;
  uint32_t v[kScanItems];
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (base + i < n) ? f(base + i) : 0u;
    c += v[i];
  }
  uint32_t total;
  const uint32_t local = block_excl_scan_u32(c, &total);
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    volatile unsigned long long *st = status;
    uint32_t excl = 0;
    if (tile == 0) {
      if (lane == 0) st[0] = scan_pack(epoch, kScanInclusive, total);
    } else {
      if (lane == 0) st[tile] = scan_pack(epoch, kScanAggregate, total);
      __threadfence();
      // look back over windows of 32 predecessors
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
        // lanes up to and including the first "inclusive" one contribute
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
      if (tile == gridDim.x - 1 && out_total) *out_total = static_cast<unsigned long long>(excl) + total;
    }
  }
  __syncthreads();
  uint32_t pos = s_prefix + local;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (base + i < n) {
      g(base + i, pos, v[i]);
      pos += v[i];
    }
}

extern std::atomic<uint32_t> g_scan_epoch;

// uint32 words a scan over n items needs in `tmp` (status words + room for the
// 64-bit results the callers keep behind them), and where those results start.
inline size_t scan_tmp_words(size_t n) { return 2 * static_cast<size_t>(scan_blocks(n)) + 16; }
inline unsigned long long *scan_total_ptr(uint32_t *tmp, size_t n) {
  return reinterpret_cast<unsigned long long *>(tmp) + scan_blocks(n);
}

// Host driver.  tmp: scan_tmp_words(n) uint32, zero-initialised when allocated
// (DevBuf::ensure_zero).  Leaves the grand total in *dev_total (device).
template <typename F, typename G>
qmcl_status device_scan(cudaStream_t s, F f, G g, size_t n, uint32_t *tmp,
                        unsigned long long *dev_total) {
  if (n == 0) {
    QMCL_CUDA(cudaMemsetAsync(dev_total, 0, sizeof(unsigned long long), s));
    return QMCL_OK;
  }
  uint32_t epoch = g_scan_epoch.fetch_add(1, std::memory_order_relaxed) & 0x3fffffffu;
  if (epoch == 0) epoch = g_scan_epoch.fetch_add(1, std::memory_order_relaxed) & 0x3fffffffu;
  scan_onepass_kernel<<<scan_blocks(n), kScanBlock, 0, s>>>(
      f, g, n, reinterpret_cast<unsigned long long *>(tmp), epoch, dev_total);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

}  // namespace qmcl
