// resample_dev.cuh — device helpers of K10-K12 shared by the stand-alone
// kernels (resample.cu) and the fused tail (tail.cu).  Plain pointers: the tail
// calls them on data written earlier in the same launch.
#pragma once

#include "state.cuh"

namespace qmcl {

// ParticlesRunningSum::lookup_index (particle.h:99-107) over the records
// (key_t = running sum before positive particle t): the record with the
// largest key <= r; among records sharing a key the first inserted wins
// (std::map::insert keeps the existing entry, particle.cpp:77).
// lo/hi: a window [lo, hi] known to contain t* (the whole list: 0, P - 1).
__device__ __forceinline__ size_t running_sum_lookup_in(const double *incl, size_t P, double r, size_t lo,
                                                        size_t hi) {
  // t* = number of t in [1, P) with key_t = incl[t-1] <= r
  while (lo < hi) {
    const size_t mid = (lo + hi) >> 1;
    if (incl[mid] <= r) lo = mid + 1;
    else hi = mid;
  }
  const size_t tstar = lo;
  if (tstar == 0) return 0;
  const double K = incl[tstar - 1];
  // first t with key_t == K, i.e. 1 + first j with incl[j] >= K.  Keys repeat
  // only where a weight was too small to move the sum: one look at the
  // neighbour (same cache line) settles every other case without a second search.
  if (tstar == 1 || incl[tstar - 2] < K) return tstar;
  lo = 0;
  hi = tstar - 1;
  while (lo < hi) {
    const size_t mid = (lo + hi) >> 1;
    if (incl[mid] < K) lo = mid + 1;
    else hi = mid;
  }
  return lo + 1;
}
__device__ __forceinline__ size_t running_sum_lookup(const double *incl, size_t P, double r) {
  return running_sum_lookup_in(incl, P, r, 0, P - 1);
}
// samples of the record list kept in shared memory by the searching kernels
constexpr size_t kLookupSamples = 1024;
// The same with the first levels of the search served from shared memory:
// samp[j] = incl[j * stride] for j < n_samp, j * stride < P - 1 (every stride-th
// key).  Narrows the window to one stride, then searches incl itself: the same
// t* (incl is non-decreasing), with log2(stride) dependent global loads instead
// of log2(P).
__device__ __forceinline__ size_t running_sum_lookup_sampled(const double *incl, size_t P, double r,
                                                             const double *samp, size_t n_samp,
                                                             size_t stride) {
  size_t a = 0, b = n_samp;  // J = number of samples <= r
  while (a < b) {
    const size_t mid = (a + b) >> 1;
    if (samp[mid] <= r) a = mid + 1;
    else b = mid;
  }
  if (a == 0) return running_sum_lookup_in(incl, P, r, 0, 0);  // incl[0] > r (or no samples): t* = 0
  const size_t lo = (a - 1) * stride + 1;
  size_t hi = a * stride;
  if (hi > P - 1) hi = P - 1;
  return running_sum_lookup_in(incl, P, r, lo, hi);
}

struct KldParams {
  float res_xy, res_t;
  double eps2, z;
  int32_t nmin, nmax;
};

// space_partitioning.cpp:52-74, the same double expressions in the same order.
__device__ __forceinline__ unsigned long long kld_limit(unsigned long long k, const KldParams &kp) {
  if (k < 2) return static_cast<unsigned long long>(kp.nmax);
  const double common = 2.0 / static_cast<double>(9 * (k - 1));
  const double to_be_cubed = 1 - common - sqrt(common) * kp.z;
  const double result =
      (static_cast<double>(k - 1) / kp.eps2) * to_be_cubed * to_be_cubed * to_be_cubed;
  const double c = ceil(result);
  int32_t fin;
  if (c >= 2147483647.0) fin = 2147483647;
  else if (c <= -2147483648.0) fin = static_cast<int32_t>(-2147483647 - 1);
  else fin = static_cast<int32_t>(c);
  if (fin < kp.nmin) return static_cast<unsigned long long>(kp.nmin);
  if (fin > kp.nmax) return static_cast<unsigned long long>(kp.nmax);
  return static_cast<unsigned long long>(fin);
}

__device__ __forceinline__ long long kld_cell(float x, float y, float t, const KldParams &kp,
                                              const BinGrid &g) {
  const long long ix = static_cast<long long>(cell_index_inf(x, kp.res_xy)) - g.x0;
  const long long iy = static_cast<long long>(cell_index_inf(y, kp.res_xy)) - g.y0;
  const long long it = static_cast<long long>(cell_index_inf(t, kp.res_t)) - g.t0;
  return (ix * g.ny + iy) * g.nt + it;
}

// particle.cpp:41-56: total = sum of weights.  The sum is a fixed tree over
// virtual blocks (deterministic); it differs from the sequential sum by
// rounding only.  Block `vblock` of `nvblocks` -> partials[vblock].
template <bool SQUARE>
__device__ __forceinline__ void weight_sum_block(const double *w, size_t n, double *partials,
                                                 unsigned vblock, unsigned nvblocks) {
  __shared__ double s_warp[8];
  double t = 0.0;
  for (size_t i = static_cast<size_t>(vblock) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(nvblocks) * blockDim.x) {
    const double v = w[i];
    t = __dadd_rn(t, SQUARE ? __dmul_rn(v, v) : v);
  }
  t = warp_sum_xor(t);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) s_warp[wid] = t;
  __syncthreads();
  if (threadIdx.x == 0) {
    double b = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) b = __dadd_rn(b, s_warp[i]);
    partials[vblock] = b;
  }
  __syncthreads();
}
// The block partials reduced by one block of 256 threads; every thread gets the total.
__device__ __forceinline__ double weight_sum_final(const double *partials, unsigned nvblocks) {
  __shared__ double s_warp[8];
  __shared__ double s_tot;
  double a = 0.0;
  for (unsigned i = threadIdx.x; i < nvblocks; i += blockDim.x)
    a = __dadd_rn(a, __ldcg(partials + i));
  a = warp_sum_xor(a);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) s_warp[wid] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) tot = __dadd_rn(tot, s_warp[i]);
    s_tot = tot;
  }
  __syncthreads();
  const double r = s_tot;
  __syncthreads();
  return r;
}
// Launch geometry of the weight sum (a function of n only).
__host__ __device__ inline unsigned weight_sum_blocks(size_t n) {
  size_t b = (n + 1023) / 1024;
  if (b < 1) b = 1;
  if (b > 4 * 148) b = 4 * 148;
  return static_cast<unsigned>(b);
}

}  // namespace qmcl
