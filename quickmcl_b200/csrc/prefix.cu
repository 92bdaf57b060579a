// prefix.cu — K9: the CDF as the reference computes it.
//
// A reassociated parallel sum rounds differently from the reference's
// sequential loop, and a resampling index flips whenever a draw falls between
// the two results (expected ~100 flips at 16 M particles).  So this
// file produces the *sequentially rounded* running sum.

#include "prefix.cuh"

namespace qmcl {

// Version 1: one warp, coalesced loads of 32 weights, lane-uniform dependent
// additions.  Exactly the reference's order by construction.
__global__ void __launch_bounds__(32)
sequential_prefix_kernel(const double *__restrict__ w, size_t n, double *__restrict__ out) {
  const int lane = threadIdx.x;
  double c = 0.0;
  for (size_t base = 0; base < n; base += 32) {
    const size_t i = base + lane;
    const double v = (i < n) ? w[i] : 0.0;
    double mine = 0.0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      c = __dadd_rn(c, __shfl_sync(0xffffffffu, v, k));
      if (lane == k) mine = c;
    }
    if (i < n) out[i] = mine;
  }
}

qmcl_status sequential_prefix(qmcl_filter *f, const double *w, size_t n, double *out) {
  if (!n) return QMCL_OK;
  sequential_prefix_kernel<<<1, 32, 0, f->map->stream>>>(w, n, out);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

}  // namespace qmcl
