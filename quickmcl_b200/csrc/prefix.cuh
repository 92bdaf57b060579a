// prefix.cuh — K9: sequentially rounded running sum (see prefix.cu).
#pragma once

#include "filter.cuh"

namespace qmcl {

// out[i] = fl(...fl(fl(carry_in + w[0]) + w[1])... + w[i]), bit-identical to
// the sequential loop; the last value is also left in f->prefix_carry.ptr[0].
qmcl_status sequential_prefix(qmcl_filter *f, const double *w, size_t n, double *out,
                              double carry_in = 0.0);

}  // namespace qmcl
