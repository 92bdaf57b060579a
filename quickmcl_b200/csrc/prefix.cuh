// prefix.cuh — the reference-faithful (sequentially rounded) running sum.
#pragma once

#include "filter.cuh"

namespace qmcl {

// out[i] = fl(fl(...fl(w[0] + w[1]) ...) + w[i]): the exact sequence of
// rounded double additions of the reference's running sums
// (particle_filter.cpp:225-236, particle.cpp:66-82).  w[i] >= 0.
qmcl_status sequential_prefix(qmcl_filter *f, const double *w, size_t n, double *out);

}  // namespace qmcl
