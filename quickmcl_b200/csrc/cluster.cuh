// cluster.cuh — bin grid / clustering entry points shared with the resamplers.
#pragma once

#include "filter.cuh"

#include <algorithm>

namespace qmcl {

// SpacePartitioning::reset + add_point for every pose in x/y/t[0..n).
qmcl_status build_bins(qmcl_filter *f, const float *x, const float *y, const float *t, size_t n);
// SpacePartitioning::assign_clusters on the bins built last.
qmcl_status label_clusters(qmcl_filter *f);

}  // namespace qmcl
