// cluster.cuh — bin grid / clustering entry points shared with the resamplers.
#pragma once

#include "filter.cuh"

#include <algorithm>

namespace qmcl {

// SpacePartitioning::reset + add_point for every pose in x/y/t[0..n).
// reuse_grid: keep the dense grid (bounding box) of the previous call — valid
// when the new poses are a subset of the previous ones; the sorted bin list
// does not depend on the grid's extent.
// grid_only: compute the grid (and size the table) but do not mark or list bins.
qmcl_status build_bins(qmcl_filter *f, const float *x, const float *y, const float *t, size_t n,
                       bool reuse_grid = false, bool grid_only = false);
// SpacePartitioning::assign_clusters on the bins built last.
qmcl_status label_clusters(qmcl_filter *f);

}  // namespace qmcl
