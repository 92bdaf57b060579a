// cluster.cuh — bin grid / clustering entry points shared with the resamplers.
#pragma once

#include "filter.cuh"

#include <algorithm>

namespace qmcl {

constexpr size_t kMaxTableCells = size_t(1) << 28;  // dense bin table: 1 GiB of int32

// SpacePartitioning::reset + add_point for every pose in x/y/t[0..n).
// reuse_grid: keep the dense grid (bounding box) of the previous call — valid
// when the new poses are a subset of the previous ones; the sorted bin list
// does not depend on the grid's extent.
// grid_only: compute the grid (and size the table) but do not mark or list bins.
// known_bbox: {x0, y0, t0, x1, y1, t1} bin bounds the caller already holds on
// the host (any box containing the poses' bins), saving the bounding-box pass
// and its read-back.  known_bins >= 0: the number of occupied bins, when the
// caller has counted them, saving that read-back.  Both single-GPU only.
qmcl_status build_bins(qmcl_filter *f, const float *x, const float *y, const float *t, size_t n,
                       bool reuse_grid = false, bool grid_only = false,
                       const int *known_bbox = nullptr, long long known_bins = -1);
// The bin bounding box of x/y/t[0..n) into f->bbox (device, 6 ints; min > max
// when n == 0).  No synchronisation.
qmcl_status launch_bin_bbox(qmcl_filter *f, const float *x, const float *y, const float *t,
                            size_t n);
// SpacePartitioning::assign_clusters on the bins built last.
qmcl_status label_clusters(qmcl_filter *f);

}  // namespace qmcl
