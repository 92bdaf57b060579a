// shard.cuh — helpers of the particle-sharded path.
//
// A sharded filter is one qmcl_filter per GPU holding a contiguous block of
// the global particle index range.  The exchanges go through the host-supplied
// qmcl_comm table (include/qmcl.h); these helpers wrap the small ones.
#pragma once

#include "state.cuh"

#include <vector>

namespace qmcl {

inline bool sharded(const qmcl_filter *f) { return f->n_shards > 1; }

// Gathers `bytes` from every rank (this rank's contribution is read from
// `src`, a device pointer when src_on_device, else a host pointer) into
// host_out[size * bytes], in rank order.  Synchronises the stream.
qmcl_status shard_gather(qmcl_filter *f, const void *src, bool src_on_device, size_t bytes,
                         void *host_out);
// Device-side variants: the gathered values stay on the device (*dev_out, valid until
// the next exchange is enqueued), resp. *dev_scalar is replaced by the sum of every
// rank's value in rank order.  No host wait.
qmcl_status shard_gather_device(qmcl_filter *f, const void *dev_src, size_t bytes, void **dev_out);
qmcl_status shard_sum_f64_device(qmcl_filter *f, double *dev_scalar);
qmcl_status shard_gather_f64(qmcl_filter *f, const double *dev_src, std::vector<double> *out);
// Sum in rank order (the fixed order makes every rank compute the same bits).
double ordered_sum(const std::vector<double> &v);
// In-place allreduce of a device buffer.
qmcl_status shard_allreduce(qmcl_filter *f, void *dev_buf, size_t count, int dtype, int op);

void shard_slice_of(size_t n, int rank, int size, size_t *first, size_t *count);

}  // namespace qmcl
