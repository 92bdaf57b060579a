// prefix.cuh — K9: sequentially rounded running sum (see prefix.cu).
#pragma once

#include "filter.cuh"

namespace qmcl {

// out[i] = fl(...fl(fl(carry_in + w[0]) + w[1])... + w[i]), bit-identical to
// the sequential loop; the last value is also left in f->prefix_carry.ptr[0].
qmcl_status sequential_prefix(qmcl_filter *f, const double *w, size_t n, double *out,
                              double carry_in = 0.0);

// The same in four steps, for shards that chain their carries:
//   prefix_local_total : chunk sums; approximate shard total -> prefix_carry[1]
//   prefix_summaries   : binade guesses from the approximate carry-in + summaries
//   prefix_resolve     : exact carry through the shard; carry-out -> prefix_carry[0]
//   prefix_apply       : expand into out[]
qmcl_status prefix_local_total(qmcl_filter *f, const double *w, size_t n);
qmcl_status prefix_summaries(qmcl_filter *f, const double *w, size_t n, double approx_carry_in);
qmcl_status prefix_summaries_dev(qmcl_filter *f, const double *w, size_t n, const double *terms,
                                 int n_terms, int stride);
qmcl_status prefix_resolve(qmcl_filter *f, const double *w, size_t n, double *out,
                           double carry_in);
qmcl_status prefix_apply(qmcl_filter *f, const double *w, size_t n, double *out);
// Shard-level shortcut for the carry chain: the closed-form summary of the whole
// shard - {binade, a} or, with one predicted binade crossing, {binade, a, wx, b} -
// as kShardSummaryWords 64-bit words in device memory (binade -1 = no closed form,
// -2 = empty shard, -3 = the exact carry-out of a shard that has been resolved
// already, prefix_exact_summary), and its replay on an exact carry on the host.
constexpr int kShardSummaryWords = 7;
qmcl_status prefix_shard_summary(qmcl_filter *f, size_t n, unsigned long long *dev_out);
qmcl_status prefix_exact_summary(qmcl_filter *f, unsigned long long *dev_out);
bool prefix_apply_summary(double carry_in, const unsigned long long summary[kShardSummaryWords],
                          double *carry_out);

namespace pfx {
struct PrefixBuffers;
}
qmcl_status plan_prefix_buffers(qmcl_filter *f, size_t n, pfx::PrefixBuffers *pb,
                                unsigned long long *nc, unsigned long long *nt);

}  // namespace qmcl
