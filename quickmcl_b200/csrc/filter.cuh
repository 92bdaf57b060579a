// filter.cuh — internal launch helpers shared by the filter translation units.
#pragma once

#include "state.cuh"

namespace qmcl {

qmcl_status launch_aos_to_soa(cudaStream_t s, const qmcl_particle *dev_in, size_t n, float *x,
                              float *y, float *t, double *w);
qmcl_status launch_soa_to_aos(cudaStream_t s, const float *x, const float *y, const float *t,
                              const double *w, size_t n, qmcl_particle *dev_out);
qmcl_status launch_export_markers(cudaStream_t s, const float *x, const float *y, const float *t,
                                  const double *w, size_t n_all, size_t n_out,
                                  unsigned long long *d_max, float *d_out);
// Deferred read-backs (state.cuh): settle_after_sync folds in whatever is
// pending when the caller has just synchronised the filter's stream; settle
// synchronises first if anything is pending.  lazy(f): deferral is in effect.
void settle_after_sync(qmcl_filter *f);
qmcl_status settle(qmcl_filter *f);
inline bool lazy(const qmcl_filter *f) { return f->lazy_sync && f->n_shards <= 1; }
// Sharded filters defer the same three read-backs (sensor total, bin count, cluster
// count): every host-synchronising exchange (shard_gather) folds them in, and the
// sharded adaptive / KLD resamplers settle before they read w_slow / w_fast.
inline bool defer_counts(const qmcl_filter *f) { return f->lazy_sync; }
// Fused tail (tail.cu).  tail_enabled: the filter defers K8 into the tail;
// tail_eligible: this call (TAIL_* mode) can run as one fused launch;
// tail_complete: fold the result in (the stream has been synchronised).
bool tail_enabled(const qmcl_filter *f);
// Sets up to this size run the fused tail; larger ones the multi-launch path, whose
// kernels each run at their own occupancy (a cooperative grid of one 256-thread
// block per SM is too thin for millions of dependent look-ups: measured at 1 M
// particles, adaptive: fused tail 1.28 ms, multi-launch 0.65 ms).
constexpr size_t kTailMaxParticles = size_t(1) << 18;
bool tail_eligible(const qmcl_filter *f, int mode);
qmcl_status launch_tail(qmcl_filter *f, int mode);
qmcl_status tail_complete(qmcl_filter *f);
// K8 + the total's read-back, when they were deferred (sensor_update_common).
qmcl_status flush_sensor(qmcl_filter *f);
// K6 of a handle_odometry call left to the next sensor update (motion.cu).
qmcl_status flush_motion(qmcl_filter *f);
// particle_filter.cpp:106-120 on the host: w_slow / w_fast from the sensor total.
void apply_sensor_total(qmcl_filter *f, double total, double n_total);
// Grow-only byte scratch of the filter (contents are not preserved).
qmcl_status filter_scratch(qmcl_filter *f, size_t bytes, void **out);

}  // namespace qmcl
