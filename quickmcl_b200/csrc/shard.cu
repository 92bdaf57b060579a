// shard.cu — qmcl_comm plumbing and the pure host arithmetic of the sharded
// path (block slices, low-variance owner ranges).

#include "shard.cuh"
#include "filter.cuh"

namespace qmcl {

qmcl_status shard_gather(qmcl_filter *f, const void *src, bool src_on_device, size_t bytes,
                         void *host_out) {
  cudaStream_t s = f->map->stream;
  const size_t G = static_cast<size_t>(f->n_shards);
  QMCL_TRY(f->xchg.ensure(G * bytes + 16));
  uint8_t *mine = f->xchg.ptr + static_cast<size_t>(f->shard_rank) * bytes;
  if (src_on_device) {
    QMCL_CUDA(cudaMemcpyAsync(mine, src, bytes, cudaMemcpyDeviceToDevice, s));
  } else {
    QMCL_CUDA(copy_h2d(mine, src, bytes, s));
  }
  if (f->comm.allgather(f->comm.ctx, f->xchg.ptr, bytes, s) != 0) {
    set_error("qmcl_comm.allgather failed");
    return QMCL_ERR_COMM;
  }
  f->n_collectives++;
  QMCL_CUDA(copy_d2h(host_out, f->xchg.ptr, G * bytes, s));
  QMCL_CUDA(cudaStreamSynchronize(s));
  settle_after_sync(f);  // read-backs left in flight (sensor total, bin and cluster counts) have arrived too
  return QMCL_OK;
}

// The same exchange with the result left on the device (rank order, in f->xchg): the
// consumer is the next kernel on the stream, the host does not wait.  The buffer is
// reused by the next exchange, which stream order puts behind that consumer.
qmcl_status shard_gather_device(qmcl_filter *f, const void *dev_src, size_t bytes, void **dev_out) {
  cudaStream_t s = f->map->stream;
  const size_t G = static_cast<size_t>(f->n_shards);
  QMCL_TRY(f->xchg.ensure(G * bytes + 16));
  uint8_t *mine = f->xchg.ptr + static_cast<size_t>(f->shard_rank) * bytes;
  QMCL_CUDA(cudaMemcpyAsync(mine, dev_src, bytes, cudaMemcpyDeviceToDevice, s));
  if (f->comm.allgather(f->comm.ctx, f->xchg.ptr, bytes, s) != 0) {
    set_error("qmcl_comm.allgather failed");
    return QMCL_ERR_COMM;
  }
  f->n_collectives++;
  *dev_out = f->xchg.ptr;
  return QMCL_OK;
}

// out = v[0] + v[1] + ... in rank order: the same bits on every rank (and the bits
// ordered_sum gives on the host).
__global__ void sum_in_rank_order_kernel(const double *__restrict__ v, int n, double *__restrict__ out) {
  double t = 0.0;
  for (int i = 0; i < n; ++i) t = __dadd_rn(t, v[i]);
  *out = t;
}
qmcl_status shard_sum_f64_device(qmcl_filter *f, double *dev_scalar) {
  void *g = nullptr;
  QMCL_TRY(shard_gather_device(f, dev_scalar, sizeof(double), &g));
  sum_in_rank_order_kernel<<<1, 1, 0, f->map->stream>>>(static_cast<const double *>(g), f->n_shards, dev_scalar);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

qmcl_status shard_gather_f64(qmcl_filter *f, const double *dev_src, std::vector<double> *out) {
  out->assign(static_cast<size_t>(f->n_shards), 0.0);
  return shard_gather(f, dev_src, true, sizeof(double), out->data());
}

double ordered_sum(const std::vector<double> &v) {
  double t = 0.0;
  for (double x : v) t += x;
  return t;
}

qmcl_status shard_allreduce(qmcl_filter *f, void *dev_buf, size_t count, int dtype, int op) {
  if (!count) return QMCL_OK;
  if (f->comm.allreduce(f->comm.ctx, dev_buf, count, dtype, op, f->map->stream) != 0) {
    set_error("qmcl_comm.allreduce failed");
    return QMCL_ERR_COMM;
  }
  f->n_collectives++;
  return QMCL_OK;
}

void shard_slice_of(size_t n, int rank, int size, size_t *first, size_t *count) {
  const size_t per = (n + size - 1) / size;
  size_t a = per * static_cast<size_t>(rank);
  if (a > n) a = n;
  size_t b = a + per;
  if (b > n) b = n;
  *first = a;
  *count = b - a;
}

}  // namespace qmcl

using namespace qmcl;

extern "C" {

qmcl_status qmcl_filter_set_comm(qmcl_filter *f, const qmcl_comm *comm) {
  if (!f) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(settle(f));  // deferred read-backs exist only on single-GPU filters
  if (!comm || comm->size <= 1) {
    f->comm = qmcl_comm{0, 1, nullptr, nullptr, nullptr, nullptr};
    f->shard_rank = 0;
    f->n_shards = 1;
    return QMCL_OK;
  }
  if (comm->rank < 0 || comm->rank >= comm->size || !comm->allgather || !comm->allreduce ||
      !comm->alltoallv)
    return QMCL_ERR_INVALID_ARG;
  f->comm = *comm;
  f->shard_rank = comm->rank;
  f->n_shards = comm->size;
  return QMCL_OK;
}

qmcl_status qmcl_filter_set_rebalance(qmcl_filter *f, int enabled) {
  if (!f) return QMCL_ERR_INVALID_ARG;
  f->rebalance = enabled != 0;
  return QMCL_OK;
}

qmcl_status qmcl_filter_set_rebalance_threshold(qmcl_filter *f, double relative_imbalance) {
  if (!f || !(relative_imbalance >= 0.0)) return QMCL_ERR_INVALID_ARG;
  f->rebalance_threshold = relative_imbalance;
  return QMCL_OK;
}

qmcl_status qmcl_filter_shard_range(const qmcl_filter *f, uint64_t *first, uint64_t *count,
                                    uint64_t *global_count) {
  if (!f) return QMCL_ERR_INVALID_ARG;
  if (first) *first = f->global_first;
  if (count) *count = f->cur().n;
  if (global_count) *global_count = f->global_n;
  return QMCL_OK;
}

qmcl_status qmcl_filter_shard_stats(const qmcl_filter *f, uint64_t out[3]) {
  if (!f || !out) return QMCL_ERR_INVALID_ARG;
  out[0] = f->n_collectives;
  out[1] = f->n_chain_rounds;
  out[2] = f->n_rebalance_skipped;
  return QMCL_OK;
}

void qmcl_shard_slice(uint64_t n, int rank, int size, uint64_t *first, uint64_t *count) {
  size_t a = 0, c = 0;
  if (size >= 1 && rank >= 0 && rank < size) shard_slice_of(n, rank, size, &a, &c);
  if (first) *first = a;
  if (count) *count = c;
}

// particle_filter.cpp:228-235: slot m uses U_m = r + m * minv (a rounded
// product, then a rounded sum) and takes the first particle whose inclusive
// running sum is >= U_m.  A shard whose running sum spans (c_lo, c_hi] therefore
// owns the slots with c_lo < U_m <= c_hi; U_m is non-decreasing in m.  The
// first shard also takes everything below, the last everything above (the
// clamped overrun).
void qmcl_lv_owner_range(uint64_t M, double r, double minv, double c_lo, double c_hi,
                         int is_first, int is_last, uint64_t *m_lo, uint64_t *m_hi) {
  auto first_above = [&](double c) {
    uint64_t lo = 0, hi = M;
    while (lo < hi) {
      const uint64_t mid = lo + (hi - lo) / 2;
      volatile double prod = static_cast<double>(mid) * minv;  // two roundings, never fused
      const double U = r + prod;
      if (U > c) hi = mid;
      else lo = mid + 1;
    }
    return lo;
  };
  const uint64_t a = is_first ? 0 : first_above(c_lo);
  const uint64_t b = is_last ? M : first_above(c_hi);
  if (m_lo) *m_lo = a;
  if (m_hi) *m_hi = b < a ? a : b;
}

}  // extern "C"
