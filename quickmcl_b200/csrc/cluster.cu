// cluster.cu — K13 (bins + connected components) and K14 (cluster moments and
// pose estimate).
//
// Replaces SpacePartitioning (src/quickmcl/space_partitioning.cpp:39-128),
// ParticleFilter::cluster / compute_cluster_statistics / get_estimated_pose
// (src/quickmcl/particle_filter.cpp:163-212, :366-391) and
// ParticleCloudStatistics (src/quickmcl/statistics.cpp:23-79).
//
// The reference keeps occupied bins in an unordered_map.  Here the bin grid is
// a DENSE int32 table over the bounding box of the particles' bin coordinates
// (x-major, then y, then theta — i.e. lexicographic key order): marking is an
// atomicAdd, the occupied-bin list is a stream compaction of the table (so it
// comes out sorted), neighbour lookups are direct loads, and connected
// components are a lock-free union-find whose roots are the smallest bin index
// of each component, which is also the canonical label the oracle reports.

#include "cluster.cuh"
#include "cluster_dev.cuh"

#include "scan.cuh"

#include <cstdlib>
#include "shard.cuh"

namespace qmcl {



__global__ void __launch_bounds__(256)
bin_bbox_kernel(const float *__restrict__ x, const float *__restrict__ y,
                const float *__restrict__ t, size_t n, BinParams bp, int *__restrict__ bbox) {
  bin_bbox_block(x, y, t, n, bp, bbox, blockIdx.x, gridDim.x);
}

__global__ void __launch_bounds__(256)
bin_mark_kernel(const float *__restrict__ x, const float *__restrict__ y,
                const float *__restrict__ t, size_t n, BinParams bp, BinGrid g,
                int32_t *__restrict__ table) {
  bin_mark_one(x, y, t, n, bp, g, table, static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x);
}

__global__ void __launch_bounds__(256)
bin_mark_occupancy_kernel(const float *__restrict__ x, const float *__restrict__ y,
                          const float *__restrict__ t, size_t n, BinParams bp, BinGrid g,
                          unsigned *__restrict__ occ_words) {
  bin_mark_occupancy(x, y, t, n, bp, g, occ_words, static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x);
}

__global__ void __launch_bounds__(256)
bin_count_kernel(const float *__restrict__ x, const float *__restrict__ y,
                 const float *__restrict__ t, size_t n, BinParams bp, BinGrid g,
                 const int32_t *__restrict__ table, int32_t *__restrict__ bin_count) {
  bin_count_one(x, y, t, n, bp, g, table, bin_count, static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x);
}

__global__ void __launch_bounds__(256)
cc_hook_kernel(const int32_t *__restrict__ table, const int32_t *__restrict__ bin_cell,
               size_t n_bins, const unsigned long long *__restrict__ n_bins_dev, BinGrid g,
               BinParams bp, int32_t *__restrict__ parent, bool linked_runs) {
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  // n_bins is a launch bound while the count is still on its way to the host
  if (n_bins_dev) n_bins = static_cast<size_t>(*n_bins_dev);
  if (tid >= n_bins) return;
  cc_hook_bin(table, bin_cell, g, bp, parent, static_cast<int>(tid), linked_runs);
}

__global__ void __launch_bounds__(256)
cc_flatten_kernel(int32_t *__restrict__ parent, size_t n_bins,
                  const unsigned long long *__restrict__ n_bins_dev) {
  const size_t b = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n_bins_dev) n_bins = static_cast<size_t>(*n_bins_dev);
  if (b >= n_bins) return;
  parent[b] = uf_find_readonly(parent, static_cast<int>(b));
}

__global__ void __launch_bounds__(256)
wmax_kernel(const double *__restrict__ w, size_t n, EstimateScalars *__restrict__ es) {
  unsigned long long m = 0;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const double v = w[i];
    if (v > 0) m = max(m, static_cast<unsigned long long>(__double_as_longlong(v)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m) atomicMax(&es->wmax_bits, m);
}

__global__ void __launch_bounds__(256)
cluster_weight_kernel(const float *__restrict__ x, const float *__restrict__ y,
                      const float *__restrict__ t, const double *__restrict__ w, size_t n,
                      BinParams bp, BinGrid g, const int32_t *__restrict__ table,
                      const int32_t *__restrict__ label,
                      const int32_t *__restrict__ cluster_of_root,
                      int32_t *__restrict__ cid_out, unsigned long long *__restrict__ cluster_w,
                      EstimateScalars *__restrict__ es, size_t n_total) {
  const int shift = fixed_point_shift(es->wmax_bits, n_total);
  cluster_weight_block(x, y, t, w, n, bp, g, table, label, cluster_of_root, cid_out, cluster_w, es,
                       shift, blockIdx.x);
}

// First maximum of the cluster weights (ties -> smallest canonical id).
__global__ void __launch_bounds__(1024)
best_cluster_kernel(const unsigned long long *__restrict__ cluster_w, size_t n_clusters,
                    const unsigned long long *__restrict__ n_clusters_dev,
                    EstimateScalars *__restrict__ es) {
  // the count may still be on its way to the host (label_clusters, deferred)
  if (n_clusters_dev) n_clusters = static_cast<size_t>(*n_clusters_dev);
  best_cluster_block(cluster_w, n_clusters, es);
}

// The same over many blocks, for cluster lists of a global cloud (tens of
// thousands of clusters and more; one block needs 24 us for 47 k of them): block
// maxima into scratch, the last block to finish folds them.  Integer comparisons
// with the index as the tie-break: the result does not depend on the order.
__global__ void __launch_bounds__(256)
best_cluster_multi_kernel(const unsigned long long *__restrict__ cluster_w, size_t n_clusters,
                          const unsigned long long *__restrict__ n_clusters_dev,
                          EstimateScalars *__restrict__ es, unsigned long long *__restrict__ part_w,
                          long long *__restrict__ part_i) {
  __shared__ unsigned long long s_w[8];
  __shared__ long long s_i[8];
  __shared__ bool s_last;
  if (n_clusters_dev) n_clusters = static_cast<size_t>(*n_clusters_dev);
  auto better = [](unsigned long long w1, long long i1, unsigned long long w2, long long i2) {
    if (i1 < 0) return false;
    if (i2 < 0) return true;
    return w1 > w2 || (w1 == w2 && i1 < i2);
  };
  auto warp_best = [&](unsigned long long &bw, long long &bi) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ow = __shfl_xor_sync(0xffffffffu, bw, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ow, oi, bw, bi)) {
        bw = ow;
        bi = oi;
      }
    }
  };
  unsigned long long bw = 0;
  long long bi = -1;
  for (size_t c = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; c < n_clusters;
       c += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const unsigned long long v = cluster_w[c];
    if (bi < 0 || v > bw) {
      bw = v;
      bi = static_cast<long long>(c);
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  warp_best(bw, bi);
  if (lane == 0) {
    s_w[wid] = bw;
    s_i[wid] = bi;
  }
  __syncthreads();
  if (wid == 0) {
    bw = lane < 8 ? s_w[lane] : 0ull;
    bi = lane < 8 ? s_i[lane] : -1ll;
    warp_best(bw, bi);
    if (lane == 0) {
      part_w[blockIdx.x] = bw;
      part_i[blockIdx.x] = bi;
      __threadfence();
      s_last = atomicAdd(&es->pad, 1u) == gridDim.x - 1;
    }
  }
  __syncthreads();
  if (!s_last || wid != 0) return;
  __threadfence();
  bw = 0;
  bi = -1;
  for (unsigned b = lane; b < gridDim.x; b += 32) {
    const unsigned long long v = __ldcg(part_w + b);
    const long long i = __ldcg(part_i + b);
    if (better(v, i, bw, bi)) {
      bw = v;
      bi = i;
    }
  }
  warp_best(bw, bi);
  if (lane == 0) {
    es->best = static_cast<int>(bi);
    es->best_w = bw;
    es->pad = 0;
  }
}

// sharded estimate: the "a bin has no cluster" flag as a double behind the moment sums
__global__ void pack_invalid_kernel(const EstimateScalars *__restrict__ es, double *__restrict__ out) {
  *out = es->invalid ? 1.0 : 0.0;
}

// statistics.cpp:23-41 for the best cluster and for all particles: block
// partials, then the last block adds them in block order (moments_final).
__global__ void __launch_bounds__(256)
moments_kernel(const float *__restrict__ x, const float *__restrict__ y,
               const float *__restrict__ t, const double *__restrict__ w,
               const int32_t *__restrict__ cid, size_t n, double *__restrict__ partials,
               EstimateScalars *__restrict__ es, double *__restrict__ sums_out,
               double *__restrict__ out, int finalise) {
  __shared__ bool s_last;
  moments_block(x, y, t, w, cid, n, partials, es->best, blockIdx.x, gridDim.x);
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&es->blocks_done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  moments_final(partials, gridDim.x, sums_out, out, finalise);
  if (threadIdx.x == 0) es->blocks_done = 0;
}

__global__ void decode_bins_kernel(const int32_t *__restrict__ bin_cell,
                                   const int32_t *__restrict__ bin_count,
                                   const int32_t *__restrict__ label, size_t n_bins, BinGrid g,
                                   int32_t *__restrict__ keys4, int32_t *__restrict__ labels) {
  const size_t b = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= n_bins) return;
  const int cell = bin_cell[b];
  keys4[4 * b + 0] = cell / (g.nt * g.ny) + g.x0;
  keys4[4 * b + 1] = (cell / g.nt) % g.ny + g.y0;
  keys4[4 * b + 2] = cell % g.nt + g.t0;
  keys4[4 * b + 3] = bin_count[b];
  labels[b] = label ? label[b] : -1;
}

static BinParams bin_params(const qmcl_filter *f) {
  BinParams bp;
  bp.res_xy = f->params.res_xy;
  bp.res_t = f->params.res_theta;
  bp.theta_lo = f->theta_lo;
  bp.theta_hi = f->theta_hi;
  return bp;
}

qmcl_status launch_bin_bbox(qmcl_filter *f, const float *x, const float *y, const float *t,
                            size_t n) {
  cudaStream_t s = f->map->stream;
  QMCL_TRY(f->bbox.ensure(8));
  bbox_init_kernel<<<1, 32, 0, s>>>(f->bbox.ptr);
  QMCL_LAUNCHED();
  if (n) {
    const unsigned bb_blocks = static_cast<unsigned>(std::min<size_t>(div_up(n, 256), 4 * kSMs));
    bin_bbox_kernel<<<bb_blocks, 256, 0, s>>>(x, y, t, n, bin_params(f), f->bbox.ptr);
    QMCL_LAUNCHED();
  }
  return QMCL_OK;
}

// SpacePartitioning::reset + add_point for x/y/t[0..n): bbox, dense table,
// sorted occupied-bin list.  Leaves parent[b] = b.
qmcl_status build_bins(qmcl_filter *f, const float *x, const float *y, const float *t, size_t n,
                       bool reuse_grid, bool grid_only, const int *known_bbox, long long known_bins) {
  cudaStream_t s = f->map->stream;
  const BinGrid prev_grid = f->grid;
  const bool have_prev = reuse_grid && f->bins_built && prev_grid.nx > 0;
  f->n_bins = 0;
  f->bins_pending = false;
  f->n_clusters = 0;
  f->clusters_pending = false;  // an older count still in flight is void
  f->bins_built = false;
  f->clusters_valid = false;
  f->est_cached = false;
  if (n == 0 && !sharded(f)) {
    f->grid = BinGrid{0, 0, 0, 0, 0, 0};
    f->bins_built = true;
    return QMCL_OK;
  }
  const BinParams bp = bin_params(f);
  StageScope sc(f, QMCL_STAGE_BINS);
  int hb[6];
  if (have_prev) {
    hb[0] = prev_grid.x0;
    hb[1] = prev_grid.y0;
    hb[2] = prev_grid.t0;
    hb[3] = prev_grid.x0 + prev_grid.nx - 1;
    hb[4] = prev_grid.y0 + prev_grid.ny - 1;
    hb[5] = prev_grid.t0 + prev_grid.nt - 1;
  } else if (known_bbox) {  // (a sharded caller passes the union over all ranks)
    for (int d = 0; d < 6; ++d) hb[d] = known_bbox[d];
  } else {
    QMCL_TRY(launch_bin_bbox(f, x, y, t, n));
    if (sharded(f)) {
      // Exchange: the bin grid is the bounding box over all shards.
      std::vector<int> all(6 * f->n_shards);
      QMCL_TRY(shard_gather(f, f->bbox.ptr, true, 6 * sizeof(int), all.data()));
      for (int d = 0; d < 6; ++d) hb[d] = all[d];
      for (int r = 1; r < f->n_shards; ++r)
        for (int d = 0; d < 3; ++d) {
          hb[d] = std::min(hb[d], all[6 * r + d]);
          hb[3 + d] = std::max(hb[3 + d], all[6 * r + 3 + d]);
        }
      if (hb[0] > hb[3]) {  // no particle anywhere
        f->grid = BinGrid{0, 0, 0, 0, 0, 0};
        f->bins_built = true;
        return QMCL_OK;
      }
    } else {
      int *h = reinterpret_cast<int *>(f->h_scalars.ptr + HS_BBOX);
      QMCL_CUDA(copy_d2h(h, f->bbox.ptr, sizeof(hb), s));
      QMCL_CUDA(cudaStreamSynchronize(s));
      settle_after_sync(f);
      for (int d = 0; d < 6; ++d) hb[d] = h[d];
    }
  }
  BinGrid g;
  g.x0 = hb[0];
  g.y0 = hb[1];
  g.t0 = hb[2];
  const long long nx = static_cast<long long>(hb[3]) - hb[0] + 1;
  const long long ny = static_cast<long long>(hb[4]) - hb[1] + 1;
  const long long nt = static_cast<long long>(hb[5]) - hb[2] + 1;
  if (nx <= 0 || ny <= 0 || nt <= 0 ||
      static_cast<double>(nx) * ny * nt > static_cast<double>(kMaxTableCells)) {
    set_error("bin grid %lld x %lld x %lld exceeds the dense-table capacity", nx, ny, nt);
    return QMCL_ERR_CAPACITY;
  }
  g.nx = static_cast<int>(nx);
  g.ny = static_cast<int>(ny);
  g.nt = static_cast<int>(nt);
  const size_t cells = static_cast<size_t>(nx) * ny * nt;
  f->grid = g;
  QMCL_TRY(f->table.ensure(cells));
  if (grid_only) {  // the caller only needs the dense grid (KLD's first-draw table)
    f->bins_built = true;
    return QMCL_OK;
  }
  size_t n_all = n;
  const size_t max_bins_local = std::min(sharded(f) ? f->global_n : n, cells);
  QMCL_TRY(f->bin_cell.ensure(max_bins_local));
  QMCL_TRY(f->bin_count.ensure(max_bins_local));
  QMCL_TRY(f->bin_label.ensure(max_bins_local));
  QMCL_TRY(f->bin_cluster.ensure(max_bins_local));
  QMCL_TRY(f->scan_tmp.ensure_zero(scan_tmp_words(cells) + 4));
  unsigned long long *d_total = &f->scalars.ptr->n_bins;  // stable: the labelling reads it
  // QMCL_LINK_THETA=0 (A/B): the forest starts without the theta runs linked
  static const bool link_theta = !(std::getenv("QMCL_LINK_THETA") && std::getenv("QMCL_LINK_THETA")[0] == '0');
  const int link_nt = link_theta ? g.nt : 0;
  f->bins_linked = link_nt > 0;
  if (sharded(f)) {
    // Exchange: bin OCCUPANCY, one 0/1 byte per cell - a quarter of the count table's
    // bytes.  The bytes are summed four at a time as int32 words: with at most 255
    // ranks no byte can carry into its neighbour, and an int32 sum is the reduction
    // NCCL runs inside the NVSwitch (a uint8 maximum takes the ring).  The
    // occupied-bin list and the labels are then computed identically on every rank;
    // the per-bin counts are made on demand (qmcl_filter_copy_bins).
    const size_t words = (cells + 3) / 4;
    QMCL_TRY(f->occ8.ensure(4 * words));
    // QMCL_BINS_SPLIT=1 (diagnostic): the marking is booked under the otherwise unused
    // "kld" stage and the exchange under "tail" in the stage breakdown of a sharded run
    static const bool split = std::getenv("QMCL_BINS_SPLIT") != nullptr;
    {
      StageScope sub(f, QMCL_STAGE_KLD, split ? 2 : 99);
      QMCL_CUDA(cudaMemsetAsync(f->occ8.ptr, 0, 4 * words, s));
      if (n) {
        bin_mark_occupancy_kernel<<<div_up(n, 256), 256, 0, s>>>(x, y, t, n, bp, g,
                                                                 reinterpret_cast<unsigned *>(f->occ8.ptr));
        QMCL_LAUNCHED();
      }
    }
    {
      StageScope sub(f, QMCL_STAGE_TAIL, split ? 2 : 99);
      QMCL_TRY(shard_allreduce(f, f->occ8.ptr, words, QMCL_DT_I32, QMCL_OP_SUM));
    }
    n_all = f->global_n;
    QMCL_TRY(device_scan(s, OccupiedByte{f->occ8.ptr},
                         EmitBinOccupancy{f->occ8.ptr, f->table.ptr, f->bin_cell.ptr, f->bin_label.ptr, link_nt}, cells,
                         f->scan_tmp.ptr, d_total));
    f->bin_counts_valid = false;
  } else {
    QMCL_CUDA(cudaMemsetAsync(f->table.ptr, 0, cells * sizeof(int32_t), s));
    if (n) {
      bin_mark_kernel<<<div_up(n, 256), 256, 0, s>>>(x, y, t, n, bp, g, f->table.ptr);
      QMCL_LAUNCHED();
    }
    QMCL_TRY(device_scan(s, OccupiedFlag{f->table.ptr},
                         EmitBin{f->table.ptr, f->bin_cell.ptr, f->bin_count.ptr, f->bin_label.ptr, link_nt},
                         cells, f->scan_tmp.ptr, d_total));
    f->bin_counts_valid = true;
  }
  const size_t max_bins = std::min(n_all, cells);
  if (known_bins >= 0 && !sharded(f)) {
    // the caller counted the bins already (KLD: new-bin flags up to the stop)
    f->n_bins = static_cast<size_t>(known_bins);
  } else {
    // The host needs the count only to size the labelling launches, for which
    // max_bins is a bound: in lazy mode the copy is left in flight.
    QMCL_CUDA(copy_d2h(f->h_scalars.ptr + HS_NBINS, d_total, sizeof(unsigned long long), s));
    f->bins_pending = true;
    f->n_bins_bound = max_bins;
    if (!defer_counts(f)) QMCL_TRY(settle(f));
  }
  f->bins_built = true;
  return QMCL_OK;
}

// SpacePartitioning::assign_clusters on the current bin list.
qmcl_status label_clusters(qmcl_filter *f) {
  cudaStream_t s = f->map->stream;
  f->n_clusters = 0;
  f->clusters_pending = false;  // an older count still in flight is void
  // With the bin count still in flight the launches are sized by its bound and
  // the kernels read the count on the device.
  const size_t nb = f->bins_pending ? f->n_bins_bound : f->n_bins;
  const unsigned long long *nb_dev = f->bins_pending ? &f->scalars.ptr->n_bins : nullptr;
  if (nb == 0) {
    f->clusters_valid = true;
    return QMCL_OK;
  }
  const BinParams bp = bin_params(f);
  // the linked start of the forest is there once: after this pass it is flattened
  const bool linked = f->bins_linked;
  f->bins_linked = false;
  StageScope sc(f, QMCL_STAGE_LABEL);
  cc_hook_kernel<<<div_up(nb, 256), 256, 0, s>>>(f->table.ptr, f->bin_cell.ptr, nb, nb_dev, f->grid,
                                                 bp, f->bin_label.ptr, linked);
  QMCL_LAUNCHED();
  cc_flatten_kernel<<<div_up(nb, 256), 256, 0, s>>>(f->bin_label.ptr, nb, nb_dev);
  QMCL_LAUNCHED();
  QMCL_TRY(f->scan_tmp.ensure_zero(scan_tmp_words(nb) + 4));
  // the count goes to a stable device word (K14 reads it there) and to its own
  // pinned slot; the host needs it only to size K14's buffers, for which the
  // bin count is a bound, so the copy is left in flight
  unsigned long long *d_total = &f->scalars.ptr->n_clusters;
  QMCL_TRY(device_scan(s, RootFlag{f->bin_label.ptr, nb_dev}, EmitRootId{f->bin_cluster.ptr}, nb,
                       f->scan_tmp.ptr, d_total));
  QMCL_CUDA(copy_d2h(f->h_scalars.ptr + HS_NCLUSTERS, d_total, sizeof(unsigned long long), s));
  f->clusters_pending = true;
  if (!defer_counts(f)) QMCL_TRY(settle(f));
  f->clusters_valid = true;
  return QMCL_OK;
}

}  // namespace qmcl

using namespace qmcl;

extern "C" {

// particle_filter.cpp:163-175
qmcl_status qmcl_filter_cluster(qmcl_filter *f) {
  if (!f) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(use_device(f->map));
  if (f->tail_pending) QMCL_TRY(settle(f));
  // bins, labels and the estimate in one fused launch when it applies (tail.cu)
  if (tail_eligible(f, 3)) return launch_tail(f, 3);
  QMCL_TRY(flush_motion(f));
  const ParticleSet &ps = f->cur();
  QMCL_TRY(build_bins(f, ps.x.ptr, ps.y.ptr, ps.t.ptr, ps.n));
  return label_clusters(f);
}

qmcl_status qmcl_filter_get_estimate(qmcl_filter *f, double pose[3], double cov[9], int *valid) {
  if (!f || !pose || !cov || !valid) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(use_device(f->map));
  cudaStream_t s = f->map->stream;
  *valid = 0;
  if (f->tail_pending) QMCL_TRY(settle(f));
  if (f->est_cached && f->clusters_valid) {  // the fused tail computed it already
    if (f->est_invalid) {
      set_error("get_estimate: a particle's bin has no cluster (reference abort()s here)");
      return QMCL_ERR_INVALID_CLUSTER;
    }
    if (f->est[0] != 0.0) {
      *valid = 1;
      pose[0] = f->est[1];
      pose[1] = f->est[2];
      pose[2] = f->est[3];
      for (int k = 0; k < 9; ++k) cov[k] = f->est[4 + k];
    }
    return QMCL_OK;
  }
  QMCL_TRY(flush_motion(f));
  QMCL_TRY(flush_sensor(f));
  // A sharded filter exchanges the cluster weights below and needs the cluster count on
  // the host for that: it waits for the counts here (a plain stream wait) instead of
  // inside an exchange of its own.
  if (sharded(f)) QMCL_TRY(settle(f));
  const ParticleSet &ps = f->cur();
  // every rank of a sharded filter must take the same branch here
  const size_t n_any = sharded(f) ? f->global_n : ps.n;
  // A cluster count still on its way to the host (label_clusters) is at least 1
  // and at most the bin count, which is all the host needs here; K14 reads the
  // exact value on the device.
  const bool nc_on_device = f->clusters_pending;
  const size_t nc_cap =
      nc_on_device ? (f->bins_pending ? f->n_bins_bound : f->n_bins) : f->n_clusters;
  if (n_any == 0 && nc_cap == 0) return QMCL_OK;  // no clusters: "ALL clusters suck"
  if (n_any > 0 && (!f->bins_built || nc_cap == 0)) {
    set_error("get_estimate: particles have no cluster (call cluster() or resample first)");
    return QMCL_ERR_INVALID_CLUSTER;
  }
  const BinParams bp = bin_params(f);
  StageScope sc(f, QMCL_STAGE_MOMENTS);
  const unsigned blocks =
      static_cast<unsigned>(std::min<size_t>(std::max<size_t>(div_up(ps.n, 256), 1), kMomentBlocks));
  // layout of cluster_acc (doubles): [n_clusters] fixed-point cluster weights |
  // [blocks * 2 * kMom] block partials | [2 * kMom] sums | [16] out | EstimateScalars
  const size_t o_part = nc_cap;
  const size_t o_sums = o_part + static_cast<size_t>(blocks) * 2 * kMom;
  const size_t o_out = o_sums + 2 * kMom;
  const size_t o_es = o_out + 16;
  const size_t total = o_es + (sizeof(EstimateScalars) + 7) / 8;
  QMCL_TRY(f->cluster_acc.ensure(total));
  QMCL_TRY(f->particle_cid.ensure(ps.n + 1));
  double *base = f->cluster_acc.ptr;
  unsigned long long *cluster_w = reinterpret_cast<unsigned long long *>(base);
  double *d_out = base + o_out;
  EstimateScalars *es = reinterpret_cast<EstimateScalars *>(base + o_es);
  QMCL_CUDA(cudaMemsetAsync(base, 0, nc_cap * sizeof(double), s));
  QMCL_CUDA(cudaMemsetAsync(base + o_sums, 0, (2 * kMom + 16) * sizeof(double) + sizeof(EstimateScalars), s));
  if (ps.n) {
    wmax_kernel<<<std::min<unsigned>(blocks, 2 * kSMs), 256, 0, s>>>(ps.w.ptr, ps.n, es);
    QMCL_LAUNCHED();
  }
  // Exchange: the fixed-point scale needs the largest weight of all shards.  The bits of
  // a non-negative double order like integers: an in-place integer maximum, consumed by
  // the next kernel - no host wait.
  if (sharded(f)) QMCL_TRY(shard_allreduce(f, &es->wmax_bits, 1, QMCL_DT_I64, QMCL_OP_MAX));
  if (ps.n) {
    cluster_weight_kernel<<<div_up(ps.n, 256), 256, 0, s>>>(
        ps.x.ptr, ps.y.ptr, ps.t.ptr, ps.w.ptr, ps.n, bp, f->grid, f->table.ptr, f->bin_label.ptr,
        f->bin_cluster.ptr, f->particle_cid.ptr, cluster_w, es,
        sharded(f) ? f->global_n : ps.n);
    QMCL_LAUNCHED();
  }
  // Exchange: cluster weights are integers, so their sum over shards is exact
  // and order-independent.
  if (sharded(f)) QMCL_TRY(shard_allreduce(f, cluster_w, f->n_clusters, QMCL_DT_I64, QMCL_OP_SUM));
  if (nc_cap >= 8192 && blocks >= 32) {
    // scratch: the moments' block partials (written after this kernel)
    const unsigned nb = std::min<unsigned>(kSMs, div_up(nc_cap, 1024));
    unsigned long long *part_w = reinterpret_cast<unsigned long long *>(base + o_part);
    best_cluster_multi_kernel<<<nb, 256, 0, s>>>(cluster_w, nc_cap,
                                                 nc_on_device ? &f->scalars.ptr->n_clusters : nullptr, es,
                                                 part_w, reinterpret_cast<long long *>(part_w + kSMs));
  } else {
    best_cluster_kernel<<<1, 1024, 0, s>>>(cluster_w, nc_cap,
                                           nc_on_device ? &f->scalars.ptr->n_clusters : nullptr, es);
  }
  QMCL_LAUNCHED();
  moments_kernel<<<blocks, 256, 0, s>>>(ps.x.ptr, ps.y.ptr, ps.t.ptr, ps.w.ptr, f->particle_cid.ptr,
                                        ps.n, base + o_part, es, base + o_sums, d_out,
                                        sharded(f) ? 0 : 1);
  QMCL_LAUNCHED();
  double h[16];
  int flag = 0;
  if (sharded(f)) {
    // Exchange: moments of the best cluster and of all particles, summed in rank order.
    // (the flag joins the sums on the device: one exchange, one host wait)
    struct Part { double sums[2 * kMom]; double invalid; };
    pack_invalid_kernel<<<1, 1, 0, s>>>(es, base + o_sums + 2 * kMom);
    QMCL_LAUNCHED();
    std::vector<Part> parts(f->n_shards);
    QMCL_TRY(shard_gather(f, base + o_sums, true, sizeof(Part), parts.data()));
    double tot[2 * kMom] = {0};
    flag = 0;
    for (const Part &pt : parts) {
      for (int k = 0; k < 2 * kMom; ++k) tot[k] += pt.sums[k];
      if (pt.invalid != 0.0) flag = 1;
    }
    finalise_estimate(tot + kMom, tot, h);
  } else {
    // one read back: out[0..13) + the invalid flag
    double *hp = f->h_scalars.ptr + HS_EST;
    QMCL_CUDA(copy_d2h(hp, d_out, 13 * sizeof(double), s));
    QMCL_CUDA(copy_d2h(hp + 13, &es->invalid, sizeof(int), s));
    QMCL_CUDA(cudaStreamSynchronize(s));
    settle_after_sync(f);
    std::memcpy(h, hp, 13 * sizeof(double));
    std::memcpy(&flag, hp + 13, sizeof(int));
  }
  if (flag) {
    set_error("get_estimate: a particle's bin has no cluster (reference abort()s here)");
    return QMCL_ERR_INVALID_CLUSTER;
  }
  if (h[0] != 0.0) {
    *valid = 1;
    pose[0] = h[1];
    pose[1] = h[2];
    pose[2] = h[3];
    for (int k = 0; k < 9; ++k) cov[k] = h[4 + k];
  }
  return QMCL_OK;
}

qmcl_status qmcl_filter_num_bins(const qmcl_filter *f, size_t *out) {
  if (!f || !out) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(settle(const_cast<qmcl_filter *>(f)));
  *out = f->n_bins;
  return QMCL_OK;
}

qmcl_status qmcl_filter_num_clusters(const qmcl_filter *f, size_t *out) {
  if (!f || !out) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(settle(const_cast<qmcl_filter *>(f)));
  *out = f->n_clusters;
  return QMCL_OK;
}

qmcl_status qmcl_filter_copy_bins(const qmcl_filter *fc, int32_t *keys4, int32_t *labels,
                                  size_t cap, size_t *n_out) {
  if (!fc) return QMCL_ERR_INVALID_ARG;
  qmcl_filter *f = const_cast<qmcl_filter *>(fc);
  QMCL_TRY(settle(f));
  if (n_out) *n_out = f->n_bins;
  const size_t n = f->n_bins < cap ? f->n_bins : cap;
  if (!n || !keys4 || !labels) return QMCL_OK;
  QMCL_TRY(use_device(f->map));
  cudaStream_t s = f->map->stream;
  if (!f->bin_counts_valid) {
    // A sharded filter exchanged occupancy only (build_bins): the counts are made now,
    // from every rank's particles - on a sharded filter this call is collective.
    const ParticleSet &ps = f->cur();
    QMCL_CUDA(cudaMemsetAsync(f->bin_count.ptr, 0, f->n_bins * sizeof(int32_t), s));
    if (ps.n) {
      bin_count_kernel<<<div_up(ps.n, 256), 256, 0, s>>>(ps.x.ptr, ps.y.ptr, ps.t.ptr, ps.n, bin_params(f),
                                                         f->grid, f->table.ptr, f->bin_count.ptr);
      QMCL_LAUNCHED();
    }
    if (sharded(f)) QMCL_TRY(shard_allreduce(f, f->bin_count.ptr, f->n_bins, QMCL_DT_I32, QMCL_OP_SUM));
    f->bin_counts_valid = true;
  }
  void *stage = nullptr;
  QMCL_TRY(filter_scratch(f, n * 5 * sizeof(int32_t), &stage));
  int32_t *d_keys = static_cast<int32_t *>(stage);
  int32_t *d_labels = d_keys + 4 * n;
  decode_bins_kernel<<<div_up(n, 256), 256, 0, s>>>(
      f->bin_cell.ptr, f->bin_count.ptr, f->clusters_valid ? f->bin_label.ptr : nullptr, n, f->grid,
      d_keys, d_labels);
  QMCL_LAUNCHED();
  QMCL_CUDA(copy_d2h(keys4, d_keys, n * 4 * sizeof(int32_t), s));
  QMCL_CUDA(copy_d2h(labels, d_labels, n * sizeof(int32_t), s));
  QMCL_CUDA(cudaStreamSynchronize(s));
  return QMCL_OK;
}

}  // extern "C"
