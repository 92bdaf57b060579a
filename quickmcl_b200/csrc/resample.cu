// resample.cu — K9-K12: CDF, low-variance / adaptive / KLD resampling, gather.
//
// Replaces ParticleFilter::resample_and_cluster, resample_low_variance,
// resample_adaptive, resample_kld (src/quickmcl/particle_filter.cpp:138-161,
// :214-348), ParticlesRunningSum (src/quickmcl/particle.cpp:66-82,
// include/quickmcl/particle.h:99-107), SpacePartitioning::limit
// (src/quickmcl/space_partitioning.cpp:52-74) and normalise_particle_weights
// (src/quickmcl/particle.cpp:41-56).
//
// Bit-exactness: the reference's CDF is a *sequential* double running sum and
// every resampling decision compares a draw against it, so the CDF here is the
// same sequence of correctly rounded additions (see prefix.cu), not a
// reassociated parallel sum.  Given that array, every step below is exact
// comparison logic and the resulting indices / KLD bin counts are bit-exact.

#include "cluster.cuh"
#include "prefix.cuh"
#include "resample_dev.cuh"
#include "scan.cuh"
#include "shard.cuh"

#include <cmath>

namespace qmcl {

// ---------------------------------------------------------------- low variance
// particle_filter.cpp:225-238: U = r + m * minv (two rounded double ops), walk
// to the first i with c_i >= U.  Deviation: i is clamped to
// M-1 where the reference's .at(i) would throw.
// Sharded use: this shard holds the running sums cdf[0..n_cdf) of its own block
// and produces the output slots m_first + [0, count) that fall into it.
__global__ void __launch_bounds__(256)
lv_select_kernel(const double *__restrict__ cdf, size_t n_cdf, size_t m_first, size_t count,
                 double r, double minv, int64_t *__restrict__ src) {
  const size_t k = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const double U = __dadd_rn(r, __dmul_rn(static_cast<double>(m_first + k), minv));
  // first i with !(U > c_i)
  size_t lo = 0, hi = n_cdf;
  while (lo < hi) {
    const size_t mid = (lo + hi) >> 1;
    if (U > cdf[mid]) lo = mid + 1;
    else hi = mid;
  }
  if (lo > n_cdf - 1) lo = n_cdf - 1;
  src[k] = static_cast<int64_t>(lo);
}

// K12: copy pose and weight of the selected source (particle_filter.cpp:237).
__global__ void __launch_bounds__(256)
gather_kernel(const float *__restrict__ ix, const float *__restrict__ iy,
              const float *__restrict__ it, const double *__restrict__ iw,
              const int64_t *__restrict__ src, size_t n_out, float *__restrict__ ox,
              float *__restrict__ oy, float *__restrict__ ot, double *__restrict__ ow) {
  const size_t m = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (m >= n_out) return;
  const int64_t s = src[m];
  if (s < 0) return;  // injected slot, already written
  ox[m] = ix[s];
  oy[m] = iy[s];
  ot[m] = it[s];
  ow[m] = iw[s];
}

// ------------------------------------------------------ running-sum lookup
struct PositiveFlag {
  const double *w;
  __device__ uint32_t operator()(size_t i) const { return w[i] > 0 ? 1u : 0u; }
};
// pos[t] = original index of the t-th positive-weight particle and
// incl[t] = running sum after it (= key of the next record).
struct EmitPositive {
  const double *cdf;
  int64_t *pos;
  double *incl;
  __device__ void operator()(size_t i, uint32_t t, uint32_t flag) const {
    if (!flag) return;
    pos[t] = static_cast<int64_t>(i);
    incl[t] = cdf[i];
  }
};

// One adaptive / KLD slot per thread (particle_filter.cpp:261-276, :313-327).
__global__ void __launch_bounds__(256)
draw_kernel(const float *__restrict__ ix, const float *__restrict__ iy,
            const float *__restrict__ it, const double *__restrict__ iw,
            const int64_t *__restrict__ pos, const double *__restrict__ incl, size_t P,
            const unsigned long long *__restrict__ P_dev,
            double w_diff, size_t n_draws, const uint32_t *__restrict__ free_xy,
            unsigned long long n_free, MapGeom g, uint64_t seed, uint32_t step,
            float *__restrict__ ox, float *__restrict__ oy, float *__restrict__ ot,
            double *__restrict__ ow, int64_t *__restrict__ src) {
  const size_t m = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  // The record count may still be on its way to the host.  Without records
  // (the host will find out and discard this output) every slot takes
  // particle 0, which keeps the poses inside the input set's bin bounds.
  if (P_dev) P = static_cast<size_t>(*P_dev);
  // (Staging every 2^k-th key in shared memory, as the fused tail does, costs more than it saves
  // here: 3907 blocks each gather 1024 scattered keys; measured at C3: select 74 -> 93 us.)
  if (m >= n_draws) return;
  uint32_t r[4];
  rng_draw(seed, step, P_DRAW, m, r);
  const double p = uniform53(r[0], r[1]);
  if (p < w_diff) {
    uint32_t q[4];
    rng_draw(seed, step, P_INJECT, m, q);
    float x, y, t;
    random_free_pose(free_xy, n_free, g, q, &x, &y, &t);
    ox[m] = x;
    oy[m] = y;
    ot[m] = t;
    ow[m] = 0.0;  // no influence on the estimate until the next update
    src[m] = -1;
    return;
  }
  const double rr = uniform53(r[2], r[3]);
  const int64_t s = P ? pos[running_sum_lookup(incl, P, rr)] : 0;
  ox[m] = ix[s];
  oy[m] = iy[s];
  ot[m] = it[s];
  ow[m] = iw[s];
  src[m] = s;
}

// ------------------------------------------------------------------------ KLD
// table[cell] = smallest draw index that fell into the cell.
__global__ void __launch_bounds__(256)
kld_first_kernel(const float *__restrict__ x, const float *__restrict__ y,
                 const float *__restrict__ t, size_t n, KldParams kp, BinGrid g,
                 int32_t *__restrict__ table) {
  const size_t m = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (m >= n) return;
  // lanes of a warp that fall into the same bin: the lowest lane has the
  // smallest draw index, one atomic per distinct bin per warp
  const long long cell = kld_cell(x[m], y[m], t[m], kp, g);
  const unsigned peers = __match_any_sync(__activemask(), cell);
  if ((__ffs(peers) - 1) == static_cast<int>(threadIdx.x & 31))
    atomicMin(table + cell, static_cast<int32_t>(m));
}

// flag(m) = draw m opened a new bin.
struct NewBinFlag {
  const float *x, *y, *t;
  KldParams kp;
  BinGrid g;
  const int32_t *table;
  __device__ uint32_t operator()(size_t m) const {
    return table[kld_cell(x[m], y[m], t[m], kp, g)] == static_cast<int32_t>(m) ? 1u : 0u;
  }
};
// After draw m there are k = prefix + flag bins; the loop stops at the first m
// with m >= limit(k) (particle_filter.cpp:329-334).
// The stop word is {m:32, k:32}: ordered by m, and it carries the number of
// bins the reference holds when it stops, so that the host reads both at once.
struct StopRule {
  KldParams kp;
  unsigned long long *first_stop;
  __device__ void operator()(size_t m, uint32_t before, uint32_t flag) const {
    const unsigned long long k = static_cast<unsigned long long>(before) + flag;
    // Every draw past the stop point qualifies, so without aggregation most
    // threads would hit one address: take the warp minimum first (m < 2^31)
    // and skip warps that cannot lower the current value.
    const unsigned cand = (m >= kld_limit(k, kp)) ? static_cast<unsigned>(m) : 0xffffffffu;
    const unsigned mask = __activemask();
    const unsigned wmin = __reduce_min_sync(mask, cand);
    if (wmin != 0xffffffffu && cand == wmin) {
      const unsigned long long word = (static_cast<unsigned long long>(wmin) << 32) | (k & 0xffffffffull);
      if (word < *reinterpret_cast<volatile unsigned long long *>(first_stop)) atomicMin(first_stop, word);
    }
  }
};

// ------------------------------------------------------------------ normalise
// particle.cpp:41-56: total = sum of weights, w /= total.  The sum is a fixed
// tree (deterministic); it differs from the sequential sum by rounding only.
// SQUARE: the sum of squared weights of effective_particles (particle.cpp:32-39),
// each product rounded before it is added, left in sc->sum_sq.
template <bool SQUARE>
__global__ void __launch_bounds__(256)
weight_sum_kernel(const double *__restrict__ w, size_t n, double *__restrict__ partials,
                  FilterScalars *__restrict__ sc) {
  __shared__ bool s_last;
  weight_sum_block<SQUARE>(w, n, partials, blockIdx.x, gridDim.x);
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&sc->blocks_done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const double tot = weight_sum_final(partials, gridDim.x);
  if (threadIdx.x == 0) {
    if (SQUARE) sc->sum_sq = tot;
    else sc->post_total = tot;
    sc->blocks_done = 0;
  }
}

__global__ void __launch_bounds__(256)
divide_by_post_total_kernel(double *__restrict__ w, size_t n, const FilterScalars *__restrict__ sc) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) w[i] = __ddiv_rn(w[i], sc->post_total);
}

__global__ void __launch_bounds__(256)
divide_by_value_kernel(double *__restrict__ w, size_t n, double total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) w[i] = __ddiv_rn(w[i], total);
}

static qmcl_status normalise_current(qmcl_filter *f) {
  ParticleSet &ps = f->cur();
  if (!ps.n && !sharded(f)) return QMCL_OK;
  cudaStream_t s = f->map->stream;
  StageScope sc(f, QMCL_STAGE_POST_NORMALISE);
  const unsigned blocks = static_cast<unsigned>(
      std::min<size_t>(std::max<size_t>(div_up(ps.n, 256 * 4), 1), 4 * kSMs));
  QMCL_TRY(f->partials.ensure(blocks));
  weight_sum_kernel<false><<<blocks, 256, 0, s>>>(ps.w.ptr, ps.n, f->partials.ptr, f->scalars.ptr);
  QMCL_LAUNCHED();
  // Exchange (sharded): the normalising total is the sum of the shards' totals, taken in
  // rank order on the device.
  if (sharded(f)) QMCL_TRY(shard_sum_f64_device(f, &f->scalars.ptr->post_total));
  if (!ps.n) return QMCL_OK;
  divide_by_post_total_kernel<<<div_up(ps.n, 256), 256, 0, s>>>(ps.w.ptr, ps.n, f->scalars.ptr);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

__global__ void fill_default_kernel(float *x, float *y, float *t, double *w, size_t from,
                                    size_t to) {
  const size_t i = from + static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= to) return;
  x[i] = 0.f;
  y[i] = 0.f;
  t[i] = 0.f;
  w[i] = 1.0;  // WeightedParticle's default weight (particle.h:46)
}

// ------------------------------------------------------------------ host side

// particle_filter.cpp:214-239
static qmcl_status resample_low_variance(qmcl_filter *f, const ParticleSet &in, ParticleSet &out) {
  cudaStream_t s = f->map->stream;
  const size_t M = in.n;
  QMCL_TRY(out.ensure(M));
  out.n = M;
  f->n_src_index = M;
  if (M == 0) return QMCL_OK;
  QMCL_TRY(f->cdf.ensure(M));
  QMCL_TRY(f->src_index.ensure(M));
  {
    StageScope sc(f, QMCL_STAGE_CDF);
    QMCL_TRY(sequential_prefix(f, in.w.ptr, M, f->cdf.ptr));
  }
  const double minv = static_cast<double>(1.0f / static_cast<float>(M));
  uint32_t r4[4];
  rng_draw(f->seed, f->step, P_LV_OFFSET, 0, r4);
  const double r = uniform53(r4[0], r4[1]) * minv;
  {
    StageScope sc(f, QMCL_STAGE_SELECT);
    lv_select_kernel<<<div_up(M, 256), 256, 0, s>>>(f->cdf.ptr, M, 0, M, r, minv,
                                                    f->src_index.ptr);
    QMCL_LAUNCHED();
  }
  {
    StageScope sc(f, QMCL_STAGE_GATHER);
    gather_kernel<<<div_up(M, 256), 256, 0, s>>>(in.x.ptr, in.y.ptr, in.t.ptr, in.w.ptr,
                                                 f->src_index.ptr, M, out.x.ptr, out.y.ptr,
                                                 out.t.ptr, out.w.ptr);
    QMCL_LAUNCHED();
  }
  return QMCL_OK;
}

// What a shard tells the others before the chain: its approximate total, its particle
// count and (optionally) the bin bounds of its particles.
struct ShardInfo {
  double approx_total, count;
  int bbox[6];
};
static_assert(sizeof(ShardInfo) == 40, "ShardInfo is exchanged as raw bytes");
__global__ void pack_info_kernel(const double *__restrict__ approx_total, double count,
                                 const int *__restrict__ bbox, ShardInfo *__restrict__ out) {
  out->approx_total = *approx_total;
  out->count = count;
  for (int d = 0; d < 6; ++d) out->bbox[d] = bbox ? bbox[d] : (d < 3 ? INT_MAX : INT_MIN);
}
// The summary's exchange carries the info block along: one host wait for both.
struct ShardSummaryMsg {
  unsigned long long w[kShardSummaryWords];
  ShardInfo info;
};
__global__ void pack_summary_msg_kernel(const unsigned long long *__restrict__ summary,
                                        const ShardInfo *__restrict__ info,
                                        ShardSummaryMsg *__restrict__ out) {
  for (int k = 0; k < kShardSummaryWords; ++k) out->w[k] = summary[k];
  out->info = *info;
}

// The exact (sequentially rounded) running sum of a sharded weight array:
// f->cdf[i] = global running sum after local particle i.  carry[r] = running
// sum before shard r's first particle (carry[G] = grand total), the same bits
// on every rank; count[r] = particles of shard r.  dev_bbox (optional): this
// rank's bin bounds (6 ints, device); their union over the ranks -> bbox_union.
qmcl_status sharded_exact_prefix(qmcl_filter *f, const double *w, size_t n,
                                 std::vector<double> *carry_out, std::vector<double> *count_out,
                                 bool need_total = true, const int *dev_bbox = nullptr,
                                 int *bbox_union = nullptr) {
  cudaStream_t s = f->map->stream;
  const int G = f->n_shards, me = f->shard_rank;
  QMCL_TRY(f->cdf.ensure(n + 1));
  StageScope sc(f, QMCL_STAGE_CDF);
  // approximate shard totals -> approximate carry-ins (binade guesses only).  The
  // totals stay on the device: the summaries' kernel adds up the ones before this
  // rank itself, and the host sees the block with the summaries' exchange.
  QMCL_TRY(prefix_local_total(f, w, n));
  QMCL_TRY(f->partials.ensure(32));
  ShardInfo *d_info = reinterpret_cast<ShardInfo *>(f->partials.ptr);               // 5 doubles
  unsigned long long *d_sum = reinterpret_cast<unsigned long long *>(f->partials.ptr + 8);  // 7 words
  ShardSummaryMsg *d_msg = reinterpret_cast<ShardSummaryMsg *>(f->partials.ptr + 16);     // 12 words
  pack_info_kernel<<<1, 1, 0, s>>>(f->prefix_carry.ptr + 1, static_cast<double>(n), dev_bbox, d_info);
  QMCL_LAUNCHED();
  void *d_all = nullptr;
  QMCL_TRY(shard_gather_device(f, d_info, sizeof(ShardInfo), &d_all));
  QMCL_TRY(prefix_summaries_dev(f, w, n, static_cast<const double *>(d_all), me,
                                static_cast<int>(sizeof(ShardInfo) / sizeof(double))));
  // The exact carry, shard after shard.  A shard publishes a closed-form summary -
  // its particles all fall into one binade of the running sum, or the sum crosses
  // into the next binade once, at a predicted element - which every rank replays
  // locally.  The first shard starts from 0 and runs through dozens of binades: it
  // resolves itself right away and publishes its exact carry-out instead.  Only
  // shards without either (two crossings, a failed prediction) cost a round of the
  // chain: at 8 ranks usually none (3-4 rounds before the crossing form existed).
  bool resolved = false;
  if (me == 0 && n) {
    QMCL_TRY(prefix_resolve(f, w, n, f->cdf.ptr, 0.0));
    QMCL_TRY(prefix_exact_summary(f, d_sum));
    resolved = true;
  } else {
    QMCL_TRY(prefix_shard_summary(f, n, d_sum));
  }
  pack_summary_msg_kernel<<<1, 1, 0, s>>>(d_sum, d_info, d_msg);
  QMCL_LAUNCHED();
  std::vector<ShardSummaryMsg> msgs(G);
  QMCL_TRY(shard_gather(f, d_msg, true, sizeof(ShardSummaryMsg), msgs.data()));
  if (bbox_union) {
    for (int d = 0; d < 6; ++d) bbox_union[d] = msgs[0].info.bbox[d];
    for (int r = 1; r < G; ++r)
      for (int d = 0; d < 3; ++d) {
        bbox_union[d] = std::min(bbox_union[d], msgs[r].info.bbox[d]);
        bbox_union[3 + d] = std::max(bbox_union[3 + d], msgs[r].info.bbox[3 + d]);
      }
  }
  std::vector<double> &carry = *carry_out;
  carry.assign(G + 1, 0.0);
  std::vector<double> round(G);
  for (int r = 0; r < G; ++r) {
    const bool known = prefix_apply_summary(carry[r], msgs[r].w, &carry[r + 1]);
    // this rank's own walk starts as soon as its carry-in is exact: it then runs
    // beside whatever rounds later shards still need
    if (r == me && !resolved) {
      QMCL_TRY(prefix_resolve(f, w, n, f->cdf.ptr, carry[r]));
      resolved = true;
    }
    if (known) continue;
    // The last shard's carry-out is the grand total.  A caller that does not use it
    // (low variance: the last shard owns every slot above its carry-in) is spared the
    // round; with weights that add up to 1 that shard fails often - whether the sum
    // reaches the next power of two at its very end is beyond any prediction.
    if (r == G - 1 && !need_total) {
      carry[G] = carry[G - 1];
      continue;
    }
    QMCL_TRY(shard_gather(f, f->prefix_carry.ptr, true, sizeof(double), round.data()));
    carry[r + 1] = round[r];
    f->n_chain_rounds++;
  }
  if (!resolved) QMCL_TRY(prefix_resolve(f, w, n, f->cdf.ptr, carry[me]));
  QMCL_TRY(prefix_apply(f, w, n, f->cdf.ptr));
  count_out->resize(G);
  for (int r = 0; r < G; ++r) (*count_out)[r] = msgs[r].info.count;
  return QMCL_OK;
}

// Moves particles so that every rank owns its equal contiguous block of the M
// global slots again.  lo/hi: the slot range [lo[r], hi[r]) rank r holds now
// (consecutive, tiling [0, M)); `set` holds this rank's slots in order.
qmcl_status rebalance_blocks(qmcl_filter *f, ParticleSet &set, const std::vector<uint64_t> &lo,
                             const std::vector<uint64_t> &hi, size_t M) {
  cudaStream_t s = f->map->stream;
  const int G = f->n_shards, me = f->shard_rank;
  bool balanced = true;
  for (int r = 0; r < G; ++r) {
    size_t a, c;
    shard_slice_of(M, r, G, &a, &c);
    balanced &= (lo[r] == a && hi[r] == a + c);
  }
  if (!f->rebalance || balanced) return QMCL_OK;
  // Only worth an all-to-all when some rank has drifted from its equal share by more than the
  // threshold (every rank evaluates the same table, so all take the same branch).  Blocks stay
  // contiguous and in rank order either way; unequal blocks are what the no-rebalance mode runs on.
  if (f->rebalance_threshold > 0.0) {
    double worst = 0.0;
    for (int r = 0; r < G; ++r) {
      size_t a, c;
      shard_slice_of(M, r, G, &a, &c);
      const double have = static_cast<double>(hi[r] - lo[r]);
      const double want = static_cast<double>(std::max<size_t>(c, 1));
      worst = std::max(worst, std::fabs(have - want) / want);
    }
    if (worst <= f->rebalance_threshold) {
      f->n_rebalance_skipped++;
      return QMCL_OK;
    }
  }
  StageScope sc(f, QMCL_STAGE_GATHER);
  const size_t n_have = static_cast<size_t>(hi[me] - lo[me]);
  size_t my_a, my_c;
  shard_slice_of(M, me, G, &my_a, &my_c);
  std::vector<uint64_t> sb(G), so(G), rb(G), ro(G);
  constexpr uint64_t kRec = sizeof(qmcl_particle);
  for (int d = 0; d < G; ++d) {
    size_t a, c;
    shard_slice_of(M, d, G, &a, &c);
    // what I send to d: [lo[me], hi[me]) ∩ [a, a+c)
    const uint64_t l = std::max<uint64_t>(lo[me], a), h = std::min<uint64_t>(hi[me], a + c);
    sb[d] = h > l ? (h - l) * kRec : 0;
    so[d] = h > l ? (l - lo[me]) * kRec : 0;
    // what I receive from d: [lo[d], hi[d]) ∩ [my_a, my_a+my_c)
    const uint64_t rl = std::max<uint64_t>(lo[d], my_a), rh = std::min<uint64_t>(hi[d], my_a + my_c);
    rb[d] = rh > rl ? (rh - rl) * kRec : 0;
    ro[d] = rh > rl ? (rl - my_a) * kRec : 0;
  }
  QMCL_TRY(f->xchg_send.ensure(std::max<size_t>(n_have, 1) * kRec));
  QMCL_TRY(f->xchg_recv.ensure(std::max<size_t>(my_c, 1) * kRec));
  QMCL_TRY(launch_soa_to_aos(s, set.x.ptr, set.y.ptr, set.t.ptr, set.w.ptr, n_have,
                             reinterpret_cast<qmcl_particle *>(f->xchg_send.ptr)));
  if (f->comm.alltoallv(f->comm.ctx, f->xchg_send.ptr, sb.data(), so.data(), f->xchg_recv.ptr,
                        rb.data(), ro.data(), s) != 0) {
    set_error("qmcl_comm.alltoallv failed");
    return QMCL_ERR_COMM;
  }
  f->n_collectives++;
  QMCL_TRY(set.ensure(my_c));
  set.n = my_c;
  QMCL_TRY(launch_aos_to_soa(s, reinterpret_cast<const qmcl_particle *>(f->xchg_recv.ptr), my_c,
                             set.x.ptr, set.y.ptr, set.t.ptr, set.w.ptr));
  f->global_first = my_a;
  return QMCL_OK;
}

// particle_filter.cpp:214-239 over a sharded particle set.  The
// global CDF is the concatenation of the shards' running sums, chained through
// one exact carry per shard; the output slots are global, and every shard
// produces the slots whose position falls into its own CDF segment
// (owner-computes: no particle leaves its GPU during selection).  The optional
// rebalancing afterwards restores equal contiguous blocks with one all-to-all.
static qmcl_status resample_low_variance_sharded(qmcl_filter *f, const ParticleSet &in,
                                                 ParticleSet &out) {
  cudaStream_t s = f->map->stream;
  const int G = f->n_shards, me = f->shard_rank;
  const size_t M = f->global_n;
  f->n_src_index = 0;
  if (M == 0) {
    out.n = 0;
    return QMCL_OK;
  }
  std::vector<double> carry, count;
  // The new set only copies input particles, so the union of the input shards' bin
  // bounds covers it: they ride along with the running sum's exchange and spare the
  // clustering that follows an exchange (and a host wait) of its own.
  QMCL_TRY(launch_bin_bbox(f, in.x.ptr, in.y.ptr, in.t.ptr, in.n));
  int in_bbox[6];
  QMCL_TRY(sharded_exact_prefix(f, in.w.ptr, in.n, &carry, &count, /*need_total=*/false, f->bbox.ptr,
                                in_bbox));
  if (in_bbox[0] <= in_bbox[3]) {
    std::memcpy(f->bbox_hint, in_bbox, sizeof(in_bbox));
    f->bbox_hint_valid = true;
  }
  struct Info { double count; };
  std::vector<Info> info(G);
  for (int r = 0; r < G; ++r) info[r].count = count[r];
  // 3. owner ranges of all shards (every rank computes the same table)
  const double minv = static_cast<double>(1.0f / static_cast<float>(M));
  uint32_t r4[4];
  rng_draw(f->seed, f->step, P_LV_OFFSET, 0, r4);
  const double r0 = uniform53(r4[0], r4[1]) * minv;
  int first_rank = -1, last_rank = -1;
  for (int r = 0; r < G; ++r)
    if (info[r].count > 0) {
      if (first_rank < 0) first_rank = r;
      last_rank = r;
    }
  std::vector<uint64_t> m_lo(G, 0), m_hi(G, 0);
  for (int r = 0; r < G; ++r) {
    if (r < first_rank) {
      m_lo[r] = m_hi[r] = 0;
    } else if (r > last_rank) {
      m_lo[r] = m_hi[r] = M;
    } else {
      // empty shards in between have carry[r] == carry[r + 1]: an empty range
      qmcl_lv_owner_range(M, r0, minv, carry[r], carry[r + 1], r == first_rank, r == last_rank,
                          &m_lo[r], &m_hi[r]);
    }
  }
  const size_t n_out = static_cast<size_t>(m_hi[me] - m_lo[me]);
  QMCL_TRY(out.ensure(n_out));
  out.n = n_out;
  QMCL_TRY(f->src_index.ensure(n_out + 1));
  if (n_out) {
    {
      StageScope sc(f, QMCL_STAGE_SELECT);
      lv_select_kernel<<<div_up(n_out, 256), 256, 0, s>>>(f->cdf.ptr, in.n, m_lo[me], n_out, r0, minv,
                                                        f->src_index.ptr);
      QMCL_LAUNCHED();
    }
    StageScope sc(f, QMCL_STAGE_GATHER);
    gather_kernel<<<div_up(n_out, 256), 256, 0, s>>>(in.x.ptr, in.y.ptr, in.t.ptr, in.w.ptr,
                                                     f->src_index.ptr, n_out, out.x.ptr, out.y.ptr,
                                                     out.t.ptr, out.w.ptr);
    QMCL_LAUNCHED();
  }
  f->n_src_index = n_out;
  f->global_first = m_lo[me];
  f->global_n = M;
  // 4. rebalance to equal blocks
  return rebalance_blocks(f, out, m_lo, m_hi, M);
}

// Builds the running-sum records of `in`; *P_out = number of records,
// *total_out = ParticlesRunningSum::get_total_weight().
// bbox_out (optional, single GPU): the bin bounding box of the input set is
// computed in the same pass and read back with the same synchronisation; it
// bounds the bins of every draw that copies an input particle.
// Enqueues the running sum of `in` (in.n > 0): records in f->pos / f->incl, their
// number in the stable device word scalars->n_records, and copies of that number
// and of the total weight on their way to HS_RS.  No synchronisation.
static qmcl_status enqueue_running_sum(qmcl_filter *f, const ParticleSet &in) {
  cudaStream_t s = f->map->stream;
  QMCL_TRY(f->cdf.ensure(in.n));
  QMCL_TRY(f->pos.ensure(in.n));
  QMCL_TRY(f->incl.ensure(in.n));
  StageScope sc(f, QMCL_STAGE_CDF);
  QMCL_TRY(sequential_prefix(f, in.w.ptr, in.n, f->cdf.ptr));
  QMCL_TRY(f->scan_tmp.ensure_zero(scan_tmp_words(in.n) + 4));
  unsigned long long *d_total = &f->scalars.ptr->n_records;
  QMCL_TRY(device_scan(s, PositiveFlag{in.w.ptr}, EmitPositive{f->cdf.ptr, f->pos.ptr, f->incl.ptr},
                       in.n, f->scan_tmp.ptr, d_total));
  double *h = f->h_scalars.ptr + HS_RS;
  QMCL_CUDA(copy_d2h(h, d_total, sizeof(unsigned long long), s));
  QMCL_CUDA(copy_d2h(h + 1, f->cdf.ptr + (in.n - 1), sizeof(double), s));
  return QMCL_OK;
}
// What enqueue_running_sum left in HS_RS, once the stream has been synchronised.
static void read_running_sum(const qmcl_filter *f, size_t *P_out, double *total_out) {
  const double *h = f->h_scalars.ptr + HS_RS;
  unsigned long long P;
  std::memcpy(&P, h, sizeof(P));
  *P_out = P;
  *total_out = h[1];
}

static qmcl_status build_running_sum(qmcl_filter *f, const ParticleSet &in, size_t *P_out,
                                     double *total_out, int *bbox_out = nullptr) {
  cudaStream_t s = f->map->stream;
  *P_out = 0;
  *total_out = 0.0;
  if (bbox_out) bbox_out[0] = 1, bbox_out[3] = 0;  // empty
  if (in.n == 0) return settle(f);
  QMCL_TRY(enqueue_running_sum(f, in));
  if (bbox_out) {
    QMCL_TRY(launch_bin_bbox(f, in.x.ptr, in.y.ptr, in.t.ptr, in.n));
    QMCL_CUDA(copy_d2h(f->h_scalars.ptr + HS_BBOX, f->bbox.ptr, 6 * sizeof(int), s));
  }
  QMCL_CUDA(cudaStreamSynchronize(s));
  settle_after_sync(f);  // the sensor total (w_slow, w_fast) rides on this synchronisation
  read_running_sum(f, P_out, total_out);
  if (bbox_out) std::memcpy(bbox_out, f->h_scalars.ptr + HS_BBOX, 6 * sizeof(int));
  return QMCL_OK;
}

// particle_filter.cpp:284-290, 338-344: both filters restart after an injection round.
static void reset_lowpass(qmcl_filter *f) {
  f->w_slow_before_reset = f->w_slow;
  f->w_fast_before_reset = f->w_fast;
  f->lowpass_was_reset = true;
  f->w_fast = 0;
  f->w_slow = 0;
}

static double adaptiveness(const qmcl_filter *f) {
  // particle_filter.cpp:246: std::max(0.0, NaN) is 0.0
  return std::max(0.0, 1.0 - f->w_fast / f->w_slow);
}

static qmcl_status launch_draws(qmcl_filter *f, const ParticleSet &in, ParticleSet &out, size_t P,
                                double w_diff, size_t n_draws,
                                const unsigned long long *P_dev = nullptr) {
  cudaStream_t s = f->map->stream;
  if (w_diff > 0 && f->map->n_free == 0) {
    set_error("adaptive injection needs a map with free cells");
    return QMCL_ERR_NO_MAP;
  }
  QMCL_TRY(f->src_index.ensure(n_draws));
  StageScope sc(f, QMCL_STAGE_SELECT);
  draw_kernel<<<div_up(n_draws, 256), 256, 0, s>>>(
      in.x.ptr, in.y.ptr, in.t.ptr, in.w.ptr, f->pos.ptr, f->incl.ptr, P, P_dev, w_diff, n_draws,
      f->map->free_cells.ptr, f->map->n_free, f->map->geom, f->seed, f->step, out.x.ptr, out.y.ptr,
      out.t.ptr, out.w.ptr, f->src_index.ptr);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

// particle_filter.cpp:241-291
static qmcl_status resample_adaptive(qmcl_filter *f, const ParticleSet &in, ParticleSet &out) {
  cudaStream_t s = f->map->stream;
  const size_t n_out = static_cast<size_t>(f->params.particle_count_min);
  // output_particles->resize(output_size): keeps what the buffer held, new
  // slots are default particles
  const size_t old_n = out.n;
  QMCL_TRY(out.x.ensure_keep(n_out, old_n, s));
  QMCL_TRY(out.y.ensure_keep(n_out, old_n, s));
  QMCL_TRY(out.t.ensure_keep(n_out, old_n, s));
  QMCL_TRY(out.w.ensure_keep(n_out, old_n, s));
  if (n_out > old_n) {
    fill_default_kernel<<<div_up(n_out - old_n, 256), 256, 0, s>>>(out.x.ptr, out.y.ptr, out.t.ptr,
                                                                   out.w.ptr, old_n, n_out);
    QMCL_LAUNCHED();
  }
  out.n = n_out;
  f->n_src_index = 0;
  size_t P;
  double total;
  int in_bbox[6];
  QMCL_TRY(build_running_sum(f, in, &P, &total, lazy(f) ? in_bbox : nullptr));
  if (!(total > 0)) return QMCL_OK;  // "No particles have any weights! Failing to resample"
  const double w_diff = adaptiveness(f);  // after the read-back: it settles w_slow / w_fast
  if (n_out) QMCL_TRY(launch_draws(f, in, out, P, w_diff, n_out));
  // Without injection the new set copies input particles only: the input set's
  // bounding box (already on the host) serves the clustering that follows.
  if (lazy(f) && n_out && !(w_diff > 0.0) && in_bbox[0] <= in_bbox[3]) {
    std::memcpy(f->bbox_hint, in_bbox, sizeof(in_bbox));
    f->bbox_hint_valid = true;
  }
  f->n_src_index = n_out;
  if (w_diff > 0.0) {
    reset_lowpass(f);
  }
  return QMCL_OK;
}

// particle_filter.cpp:293-348.  All particle_count_max candidate draws are
// made at once (counter-based RNG: draw m does not depend on draw m-1); the
// sequential stop rule "first m with m >= limit(bins after m)" becomes a
// first-occurrence mark per bin, a prefix count and a min-reduction.
static qmcl_status resample_kld(qmcl_filter *f, const ParticleSet &in, ParticleSet &out,
                                bool *clustered) {
  cudaStream_t s = f->map->stream;
  *clustered = false;
  const size_t n_max = static_cast<size_t>(f->params.particle_count_max);
  out.n = 0;  // output_set->particles.clear()
  f->n_src_index = 0;
  f->n_bins = 0;  // space_partitioning.reset()
  f->bins_pending = false;
  f->n_clusters = 0;
  f->clusters_pending = false;  // an older count still in flight is void
  f->bins_built = true;
  f->grid = BinGrid{0, 0, 0, 0, 0, 0};
  size_t P = 0;
  double total = 0.0;
  int in_bbox[6] = {1, 0, 0, 0, 0, 0};  // empty
  auto bbox_usable = [&]() {
    if (in_bbox[0] > in_bbox[3]) return false;
    const double cells_in = (static_cast<double>(in_bbox[3]) - in_bbox[0] + 1) *
                            (static_cast<double>(in_bbox[4]) - in_bbox[1] + 1) *
                            (static_cast<double>(in_bbox[5]) - in_bbox[2] + 1);
    return cells_in <= static_cast<double>(kMaxTableCells);
  };
  // Streamlined order (single GPU, deferred read-backs): the input set's bin
  // bounds were taken during the sensor update, so nothing here has to wait for
  // the running sum — it is enqueued, the host picks up the sensor total as soon
  // as the sensor kernel is through (an event, not a stream synchronisation) and
  // goes on enqueueing while the running sum executes.  Whether there was
  // anything to draw from is checked at the stop-rule read-back below.
  bool check_later = false;
  double w_diff = 0.0;
  if (lazy(f) && in.n && n_max > 0 && n_max <= 0x7fffffffull &&
      f->early_bbox_version == f->pose_version) {
    QMCL_TRY(enqueue_running_sum(f, in));
    if (f->total_pending) {
      QMCL_CUDA(cudaEventSynchronize(f->total_copied));
      // everything pending was enqueued before that event
      settle_after_sync(f);
    }
    std::memcpy(in_bbox, f->h_scalars.ptr + HS_BBOX_EARLY, sizeof(in_bbox));
    w_diff = adaptiveness(f);
    if (!(w_diff > 0.0) && bbox_usable()) {
      check_later = true;
    } else {  // injected poses leave the input set's bounds: wait for the candidates' own
      QMCL_CUDA(cudaStreamSynchronize(s));
      settle_after_sync(f);
      read_running_sum(f, &P, &total);
    }
  } else {
    QMCL_TRY(build_running_sum(f, in, &P, &total, lazy(f) ? in_bbox : nullptr));
    w_diff = adaptiveness(f);  // after the read-back: it settles w_slow / w_fast
  }
  if (!check_later && !(total > 0)) return QMCL_OK;
  if (n_max == 0) return QMCL_OK;
  if (n_max > 0x7fffffffull) return QMCL_ERR_CAPACITY;
  QMCL_TRY(out.ensure(n_max));
  QMCL_TRY(launch_draws(f, in, out, P, w_diff, n_max,
                        check_later ? &f->scalars.ptr->n_records : nullptr));
  // The dense bin grid spanned by all candidates.  Without injection every
  // candidate is a copy of an input particle, so the input set's bounding box
  // contains it.
  const int *known_bbox = nullptr;
  if (lazy(f) && !(w_diff > 0.0) && bbox_usable()) known_bbox = in_bbox;
  QMCL_TRY(build_bins(f, out.x.ptr, out.y.ptr, out.t.ptr, n_max, false, true, known_bbox));
  const BinGrid g = f->grid;
  const size_t cells = static_cast<size_t>(g.nx) * g.ny * g.nt;
  KldParams kp;
  kp.res_xy = f->params.res_xy;
  kp.res_t = f->params.res_theta;
  kp.eps2 = 2 * f->params.kld_epsilon;  // space_partitioning.cpp:31
  kp.z = f->params.kld_z;
  kp.nmin = f->params.particle_count_min;
  kp.nmax = f->params.particle_count_max;
  QMCL_TRY(f->scan_tmp.ensure_zero(scan_tmp_words(n_max) + 8));
  unsigned long long *d_total = &f->scalars.ptr->kld_new;
  unsigned long long *d_stop = &f->scalars.ptr->kld_stop;
  {
    StageScope sc(f, QMCL_STAGE_KLD);
    QMCL_CUDA(cudaMemsetAsync(f->table.ptr, 0x7f, cells * sizeof(int32_t), s));
    kld_first_kernel<<<div_up(n_max, 256), 256, 0, s>>>(out.x.ptr, out.y.ptr, out.t.ptr, n_max, kp,
                                                        g, f->table.ptr);
    QMCL_LAUNCHED();
    QMCL_CUDA(cudaMemsetAsync(d_stop, 0xff, sizeof(unsigned long long), s));
    QMCL_TRY(device_scan(s, NewBinFlag{out.x.ptr, out.y.ptr, out.t.ptr, kp, g, f->table.ptr},
                         StopRule{kp, d_stop}, n_max, f->scan_tmp.ptr, d_total));
  }
  unsigned long long *h = reinterpret_cast<unsigned long long *>(f->h_scalars.ptr + HS_STOP);
  QMCL_CUDA(copy_d2h(h, d_stop, sizeof(unsigned long long), s));
  QMCL_CUDA(copy_d2h(h + 1, d_total, sizeof(unsigned long long), s));
  QMCL_CUDA(cudaStreamSynchronize(s));
  settle_after_sync(f);
  if (check_later) {
    read_running_sum(f, &P, &total);
    if (!(total > 0)) {  // nothing to draw from: the state of the early return above
      f->grid = BinGrid{0, 0, 0, 0, 0, 0};
      f->n_bins = 0;
      f->bins_built = true;
      return QMCL_OK;
    }
  }
  // no stop (all ones): every candidate is kept and every new-bin flag counts
  const bool stopped = h[0] != ~0ull;
  const size_t n_out = stopped ? static_cast<size_t>(h[0] >> 32) + 1 : n_max;
  const long long n_bins_kept = static_cast<long long>(stopped ? (h[0] & 0xffffffffull) : h[1]);
  out.n = n_out;
  f->n_src_index = n_out;
  if (w_diff > 0.0) {
    reset_lowpass(f);
  }
  // the bins the reference holds at this point are those of the kept draws; they
  // lie inside the candidates' bounding box, so the grid is reused
  QMCL_TRY(build_bins(f, out.x.ptr, out.y.ptr, out.t.ptr, n_out, true, false, nullptr,
                      lazy(f) ? n_bins_kept : -1));
  QMCL_TRY(label_clusters(f));
  *clustered = true;
  return QMCL_OK;
}

// ===================================================================== sharded
// Adaptive and KLD resampling over a sharded particle set.  Draws
// come from the counter-based stream, so every rank can enumerate all of them;
// a rank keeps the draws whose running-sum record lies in its own block
// (owner computes), one all-to-all then delivers every output slot m to the
// rank that owns slot m's block, and the KLD stop rule runs on the
// slot-ordered shards with a min-allreduce of the per-bin first draw index.

constexpr int kMaxShards = 32;

// Keys (running sum before a positive-weight particle) at the edges of each
// shard's record list, the same on every rank.
struct ShardKeys {
  int G, me;
  unsigned long long per;  // slots per block (ceil)
  int has[kMaxShards];
  double first[kMaxShards];   // key of the shard's first record
  double second[kMaxShards];  // smallest key > first (inf if none)
  double last[kMaxShards];    // key of its last record
};

// The shard holding the record ParticlesRunningSum::lookup_index(r) returns:
// the record with the largest key <= r, the first one among equal keys
// (particle.h:99-107, particle.cpp:77).  Equal keys across a shard boundary
// happen when the weights in between are too small to move the sum.
__device__ __forceinline__ int draw_owner(const ShardKeys &T, double rr) {
  int p = 0;
  for (int s = 0; s < T.G; ++s)
    if (T.has[s] && T.first[s] <= rr) p = s;
  if (rr < T.second[p]) {  // the record found has key == first[p]: look for earlier equals
    const double K = T.first[p];
    for (int s = p - 1; s >= 0; --s) {
      if (!T.has[s]) continue;
      if (T.last[s] != K) break;
      p = s;
      if (T.first[s] != K) break;
    }
  }
  return p;
}
__device__ __forceinline__ int slot_owner(const ShardKeys &T, size_t m) {
  const unsigned long long o = m / T.per;
  return static_cast<int>(o < static_cast<unsigned long long>(T.G) ? o : T.G - 1);
}

struct SlotRec {  // one output slot in flight (32 bytes)
  long long m;
  float x, y, t;
  uint32_t src;  // global source index, 0xffffffff = injected
  double w;
};

struct DrawCtx {
  ShardKeys T;
  uint64_t seed;
  uint32_t step;
  double w_diff;
};

struct OwnedFlag {
  DrawCtx c;
  __device__ uint32_t operator()(size_t m) const {
    uint32_t r[4];
    rng_draw(c.seed, c.step, P_DRAW, m, r);
    if (uniform53(r[0], r[1]) < c.w_diff) return slot_owner(c.T, m) == c.T.me ? 1u : 0u;
    return draw_owner(c.T, uniform53(r[2], r[3])) == c.T.me ? 1u : 0u;
  }
};

struct EmitOwned {
  DrawCtx c;
  const float *ix, *iy, *it;
  const double *iw;
  const int64_t *pos;
  const double *incl;
  size_t P;
  size_t global_first;
  const uint32_t *free_xy;
  unsigned long long n_free;
  MapGeom g;
  SlotRec *rec;
  __device__ void operator()(size_t m, uint32_t at, uint32_t flag) const {
    if (!flag) return;
    uint32_t r[4];
    rng_draw(c.seed, c.step, P_DRAW, m, r);
    SlotRec o;
    o.m = static_cast<long long>(m);
    if (uniform53(r[0], r[1]) < c.w_diff) {
      uint32_t q[4];
      rng_draw(c.seed, c.step, P_INJECT, m, q);
      random_free_pose(free_xy, n_free, g, q, &o.x, &o.y, &o.t);
      o.w = 0.0;
      o.src = 0xffffffffu;
    } else {
      const int64_t sidx = pos[running_sum_lookup(incl, P, uniform53(r[2], r[3]))];
      o.x = ix[sidx];
      o.y = iy[sidx];
      o.t = it[sidx];
      o.w = iw[sidx];
      o.src = static_cast<uint32_t>(global_first + static_cast<size_t>(sidx));
    }
    rec[at] = o;
  }
};

// {has, first, second, last} of this shard's record list.
__global__ void shard_keys_kernel(const double *__restrict__ incl, size_t P, double carry_in,
                                  double *__restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  out[0] = P ? 1.0 : 0.0;
  out[1] = carry_in;
  // keys: carry_in, incl[0], ..., incl[P-2]; second = first key > carry_in
  size_t lo = 0, hi = P ? P - 1 : 0;
  while (lo < hi) {
    const size_t mid = (lo + hi) >> 1;
    if (incl[mid] > carry_in) hi = mid;
    else lo = mid + 1;
  }
  out[2] = (P >= 2 && lo < P - 1) ? incl[lo] : __longlong_as_double(0x7ff0000000000000ll);
  out[3] = (P >= 2) ? incl[P - 2] : carry_in;
}

// bounds[d] = first record whose slot is >= d * per (records are sorted by slot).
__global__ void dest_bounds_kernel(const SlotRec *__restrict__ rec,
                                   const unsigned long long *__restrict__ n_rec_dev,
                                   unsigned long long per, int G,
                                   unsigned long long *__restrict__ bounds) {
  const int d = threadIdx.x;
  if (d > G) return;
  const size_t n_rec = static_cast<size_t>(*n_rec_dev);
  const long long key = static_cast<long long>(per) * d;
  size_t lo = 0, hi = n_rec;
  if (d == G) lo = n_rec;
  while (lo < hi) {
    const size_t mid = (lo + hi) >> 1;
    if (rec[mid].m < key) lo = mid + 1;
    else hi = mid;
  }
  bounds[d] = lo;
}

__global__ void __launch_bounds__(256)
scatter_slots_kernel(const SlotRec *__restrict__ rec, size_t n_rec, size_t slot_first,
                     float *__restrict__ x, float *__restrict__ y, float *__restrict__ t,
                     double *__restrict__ w, int64_t *__restrict__ src) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_rec) return;
  const SlotRec r = rec[i];
  const size_t k = static_cast<size_t>(r.m) - slot_first;
  x[k] = r.x;
  y[k] = r.y;
  t[k] = r.t;
  w[k] = r.w;
  src[k] = r.src == 0xffffffffu ? -1 : static_cast<int64_t>(r.src);
}

// Fills `out` with this rank's block of the n_draws output slots.  *ok = false
// (on every rank alike) when no particle has weight.
static qmcl_status resample_draws_sharded(qmcl_filter *f, const ParticleSet &in, ParticleSet &out,
                                          size_t n_draws, double w_diff, bool *ok) {
  cudaStream_t s = f->map->stream;
  const int G = f->n_shards, me = f->shard_rank;
  *ok = false;
  if (G > kMaxShards) {
    set_error("sharded adaptive/KLD resampling supports at most %d shards", kMaxShards);
    return QMCL_ERR_CAPACITY;
  }
  size_t my_first, my_count;
  shard_slice_of(n_draws, me, G, &my_first, &my_count);
  // 1. exact global running sum + local records
  std::vector<double> carry, count;
  QMCL_TRY(sharded_exact_prefix(f, in.w.ptr, in.n, &carry, &count));
  if (!(carry[G] > 0)) return QMCL_OK;  // "No particles have any weights!"
  if (w_diff > 0 && f->map->n_free == 0) {
    set_error("adaptive injection needs a map with free cells");
    return QMCL_ERR_NO_MAP;
  }
  QMCL_TRY(f->pos.ensure(in.n + 1));
  QMCL_TRY(f->incl.ensure(in.n + 1));
  QMCL_TRY(f->scan_tmp.ensure_zero(scan_tmp_words(std::max(in.n, n_draws)) + 16 + 2 * (kMaxShards + 2)));
  unsigned long long P = 0;
  {
    StageScope sc(f, QMCL_STAGE_CDF);
    unsigned long long *d_total = scan_total_ptr(f->scan_tmp.ptr, std::max(in.n, n_draws));
    QMCL_TRY(device_scan(s, PositiveFlag{in.w.ptr}, EmitPositive{f->cdf.ptr, f->pos.ptr, f->incl.ptr},
                         in.n, f->scan_tmp.ptr, d_total));
    QMCL_CUDA(copy_d2h(&P, d_total, sizeof(P), s));
    QMCL_CUDA(cudaStreamSynchronize(s));
  }
  // 2. the shards' edge keys
  struct Keys { double has, first, second, last; };
  std::vector<Keys> keys(G);
  {
    QMCL_TRY(f->partials.ensure(8));
    shard_keys_kernel<<<1, 32, 0, s>>>(f->incl.ptr, P, carry[me], f->partials.ptr);
    QMCL_LAUNCHED();
    QMCL_TRY(shard_gather(f, f->partials.ptr, true, sizeof(Keys), keys.data()));
  }
  DrawCtx ctx;
  ctx.T.G = G;
  ctx.T.me = me;
  ctx.T.per = std::max<size_t>((n_draws + G - 1) / G, 1);
  for (int r = 0; r < G; ++r) {
    ctx.T.has[r] = keys[r].has != 0.0;
    ctx.T.first[r] = keys[r].first;
    ctx.T.second[r] = keys[r].second;
    ctx.T.last[r] = keys[r].last;
  }
  ctx.seed = f->seed;
  ctx.step = f->step;
  ctx.w_diff = w_diff;
  // 3. the draws this rank owns, compacted in slot order
  StageScope sc(f, QMCL_STAGE_SELECT);
  const OwnedFlag flag{ctx};
  const size_t worst = n_draws * sizeof(SlotRec);
  unsigned long long *d_total = nullptr;
  std::vector<unsigned long long> bounds(G + 1);
  if (worst <= (size_t(1) << 30)) {
    // One pass: the send buffer is sized for the case that this rank owns every
    // draw (<= 1 GiB), so the owner test and the Philox draw run once per slot
    // and the count comes back together with the destination bounds.
    QMCL_TRY(f->xchg_send.ensure(worst + sizeof(SlotRec)));
    QMCL_TRY(f->scan_tmp.ensure_zero(scan_tmp_words(n_draws) + 2 * (kMaxShards + 4)));
    d_total = scan_total_ptr(f->scan_tmp.ptr, n_draws);
    SlotRec *send = reinterpret_cast<SlotRec *>(f->xchg_send.ptr);
    const EmitOwned emit{ctx,        in.x.ptr,          in.y.ptr,
                         in.t.ptr,   in.w.ptr,          f->pos.ptr,
                         f->incl.ptr, static_cast<size_t>(P), f->global_first,
                         f->map->free_cells.ptr, f->map->n_free, f->map->geom, send};
    QMCL_TRY(device_scan(s, flag, emit, n_draws, f->scan_tmp.ptr, d_total));
  } else {
    // (three-kernel scan: the host needs the count between the two passes; its
    // block sums live in their own buffer, never in the status words of scan_tmp)
    const unsigned blocks = scan_blocks(n_draws);
    QMCL_TRY(f->scan_legacy.ensure(blocks + 16 + 2 * (kMaxShards + 2)));
    d_total = reinterpret_cast<unsigned long long *>(f->scan_legacy.ptr + ((blocks + 1) & ~1u));
    scan_reduce_kernel<<<blocks, kScanBlock, 0, s>>>(flag, n_draws, f->scan_legacy.ptr);
    QMCL_LAUNCHED();
    scan_block_sums_kernel<0><<<1, 1024, 0, s>>>(f->scan_legacy.ptr, blocks, d_total);
    QMCL_LAUNCHED();
    unsigned long long n_owned = 0;
    QMCL_CUDA(copy_d2h(&n_owned, d_total, sizeof(n_owned), s));
    QMCL_CUDA(cudaStreamSynchronize(s));
    QMCL_TRY(f->xchg_send.ensure((n_owned + 1) * sizeof(SlotRec)));
    SlotRec *send = reinterpret_cast<SlotRec *>(f->xchg_send.ptr);
    const EmitOwned emit{ctx,        in.x.ptr,          in.y.ptr,
                         in.t.ptr,   in.w.ptr,          f->pos.ptr,
                         f->incl.ptr, static_cast<size_t>(P), f->global_first,
                         f->map->free_cells.ptr, f->map->n_free, f->map->geom, send};
    scan_apply_kernel<<<blocks, kScanBlock, 0, s>>>(flag, emit, n_draws, f->scan_legacy.ptr);
    QMCL_LAUNCHED();
  }
  // 4. deliver every slot to the rank that owns its block
  {
    unsigned long long *d_bounds = d_total + 2;  // inside the allocations made above
    dest_bounds_kernel<<<1, 64, 0, s>>>(reinterpret_cast<const SlotRec *>(f->xchg_send.ptr), d_total,
                                        ctx.T.per, G, d_bounds);
    QMCL_LAUNCHED();
    QMCL_CUDA(copy_d2h(bounds.data(), d_bounds, (G + 1) * sizeof(unsigned long long), s));
    QMCL_CUDA(cudaStreamSynchronize(s));
  }
  std::vector<uint64_t> sb(G), so(G), rb(G), ro(G), matrix(static_cast<size_t>(G) * G);
  for (int d = 0; d < G; ++d) {
    sb[d] = (bounds[d + 1] - bounds[d]) * sizeof(SlotRec);
    so[d] = bounds[d] * sizeof(SlotRec);
  }
  QMCL_TRY(shard_gather(f, sb.data(), false, G * sizeof(uint64_t), matrix.data()));
  uint64_t at = 0;
  for (int r = 0; r < G; ++r) {
    rb[r] = matrix[static_cast<size_t>(r) * G + me];
    ro[r] = at;
    at += rb[r];
  }
  if (at != my_count * sizeof(SlotRec)) {
    set_error("sharded resampling: rank %d received %llu slots, expected %llu", me,
              static_cast<unsigned long long>(at / sizeof(SlotRec)),
              static_cast<unsigned long long>(my_count));
    return QMCL_ERR_COMM;
  }
  QMCL_TRY(f->xchg_recv.ensure((my_count + 1) * sizeof(SlotRec)));
  if (f->comm.alltoallv(f->comm.ctx, f->xchg_send.ptr, sb.data(), so.data(), f->xchg_recv.ptr,
                        rb.data(), ro.data(), s) != 0) {
    set_error("qmcl_comm.alltoallv failed");
    return QMCL_ERR_COMM;
  }
  f->n_collectives++;
  QMCL_TRY(out.ensure(my_count));
  QMCL_TRY(f->src_index.ensure(my_count + 1));
  out.n = my_count;
  if (my_count) {
    scatter_slots_kernel<<<div_up(my_count, 256), 256, 0, s>>>(
        reinterpret_cast<const SlotRec *>(f->xchg_recv.ptr), my_count, my_first, out.x.ptr,
        out.y.ptr, out.t.ptr, out.w.ptr, f->src_index.ptr);
    QMCL_LAUNCHED();
  }
  f->n_src_index = my_count;
  f->global_first = my_first;
  f->global_n = n_draws;
  *ok = true;
  return QMCL_OK;
}

// particle_filter.cpp:241-291, sharded.
static qmcl_status resample_adaptive_sharded(qmcl_filter *f, const ParticleSet &in,
                                             ParticleSet &out) {
  QMCL_TRY(settle(f));  // the sensor total of the last update feeds w_slow / w_fast
  const double w_diff = adaptiveness(f);
  const size_t n_out = static_cast<size_t>(f->params.particle_count_min);
  bool ok = false;
  f->n_src_index = 0;
  QMCL_TRY(resample_draws_sharded(f, in, out, n_out, w_diff, &ok));
  if (!ok) {
    // failure path of the reference: the output buffer is resized, not filled;
    // here every slot of the block is a default particle
    size_t a, c;
    shard_slice_of(n_out, f->shard_rank, f->n_shards, &a, &c);
    QMCL_TRY(out.ensure(c));
    if (c) {
      fill_default_kernel<<<div_up(c, 256), 256, 0, f->map->stream>>>(out.x.ptr, out.y.ptr,
                                                                     out.t.ptr, out.w.ptr, 0, c);
      QMCL_LAUNCHED();
    }
    out.n = c;
    f->global_first = a;
    f->global_n = n_out;
    return QMCL_OK;
  }
  if (w_diff > 0.0) {
    reset_lowpass(f);
  }
  return QMCL_OK;
}

// Sharded variants of the KLD scan functors: slot index and bin count are global.
struct NewBinFlagAt {
  const float *x, *y, *t;
  KldParams kp;
  BinGrid g;
  const int32_t *table;
  size_t slot_first;
  __device__ uint32_t operator()(size_t i) const {
    return table[kld_cell(x[i], y[i], t[i], kp, g)] == static_cast<int32_t>(slot_first + i) ? 1u : 0u;
  }
};
struct StopRuleAt {
  KldParams kp;
  unsigned long long *first_stop;
  size_t slot_first;
  unsigned long long bins_before;  // new bins opened by the earlier shards
  __device__ void operator()(size_t i, uint32_t before, uint32_t flag) const {
    const unsigned long long k = bins_before + before + flag;
    const size_t m = slot_first + i;
    const unsigned cand = (m >= kld_limit(k, kp)) ? static_cast<unsigned>(m) : 0xffffffffu;
    const unsigned mask = __activemask();
    const unsigned wmin = __reduce_min_sync(mask, cand);
    if (wmin != 0xffffffffu && cand == wmin &&
        static_cast<unsigned long long>(wmin) < *reinterpret_cast<volatile unsigned long long *>(first_stop))
      atomicMin(first_stop, static_cast<unsigned long long>(wmin));
  }
};
__global__ void __launch_bounds__(256)
kld_first_at_kernel(const float *__restrict__ x, const float *__restrict__ y,
                    const float *__restrict__ t, size_t n, size_t slot_first, KldParams kp,
                    BinGrid g, int32_t *__restrict__ table) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const long long cell = kld_cell(x[i], y[i], t[i], kp, g);
  const unsigned peers = __match_any_sync(__activemask(), cell);
  if ((__ffs(peers) - 1) == static_cast<int>(threadIdx.x & 31))
    atomicMin(table + cell, static_cast<int32_t>(slot_first + i));
}

// particle_filter.cpp:293-348, sharded.
static qmcl_status resample_kld_sharded(qmcl_filter *f, const ParticleSet &in, ParticleSet &out,
                                        bool *clustered) {
  cudaStream_t s = f->map->stream;
  const int G = f->n_shards, me = f->shard_rank;
  *clustered = false;
  QMCL_TRY(settle(f));  // the sensor total of the last update feeds w_slow / w_fast
  const double w_diff = adaptiveness(f);
  const size_t n_max = static_cast<size_t>(f->params.particle_count_max);
  f->n_src_index = 0;
  f->n_bins = 0;
  f->bins_pending = false;
  f->n_clusters = 0;
  f->clusters_pending = false;  // an older count still in flight is void
  f->bins_built = true;
  f->grid = BinGrid{0, 0, 0, 0, 0, 0};
  if (n_max > 0x7fffffffull) return QMCL_ERR_CAPACITY;
  bool ok = false;
  if (n_max) QMCL_TRY(resample_draws_sharded(f, in, out, n_max, w_diff, &ok));
  if (!ok) {  // output_set->particles.clear()
    out.n = 0;
    f->global_first = 0;
    f->global_n = 0;
    return QMCL_OK;
  }
  const size_t slot_first = f->global_first, n_loc = out.n;
  // the dense bin grid spanned by all candidates of all shards
  QMCL_TRY(build_bins(f, out.x.ptr, out.y.ptr, out.t.ptr, n_loc, false, true));
  const BinGrid g = f->grid;
  const size_t cells = static_cast<size_t>(g.nx) * g.ny * g.nt;
  KldParams kp;
  kp.res_xy = f->params.res_xy;
  kp.res_t = f->params.res_theta;
  kp.eps2 = 2 * f->params.kld_epsilon;
  kp.z = f->params.kld_z;
  kp.nmin = f->params.particle_count_min;
  kp.nmax = f->params.particle_count_max;
  const unsigned kld_blocks = scan_blocks(std::max<size_t>(n_loc, 1));
  QMCL_TRY(f->scan_legacy.ensure(kld_blocks + 16));
  unsigned long long *d_total =
      reinterpret_cast<unsigned long long *>(f->scan_legacy.ptr + ((kld_blocks + 1) & ~1u));
  unsigned long long *d_stop = d_total + 1;
  unsigned long long stop = ~0ull;
  {
    StageScope sc(f, QMCL_STAGE_KLD);
    // smallest global draw index per bin: local atomicMin, then min over shards
    QMCL_CUDA(cudaMemsetAsync(f->table.ptr, 0x7f, cells * sizeof(int32_t), s));
    if (n_loc) {
      kld_first_at_kernel<<<div_up(n_loc, 256), 256, 0, s>>>(out.x.ptr, out.y.ptr, out.t.ptr, n_loc,
                                                             slot_first, kp, g, f->table.ptr);
      QMCL_LAUNCHED();
    }
    QMCL_TRY(shard_allreduce(f, f->table.ptr, cells, QMCL_DT_I32, QMCL_OP_MIN));
    // new bins opened by the earlier shards
    const NewBinFlagAt nb{out.x.ptr, out.y.ptr, out.t.ptr, kp, g, f->table.ptr, slot_first};
    const unsigned blocks = scan_blocks(std::max<size_t>(n_loc, 1));
    QMCL_CUDA(cudaMemsetAsync(d_total, 0, sizeof(unsigned long long), s));
    if (n_loc) {
      scan_reduce_kernel<<<blocks, kScanBlock, 0, s>>>(nb, n_loc, f->scan_legacy.ptr);
      QMCL_LAUNCHED();
      scan_block_sums_kernel<0><<<1, 1024, 0, s>>>(f->scan_legacy.ptr, blocks, d_total);
      QMCL_LAUNCHED();
    }
    std::vector<unsigned long long> opened(G);
    QMCL_TRY(shard_gather(f, d_total, true, sizeof(unsigned long long), opened.data()));
    unsigned long long before = 0;
    for (int r = 0; r < me; ++r) before += opened[r];
    QMCL_CUDA(cudaMemsetAsync(d_stop, 0xff, sizeof(unsigned long long), s));
    if (n_loc) {
      scan_apply_kernel<<<blocks, kScanBlock, 0, s>>>(nb, StopRuleAt{kp, d_stop, slot_first, before},
                                                     n_loc, f->scan_legacy.ptr);
      QMCL_LAUNCHED();
    }
    std::vector<unsigned long long> stops(G);
    QMCL_TRY(shard_gather(f, d_stop, true, sizeof(unsigned long long), stops.data()));
    for (unsigned long long v : stops) stop = std::min(stop, v);
  }
  const size_t n_out = (stop >= n_max) ? n_max : static_cast<size_t>(stop) + 1;
  if (w_diff > 0.0) {
    reset_lowpass(f);
  }
  // keep the first n_out slots, then give every rank an equal block again
  std::vector<uint64_t> lo(G), hi(G);
  for (int r = 0; r < G; ++r) {
    size_t a, c;
    shard_slice_of(n_max, r, G, &a, &c);
    lo[r] = std::min<uint64_t>(a, n_out);
    hi[r] = std::min<uint64_t>(a + c, n_out);
  }
  out.n = static_cast<size_t>(hi[me] - lo[me]);
  f->n_src_index = out.n;
  f->global_first = lo[me];
  f->global_n = n_out;
  QMCL_TRY(rebalance_blocks(f, out, lo, hi, n_out));
  // the bins the reference holds at this point are those of the kept draws
  QMCL_TRY(build_bins(f, out.x.ptr, out.y.ptr, out.t.ptr, out.n, true));
  QMCL_TRY(label_clusters(f));
  *clustered = true;
  return QMCL_OK;
}

}  // namespace qmcl

using namespace qmcl;

extern "C" {

static qmcl_status resample_and_cluster_impl(qmcl_filter *f);

// particle_filter.cpp:138-161.  The reference has no failing path here; this
// one has (capacity of the dense bin table, injection without free cells, CUDA
// or exchange errors), so the call is a transaction: the new set replaces the
// current one only when every step succeeded.  After an error the filter still
// holds the particles it was called with (their clustering is void).
qmcl_status qmcl_filter_resample_and_cluster(qmcl_filter *f) {
  if (!f) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(use_device(f->map));
  if (f->tail_pending) QMCL_TRY(settle(f));  // an earlier fused call whose outcome was never read
  const int current = f->current;
  const uint32_t step = f->step;
  const uint64_t pose_version = f->pose_version;
  const size_t global_first = f->global_first, global_n = f->global_n;
  f->lowpass_was_reset = false;
  const qmcl_status st = resample_and_cluster_impl(f);
  if (st == QMCL_OK) return st;
  // whatever was enqueued no longer matters; the sensor total of the update
  // before this call still counts (w_slow, w_fast)
  cudaStreamSynchronize(f->map->stream);
  settle_after_sync(f);
  if (f->lowpass_was_reset) {  // the injection round's reset (particle_filter.cpp:284-290) is undone
    f->w_slow = f->w_slow_before_reset;
    f->w_fast = f->w_fast_before_reset;
  }
  f->current = current;
  f->step = step;
  f->pose_version = pose_version + 1;  // bin bounds taken earlier are not trusted any more
  f->global_first = global_first;
  f->global_n = global_n;
  f->n_src_index = 0;
  f->n_bins = 0;
  f->n_clusters = 0;
  f->bins_built = false;
  f->clusters_valid = false;
  f->bbox_hint_valid = false;
  f->grid = BinGrid{0, 0, 0, 0, 0, 0};
  return st;
}

static qmcl_status resample_and_cluster_impl(qmcl_filter *f) {
  // one fused launch for the whole call (and the estimate) when it applies
  if (f->params.resample_type >= 0 && f->params.resample_type <= 2 &&
      tail_eligible(f, f->params.resample_type))
    return launch_tail(f, f->params.resample_type);
  QMCL_TRY(flush_motion(f));
  QMCL_TRY(flush_sensor(f));
  const ParticleSet &in = f->sets[f->current];
  ParticleSet &out = f->sets[(f->current + 1) % 2];
  f->current = (f->current + 1) % 2;
  f->clusters_valid = false;
  f->bbox_hint_valid = false;
  bool clustered = false;
  switch (f->params.resample_type) {
  case QMCL_RESAMPLE_LOW_VARIANCE:
    if (sharded(f)) QMCL_TRY(resample_low_variance_sharded(f, in, out));
    else QMCL_TRY(resample_low_variance(f, in, out));
    break;
  case QMCL_RESAMPLE_ADAPTIVE:
    if (sharded(f)) QMCL_TRY(resample_adaptive_sharded(f, in, out));
    else QMCL_TRY(resample_adaptive(f, in, out));
    break;
  case QMCL_RESAMPLE_KLD:
    if (sharded(f)) QMCL_TRY(resample_kld_sharded(f, in, out, &clustered));
    else QMCL_TRY(resample_kld(f, in, out, &clustered));
    break;
  default:
    return QMCL_ERR_INVALID_ARG;
  }
  f->step++;
  f->pose_version++;
  if (!sharded(f)) {
    f->global_first = 0;
    f->global_n = out.n;
  }
  if (!clustered) {
    if (f->params.resample_type == QMCL_RESAMPLE_KLD) {
      // failed KLD: bins stay reset, nothing to label
      f->clusters_valid = true;
    } else {
      const int *known_bbox = nullptr;
      if (f->bbox_hint_valid) {
        const double cells = (static_cast<double>(f->bbox_hint[3]) - f->bbox_hint[0] + 1) *
                             (static_cast<double>(f->bbox_hint[4]) - f->bbox_hint[1] + 1) *
                             (static_cast<double>(f->bbox_hint[5]) - f->bbox_hint[2] + 1);
        if (cells <= static_cast<double>(kMaxTableCells)) known_bbox = f->bbox_hint;
      }
      f->bbox_hint_valid = false;
      QMCL_TRY(build_bins(f, out.x.ptr, out.y.ptr, out.t.ptr, out.n, false, false, known_bbox));
      QMCL_TRY(label_clusters(f));
    }
  }
  return normalise_current(f);
}

// effective_particles (particle.cpp:32-39): 1 / sum of squared weights.  One
// launch; the sum is a fixed tree over the particle index (deterministic), so it
// differs from the reference's sequential sum by rounding only.  An empty set
// gives 1/0 = inf, as the reference's loop does.
qmcl_status qmcl_filter_effective_particles(const qmcl_filter *fc, double *out) {
  if (!fc || !out) return QMCL_ERR_INVALID_ARG;
  qmcl_filter *f = const_cast<qmcl_filter *>(fc);
  QMCL_TRY(use_device(f->map));
  QMCL_TRY(settle(f));
  cudaStream_t s = f->map->stream;
  const ParticleSet &ps = f->cur();
  double sum = 0.0;
  if (ps.n) {
    const unsigned blocks = static_cast<unsigned>(
        std::min<size_t>(std::max<size_t>(div_up(ps.n, 256 * 4), 1), 4 * kSMs));
    QMCL_TRY(f->partials.ensure(blocks));
    weight_sum_kernel<true><<<blocks, 256, 0, s>>>(ps.w.ptr, ps.n, f->partials.ptr, f->scalars.ptr);
    QMCL_LAUNCHED();
  } else if (sharded(f)) {
    QMCL_CUDA(cudaMemsetAsync(&f->scalars.ptr->sum_sq, 0, sizeof(double), s));
  }
  if (sharded(f)) {
    std::vector<double> sums;
    QMCL_TRY(shard_gather_f64(f, &f->scalars.ptr->sum_sq, &sums));
    sum = ordered_sum(sums);
  } else if (ps.n) {
    QMCL_CUDA(copy_d2h(&sum, &f->scalars.ptr->sum_sq, sizeof(double), s));
    QMCL_CUDA(cudaStreamSynchronize(s));
  }
  *out = 1.0 / sum;
  return QMCL_OK;
}

qmcl_status qmcl_filter_copy_resample_indices(const qmcl_filter *fc, int64_t *out, size_t cap,
                                              size_t *n_out) {
  if (!fc) return QMCL_ERR_INVALID_ARG;
  qmcl_filter *f = const_cast<qmcl_filter *>(fc);
  QMCL_TRY(settle(f));
  if (n_out) *n_out = f->n_src_index;
  const size_t n = f->n_src_index < cap ? f->n_src_index : cap;
  if (!n || !out) return QMCL_OK;
  QMCL_TRY(use_device(f->map));
  QMCL_CUDA(copy_d2h(out, f->src_index.ptr, n * sizeof(int64_t), f->map->stream));
  QMCL_CUDA(cudaStreamSynchronize(f->map->stream));
  return QMCL_OK;
}

}  // extern "C"
