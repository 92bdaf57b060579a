// tail.cu — the fused tail kernel (see tail.cuh) and its host side.
//
// Replaces, in one launch, the chain that follows the sensor kernel in
// ParticleFilter::update_importance_from_observations, resample_and_cluster,
// cluster and get_estimated_pose (src/quickmcl/particle_filter.cpp:101-120,
// :138-348, :163-212, :366-391; particle.cpp:41-82; space_partitioning.cpp:39-128;
// statistics.cpp:23-79).

#include "tail.cuh"

#include "cluster.cuh"
#include "prefix.cuh"
#include "scan.cuh"
#include "shard.cuh"

#include <cooperative_groups.h>

#include <cstdlib>

namespace cg = cooperative_groups;

#ifndef QMCL_TAIL_CG_SYNC
#define QMCL_TAIL_CG_SYNC 0  // 1: cooperative_groups grid.sync() instead of the backing-off barrier
#endif

namespace qmcl {

constexpr size_t kSoloCells = 2048;  // dense grids up to this size are listed and labelled by one block

// ------------------------------------------------------------------ teams
// Grid barrier of the cooperative launch.  cooperative_groups' grid.sync() polls
// its flag without pause; with 295 blocks waiting for one busy warp (the exact
// carry walk, the single-block labelling) that polling alone slowed the worker
// down 2-3x (measured: resolve phase 36 us stand-alone, 61 us with 147 waiting
// blocks, 94 us with 295).  This barrier backs off with nanosleep between polls.
// The counter only ever grows; every block passes the same number of barriers,
// so a launch leaves it at a multiple of the grid size.
struct GridTeam {
  unsigned long long *counter;
  __device__ explicit GridTeam(const TailJob &J) : counter(&J.dev->barrier) {}
  __device__ static unsigned rank() { return blockIdx.x; }
  __device__ static unsigned size() { return gridDim.x; }
  // max_sleep_ns: the longest pause between two polls; phases in which one block
  // works for tens of microseconds pass a larger one.
  __device__ void sync(unsigned max_sleep_ns = 320) const {
#if QMCL_TAIL_CG_SYNC
    cg::this_grid().sync();
#else
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned long long n = gridDim.x;
      const unsigned long long old = atomicAdd(counter, 1ull);
      const unsigned long long target = (old / n + 1) * n;
      unsigned ns = 20;
      while (*reinterpret_cast<volatile unsigned long long *>(counter) < target) {
        __nanosleep(ns);
        if (ns < max_sleep_ns) ns *= 2;
      }
      __threadfence();
    }
    __syncthreads();
#endif
  }
};
struct BlockTeam {  // one block per job
  __device__ explicit BlockTeam(const TailJob &) {}
  __device__ static unsigned rank() { return 0; }
  __device__ static unsigned size() { return 1; }
  __device__ void sync(unsigned = 0) const { __syncthreads(); }
};

// ---------------------------------------------------------- block helpers
// Sum over the 256 threads of a block; every thread receives it.  Any order
// will do for the callers (integers, or doubles of which only the exponent is
// used).
__device__ __forceinline__ double block_sum_f64(double v) {
  __shared__ double s[8];
  __shared__ double s_out;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s[i];
    s_out = t;
  }
  __syncthreads();
  const double r = s_out;
  __syncthreads();
  return r;
}
__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v) {
  __shared__ unsigned long long s[8];
  __shared__ unsigned long long s_out;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += s[i];
    s_out = t;
  }
  __syncthreads();
  const unsigned long long r = s_out;
  __syncthreads();
  return r;
}
// Sum of cnt[0 .. upto) (every thread receives it).
__device__ __forceinline__ unsigned long long block_prefix_of(const uint32_t *cnt, size_t upto) {
  unsigned long long c = 0;
  for (size_t j = threadIdx.x; j < upto; j += kTailThreads) c += cnt[j];
  return block_sum_u64(c);
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// Block 0 stamps the end of a phase (after the team barrier that closes it).
#define TAIL_STAMP(k)                                              \
  do {                                                             \
    if (rank == 0 && tid == 0) D->stamps[k] = global_timer_ns();   \
  } while (0)

// low_pass_filter.h:39-50
__device__ __forceinline__ void lowpass_update_dev(double *value, double alpha, double x) {
  if (*value == 0) *value = x;
  else *value = __dadd_rn(*value, __dmul_rn(alpha, __dadd_rn(x, -*value)));
}

// ------------------------------------------------------------------ phases
// A device-wide compaction over n items in tiles of 2048 (256 threads x 8
// contiguous items, the geometry of scan.cuh): phase A counts per tile, phase B
// emits g(i, exclusive_prefix, flag).  Between the two the caller synchronises
// the team.  Returns (phase B) the grand total to every thread.
// Items per thread of a compaction over n items: as few as keeps every block of
// the team busy (each item is a chain of dependent loads), at most 8.
template <class Team>
__device__ __forceinline__ int compact_ipt(size_t n) {
  const size_t per = static_cast<size_t>(Team::size()) * kScanBlock;
  const size_t want = (n + per - 1) / per;
  return static_cast<int>(want < 1 ? 1 : (want > kScanItems ? kScanItems : want));
}
template <class Team, typename F>
__device__ __forceinline__ void compact_count(F f, size_t n, uint32_t *tile_cnt, int ipt = kScanItems) {
  const size_t tile_items = static_cast<size_t>(kScanBlock) * ipt;
  const size_t tiles = (n + tile_items - 1) / tile_items;
  for (size_t tile = Team::rank(); tile < tiles; tile += Team::size()) {
    const size_t base = (tile * kScanBlock + threadIdx.x) * ipt;
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
      if (i < ipt && base + i < n) c += f(base + i);
    uint32_t total;
    block_excl_scan_u32(c, &total);
    if (threadIdx.x == 0) tile_cnt[tile] = total;
  }
}
template <typename F, typename G>
__device__ __forceinline__ void compact_emit_tile(F f, G g, size_t n, const uint32_t *tile_cnt, size_t tile,
                                                  int ipt = kScanItems) {
  const size_t base = (tile * kScanBlock + threadIdx.x) * ipt;
  const unsigned long long excl = block_prefix_of(tile_cnt, tile);
  uint32_t v[kScanItems];
  uint32_t c = 0;
#pragma unroll
  for (int i = 0; i < kScanItems; ++i) {
    v[i] = (i < ipt && base + i < n) ? f(base + i) : 0u;
    c += v[i];
  }
  uint32_t total;
  uint32_t at = static_cast<uint32_t>(excl) + block_excl_scan_u32(c, &total);
#pragma unroll
  for (int i = 0; i < kScanItems; ++i)
    if (i < ipt && base + i < n) {
      g(base + i, at, v[i]);
      at += v[i];
    }
}
template <class Team, typename F, typename G>
__device__ __forceinline__ unsigned long long compact_emit(F f, G g, size_t n, const uint32_t *tile_cnt,
                                                           int ipt = kScanItems) {
  const size_t tile_items = static_cast<size_t>(kScanBlock) * ipt;
  const size_t tiles = (n + tile_items - 1) / tile_items;
  for (size_t tile = Team::rank(); tile < tiles; tile += Team::size())
    compact_emit_tile(f, g, n, tile_cnt, tile, ipt);
  return block_prefix_of(tile_cnt, tiles);
}

// Union-find helpers on a parent array in SHARED memory (the global-memory
// versions in cluster_dev.cuh read through the L2 with ld.cg).
__device__ __forceinline__ int uf_root_plain(const int32_t *parent, int v) {
  int p = parent[v];
  while (p != v) {
    v = p;
    p = parent[v];
  }
  return v;
}

struct TailPositive {
  const double *w;
  __device__ uint32_t operator()(size_t i) const { return w[i] > 0 ? 1u : 0u; }
};
struct TailEmitPositive {
  const double *cdf;
  int64_t *pos;
  double *incl;
  __device__ void operator()(size_t i, uint32_t t, uint32_t flag) const {
    if (!flag) return;
    pos[t] = static_cast<int64_t>(i);
    incl[t] = cdf[i];
  }
};
struct TailNewBin {
  const float *x, *y, *t;
  KldParams kp;
  BinGrid g;
  const int32_t *first;
  __device__ uint32_t operator()(size_t m) const {
    return first[kld_cell(x[m], y[m], t[m], kp, g)] == static_cast<int32_t>(m) ? 1u : 0u;
  }
};
// particle_filter.cpp:329-334: stop at the first m with m >= limit(bins after m).
struct TailStopRule {
  KldParams kp;
  unsigned long long *stop;
  __device__ void operator()(size_t m, uint32_t before, uint32_t flag) const {
    const unsigned long long k = static_cast<unsigned long long>(before) + flag;
    const unsigned cand = (m >= kld_limit(k, kp)) ? static_cast<unsigned>(m) : 0xffffffffu;
    const unsigned mask = __activemask();
    const unsigned wmin = __reduce_min_sync(mask, cand);
    if (wmin != 0xffffffffu && cand == wmin) {
      const unsigned long long word = (static_cast<unsigned long long>(wmin) << 32) | (k & 0xffffffffull);
      if (word < *reinterpret_cast<volatile unsigned long long *>(stop)) atomicMin(stop, word);
    }
  }
};
struct TailOccupied {
  const int32_t *table;
  __device__ uint32_t operator()(size_t i) const { return table[i] > 0 ? 1u : 0u; }
};
struct TailEmitBin {
  int32_t *table, *bin_cell, *bin_count, *parent;
  __device__ void operator()(size_t i, uint32_t at, uint32_t flag) const {
    if (!flag) return;
    bin_cell[at] = static_cast<int32_t>(i);
    bin_count[at] = table[i];
    parent[at] = static_cast<int32_t>(at);
    table[i] = static_cast<int32_t>(at) + 1;
  }
};
struct TailRoot {
  const int32_t *label;
  __device__ uint32_t operator()(size_t b) const { return label[b] == static_cast<int32_t>(b) ? 1u : 0u; }
};
struct TailEmitRoot {
  int32_t *cluster_of_root;
  unsigned long long *cluster_w;
  __device__ void operator()(size_t b, uint32_t at, uint32_t flag) const {
    if (!flag) return;
    cluster_of_root[b] = static_cast<int32_t>(at);
    cluster_w[at] = 0ull;  // the fixed-point cluster weights start from zero
  }
};

template <class Team>
__device__ void tail_finish(const TailJob &J, unsigned long long status, double total, double w_diff,
                            const BinGrid &g) {
  if (Team::rank() != 0 || threadIdx.x != 0) return;
  TailResult *r = J.result;
  const TailDev *D = J.dev;
  r->status = status;
  r->n_out = D->n_out;
  r->n_bins = D->n_bins;
  r->n_clusters = D->n_clusters;
  r->n_records = D->n_records;
  r->invalid = static_cast<unsigned long long>(D->es.invalid);
  r->grid[0] = g.x0;
  r->grid[1] = g.y0;
  r->grid[2] = g.t0;
  r->grid[3] = g.nx;
  r->grid[4] = g.ny;
  r->grid[5] = g.nt;
  r->total_weight = total;
  r->w_diff = w_diff;
  for (int k = 0; k < 13; ++k) r->est[k] = D->out[k];
  for (int k = 0; k < kTailStamps; ++k) r->stamps[k] = D->stamps[k];
  r->stamps[TS_END] = global_timer_ns();
  __threadfence_system();
  r->serial = J.serial;
  __threadfence_system();
}

template <class Team>
__device__ void run_tail(const TailJob &J) {
  using namespace pfx;
  // block 0's scratch for the single-block phases (KLD stop rule, bin list + labelling)
  __shared__ __align__(16) int32_t s_solo[3 * kSoloCells];
  const Team team(J);
  const unsigned rank = Team::rank(), size = Team::size();
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  TailDev *D = J.dev;
  const size_t n_in = static_cast<size_t>(J.n_in);
  const bool resample = J.mode != TAIL_CLUSTER;
  const bool draws = J.mode == TAIL_ADAPTIVE || J.mode == TAIL_KLD;

  // ---- scalars every thread derives alike: sensor total, low-pass filters
  // (particle_filter.cpp:101-120), injection probability (:246), bin grid
  double total = 0.0, w_slow = J.w_slow, w_fast = J.w_fast;
  if (J.normalise_pending) {
    total = J.sc->total_weight;
    if (total > 0) {
      const double avg = __ddiv_rn(total, static_cast<double>(n_in));
      lowpass_update_dev(&w_slow, J.alpha_slow, avg);
      lowpass_update_dev(&w_fast, J.alpha_fast, avg);
    }
  }
  double w_diff = 0.0;
  if (draws) {
    const double d = __dadd_rn(1.0, -__ddiv_rn(w_fast, w_slow));
    w_diff = (0.0 < d) ? d : 0.0;  // std::max(0.0, NaN) is 0.0
  }
  int bb[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) bb[k] = J.bbox[k];
  if (w_diff > 0.0) {  // injected poses are free-cell corners: inside the map's extent
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      bb[k] = min(bb[k], J.map_bbox[k]);
      bb[3 + k] = max(bb[3 + k], J.map_bbox[3 + k]);
    }
  }
  BinGrid g{0, 0, 0, 0, 0, 0};
  const long long nx = static_cast<long long>(bb[3]) - bb[0] + 1;
  const long long ny = static_cast<long long>(bb[4]) - bb[1] + 1;
  const long long nt = static_cast<long long>(bb[5]) - bb[2] + 1;
  const double cells_d = static_cast<double>(nx) * static_cast<double>(ny) * static_cast<double>(nt);
  const bool grid_ok = nx > 0 && ny > 0 && nt > 0 && cells_d <= static_cast<double>(J.table_cap);
  size_t cells = 0;
  if (grid_ok) {
    g = BinGrid{bb[0], bb[1], bb[2], static_cast<int>(nx), static_cast<int>(ny), static_cast<int>(nt)};
    cells = static_cast<size_t>(nx) * ny * nt;
  }
  if (rank == 0 && tid == 0) {
    D->stop = ~0ull;
    D->n_new = 0;
    D->n_out = 0;
    D->n_bins = 0;
    D->n_clusters = 0;
    D->n_records = 0;
    D->wmax_raw = 0;
    D->slot_counter = 0;
    D->es.wmax_bits = 0;
    D->es.best_w = 0;
    D->es.best = -1;
    D->es.invalid = 0;
    D->es.blocks_done = 0;
    for (int k = 0; k < 16; ++k) D->out[k] = 0.0;
    for (int k = 0; k < kTailStamps; ++k) D->stamps[k] = 0;
    D->stamps[TS_START] = global_timer_ns();
  }
  // Conditions under which the multi-launch path has to take over; all known
  // before anything is touched.
  if (!grid_ok || (w_diff > 0.0 && J.n_free == 0)) {
    team.sync();  // D is initialised before tail_finish reads it
    tail_finish<Team>(J, TAIL_FALLBACK, total, w_diff, g);
    return;
  }

  // ---- phase 0: K8 normalise (particle.cpp:50-64), chunk sums, positive counts; clear the tables
  const double inv_n = __ddiv_rn(1.0, static_cast<double>(n_in));
  if (resample) {
    for (size_t tile = rank; tile < J.nt; tile += size) {
      const size_t chunk = tile * kTileChunks + wid;
      const size_t base = chunk * kChunk + static_cast<size_t>(lane) * kLaneItems;
      double s = 0.0;
      unsigned npos = 0;
      if (chunk * kChunk < n_in) {
        double v[kLaneItems];
        load_items(J.iw, base, n_in, v);
        if (J.normalise_pending) {
#pragma unroll
          for (int i = 0; i < kLaneItems; ++i)
            if (base + i < n_in) {
              v[i] = (total > 0) ? __ddiv_rn(v[i], total) : inv_n;
              J.iw[base + i] = v[i];
            }
        }
#pragma unroll
        for (int i = 0; i < kLaneItems; ++i) {
          s += v[i];
          npos += (base + i < n_in && v[i] > 0) ? 1u : 0u;
        }
        s = warp_sum_xor(s);
        if (lane == 0) J.pb.chunk_sum[chunk] = s;
      }
      const double ts = block_sum_f64(lane == 0 ? s : 0.0);
      const unsigned long long tp = block_sum_u64(npos);
      if (tid == 0) {
        J.tile_sum[tile] = ts;
        J.tile_pos[tile] = static_cast<uint32_t>(tp);
      }
    }
  } else if (J.normalise_pending) {
    for (size_t i = static_cast<size_t>(rank) * kTailThreads + tid; i < n_in;
         i += static_cast<size_t>(size) * kTailThreads)
      J.iw[i] = (total > 0) ? __ddiv_rn(J.iw[i], total) : inv_n;
  }
  for (size_t i = static_cast<size_t>(rank) * kTailThreads + tid; i < cells;
       i += static_cast<size_t>(size) * kTailThreads) {
    J.table[i] = 0;
    if (J.mode == TAIL_KLD) J.first[i] = 0x7f7f7f7f;
  }
  team.sync();
  TAIL_STAMP(TS_NORMALISE);

  unsigned long long n_out = n_in;  // cluster-only: the current set
  if (resample) {
    // ---- phase 1: binade guesses (approximate carry-in per chunk) and summaries
    for (size_t tile = rank; tile < J.nt; tile += size) {
      double c = 0.0;
      for (size_t j = tid; j < tile; j += kTailThreads) c += J.tile_sum[j];
      const double before = block_sum_f64(c);
      if (tid == 0) {
        double run = before;
        for (int k = 0; k < kTileChunks; ++k) {
          const size_t ch = tile * kTileChunks + k;
          if (ch >= J.nc) break;
          J.pb.approx_in[ch] = run;
          run += J.pb.chunk_sum[ch];
        }
      }
      __syncthreads();
      prefix_summaries_tile(J.iw, n_in, J.nc, J.pb, tile);
      __syncthreads();
    }
    team.sync();
    TAIL_STAMP(TS_SUMMARIES);
    // ---- phase 2: the exact carry through all tiles (one warp)
    static_assert(sizeof(ResolveScratch) <= sizeof(s_solo), "the resolve scratch is lent from s_solo");
    if (rank == 0 && wid == 0)
      prefix_resolve_warp(J.iw, n_in, J.nc, J.nt, 0.0, J.pb, J.cdf, lane, reinterpret_cast<ResolveScratch *>(s_solo));
    team.sync(2560);
    TAIL_STAMP(TS_RESOLVE);
    // ---- phase 3: expand into the running sums; records of the positive weights
    // (a tile's records read only that tile's running sums, which this block has just written)
    for (size_t tile = rank; tile < J.nt; tile += size) {
      prefix_apply_tile(J.iw, n_in, J.nc, J.pb, J.cdf, tile);
      if (draws) {
        __syncthreads();
        compact_emit_tile(TailPositive{J.iw}, TailEmitPositive{J.cdf, J.pos, J.incl}, n_in, J.tile_pos, tile);
      }
    }
    if (draws && rank == 0) {
      const unsigned long long P = block_prefix_of(J.tile_pos, J.nt);
      if (tid == 0) D->n_records = P;
    }
    team.sync();
    TAIL_STAMP(TS_APPLY);

    const double cdf_total = J.cdf[n_in - 1];
    if (draws && !(cdf_total > 0)) {  // "No particles have any weights! Failing to resample"
      if (J.mode == TAIL_ADAPTIVE) {
        // the reference leaves a resized, unfilled buffer behind: the multi-launch path restates that
        tail_finish<Team>(J, J.normalise_pending ? TAIL_FALLBACK_NORMALISED : TAIL_FALLBACK, total,
                          w_diff, g);
      } else {
        tail_finish<Team>(J, TAIL_NO_WEIGHT, total, w_diff, BinGrid{0, 0, 0, 0, 0, 0});
      }
      return;
    }

    // ---- phase 4: K10 / K12 selection and gather
    if (J.mode == TAIL_LOW_VARIANCE) {
      // particle_filter.cpp:214-239
      const size_t M = n_in;
      const double minv = static_cast<double>(1.0f / static_cast<float>(M));
      uint32_t r4[4];
      rng_draw(J.seed, J.step, P_LV_OFFSET, 0, r4);
      const double r = __dmul_rn(uniform53(r4[0], r4[1]), minv);
      for (size_t m = static_cast<size_t>(rank) * kTailThreads + tid; m < M;
           m += static_cast<size_t>(size) * kTailThreads) {
        const double U = __dadd_rn(r, __dmul_rn(static_cast<double>(m), minv));
        size_t lo = 0, hi = M;
        while (lo < hi) {
          const size_t mid = (lo + hi) >> 1;
          if (U > J.cdf[mid]) lo = mid + 1;
          else hi = mid;
        }
        if (lo > M - 1) lo = M - 1;
        J.src[m] = static_cast<int64_t>(lo);
        J.ox[m] = J.ix[lo];
        J.oy[m] = J.iy[lo];
        J.ot[m] = J.it[lo];
        J.ow[m] = J.iw[lo];
      }
      n_out = M;
    } else {
      // particle_filter.cpp:261-276, :313-327; KLD: smallest draw index per bin.
      // KLD looks at the candidates in (at most) two rounds: a probe of the first
      // J.kld_probe draws, which holds the stop in all but the first cycles of a
      // tracking filter, then all particle_count_max of them.  Draw m and the
      // stop test at m depend on draws <= m only, so the outcome is the same.
      const size_t n_draws = static_cast<size_t>(J.mode == TAIL_KLD ? J.n_max : J.n_min);
      const size_t P = static_cast<size_t>(D->n_records);
      // every 2^k-th key of the record list in shared memory: the first ten levels of each search
      __shared__ double s_samp[kLookupSamples];
      size_t samp_stride = 1, n_samp = 0;
      if (P > 1) {
        samp_stride = (P - 1 + kLookupSamples - 1) / kLookupSamples;
        n_samp = (P - 1 + samp_stride - 1) / samp_stride;
        for (size_t j = tid; j < n_samp; j += kTailThreads) s_samp[j] = J.incl[j * samp_stride];
      }
      __syncthreads();
      size_t done = 0;
      size_t round_end = n_draws;
      if (J.mode == TAIL_KLD && J.kld_probe > 0 && J.kld_probe < n_draws) round_end = static_cast<size_t>(J.kld_probe);
      while (true) {
        const size_t first_m = done & ~static_cast<size_t>(31);
        const size_t rounded = (round_end + 31) & ~static_cast<size_t>(31);
        for (size_t m = first_m + static_cast<size_t>(rank) * kTailThreads + tid; m < rounded;
             m += static_cast<size_t>(size) * kTailThreads) {
          long long cell = -1 - static_cast<long long>(lane);  // lanes outside the round: distinct, never marked
          if (m >= done && m < round_end) {
            uint32_t rr4[4];
            rng_draw(J.seed, J.step, P_DRAW, m, rr4);
            float x, y, t;
            double w;
            int64_t sidx;
            if (uniform53(rr4[0], rr4[1]) < w_diff) {
              uint32_t q[4];
              rng_draw(J.seed, J.step, P_INJECT, m, q);
              random_free_pose(J.free_cells, J.n_free, J.geom, q, &x, &y, &t);
              w = 0.0;  // no influence on the estimate until the next update
              sidx = -1;
            } else {
              sidx = J.pos[running_sum_lookup_sampled(J.incl, P, uniform53(rr4[2], rr4[3]), s_samp, n_samp,
                                                      samp_stride)];
              x = J.ix[sidx];
              y = J.iy[sidx];
              t = J.it[sidx];
              w = J.iw[sidx];
            }
            J.ox[m] = x;
            J.oy[m] = y;
            J.ot[m] = t;
            J.ow[m] = w;
            J.src[m] = sidx;
            if (J.mode == TAIL_KLD) cell = kld_cell(x, y, t, J.kp, g);
          }
          if (J.mode == TAIL_KLD) {
            const unsigned peers = __match_any_sync(0xffffffffu, cell);
            if (cell >= 0 && (__ffs(peers) - 1) == lane) atomicMin(J.first + cell, static_cast<int32_t>(m));
          }
        }
        n_out = n_draws;
        team.sync();
        if (J.mode != TAIL_KLD) break;
        if (done == 0) TAIL_STAMP(TS_SELECT);
        // ---- phases 5, 6: K11 KLD stop rule (particle_filter.cpp:313-335) over [0, round_end)
        if (cells <= J.solo_cells) {
          // A small table (a tracking cloud): the stop follows from the table alone.  With the
          // first-draw indices of the occupied cells sorted, f_1 < f_2 < ... < f_K, the number of
          // bins after draw m is k(m) = #{j : f_j <= m}; on [f_j, f_j+1) it is j, so the first m
          // of that stretch with m >= limit(j) is max(f_j, limit(j)), and the stop is the smallest
          // such m over all j.  One block, K values: no pass over the candidates at all.
          if (rank == 0) {
            int32_t *s_first = s_solo;
            constexpr int kPer = kSoloCells / kTailThreads;
            int32_t fv[kPer];
            uint32_t c = 0, total = 0;
#pragma unroll
            for (int i = 0; i < kPer; ++i) {
              const size_t idx = static_cast<size_t>(tid) * kPer + i;
              fv[i] = idx < cells ? J.first[idx] : 0x7f7f7f7f;
              c += (fv[i] != 0x7f7f7f7f) ? 1u : 0u;
            }
            uint32_t at = block_excl_scan_u32(c, &total);
            const uint32_t K = total;
            uint32_t n2 = 32;
            while (n2 < K) n2 <<= 1;
            for (uint32_t i = tid; i < n2; i += kTailThreads) s_first[i] = 0x7fffffff;
            __syncthreads();
#pragma unroll
            for (int i = 0; i < kPer; ++i)
              if (fv[i] != 0x7f7f7f7f) s_first[at++] = fv[i];
            __syncthreads();
            for (uint32_t k = 2; k <= n2; k <<= 1)
              for (uint32_t j = k >> 1; j > 0; j >>= 1) {
                for (uint32_t i = tid; i < n2; i += kTailThreads) {
                  const uint32_t l = i ^ j;
                  if (l > i) {
                    const int32_t a = s_first[i], b = s_first[l];
                    if ((a > b) == ((i & k) == 0)) {
                      s_first[i] = b;
                      s_first[l] = a;
                    }
                  }
                }
                __syncthreads();
              }
            unsigned long long best = ~0ull;
            for (uint32_t j = tid; j < K; j += kTailThreads) {
              const unsigned long long fj = static_cast<unsigned long long>(s_first[j]);
              const unsigned long long upper =
                  (j + 1 < K) ? static_cast<unsigned long long>(s_first[j + 1]) : round_end;
              const unsigned long long lim = kld_limit(j + 1, J.kp);
              const unsigned long long m = fj > lim ? fj : lim;
              if (m < upper) best = min(best, (m << 32) | (j + 1));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
            if (lane == 0 && best != ~0ull) atomicMin(&D->stop, best);
            if (tid == 0) D->n_new = K;
          }
          team.sync(640);
        } else {
          const TailNewBin nb{J.ox, J.oy, J.ot, J.kp, g, J.first};
          const int ipt = compact_ipt<Team>(round_end);
          compact_count<Team>(nb, round_end, J.tile_cnt, ipt);
          team.sync();
          const unsigned long long n_new =
              compact_emit<Team>(nb, TailStopRule{J.kp, &D->stop}, round_end, J.tile_cnt, ipt);
          if (rank == 0 && tid == 0) D->n_new = n_new;
          team.sync();
        }
        const unsigned long long stop = *reinterpret_cast<volatile unsigned long long *>(&D->stop);
        if (stop != ~0ull) {
          n_out = (stop >> 32) + 1;
          break;
        }
        if (round_end == n_draws) break;  // no stop at all: every candidate is kept
        done = round_end;
        round_end = n_draws;
      }
      if (J.mode == TAIL_KLD) TAIL_STAMP(TS_KLD);
      else TAIL_STAMP(TS_SELECT);
    }
    if (J.mode == TAIL_LOW_VARIANCE) {
      team.sync();
      TAIL_STAMP(TS_SELECT);
    }
  }
  if (rank == 0 && tid == 0) D->n_out = n_out;

  // ---- phase 7: SpacePartitioning::add_point for the kept particles
  const float *cx = resample ? J.ox : J.ix, *cy = resample ? J.oy : J.iy, *ct = resample ? J.ot : J.it;
  double *cw = resample ? J.ow : J.iw;
  {
    const size_t rounded = (static_cast<size_t>(n_out) + 31) & ~static_cast<size_t>(31);
    for (size_t i = static_cast<size_t>(rank) * kTailThreads + tid; i < rounded;
         i += static_cast<size_t>(size) * kTailThreads)
      bin_mark_one(cx, cy, ct, static_cast<size_t>(n_out), J.bp, g, J.table, i);
  }
  team.sync();
  TAIL_STAMP(TS_MARK);
  // ---- phases 8-11: the occupied-bin list (sorted), table -> bin id + 1; K13 connected
  // components (space_partitioning.cpp:89-128); dense cluster ids.  A small grid (a
  // tracking cloud) is handled by block 0 alone between block barriers, with the
  // union-find forest in shared memory: six team barriers and the latency of
  // contended global atomics on a single component go away.
  unsigned long long n_bins = 0, n_clusters = 0;  // valid on rank 0 (everywhere on the team path)
  if (cells <= J.solo_cells) {
    if (rank == 0) {
      // everything in shared memory: the dense table, the bin list, the union-find forest
      int32_t *s_tab = s_solo, *s_cell = s_solo + kSoloCells, *s_parent = s_solo + 2 * kSoloCells;
      for (size_t i = tid; i < cells; i += kTailThreads) s_tab[i] = J.table[i];
      __syncthreads();
      uint32_t v[kSoloCells / kTailThreads];
      uint32_t c = 0, total = 0;
      constexpr int kPer = kSoloCells / kTailThreads;  // contiguous cells per thread
#pragma unroll
      for (int i = 0; i < kPer; ++i) {
        const size_t idx = static_cast<size_t>(tid) * kPer + i;
        v[i] = (idx < cells && s_tab[idx] > 0) ? 1u : 0u;
        c += v[i];
      }
      uint32_t at = block_excl_scan_u32(c, &total);
      n_bins = total;
#pragma unroll
      for (int i = 0; i < kPer; ++i)
        if (v[i]) {
          const size_t idx = static_cast<size_t>(tid) * kPer + i;
          J.bin_cell[at] = static_cast<int32_t>(idx);
          s_cell[at] = static_cast<int32_t>(idx);
          J.bin_count[at] = s_tab[idx];
          s_parent[at] = static_cast<int32_t>(at);
          s_tab[idx] = static_cast<int32_t>(at) + 1;
          ++at;
        }
      __syncthreads();
      for (size_t b = tid; b < n_bins; b += kTailThreads)
        cc_hook_bin(s_tab, s_cell, g, J.bp, s_parent, static_cast<int>(b));
      __syncthreads();
      int32_t root[kPer];
#pragma unroll
      for (int i = 0; i < kPer; ++i) {
        const size_t b = static_cast<size_t>(tid) + static_cast<size_t>(i) * kTailThreads;
        root[i] = b < n_bins ? uf_root_plain(s_parent, static_cast<int>(b)) : -1;
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kPer; ++i) {
        const size_t b = static_cast<size_t>(tid) + static_cast<size_t>(i) * kTailThreads;
        if (b < n_bins) {
          s_parent[b] = root[i];
          J.bin_label[b] = root[i];
        }
      }
      __syncthreads();
      // dense cluster ids in root order
      c = 0;
#pragma unroll
      for (int i = 0; i < kPer; ++i) {
        const size_t b = static_cast<size_t>(tid) * kPer + i;
        v[i] = (b < n_bins && s_parent[b] == static_cast<int32_t>(b)) ? 1u : 0u;
        c += v[i];
      }
      at = block_excl_scan_u32(c, &total);
      n_clusters = total;
#pragma unroll
      for (int i = 0; i < kPer; ++i)
        if (v[i]) {
          J.bin_cluster[static_cast<size_t>(tid) * kPer + i] = static_cast<int32_t>(at);
          J.cluster_w[at] = 0ull;
          ++at;
        }
      for (size_t i = tid; i < cells; i += kTailThreads) J.table[i] = s_tab[i];
      if (tid == 0) {
        D->n_bins = n_bins;
        D->n_clusters = n_clusters;
      }
    }
    TAIL_STAMP(TS_BINS);
  } else {
    const int ipt_c = compact_ipt<Team>(cells);
    compact_count<Team>(TailOccupied{J.table}, cells, J.tile_cnt, ipt_c);
    team.sync();
    n_bins = compact_emit<Team>(TailOccupied{J.table},
                                TailEmitBin{J.table, J.bin_cell, J.bin_count, J.bin_label}, cells, J.tile_cnt,
                                ipt_c);
    if (rank == 0 && tid == 0) D->n_bins = n_bins;
    team.sync();
    TAIL_STAMP(TS_BINS);
    for (size_t b = static_cast<size_t>(rank) * kTailThreads + tid; b < n_bins;
         b += static_cast<size_t>(size) * kTailThreads)
      cc_hook_bin(J.table, J.bin_cell, g, J.bp, J.bin_label, static_cast<int>(b));
    team.sync();
    for (size_t b = static_cast<size_t>(rank) * kTailThreads + tid; b < n_bins;
         b += static_cast<size_t>(size) * kTailThreads)
      J.bin_label[b] = uf_find_readonly(J.bin_label, static_cast<int>(b));
    team.sync();
    const int ipt_b = compact_ipt<Team>(n_bins);
    compact_count<Team>(TailRoot{J.bin_label}, n_bins, J.tile_cnt, ipt_b);
    team.sync();
    n_clusters = compact_emit<Team>(TailRoot{J.bin_label}, TailEmitRoot{J.bin_cluster, J.cluster_w}, n_bins,
                                    J.tile_cnt, ipt_b);
    if (rank == 0 && tid == 0) D->n_clusters = n_clusters;
  }
  TAIL_STAMP(TS_LABEL);

  // ---- phase 12: total of the new set (particle.cpp:41-48) and its largest weight
  const size_t n = static_cast<size_t>(n_out);
  const unsigned ws_blocks = weight_sum_blocks(n);
  double *mom_partials = J.partials + 4 * 148;
  {
    unsigned long long m = 0;
    if (resample) {
      for (unsigned vb = rank; vb < ws_blocks; vb += size) weight_sum_block<false>(cw, n, J.partials, vb, ws_blocks);
    }
    for (size_t i = static_cast<size_t>(rank) * kTailThreads + tid; i < n;
         i += static_cast<size_t>(size) * kTailThreads) {
      const double v = cw[i];
      if (v > 0) m = max(m, static_cast<unsigned long long>(__double_as_longlong(v)));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0 && m) atomicMax(&D->wmax_raw, m);
  }
  team.sync(cells <= J.solo_cells ? 1280 : 320);  // block 0 may still be labelling on its own
  TAIL_STAMP(TS_WEIGHT_SUM);
  // ---- phase 13: normalise (particle_filter.cpp:160), cluster weights (K14 step 2)
  {
    double post_total = 1.0;
    if (resample) post_total = weight_sum_final(J.partials, ws_blocks);
    unsigned long long wmax_bits = *reinterpret_cast<volatile unsigned long long *>(&D->wmax_raw);
    if (resample && wmax_bits) {
      // x -> fl(x / total) is monotone: the largest normalised weight is the normalised largest weight
      const double q = __ddiv_rn(__longlong_as_double(static_cast<long long>(wmax_bits)), post_total);
      wmax_bits = (q > 0) ? static_cast<unsigned long long>(__double_as_longlong(q)) : 0ull;
    }
    if (rank == 0 && tid == 0) {
      D->es.wmax_bits = wmax_bits;
      J.result->post_total = post_total;
    }
    const int shift = fixed_point_shift(wmax_bits, n);
    const size_t vblocks = (n + kTailThreads - 1) / kTailThreads;
    for (size_t vb = rank; vb < vblocks; vb += size) {
      const size_t i = vb * kTailThreads + tid;
      if (resample && i < n) cw[i] = __ddiv_rn(cw[i], post_total);
      cluster_weight_block(cx, cy, ct, cw, n, J.bp, g, J.table, J.bin_label, J.bin_cluster,
                           J.particle_cid, J.cluster_w, &D->es, shift, vb);
    }
  }
  team.sync();
  TAIL_STAMP(TS_CLUSTER_W);
  // ---- phase 14: the cluster with the largest weight (particle_filter.cpp:186-200)
  if (rank == 0) best_cluster_block(J.cluster_w, static_cast<size_t>(n_clusters), &D->es);
  team.sync();
  TAIL_STAMP(TS_BEST);
  // ---- phases 15, 16: moments of the best cluster and of all particles, estimate
  const unsigned mom_blocks =
      static_cast<unsigned>(min(max((n + kTailThreads - 1) / kTailThreads, static_cast<size_t>(1)),
                                static_cast<size_t>(kMomentBlocks)));
  {
    const int best = *reinterpret_cast<volatile int *>(&D->es.best);
    for (unsigned vb = rank; vb < mom_blocks; vb += size)
      moments_block(cx, cy, ct, cw, J.particle_cid, n, mom_partials, best, vb, mom_blocks);
  }
  team.sync();
  TAIL_STAMP(TS_MOMENTS);
  if (rank == 0) {
    // no clusters ("ALL clusters suck", particle_filter.cpp:206-211): out stays {0}
    moments_final(mom_partials, mom_blocks, D->sums, D->out, n_clusters > 0 ? 1 : 0);
    __syncthreads();
  }
  tail_finish<Team>(J, TAIL_OK, total, w_diff, g);
}

__global__ void __launch_bounds__(kTailThreads, 1) tail_grid_kernel(const __grid_constant__ TailJob job) {
  run_tail<GridTeam>(job);
}
// A small filter alone: one block, block barriers, an ordinary launch.
__global__ void __launch_bounds__(kTailThreads) tail_block_kernel(const __grid_constant__ TailJob job) {
  run_tail<BlockTeam>(job);
}
__global__ void __launch_bounds__(kTailThreads) tail_batch_kernel(const TailJob *__restrict__ jobs) {
  __shared__ TailJob s_job;
  // the job is read through shared memory: one copy per block instead of per thread
  const uint64_t *src = reinterpret_cast<const uint64_t *>(jobs + blockIdx.x);
  uint64_t *dst = reinterpret_cast<uint64_t *>(&s_job);
  for (unsigned k = threadIdx.x; k < sizeof(TailJob) / 8; k += blockDim.x) dst[k] = src[k];
  __syncthreads();
  run_tail<BlockTeam>(s_job);
}

}  // namespace qmcl

extern "C" {

#ifdef QMCL_RESOLVE_TRACE
// Diagnostic build only: the resolving warp's event trace of the last fused tail (prefix_dev.cuh).
qmcl_status qmcl_debug_resolve_trace(unsigned long long *out, size_t cap, size_t *n_out) {
  unsigned int n = 0;
  QMCL_CUDA(cudaDeviceSynchronize());
  QMCL_CUDA(cudaMemcpyFromSymbol(&n, qmcl::pfx::g_resolve_trace_n, sizeof(n)));
  if (n > cap) n = static_cast<unsigned int>(cap);
  if (n) QMCL_CUDA(cudaMemcpyFromSymbol(out, qmcl::pfx::g_resolve_trace, n * sizeof(unsigned long long)));
  *n_out = n;
  return QMCL_OK;
}
#endif

// Measurement hook: device time of every phase of the fused tail (ns, from
// %globaltimer stamps taken by block 0 at the phase boundaries), summed over the
// *calls fused calls this filter has completed since the last reset.
qmcl_status qmcl_filter_tail_profile(qmcl_filter *f, uint64_t out[16], uint64_t *calls,
                                     uint64_t *fallbacks, int reset) {
  if (!f || !out) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(qmcl::settle(f));
  for (int k = 0; k < 16; ++k) out[k] = f->tail_phase_ns[k];
  if (calls) *calls = f->tail_phase_calls;
  if (fallbacks) *fallbacks = f->tail_fallbacks;
  if (reset) {
    for (int k = 0; k < 16; ++k) f->tail_phase_ns[k] = 0;
    f->tail_phase_calls = 0;
  }
  return QMCL_OK;
}

const char *qmcl_tail_phase_name(int k) {
  static const char *names[16] = {"start", "normalise+chunk_sums", "summaries", "resolve", "apply+records",
                                  "select+gather", "kld_stop_rule", "bin_mark", "bin_list", "label",
                                  "weight_sum", "normalise+cluster_weights", "best_cluster", "moments",
                                  "estimate", "?"};
  return (k >= 0 && k < 16) ? names[k] : "?";
}

}  // extern "C"

// ===================================================================== host
namespace qmcl {

namespace {

constexpr size_t kTailMaxCells = size_t(1) << 26;      // dense tables: 2 x 256 MiB at most
constexpr size_t kTailOneBlock = 8192;                 // sets up to this size: the tail runs in one block

int tail_blocks_per_sm() {
  static int cached = -1;
  if (cached >= 0) return cached;
  int occ = 0, dev = 0, coop = 0;
  // without cooperative launch (some virtualised devices) the multi-launch path serves every call
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop ||
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tail_grid_kernel, kTailThreads, 0) != cudaSuccess)
    occ = 0;
  int want = 1;  // A/B at C2: one or two blocks per SM time the same; one leaves every register to the block
  if (const char *e = std::getenv("QMCL_TAIL_BLOCKS_PER_SM")) want = std::atoi(e);
  cached = std::max(0, std::min(occ, want));
  return cached;
}

// Bin bounds of the map's extent: every injected pose (a free cell's corner,
// map_base.cpp:58-64, heading in [-pi, pi)) falls inside.
void map_bin_bounds(const qmcl_filter *f, int out[6]) {
  const MapGeom &g = f->map->geom;
  const float rxy = f->params.res_xy, rt = f->params.res_theta;
  const float x0 = g.ox, y0 = g.oy;
  const float x1 = static_cast<float>(g.w - 1) * g.res + g.ox;
  const float y1 = static_cast<float>(g.h - 1) * g.res + g.oy;
  auto key = [](float v, float r) { return static_cast<int>(v / r); };  // scaled_map.h:43-48
  out[0] = std::min(key(x0, rxy), key(x1, rxy));
  out[1] = std::min(key(y0, rxy), key(y1, rxy));
  out[2] = key(static_cast<float>(-kPi), rt);
  out[3] = std::max(key(x0, rxy), key(x1, rxy));
  out[4] = std::max(key(y0, rxy), key(y1, rxy));
  out[5] = key(3.1415925f, rt);
}

}  // namespace

bool tail_enabled(const qmcl_filter *f) {
  return lazy(f) && f->use_tail && f->timer.level < 2 && f->prefix_mode != 1 && tail_blocks_per_sm() > 0;
}

bool tail_eligible(const qmcl_filter *f, int mode) {
  if (!tail_enabled(f) || !f->map->has_map) return false;
  const ParticleSet &in = f->cur();
  if (in.n == 0 || in.n > kTailMaxParticles) return false;
  const long long n_min = f->params.particle_count_min, n_max = f->params.particle_count_max;
  if (mode == TAIL_ADAPTIVE && (n_min <= 0 || static_cast<size_t>(n_min) > kTailMaxParticles)) return false;
  if (mode == TAIL_KLD && (n_max <= 0 || static_cast<size_t>(n_max) > kTailMaxParticles)) return false;
  // injection draws free cells; without any the multi-launch path reports the error
  if ((mode == TAIL_ADAPTIVE || mode == TAIL_KLD) && f->map->n_free == 0) return false;
  return true;
}

// Fills the tail job of filter f (buffers sized, every pointer set) and does the
// host-side bookkeeping of its launch: the caller launches it — tail_grid_kernel
// for one filter, tail_batch_kernel for many — right after.
qmcl_status tail_prepare(qmcl_filter *f, int mode, TailJob *Jp) {
  QMCL_TRY(flush_motion(f));  // a motion no sensor update consumed
  ParticleSet &in = f->cur();
  const bool resample = mode != TAIL_CLUSTER;
  ParticleSet &out = f->sets[(f->current + 1) % 2];
  const size_t n_in = in.n;
  const size_t n_min = static_cast<size_t>(std::max(0, f->params.particle_count_min));
  const size_t n_max = static_cast<size_t>(std::max(0, f->params.particle_count_max));
  const size_t n_draws = mode == TAIL_KLD ? n_max : (mode == TAIL_ADAPTIVE ? n_min : n_in);
  const size_t n_any = std::max(n_in, n_draws);

  TailJob &J = *Jp;
  std::memset(&J, 0, sizeof(J));
  J.mode = mode;
  J.normalise_pending = f->normalise_pending ? 1 : 0;
  // the bin bounds of the input set, taken on the device
  if (f->early_bbox_version != f->pose_version) {
    QMCL_TRY(launch_bin_bbox(f, in.x.ptr, in.y.ptr, in.t.ptr, in.n));
    f->early_bbox_version = f->pose_version;
  }
  map_bin_bounds(f, J.map_bbox);
  const double span = (static_cast<double>(J.map_bbox[3]) - J.map_bbox[0] + 5) *
                      (static_cast<double>(J.map_bbox[4]) - J.map_bbox[1] + 5) *
                      (static_cast<double>(J.map_bbox[5]) - J.map_bbox[2] + 1);
  const size_t table_cap =
      span > static_cast<double>(kTailMaxCells) ? kTailMaxCells : std::max<size_t>(static_cast<size_t>(span), 64);
  const size_t bins_cap = std::min(n_any, table_cap);

  if (resample) QMCL_TRY(out.ensure(n_draws));
  QMCL_TRY(f->cdf.ensure(n_in));
  QMCL_TRY(f->pos.ensure(n_in));
  QMCL_TRY(f->incl.ensure(n_in));
  QMCL_TRY(f->src_index.ensure(n_any));
  QMCL_TRY(f->table.ensure(table_cap));
  if (mode == TAIL_KLD) QMCL_TRY(f->tail_first.ensure(table_cap));
  QMCL_TRY(f->bin_cell.ensure(bins_cap));
  QMCL_TRY(f->bin_count.ensure(bins_cap));
  QMCL_TRY(f->bin_label.ensure(bins_cap));
  QMCL_TRY(f->bin_cluster.ensure(bins_cap));
  QMCL_TRY(f->tail_cluster_w.ensure(bins_cap));
  QMCL_TRY(f->particle_cid.ensure(n_any + 1));
  QMCL_TRY(f->tail_dev.ensure_zero(1));  // the barrier counter starts at zero
  if (!f->h_tail) QMCL_CUDA(cudaMallocHost(reinterpret_cast<void **>(&f->h_tail), sizeof(TailResult)));
  QMCL_TRY(plan_prefix_buffers(f, n_in, &J.pb, &J.nc, &J.nt));
  // tiles of the compactions: as small as one item per thread
  const size_t tiles = (std::max(std::max(n_any, table_cap), bins_cap) + kScanBlock - 1) / kScanBlock + 1;
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    const size_t at = off;
    off = (off + bytes + 15) & ~static_cast<size_t>(15);
    return at;
  };
  const size_t o_ts = carve((J.nt + 1) * 8), o_tp = carve((J.nt + 1) * 4), o_tc = carve(tiles * 4),
               o_pa = carve(kTailPartials * 8);
  QMCL_TRY(f->tail_scratch.ensure(off + 16));
  uint8_t *b = f->tail_scratch.ptr;
  J.tile_sum = reinterpret_cast<double *>(b + o_ts);
  J.tile_pos = reinterpret_cast<uint32_t *>(b + o_tp);
  J.tile_cnt = reinterpret_cast<uint32_t *>(b + o_tc);
  J.partials = reinterpret_cast<double *>(b + o_pa);

  J.ix = in.x.ptr;
  J.iy = in.y.ptr;
  J.it = in.t.ptr;
  J.iw = in.w.ptr;
  J.ox = out.x.ptr;
  J.oy = out.y.ptr;
  J.ot = out.t.ptr;
  J.ow = out.w.ptr;
  J.n_in = n_in;
  J.n_min = n_min;
  J.n_max = n_max;
  J.kld_probe = f->kld_probe;
  {
    static long solo = -1;  // QMCL_TAIL_SOLO_CELLS: A/B knob, 0 = always the team-wide labelling
    if (solo < 0) {
      const char *e = std::getenv("QMCL_TAIL_SOLO_CELLS");
      solo = e ? std::min<long>(std::atol(e), static_cast<long>(kSoloCells)) : static_cast<long>(kSoloCells);
      if (solo < 0) solo = 0;
    }
    J.solo_cells = static_cast<unsigned long long>(solo);
  }
  J.kp.res_xy = f->params.res_xy;
  J.kp.res_t = f->params.res_theta;
  J.kp.eps2 = 2 * f->params.kld_epsilon;  // space_partitioning.cpp:31
  J.kp.z = f->params.kld_z;
  J.kp.nmin = f->params.particle_count_min;
  J.kp.nmax = f->params.particle_count_max;
  J.bp.res_xy = f->params.res_xy;
  J.bp.res_t = f->params.res_theta;
  J.bp.theta_lo = f->theta_lo;
  J.bp.theta_hi = f->theta_hi;
  J.geom = f->map->geom;
  J.free_cells = f->map->free_cells.ptr;
  J.n_free = f->map->n_free;
  J.seed = f->seed;
  J.step = f->step;
  J.w_slow = f->w_slow;
  J.w_fast = f->w_fast;
  J.alpha_slow = f->params.alpha_slow;
  J.alpha_fast = f->params.alpha_fast;
  J.sc = f->scalars.ptr;
  J.bbox = f->bbox.ptr;
  J.cdf = f->cdf.ptr;
  J.pos = f->pos.ptr;
  J.incl = f->incl.ptr;
  J.src = f->src_index.ptr;
  J.first = f->tail_first.ptr;
  J.table = f->table.ptr;
  J.table_cap = table_cap;
  J.bin_cell = f->bin_cell.ptr;
  J.bin_count = f->bin_count.ptr;
  J.bin_label = f->bin_label.ptr;
  J.bin_cluster = f->bin_cluster.ptr;
  J.bins_cap = bins_cap;
  J.cluster_w = f->tail_cluster_w.ptr;
  J.particle_cid = f->particle_cid.ptr;
  J.dev = f->tail_dev.ptr;
  J.result = f->h_tail;
  J.serial = ++f->tail_serial;

  g_d2h_bytes.fetch_add(sizeof(TailResult), std::memory_order_relaxed);  // written by the kernel itself

  // bookkeeping of a call whose outcome the host has not seen yet
  f->tail_pending = true;
  f->tail_mode = mode;
  f->tail_had_normalise = f->normalise_pending;
  f->tail_n_total = static_cast<double>(f->normalise_n);
  f->normalise_pending = false;
  f->tail_prev_current = f->current;
  f->tail_prev_step = f->step;
  f->est_cached = false;
  f->clusters_valid = false;
  f->bins_built = false;
  f->bins_pending = f->clusters_pending = false;
  if (resample) {
    f->current = (f->current + 1) % 2;
    f->step++;
    f->pose_version++;
    out.n = n_draws;  // an upper bound until the result arrives
  }
  return QMCL_OK;
}

qmcl_status launch_tail(qmcl_filter *f, int mode) {
  TailJob J;
  QMCL_TRY(tail_prepare(f, mode, &J));
  StageScope sc(f, QMCL_STAGE_TAIL, 1);
  if (J.n_in <= kTailOneBlock && (mode == TAIL_CLUSTER || mode == TAIL_LOW_VARIANCE ||
                                  (mode == TAIL_KLD ? J.n_max : J.n_min) <= kTailOneBlock)) {
    // small sets (QuickMCL's defaults): one block between block barriers beats 148 blocks
    // between grid barriers, and leaves the other SMs to concurrent filters
    tail_block_kernel<<<1, kTailThreads, 0, f->map->stream>>>(J);
    QMCL_LAUNCHED();
    return QMCL_OK;
  }
  void *args[] = {&J};
  const unsigned blocks = static_cast<unsigned>(kSMs * tail_blocks_per_sm());
  QMCL_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(tail_grid_kernel), dim3(blocks),
                                        dim3(kTailThreads), args, 0, f->map->stream));
  QMCL_LAUNCHED();
  return QMCL_OK;
}

// One block per job: the batch form (jobs in device memory).
qmcl_status launch_tail_batch(qmcl_map *m, const TailJob *dev_jobs, unsigned n_jobs) {
  if (!n_jobs) return QMCL_OK;
  tail_batch_kernel<<<n_jobs, kTailThreads, 0, m->stream>>>(dev_jobs);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

// Folds the result of the pending tail into the host's view of the filter.  The
// stream has been synchronised.  A fall-back re-runs the call on the multi-launch
// path and returns its status.
qmcl_status tail_complete(qmcl_filter *f) {
  if (!f->tail_pending) return QMCL_OK;
  f->tail_pending = false;
  const TailResult r = *f->h_tail;
  if (r.serial != f->tail_serial) {
    set_error("fused tail: result block is stale (serial %llu, expected %llu)",
              static_cast<unsigned long long>(r.serial), static_cast<unsigned long long>(f->tail_serial));
    return QMCL_ERR_CUDA;
  }
  const int mode = f->tail_mode;
  const bool resample = mode != TAIL_CLUSTER;
  {  // phase durations, accumulated for qmcl_filter_tail_profile
    unsigned long long prev = r.stamps[TS_START];
    for (int k = 1; k < kTailStamps; ++k) {
      const unsigned long long t = r.stamps[k];
      if (t >= prev && t != 0) {
        f->tail_phase_ns[k] += t - prev;
        prev = t;
      }
    }
    f->tail_phase_calls++;
  }
  if (r.status == TAIL_FALLBACK || r.status == TAIL_FALLBACK_NORMALISED) {
    if (resample) {  // undo launch_tail's bookkeeping; the multi-launch path redoes it
      f->current = f->tail_prev_current;
      f->step = f->tail_prev_step;
    }
    if (f->tail_had_normalise) {
      if (r.status == TAIL_FALLBACK_NORMALISED) apply_sensor_total(f, r.total_weight, f->tail_n_total);
      else f->normalise_pending = true;
    }
    f->tail_fallbacks++;
    const bool saved = f->use_tail;
    f->use_tail = false;
    const qmcl_status st = resample ? qmcl_filter_resample_and_cluster(f) : qmcl_filter_cluster(f);
    f->use_tail = saved;
    return st;
  }
  if (f->tail_had_normalise) apply_sensor_total(f, r.total_weight, f->tail_n_total);
  ParticleSet &cur = f->cur();
  if (resample) {
    cur.n = static_cast<size_t>(r.n_out);
    f->n_src_index = cur.n;
    if (mode == TAIL_KLD) {
      // next time, look at twice as many candidates as this stop needed (a power of two)
      size_t probe = 4096;
      while (probe < 2 * cur.n) probe *= 2;
      f->kld_probe = probe;
    }
    f->global_first = 0;
    f->global_n = cur.n;
    if ((mode == TAIL_ADAPTIVE || mode == TAIL_KLD) && r.status == TAIL_OK && r.w_diff > 0.0) {
      // particle_filter.cpp:284-290, 338-344
      f->w_fast = 0;
      f->w_slow = 0;
    }
  }
  f->grid = BinGrid{static_cast<int>(r.grid[0]), static_cast<int>(r.grid[1]), static_cast<int>(r.grid[2]),
                    static_cast<int>(r.grid[3]), static_cast<int>(r.grid[4]), static_cast<int>(r.grid[5])};
  f->n_bins = static_cast<size_t>(r.n_bins);
  f->n_clusters = static_cast<size_t>(r.n_clusters);
  f->bins_pending = f->clusters_pending = false;
  f->bins_built = true;
  f->clusters_valid = true;
  f->bbox_hint_valid = false;
  f->est_cached = true;
  f->est_invalid = r.invalid != 0;
  std::memcpy(f->est, r.est, sizeof(f->est));
  return QMCL_OK;
}

}  // namespace qmcl
