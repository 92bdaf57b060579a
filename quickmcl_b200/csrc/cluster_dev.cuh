// cluster_dev.cuh — device side of K13 / K14 (see cluster.cu): bin keys, the
// lock-free union-find and the hub-cover hook, fixed-point cluster weights,
// ordered moments.  Bodies are written against a *virtual* block index so that
// the stand-alone kernels (cluster.cu) and the fused tail (tail.cu), which walks
// the same virtual blocks with however many real blocks it has, produce the
// same bits.  Plain pointers: the tail calls them on data written earlier in
// the same launch.
#pragma once

#include "state.cuh"

#include <climits>

namespace qmcl {

struct BinParams {
  float res_xy, res_t;
  int theta_lo, theta_hi;
};

// space_partitioning.cpp:45-50 via scaled_map.h:43-48: float division, then
// truncation toward zero, no offset.
__device__ __forceinline__ void bin_key(float x, float y, float t, const BinParams &bp, int *kx,
                                        int *ky, int *kt) {
  *kx = cell_index_inf(x, bp.res_xy);
  *ky = cell_index_inf(y, bp.res_xy);
  *kt = cell_index_inf(t, bp.res_t);
}

__device__ __forceinline__ long long grid_index(const BinGrid &g, int kx, int ky, int kt) {
  const long long ix = static_cast<long long>(kx) - g.x0;
  const long long iy = static_cast<long long>(ky) - g.y0;
  const long long it = static_cast<long long>(kt) - g.t0;
  if (ix < 0 || iy < 0 || it < 0 || ix >= g.nx || iy >= g.ny || it >= g.nt) return -1;
  return (ix * g.ny + iy) * g.nt + it;
}

// bbox[0..2] = min (x,y,t), bbox[3..5] = max.
__device__ __forceinline__ void bin_bbox_block(const float *x, const float *y, const float *t, size_t n,
                                               const BinParams &bp, int *bbox, unsigned vblock,
                                               unsigned nvblocks) {
  int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
  for (size_t i = static_cast<size_t>(vblock) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(nvblocks) * blockDim.x) {
    int k[3];
    bin_key(x[i], y[i], t[i], bp, &k[0], &k[1], &k[2]);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      mn[d] = min(mn[d], k[d]);
      mx[d] = max(mx[d], k[d]);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = min(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = max(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  }
  __shared__ int s_mn[8][3], s_mx[8][3];
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      s_mn[wid][d] = mn[d];
      s_mx[wid][d] = mx[d];
    }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    int a = INT_MAX, b = INT_MIN;
    for (int k = 0; k < 8; ++k) {
      a = min(a, s_mn[k][threadIdx.x]);
      b = max(b, s_mx[k][threadIdx.x]);
    }
    atomicMin(bbox + threadIdx.x, a);
    atomicMax(bbox + 3 + threadIdx.x, b);
  }
}

static __global__ void bbox_init_kernel(int *bbox) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = INT_MAX;
  else if (threadIdx.x < 6) bbox[threadIdx.x] = INT_MIN;
}

// SpacePartitioning::add_point for every particle: table[idx] counts members.
// One particle slot per thread; every lane of the warp must call (i >= n: no mark).
__device__ __forceinline__ void bin_mark_one(const float *x, const float *y, const float *t, size_t n,
                                             const BinParams &bp, const BinGrid &g, int32_t *table,
                                             size_t i) {
  if (i >= n) return;
  int kx, ky, kt;
  bin_key(x[i], y[i], t[i], bp, &kx, &ky, &kt);
  const long long idx = grid_index(g, kx, ky, kt);
  // one atomic per distinct bin per warp (a tracking cloud has ~100 bins for
  // 10^5 particles, so per-particle atomics would serialise on a few addresses)
  const unsigned peers = __match_any_sync(__activemask(), idx);
  if (idx >= 0 && (__ffs(peers) - 1) == static_cast<int>(threadIdx.x & 31))
    atomicAdd(table + idx, __popc(peers));
}

// Sharded filters exchange bin *occupancy* only: one 0/1 byte per cell, summed
// over the ranks word-wise (a quarter of the int32 count table; the per-bin
// counts are made on demand, qmcl_filter_copy_bins).  Bytes are set through the
// word that holds them.
__device__ __forceinline__ void bin_mark_occupancy(const float *x, const float *y, const float *t, size_t n,
                                                   const BinParams &bp, const BinGrid &g,
                                                   unsigned *occ_words, size_t i) {
  if (i >= n) return;
  int kx, ky, kt;
  bin_key(x[i], y[i], t[i], bp, &kx, &ky, &kt);
  const long long idx = grid_index(g, kx, ky, kt);
  const unsigned peers = __match_any_sync(__activemask(), idx);
  if (idx >= 0 && (__ffs(peers) - 1) == static_cast<int>(threadIdx.x & 31))
    atomicOr(occ_words + (idx >> 2), 1u << (8 * static_cast<unsigned>(idx & 3)));
}
struct OccupiedByte {
  const uint8_t *occ;
  __device__ uint32_t operator()(size_t i) const { return occ[i] ? 1u : 0u; }
};
// Compacted bin list from the occupancy bytes; every cell of the table is written
// (cell -> bin id + 1, 0 for an empty cell), so the table needs no clearing.
struct EmitBinOccupancy {
  const uint8_t *occ;
  int32_t *table;
  int32_t *bin_cell;
  int32_t *parent;
  int nt;  // runs along theta start out linked (see EmitBin); 0: off
  size_t run_next = ~static_cast<size_t>(0);
  int32_t run_head = 0;
  __device__ void operator()(size_t i, uint32_t pos, uint32_t flag) {
    table[i] = flag ? static_cast<int32_t>(pos) + 1 : 0;
    if (!flag) return;
    bin_cell[pos] = static_cast<int32_t>(i);
    const bool linked = nt > 0 && pos > 0 && (i % static_cast<size_t>(nt)) != 0 && occ[i - 1] != 0;
    int32_t p = static_cast<int32_t>(pos);
    if (linked) p = (i == run_next) ? run_head : p - 1;
    if (!linked || i != run_next) run_head = static_cast<int32_t>(pos);
    run_next = i + 1;
    parent[pos] = p;
  }
};
// Per-bin particle counts of this rank's particles (table: cell -> bin id + 1).
__device__ __forceinline__ void bin_count_one(const float *x, const float *y, const float *t, size_t n,
                                              const BinParams &bp, const BinGrid &g, const int32_t *table,
                                              int32_t *bin_count, size_t i) {
  if (i >= n) return;
  int kx, ky, kt;
  bin_key(x[i], y[i], t[i], bp, &kx, &ky, &kt);
  const long long idx = grid_index(g, kx, ky, kt);
  const unsigned peers = __match_any_sync(__activemask(), idx);
  if (idx >= 0 && (__ffs(peers) - 1) == static_cast<int>(threadIdx.x & 31)) {
    const int b = table[idx] - 1;
    if (b >= 0) atomicAdd(bin_count + b, __popc(peers));
  }
}

struct OccupiedFlag {
  const int32_t *table;
  __device__ uint32_t operator()(size_t i) const { return table[i] > 0 ? 1u : 0u; }
};
// Writes the compacted bin list and turns the table into cell -> bin id + 1.
// The forest starts with the runs along theta already linked: a bin whose predecessor
// in its (x, y) column is occupied - the previous bin of the sorted list, one cell back
// in memory - points at it.  That edge is part of the neighbour relation in either
// reading of it (cc_hook_bin), the link keeps the forest's invariant (parent < self),
// and on a dense grid (global localisation: 36 bins per column) it turns most of the
// labelling's scattered unions into walks along consecutive memory.  nt = 0: off.
struct EmitBin {
  int32_t *table;
  int32_t *bin_cell;
  int32_t *bin_count;
  int32_t *parent;
  int nt;
  // the scan hands a thread its items in order: the bin at which the thread's current
  // run began, so that the run's later bins point there directly (short chains)
  size_t run_next = ~static_cast<size_t>(0);
  int32_t run_head = 0;
  __device__ void operator()(size_t i, uint32_t pos, uint32_t flag) {
    if (!flag) return;
    bin_cell[pos] = static_cast<int32_t>(i);
    bin_count[pos] = table[i];
    // table[i - 1] is a count (> 0) or already that cell's bin id + 1 (> 0) when occupied
    const bool linked = nt > 0 && pos > 0 && (i % static_cast<size_t>(nt)) != 0 && table[i - 1] != 0;
    int32_t p = static_cast<int32_t>(pos);
    if (linked) p = (i == run_next) ? run_head : p - 1;
    if (!linked || i != run_next) run_head = static_cast<int32_t>(pos);
    run_next = i + 1;
    parent[pos] = p;
    table[i] = static_cast<int32_t>(pos) + 1;
  }
};

// Lock-free union-find.  Invariant: parent[v] < v for every non-root, roots
// point to themselves, a node stops being a root exactly once (the CAS in
// uf_union).  Path halving replaces parent[v] by an ancestor, which keeps the
// invariant under any interleaving; a stale (L1) read only shows an older,
// less merged forest and at worst costs a CAS retry.
__device__ __forceinline__ int uf_find(int32_t *parent, int v) {
  int p = parent[v];
  while (p != v) {
    const int gp = parent[p];
    if (gp != p) parent[v] = gp;  // path halving (benign race)
    v = p;
    p = gp;
  }
  return v;
}
// Read-only find for the flatten pass: there every thread publishes the final
// root of its own bin, and a halving store from another thread arriving later
// would overwrite that root with an intermediate ancestor.
__device__ __forceinline__ int uf_find_readonly(const int32_t *parent, int v) {
  int p = __ldcg(parent + v);
  while (p != v) {
    v = p;
    p = __ldcg(parent + v);
  }
  return v;
}

__device__ __forceinline__ void uf_union(int32_t *parent, int a, int b) {
  while (true) {
    a = uf_find(parent, a);
    b = uf_find(parent, b);
    if (a == b) return;
    if (a < b) {
      const int tmp = a;
      a = b;
      b = tmp;
    }
    // hook the larger root under the smaller one: roots end up being the
    // smallest bin index of their component
    if (atomicCAS(parent + a, a, b) == a) return;
  }
}

// SpacePartitioning::do_cluster's neighbour relation (space_partitioning.cpp:
// 102-128): theta' = theta + {-1,0,1} wrapped with "< lo -> hi, > hi -> lo",
// then (x+dx, y+dy) for dx,dy in {-1,0,1}.  One thread per bin.
//
// Every reciprocal edge has to be accounted for once, by the bin with the
// larger index.  Most edges need no union of their own: two occupied cells
// that are both earlier neighbours of this bin and neighbours of each other are
// already joined through edges between earlier bins (induction over the bin
// index), so the bin only unions with a set of "hubs" such that every other
// occupied earlier neighbour is adjacent to a hub.  The hubs are picked
// greedily in a fixed order that starts with A = (x-1, y, t), which is adjacent
// to all 13 other candidates, then B = (x, y-1, t).  On a dense cloud (global
// localisation) that is one table read and one union per bin instead of 27 and
// up to 13; components, roots and cluster ids are the same.
// Bins outside [theta_lo, theta_hi] (theta == -pi rounds to bin -18 at 10
// degrees) see neighbours that do not see them back (space_partitioning.cpp:
// 110-114 wraps them onto the other end), so they union everything they see.
namespace hook {
// candidates of a regular bin: offsets whose cell precedes it in (x, y, t) order
// (the two in-column ones are checked by index at run time because of the wrap)
constexpr int kCand = 14;
constexpr int kDx[kCand] = {-1, 0, -1, -1, -1, -1, -1, -1, -1, -1, 0, 0, 0, 0};
constexpr int kDy[kCand] = {0, -1, 0, 0, -1, -1, -1, 1, 1, 1, -1, -1, 0, 0};
constexpr int kDt[kCand] = {0, 0, -1, 1, -1, 0, 1, -1, 0, 1, -1, 1, -1, 1};
constexpr bool adjacent(int i, int j) {
  const int ax = kDx[i] - kDx[j], ay = kDy[i] - kDy[j], at = kDt[i] - kDt[j];
  // theta steps -1 and +1 are two apart (adjacent only on rings of three bins
  // or fewer; treating them as apart is the safe side)
  return ax >= -1 && ax <= 1 && ay >= -1 && ay <= 1 && at >= -1 && at <= 1;
}
constexpr unsigned adjacency_mask(int i) {
  unsigned m = 0;
  for (int j = 0; j < kCand; ++j)
    if (adjacent(i, j)) m |= 1u << j;
  return m;
}
// offsets as arithmetic on the (unrolled) candidate number, checked below
#define QMCL_HOOK_DX(k) ((k) == 1 || (k) >= 10 ? 0 : -1)
#define QMCL_HOOK_DY(k) ((k) == 1 || (k) == 10 || (k) == 11 ? -1 : (k) >= 12 || (k) == 0 || (k) == 2 || (k) == 3 ? 0 : (k) <= 6 ? -1 : 1)
#define QMCL_HOOK_DT(k) ((k) <= 1 ? 0 : (k) == 2 ? -1 : (k) == 3 ? 1 : (k) <= 9 ? ((k) - 4) % 3 - 1 : ((k) % 2 ? 1 : -1))
constexpr bool offsets_ok() {
  for (int k = 0; k < kCand; ++k)
    if (QMCL_HOOK_DX(k) != kDx[k] || QMCL_HOOK_DY(k) != kDy[k] || QMCL_HOOK_DT(k) != kDt[k]) return false;
  return true;
}
static_assert(offsets_ok(), "QMCL_HOOK_D* out of date");
// adjacency_mask(k), spelled out so that device code sees plain literals
#define QMCL_HOOK_ADJ(k) \
  ((k) == 0 ? 0x3fffu : (k) == 1 ? 0x3c7fu : (k) == 2 ? 0x15b7u : (k) == 3 ? 0x2b6bu : (k) == 4 ? 0x1437u : \
   (k) == 5 ? 0x3c7fu : (k) == 6 ? 0x286bu : (k) == 7 ? 0x1185u : (k) == 8 ? 0x338du : (k) == 9 ? 0x2309u : \
   (k) == 10 ? 0x1437u : (k) == 11 ? 0x286bu : (k) == 12 ? 0x15b7u : 0x2b6bu)
constexpr bool masks_ok() {
  for (int k = 0; k < kCand; ++k)
    if (QMCL_HOOK_ADJ(k) != adjacency_mask(k)) return false;
  return true;
}
static_assert(masks_ok(), "QMCL_HOOK_ADJ out of date");
static_assert(adjacency_mask(0) == (1u << kCand) - 1, "A must be adjacent to every candidate");
}  // namespace hook

// linked_runs: the forest came with the theta runs linked (EmitBin).  A bin whose theta
// predecessor is occupied, and whose hub A = (x-1, y, t) has an occupied predecessor
// too, then needs no union of its own: the two predecessors were joined one step down
// (by this rule or by their union), and each is linked to its successor.  On a dense
// grid that leaves one union per pair of neighbouring runs instead of one per bin.
__device__ __forceinline__ void cc_hook_bin(const int32_t *table, const int32_t *bin_cell,
                                            const BinGrid &g, const BinParams &bp, int32_t *parent,
                                            int b, bool linked_runs = false) {
  const int cell = bin_cell[b];
  const int it = cell % g.nt;
  const int iy = (cell / g.nt) % g.ny;
  const int ix = cell / (g.nt * g.ny);
  const int kt0 = it + g.t0;
  const bool regular = kt0 >= bp.theta_lo && kt0 <= bp.theta_hi;
  auto wrap = [&](int kt) { return kt < bp.theta_lo ? bp.theta_hi : (kt > bp.theta_hi ? bp.theta_lo : kt); };
  // bin id + 1 of the cell, 0 when empty or outside the grid
  auto occupant = [&](int jx, int jy, int kt) -> int {
    const int jt = kt - g.t0;
    if (jx < 0 || jx >= g.nx || jy < 0 || jy >= g.ny || jt < 0 || jt >= g.nt) return 0;
    return table[(static_cast<size_t>(jx) * g.ny + jy) * g.nt + jt];
  };
  const int kts[3] = {wrap(kt0 - 1), wrap(kt0), wrap(kt0 + 1)};
  if (!regular) {
    int v[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) v[k] = occupant(ix + (k / 3) % 3 - 1, iy + k % 3 - 1, kts[k / 9]);
#pragma unroll
    for (int k = 0; k < 27; ++k)
      if (v[k] > 0 && v[k] - 1 != b) uf_union(parent, b, v[k] - 1);
    return;
  }
  // Loads first, unions after: the unions' atomics would otherwise serialise
  // the (independent) table reads.
  int v[hook::kCand];
  v[0] = occupant(ix - 1, iy, kt0);
  v[1] = occupant(ix, iy - 1, kt0);
  if (v[0] > 0) {  // adjacent to every other candidate
    if (linked_runs && it > 0 && kt0 - 1 >= bp.theta_lo && table[cell - 1] > 0 &&
        occupant(ix - 1, iy, kt0 - 1) > 0)
      return;
    uf_union(parent, b, v[0] - 1);
    return;
  }
#pragma unroll
  for (int k = 2; k < hook::kCand; ++k) {
    v[k] = occupant(ix + QMCL_HOOK_DX(k), iy + QMCL_HOOK_DY(k), kts[QMCL_HOOK_DT(k) + 1]);
    if (v[k] - 1 >= b) v[k] = 0;  // in-column neighbours across the wrap: not an earlier bin
  }
  unsigned hubs = 0;
#pragma unroll
  for (int k = 1; k < hook::kCand; ++k) {
    if (v[k] > 0 && !(QMCL_HOOK_ADJ(k) & hubs)) {
      uf_union(parent, b, v[k] - 1);
      hubs |= 1u << k;
    }
  }
}


struct RootFlag {
  const int32_t *label;
  const unsigned long long *n_bins_dev;  // non-null: the scan runs over a bound
  __device__ uint32_t operator()(size_t b) const {
    if (n_bins_dev && b >= *n_bins_dev) return 0u;
    return label[b] == static_cast<int32_t>(b) ? 1u : 0u;
  }
};
struct EmitRootId {
  int32_t *cluster_of_root;
  __device__ void operator()(size_t b, uint32_t pos, uint32_t flag) const {
    if (flag) cluster_of_root[b] = static_cast<int32_t>(pos);
  }
};

// ---------------------------------------------------------------- K14
// Pose estimate = weighted mean of the cluster with the largest total weight,
// covariance = statistics of all particles (particle_filter.cpp:186-212,
// statistics.cpp:23-79).  Only two sets of moments are ever read — the best
// cluster's and the global one — so instead of accumulating ten doubles per
// cluster with floating-point atomics (contended, and order-dependent), this is
// done in three deterministic steps:
//   1. wmax_kernel: largest weight (fixes the fixed-point scale);
//   2. cluster_weight_kernel: per-cluster total weight as an *integer*
//      fixed-point sum (order-independent, warp-aggregated), cluster id per
//      particle kept for step 3;
//   3. best_cluster_kernel + moments_kernel: argmax, then the moments of the
//      best cluster and of all particles, reduced in a fixed order.
constexpr int kMom = 10;  // 0 W, 1..4 S = w*(x, y, sin, cos), 5..8 C, 9 count
constexpr int kMomentBlocks = 4 * kSMs;

struct EstimateScalars {
  unsigned long long wmax_bits;   // largest weight, as the bits of a non-negative double
  unsigned long long best_w;      // fixed-point total of the best cluster
  int best;                       // its id, -1 if none
  int invalid;                    // a particle's bin has no cluster
  unsigned int blocks_done;
  unsigned int pad;
};


// Scale 2^k with n * wmax * 2^k < 2^62.
__device__ __forceinline__ int fixed_point_shift(unsigned long long wmax_bits, size_t n) {
  if (wmax_bits == 0) return 0;
  const int e = static_cast<int>(wmax_bits >> 52) - 1023 + 1;  // wmax < 2^e
  const int ln = 64 - __clzll(static_cast<unsigned long long>(n));  // n < 2^ln
  return 62 - e - ln;
}

// One block of 256 particle slots (vblock); all threads of the block must call.
QMCL_PHASE_FN void cluster_weight_block(
    const float *x, const float *y, const float *t, const double *w, size_t n, const BinParams &bp,
    const BinGrid &g, const int32_t *table, const int32_t *label, const int32_t *cluster_of_root,
    int32_t *cid_out, unsigned long long *cluster_w, EstimateScalars *es, int shift, size_t vblock) {
  const size_t i = vblock * blockDim.x + threadIdx.x;
  int cid = -1;
  unsigned long long fx = 0;
  if (i < n) {
    int kx, ky, kt;
    bin_key(x[i], y[i], t[i], bp, &kx, &ky, &kt);
    const long long idx = grid_index(g, kx, ky, kt);
    const int b = idx >= 0 ? table[idx] - 1 : -1;
    if (b < 0) {
      atomicExch(&es->invalid, 1);  // particle_filter.cpp:375-379
    } else {
      cid = cluster_of_root[label[b]];
      const double wi = w[i];
      if (wi > 0) fx = static_cast<unsigned long long>(__double2ll_rn(scalbn(wi, shift)));
    }
    cid_out[i] = cid;
  }
  // aggregate lanes of the same cluster: three 21-bit limbs through REDUX
  const unsigned mask = __match_any_sync(0xffffffffu, cid);
  const unsigned l0 = __reduce_add_sync(mask, static_cast<unsigned>(fx & 0x1fffffu));
  const unsigned l1 = __reduce_add_sync(mask, static_cast<unsigned>((fx >> 21) & 0x1fffffu));
  const unsigned l2 = __reduce_add_sync(mask, static_cast<unsigned>(fx >> 42));
  const unsigned long long sum = static_cast<unsigned long long>(l0) +
                                 (static_cast<unsigned long long>(l1) << 21) +
                                 (static_cast<unsigned long long>(l2) << 42);
  const bool leader = (__ffs(mask) - 1) == static_cast<int>(threadIdx.x & 31);
  // block-uniform clusters (tracking): one atomic per block
  __shared__ unsigned long long s_sum[8];
  __shared__ int s_cid[8];
  const int wid = threadIdx.x >> 5;
  const bool warp_uniform = mask == 0xffffffffu;
  if ((threadIdx.x & 31) == 0) {
    s_sum[wid] = sum;
    s_cid[wid] = warp_uniform ? cid : -2;
  }
  __syncthreads();
  bool block_uniform = true;
#pragma unroll
  for (int k = 0; k < 8; ++k) block_uniform &= (s_cid[k] == s_cid[0]) && s_cid[k] != -2;
  if (block_uniform) {
    if (threadIdx.x == 0 && s_cid[0] >= 0) {
      unsigned long long tot = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) tot += s_sum[k];
      if (tot) atomicAdd(cluster_w + s_cid[0], tot);
    }
  } else if (leader && cid >= 0 && sum) {
    atomicAdd(cluster_w + cid, sum);
  }
  __syncthreads();  // s_sum / s_cid are reused by the caller's next block
}

// First maximum of the cluster weights (ties -> smallest canonical id).
// Run by one whole block (any multiple of 32 threads up to 1024).
QMCL_PHASE_FN void best_cluster_block(const unsigned long long *cluster_w,
                                                   size_t n_clusters, EstimateScalars *es) {
  __shared__ unsigned long long s_w[32];
  __shared__ long long s_i[32];
  unsigned long long bw = 0;
  long long bi = -1;
  for (size_t c = threadIdx.x; c < n_clusters; c += blockDim.x) {
    const unsigned long long v = cluster_w[c];
    if (bi < 0 || v > bw) {
      bw = v;
      bi = static_cast<long long>(c);
    }
  }
  auto better = [](unsigned long long w1, long long i1, unsigned long long w2, long long i2) {
    if (i1 < 0) return false;
    if (i2 < 0) return true;
    return w1 > w2 || (w1 == w2 && i1 < i2);
  };
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long ow = __shfl_xor_sync(0xffffffffu, bw, o);
    const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (better(ow, oi, bw, bi)) {
      bw = ow;
      bi = oi;
    }
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) {
    s_w[wid] = bw;
    s_i[wid] = bi;
  }
  __syncthreads();
  if (wid == 0) {
    const bool have = lane < static_cast<int>(blockDim.x >> 5);
    bw = have ? s_w[lane] : 0ull;
    bi = have ? s_i[lane] : -1ll;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long ow = __shfl_xor_sync(0xffffffffu, bw, o);
      const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (better(ow, oi, bw, bi)) {
        bw = ow;
        bi = oi;
      }
    }
    if (lane == 0) {
      es->best = static_cast<int>(bi);
      es->best_w = bw;
    }
  }
}

__device__ __forceinline__ double block_sum_256(double v, double *smem) {
  v = warp_sum_xor(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) smem[wid] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) t = __dadd_rn(t, smem[i]);
  }
  __syncthreads();
  return t;
}

// statistics.cpp:50-79 + particle_filter.cpp:201-211 on the reduced moments
// (b = best cluster, gsum = all particles).  out = {valid, pose[3], cov[9]}.
__host__ __device__ inline void finalise_estimate(const double *b, const double *gsum, double *out) {
  for (int k = 0; k < 13; ++k) out[k] = 0.0;
  if (!(b[0] > 0)) return;
  out[0] = 1.0;
  out[1] = b[1] / b[0];
  out[2] = b[2] / b[0];
  out[3] = atan2(b[3], b[4]);
  const double W = gsum[0];
  const double m0 = gsum[1] / W, m1 = gsum[2] / W;
  double *cov = out + 4;
  cov[0] = gsum[5] / W - m0 * m0;
  cov[1] = gsum[6] / W - m0 * m1;
  cov[3] = gsum[7] / W - m1 * m0;
  cov[4] = gsum[8] / W - m1 * m1;
  const double R = sqrt(gsum[3] * gsum[3] + gsum[4] * gsum[4]);  // un-normalised, as the reference
  cov[8] = 1 - R;
}

// statistics.cpp:23-41 for the best cluster and for all particles.  Every
// thread sums its grid-stride particles in index order, blocks reduce in a
// fixed tree, the last block adds the block partials in block order: the
// result depends on n only, not on scheduling.  partials: [gridDim][2*kMom].
// When `finalise` is set the last block also writes out[0..13).
// Partial moments of virtual block `vblock` of `nvblocks` -> partials[vblock][2*kMom].
QMCL_PHASE_FN void moments_block(const float *x, const float *y, const float *t,
                                              const double *w, const int32_t *cid, size_t n,
                                              double *partials, int best, unsigned vblock,
                                              unsigned nvblocks) {
  __shared__ double s_mom[2 * kMom][8];
  double vg[kMom], vb[kMom];
#pragma unroll
  for (int k = 0; k < kMom; ++k) vg[k] = vb[k] = 0.0;
  for (size_t i = static_cast<size_t>(vblock) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(nvblocks) * blockDim.x) {
    const int c = cid[i];
    if (c < 0) continue;
    const float px = x[i], py = y[i], pt = t[i];
    const double wi = w[i];
    double sn, cs;
    sincos(static_cast<double>(pt), &sn, &cs);
    const double d0 = static_cast<double>(px), d1 = static_cast<double>(py);
    double v[kMom];
    v[0] = wi;
    v[1] = wi * d0;
    v[2] = wi * d1;
    v[3] = wi * sn;
    v[4] = wi * cs;
    v[5] = (wi * d0) * d0;
    v[6] = (wi * d0) * d1;
    v[7] = (wi * d1) * d0;
    v[8] = (wi * d1) * d1;
    v[9] = 1.0;
    const bool in_best = c == best;
#pragma unroll
    for (int k = 0; k < kMom; ++k) {
      vg[k] = __dadd_rn(vg[k], v[k]);
      if (in_best) vb[k] = __dadd_rn(vb[k], v[k]);
    }
  }
  // block_sum_256 for all twenty sums with one exchange: the warp butterflies, then
  // the eight warp sums of each moment added in warp order (the same additions)
  double *mine = partials + static_cast<size_t>(vblock) * 2 * kMom;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kMom; ++k) {
    const double sg = warp_sum_xor(vg[k]);
    const double sb = warp_sum_xor(vb[k]);
    if (lane == 0) {
      s_mom[k][wid] = sg;
      s_mom[kMom + k][wid] = sb;
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * kMom) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) t = __dadd_rn(t, s_mom[threadIdx.x][i]);
    mine[threadIdx.x] = t;
  }
  __syncthreads();
}
// The block partials added in block order (one block; all its threads call).
QMCL_PHASE_FN void moments_final(const double *partials, unsigned nvblocks,
                                              double *sums_out, double *out, int finalise) {
  __shared__ double s_tot[2 * kMom];
  if (threadIdx.x < 2 * kMom) {
    double a = 0.0;
    for (unsigned b = 0; b < nvblocks; ++b)
      a = __dadd_rn(a, __ldcg(partials + static_cast<size_t>(b) * 2 * kMom + threadIdx.x));
    s_tot[threadIdx.x] = a;
    sums_out[threadIdx.x] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0 && finalise) finalise_estimate(s_tot + kMom, s_tot, out);
  __syncthreads();
}


}  // namespace qmcl
