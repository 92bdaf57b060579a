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

#include "scan.cuh"
#include "shard.cuh"

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
__global__ void __launch_bounds__(256)
bin_bbox_kernel(const float *__restrict__ x, const float *__restrict__ y,
                const float *__restrict__ t, size_t n, BinParams bp, int *__restrict__ bbox) {
  int mn[3] = {INT_MAX, INT_MAX, INT_MAX}, mx[3] = {INT_MIN, INT_MIN, INT_MIN};
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
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

__global__ void bbox_init_kernel(int *bbox) {
  if (threadIdx.x < 3) bbox[threadIdx.x] = INT_MAX;
  else if (threadIdx.x < 6) bbox[threadIdx.x] = INT_MIN;
}

// SpacePartitioning::add_point for every particle: table[idx] counts members.
__global__ void __launch_bounds__(256)
bin_mark_kernel(const float *__restrict__ x, const float *__restrict__ y,
                const float *__restrict__ t, size_t n, BinParams bp, BinGrid g,
                int32_t *__restrict__ table) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
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

struct OccupiedFlag {
  const int32_t *table;
  __device__ uint32_t operator()(size_t i) const { return table[i] > 0 ? 1u : 0u; }
};
// Writes the compacted bin list and turns the table into cell -> bin id + 1.
struct EmitBin {
  int32_t *table;
  int32_t *bin_cell;
  int32_t *bin_count;
  int32_t *parent;
  __device__ void operator()(size_t i, uint32_t pos, uint32_t flag) const {
    if (!flag) return;
    bin_cell[pos] = static_cast<int32_t>(i);
    bin_count[pos] = table[i];
    parent[pos] = static_cast<int32_t>(pos);
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

__global__ void __launch_bounds__(256)
cc_hook_kernel(const int32_t *__restrict__ table, const int32_t *__restrict__ bin_cell,
               size_t n_bins, const unsigned long long *__restrict__ n_bins_dev, BinGrid g,
               BinParams bp, int32_t *__restrict__ parent) {
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  // n_bins is a launch bound while the count is still on its way to the host
  if (n_bins_dev) n_bins = static_cast<size_t>(*n_bins_dev);
  if (tid >= n_bins) return;
  const int b = static_cast<int>(tid);
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

__global__ void __launch_bounds__(256)
cc_flatten_kernel(int32_t *__restrict__ parent, size_t n_bins,
                  const unsigned long long *__restrict__ n_bins_dev) {
  const size_t b = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (n_bins_dev) n_bins = static_cast<size_t>(*n_bins_dev);
  if (b >= n_bins) return;
  parent[b] = uf_find_readonly(parent, static_cast<int>(b));
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

// Scale 2^k with n * wmax * 2^k < 2^62.
__device__ __forceinline__ int fixed_point_shift(unsigned long long wmax_bits, size_t n) {
  if (wmax_bits == 0) return 0;
  const int e = static_cast<int>(wmax_bits >> 52) - 1023 + 1;  // wmax < 2^e
  const int ln = 64 - __clzll(static_cast<unsigned long long>(n));  // n < 2^ln
  return 62 - e - ln;
}

__global__ void __launch_bounds__(256)
cluster_weight_kernel(const float *__restrict__ x, const float *__restrict__ y,
                      const float *__restrict__ t, const double *__restrict__ w, size_t n,
                      BinParams bp, BinGrid g, const int32_t *__restrict__ table,
                      const int32_t *__restrict__ label,
                      const int32_t *__restrict__ cluster_of_root,
                      int32_t *__restrict__ cid_out, unsigned long long *__restrict__ cluster_w,
                      EstimateScalars *__restrict__ es, size_t n_total) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int shift = fixed_point_shift(es->wmax_bits, n_total);
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
    return;
  }
  if (leader && cid >= 0 && sum) atomicAdd(cluster_w + cid, sum);
}

// First maximum of the cluster weights (ties -> smallest canonical id).
__global__ void __launch_bounds__(1024)
best_cluster_kernel(const unsigned long long *__restrict__ cluster_w, size_t n_clusters,
                    const unsigned long long *__restrict__ n_clusters_dev,
                    EstimateScalars *__restrict__ es) {
  __shared__ unsigned long long s_w[32];
  __shared__ long long s_i[32];
  // the count may still be on its way to the host (label_clusters, deferred)
  if (n_clusters_dev) n_clusters = static_cast<size_t>(*n_clusters_dev);
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
    bw = s_w[lane];
    bi = s_i[lane];
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
__host__ __device__ void finalise_estimate(const double *b, const double *gsum, double *out) {
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
__global__ void __launch_bounds__(256)
moments_kernel(const float *__restrict__ x, const float *__restrict__ y,
               const float *__restrict__ t, const double *__restrict__ w,
               const int32_t *__restrict__ cid, size_t n, double *__restrict__ partials,
               EstimateScalars *__restrict__ es, double *__restrict__ sums_out,
               double *__restrict__ out, int finalise) {
  __shared__ double smem[8];
  __shared__ bool s_last;
  const int best = es->best;
  double vg[kMom], vb[kMom];
#pragma unroll
  for (int k = 0; k < kMom; ++k) vg[k] = vb[k] = 0.0;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
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
  double *mine = partials + static_cast<size_t>(blockIdx.x) * 2 * kMom;
#pragma unroll
  for (int k = 0; k < kMom; ++k) {
    const double sg = block_sum_256(vg[k], smem);
    const double sb = block_sum_256(vb[k], smem);
    if (threadIdx.x == 0) {
      mine[k] = sg;
      mine[kMom + k] = sb;
    }
  }
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(&es->blocks_done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  __shared__ double s_tot[2 * kMom];
  if (threadIdx.x < 2 * kMom) {
    double a = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b)
      a = __dadd_rn(a, __ldcg(partials + static_cast<size_t>(b) * 2 * kMom + threadIdx.x));
    s_tot[threadIdx.x] = a;
    sums_out[threadIdx.x] = a;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    es->blocks_done = 0;
    if (finalise) finalise_estimate(s_tot + kMom, s_tot, out);
  }
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
  } else if (known_bbox && !sharded(f)) {
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
  QMCL_CUDA(cudaMemsetAsync(f->table.ptr, 0, cells * sizeof(int32_t), s));
  if (n) {
    bin_mark_kernel<<<div_up(n, 256), 256, 0, s>>>(x, y, t, n, bp, g, f->table.ptr);
    QMCL_LAUNCHED();
  }
  size_t n_all = n;
  if (sharded(f)) {
    // Exchange: bin occupancy counts summed over the shards; the occupied-bin
    // list and the labels are then computed identically on every rank.
    QMCL_TRY(shard_allreduce(f, f->table.ptr, cells, QMCL_DT_I32, QMCL_OP_SUM));
    n_all = f->global_n;
  }
  const size_t max_bins = std::min(n_all, cells);
  QMCL_TRY(f->bin_cell.ensure(max_bins));
  QMCL_TRY(f->bin_count.ensure(max_bins));
  QMCL_TRY(f->bin_label.ensure(max_bins));
  QMCL_TRY(f->bin_cluster.ensure(max_bins));
  QMCL_TRY(f->scan_tmp.ensure_zero(scan_tmp_words(cells) + 4));
  unsigned long long *d_total = &f->scalars.ptr->n_bins;  // stable: the labelling reads it
  QMCL_TRY(device_scan(s, OccupiedFlag{f->table.ptr},
                       EmitBin{f->table.ptr, f->bin_cell.ptr, f->bin_count.ptr, f->bin_label.ptr},
                       cells, f->scan_tmp.ptr, d_total));
  if (known_bins >= 0 && !sharded(f)) {
    // the caller counted the bins already (KLD: new-bin flags up to the stop)
    f->n_bins = static_cast<size_t>(known_bins);
  } else {
    // The host needs the count only to size the labelling launches, for which
    // max_bins is a bound: in lazy mode the copy is left in flight.
    QMCL_CUDA(copy_d2h(f->h_scalars.ptr + HS_NBINS, d_total, sizeof(unsigned long long), s));
    f->bins_pending = true;
    f->n_bins_bound = max_bins;
    if (!lazy(f)) QMCL_TRY(settle(f));
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
  StageScope sc(f, QMCL_STAGE_LABEL);
  cc_hook_kernel<<<div_up(nb, 256), 256, 0, s>>>(f->table.ptr, f->bin_cell.ptr, nb, nb_dev, f->grid,
                                                 bp, f->bin_label.ptr);
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
  if (!lazy(f)) QMCL_TRY(settle(f));
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
  const ParticleSet &ps = f->cur();
  QMCL_TRY(build_bins(f, ps.x.ptr, ps.y.ptr, ps.t.ptr, ps.n));
  return label_clusters(f);
}

qmcl_status qmcl_filter_get_estimate(qmcl_filter *f, double pose[3], double cov[9], int *valid) {
  if (!f || !pose || !cov || !valid) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(use_device(f->map));
  cudaStream_t s = f->map->stream;
  *valid = 0;
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
  if (sharded(f)) {
    // Exchange: the fixed-point scale needs the largest weight of all shards.
    std::vector<unsigned long long> all(f->n_shards);
    QMCL_TRY(shard_gather(f, &es->wmax_bits, true, sizeof(unsigned long long), all.data()));
    unsigned long long gmax = 0;
    for (unsigned long long v : all) gmax = std::max(gmax, v);
    QMCL_CUDA(copy_h2d(&es->wmax_bits, &gmax, sizeof(gmax), s));
  }
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
  best_cluster_kernel<<<1, 1024, 0, s>>>(cluster_w, nc_cap,
                                         nc_on_device ? &f->scalars.ptr->n_clusters : nullptr, es);
  QMCL_LAUNCHED();
  moments_kernel<<<blocks, 256, 0, s>>>(ps.x.ptr, ps.y.ptr, ps.t.ptr, ps.w.ptr, f->particle_cid.ptr,
                                        ps.n, base + o_part, es, base + o_sums, d_out,
                                        sharded(f) ? 0 : 1);
  QMCL_LAUNCHED();
  double h[16];
  int flag = 0;
  if (sharded(f)) {
    // Exchange: moments of the best cluster and of all particles, summed in rank order.
    struct Part { double sums[2 * kMom]; double invalid; } mine;
    QMCL_CUDA(copy_d2h(mine.sums, base + o_sums, sizeof(mine.sums), s));
    QMCL_CUDA(copy_d2h(&flag, &es->invalid, sizeof(int), s));
    QMCL_CUDA(cudaStreamSynchronize(s));
    mine.invalid = flag;
    std::vector<Part> parts(f->n_shards);
    QMCL_TRY(shard_gather(f, &mine, false, sizeof(Part), parts.data()));
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
