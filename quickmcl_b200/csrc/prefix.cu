// prefix.cu — K9: the CDF exactly as the reference computes it, in parallel.
//
// The reference's running sums (particle_filter.cpp:225-236, particle.cpp:66-82)
// are *sequential* double additions, c_i = fl(c_{i-1} + w_i).  A reassociated
// parallel sum rounds differently and a resampling index flips whenever a draw
// falls between the two results (expected ~100 flips at 16 M particles,
// SURVEY §7), so the north star's "indices bit-exact given identical weights"
// needs the sequentially rounded values.  A literal sequential loop costs
// ~12 ns per particle on one warp (1.3 ms at 100 k, 200 ms at 16 M).
//
// This file computes the same bits in parallel.  While the running sum stays
// inside one binade [2^E, 2^(E+1)) it is an integer K in [2^52, 2^53) times
// u = 2^(E-52), and fl(c + w) = (K + round_to_nearest_even(w / u)) * u.  The
// increment of one element is therefore an integer that depends on K only
// through its parity, and only when w/u lies exactly half-way between two
// integers (round-half-even).  Such "parity-dependent offsets"
//     F(K) = K + (K even ? a0 : a1)
// are closed under composition, so chunks of elements can be summarised and
// combined with an ordinary parallel scan:
//
//   phase 0  per-chunk plain sums -> approximate carry-in per chunk; its
//            exponent is the chunk's binade *guess* E
//   phase 1  per-chunk (256 elements, one warp) and per-tile (8 chunks)
//            summaries (a0, a1) under that guess
//   phase 2  one warp walks the tiles in order with the exact carry: a summary
//            is used only if the carry really is in binade E and K stays below
//            2^53 (weights are >= 0, so the sum is monotone and the end points
//            bound everything in between); 32 tiles are composed at a time.
//            Chunks that fail the check — the ones in which the sum crosses
//            into the next binade, the very first chunk, and anything with a
//            negative / subnormal / non-finite weight — are evaluated with
//            true sequential double additions (256 adds each), except that a
//            crossing predicted by phase 1 from the approximate carry (which
//            element crosses) is verified with the exact carry and then costs
//            one addition (prefix_dev.cuh)
//   phase 3  every fast chunk expands its carry-in into the 256 running sums
//            with integer arithmetic and writes them as doubles
//
// The result is bit-identical to the sequential loop for every input (checked
// against the one-warp sequential kernel below and numpy.cumsum in
// tests/test_prefix_gpu.py, including ties, zeros, 600 binades of dynamic
// range and invalid inputs).
//
// Sharding: `carry_in` is the exact running sum before this shard's first
// particle and `*carry_out` the one after its last, so shards chain through
// one double (quickmcl_b200/sharded.py).

#include "prefix.cuh"
#include "prefix_dev.cuh"

namespace qmcl {

namespace {

using namespace pfx;

// Exclusive scan of the chunk sums by one block, offset by the carry-in.
__global__ void __launch_bounds__(1024)
prefix_scan_chunk_sums_kernel(const double *__restrict__ chunk_sum, size_t nc, double carry_in,
                              double *__restrict__ approx_in, double *__restrict__ total_out,
                              const double *__restrict__ carry_terms = nullptr, int n_terms = 0,
                              int term_stride = 0) {
  __shared__ double warp_tot[32];
  __shared__ double chunk_total;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double carry = carry_in;
  // sharded filters: the approximate totals of the shards before this one, as they
  // came out of the exchange (device memory), added in rank order
  for (int r = 0; r < n_terms; ++r) carry += carry_terms[static_cast<size_t>(r) * term_stride];
  for (size_t base = 0; base < nc; base += 1024) {
    const size_t i = base + threadIdx.x;
    const double v = (i < nc) ? chunk_sum[i] : 0.0;
    double incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) warp_tot[wid] = incl;
    __syncthreads();
    if (wid == 0) {
      const double t = warp_tot[lane];
      double ti = t;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double up = __shfl_up_sync(0xffffffffu, ti, o);
        if (lane >= o) ti += up;
      }
      warp_tot[lane] = ti - t;
      if (lane == 31) chunk_total = ti;
    }
    __syncthreads();
    if (i < nc) approx_in[i] = carry + warp_tot[wid] + (incl - v);
    carry += chunk_total;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(256)
prefix_chunk_sums_kernel(const double *__restrict__ w, size_t n, double *__restrict__ chunk_sum) {
  if (static_cast<size_t>(blockIdx.x) * kTile >= n) return;
  prefix_chunk_sums_tile(w, n, chunk_sum, blockIdx.x);
}

__global__ void __launch_bounds__(256)
prefix_summaries_kernel(const double *__restrict__ w, size_t n, size_t nc, PrefixBuffers pb) {
  prefix_summaries_tile(w, n, nc, pb, blockIdx.x);
}

__global__ void __launch_bounds__(32)
prefix_resolve_kernel(const double *__restrict__ w, size_t n, size_t nc, size_t nt, double carry_in,
                      PrefixBuffers pb, double *__restrict__ out) {
  __shared__ ResolveScratch scratch;
  prefix_resolve_warp(w, n, nc, nt, carry_in, pb, out, threadIdx.x, &scratch);
}

__global__ void __launch_bounds__(256)
prefix_apply_kernel(const double *__restrict__ w, size_t n, size_t nc, PrefixBuffers pb,
                    double *__restrict__ out) {
  prefix_apply_tile(w, n, nc, pb, out, blockIdx.x);
}

// Summary of a whole shard: the composition of its tile summaries, in the same form
// as a tile's: {E, a} when every tile was summarised under one binade, or - with
// one predicted crossing in one of the tiles - {E, a, wx, b}: a up to the crossing
// element, the literal addition of wx (which must take the sum into binade E + 1),
// b over everything after it (tiles summarised under E + 1).  Two crossings, or a
// tile without a summary: no closed form.
//   out[0] = binade (biased exponent; -1 = no closed form, -2 = empty shard)
//   out[1], out[2] = a0, a1     out[3] = 1 with a crossing
//   out[4] = bits of wx         out[5], out[6] = b0, b1
struct ShardPiece {
  int E;    // binade at the start; -2 = identity (nothing yet), -1 = no closed form
  int x;    // 1: holds a crossing
  Fn a, b;
  double wx;
};
__device__ __forceinline__ ShardPiece piece_identity() {
  return ShardPiece{-2, 0, fn_identity(), fn_identity(), 0.0};
}
// p first, then q
__device__ __forceinline__ ShardPiece piece_compose(const ShardPiece &p, const ShardPiece &q) {
  if (p.E == -2) return q;
  if (q.E == -2) return p;
  ShardPiece r = p;
  if (p.E == kInvalidE || q.E == kInvalidE) {
    r.E = kInvalidE;
  } else if (!p.x) {
    if (q.E != p.E) {
      r.E = kInvalidE;
    } else {
      r.a = fn_compose(p.a, q.a);
      r.x = q.x;
      r.wx = q.wx;
      r.b = q.b;
    }
  } else if (q.x || q.E != p.E + 1) {
    r.E = kInvalidE;
  } else {
    r.b = fn_compose(p.b, q.a);
  }
  return r;
}
__global__ void __launch_bounds__(256)
prefix_shard_summary_kernel(PrefixBuffers pb, size_t nt, unsigned long long *__restrict__ out) {
  __shared__ ShardPiece s_piece[256];
  // thread t composes a contiguous run of tiles, in order; thread 0 then composes the runs
  const size_t per = (nt + blockDim.x - 1) / blockDim.x;
  const size_t t0 = static_cast<size_t>(threadIdx.x) * per;
  const size_t t1 = (t0 + per < nt) ? t0 + per : nt;
  ShardPiece acc = piece_identity();
  for (size_t t = t0; t < t1; ++t) {
    ShardPiece q;
    q.E = pb.tile_E[t];
    q.x = pb.tile_jx[t] >= 0 ? 1 : 0;
    q.a = Fn{pb.tile_a0[t], pb.tile_a1[t]};
    q.b = Fn{pb.tile_b0[t], pb.tile_b1[t]};
    q.wx = pb.tile_wx[t];
    acc = piece_compose(acc, q);
  }
  s_piece[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    ShardPiece tot = piece_identity();
    for (int k = 0; k < 256; ++k) tot = piece_compose(tot, s_piece[k]);
    out[0] = static_cast<unsigned long long>(static_cast<long long>(tot.E == -2 ? -2 : (tot.E < 0 ? -1 : tot.E)));
    out[1] = tot.a.a0;
    out[2] = tot.a.a1;
    out[3] = tot.x ? 1ull : 0ull;
    out[4] = static_cast<unsigned long long>(__double_as_longlong(tot.wx));
    out[5] = tot.b.a0;
    out[6] = tot.b.a1;
  }
}
// A shard whose exact carry-out is known already (the first one: its carry-in is 0):
// out[0] = -3, out[1] = bits of the carry-out.
__global__ void prefix_exact_summary_kernel(const double *__restrict__ carry_out,
                                            unsigned long long *__restrict__ out) {
  out[0] = static_cast<unsigned long long>(-3ll);
  out[1] = static_cast<unsigned long long>(__double_as_longlong(*carry_out));
  for (int k = 2; k < 7; ++k) out[k] = 0ull;
}

// The literal loop on one warp: coalesced loads of 32 weights, lane-uniform
// dependent additions.  Kept as the in-library cross-check of the parallel
// path (qmcl_filter_set_prefix_mode(f, 1)).
__global__ void __launch_bounds__(32)
sequential_prefix_kernel(const double *__restrict__ w, size_t n, double carry_in,
                         double *__restrict__ out, double *__restrict__ carry_out) {
  const int lane = threadIdx.x;
  double c = carry_in;
  for (size_t base = 0; base < n; base += 32) {
    const size_t i = base + lane;
    const double v = (i < n) ? w[i] : 0.0;
    double mine = 0.0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      c = __dadd_rn(c, __shfl_sync(0xffffffffu, v, k));
      if (lane == k) mine = c;
    }
    if (i < n) out[i] = mine;
  }
  if (lane == 0) *carry_out = c;
}

__global__ void set_double_kernel(double *p, double v) { *p = v; }

}  // namespace

namespace {

struct PrefixPlan {
  size_t nc, nt;
  PrefixBuffers pb;
};

qmcl_status plan_prefix(qmcl_filter *f, size_t n, PrefixPlan *pl) {
  QMCL_TRY(f->prefix_carry.ensure(2));
  const size_t nc = (n + kChunk - 1) / kChunk;
  const size_t nt = (n + kTile - 1) / kTile;
  size_t off = 0;
  auto carve = [&](size_t bytes) {
    const size_t at = off;
    off = align_up(off + bytes, 16);
    return at;
  };
  const size_t o_sum = carve(nc * 8), o_apx = carve(nc * 8), o_ca0 = carve(nc * 8),
               o_ca1 = carve(nc * 8), o_cb0 = carve(nc * 8), o_cb1 = carve(nc * 8),
               o_cwx = carve(nc * 8), o_ccin = carve(nc * 8), o_ta0 = carve(nt * 8),
               o_ta1 = carve(nt * 8), o_tb0 = carve(nt * 8), o_tb1 = carve(nt * 8),
               o_twx = carve(nt * 8), o_tcin = carve(nt * 8), o_cE = carve(nc * 2),
               o_cjx = carve(nc * 2), o_tE = carve(nt * 2), o_tjx = carve(nt * 2), o_cm = carve(nc),
               o_tm = carve(nt);
  QMCL_TRY(f->prefix_scratch.ensure(off + 16));
  uint8_t *b = f->prefix_scratch.ptr;
  PrefixBuffers &pb = pl->pb;
  pb.chunk_sum = reinterpret_cast<double *>(b + o_sum);
  pb.approx_in = reinterpret_cast<double *>(b + o_apx);
  pb.chunk_a0 = reinterpret_cast<unsigned long long *>(b + o_ca0);
  pb.chunk_a1 = reinterpret_cast<unsigned long long *>(b + o_ca1);
  pb.chunk_b0 = reinterpret_cast<unsigned long long *>(b + o_cb0);
  pb.chunk_b1 = reinterpret_cast<unsigned long long *>(b + o_cb1);
  pb.chunk_wx = reinterpret_cast<double *>(b + o_cwx);
  pb.chunk_cin = reinterpret_cast<double *>(b + o_ccin);
  pb.tile_a0 = reinterpret_cast<unsigned long long *>(b + o_ta0);
  pb.tile_a1 = reinterpret_cast<unsigned long long *>(b + o_ta1);
  pb.tile_b0 = reinterpret_cast<unsigned long long *>(b + o_tb0);
  pb.tile_b1 = reinterpret_cast<unsigned long long *>(b + o_tb1);
  pb.tile_wx = reinterpret_cast<double *>(b + o_twx);
  pb.tile_cin = reinterpret_cast<double *>(b + o_tcin);
  pb.chunk_E = reinterpret_cast<int16_t *>(b + o_cE);
  pb.chunk_jx = reinterpret_cast<int16_t *>(b + o_cjx);
  pb.tile_E = reinterpret_cast<int16_t *>(b + o_tE);
  pb.tile_jx = reinterpret_cast<int16_t *>(b + o_tjx);
  pb.chunk_mode = b + o_cm;
  pb.tile_mode = b + o_tm;
  pb.carry_out = f->prefix_carry.ptr;
  // prefix mode 2 = no predicted crossings: every binade-crossing chunk is added literally
  // (bit-identical; kept for A/B and as a cross-check of the speculation).
  pb.speculate = f->prefix_mode == 2 ? 0 : 1;
  pl->nc = nc;
  pl->nt = nt;
  return QMCL_OK;
}

}  // namespace

// The scratch layout of a running sum over n elements, for the fused tail.
qmcl_status plan_prefix_buffers(qmcl_filter *f, size_t n, pfx::PrefixBuffers *pb,
                                unsigned long long *nc, unsigned long long *nt) {
  PrefixPlan pl;
  QMCL_TRY(plan_prefix(f, n, &pl));
  *pb = pl.pb;
  *nc = pl.nc;
  *nt = pl.nt;
  return QMCL_OK;
}

// Phase 0: chunk sums; the shard's approximate total is left in prefix_carry[1].
qmcl_status prefix_local_total(qmcl_filter *f, const double *w, size_t n) {
  cudaStream_t s = f->map->stream;
  PrefixPlan pl;
  QMCL_TRY(plan_prefix(f, n, &pl));
  if (!n) {
    set_double_kernel<<<1, 1, 0, s>>>(f->prefix_carry.ptr + 1, 0.0);
    QMCL_LAUNCHED();
    return QMCL_OK;
  }
  prefix_chunk_sums_kernel<<<static_cast<unsigned>(pl.nt), 256, 0, s>>>(w, n, pl.pb.chunk_sum);
  QMCL_LAUNCHED();
  prefix_scan_chunk_sums_kernel<<<1, 1024, 0, s>>>(pl.pb.chunk_sum, pl.nc, 0.0, pl.pb.approx_in,
                                                  f->prefix_carry.ptr + 1);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

// Phase 1: binade guesses from the approximate carry-in, chunk/tile summaries.
qmcl_status prefix_summaries(qmcl_filter *f, const double *w, size_t n, double approx_carry_in) {
  if (!n) return QMCL_OK;
  cudaStream_t s = f->map->stream;
  PrefixPlan pl;
  QMCL_TRY(plan_prefix(f, n, &pl));
  prefix_scan_chunk_sums_kernel<<<1, 1024, 0, s>>>(pl.pb.chunk_sum, pl.nc, approx_carry_in,
                                                  pl.pb.approx_in, nullptr);
  QMCL_LAUNCHED();
  prefix_summaries_kernel<<<static_cast<unsigned>(pl.nt), 256, 0, s>>>(w, n, pl.nc, pl.pb);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

// The same with the approximate carry-in still on the device: the sum of
// terms[0], terms[stride], ... (n_terms of them), e.g. the gathered shard totals.
qmcl_status prefix_summaries_dev(qmcl_filter *f, const double *w, size_t n, const double *terms,
                                 int n_terms, int stride) {
  if (!n) return QMCL_OK;
  cudaStream_t s = f->map->stream;
  PrefixPlan pl;
  QMCL_TRY(plan_prefix(f, n, &pl));
  prefix_scan_chunk_sums_kernel<<<1, 1024, 0, s>>>(pl.pb.chunk_sum, pl.nc, 0.0, pl.pb.approx_in, nullptr,
                                                  terms, n_terms, stride);
  QMCL_LAUNCHED();
  prefix_summaries_kernel<<<static_cast<unsigned>(pl.nt), 256, 0, s>>>(w, n, pl.nc, pl.pb);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

// After prefix_summaries: the shard's summary as kShardSummaryWords 64-bit words in
// dev_out (device memory); layout at prefix_shard_summary_kernel.
qmcl_status prefix_shard_summary(qmcl_filter *f, size_t n, unsigned long long *dev_out) {
  cudaStream_t s = f->map->stream;
  if (!n) {
    // an empty shard adds nothing under any binade: mark as "identity"
    const unsigned long long ident[kShardSummaryWords] = {~0ull - 1, 0ull, 0ull, 0ull, 0ull, 0ull, 0ull};  // binade word -2
    QMCL_CUDA(cudaMemcpyAsync(dev_out, ident, sizeof(ident), cudaMemcpyHostToDevice, s));
    QMCL_CUDA(cudaStreamSynchronize(s));
    return QMCL_OK;
  }
  PrefixPlan pl;
  QMCL_TRY(plan_prefix(f, n, &pl));
  prefix_shard_summary_kernel<<<1, 256, 0, s>>>(pl.pb, pl.nt, dev_out);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

// After prefix_resolve: the "summary" of a shard whose exact carry-out is known.
qmcl_status prefix_exact_summary(qmcl_filter *f, unsigned long long *dev_out) {
  prefix_exact_summary_kernel<<<1, 1, 0, f->map->stream>>>(f->prefix_carry.ptr, dev_out);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

// Host replay of one summary on an exact carry (apply_summary of prefix_dev.cuh in
// host arithmetic: the same integer operations, and one IEEE double addition for a
// crossing): true and *carry_out set when the carry is in the summary's binade and
// the summary's conditions hold.
bool prefix_apply_summary(double carry_in, const unsigned long long summary[kShardSummaryWords],
                          double *carry_out) {
  const long long Eb = static_cast<long long>(summary[0]);
  if (Eb == -2) {  // empty shard
    *carry_out = carry_in;
    return true;
  }
  if (Eb == -3) {  // resolved by its owner already (the first shard, carry-in 0)
    if (carry_in != 0.0) return false;
    std::memcpy(carry_out, &summary[1], sizeof(double));
    return true;
  }
  if (Eb < 0) return false;
  unsigned long long bits;
  std::memcpy(&bits, &carry_in, sizeof(bits));
  if (carry_in <= 0.0 || (bits >> 63)) return false;
  const long long ce = static_cast<long long>(bits >> 52);
  if (ce == 0 || ce == 0x7ff || ce != Eb) return false;
  const unsigned long long K = (bits & kMantMask) | kImplicit;
  const unsigned long long K1 = K + ((K & 1ull) ? summary[2] : summary[1]);
  if (K1 >= kKLimit || K1 < K) return false;
  const unsigned long long b1 = (static_cast<unsigned long long>(Eb) << 52) | (K1 & kMantMask);
  if (!summary[3]) {
    std::memcpy(carry_out, &b1, sizeof(b1));
    return true;
  }
  // the crossing: one literal addition, which must land exactly one binade up
  double c1, wx;
  std::memcpy(&c1, &b1, sizeof(c1));
  std::memcpy(&wx, &summary[4], sizeof(wx));
  volatile double c2v = c1 + wx;  // a rounded IEEE addition (never contracted, never extended)
  const double c2 = c2v;
  unsigned long long b2;
  std::memcpy(&b2, &c2, sizeof(b2));
  if ((b2 >> 63) || static_cast<long long>(b2 >> 52) != Eb + 1 || Eb + 1 >= 0x7ff) return false;
  const unsigned long long K2 = (b2 & kMantMask) | kImplicit;
  const unsigned long long K3 = K2 + ((K2 & 1ull) ? summary[6] : summary[5]);
  if (K3 >= kKLimit || K3 < K2) return false;
  const unsigned long long ob = (static_cast<unsigned long long>(Eb + 1) << 52) | (K3 & kMantMask);
  std::memcpy(carry_out, &ob, sizeof(ob));
  return true;
}

// Phase 2: exact carry through the shard; carry_out -> prefix_carry[0].
qmcl_status prefix_resolve(qmcl_filter *f, const double *w, size_t n, double *out,
                           double carry_in) {
  cudaStream_t s = f->map->stream;
  PrefixPlan pl;
  QMCL_TRY(plan_prefix(f, n, &pl));
  if (!n) {
    set_double_kernel<<<1, 1, 0, s>>>(f->prefix_carry.ptr, carry_in);
    QMCL_LAUNCHED();
    return QMCL_OK;
  }
  prefix_resolve_kernel<<<1, 32, 0, s>>>(w, n, pl.nc, pl.nt, carry_in, pl.pb, out);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

// Phase 3: expand.
qmcl_status prefix_apply(qmcl_filter *f, const double *w, size_t n, double *out) {
  if (!n) return QMCL_OK;
  PrefixPlan pl;
  QMCL_TRY(plan_prefix(f, n, &pl));
  prefix_apply_kernel<<<static_cast<unsigned>(pl.nt), 256, 0, f->map->stream>>>(w, n, pl.nc, pl.pb,
                                                                               out);
  QMCL_LAUNCHED();
  return QMCL_OK;
}

qmcl_status sequential_prefix(qmcl_filter *f, const double *w, size_t n, double *out,
                              double carry_in) {
  cudaStream_t s = f->map->stream;
  QMCL_TRY(f->prefix_carry.ensure(2));
  if (f->prefix_mode == 1 && n) {
    sequential_prefix_kernel<<<1, 32, 0, s>>>(w, n, carry_in, out, f->prefix_carry.ptr);
    QMCL_LAUNCHED();
    return QMCL_OK;
  }
  if (n) {
    PrefixPlan pl;
    QMCL_TRY(plan_prefix(f, n, &pl));
    prefix_chunk_sums_kernel<<<static_cast<unsigned>(pl.nt), 256, 0, s>>>(w, n, pl.pb.chunk_sum);
    QMCL_LAUNCHED();
  }
  QMCL_TRY(prefix_summaries(f, w, n, carry_in));
  QMCL_TRY(prefix_resolve(f, w, n, out, carry_in));
  return prefix_apply(f, w, n, out);
}

}  // namespace qmcl

using namespace qmcl;

extern "C" {

qmcl_status qmcl_filter_set_prefix_mode(qmcl_filter *f, int mode) {
  if (!f || mode < 0 || mode > 2) return QMCL_ERR_INVALID_ARG;
  f->prefix_mode = mode;
  return QMCL_OK;
}

// Test hook: the running sum of an arbitrary host array.
qmcl_status qmcl_filter_debug_prefix(qmcl_filter *f, const double *w, size_t n, double carry_in,
                                     double *out, double *carry_out) {
  if (!f || (!w && n) || (!out && n)) return QMCL_ERR_INVALID_ARG;
  QMCL_TRY(use_device(f->map));
  cudaStream_t s = f->map->stream;
  void *stage = nullptr;
  QMCL_TRY(filter_scratch(f, 2 * n * sizeof(double) + 64, &stage));
  double *d_w = static_cast<double *>(stage);
  double *d_o = d_w + n + (n & 1);
  if (n) QMCL_CUDA(copy_h2d(d_w, w, n * sizeof(double), s));
  QMCL_TRY(sequential_prefix(f, d_w, n, d_o, carry_in));
  if (n) QMCL_CUDA(copy_d2h(out, d_o, n * sizeof(double), s));
  double co = carry_in;
  if (n) QMCL_CUDA(copy_d2h(&co, f->prefix_carry.ptr, sizeof(double), s));
  QMCL_CUDA(cudaStreamSynchronize(s));
  if (carry_out) *carry_out = co;
  return QMCL_OK;
}

}  // extern "C"
