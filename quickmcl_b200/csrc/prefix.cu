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
//            true sequential double additions (256 adds each)
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

namespace qmcl {

namespace {

constexpr int kChunk = 256;                 // elements per warp-chunk (8 per lane)
constexpr int kLaneItems = 8;
constexpr int kTileChunks = 8;              // chunks per tile (block of 256 threads)
constexpr int kTile = kChunk * kTileChunks; // 2048 elements
constexpr unsigned long long kMantMask = (1ull << 52) - 1;
constexpr unsigned long long kImplicit = 1ull << 52;
constexpr unsigned long long kKLimit = 1ull << 53;
constexpr int kInvalidE = -1;

enum : uint8_t { MODE_FAST = 0, MODE_SLOW = 1, MODE_MIXED = 2 };

// Parity-dependent offset: K -> K + (K even ? a0 : a1).
struct Fn {
  unsigned long long a0, a1;
};
__device__ __forceinline__ Fn fn_identity() { return Fn{0ull, 0ull}; }
// f first, then g.
__device__ __forceinline__ Fn fn_compose(const Fn &f, const Fn &g) {
  Fn h;
  h.a0 = f.a0 + ((f.a0 & 1ull) ? g.a1 : g.a0);
  h.a1 = f.a1 + (((f.a1 + 1ull) & 1ull) ? g.a1 : g.a0);
  return h;
}
__device__ __forceinline__ unsigned long long fn_apply(const Fn &f, unsigned long long K) {
  return K + ((K & 1ull) ? f.a1 : f.a0);
}
__device__ __forceinline__ Fn fn_shfl_up(const Fn &f, int o) {
  Fn r;
  r.a0 = __shfl_up_sync(0xffffffffu, f.a0, o);
  r.a1 = __shfl_up_sync(0xffffffffu, f.a1, o);
  return r;
}
__device__ __forceinline__ Fn fn_shfl(const Fn &f, int src) {
  Fn r;
  r.a0 = __shfl_sync(0xffffffffu, f.a0, src);
  r.a1 = __shfl_sync(0xffffffffu, f.a1, src);
  return r;
}
// Inclusive scan over the lanes of a warp (lane 0 first).
__device__ __forceinline__ Fn fn_warp_scan(Fn v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const Fn prev = fn_shfl_up(v, o);
    if (lane >= o) v = fn_compose(prev, v);
  }
  return v;
}

// Increment of one weight for a running sum in the binade with biased
// exponent Eb.  Returns false when the closed form does not apply.
__device__ __forceinline__ bool elem_fn(double w, int Eb, Fn *out) {
  const long long bits = __double_as_longlong(w);
  *out = fn_identity();
  if ((bits & 0x7fffffffffffffffll) == 0) return true;  // +-0 adds nothing
  if (bits < 0) return false;                            // negative weight
  const int ew = static_cast<int>(bits >> 52);
  if (ew == 0 || ew == 0x7ff) return false;              // subnormal, inf, nan
  const int s = Eb - ew;
  if (s <= 0) return false;                              // w >= 2^E: leaves the binade
  if (s >= 54) return true;                              // w < u/2: absorbed
  const unsigned long long m = (static_cast<unsigned long long>(bits) & kMantMask) | kImplicit;
  const unsigned long long q = m >> s;
  const unsigned long long rem = m & ((1ull << s) - 1ull);
  const unsigned long long half = 1ull << (s - 1);
  if (rem > half) {
    out->a0 = out->a1 = q + 1ull;
  } else if (rem < half) {
    out->a0 = out->a1 = q;
  } else {  // exact tie: round to even
    out->a0 = q + (q & 1ull);
    out->a1 = q + ((q + 1ull) & 1ull);
  }
  return true;
}

__device__ __forceinline__ int biased_exponent_if_normal(double c) {
  const long long bits = __double_as_longlong(c);
  if (bits <= 0) return kInvalidE;
  const int e = static_cast<int>(bits >> 52);
  return (e == 0 || e == 0x7ff) ? kInvalidE : e;
}
__device__ __forceinline__ unsigned long long mantissa_int(double c) {
  return (static_cast<unsigned long long>(__double_as_longlong(c)) & kMantMask) | kImplicit;
}
__device__ __forceinline__ double make_double(int Eb, unsigned long long K) {
  return __longlong_as_double(
      static_cast<long long>((static_cast<unsigned long long>(Eb) << 52) | (K & kMantMask)));
}

__device__ __forceinline__ void load_items(const double *__restrict__ w, size_t base, size_t n,
                                           double v[kLaneItems]) {
  if (base + kLaneItems <= n && (reinterpret_cast<uintptr_t>(w + base) & 15u) == 0) {
    const double2 *p = reinterpret_cast<const double2 *>(w + base);
#pragma unroll
    for (int i = 0; i < kLaneItems / 2; ++i) {
      const double2 t = p[i];
      v[2 * i] = t.x;
      v[2 * i + 1] = t.y;
    }
  } else {
#pragma unroll
    for (int i = 0; i < kLaneItems; ++i) v[i] = (base + i < n) ? w[base + i] : 0.0;
  }
}

// ---- phase 0: plain chunk sums (any order; only the exponent is used)
__global__ void __launch_bounds__(256)
prefix_chunk_sums_kernel(const double *__restrict__ w, size_t n, double *__restrict__ chunk_sum) {
  const int lane = threadIdx.x & 31;
  const size_t chunk = static_cast<size_t>(blockIdx.x) * kTileChunks + (threadIdx.x >> 5);
  const size_t base = chunk * kChunk + static_cast<size_t>(lane) * kLaneItems;
  if (chunk * kChunk >= n) return;
  double v[kLaneItems];
  load_items(w, base, n, v);
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < kLaneItems; ++i) s += v[i];
  s = warp_sum_xor(s);
  if (lane == 0) chunk_sum[chunk] = s;
}

// Exclusive scan of the chunk sums by one block, offset by the carry-in.
__global__ void __launch_bounds__(1024)
prefix_scan_chunk_sums_kernel(const double *__restrict__ chunk_sum, size_t nc, double carry_in,
                              double *__restrict__ approx_in, double *__restrict__ total_out) {
  __shared__ double warp_tot[32];
  __shared__ double chunk_total;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double carry = carry_in;
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

struct PrefixBuffers {
  double *chunk_sum;     // nc
  double *approx_in;     // nc
  unsigned long long *chunk_a0, *chunk_a1;  // nc
  double *chunk_cin;     // nc   (phase 2, MIXED tiles only)
  unsigned long long *tile_a0, *tile_a1;    // nt
  double *tile_cin;      // nt
  int16_t *chunk_E;      // nc   biased exponent guess, -1 = no closed form
  int16_t *tile_E;       // nt
  uint8_t *chunk_mode;   // nc
  uint8_t *tile_mode;    // nt
  double *carry_out;     // 1
};

// ---- phase 1: chunk and tile summaries under the guessed binade
__global__ void __launch_bounds__(256)
prefix_summaries_kernel(const double *__restrict__ w, size_t n, size_t nc, PrefixBuffers pb) {
  __shared__ unsigned long long s_a0[kTileChunks], s_a1[kTileChunks];
  __shared__ int s_E[kTileChunks];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const size_t chunk = static_cast<size_t>(blockIdx.x) * kTileChunks + wid;
  int Eb = kInvalidE;
  Fn f = fn_identity();
  if (chunk < nc) {
    Eb = biased_exponent_if_normal(pb.approx_in[chunk]);
    const size_t base = chunk * kChunk + static_cast<size_t>(lane) * kLaneItems;
    double v[kLaneItems];
    load_items(w, base, n, v);
    bool ok = Eb != kInvalidE;
    if (ok) {
#pragma unroll
      for (int i = 0; i < kLaneItems; ++i) {
        Fn e;
        ok &= elem_fn(v[i], Eb, &e);
        f = fn_compose(f, e);
      }
    }
    if (!__all_sync(0xffffffffu, ok)) Eb = kInvalidE;
    f = fn_warp_scan(f, lane);
    if (lane == 31) {
      pb.chunk_a0[chunk] = f.a0;
      pb.chunk_a1[chunk] = f.a1;
      pb.chunk_E[chunk] = static_cast<int16_t>(Eb);
    }
  }
  if (lane == 31) {
    s_a0[wid] = f.a0;
    s_a1[wid] = f.a1;
    s_E[wid] = (chunk < nc) ? Eb : -2;  // -2: chunk beyond the end (identity)
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    Fn t = fn_identity();
    int E = s_E[0];
    for (int k = 0; k < kTileChunks; ++k) {
      if (s_E[k] == -2) break;
      if (s_E[k] != E) E = kInvalidE;
      t = fn_compose(t, Fn{s_a0[k], s_a1[k]});
    }
    pb.tile_a0[blockIdx.x] = t.a0;
    pb.tile_a1[blockIdx.x] = t.a1;
    pb.tile_E[blockIdx.x] = static_cast<int16_t>(E);
  }
}

// Exact running sums of one chunk from the exact carry `c` (warp-uniform), for
// chunks whose summary cannot be used: the literal sequence of double
// additions.  The warp stages the chunk in shared memory, lane 0 runs the
// dependent chain (bound by the DADD latency, the loads are issued ahead in
// groups of 16), the warp writes the result back coalesced.  (A closed-form
// variant that re-summarised the lanes around the crossing was tried; on one
// latency-bound warp its ~800 instructions per round were slower than these
// 256 additions.)
__device__ double exact_chunk(const double *__restrict__ w, size_t e0, int cnt, double c,
                              double *__restrict__ out, int lane, double *s_w, double *s_o) {
#pragma unroll
  for (int i = 0; i < kChunk / 32; ++i) {
    const int k = lane + 32 * i;
    s_w[k] = (k < cnt) ? w[e0 + k] : 0.0;
  }
  __syncwarp();
  if (lane == 0) {
    double cc = c;
    for (int base = 0; base < kChunk; base += 16) {
      double v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) v[i] = s_w[base + i];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        cc = __dadd_rn(cc, v[i]);
        v[i] = cc;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) s_o[base + i] = v[i];
    }
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < kChunk / 32; ++i) {
    const int k = lane + 32 * i;
    if (k < cnt) out[e0 + k] = s_o[k];
  }
  const double r = s_o[cnt - 1];
  __syncwarp();
  return r;
}

// ---- phase 2: the exact carry through all tiles (one warp)
__global__ void __launch_bounds__(32)
prefix_resolve_kernel(const double *__restrict__ w, size_t n, size_t nc, size_t nt, double carry_in,
                      PrefixBuffers pb, double *__restrict__ out) {
  // chunk summaries of the current batch of 32 tiles, staged on demand
  __shared__ unsigned long long s_a0[32 * kTileChunks], s_a1[32 * kTileChunks];
  __shared__ int16_t s_E[32 * kTileChunks];
  __shared__ double s_w[kChunk], s_o[kChunk];
  const int lane = threadIdx.x;
  double c = carry_in;  // uniform across the warp
  for (size_t base = 0; base < nt; base += 32) {
    const size_t t = base + lane;
    Fn f = fn_identity();
    int E = -2;
    if (t < nt) {
      f.a0 = pb.tile_a0[t];
      f.a1 = pb.tile_a1[t];
      E = pb.tile_E[t];
    }
    const int Ec = biased_exponent_if_normal(c);
    const bool lane_ok = (E == -2) || (E == Ec && Ec != kInvalidE);
    if (__all_sync(0xffffffffu, lane_ok)) {
      const Fn incl = fn_warp_scan(f, lane);
      const Fn total = fn_shfl(incl, 31);
      const unsigned long long K = mantissa_int(c);
      const unsigned long long Kend = fn_apply(total, K);
      if (Kend < kKLimit) {
        Fn excl = fn_shfl_up(incl, 1);
        if (lane == 0) excl = fn_identity();
        if (t < nt) {
          pb.tile_cin[t] = make_double(Ec, fn_apply(excl, K));
          pb.tile_mode[t] = MODE_FAST;
        }
        c = make_double(Ec, Kend);
        continue;
      }
    }
    // Walk these (up to) 32 tiles one by one; stage their chunk summaries first.
    const size_t t_end = (base + 32 < nt) ? base + 32 : nt;
    {
      const size_t cb = base * kTileChunks;
      for (int i = lane; i < 32 * kTileChunks; i += 32) {
        const size_t ch = cb + i;
        if (ch < nc) {
          s_a0[i] = pb.chunk_a0[ch];
          s_a1[i] = pb.chunk_a1[ch];
          s_E[i] = pb.chunk_E[ch];
        }
      }
      __syncwarp();
    }
    for (size_t tt = base; tt < t_end; ++tt) {
      const Fn tf = fn_shfl(f, static_cast<int>(tt - base));
      const int tE = __shfl_sync(0xffffffffu, E, static_cast<int>(tt - base));
      int cE = biased_exponent_if_normal(c);
      if (tE != kInvalidE && tE == cE) {
        const unsigned long long Kend = fn_apply(tf, mantissa_int(c));
        if (Kend < kKLimit) {
          if (lane == 0) {
            pb.tile_cin[tt] = c;
            pb.tile_mode[tt] = MODE_FAST;
          }
          c = make_double(cE, Kend);
          continue;
        }
      }
      if (lane == 0) {
        pb.tile_cin[tt] = c;
        pb.tile_mode[tt] = MODE_MIXED;
      }
      const size_t ch0 = tt * kTileChunks;
      const size_t ch1 = (ch0 + kTileChunks < nc) ? ch0 + kTileChunks : nc;
      for (size_t ch = ch0; ch < ch1; ++ch) {
        const int si = static_cast<int>(ch - base * kTileChunks);
        cE = biased_exponent_if_normal(c);
        const int chE = s_E[si];
        if (chE != kInvalidE && chE == cE) {
          const unsigned long long Kend = fn_apply(Fn{s_a0[si], s_a1[si]}, mantissa_int(c));
          if (Kend < kKLimit) {
            if (lane == 0) {
              pb.chunk_cin[ch] = c;
              pb.chunk_mode[ch] = MODE_FAST;
            }
            c = make_double(cE, Kend);
            continue;
          }
        }
        const size_t e0 = ch * kChunk;
        const int cnt = static_cast<int>((e0 + kChunk <= n) ? kChunk : n - e0);
        if (lane == 0) {
          pb.chunk_mode[ch] = MODE_SLOW;  // written here, phase 3 skips it
          pb.chunk_cin[ch] = c;
        }
        c = exact_chunk(w, e0, cnt, c, out, lane, s_w, s_o);
      }
      __syncwarp();
    }
  }
  if (lane == 0) *pb.carry_out = c;
}

// ---- phase 3: expand the carries into the running sums
__global__ void __launch_bounds__(256)
prefix_apply_kernel(const double *__restrict__ w, size_t n, size_t nc, PrefixBuffers pb,
                    double *__restrict__ out) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const size_t tile = blockIdx.x;
  const size_t chunk = tile * kTileChunks + wid;
  if (chunk >= nc) return;
  int Eb;
  unsigned long long K;
  if (pb.tile_mode[tile] == MODE_FAST) {
    Eb = pb.tile_E[tile];
    K = mantissa_int(pb.tile_cin[tile]);
    Fn lane_cf = fn_identity();
    if (lane < wid) {
      lane_cf.a0 = pb.chunk_a0[tile * kTileChunks + lane];
      lane_cf.a1 = pb.chunk_a1[tile * kTileChunks + lane];
    }
    for (int j = 0; j < wid; ++j) K = fn_apply(fn_shfl(lane_cf, j), K);
  } else {
    if (pb.chunk_mode[chunk] == MODE_SLOW) return;  // written by phase 2
    Eb = pb.chunk_E[chunk];
    K = mantissa_int(pb.chunk_cin[chunk]);
  }
  const size_t base = chunk * kChunk + static_cast<size_t>(lane) * kLaneItems;
  double v[kLaneItems];
  load_items(w, base, n, v);
  Fn e[kLaneItems];
  Fn f = fn_identity();
#pragma unroll
  for (int i = 0; i < kLaneItems; ++i) {
    elem_fn(v[i], Eb, &e[i]);
    f = fn_compose(f, e[i]);
  }
  const Fn incl = fn_warp_scan(f, lane);
  Fn excl = fn_shfl_up(incl, 1);
  if (lane == 0) excl = fn_identity();
  unsigned long long Kl = fn_apply(excl, K);
  double r[kLaneItems];
#pragma unroll
  for (int i = 0; i < kLaneItems; ++i) {
    Kl = fn_apply(e[i], Kl);
    r[i] = make_double(Eb, Kl);
  }
  if (base + kLaneItems <= n && (reinterpret_cast<uintptr_t>(out + base) & 15u) == 0) {
    double2 *p = reinterpret_cast<double2 *>(out + base);
#pragma unroll
    for (int i = 0; i < kLaneItems / 2; ++i) p[i] = make_double2(r[2 * i], r[2 * i + 1]);
  } else {
#pragma unroll
    for (int i = 0; i < kLaneItems; ++i)
      if (base + i < n) out[base + i] = r[i];
  }
}

// Summary of a whole shard: the composition of its tile summaries, usable only
// if every tile was summarised under one and the same binade.  out[0] = binade
// (biased exponent, -1 = no closed form), out[1], out[2] = a0, a1.
__global__ void __launch_bounds__(256)
prefix_shard_summary_kernel(PrefixBuffers pb, size_t nt, unsigned long long *__restrict__ out) {
  __shared__ unsigned long long s_a0[8], s_a1[8];
  __shared__ int s_E[8];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  // thread t composes a contiguous run of tiles, in order
  const size_t per = (nt + blockDim.x - 1) / blockDim.x;
  const size_t t0 = static_cast<size_t>(threadIdx.x) * per;
  const size_t t1 = (t0 + per < nt) ? t0 + per : nt;
  Fn f = fn_identity();
  int E = -2;  // -2: nothing yet
  for (size_t t = t0; t < t1; ++t) {
    const int tE = pb.tile_E[t];
    if (E == -2) E = tE;
    else if (tE != E) E = kInvalidE;
    f = fn_compose(f, Fn{pb.tile_a0[t], pb.tile_a1[t]});
  }
  // combine E across the warp / block: all non-empty must agree
  auto merge_E = [](int a, int b) { return a == -2 ? b : (b == -2 ? a : (a == b ? a : kInvalidE)); };
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) E = merge_E(E, __shfl_xor_sync(0xffffffffu, E, o));
  f = fn_warp_scan(f, lane);
  if (lane == 31) {
    s_a0[wid] = f.a0;
    s_a1[wid] = f.a1;
  }
  if (lane == 0) s_E[wid] = E;
  __syncthreads();
  if (threadIdx.x == 0) {
    Fn tot = fn_identity();
    int Eb = -2;
    for (int k = 0; k < 8; ++k) {
      tot = fn_compose(tot, Fn{s_a0[k], s_a1[k]});
      Eb = merge_E(Eb, s_E[k]);
    }
    out[0] = static_cast<unsigned long long>(static_cast<long long>(Eb < 0 ? -1 : Eb));
    out[1] = tot.a0;
    out[2] = tot.a1;
  }
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

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

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
               o_ca1 = carve(nc * 8), o_ccin = carve(nc * 8), o_ta0 = carve(nt * 8),
               o_ta1 = carve(nt * 8), o_tcin = carve(nt * 8), o_cE = carve(nc * 2),
               o_tE = carve(nt * 2), o_cm = carve(nc), o_tm = carve(nt);
  QMCL_TRY(f->prefix_scratch.ensure(off + 16));
  uint8_t *b = f->prefix_scratch.ptr;
  PrefixBuffers &pb = pl->pb;
  pb.chunk_sum = reinterpret_cast<double *>(b + o_sum);
  pb.approx_in = reinterpret_cast<double *>(b + o_apx);
  pb.chunk_a0 = reinterpret_cast<unsigned long long *>(b + o_ca0);
  pb.chunk_a1 = reinterpret_cast<unsigned long long *>(b + o_ca1);
  pb.chunk_cin = reinterpret_cast<double *>(b + o_ccin);
  pb.tile_a0 = reinterpret_cast<unsigned long long *>(b + o_ta0);
  pb.tile_a1 = reinterpret_cast<unsigned long long *>(b + o_ta1);
  pb.tile_cin = reinterpret_cast<double *>(b + o_tcin);
  pb.chunk_E = reinterpret_cast<int16_t *>(b + o_cE);
  pb.tile_E = reinterpret_cast<int16_t *>(b + o_tE);
  pb.chunk_mode = b + o_cm;
  pb.tile_mode = b + o_tm;
  pb.carry_out = f->prefix_carry.ptr;
  pl->nc = nc;
  pl->nt = nt;
  return QMCL_OK;
}

}  // namespace

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

// After prefix_summaries: the shard's summary {binade, a0, a1} as three 64-bit
// words in dev_out (device memory).
qmcl_status prefix_shard_summary(qmcl_filter *f, size_t n, unsigned long long *dev_out) {
  cudaStream_t s = f->map->stream;
  if (!n) {
    // an empty shard adds nothing under any binade: mark as "identity"
    const unsigned long long ident[3] = {~0ull - 1, 0ull, 0ull};  // binade word -2
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

// Host replay of one summary on an exact carry: true and *carry_out set when
// the carry is in the summary's binade and stays there.
bool prefix_apply_summary(double carry_in, const unsigned long long summary[3], double *carry_out) {
  const long long Eb = static_cast<long long>(summary[0]);
  if (Eb == -2) {  // empty shard
    *carry_out = carry_in;
    return true;
  }
  if (Eb < 0) return false;
  unsigned long long bits;
  std::memcpy(&bits, &carry_in, sizeof(bits));
  if (carry_in <= 0.0 || (bits >> 63)) return false;
  const long long ce = static_cast<long long>(bits >> 52);
  if (ce == 0 || ce == 0x7ff || ce != Eb) return false;
  const unsigned long long K = (bits & kMantMask) | kImplicit;
  const unsigned long long Kend = K + ((K & 1ull) ? summary[2] : summary[1]);
  if (Kend >= kKLimit || Kend < K) return false;
  const unsigned long long ob = (static_cast<unsigned long long>(Eb) << 52) | (Kend & kMantMask);
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
  if (!f || mode < 0 || mode > 1) return QMCL_ERR_INVALID_ARG;
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
