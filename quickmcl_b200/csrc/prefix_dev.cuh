// prefix_dev.cuh — device side of K9 (see prefix.cu for the algorithm): the
// parity-dependent-offset algebra and the per-tile bodies of the four phases,
// shared by the stand-alone kernels (prefix.cu) and the fused tail (tail.cu).
#pragma once

#include "common.cuh"

namespace qmcl {
namespace pfx {

constexpr int kChunk = 256;                 // elements per warp-chunk (8 per lane)
constexpr int kLaneItems = 8;
constexpr int kTileChunks = 8;              // chunks per tile (block of 256 threads)
constexpr int kTile = kChunk * kTileChunks; // 2048 elements
constexpr unsigned long long kMantMask = (1ull << 52) - 1;
constexpr unsigned long long kImplicit = 1ull << 52;
constexpr unsigned long long kKLimit = 1ull << 53;
constexpr int kInvalidE = -1;

enum : uint8_t { MODE_FAST = 0, MODE_SLOW = 1, MODE_MIXED = 2, MODE_CROSS = 3 };

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
#ifdef QMCL_ELEM_FN_INT
// Integer form (round 1): shifts of the mantissa.  ~45 instructions.
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
#else
// The same in double arithmetic, a third of the instructions.  x = w / ulp is
// exact (a power-of-two scaling; x < 2^52).  K + x rounded to nearest-even is
// K + rint(x) unless x lies exactly half-way between two integers; then an even
// K takes the even neighbour of x, which is rint(x), and an odd K the other one.
// x - rint(x) is exact (at most 1/2 in magnitude and a multiple of ulp(x)), so
// the tie test is exact too.  Binades below 2^-971 (Eb < 52: the scale factor
// would overflow) get no closed form.  Checked against the integer form on 2e7
// random / tie / boundary cases (tools/k9_model.cpp).
__device__ __forceinline__ bool elem_fn(double w, int Eb, Fn *out) {
  *out = fn_identity();
  if (w == 0.0) return true;  // +-0 adds nothing
  const double top = __longlong_as_double(static_cast<long long>(Eb) << 52);  // 2^E: w >= it leaves the binade
  // negative, subnormal, inf, nan all fail one of the two comparisons
  if (!(w >= 2.2250738585072014e-308 && w < top) || Eb < 52) return false;
  const double scale = __longlong_as_double(static_cast<long long>(2098 - Eb) << 52);  // 2^(52 - E)
  const double x = __dmul_rn(w, scale);
  const double r = rint(x);
  const double d = __dadd_rn(x, -r);
  const unsigned long long a0 = static_cast<unsigned long long>(__double2ll_rn(r));
  unsigned long long a1 = a0;
  if (fabs(d) == 0.5) a1 = d > 0.0 ? a0 + 1ull : a0 - 1ull;
  out->a0 = a0;
  out->a1 = a1;
  return true;
}
#endif

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

__device__ __forceinline__ void load_items(const double *w, size_t base, size_t n,
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
// Body of one tile (a 256-thread block).  The bodies take plain pointers: the
// fused tail kernel (tail.cu) calls them on data written earlier in the same
// launch, which must not travel through the non-coherent load path.
__device__ __forceinline__ void prefix_chunk_sums_tile(const double *w, size_t n, double *chunk_sum,
                                                       size_t tile) {
  const int lane = threadIdx.x & 31;
  const size_t chunk = tile * kTileChunks + (threadIdx.x >> 5);
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

// A summary is {E, a} — "in binade E the integer sum grows by a(parity)" — or,
// with a predicted crossing (jx >= 0), {E, a, wx, b}: a over the elements before
// element jx, then the literal addition of wx, which must take the sum into
// binade E + 1, then b over the elements after it.  Tiles: jx is the index of
// the crossing CHUNK inside the tile.
struct PrefixBuffers {
  double *chunk_sum;     // nc
  double *approx_in;     // nc
  unsigned long long *chunk_a0, *chunk_a1, *chunk_b0, *chunk_b1;  // nc
  double *chunk_wx;      // nc
  double *chunk_cin;     // nc   (phase 2, MIXED tiles only)
  unsigned long long *tile_a0, *tile_a1, *tile_b0, *tile_b1;      // nt
  double *tile_wx;       // nt
  double *tile_cin;      // nt
  int16_t *chunk_E;      // nc   biased exponent guess, -1 = no closed form
  int16_t *chunk_jx;     // nc   predicted crossing element (index inside the chunk), -1 = none
  int16_t *tile_E;       // nt
  int16_t *tile_jx;      // nt   chunk (index inside the tile) with the predicted crossing, -1 = none
  uint8_t *chunk_mode;   // nc
  uint8_t *tile_mode;    // nt
  double *carry_out;     // 1
  int speculate;         // 1: predicted crossings are summarised; 0: such chunks are added literally
};

// Shared memory of the resolving warp: one chunk of weights / sums for the literal
// additions.  The fused tail lends it from a buffer of a later phase.
struct ResolveScratch {
  double w[kChunk], o[kChunk];
  // the summaries of the 32 tiles being walked
  unsigned long long ta0[33], ta1[33], tb0[32], tb1[32];  // 33: the walk fetches one tile ahead
  double twx[32];
  int tE[33], tjx[33];
};

// One summary applied to the exact carry c (binade cE, valid).  True and *c_out
// set when it holds: the sum is inside the binade at the end of `a` — weights
// are >= 0 wherever a summary exists, so that bounds everything before — and,
// with a crossing, right after wx it is one binade up and stays there through b.
__device__ __forceinline__ bool apply_summary(double c, int cE, const Fn &a, int jx, double wx, const Fn &b,
                                              double *c_out) {
  const unsigned long long K1 = fn_apply(a, mantissa_int(c));
  if (K1 >= kKLimit) return false;
  if (jx < 0) {
    *c_out = make_double(cE, K1);
    return true;
  }
  const double c2 = __dadd_rn(make_double(cE, K1), wx);
  if (biased_exponent_if_normal(c2) != cE + 1) return false;
  const unsigned long long K3 = fn_apply(b, mantissa_int(c2));
  if (K3 >= kKLimit) return false;
  *c_out = make_double(cE + 1, K3);
  return true;
}

// ---- phase 1: chunk and tile summaries under the guessed binade
// The chunk in which the running sum crosses into the next binade (about
// log2(N) chunks of a normalised weight array) has no summary of the plain
// kind: where it crosses depends on the exact carry.  But the approximate
// carry-in is good to a few thousand ulps while one weight is worth ~2^53 / N
// of them, so the crossing ELEMENT can be predicted, and phase 2 checks the
// prediction with the exact carry before using it (apply_summary).  A wrong
// prediction only costs the literal additions of that chunk.
QMCL_PHASE_FN void prefix_summaries_tile(const double *w, size_t n, size_t nc,
                                                      const PrefixBuffers &pb, size_t tile) {
  __shared__ unsigned long long s_a0[kTileChunks], s_a1[kTileChunks], s_b0[kTileChunks], s_b1[kTileChunks];
  __shared__ double s_wx[kTileChunks];
  __shared__ int s_E[kTileChunks], s_jx[kTileChunks];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const size_t chunk = tile * kTileChunks + wid;
  int Eb = kInvalidE, jx = -1;
  Fn f = fn_identity(), tail = fn_identity();
  double wx = 0.0;
  if (chunk < nc) {
    const double approx = pb.approx_in[chunk];
    Eb = biased_exponent_if_normal(approx);
    const size_t base = chunk * kChunk + static_cast<size_t>(lane) * kLaneItems;
    double v[kLaneItems];
    load_items(w, base, n, v);
    Fn e[kLaneItems];
    bool ok = Eb != kInvalidE;
    if (ok) {
#pragma unroll
      for (int i = 0; i < kLaneItems; ++i) {
        ok &= elem_fn(v[i], Eb, &e[i]);
        f = fn_compose(f, e[i]);
      }
    }
    if (!__all_sync(0xffffffffu, ok)) Eb = kInvalidE;
    f = fn_warp_scan(f, lane);
    if (Eb != kInvalidE && Eb < 0x7fd && pb.speculate) {
      const unsigned long long K = mantissa_int(approx);  // approximate, like the carry it stands for
      const unsigned int over = __ballot_sync(0xffffffffu, fn_apply(f, K) >= kKLimit);
      if (over) {
        // the lane, then the element, at which the approximate sum leaves the binade
        const int L = __ffs(over) - 1;
        Fn excl = fn_shfl_up(f, 1);
        if (lane == 0) excl = fn_identity();
        unsigned long long k = fn_apply(excl, K);
        int ix = kLaneItems;
        double mine = 0.0;
#pragma unroll
        for (int i = 0; i < kLaneItems; ++i) {
          k = fn_apply(e[i], k);
          if (ix == kLaneItems && k >= kKLimit) {
            ix = i;
            mine = v[i];
          }
        }
        ix = __shfl_sync(0xffffffffu, ix, L);
        wx = __shfl_sync(0xffffffffu, mine, L);
        jx = L * kLaneItems + ix;
        Fn head = fn_identity();
#pragma unroll
        for (int i = 0; i < kLaneItems; ++i) {
          const int idx = lane * kLaneItems + i;
          if (idx < jx) {
            head = fn_compose(head, e[i]);
          } else if (idx > jx) {
            Fn e1;
            elem_fn(v[i], Eb + 1, &e1);  // valid under E, hence under E + 1
            tail = fn_compose(tail, e1);
          }
        }
        f = fn_warp_scan(head, lane);
        tail = fn_warp_scan(tail, lane);
      }
    }
    if (lane == 31) {
      pb.chunk_a0[chunk] = f.a0;
      pb.chunk_a1[chunk] = f.a1;
      pb.chunk_b0[chunk] = tail.a0;
      pb.chunk_b1[chunk] = tail.a1;
      pb.chunk_wx[chunk] = wx;
      pb.chunk_E[chunk] = static_cast<int16_t>(Eb);
      pb.chunk_jx[chunk] = static_cast<int16_t>(jx);
    }
  }
  if (lane == 31) {
    s_a0[wid] = f.a0;
    s_a1[wid] = f.a1;
    s_b0[wid] = tail.a0;
    s_b1[wid] = tail.a1;
    s_wx[wid] = wx;
    s_E[wid] = (chunk < nc) ? Eb : -2;  // -2: chunk beyond the end (identity)
    s_jx[wid] = jx;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // the tile's summary: its chunks' in order; at most one crossing, after which the
    // guesses must be one binade up
    Fn a = fn_identity(), b = fn_identity();
    int E = s_E[0], kx = -1;
    double twx = 0.0;
    for (int k = 0; k < kTileChunks && E != kInvalidE; ++k) {
      if (s_E[k] == -2) break;
      const Fn ca{s_a0[k], s_a1[k]};
      if (kx < 0) {
        if (s_E[k] != E) {
          E = kInvalidE;
        } else if (s_jx[k] < 0) {
          a = fn_compose(a, ca);
        } else {
          a = fn_compose(a, ca);
          twx = s_wx[k];
          b = Fn{s_b0[k], s_b1[k]};
          kx = k;
        }
      } else if (s_E[k] != E + 1 || s_jx[k] >= 0) {
        E = kInvalidE;
      } else {
        b = fn_compose(b, ca);
      }
    }
    pb.tile_a0[tile] = a.a0;
    pb.tile_a1[tile] = a.a1;
    pb.tile_b0[tile] = b.a0;
    pb.tile_b1[tile] = b.a1;
    pb.tile_wx[tile] = twx;
    pb.tile_E[tile] = static_cast<int16_t>(E);
    pb.tile_jx[tile] = static_cast<int16_t>(kx);
  }
}

// Exact running sums of one chunk from the exact carry `c` (warp-uniform), for
// chunks whose summary cannot be used: the literal sequence of double
// additions.  The warp stages the chunk in shared memory, lane 0 runs the
// dependent chain (bound by the DADD latency, the loads are issued ahead in
// groups of 16), the warp writes the result back coalesced: 1.8-2.1 us per chunk on
// the B200 (profiles/r2_resolve_trace.txt).  A closed-form variant that
// re-summarised the lanes around the crossing inside this phase was tried twice;
// on one latency-bound warp its ~1500 instructions took 4.4 us.
__device__ __forceinline__ double exact_chunk(const double *w, size_t e0, int cnt, double c, double *out,
                                              int lane, double *s_w, double *s_o) {
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
#ifdef QMCL_RESOLVE_TRACE
// Diagnostic build only: (event id, %globaltimer) pairs of the resolving warp (tools/resolve_trace.py).
static __device__ unsigned long long g_resolve_trace[4096];
static __device__ unsigned int g_resolve_trace_n;
#define RTRACE(id)                                                                 \
  do {                                                                             \
    if (lane == 0) {                                                               \
      unsigned long long t_;                                                       \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                       \
      const unsigned k_ = g_resolve_trace_n;                                       \
      if (k_ + 2 <= 4096) {                                                        \
        g_resolve_trace[k_] = (id);                                                \
        g_resolve_trace[k_ + 1] = t_;                                              \
        g_resolve_trace_n = k_ + 2;                                                \
      }                                                                            \
    }                                                                              \
  } while (0)
#else
#define RTRACE(id) do { } while (0)
#endif
// Run by ONE warp (lane = its lane index); only __syncwarp inside.
QMCL_PHASE_FN void prefix_resolve_warp(const double *w, size_t n, size_t nc, size_t nt,
                                                    double carry_in, const PrefixBuffers &pb,
                                                    double *out, int lane, ResolveScratch *S) {
  double c = carry_in;  // uniform across the warp
#ifdef QMCL_RESOLVE_TRACE
  if (lane == 0) g_resolve_trace_n = 0;
#endif
  RTRACE(1);
  for (size_t base = 0; base < nt; base += 32) {
    // the summaries of 32 tiles, one per lane
    const size_t t = base + lane;
    Fn f = fn_identity(), g = fn_identity();
    double twx = 0.0;
    int E = -2, tjx = -1;
    if (t < nt) {
      f.a0 = pb.tile_a0[t];
      f.a1 = pb.tile_a1[t];
      g.a0 = pb.tile_b0[t];
      g.a1 = pb.tile_b1[t];
      twx = pb.tile_wx[t];
      E = pb.tile_E[t];
      tjx = pb.tile_jx[t];
    }
    RTRACE(2);
    const int Ec = biased_exponent_if_normal(c);
    const bool lane_ok = (E == -2) || (E == Ec && Ec != kInvalidE && tjx < 0);
    if (__all_sync(0xffffffffu, lane_ok)) {
      // no crossing expected in any of them: one scan
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
        RTRACE(3);
        continue;
      }
    }
    RTRACE(3);
    // Walk these (up to) 32 tiles one by one, every lane alike.  A lone warp issues one
    // instruction every ~5 cycles whatever it is (measured: profiles/r2_resolve_trace.txt), so
    // the step of the common kind is kept to a couple of dozen instructions: the carry travels
    // as (cE, K), the next tile's summary is fetched from shared memory ahead of its use, and
    // each lane just remembers the carry of "its" tile — the carries are stored after the walk.
    __syncwarp();
    S->ta0[lane] = f.a0;
    S->ta1[lane] = f.a1;
    S->tb0[lane] = g.a0;
    S->tb1[lane] = g.a1;
    S->twx[lane] = twx;
    S->tE[lane] = E;
    S->tjx[lane] = tjx;
    __syncwarp();
    const int n_win = static_cast<int>(((base + 32 < nt) ? base + 32 : nt) - base);
    int cE = biased_exponent_if_normal(c);
    unsigned long long K = mantissa_int(c);
    int nE = S->tE[0], nj = S->tjx[0];
    unsigned long long na0 = S->ta0[0], na1 = S->ta1[0];
    double my_cin = 0.0;
    uint8_t my_mode = MODE_MIXED;
    for (int i = 0; i < n_win; ++i) {
      const int tE = nE, tj = nj;
      const unsigned long long a0 = na0, a1 = na1;
      nE = S->tE[i + 1];
      nj = S->tjx[i + 1];
      na0 = S->ta0[i + 1];
      na1 = S->ta1[i + 1];
      if (lane == i) my_cin = c;
      if (tE == cE && cE != kInvalidE) {
        const unsigned long long K1 = K + ((K & 1ull) ? a1 : a0);
        if (K1 < kKLimit) {
          if (tj < 0) {
            if (lane == i) my_mode = MODE_FAST;
            K = K1;
            c = make_double(cE, K1);
            continue;
          }
          // predicted crossing: one literal addition, then the rest of the tile one binade up
          const double c2 = __dadd_rn(make_double(cE, K1), S->twx[i]);
          if (biased_exponent_if_normal(c2) == cE + 1) {
            const unsigned long long K3 = fn_apply(Fn{S->tb0[i], S->tb1[i]}, mantissa_int(c2));
            if (K3 < kKLimit) {
              if (lane == i) my_mode = MODE_CROSS;
              ++cE;
              K = K3;
              c = make_double(cE, K3);
              continue;
            }
          }
        }
      }
      const size_t tt = base + i;
      // open the tile (my_mode stays MODE_MIXED): its chunks' summaries, one per lane
      RTRACE(5);
      const size_t ch0 = tt * kTileChunks;
      const size_t ch1 = (ch0 + kTileChunks < nc) ? ch0 + kTileChunks : nc;
      Fn ca = fn_identity(), cb = fn_identity();
      double cwx = 0.0;
      int chE = kInvalidE, cjx = -1;
      if (ch0 + lane < ch1) {
        const size_t ch = ch0 + lane;
        ca.a0 = pb.chunk_a0[ch];
        ca.a1 = pb.chunk_a1[ch];
        cb.a0 = pb.chunk_b0[ch];
        cb.a1 = pb.chunk_b1[ch];
        cwx = pb.chunk_wx[ch];
        chE = pb.chunk_E[ch];
        cjx = pb.chunk_jx[ch];
      }
      RTRACE(4);
      for (size_t ch = ch0; ch < ch1; ++ch) {
        const int k = static_cast<int>(ch - ch0);
        cE = biased_exponent_if_normal(c);
        const int kE = __shfl_sync(0xffffffffu, chE, k);
        const int kj = __shfl_sync(0xffffffffu, cjx, k);
        if (kE != kInvalidE && kE == cE) {
          double c_next;
          if (apply_summary(c, cE, fn_shfl(ca, k), kj, __shfl_sync(0xffffffffu, cwx, k), fn_shfl(cb, k), &c_next)) {
            if (lane == 0) {
              pb.chunk_cin[ch] = c;
              pb.chunk_mode[ch] = kj < 0 ? MODE_FAST : MODE_CROSS;
            }
            c = c_next;
            RTRACE(6);
            continue;
          }
        }
        const size_t e0 = ch * kChunk;
        const int cnt = static_cast<int>((e0 + kChunk <= n) ? kChunk : n - e0);
        if (lane == 0) {
          pb.chunk_mode[ch] = MODE_SLOW;  // written here, phase 3 skips it
          pb.chunk_cin[ch] = c;
        }
        RTRACE(7);
        c = exact_chunk(w, e0, cnt, c, out, lane, S->w, S->o);
        RTRACE(8);
      }
      cE = biased_exponent_if_normal(c);
      K = mantissa_int(c);
    }
    if (lane < n_win) {
      pb.tile_cin[base + lane] = my_cin;
      pb.tile_mode[base + lane] = my_mode;
    }
  }
  RTRACE(9);
  if (lane == 0) *pb.carry_out = c;
}

// ---- phase 3: expand the carries into the running sums
QMCL_PHASE_FN void prefix_apply_tile(const double *w, size_t n, size_t nc,
                                                  const PrefixBuffers &pb, double *out, size_t tile) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const size_t chunk = tile * kTileChunks + wid;
  if (chunk >= nc) return;
  // the chunk's exact carry-in (Eb, K) and its crossing element, if phase 2 confirmed one
  int Eb, jx = kChunk;
  unsigned long long K;
  const uint8_t tmode = pb.tile_mode[tile];
  if (tmode != MODE_MIXED) {
    // from the tile's carry-in through the summaries of the chunks before this one
    Eb = pb.tile_E[tile];
    const int kx = tmode == MODE_CROSS ? pb.tile_jx[tile] : -1;
    double c = pb.tile_cin[tile];
    Fn a = fn_identity(), b = fn_identity();
    double cwx = 0.0;
    if (lane < wid) {
      const size_t ch = tile * kTileChunks + lane;
      a.a0 = pb.chunk_a0[ch];
      a.a1 = pb.chunk_a1[ch];
      if (lane == kx) {
        b.a0 = pb.chunk_b0[ch];
        b.a1 = pb.chunk_b1[ch];
        cwx = pb.chunk_wx[ch];
      }
    }
    for (int j = 0; j < wid; ++j) {
      unsigned long long k = fn_apply(fn_shfl(a, j), mantissa_int(c));
      if (j == kx) {
        const double c2 = __dadd_rn(make_double(Eb, k), __shfl_sync(0xffffffffu, cwx, j));
        k = fn_apply(fn_shfl(b, j), mantissa_int(c2));
        ++Eb;
      }
      c = make_double(Eb, k);
    }
    K = mantissa_int(c);
    if (wid == kx) jx = pb.chunk_jx[chunk];
  } else {
    const uint8_t mode = pb.chunk_mode[chunk];
    if (mode == MODE_SLOW) return;  // written by phase 2
    Eb = pb.chunk_E[chunk];
    K = mantissa_int(pb.chunk_cin[chunk]);
    if (mode == MODE_CROSS) jx = pb.chunk_jx[chunk];
  }
  const size_t base = chunk * kChunk + static_cast<size_t>(lane) * kLaneItems;
  double v[kLaneItems];
  load_items(w, base, n, v);
  // elements before jx under Eb, element jx added literally, elements after it under Eb + 1
  Fn e[kLaneItems];
  Fn head = fn_identity(), tail = fn_identity();
  double mine = 0.0;
#pragma unroll
  for (int i = 0; i < kLaneItems; ++i) {
    const int idx = lane * kLaneItems + i;
    e[i] = fn_identity();
    if (idx == jx) mine = v[i];
    else elem_fn(v[i], idx < jx ? Eb : Eb + 1, &e[i]);
    if (idx < jx) head = fn_compose(head, e[i]);
    else tail = fn_compose(tail, e[i]);
  }
  const Fn hs = fn_warp_scan(head, lane);
  Fn hx = fn_shfl_up(hs, 1);
  if (lane == 0) hx = fn_identity();
  unsigned long long Kh = fn_apply(hx, K), Kt = 0;
  double c2 = 0.0;
  if (jx < kChunk) {
    const Fn ts = fn_warp_scan(tail, lane);
    Fn tx = fn_shfl_up(ts, 1);
    if (lane == 0) tx = fn_identity();
    c2 = __dadd_rn(make_double(Eb, fn_apply(fn_shfl(hs, 31), K)), __shfl_sync(0xffffffffu, mine, jx / kLaneItems));
    Kt = fn_apply(tx, mantissa_int(c2));
  }
  double r[kLaneItems];
#pragma unroll
  for (int i = 0; i < kLaneItems; ++i) {
    const int idx = lane * kLaneItems + i;
    if (idx < jx) {
      Kh = fn_apply(e[i], Kh);
      r[i] = make_double(Eb, Kh);
    } else if (idx > jx) {
      Kt = fn_apply(e[i], Kt);
      r[i] = make_double(Eb + 1, Kt);
    } else {
      r[i] = c2;
    }
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


inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace pfx
}  // namespace qmcl
