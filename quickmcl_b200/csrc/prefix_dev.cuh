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
  int lanes;             // 1: binade-crossing chunks are resolved lane-wise (lane_chunk)
};

// ---- phase 1: chunk and tile summaries under the guessed binade
QMCL_PHASE_FN void prefix_summaries_tile(const double *w, size_t n, size_t nc,
                                                      const PrefixBuffers &pb, size_t tile) {
  __shared__ unsigned long long s_a0[kTileChunks], s_a1[kTileChunks];
  __shared__ int s_E[kTileChunks];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const size_t chunk = tile * kTileChunks + wid;
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
    pb.tile_a0[tile] = t.a0;
    pb.tile_a1[tile] = t.a1;
    pb.tile_E[tile] = static_cast<int16_t>(E);
  }
}

// Phase 2 for a chunk whose summary does not apply because the running sum
// crosses into the next binade inside it (about log2(N) chunks of a normalised
// weight array).  Instead of 256 dependent additions the warp evaluates every
// lane's 8 elements in closed form under the carry's binade E and under E + 1:
// the lanes before the one in which the sum really crosses follow E, the lanes
// after it E + 1, and only the crossing lane's 8 elements are added one by one
// (the same additions on every lane, its elements broadcast by shuffles).  The
// chunk's running sums are written here.  Returns false — nothing written —
// when this does not apply (carry not normal, a second crossing, a negative /
// subnormal / non-finite weight after the crossing); the caller then falls back
// to the literal additions.  Bit-identical to them whenever it returns true
// (CPU model of the algorithm: 540 adversarial arrays, round 1; GPU:
// tests/test_prefix_gpu.py against the one-warp sequential kernel and numpy).
__device__ __forceinline__ bool lane_chunk(const double *w, size_t n, size_t ch, double c, double *out,
                                           int lane, double *c_out) {
  const int cE = biased_exponent_if_normal(c);
  if (cE == kInvalidE || cE >= 0x7fd) return false;
  const size_t base = ch * kChunk + static_cast<size_t>(lane) * kLaneItems;
  double v[kLaneItems];
  load_items(w, base, n, v);
  Fn e0[kLaneItems], e1[kLaneItems];
  Fn F0 = fn_identity(), F1 = fn_identity();
  bool ok0 = true, ok1 = true;
#pragma unroll
  for (int i = 0; i < kLaneItems; ++i) {
    ok0 &= elem_fn(v[i], cE, &e0[i]);
    F0 = fn_compose(F0, e0[i]);
    ok1 &= elem_fn(v[i], cE + 1, &e1[i]);
    F1 = fn_compose(F1, e1[i]);
  }
  const Fn incl = fn_warp_scan(F0, lane);
  Fn excl = fn_shfl_up(incl, 1);
  if (lane == 0) excl = fn_identity();
  const unsigned long long K0 = mantissa_int(c);
  const unsigned long long Kin = fn_apply(excl, K0), Kout = fn_apply(incl, K0);
  // weights are >= 0 wherever a closed form exists, so K only grows: a lane is
  // good iff its own end point is still inside the binade
  const bool fine = ok0 && Kout < kKLimit;
  const unsigned int bad = __ballot_sync(0xffffffffu, !fine);
  int L = 32, myE = cE;
  double cin;
  if (bad == 0) {
    cin = make_double(cE, Kin);
    *c_out = make_double(cE, __shfl_sync(0xffffffffu, Kout, 31));
  } else {
    L = __ffs(bad) - 1;
    const double cL = make_double(cE, __shfl_sync(0xffffffffu, Kin, L));
    double cc = cL;  // the crossing lane's 8 elements: true additions (padding zeros add nothing)
#pragma unroll
    for (int i = 0; i < kLaneItems; ++i) cc = __dadd_rn(cc, __shfl_sync(0xffffffffu, v[i], L));
    if (biased_exponent_if_normal(cc) != cE + 1) return false;
    const Fn G = lane > L ? F1 : fn_identity();
    const bool okG = lane > L ? ok1 : true;
    const Fn incl2 = fn_warp_scan(G, lane);
    Fn excl2 = fn_shfl_up(incl2, 1);
    if (lane == 0) excl2 = fn_identity();
    const unsigned long long K1 = mantissa_int(cc);
    const unsigned long long Kin2 = fn_apply(excl2, K1), Kout2 = fn_apply(incl2, K1);
    if (!__all_sync(0xffffffffu, okG && Kout2 < kKLimit)) return false;  // e.g. a second crossing
    cin = lane < L ? make_double(cE, Kin) : (lane == L ? cL : make_double(cE + 1, Kin2));
    myE = lane < L ? cE : cE + 1;
    *c_out = make_double(cE + 1, __shfl_sync(0xffffffffu, Kout2, 31));
  }
  double r[kLaneItems];
  if (lane == L) {
    double cc = cin;
#pragma unroll
    for (int i = 0; i < kLaneItems; ++i) {
      cc = __dadd_rn(cc, v[i]);
      r[i] = cc;
    }
  } else {
    unsigned long long K = mantissa_int(cin);
#pragma unroll
    for (int i = 0; i < kLaneItems; ++i) {
      K = fn_apply(myE == cE ? e0[i] : e1[i], K);
      r[i] = make_double(myE, K);
    }
  }
#pragma unroll
  for (int i = 0; i < kLaneItems; ++i)
    if (base + i < n) out[base + i] = r[i];
  return true;
}

// Exact running sums of one chunk from the exact carry `c` (warp-uniform), for
// chunks whose summary cannot be used: the literal sequence of double
// additions.  The warp stages the chunk in shared memory, lane 0 runs the
// dependent chain (bound by the DADD latency, the loads are issued ahead in
// groups of 16), the warp writes the result back coalesced.  (A closed-form
// variant that re-summarised the lanes around the crossing was tried; on one
// latency-bound warp its ~800 instructions per round were slower than these
// 256 additions.)
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
// Run by ONE warp (lane = its lane index); only __syncwarp inside.
QMCL_PHASE_FN void prefix_resolve_warp(const double *w, size_t n, size_t nc, size_t nt,
                                                    double carry_in, const PrefixBuffers &pb,
                                                    double *out, int lane) {
  // chunk summaries of the current batch of 32 tiles, staged on demand
  __shared__ unsigned long long s_a0[32 * kTileChunks], s_a1[32 * kTileChunks];
  __shared__ int16_t s_E[32 * kTileChunks];
  __shared__ double s_w[kChunk], s_o[kChunk];
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
        if (pb.lanes) {
          double c_next;
          if (lane_chunk(w, n, ch, c, out, lane, &c_next)) {
            if (lane == 0) {
              pb.chunk_mode[ch] = MODE_SLOW;  // written here, phase 3 skips it
              pb.chunk_cin[ch] = c;
            }
            c = c_next;
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
QMCL_PHASE_FN void prefix_apply_tile(const double *w, size_t n, size_t nc,
                                                  const PrefixBuffers &pb, double *out, size_t tile) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
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


inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace pfx
}  // namespace qmcl
