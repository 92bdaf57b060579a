// CPU model of K9 (csrc/prefix.cu) with the proposed "lane summaries" for the
// chunks in which the running sum crosses a power of two.  Plain loops stand in
// for the warp: lane l of a chunk owns elements [8l, 8l+8).  The model is run
// against the literal sequential sum on adversarial inputs; it validates the
// ALGORITHM (the device code is a transliteration with shuffles).
//
//   g++ -O2 -std=c++17 -o model model.cpp && ./model
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

typedef unsigned long long u64;
constexpr int kChunk = 256, kLaneItems = 8, kTileChunks = 8;
constexpr u64 kMantMask = (1ull << 52) - 1, kImplicit = 1ull << 52, kKLimit = 1ull << 53;
constexpr int kInvalidE = -1;
constexpr int kMaxSlots = 128;

static long long bits_of(double v) { long long b; std::memcpy(&b, &v, 8); return b; }
static double from_bits(u64 b) { double v; std::memcpy(&v, &b, 8); return v; }

struct Fn { u64 a0, a1; };
static Fn fn_identity() { return Fn{0, 0}; }
static Fn fn_compose(const Fn &f, const Fn &g) {  // f first, then g
  Fn h;
  h.a0 = f.a0 + ((f.a0 & 1ull) ? g.a1 : g.a0);
  h.a1 = f.a1 + (((f.a1 + 1ull) & 1ull) ? g.a1 : g.a0);
  return h;
}
static u64 fn_apply(const Fn &f, u64 K) { return K + ((K & 1ull) ? f.a1 : f.a0); }

static bool elem_fn(double w, int Eb, Fn *out) {
  const long long bits = bits_of(w);
  *out = fn_identity();
  if ((bits & 0x7fffffffffffffffll) == 0) return true;
  if (bits < 0) return false;
  const int ew = static_cast<int>(bits >> 52);
  if (ew == 0 || ew == 0x7ff) return false;
  const int s = Eb - ew;
  if (s <= 0) return false;
  if (s >= 54) return true;
  const u64 m = (static_cast<u64>(bits) & kMantMask) | kImplicit;
  const u64 q = m >> s;
  const u64 rem = m & ((1ull << s) - 1ull);
  const u64 half = 1ull << (s - 1);
  if (rem > half) out->a0 = out->a1 = q + 1;
  else if (rem < half) out->a0 = out->a1 = q;
  else { out->a0 = q + (q & 1ull); out->a1 = q + ((q + 1ull) & 1ull); }
  return true;
}
static int exponent_if_normal(double c) {
  const long long b = bits_of(c);
  if (b <= 0) return kInvalidE;
  const int e = static_cast<int>(b >> 52);
  return (e == 0 || e == 0x7ff) ? kInvalidE : e;
}
static u64 mantissa_int(double c) { return (static_cast<u64>(bits_of(c)) & kMantMask) | kImplicit; }
static double make_double(int Eb, u64 K) { return from_bits((static_cast<u64>(Eb) << 52) | (K & kMantMask)); }

struct Slot {
  int Elo;
  uint32_t ok[2];
  Fn fn[2][32];
  double cin[32];
};

struct Stats { long fast_tiles = 0, fast_chunks = 0, lane_chunks = 0, exact_chunks = 0; };

// Returns the running sums; counts which path each chunk took.
static std::vector<double> model(const std::vector<double> &w, double carry_in, Stats *st) {
  const size_t n = w.size();
  const size_t nc = (n + kChunk - 1) / kChunk, nt = (nc + kTileChunks - 1) / kTileChunks;
  auto at = [&](size_t i) { return i < n ? w[i] : 0.0; };
  // phase 0: plain chunk sums (any order) and approximate carry-ins
  std::vector<double> chunk_sum(nc), approx_in(nc);
  for (size_t c = 0; c < nc; ++c) {
    double s = 0;
    for (int l = 31; l >= 0; --l) {  // deliberately not the sequential order
      double t = 0;
      for (int i = 0; i < kLaneItems; ++i) t += at(c * kChunk + l * kLaneItems + i);
      s += t;
    }
    chunk_sum[c] = s;
  }
  {
    double a = carry_in;
    for (size_t c = 0; c < nc; ++c) { approx_in[c] = a; a += chunk_sum[c]; }
  }
  // phase 1: chunk / tile summaries under the guess
  std::vector<Fn> chunk_f(nc), tile_f(nt);
  std::vector<int> chunk_E(nc), tile_E(nt);
  for (size_t c = 0; c < nc; ++c) {
    int Eb = exponent_if_normal(approx_in[c]);
    Fn f = fn_identity();
    bool ok = Eb != kInvalidE;
    if (ok)
      for (int k = 0; k < kChunk; ++k) {
        Fn e;
        ok &= elem_fn(at(c * kChunk + k), Eb, &e);
        f = fn_compose(f, e);
      }
    if (!ok) Eb = kInvalidE;
    chunk_f[c] = f;
    chunk_E[c] = Eb;
  }
  for (size_t t = 0; t < nt; ++t) {
    Fn f = fn_identity();
    int E = chunk_E[t * kTileChunks];
    for (int k = 0; k < kTileChunks; ++k) {
      const size_t c = t * kTileChunks + k;
      if (c >= nc) break;
      if (chunk_E[c] != E) E = kInvalidE;
      f = fn_compose(f, chunk_f[c]);
    }
    tile_f[t] = f;
    tile_E[t] = E;
  }
  // phase 1b (new): lane summaries under Elo and Elo + 1 for the chunks in which the
  // approximate sum changes binade
  std::vector<int> chunk_slot(nc, -1);
  std::vector<Slot> slots;
  for (size_t c = 0; c < nc; ++c) {
    const int Elo = exponent_if_normal(approx_in[c]);
    const int Eout = exponent_if_normal(approx_in[c] + chunk_sum[c]);
    if (Elo == kInvalidE || Eout == Elo) continue;
    if (static_cast<int>(slots.size()) >= kMaxSlots) continue;
    Slot s;
    s.Elo = Elo;
    for (int k = 0; k < 2; ++k) {
      s.ok[k] = 0;
      for (int l = 0; l < 32; ++l) {
        Fn f = fn_identity();
        bool ok = true;
        for (int i = 0; i < kLaneItems; ++i) {
          Fn e;
          ok &= elem_fn(at(c * kChunk + l * kLaneItems + i), Elo + k, &e);
          f = fn_compose(f, e);
        }
        s.fn[k][l] = f;
        if (ok) s.ok[k] |= 1u << l;
      }
    }
    chunk_slot[c] = static_cast<int>(slots.size());
    slots.push_back(s);
  }
  // phase 2 + 3
  std::vector<double> out(n);
  auto expand_fast = [&](size_t c, double cin, int Eb) {  // closed form of one chunk
    u64 K = mantissa_int(cin);
    for (int k = 0; k < kChunk; ++k) {
      const size_t i = c * kChunk + k;
      Fn e;
      elem_fn(at(i), Eb, &e);
      K = fn_apply(e, K);
      if (i < n) out[i] = make_double(Eb, K);
    }
  };
  auto exact_chunk = [&](size_t c, double cin) {
    double cc = cin;
    for (int k = 0; k < kChunk; ++k) {
      const size_t i = c * kChunk + k;
      if (i < n) { cc += w[i]; out[i] = cc; }
    }
    return cc;
  };
  // the proposed lane path; returns false when it does not apply
  auto lane_chunk = [&](size_t ch, double c, double *c_out) -> bool {
    const int slot = chunk_slot[ch];
    if (slot < 0) return false;
    Slot &s = slots[slot];
    const int cE = exponent_if_normal(c);
    if (cE == kInvalidE || (cE != s.Elo && cE != s.Elo + 1)) return false;
    const int set = cE - s.Elo;
    // inclusive / exclusive scans over the lanes
    u64 K0 = mantissa_int(c), Kin[32], Kout[32];
    {
      u64 K = K0;
      for (int l = 0; l < 32; ++l) { Kin[l] = K; K = fn_apply(s.fn[set][l], K); Kout[l] = K; }
    }
    int L = -1;
    for (int l = 0; l < 32; ++l)
      if (!((s.ok[set] >> l) & 1u) || Kout[l] >= kKLimit) { L = l; break; }
    double cin[32];
    if (L < 0) {
      for (int l = 0; l < 32; ++l) cin[l] = make_double(cE, Kin[l]);
      *c_out = make_double(cE, Kout[31]);
    } else {
      if (set != 0) return false;  // the second set would be Elo + 2
      for (int l = 0; l < L; ++l) cin[l] = make_double(cE, Kin[l]);
      const double cL = make_double(cE, Kin[L]);
      cin[L] = cL;
      double cc = cL;
      for (int i = 0; i < kLaneItems; ++i) {
        const size_t idx = ch * kChunk + L * kLaneItems + i;
        if (idx < n) cc += w[idx];
      }
      if (exponent_if_normal(cc) != cE + 1) return false;
      u64 K = mantissa_int(cc);
      for (int l = L + 1; l < 32; ++l) {
        if (!((s.ok[1] >> l) & 1u)) return false;
        cin[l] = make_double(cE + 1, K);
        K = fn_apply(s.fn[1][l], K);
        if (K >= kKLimit) return false;  // a second crossing
      }
      *c_out = make_double(cE + 1, K);
    }
    // phase 3 for this chunk: every lane expands from its own carry-in — closed form
    // under the binade of that carry when it holds for the lane, true additions otherwise
    for (int l = 0; l < 32; ++l) {
      const int Eb = exponent_if_normal(cin[l]);
      Fn e[kLaneItems];
      bool ok = Eb != kInvalidE;
      u64 K = ok ? mantissa_int(cin[l]) : 0;
      if (ok)
        for (int i = 0; i < kLaneItems; ++i) {
          ok &= elem_fn(at(ch * kChunk + l * kLaneItems + i), Eb, &e[i]);
          if (ok) { K = fn_apply(e[i], K); ok &= K < kKLimit; }
        }
      if (ok) {
        K = mantissa_int(cin[l]);
        for (int i = 0; i < kLaneItems; ++i) {
          const size_t idx = ch * kChunk + l * kLaneItems + i;
          K = fn_apply(e[i], K);
          if (idx < n) out[idx] = make_double(Eb, K);
        }
      } else {
        double cc = cin[l];
        for (int i = 0; i < kLaneItems; ++i) {
          const size_t idx = ch * kChunk + l * kLaneItems + i;
          if (idx < n) { cc += w[idx]; out[idx] = cc; }
        }
      }
    }
    return true;
  };
  double c = carry_in;
  for (size_t t = 0; t < nt; ++t) {
    int cE = exponent_if_normal(c);
    if (tile_E[t] != kInvalidE && tile_E[t] == cE) {
      const u64 Kend = fn_apply(tile_f[t], mantissa_int(c));
      if (Kend < kKLimit) {
        u64 K = mantissa_int(c);
        for (int k = 0; k < kTileChunks; ++k) {
          const size_t ch = t * kTileChunks + k;
          if (ch >= nc) break;
          expand_fast(ch, make_double(cE, K), cE);
          K = fn_apply(chunk_f[ch], K);
        }
        c = make_double(cE, Kend);
        st->fast_tiles++;
        continue;
      }
    }
    for (int k = 0; k < kTileChunks; ++k) {
      const size_t ch = t * kTileChunks + k;
      if (ch >= nc) break;
      cE = exponent_if_normal(c);
      if (chunk_E[ch] != kInvalidE && chunk_E[ch] == cE) {
        const u64 Kend = fn_apply(chunk_f[ch], mantissa_int(c));
        if (Kend < kKLimit) {
          expand_fast(ch, c, cE);
          c = make_double(cE, Kend);
          st->fast_chunks++;
          continue;
        }
      }
      double c_next;
      if (lane_chunk(ch, c, &c_next)) {
        c = c_next;
        st->lane_chunks++;
        continue;
      }
      c = exact_chunk(ch, c);
      st->exact_chunks++;
    }
  }
  return out;
}

static std::vector<double> sequential(const std::vector<double> &w, double carry_in) {
  std::vector<double> out(w.size());
  volatile double c = carry_in;
  for (size_t i = 0; i < w.size(); ++i) { c = c + w[i]; out[i] = c; }
  return out;
}

int main() {
  std::mt19937_64 rng(12345);
  std::uniform_real_distribution<double> U(0.0, 1.0);
  Stats st;
  long cases = 0, bad = 0;
  auto check = [&](const std::vector<double> &w, double carry, const char *name) {
    const auto got = model(w, carry, &st), want = sequential(w, carry);
    ++cases;
    for (size_t i = 0; i < w.size(); ++i)
      if (bits_of(got[i]) != bits_of(want[i])) {
        if (bad < 10) std::printf("MISMATCH %s n=%zu i=%zu got %a want %a\n", name, w.size(), i, got[i], want[i]);
        ++bad;
        break;
      }
  };
  for (int rep = 0; rep < 60; ++rep) {
    const size_t n = 1 + rng() % 70000;
    std::vector<double> w(n);
    // normalised weights (the filter's case): crosses ~17 binades
    double tot = 0;
    for (auto &v : w) { v = U(rng) < 0.1 ? 0.0 : std::pow(U(rng), 3.0); tot += v; }
    for (auto &v : w) v /= tot;
    check(w, 0.0, "normalised");
    check(w, U(rng), "normalised+carry");
    // ties: weights that are exact multiples of small powers of two
    for (auto &v : w) v = std::ldexp(static_cast<double>(rng() % 5), -static_cast<int>(40 + rng() % 20));
    check(w, 0.0, "ties");
    check(w, 1.0 - std::ldexp(1.0, -30), "ties near 1");
    // wide dynamic range
    for (auto &v : w) v = std::ldexp(U(rng), static_cast<int>(rng() % 120) - 100);
    check(w, 0.0, "wide");
    // crossing placed right at lane / chunk boundaries
    for (auto &v : w) v = std::ldexp(1.0, -20);
    check(w, 1.0 - 8 * std::ldexp(1.0, -20) * (1 + rng() % 64), "boundary");
    check(w, 0.5 - 256 * std::ldexp(1.0, -20) * (1 + rng() % 8), "chunk boundary");
    // a few invalid inputs (negative, subnormal) sprinkled in
    for (auto &v : w) v = std::pow(U(rng), 2.0) * 1e-4;
    for (int k = 0; k < 5; ++k) w[rng() % n] = (k & 1) ? -1e-6 : 4.9e-324;
    check(w, 0.0, "invalid");
    // big steps: elements of the order of the sum itself (several binades per chunk)
    for (auto &v : w) v = U(rng) < 0.02 ? std::ldexp(U(rng), static_cast<int>(rng() % 8)) : 1e-9 * U(rng);
    check(w, 0.0, "big steps");
  }
  std::printf("cases %ld, mismatches %ld; chunks: fast tiles %ld, fast chunks %ld, lane path %ld, exact %ld\n",
              cases, bad, st.fast_tiles, st.fast_chunks, st.lane_chunks, st.exact_chunks);
  return bad ? 1 : 0;
}
