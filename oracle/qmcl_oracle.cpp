// qmcl_oracle.cpp — CPU restatement of QuickMCL's particle-filter hot path.
//
// TEST INFRASTRUCTURE ONLY (see qmcl_oracle.h).  Dependency-free C++17; built
// with -O3 -DNDEBUG -ffp-contract=off so that every float/double expression
// rounds exactly where the reference's x86-64 (no FMA) Release build rounds.
//
// All file:line citations are relative to /root/reference.
// Parity status: helpers are pinned by the reference's gtest golden vectors
// (tests/test_oracle_golden.py); the RNG-driven stages are "parity unpinned"
// by the reference (no test seeds its RNG), see header.

#include "qmcl_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <map>
#include <queue>
#include <vector>

namespace {

constexpr double kPi = 3.14159265358979323846;  // M_PI

double g_last_seconds = 0.0;
struct ScopedTimer {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  ~ScopedTimer() {
    g_last_seconds = std::chrono::duration<double>(
                         std::chrono::steady_clock::now() - t0)
                         .count();
  }
};

// ---------------------------------------------------------------- utils.h
// utils.h:28-35.  For a float argument the expression is still evaluated in
// double (M_PI is double) and converted back on return.
double norm_angle_d(double a) {
  return a - 2 * kPi * std::floor((a + kPi) / (2 * kPi));
}
float norm_angle_f(float a) {
  return static_cast<float>(a - 2 * kPi * std::floor((a + kPi) / (2 * kPi)));
}
// utils.h:38-61
double angle_delta_d(double a, double b) {
  a = norm_angle_d(a);
  b = norm_angle_d(b);
  double d1 = a - b;
  double d2 = 2 * kPi - std::fabs(d1);
  if (0 < d1) d2 = -d2;
  return (std::fabs(d1) < std::fabs(d2)) ? d1 : d2;
}
float angle_delta_f(float a, float b) {
  a = norm_angle_f(a);
  b = norm_angle_f(b);
  float d1 = a - b;
  // "auto d2 = 2 * M_PI - std::abs(d1)" is double even for float inputs.
  double d2 = 2 * kPi - std::fabs(d1);
  if (0 < d1) d2 = -d2;
  if (std::fabs(d1) < std::fabs(d2)) return d1;
  return static_cast<float>(d2);
}

// ------------------------------------------------------------ scaled_map.h
// scaled_map.h:43-48 + :94-98: (pos - offset) / res in float, then the C++
// float->int conversion (truncation toward zero).
inline int cell_of(float pos, float offset, float res) {
  return static_cast<int>((pos - offset) / res);
}
inline int cell_of_inf(float pos, float res) {  // ScaledInfiniteMapLogic
  return static_cast<int>(pos / res);
}

// ------------------------------------------------------------------ Philox
// Philox4x32-10 (Salmon et al., SC'11), the shared RNG contract.
inline void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                   uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = 0xD2511F53ull * c0;
    uint64_t p1 = 0xCD9E8D57ull * c2;
    uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = static_cast<uint32_t>(p1);
    uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = static_cast<uint32_t>(p0);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum Purpose : uint32_t {
  P_INIT = 1, P_GLOBAL = 2, P_MOTION = 3, P_LV_OFFSET = 4, P_DRAW = 5,
  P_INJECT = 6
};

struct Rng {
  uint64_t seed = 0;
  uint32_t step = 0;
  void draw(uint64_t index, uint32_t purpose, uint32_t out[4]) const {
    philox(static_cast<uint32_t>(index), static_cast<uint32_t>(index >> 32),
           step, purpose, static_cast<uint32_t>(seed),
           static_cast<uint32_t>(seed >> 32), out);
  }
};

inline double u53(uint32_t lo, uint32_t hi) {
  uint64_t v = (static_cast<uint64_t>(hi) << 32) | lo;
  return static_cast<double>(v >> 11) * 0x1.0p-53;
}
// Box-Muller in double from two 32-bit words.
inline void normal_pair(uint32_t a, uint32_t b, double *z0, double *z1) {
  double u1 = (static_cast<double>(a) + 1.0) * 0x1.0p-32;  // (0,1]
  double u2 = static_cast<double>(b) * 0x1.0p-32;          // [0,1)
  double rad = std::sqrt(-2.0 * std::log(u1));
  double ang = (2 * kPi) * u2;
  *z0 = rad * std::cos(ang);
  *z1 = rad * std::sin(ang);
}
inline uint64_t mulhi64(uint64_t a, uint64_t b) {
  return static_cast<uint64_t>((static_cast<unsigned __int128>(a) * b) >> 64);
}

// -------------------------------------------------------------- statistics
struct Stats {
  double total_weight = 0.0;
  uint64_t sample_count = 0;
  double mean[4] = {0, 0, 0, 0};
  double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};

  // statistics.cpp:23-41; pose_2d.h:137-153
  void add(const orc_particle &p) {
    sample_count++;
    total_weight += p.weight;
    double d[4] = {static_cast<double>(p.x), static_cast<double>(p.y),
                   std::sin(static_cast<double>(p.theta)),
                   std::cos(static_cast<double>(p.theta))};
    for (int k = 0; k < 4; ++k) mean[k] += p.weight * d[k];
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j) cov[i * 3 + j] += (p.weight * d[i]) * d[j];
  }
  // statistics.cpp:50-57
  void add(const Stats &s) {
    total_weight += s.total_weight;
    sample_count += s.sample_count;
    for (int k = 0; k < 4; ++k) mean[k] += s.mean[k];
    for (int k = 0; k < 9; ++k) cov[k] += s.cov[k];
  }
  // statistics.cpp:59-79
  void finalise() {
    mean[0] /= total_weight;
    mean[1] /= total_weight;
    double tmp[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 2; ++i)
      for (int j = 0; j < 2; ++j)
        tmp[i * 3 + j] = cov[i * 3 + j] / total_weight - mean[i] * mean[j];
    double R = std::sqrt(mean[2] * mean[2] + mean[3] * mean[3]);
    tmp[8] = 1 - R;
    mean[2] = std::atan2(mean[2], mean[3]);
    mean[3] = 0;
    std::memcpy(cov, tmp, sizeof(cov));
  }
  void pack(double out[15]) const {
    out[0] = total_weight;
    out[1] = static_cast<double>(sample_count);
    for (int k = 0; k < 4; ++k) out[2 + k] = mean[k];
    for (int k = 0; k < 9; ++k) out[6 + k] = cov[k];
  }
  void unpack(const double in[15]) {
    total_weight = in[0];
    sample_count = static_cast<uint64_t>(in[1]);
    for (int k = 0; k < 4; ++k) mean[k] = in[2 + k];
    for (int k = 0; k < 9; ++k) cov[k] = in[6 + k];
  }
};

// ------------------------------------------------------- running sum (CDF)
// particle.cpp:66-82 + particle.h:99-107.  std::map keeps the first record
// for a duplicated key, which is what the reference's insert() does.
struct RunningSum {
  std::map<double, size_t> intervals;
  double total = 0.0;
  explicit RunningSum(const std::vector<orc_particle> &ps) {
    double running = 0;
    for (size_t i = 0; i < ps.size(); ++i) {
      double pw = ps[i].weight;
      if (pw <= 0) continue;
      intervals.insert(std::make_pair(running, i));
      running += pw;
    }
    total = running;
  }
  size_t lookup(double w) const {
    auto it = intervals.lower_bound(w);
    if (it == intervals.end() || it->first > w) --it;
    return it->second;
  }
};

// ---------------------------------------------- low pass (low_pass_filter.h)
template <typename T> struct LowPass {
  T value = 0;
  T alpha = static_cast<T>(0.5);
  void update(T x) {  // :39-50
    if (value == 0) value = x;
    else value = value + alpha * (x - value);
  }
};

// -------------------------------------------- bins (space_partitioning.cpp)
struct BinKey {
  int x, y, t;
  bool operator<(const BinKey &o) const {
    if (x != o.x) return x < o.x;
    if (y != o.y) return y < o.y;
    return t < o.t;
  }
};

struct Bins {
  float res_xy = 0.5f, res_t = 0.17453292f;
  int t_lo = -17, t_hi = 17;
  double eps2 = 0.02, kld_z = 0.01;
  int32_t nmin = 100, nmax = 5000;
  // value: {count, component label}
  struct Val { int64_t count = 0; int32_t label = -1; };
  std::map<BinKey, Val> cells;
  size_t cluster_count = 0;

  // space_partitioning.cpp:23-37
  void configure(const orc_pf_params &p) {
    res_xy = p.res_xy;
    res_t = p.res_theta;
    eps2 = 2 * p.kld_epsilon;
    kld_z = p.kld_z;
    nmin = p.particle_count_min;
    nmax = p.particle_count_max;
    t_lo = cell_of_inf(static_cast<float>(-kPi + 0.001), res_t);
    t_hi = cell_of_inf(static_cast<float>(kPi - 0.001), res_t);
    reset();
  }
  void reset() { cells.clear(); cluster_count = 0; }
  BinKey key_of(float x, float y, float t) const {  // :45-50
    return BinKey{cell_of_inf(x, res_xy), cell_of_inf(y, res_xy),
                  cell_of_inf(t, res_t)};
  }
  void add_point(float x, float y, float t) { cells[key_of(x, y, t)].count++; }

  // space_partitioning.cpp:52-74
  static uint64_t limit_for(uint64_t k, double eps2, double z, int32_t nmin,
                            int32_t nmax) {
    if (k < 2) return static_cast<uint64_t>(nmax);
    double common = 2.0 / (9 * (k - 1));
    double to_be_cubed = 1 - common - std::sqrt(common) * z;
    double result =
        ((k - 1) / eps2) * to_be_cubed * to_be_cubed * to_be_cubed;
    int32_t fin = static_cast<int32_t>(std::ceil(result));
    if (fin < nmin) return static_cast<uint64_t>(nmin);
    if (fin > nmax) return static_cast<uint64_t>(nmax);
    return static_cast<uint64_t>(fin);
  }
  uint64_t limit() const {
    return limit_for(cells.size(), eps2, kld_z, nmin, nmax);
  }

  // space_partitioning.cpp:89-128.  The reference labels connected components
  // by a recursive flood fill over the neighbour relation "theta +-1 with
  // wrap, x +-1, y +-1".  Restated as union-find over the same edges (no
  // recursion: the reference overflows its stack on million-bin components),
  // with components numbered in sorted-key order instead of unordered_map
  // order (ids are not canonical in the reference, only the partition is).
  // Deviation: for theta bins outside [t_lo, t_hi] (a heading of exactly
  // +-pi as a float) the reference's relation is not symmetric and its result
  // depends on hash-map iteration order; the symmetric closure is used here.
  void assign_clusters() {
    std::vector<std::map<BinKey, Val>::iterator> its;
    its.reserve(cells.size());
    int32_t idx = 0;
    for (auto it = cells.begin(); it != cells.end(); ++it) {
      it->second.label = idx++;  // temporarily: own sorted index
      its.push_back(it);
    }
    std::vector<int32_t> parent(its.size());
    for (size_t i = 0; i < parent.size(); ++i) parent[i] = static_cast<int32_t>(i);
    auto find = [&](int32_t v) {
      while (parent[v] != v) {
        parent[v] = parent[parent[v]];
        v = parent[v];
      }
      return v;
    };
    for (size_t i = 0; i < its.size(); ++i) {
      const BinKey &k = its[i]->first;
      for (int dt = -1; dt <= 1; ++dt) {
        int t = k.t + dt;
        if (t < t_lo) t = t_hi;
        else if (t > t_hi) t = t_lo;
        for (int dx = -1; dx <= 1; ++dx)
          for (int dy = -1; dy <= 1; ++dy) {
            auto nb = cells.find(BinKey{k.x + dx, k.y + dy, t});
            if (nb == cells.end()) continue;
            int32_t a = find(static_cast<int32_t>(i)), b = find(nb->second.label);
            if (a == b) continue;
            if (a < b) parent[b] = a;
            else parent[a] = b;
          }
      }
    }
    // number components by their smallest sorted index
    std::vector<int32_t> id_of_root(its.size(), -1);
    cluster_count = 0;
    std::vector<int32_t> labels(its.size());
    for (size_t i = 0; i < its.size(); ++i) {
      int32_t r = find(static_cast<int32_t>(i));
      if (id_of_root[r] < 0) id_of_root[r] = static_cast<int32_t>(cluster_count++);
      labels[i] = id_of_root[r];
    }
    for (size_t i = 0; i < its.size(); ++i) its[i]->second.label = labels[i];
  }
  int32_t cluster_of(float x, float y, float t) const {  // :76-87
    auto it = cells.find(key_of(x, y, t));
    return it == cells.end() ? -1 : it->second.label;
  }
};

}  // namespace

// ======================================================================= map
struct orc_map {
  orc_likelihood_params params{0.5f, 0.5f, 30, 14.0f, 2.0f, 0.01f};
  float z_hit_pre = 2 * 0.01f * 0.01f;
  double max_distance_probability = 0;
  uint32_t w = 0, h = 0;
  float res = 0, ox = 0, oy = 0;
  std::vector<float> distance;  // metres
  std::vector<float> field;
  std::vector<int8_t> state;
  std::vector<int32_t> free_cells;  // x,y pairs, row-major order

  void rebuild_free_cells() {  // map_base.cpp:46-55
    free_cells.clear();
    for (uint32_t y = 0; y < h; ++y)
      for (uint32_t x = 0; x < w; ++x)
        if (state[static_cast<size_t>(y) * w + x] == 0) {
          free_cells.push_back(static_cast<int32_t>(x));
          free_cells.push_back(static_cast<int32_t>(y));
        }
  }
  int8_t state_at(float x, float y) const {  // map_base.cpp:66-77
    int cx = cell_of(x, ox, res), cy = cell_of(y, oy, res);
    if (0 <= cx && 0 <= cy && cx < static_cast<int>(w) &&
        cy < static_cast<int>(h))
      return state[static_cast<size_t>(cy) * w + cx];
    return 2;
  }
  double probability_at(float x, float y) const {  // map_likelihood.cpp:148-160
    int cx = cell_of(x, ox, res), cy = cell_of(y, oy, res);
    double p = max_distance_probability;
    if (0 <= cx && 0 <= cy && cx < static_cast<int>(w) &&
        cy < static_cast<int>(h))
      p = field[static_cast<size_t>(cy) * w + cx];
    return p;
  }
  // map_base.cpp:58-64 with the shared RNG contract: cell from words 0-1,
  // heading from word 2.  Cell *corner* (scaled_map.h:101-105).
  void random_pose(const uint32_t r[4], float *x, float *y, float *t) const {
    uint64_t nfree = free_cells.size() / 2;
    uint64_t idx = mulhi64((static_cast<uint64_t>(r[1]) << 32) | r[0], nfree);
    float cx = static_cast<float>(free_cells[2 * idx]);
    float cy = static_cast<float>(free_cells[2 * idx + 1]);
    *x = cx * res + ox;
    *y = cy * res + oy;
    double u = static_cast<double>(r[2]) * 0x1.0p-32;
    float th = static_cast<float>(-kPi + (2 * kPi) * u);
    if (th > 3.1415925f) th = 3.1415925f;  // keep heading inside [-pi, pi)
    *t = th;
  }
};

namespace {

// distance_calculator.cpp:86-101
std::vector<float> make_distance_cache(float resolution, float max_dist,
                                       int *n_out) {
  long size = static_cast<long>(max_dist / resolution);
  int n = static_cast<int>(size + 2);
  std::vector<float> c(static_cast<size_t>(n) * n);
  for (int col = 0; col < n; ++col)
    for (int row = 0; row < n; ++row)
      c[static_cast<size_t>(row) * n + col] =
          std::sqrt(static_cast<float>(col * col + row * row));
  *n_out = n;
  return c;
}

// distance_calculator.cpp:184-209
void state_map(const int8_t *occ, uint32_t w, uint32_t h,
               std::vector<int8_t> *state) {
  state->resize(static_cast<size_t>(w) * h);
  for (size_t i = 0; i < state->size(); ++i) {
    int8_t v = occ[i];
    int8_t s;
    if (v == -1) s = 2;
    else if (v < 35) s = 0;
    else if (v > 65) s = 1;
    else s = 2;
    (*state)[i] = s;
  }
}

// distance_calculator.cpp:103-182 — the AMCL-style priority-queue brushfire.
// std::priority_queue + the same comparator + the same push order reproduce
// the reference's tie-breaking on the same libstdc++.
struct QCell { long sx, sy, x, y; };
struct QCmp {
  const std::vector<float> *dist;
  uint32_t w;
  bool operator()(const QCell &a, const QCell &b) const {
    return (*dist)[static_cast<size_t>(a.y) * w + a.x] >
           (*dist)[static_cast<size_t>(b.y) * w + b.x];
  }
};
void brushfire(float max_dist, const int8_t *occ, uint32_t w, uint32_t h,
               float resolution, std::vector<float> *out) {
  int cn = 0;
  std::vector<float> cache = make_distance_cache(resolution, max_dist, &cn);
  std::vector<uint8_t> visited(static_cast<size_t>(w) * h, 0);
  out->assign(static_cast<size_t>(w) * h, 0.f);
  std::priority_queue<QCell, std::vector<QCell>, QCmp> q(QCmp{out, w});
  for (long row = 0; row < static_cast<long>(h); ++row)
    for (long col = 0; col < static_cast<long>(w); ++col) {
      size_t i = static_cast<size_t>(row) * w + col;
      int8_t v = occ[i];
      if (v > 65 && v != -1) {
        (*out)[i] = 0;
        q.push(QCell{col, row, col, row});
        visited[i] = 1;
      } else {
        (*out)[i] = max_dist;
      }
    }
  auto visit = [&](const QCell &parent, long x, long y) {
    size_t i = static_cast<size_t>(y) * w + x;
    if (visited[i]) return;
    visited[i] = 1;
    long dx = std::labs(parent.sx - x), dy = std::labs(parent.sy - y);
    float d = cache[static_cast<size_t>(dy) * cn + dx] * resolution;
    if (d > max_dist) return;
    (*out)[i] = d;
    q.push(QCell{parent.sx, parent.sy, x, y});
  };
  while (!q.empty()) {
    QCell c = q.top();
    q.pop();
    if (c.x > 0) visit(c, c.x - 1, c.y);
    if (c.x < static_cast<long>(w) - 1) visit(c, c.x + 1, c.y);
    if (c.y > 0) visit(c, c.x, c.y - 1);
    if (c.y < static_cast<long>(h) - 1) visit(c, c.x, c.y + 1);
  }
}

// Exact squared Euclidean distance transform in integers, separable:
// vertical nearest-obstacle distance per column, then a windowed row
// minimisation.  k = min(dx^2+dy^2) over obstacles with |dx|,|dy| <= R;
// -1 when none.  This is the semantics BASELINE.json's north_star asks the
// kernels to implement ("separable exact Euclidean distance transform").
void exact_d2(const int8_t *occ, uint32_t w, uint32_t h, int R,
              std::vector<int32_t> *out) {
  const int32_t INF = 1 << 29;
  std::vector<int32_t> g(static_cast<size_t>(w) * h);
  for (uint32_t x = 0; x < w; ++x) {
    int32_t d = INF;
    for (uint32_t y = 0; y < h; ++y) {
      size_t i = static_cast<size_t>(y) * w + x;
      d = (occ[i] > 65) ? 0 : (d >= INF ? INF : d + 1);
      g[i] = d;
    }
    d = INF;
    for (uint32_t yy = h; yy-- > 0;) {
      size_t i = static_cast<size_t>(yy) * w + x;
      d = (occ[i] > 65) ? 0 : (d >= INF ? INF : d + 1);
      g[i] = std::min(g[i], d);
    }
  }
  out->assign(static_cast<size_t>(w) * h, -1);
  for (uint32_t y = 0; y < h; ++y)
    for (uint32_t x = 0; x < w; ++x) {
      int64_t best = -1;
      int lo = std::max<int>(0, static_cast<int>(x) - R);
      int hi = std::min<int>(static_cast<int>(w) - 1, static_cast<int>(x) + R);
      for (int xx = lo; xx <= hi; ++xx) {
        int32_t gy = g[static_cast<size_t>(y) * w + xx];
        if (gy > R) continue;
        int64_t dx = static_cast<int64_t>(xx) - x;
        int64_t k = dx * dx + static_cast<int64_t>(gy) * gy;
        if (best < 0 || k < best) best = k;
      }
      (*out)[static_cast<size_t>(y) * w + x] = static_cast<int32_t>(best);
    }
}

// d = sqrtf(k) * res, clipped exactly like distance_calculator.cpp:150-157.
void distance_from_d2(const std::vector<int32_t> &k, float res, float max_dist,
                      std::vector<float> *out) {
  out->resize(k.size());
  for (size_t i = 0; i < k.size(); ++i) {
    float d = max_dist;
    if (k[i] >= 0) {
      float cand = std::sqrt(static_cast<float>(k[i])) * res;
      if (!(cand > max_dist)) d = cand;
    }
    (*out)[i] = d;
  }
}

// map_likelihood.cpp:60.  Eigen's packet exp is within ~1-2 ulp of expf; the
// shared contract uses the correctly rounded value, (float)exp((double)x).
void field_from_distance(const std::vector<float> &d, float z_hit_pre,
                         std::vector<float> *field) {
  field->resize(d.size());
  for (size_t i = 0; i < d.size(); ++i) {
    float arg = -(d[i] * d[i]) / z_hit_pre;
    (*field)[i] = static_cast<float>(std::exp(static_cast<double>(arg)));
  }
}

}  // namespace

// ==================================================================== filter
struct orc_filter {
  orc_map *map = nullptr;
  orc_pf_params params{2, 100, 5000, 2, 0.1, 0.001, 0.01, 0.01, 0.5f,
                       0.17453292f};
  std::vector<orc_particle> sets[2];
  int current = 0;
  Bins bins;
  Rng rng;
  LowPass<double> w_slow, w_fast;
  int trig_mode = ORC_TRIG_CR;
  std::vector<int64_t> last_indices;
  double last_total = 0;

  std::vector<orc_particle> &cur() { return sets[current]; }
  const std::vector<orc_particle> &cur() const { return sets[current]; }
};

namespace {

orc_particle make_particle(float x, float y, float t, double w) {
  orc_particle p;
  p.x = x; p.y = y; p.theta = t; p.pad_ = 0; p.weight = w;
  return p;
}

// particle.cpp:41-64
void normalise_total(std::vector<orc_particle> &ps, double total) {
  for (auto &p : ps) p.weight /= total;
}
void normalise(std::vector<orc_particle> &ps) {
  double total = 0;
  for (const auto &p : ps) total += p.weight;
  normalise_total(ps, total);
}
void equalise(std::vector<orc_particle> &ps) {
  double avg = 1.0 / ps.size();
  for (auto &p : ps) p.weight = avg;
}

// map_likelihood.cpp:63-130
double update_importance(const orc_map &m, const float *xy, size_t npoints,
                         orc_particle *ps, size_t n, int trig_mode) {
  size_t step = 1;
  if (m.params.num_beams)
    step = npoints / static_cast<size_t>(m.params.num_beams);
  if (step < 1) step = 1;
  const float zr = m.params.z_rand / m.params.max_laser_distance;
  double total = 0;
  for (size_t k = 0; k < n; ++k) {
    orc_particle &p = ps[k];
    // pose_2d.h:118-126: translate then rotate, float sin/cos.
    float c, s;
    if (trig_mode == ORC_TRIG_LIBM) {
      c = std::cos(p.theta);
      s = std::sin(p.theta);
    } else {
      c = static_cast<float>(std::cos(static_cast<double>(p.theta)));
      s = static_cast<float>(std::sin(static_cast<double>(p.theta)));
    }
    double weight = 0;
    for (size_t i = 0; i < npoints; i += step) {
      float px = xy[2 * i], py = xy[2 * i + 1];
      // Eigen Isometry2f * Vector2f: result starts as the translation and
      // the 2-term linear product is added to it.
      float mx = p.x + (c * px + (-s) * py);
      float my = p.y + (s * px + c * py);
      double pz = m.probability_at(mx, my);
      double pw = m.params.z_hit * pz + zr;
      weight += pw * pw * pw;
    }
    total += weight;
    int8_t st = m.state_at(p.x, p.y);
    if (st == 2) weight = 0;
    else if (st == 1) weight *= 0.01;
    p.weight = weight;
  }
  return total;
}

// odometry.cpp:44-102 with the shared RNG contract (one Philox block per
// particle: words 0-1 -> (trans, rot1) normals, words 2-3 -> (rot2, unused)).
void sample_odometry(const double nw[3], const double od[3],
                     std::vector<orc_particle> &ps, const Rng &rng,
                     const float alpha[4]) {
  double xdiff = nw[0] - od[0];
  double ydiff = nw[1] - od[1];
  double delta_trans = std::sqrt(xdiff * xdiff + ydiff * ydiff);
  double delta_rot1 = 0.0;
  if (delta_trans >= 0.01)
    delta_rot1 = angle_delta_d(std::atan2(ydiff, xdiff), od[2]);
  double delta_theta_change = angle_delta_d(nw[2], od[2]);
  double delta_rot2 = angle_delta_d(delta_theta_change, delta_rot1);
  double r1n = std::min(std::fabs(delta_rot1),
                        std::fabs(angle_delta_d(kPi, delta_rot1)));
  double r2n = std::min(std::fabs(delta_rot2),
                        std::fabs(angle_delta_d(kPi, delta_rot2)));
  double trans_var = alpha[2] * delta_trans * delta_trans +
                     alpha[3] * r1n * r1n + alpha[3] * r2n * r2n;
  double rot1_var = alpha[0] * r1n * r1n + alpha[1] * delta_trans * delta_trans;
  double rot2_var = alpha[0] * r2n * r2n + alpha[1] * delta_trans * delta_trans;
  double sd_t = std::sqrt(trans_var), sd_r1 = std::sqrt(rot1_var),
         sd_r2 = std::sqrt(rot2_var);
  for (size_t i = 0; i < ps.size(); ++i) {
    uint32_t r[4];
    rng.draw(i, P_MOTION, r);
    double z0, z1, z2, z3;
    normal_pair(r[0], r[1], &z0, &z1);
    normal_pair(r[2], r[3], &z2, &z3);
    double trans_hat = delta_trans - z0 * sd_t;
    double rot1_hat = angle_delta_d(delta_rot1, z1 * sd_r1);
    double rot2_hat = angle_delta_d(delta_rot2, z2 * sd_r2);
    orc_particle &p = ps[i];
    float th = p.theta;
    p.x += static_cast<float>(trans_hat * std::cos(th + rot1_hat));
    p.y += static_cast<float>(trans_hat * std::sin(th + rot1_hat));
    p.theta += static_cast<float>(rot1_hat + rot2_hat);
    p.theta = norm_angle_f(p.theta);
  }
}

// Symmetric 3x3 eigen-decomposition (cyclic Jacobi, double), eigenvalues
// ascending like Eigen's SelfAdjointEigenSolver (random.h:60-62).  Eigenvector
// signs are not canonical in either implementation.
void cov_factor(const float cov[9], float out[9]) {
  double a[3][3], v[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) a[i][j] = 0.5 * (static_cast<double>(cov[i * 3 + j]) + cov[j * 3 + i]);
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
    if (off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        if (a[p][q] == 0.0) continue;
        double theta = (a[q][q] - a[p][p]) / (2 * a[p][q]);
        double t = (theta >= 0 ? 1.0 : -1.0) /
                   (std::fabs(theta) + std::sqrt(theta * theta + 1));
        double c = 1 / std::sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < 3; ++k) {
          double akp = a[k][p], akq = a[k][q];
          a[k][p] = c * akp - s * akq;
          a[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) {
          double apk = a[p][k], aqk = a[q][k];
          a[p][k] = c * apk - s * aqk;
          a[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 3; ++k) {
          double vkp = v[k][p], vkq = v[k][q];
          v[k][p] = c * vkp - s * vkq;
          v[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int order[3] = {0, 1, 2};
  std::sort(order, order + 3, [&](int i, int j) { return a[i][i] < a[j][j]; });
  for (int col = 0; col < 3; ++col) {
    int e = order[col];
    double lam = a[e][e] > 0 ? a[e][e] : 0.0;
    double sq = std::sqrt(lam);
    for (int row = 0; row < 3; ++row)
      out[row * 3 + col] = static_cast<float>(v[row][e] * sq);
  }
}

// particle_filter.cpp:163-175
void do_cluster(orc_filter *f) {
  f->bins.reset();
  for (const auto &p : f->cur()) f->bins.add_point(p.x, p.y, p.theta);
  f->bins.assign_clusters();
}

// particle_filter.cpp:214-239.  Deviation: the reference's
// .at(i) throws when the walk runs off the end; here i is clamped to M-1.
void resample_low_variance(orc_filter *f,
                           const std::vector<orc_particle> &in,
                           std::vector<orc_particle> *out) {
  const size_t M = in.size();
  out->resize(M);
  f->last_indices.assign(M, 0);
  if (M == 0) return;
  double minv = 1.0f / M;  // float reciprocal widened to double
  uint32_t r4[4];
  f->rng.draw(0, P_LV_OFFSET, r4);
  double r = u53(r4[0], r4[1]) * minv;
  double c = in[0].weight;
  size_t i = 0;
  for (size_t m = 0; m < M; ++m) {
    double U = r + m * minv;
    while (U > c && i + 1 < M) {
      i++;
      c += in[i].weight;
    }
    (*out)[m] = in[i];
    f->last_indices[m] = static_cast<int64_t>(i);
  }
}

// One adaptive/KLD slot, particle_filter.cpp:261-276 / :313-327.
orc_particle draw_slot(orc_filter *f, const std::vector<orc_particle> &in,
                       const RunningSum &cdf, double w_diff, size_t m,
                       int64_t *src) {
  uint32_t r[4];
  f->rng.draw(m, P_DRAW, r);
  double p = u53(r[0], r[1]);
  if (p < w_diff) {
    uint32_t q[4];
    f->rng.draw(m, P_INJECT, q);
    float x, y, t;
    f->map->random_pose(q, &x, &y, &t);
    *src = -1;
    return make_particle(x, y, t, 0.0);
  }
  double rr = u53(r[2], r[3]);
  size_t idx = cdf.lookup(rr);
  *src = static_cast<int64_t>(idx);
  return in[idx];
}

double w_diff_of(const orc_filter *f) {  // particle_filter.cpp:246,298
  return std::max(0.0, 1.0 - f->w_fast.value / f->w_slow.value);
}

// particle_filter.cpp:241-291
void resample_adaptive(orc_filter *f, const std::vector<orc_particle> &in,
                       std::vector<orc_particle> *out) {
  const double w_diff = w_diff_of(f);
  const size_t n_out = static_cast<size_t>(f->params.particle_count_min);
  out->resize(n_out, make_particle(0, 0, 0, 1.0));
  f->last_indices.clear();
  RunningSum cdf(in);
  if (cdf.total <= 0) return;  // reference logs an error and returns
  f->last_indices.resize(n_out);
  for (size_t m = 0; m < n_out; ++m)
    (*out)[m] = draw_slot(f, in, cdf, w_diff, m, &f->last_indices[m]);
  if (w_diff > 0.0) {
    f->w_fast.value = 0;
    f->w_slow.value = 0;
  }
}

// particle_filter.cpp:293-348
void resample_kld(orc_filter *f, const std::vector<orc_particle> &in,
                  std::vector<orc_particle> *out) {
  f->bins.reset();
  const double w_diff = w_diff_of(f);
  const size_t n_max = static_cast<size_t>(f->params.particle_count_max);
  out->clear();
  out->reserve(n_max);
  f->last_indices.clear();
  RunningSum cdf(in);
  if (cdf.total <= 0) return;
  for (size_t m = 0; m < n_max; ++m) {
    int64_t src;
    orc_particle p = draw_slot(f, in, cdf, w_diff, m, &src);
    out->push_back(p);
    f->last_indices.push_back(src);
    f->bins.add_point(p.x, p.y, p.theta);
    if (m >= f->bins.limit()) break;
  }
  if (w_diff > 0.0) {
    f->w_fast.value = 0;
    f->w_slow.value = 0;
  }
  f->bins.assign_clusters();
}

}  // namespace

// ===================================================================== C API
extern "C" {

double orc_last_seconds(void) { return g_last_seconds; }

// laser_handler.cpp:108-121 (laser_geometry projection + mount transform).
size_t orc_project_scan(const float *ranges, size_t n, float angle_min,
                        float angle_increment, float range_min,
                        float range_max, const double mount[3],
                        float *out_xy) {
  const double cy = std::cos(mount[2]), sy = std::sin(mount[2]);
  size_t k = 0;
  for (size_t i = 0; i < n; ++i) {
    const double r = ranges[i];
    if (!(r < static_cast<double>(range_max) &&
          r >= static_cast<double>(range_min)))
      continue;
    const double a = static_cast<double>(angle_min) +
                     static_cast<double>(i) * static_cast<double>(angle_increment);
    const float px = static_cast<float>(r * std::cos(a));
    const float py = static_cast<float>(r * std::sin(a));
    // tf2::Transform: basis row . v + origin, in double
    const double x = (cy * px + (-sy) * py) + mount[0];
    const double y = (sy * px + cy * py) + mount[1];
    out_xy[2 * k] = static_cast<float>(x);
    out_xy[2 * k + 1] = static_cast<float>(y);
    ++k;
  }
  return k;
}

void orc_distance_cache(float resolution, float max_dist, int *size_out,
                        float *out, size_t cap) {
  int n = 0;
  std::vector<float> c = make_distance_cache(resolution, max_dist, &n);
  *size_out = n;
  for (size_t i = 0; i < c.size() && i < cap; ++i) out[i] = c[i];
}
void orc_world_to_map(const float *pos, const float *offset, const float *res,
                      int dims, int *out) {
  for (int d = 0; d < dims; ++d)
    out[d] = offset ? cell_of(pos[d], offset[d], res[d])
                    : cell_of_inf(pos[d], res[d]);
}
void orc_map_to_world(const int *cell, const float *offset, const float *res,
                      int dims, float *out) {
  for (int d = 0; d < dims; ++d) {
    float v = static_cast<float>(cell[d]) * res[d];  // scaled_map.h:51-55
    if (offset) v = v + offset[d];                   // :101-105
    out[d] = v;
  }
}
int orc_is_in_map(const int *cell, const int *size, int dims) {
  for (int d = 0; d < dims; ++d)
    if (!(0 <= cell[d] && cell[d] < size[d])) return 0;
  return 1;
}
double orc_normalise_angle_f64(double a) { return norm_angle_d(a); }
float orc_normalise_angle_f32(float a) { return norm_angle_f(a); }
double orc_angle_delta_f64(double a, double b) { return angle_delta_d(a, b); }
float orc_angle_delta_f32(float a, float b) { return angle_delta_f(a, b); }

void orc_transform_point_f32(const float pose[3], const float pt[2],
                             float out[2], int trig_mode) {
  float c, s;
  if (trig_mode == ORC_TRIG_LIBM) {
    c = std::cos(pose[2]);
    s = std::sin(pose[2]);
  } else {
    c = static_cast<float>(std::cos(static_cast<double>(pose[2])));
    s = static_cast<float>(std::sin(static_cast<double>(pose[2])));
  }
  out[0] = pose[0] + (c * pt[0] + (-s) * pt[1]);
  out[1] = pose[1] + (s * pt[0] + c * pt[1]);
}
void orc_to_decomposed(const float pose[3], double out[4]) {
  out[0] = pose[0];
  out[1] = pose[1];
  out[2] = std::sin(static_cast<double>(pose[2]));
  out[3] = std::cos(static_cast<double>(pose[2]));
}
double orc_effective_particles(const orc_particle *p, size_t n) {
  double sum = 0;  // particle.cpp:32-39
  for (size_t i = 0; i < n; ++i) sum += p[i].weight * p[i].weight;
  return 1.0 / sum;
}
void orc_normalise_weights(orc_particle *p, size_t n) {
  double total = 0;
  for (size_t i = 0; i < n; ++i) total += p[i].weight;
  for (size_t i = 0; i < n; ++i) p[i].weight /= total;
}
void orc_normalise_weights_total(orc_particle *p, size_t n, double total) {
  for (size_t i = 0; i < n; ++i) p[i].weight /= total;
}
void orc_equalise_weights(orc_particle *p, size_t n) {
  double avg = 1.0 / n;
  for (size_t i = 0; i < n; ++i) p[i].weight = avg;
}
double orc_running_sum_lookup(const orc_particle *p, size_t n,
                              const double *queries, size_t nq,
                              uint64_t *idx_out) {
  std::vector<orc_particle> v(p, p + n);
  RunningSum rs(v);
  for (size_t i = 0; i < nq; ++i) idx_out[i] = rs.lookup(queries[i]);
  return rs.total;
}
void orc_statistics(const orc_particle *p, size_t n, int finalise,
                    double out[15]) {
  Stats s;
  for (size_t i = 0; i < n; ++i) s.add(p[i]);
  if (finalise) s.finalise();
  s.pack(out);
}
void orc_statistics_merge(const double a[15], const double b[15], int finalise,
                          double out[15]) {
  Stats sa, sb;
  sa.unpack(a);
  sb.unpack(b);
  sa.add(sb);
  if (finalise) sa.finalise();
  sa.pack(out);
}
double orc_lowpass_run(double alpha, const double *xs, size_t n,
                       int as_float) {
  if (as_float) {
    LowPass<float> f;
    f.alpha = static_cast<float>(alpha);
    for (size_t i = 0; i < n; ++i) f.update(static_cast<float>(xs[i]));
    return f.value;
  }
  LowPass<double> f;
  f.alpha = alpha;
  for (size_t i = 0; i < n; ++i) f.update(xs[i]);
  return f.value;
}
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2],
                       uint32_t out[4]) {
  philox(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}
void orc_rng_words(uint64_t seed, uint32_t step, uint32_t purpose,
                   uint64_t first_index, size_t count, uint32_t *out4) {
  Rng rng;
  rng.seed = seed;
  rng.step = step;
  for (size_t i = 0; i < count; ++i) rng.draw(first_index + i, purpose, out4 + 4 * i);
}
void orc_normal_pairs(const uint32_t *words2, size_t count, double *out2) {
  for (size_t i = 0; i < count; ++i)
    normal_pair(words2[2 * i], words2[2 * i + 1], &out2[2 * i], &out2[2 * i + 1]);
}
uint64_t orc_kld_limit(uint64_t k, double eps, double z, int32_t nmin,
                       int32_t nmax) {
  return Bins::limit_for(k, 2 * eps, z, nmin, nmax);
}

// ------------------------------------------------------------------- map API
orc_map *orc_map_create(void) {
  orc_map *m = new orc_map();
  orc_map_set_params(m, &m->params);
  return m;
}
void orc_map_destroy(orc_map *m) { delete m; }

// map_likelihood.cpp:40-49 (all float; exp via the correctly-rounded contract)
void orc_map_set_params(orc_map *m, const orc_likelihood_params *p) {
  m->params = *p;
  m->z_hit_pre = 2 * p->sigma_hit * p->sigma_hit;
  float arg =
      -(p->max_obstacle_distance * p->max_obstacle_distance) / m->z_hit_pre;
  m->max_distance_probability =
      static_cast<float>(std::exp(static_cast<double>(arg)));
}

// map_base.cpp:30-56 + map_likelihood.cpp:51-61
void orc_map_load(orc_map *m, const int8_t *occ, uint32_t w, uint32_t h,
                  float res, double ox, double oy, int field_mode) {
  ScopedTimer t;
  m->w = w; m->h = h; m->res = res;
  m->ox = static_cast<float>(ox);
  m->oy = static_cast<float>(oy);
  state_map(occ, w, h, &m->state);
  m->rebuild_free_cells();
  const float max_dist = m->params.max_obstacle_distance;
  if (field_mode == ORC_FIELD_BRUSHFIRE) {
    brushfire(max_dist, occ, w, h, res, &m->distance);
  } else {
    int R = static_cast<int>(max_dist / res) + 1;
    std::vector<int32_t> k;
    exact_d2(occ, w, h, R, &k);
    distance_from_d2(k, res, max_dist, &m->distance);
  }
  field_from_distance(m->distance, m->z_hit_pre, &m->field);
}
void orc_map_load_field(orc_map *m, const float *field, const int8_t *state,
                        uint32_t w, uint32_t h, float res, double ox,
                        double oy) {
  m->w = w; m->h = h; m->res = res;
  m->ox = static_cast<float>(ox);
  m->oy = static_cast<float>(oy);
  m->field.assign(field, field + static_cast<size_t>(w) * h);
  m->state.assign(state, state + static_cast<size_t>(w) * h);
  m->distance.clear();
  m->rebuild_free_cells();
}
void orc_map_copy_distance(const orc_map *m, float *out) {
  std::memcpy(out, m->distance.data(), m->distance.size() * sizeof(float));
}
void orc_map_copy_field(const orc_map *m, float *out) {
  std::memcpy(out, m->field.data(), m->field.size() * sizeof(float));
}
void orc_map_copy_state(const orc_map *m, int8_t *out) {
  std::memcpy(out, m->state.data(), m->state.size());
}
size_t orc_map_num_free(const orc_map *m) { return m->free_cells.size() / 2; }
void orc_map_copy_free_cells(const orc_map *m, int32_t *xy_out) {
  std::memcpy(xy_out, m->free_cells.data(),
              m->free_cells.size() * sizeof(int32_t));
}
int8_t orc_map_state_at(const orc_map *m, float x, float y) {
  return m->state_at(x, y);
}
// map_likelihood.cpp:132-146
void orc_map_likelihood_grid(const orc_map *m, int8_t *out) {
  for (size_t i = 0; i < m->field.size(); ++i)
    out[i] = static_cast<int8_t>(100 * m->field[i] + 0.5f);
}
double orc_map_update_importance(const orc_map *m, const float *xy,
                                 size_t npoints, orc_particle *p, size_t n,
                                 int trig_mode) {
  ScopedTimer t;
  return update_importance(*m, xy, npoints, p, n, trig_mode);
}
void orc_bruteforce_d2(const int8_t *occ, uint32_t w, uint32_t h, int radius,
                       int32_t *out) {
  for (int y = 0; y < static_cast<int>(h); ++y)
    for (int x = 0; x < static_cast<int>(w); ++x) {
      int64_t best = -1;
      for (int yy = std::max(0, y - radius);
           yy <= std::min(static_cast<int>(h) - 1, y + radius); ++yy)
        for (int xx = std::max(0, x - radius);
             xx <= std::min(static_cast<int>(w) - 1, x + radius); ++xx)
          if (occ[static_cast<size_t>(yy) * w + xx] > 65) {
            int64_t k = static_cast<int64_t>(xx - x) * (xx - x) +
                        static_cast<int64_t>(yy - y) * (yy - y);
            if (best < 0 || k < best) best = k;
          }
      out[static_cast<size_t>(y) * w + x] = static_cast<int32_t>(best);
    }
}

// ---------------------------------------------------------------- filter API
orc_filter *orc_filter_create(orc_map *m, uint64_t seed) {
  orc_filter *f = new orc_filter();
  f->map = m;
  f->rng.seed = seed;
  orc_filter_set_params(f, &f->params);
  return f;
}
void orc_filter_destroy(orc_filter *f) { delete f; }

// particle_filter.cpp:177-184
void orc_filter_set_params(orc_filter *f, const orc_pf_params *p) {
  f->params = *p;
  f->bins.configure(*p);
  f->w_fast.alpha = p->alpha_fast;
  f->w_slow.alpha = p->alpha_slow;
}
void orc_filter_set_rng(orc_filter *f, uint64_t seed, uint32_t step) {
  f->rng.seed = seed;
  f->rng.step = step;
}
void orc_filter_set_trig_mode(orc_filter *f, int trig_mode) {
  f->trig_mode = trig_mode;
}
void orc_cov_factor(const float cov[9], float out[9]) { cov_factor(cov, out); }

// particle_filter.cpp:41-69 + random.h:64-70.  z = 3 standard normals
// (float), v = mean + T*z with the 3-term products summed as a0+(a1+a2)
// (Eigen's unrolled redux split).
void orc_filter_initialise_factor(orc_filter *f, const float pose[3],
                                  const float T[9]) {
  ScopedTimer t;
  size_t n = static_cast<size_t>(f->params.particle_count_min);
  auto &cloud = f->cur();
  cloud.clear();
  cloud.reserve(n);
  double weight = 1.0 / n;
  for (size_t i = 0; i < n; ++i) {
    uint32_t r[4];
    f->rng.draw(i, P_INIT, r);
    double d0, d1, d2, d3;
    normal_pair(r[0], r[1], &d0, &d1);
    normal_pair(r[2], r[3], &d2, &d3);
    float z[3] = {static_cast<float>(d0), static_cast<float>(d1),
                  static_cast<float>(d2)};
    float v[3];
    for (int k = 0; k < 3; ++k) {
      float prod = T[k * 3 + 0] * z[0] + (T[k * 3 + 1] * z[1] + T[k * 3 + 2] * z[2]);
      v[k] = pose[k] + prod;
    }
    cloud.push_back(make_particle(v[0], v[1], norm_angle_f(v[2]), weight));
  }
  f->rng.step++;
}
void orc_filter_initialise(orc_filter *f, const float pose[3],
                           const float cov[9]) {
  float T[9];
  cov_factor(cov, T);
  orc_filter_initialise_factor(f, pose, T);
}
// particle_filter.cpp:71-83
void orc_filter_global_localization(orc_filter *f) {
  ScopedTimer t;
  size_t n = static_cast<size_t>(f->params.particle_count_max);
  auto &cloud = f->cur();
  cloud.clear();
  cloud.reserve(n);
  double weight = 1.0 / n;
  for (size_t i = 0; i < n; ++i) {
    uint32_t r[4];
    f->rng.draw(i, P_GLOBAL, r);
    float x, y, th;
    f->map->random_pose(r, &x, &y, &th);
    cloud.push_back(make_particle(x, y, th, weight));
  }
  f->rng.step++;
}
void orc_filter_handle_odometry(orc_filter *f, const double odom_new[3],
                                const double odom_old[3],
                                const float alpha[4]) {
  ScopedTimer t;
  sample_odometry(odom_new, odom_old, f->cur(), f->rng, alpha);
  f->rng.step++;
}
// particle_filter.cpp:95-136
void orc_filter_update_sensor(orc_filter *f, const float *xy, size_t npoints) {
  ScopedTimer t;
  auto &ps = f->cur();
  double total = update_importance(*f->map, xy, npoints, ps.data(), ps.size(),
                                   f->trig_mode);
  f->last_total = total;
  if (total > 0) {
    normalise_total(ps, total);
    double avg = total / ps.size();
    f->w_slow.update(avg);
    f->w_fast.update(avg);
  } else {
    equalise(ps);
  }
}
// particle_filter.cpp:138-161
void orc_filter_resample_and_cluster(orc_filter *f) {
  ScopedTimer t;
  const auto &in = f->sets[f->current];
  auto &out = f->sets[(f->current + 1) % 2];
  f->current = (f->current + 1) % 2;
  switch (f->params.resample_type) {
  case 0:
    resample_low_variance(f, in, &out);
    do_cluster(f);
    break;
  case 1:
    resample_adaptive(f, in, &out);
    do_cluster(f);
    break;
  default:
    resample_kld(f, in, &out);
    break;
  }
  f->rng.step++;
  normalise(f->cur());
}
void orc_filter_cluster(orc_filter *f) {
  ScopedTimer t;
  do_cluster(f);
}
// particle_filter.cpp:186-212 + :366-391.  Returns 1 valid, 0 no good
// cluster, -1 a particle has no cluster (the reference abort()s).
int orc_filter_get_estimate(const orc_filter *f, double pose[3],
                            double cov[9]) {
  ScopedTimer t;
  std::vector<Stats> clusters(f->bins.cluster_count);
  for (const auto &p : f->cur()) {
    int32_t id = f->bins.cluster_of(p.x, p.y, p.theta);
    if (id < 0) return -1;
    clusters[id].add(p);
  }
  Stats global;
  for (auto &c : clusters) {
    global.add(c);
    c.finalise();
  }
  global.finalise();
  double best_w = -1;
  const Stats *best = nullptr;
  for (const auto &c : clusters)
    if (best_w < c.total_weight) {
      best_w = c.total_weight;
      best = &c;
    }
  if (best && best_w > 0) {
    pose[0] = best->mean[0];
    pose[1] = best->mean[1];
    pose[2] = best->mean[2];
    std::memcpy(cov, global.cov, sizeof(double) * 9);
    return 1;
  }
  return 0;
}
size_t orc_filter_num_particles(const orc_filter *f) { return f->cur().size(); }
void orc_filter_copy_particles(const orc_filter *f, orc_particle *out,
                               size_t cap) {
  size_t n = std::min(cap, f->cur().size());
  std::memcpy(out, f->cur().data(), n * sizeof(orc_particle));
}
void orc_filter_set_particles(orc_filter *f, const orc_particle *p, size_t n) {
  f->cur().assign(p, p + n);
}
size_t orc_filter_copy_resample_indices(const orc_filter *f, int64_t *out,
                                        size_t cap) {
  size_t n = std::min(cap, f->last_indices.size());
  std::memcpy(out, f->last_indices.data(), n * sizeof(int64_t));
  return f->last_indices.size();
}
size_t orc_filter_num_bins(const orc_filter *f) { return f->bins.cells.size(); }
size_t orc_filter_copy_bins(const orc_filter *f, int32_t *keys4,
                            int32_t *labels, size_t cap) {
  // canonical label: index (in sorted-key order) of the first bin of the
  // component.  assign_clusters() numbers components in that order already.
  std::vector<int32_t> first(f->bins.cluster_count, -1);
  size_t i = 0;
  for (const auto &kv : f->bins.cells) {
    if (i >= cap) break;
    keys4[4 * i + 0] = kv.first.x;
    keys4[4 * i + 1] = kv.first.y;
    keys4[4 * i + 2] = kv.first.t;
    keys4[4 * i + 3] = static_cast<int32_t>(kv.second.count);
    int32_t lab = kv.second.label;
    if (lab >= 0) {
      if (first[lab] < 0) first[lab] = static_cast<int32_t>(i);
      labels[i] = first[lab];
    } else {
      labels[i] = -1;
    }
    ++i;
  }
  return f->bins.cells.size();
}
size_t orc_filter_num_clusters(const orc_filter *f) {
  return f->bins.cluster_count;
}
void orc_filter_get_lowpass(const orc_filter *f, double out[2]) {
  out[0] = f->w_slow.value;
  out[1] = f->w_fast.value;
}
void orc_filter_set_lowpass(orc_filter *f, double w_slow, double w_fast) {
  f->w_slow.value = w_slow;
  f->w_fast.value = w_fast;
}
double orc_filter_last_total_weight(const orc_filter *f) {
  return f->last_total;
}

}  // extern "C"
