// state.cuh — the two opaque handles behind the C ABI.
#pragma once

#include "common.cuh"

#include <vector>

// Device-resident scalars of one filter.
struct FilterScalars {
  double total_weight;   // pre-penalty total of the last sensor update
  double post_total;     // sum of weights used by normalise()
  unsigned int blocks_done;
  unsigned int pad0;
  unsigned long long n_out;  // KLD output size
  unsigned long long n_clusters;  // label_clusters' count (read on the device by K14)
  unsigned long long n_bins;      // build_bins' count (read on the device by the labelling)
  unsigned long long n_records;   // running-sum records (read on the device by the draws)
  double sum_sq;                  // sum of squared weights (effective_particles)
  // Results of the multi-launch KLD scan.  They live here, not behind the scan's
  // status words: a later, longer scan over the same buffer would read a stop
  // word {m:32, k:32} left there as a tile status.
  unsigned long long kld_new;     // new-bin flags over all candidates
  unsigned long long kld_stop;    // stop word
};

// quickmcl::MapLikelihood (map_likelihood.h:35-93) + MapBase (map_base.h:31-69)
struct qmcl_map {
  int device = 0;
  cudaStream_t stream = nullptr;

  qmcl_likelihood_params params{0.5f, 0.5f, 30, 14.0f, 2.0f, 0.01f};
  // map_likelihood.cpp:45-48, refreshed by set_params
  float z_hit_pre = 2 * 0.01f * 0.01f;
  double max_distance_probability = 0.0;

  bool has_map = false;
  bool has_distance = false;
  qmcl::MapGeom geom{0, 0, 0.f, 0.f, 0.f};

  // HBM layout (row-major [y][x], as map_types.h:27-32):
  //   field    float32  W*H   likelihood table exp(-d^2 / 2 sigma^2)
  //   distance float32  W*H   clipped distance in metres (only after load())
  //   state    int8     W*H   0 free / 1 occupied / 2 unknown (map_state.h)
  //   free     uint32   nfree linear indices (y*w + x) of the free cells, row-major order
  qmcl::DevBuf<float> field;
  qmcl::DevBuf<float> distance;
  qmcl::DevBuf<int8_t> state;
  qmcl::DevBuf<uint32_t> free_cells;
  uint64_t n_free = 0;

  // scratch for load()
  qmcl::DevBuf<int8_t> occ;
  qmcl::DevBuf<int32_t> edt_g;
  qmcl::DevBuf<uint32_t> scan_tmp;
  qmcl::DevBuf<float> lut;  // d^2 -> distance | likelihood

  // sensor-model palette path: uint16 d^2 code per cell + pw^3 table
  qmcl::DevBuf<uint16_t> code;
  qmcl::DevBuf<double> pw3_lut;
  qmcl::DevBuf<float> pw3_lut_f32;
  bool has_code = false;
  int n_codes = 0;
  // PAL8 (the default value path when it applies): when the float pw^3 table
  // holds at most 254 distinct values over the d^2 codes that can occur, the
  // map is also kept as one byte per cell (index into that palette), 16 x 8
  // cells per 128-byte line.  Same floats as PAL32, hence the same weights.
  qmcl::DevBuf<uint8_t> code8;
  qmcl::DevBuf<uint8_t> pal8_rank;   // d^2 code -> palette index (255 = cannot occur)
  qmcl::DevBuf<float> pal8_lut;      // [0, pal8_n) palette, [pal8_n] outside the map, [pal8_n+1] 0
  int pal8_n = 0;                    // 0 = not available for these parameters
  bool pal8_valid = false;           // built for the current pw^3 table
  int div_verified = 0;           // cheapest division form verified against IEEE division at load
  int div_forced = -1;            // qmcl_map_set_division_mode: -1 auto, else an upper bound
  int sensor_path = 0;            // 0 auto (3, else 4), 1 float field, 2 exact double palette, 3 one-byte palette, 4 float table
  bool lut_params_valid = false;
  float lut_z_hit = 0.f, lut_zr = 0.f;
  double lut_p_max = 0.0;

  // qmcl_map_load: the occupancy grid is copied in row slabs on copy_stream while
  // earlier slabs are being transformed on `stream`
  // (fast path: column passes on col_stream, row passes on `stream`, free-cell list on aux_stream)
  cudaStream_t copy_stream = nullptr, col_stream = nullptr, aux_stream = nullptr;
  std::vector<cudaEvent_t> slab_events, col_events;
  qmcl::PinnedBuf<unsigned long long> h_count;  // [0] free-cell count, [1] division self-check flags
  qmcl::DevBuf<unsigned long long> free_totals; // running free-cell count after each slab, then the flags

  int refs = 1;  // the creator + one per filter

  // scratch for the host-vector update_importance overload
  qmcl::DevBuf<float> tmp_x, tmp_y, tmp_t;
  qmcl::DevBuf<double> tmp_w;
  qmcl::DevBuf<uint8_t> tmp_aos;
  qmcl::DevBuf<float> tmp_cloud;
  qmcl::DevBuf<FilterScalars> tmp_scalars;
};

struct ParticleSet {
  // SoA in HBM: 12 B pose + 8 B weight per particle.
  qmcl::DevBuf<float> x, y, t;
  qmcl::DevBuf<double> w;
  size_t n = 0;
  qmcl_status ensure(size_t cap) {
    QMCL_TRY(x.ensure(cap));
    QMCL_TRY(y.ensure(cap));
    QMCL_TRY(t.ensure(cap));
    QMCL_TRY(w.ensure(cap));
    return QMCL_OK;
  }
};

namespace qmcl {
struct TailResult;
struct TailDev;
}
using qmcl::TailDev;
using qmcl::TailResult;

// Dense bin grid: bounding box of the occupied KLD/cluster bins.
struct BinGrid {
  int x0, y0, t0;
  int nx, ny, nt;
};

// CUDA-event stage timers (qmcl_filter_set_timing / get_timing).
struct StageTimer {
  int level = 0;
  struct Pending { int stage; cudaEvent_t a, b; };
  std::vector<cudaEvent_t> pool;
  std::vector<Pending> pending;
  double ms[QMCL_N_STAGES] = {};
  uint64_t count[QMCL_N_STAGES] = {};
  cudaEvent_t take() {
    if (!pool.empty()) {
      cudaEvent_t e = pool.back();
      pool.pop_back();
      return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
  }
  // Folds finished pairs into ms[]; the stream must be idle.
  void resolve() {
    for (const Pending &p : pending) {
      float t = 0.f;
      if (cudaEventElapsedTime(&t, p.a, p.b) == cudaSuccess) {
        ms[p.stage] += t;
        count[p.stage]++;
      }
      pool.push_back(p.a);
      pool.push_back(p.b);
    }
    pending.clear();
  }
  void release() {
    resolve();
    for (cudaEvent_t e : pool) cudaEventDestroy(e);
    pool.clear();
  }
};

// Slots of qmcl_filter::h_scalars (pinned host memory, 8-byte words).  The
// deferred values own their slots; the others are consumed right after the
// synchronisation that follows the copy.
enum : int {
  HS_TOTAL = 0,      // pre-penalty sensor total (deferred)
  HS_NCLUSTERS = 1,  // cluster count (deferred)
  HS_RS = 2,         // running sum: record count, total weight
  HS_BBOX = 4,       // 6 ints
  HS_STOP = 8,       // KLD: packed stop word, number of new-bin flags
  HS_NBINS = 10,
  HS_BBOX_EARLY = 12,  // 6 ints: bin bounds taken ahead of the sensor kernel (KLD)
  HS_EST = 16,       // estimate: 13 doubles + invalid flag
  HS_COUNT = 32
};

// quickmcl::ParticleFilter (particle_filter.h)
struct qmcl_filter {
  qmcl_map *map = nullptr;
  qmcl_pf_params params{QMCL_RESAMPLE_KLD, 100, 5000, 2, 0.1, 0.001, 0.01, 0.01, 0.5f,
                        0.17453292f};
  uint64_t seed = 0;
  uint32_t step = 0;
  int shard_rank = 0, n_shards = 1;
  qmcl_comm comm{0, 1, nullptr, nullptr, nullptr, nullptr};
  bool rebalance = true;
  double rebalance_threshold = 0.0;  // relative imbalance below which the all-to-all is skipped (0: never)
  uint64_t n_rebalance_skipped = 0;
  qmcl::DevBuf<uint8_t> xchg;        // collective payloads
  qmcl::DevBuf<uint8_t> xchg_send, xchg_recv;  // particle rebalancing (AoS)
  uint64_t n_collectives = 0;
  uint64_t n_chain_rounds = 0;  // carry-chain rounds that needed an exchange
  size_t global_first = 0;  // global index of local particle 0
  size_t global_n = 0;      // particles over all shards

  ParticleSet sets[2];
  int current = 0;

  // low_pass_filter.h: two host scalars
  double w_slow = 0.0, w_fast = 0.0;
  double last_total = 0.0;
  // roll-back of a failed resample_and_cluster (resample.cu)
  bool lowpass_was_reset = false;
  double w_slow_before_reset = 0.0, w_fast_before_reset = 0.0;

  // Deferred 8-byte read-backs (single-GPU filters only).  A value the host
  // needs only later — the sensor total for the two low-pass scalars, the
  // cluster count — is copied to its own pinned slot without a stream
  // synchronisation and folded in by settle()/settle_after_sync() at the next
  // point that synchronises anyway (filter.cu).  lazy_sync = false restores a
  // synchronisation per value (qmcl_filter_set_sync_mode).
  bool lazy_sync = true;
  // Fused tail (tail.cu): one cooperative launch for everything after the sensor
  // kernel.  While its result block is in flight the host knows the outcome of
  // resample_and_cluster / cluster only as an upper bound (tail_pending); every
  // entry point settles first.
  bool use_tail = true;
  bool motion_pending = false;      // K6 of the last handle_odometry deferred into the sensor update's front kernel
  double pending_motion[6] = {};    // its MotionParams (motion_dev.cuh)
  uint32_t pending_motion_step = 0; // the RNG step it was called at
  bool normalise_pending = false;   // K8 (divide by the sensor total) deferred into the tail
  size_t normalise_n = 0;           // particles of that sensor update
  bool tail_pending = false;
  int tail_mode = 0;
  bool tail_had_normalise = false;
  double tail_n_total = 0.0;
  int tail_prev_current = 0;
  uint32_t tail_prev_step = 0;
  unsigned long long tail_serial = 0;
  size_t kld_probe = 8192;          // candidates the KLD tail examines first (tail.cu)
  uint64_t tail_fallbacks = 0;
  uint64_t tail_phase_ns[16] = {};  // per-phase device time, summed over tail_phase_calls calls
  uint64_t tail_phase_calls = 0;
  TailResult *h_tail = nullptr;  // pinned, written by the kernel
  qmcl::DevBuf<TailDev> tail_dev;
  qmcl::DevBuf<uint8_t> tail_scratch;
  qmcl::DevBuf<int32_t> tail_first;
  qmcl::DevBuf<unsigned long long> tail_cluster_w;
  // the estimate the last tail computed (valid while clusters_valid)
  bool est_cached = false;
  bool est_invalid = false;
  double est[13] = {};
  bool total_pending = false;
  double pending_n_total = 0.0;
  bool clusters_pending = false;
  bool bins_pending = false;   // n_bins in flight; n_bins_bound sizes launches meanwhile
  size_t n_bins_bound = 0;
  cudaEvent_t staging_read = nullptr;  // the last H2D copy out of h_beams
  cudaEvent_t total_copied = nullptr;  // the sensor total has reached HS_TOTAL
  // Bumped by every call that moves particles or replaces the set; the bin bounds
  // taken during the sensor update serve the resampling only at the same version.
  uint64_t pose_version = 0, early_bbox_version = ~0ull;
  int bbox_hint[6] = {0, 0, 0, 0, 0, 0};  // bin bounds of the set just resampled from
  bool bbox_hint_valid = false;

  // space_partitioning.cpp:34-35
  int theta_lo = -17, theta_hi = 17;

  qmcl::DevBuf<float> cloud;        // the scan as received, float2
  qmcl::DevBuf<float> ranges;       // raw LaserScan ranges (update_sensor_scan)
  size_t n_cloud = 0;               // points of the last projected scan
  qmcl::DevBuf<float> beams;        // compacted used beams, float2
  qmcl::DevBuf<double> partials;    // per-block partial sums
  qmcl::DevBuf<FilterScalars> scalars;
  qmcl::PinnedBuf<float> h_beams;
  qmcl::PinnedBuf<double> h_scalars;

  // resampling scratch
  qmcl::DevBuf<double> cdf;          // sequentially rounded inclusive running sum
  qmcl::DevBuf<int64_t> pos;         // positive-weight particles, original indices
  qmcl::DevBuf<double> incl;         // running sum after each positive-weight particle
  qmcl::DevBuf<int64_t> src_index;
  size_t n_src_index = 0;
  qmcl::DevBuf<uint8_t> scratch;
  qmcl::DevBuf<uint8_t> prefix_scratch;  // chunk / tile summaries of prefix.cu
  qmcl::DevBuf<double> prefix_carry;     // [0] = running sum after the last element
  int prefix_mode = 0;                   // 0 parallel exact, 1 one-warp sequential loop

  // bins / clusters: dense table over the bin bounding box + sorted bin list
  BinGrid grid{0, 0, 0, 0, 0, 0};
  qmcl::DevBuf<int> bbox;
  qmcl::DevBuf<int32_t> table;        // cell -> bin id + 1 (0 = empty)
  qmcl::DevBuf<int32_t> bin_cell;     // bin -> dense cell index (ascending)
  qmcl::DevBuf<int32_t> bin_count;    // particles per bin
  qmcl::DevBuf<uint8_t> occ8;         // sharded filters: bin occupancy, one byte per cell (the exchanged form)
  bool bins_linked = false;           // the labelling forest starts with the theta runs linked (build_bins)
  bool bin_counts_valid = true;       // false: bin_count is made on demand (qmcl_filter_copy_bins)
  qmcl::DevBuf<int32_t> bin_label;    // union-find parent, then component root
  qmcl::DevBuf<int32_t> bin_cluster;  // root bin -> dense cluster id
  qmcl::DevBuf<uint32_t> scan_tmp;     // status words of the single-pass scans (+ results)
  qmcl::DevBuf<uint32_t> scan_legacy;  // block sums of the three-kernel scans (sharded paths)
  qmcl::DevBuf<double> cluster_acc;
  qmcl::DevBuf<int32_t> particle_cid;  // cluster id per particle (K14)
  size_t n_bins = 0, n_clusters = 0;
  bool bins_built = false;
  bool clusters_valid = false;

  StageTimer timer;

  ParticleSet &cur() { return sets[current]; }
  const ParticleSet &cur() const { return sets[current]; }
};

namespace qmcl {
qmcl_status use_device(const qmcl_map *m);

// RAII: times the enclosed launches on the filter's stream when the timing
// level is at least `min_level`.
struct StageScope {
  qmcl_filter *f;
  int stage;
  cudaEvent_t a = nullptr;
  StageScope(qmcl_filter *f_, int stage_, int min_level = 2) : f(f_), stage(stage_) {
    if (f->timer.level < min_level) return;
    a = f->timer.take();
    cudaEventRecord(a, f->map->stream);
  }
  ~StageScope() {
    if (!a) return;
    cudaEvent_t b = f->timer.take();
    cudaEventRecord(b, f->map->stream);
    f->timer.pending.push_back({stage, a, b});
  }
};
}
