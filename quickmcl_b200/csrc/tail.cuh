// tail.cuh — the fused "tail" of an update: everything after the sensor kernel
// (normalise, running sum, resampling, KLD stop rule, bins, connected
// components, post-normalise, cluster weights, moments, estimate) as ONE launch
// whose phases are separated by team barriers instead of kernel boundaries.
//
// Why: a KLD update at 100 k particles is ~25 dependent kernels of 3-10 us
// each.  Measured on the B200 (profiles/r2_probe_sync.txt): a dependent launch
// costs 3.0 us, a grid-wide barrier of a co-resident grid 1.2 us, and a kernel
// has to spin up and drain 148 SMs either way.  The phases below run the very
// same device bodies as the stand-alone kernels (prefix_dev.cuh,
// resample_dev.cuh, cluster_dev.cuh) over the same *virtual* block geometry, so
// every result is bit-identical to the multi-launch path, which stays in the
// library as the cross-check (qmcl_filter_set_sync_mode(f, 0)) and as the path
// for sharded filters and very large sets.
//
// A *team* is the set of blocks working on one filter: the whole cooperative
// grid (single filter) or one block (the multi-robot batch, one filter per
// block).  Launch sizes never depend on values the device computes: counts
// (records, KLD stop index, bins, clusters) are read from device words, and the
// host learns them — together with the estimate — from one block of pinned
// memory the last phase writes.
#pragma once

#include "cluster_dev.cuh"
#include "prefix_dev.cuh"
#include "resample_dev.cuh"

namespace qmcl {

enum : int { TAIL_LOW_VARIANCE = 0, TAIL_ADAPTIVE = 1, TAIL_KLD = 2, TAIL_CLUSTER = 3 };
enum : unsigned long long {
  TAIL_OK = 0,
  TAIL_FALLBACK = 1,             // nothing was touched: run the multi-launch path instead
  TAIL_FALLBACK_NORMALISED = 2,  // the same, but the weights were already divided by the sensor total
  TAIL_NO_WEIGHT = 3             // KLD: "No particles have any weights" (particle_filter.cpp:307-311)
};

// Phase boundaries of the tail, stamped by block 0 (qmcl_filter_tail_profile).
enum : int {
  TS_START = 0, TS_NORMALISE, TS_SUMMARIES, TS_RESOLVE, TS_APPLY, TS_SELECT, TS_KLD, TS_MARK, TS_BINS,
  TS_LABEL, TS_WEIGHT_SUM, TS_CLUSTER_W, TS_BEST, TS_MOMENTS, TS_END, kTailStamps = 16
};

// What the last phase leaves for the host (pinned, mapped memory; 8-byte words).
struct TailResult {
  unsigned long long status;
  unsigned long long n_out;       // particles of the new set
  unsigned long long n_bins;
  unsigned long long n_clusters;
  unsigned long long n_records;   // running-sum records
  unsigned long long invalid;     // a particle's bin has no cluster (particle_filter.cpp:375-379)
  long long grid[6];              // x0, y0, t0, nx, ny, nt of the dense bin grid
  double total_weight;            // pre-penalty sensor total (when the normalisation was pending)
  double w_diff;                  // injection probability used (adaptive / KLD)
  double post_total;              // normalising total of the new set
  double est[13];                 // valid, pose[3], cov[9]
  unsigned long long stamps[kTailStamps];  // %globaltimer (ns) at the phase boundaries
  unsigned long long serial;      // launch serial number, written last
};

// Device-side scalars the phases hand to each other.
struct TailDev {
  unsigned long long barrier;     // grid barrier: arrivals so far (never reset)
  unsigned long long stop;        // KLD stop word {m:32, k:32}
  unsigned long long n_new;       // KLD new-bin flags over all candidates
  unsigned long long n_out;
  unsigned long long n_bins;
  unsigned long long n_clusters;
  unsigned long long n_records;
  unsigned long long wmax_raw;    // largest weight before the post-normalisation (bits)
  unsigned int slot_counter;      // prefix lane summaries
  unsigned int pad;
  EstimateScalars es;
  unsigned long long stamps[kTailStamps];
  double sums[2 * kMom];
  double out[16];
};

struct TailJob {
  int mode;               // TAIL_*
  int normalise_pending;  // divide the input weights by sc->total_weight first (K8)
  // current / input set and output set (SoA)
  float *ix, *iy, *it;
  double *iw;
  float *ox, *oy, *ot;
  double *ow;
  unsigned long long n_in;
  unsigned long long n_min, n_max;  // particle_count_min / max
  unsigned long long kld_probe;     // KLD: candidates examined in the first round (0 = all)
  unsigned long long solo_cells;    // grids up to this many cells are listed and labelled by block 0 alone
  KldParams kp;
  BinParams bp;
  MapGeom geom;
  const uint32_t *free_cells;
  unsigned long long n_free;
  int map_bbox[6];        // bin bounds of the map's extent (injected poses lie inside)
  uint64_t seed;
  uint32_t step;
  // low-pass filters before this update's total is applied (low_pass_filter.h:39-50)
  double w_slow, w_fast, alpha_slow, alpha_fast;
  // scratch of the filter
  const FilterScalars *sc;
  const int *bbox;        // device: bin bounds of the input set (6 ints)
  double *cdf;
  int64_t *pos;
  double *incl;
  int64_t *src;
  pfx::PrefixBuffers pb;
  unsigned long long nc, nt;  // chunks / tiles of the running sum
  double *tile_sum;           // nt
  uint32_t *tile_pos;         // nt
  uint32_t *tile_cnt;         // scan tiles of max(n_max, table_cap, bins_cap)
  int32_t *first;             // KLD first-draw table
  int32_t *table;
  unsigned long long table_cap;
  int32_t *bin_cell, *bin_count, *bin_label, *bin_cluster;
  unsigned long long bins_cap;
  unsigned long long *cluster_w;  // bins_cap
  int32_t *particle_cid;
  double *partials;        // weight-sum partials (4 * 148), then moments partials (4 * 148 * 2 * kMom)
  TailDev *dev;
  TailResult *result;      // pinned host memory, device-accessible
  unsigned long long serial;
};

// Host side (tail.cu): fill / launch.  tail_prepare does the bookkeeping of a
// launch, so the kernel must follow on the filter's stream.
qmcl_status tail_prepare(qmcl_filter *f, int mode, TailJob *J);
qmcl_status launch_tail_batch(qmcl_map *m, const TailJob *dev_jobs, unsigned n_jobs);

constexpr unsigned kTailThreads = 256;
// doubles `partials` must hold
constexpr size_t kTailPartials = 4 * 148 + static_cast<size_t>(4 * 148) * 2 * kMom;

}  // namespace qmcl
