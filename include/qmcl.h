/*
 * qmcl.h — C ABI of the B200-native Monte Carlo localisation core.
 *
 * This is the drop-in boundary for QuickMCL's particle-filter hot path: the
 * entry points are what an FFI for quickmcl::Map (include/quickmcl/map.h:31-62)
 * and quickmcl::IParticleFilter (include/quickmcl/i_particle_filter.h:30-85)
 * would bind.  Plain pointers and sizes only; every function returns a
 * qmcl_status (0 = ok).  All host pointers are ordinary (pageable or pinned)
 * host memory unless a parameter is named dev_*.
 *
 * There is no CPU fallback behind this ABI: every entry point that computes
 * runs hand-written sm_100a kernels, and fails with QMCL_ERR_NO_DEVICE when no
 * CUDA device is present.
 *
 * Reference file:line citations are relative to the QuickMCL 0.5.0 tree.
 */
#ifndef QMCL_H
#define QMCL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum qmcl_status {
  QMCL_OK = 0,
  QMCL_ERR_INVALID_ARG = 1,
  QMCL_ERR_NO_DEVICE = 2,
  QMCL_ERR_CUDA = 3,
  QMCL_ERR_NO_MAP = 4,
  /* get_estimate met a particle whose bin has no cluster: the reference
   * abort()s here (particle_filter.cpp:375-379). */
  QMCL_ERR_INVALID_CLUSTER = 5,
  QMCL_ERR_CAPACITY = 6,
  QMCL_ERR_COMM = 7,        /* a qmcl_comm collective failed */
  QMCL_ERR_UNSUPPORTED = 8  /* the call is not available on a sharded filter */
} qmcl_status;

/* quickmcl::WeightedParticle (particle.h:32-47): Pose2D<float> + double. */
typedef struct qmcl_particle {
  float x, y, theta;
  float pad_;
  double weight;
} qmcl_particle;

/* Parameters::LikelihoodMap (parameters.h:90-113). */
typedef struct qmcl_likelihood_params {
  float z_hit;
  float z_rand;
  int32_t num_beams; /* 0 = use every point */
  float max_laser_distance;
  float max_obstacle_distance;
  float sigma_hit;
} qmcl_likelihood_params;

/* ResampleType (parameters.h:28-35). */
enum { QMCL_RESAMPLE_LOW_VARIANCE = 0, QMCL_RESAMPLE_ADAPTIVE = 1, QMCL_RESAMPLE_KLD = 2 };

/* Parameters::ParticleFilter (parameters.h:55-85). */
typedef struct qmcl_pf_params {
  int32_t resample_type;
  int32_t particle_count_min;
  int32_t particle_count_max;
  int32_t resample_count;
  double alpha_fast;
  double alpha_slow;
  double kld_epsilon;
  double kld_z;
  float res_xy;    /* space_partitioning_resolution_xy */
  float res_theta; /* space_partitioning_resolution_theta */
} qmcl_pf_params;

typedef struct qmcl_map qmcl_map;
typedef struct qmcl_filter qmcl_filter;

const char *qmcl_version(void);
const char *qmcl_last_error(void);
/* Number of kernels this library has launched so far in this process. */
uint64_t qmcl_kernel_launch_count(void);

/* ------------------------------------------------------------------ Map ----
 * quickmcl::Map / MapLikelihood (map.h:31-62, map_likelihood.h:35-93). */

/* create_likelihood_map() (map_factory.h:29). device < 0 = current device. */
qmcl_status qmcl_map_create(int device, qmcl_map **out);
qmcl_status qmcl_map_destroy(qmcl_map *map);
/* All work of this map (and of filters created on it) is enqueued on
 * `cuda_stream` (a cudaStream_t; NULL = the legacy default stream). */
qmcl_status qmcl_map_set_stream(qmcl_map *map, void *cuda_stream);
/* Sensor-model value path: 0 = automatic (for a map built by qmcl_map_load:
 * path 3 when it applies, else path 4; <= 5e-7 relative),
 * 1 = gather the float32 field and do the reference's double arithmetic per
 * evaluation (exact), 2 = uint16 d^2 code per cell with a double pw^3 table
 * (exact), 3 = one-byte palette index per cell with the float pw^3 values
 * (exists when those take at most 254 distinct values over the squared
 * distances that can occur, e.g. sigma_hit <= 0.2 m at 5 cm cells; otherwise
 * path 4 runs), 4 = uint16 d^2 code per cell with a float pw^3 table.
 * Paths 3 and 4 produce identical bits. */
qmcl_status qmcl_map_set_sensor_path(qmcl_map *map, int path);
/* Cell-index division (m - o) / res of map_base.cpp:66-77.  At load the
 * library checks two multiply-based forms against IEEE division around every
 * cell boundary of the map and uses the cheapest one whose truncated quotient
 * never differs: 2 = a*r_hi + a*r_lo (two ops), 1 = Markstein's three-op
 * correctly rounded quotient, 0 = IEEE division.  `mode` caps that choice
 * (-1 = no cap; a form the check rejected is never used); get_ returns the
 * verified form of the loaded map. */
qmcl_status qmcl_map_set_division_mode(qmcl_map *map, int mode);
qmcl_status qmcl_map_get_division_mode(const qmcl_map *map, int *verified);

/* Map::set_parameters (map_likelihood.cpp:40-49). z_hit/z_rand/num_beams/
 * max_laser_distance act on the next sensor update; sigma_hit and
 * max_obstacle_distance on the next qmcl_map_load (map.h:36-39). */
qmcl_status qmcl_map_set_params(qmcl_map *map, const qmcl_likelihood_params *p);

/* Map::new_map (map_base.cpp:30-56, map_likelihood.cpp:51-61):
 * occupancy (row-major [y][x], nav_msgs/OccupancyGrid values) -> tri-state
 * map, free-cell list, distance field, Gaussian table.  The distance field is
 * the exact clipped Euclidean distance transform (north_star); the reference's
 * brushfire (distance_calculator.cpp:103-182) is an approximation of it. */
qmcl_status qmcl_map_load(qmcl_map *map, const int8_t *occ, uint32_t width,
                          uint32_t height, float resolution, double origin_x,
                          double origin_y);
/* Parity hook: install a likelihood field and state map computed elsewhere
 * (e.g. the reference's own brushfire field) instead of building one. */
qmcl_status qmcl_map_load_field(qmcl_map *map, const float *field,
                                const int8_t *state, uint32_t width,
                                uint32_t height, float resolution,
                                double origin_x, double origin_y);
qmcl_status qmcl_map_size(const qmcl_map *map, uint32_t *width, uint32_t *height);
qmcl_status qmcl_map_copy_field(const qmcl_map *map, float *out);
qmcl_status qmcl_map_copy_distance(const qmcl_map *map, float *out);
qmcl_status qmcl_map_copy_state(const qmcl_map *map, int8_t *out);
qmcl_status qmcl_map_num_free_cells(const qmcl_map *map, uint64_t *out);
qmcl_status qmcl_map_copy_free_cells(const qmcl_map *map, int32_t *xy_out);
/* Map::get_map_state_at (map_base.cpp:66-77). */
qmcl_status qmcl_map_state_at(const qmcl_map *map, float x, float y, int8_t *out);
/* Map::get_likelihood_as_gridmap (map_likelihood.cpp:132-146). */
qmcl_status qmcl_map_likelihood_grid(const qmcl_map *map, int8_t *out);
/* Map::update_importance on a host particle vector (map_likelihood.cpp:63-130):
 * H2D, kernel, D2H.  total_out receives the pre-penalty total weight. */
qmcl_status qmcl_map_update_importance(qmcl_map *map, const float *cloud_xy,
                                       size_t n_points, qmcl_particle *particles,
                                       size_t n_particles, double *total_out);
/* Map::generate_random_pose (map_base.cpp:58-64) for draw `index` of the
 * shared RNG contract (Philox key `seed`, counter (index, step, GLOBAL)). */
qmcl_status qmcl_map_generate_random_pose(const qmcl_map *map, uint64_t seed,
                                          uint32_t step, uint64_t index,
                                          float pose_out[3]);

/* --------------------------------------------------------------- Filter ----
 * quickmcl::IParticleFilter (i_particle_filter.h:30-85). */

/* create_particle_filter(map) (particle_filter_factory.h:30-31). */
qmcl_status qmcl_filter_create(qmcl_map *map, uint64_t seed, qmcl_filter **out);
qmcl_status qmcl_filter_destroy(qmcl_filter *f);
/* ---- particle sharding (north_star / SURVEY 8e) ----------------------------
 * One filter handle per GPU owns a contiguous block of the global particle
 * index range; the blocks are ordered by rank, so the global CDF is the
 * concatenation of the shards' CDFs.  The map, state map and scan are
 * replicated.  The handful of exchanges per cycle (weight total, running-sum
 * carry, bin occupancy, cluster weights, estimate moments, particle
 * rebalancing) go through these host-supplied collectives; the host decides
 * what carries them (torch.distributed over NCCL in quickmcl_b200/sharded.py;
 * an MPI or raw NCCL host fills the same table).
 *
 * Every buffer is a DEVICE pointer.  A call must be ordered after the work
 * already enqueued on `stream` (the filter's cudaStream_t) and the filter's
 * later work on that stream must see its result.  Return 0 on success. */
enum { QMCL_DT_I32 = 0, QMCL_DT_I64 = 1, QMCL_DT_F64 = 2 };
enum { QMCL_OP_SUM = 0, QMCL_OP_MIN = 1, QMCL_OP_MAX = 2 };
typedef struct qmcl_comm {
  int rank, size;
  void *ctx;
  /* in place: buf holds size * bytes_per_rank bytes, this rank's part at
   * rank * bytes_per_rank */
  int (*allgather)(void *ctx, void *dev_buf, size_t bytes_per_rank, void *stream);
  /* in place over `count` elements of dtype (QMCL_DT_*) with op (QMCL_OP_*) */
  int (*allreduce)(void *ctx, void *dev_buf, size_t count, int dtype, int op, void *stream);
  /* byte counts and offsets per peer (arrays of `size`) */
  int (*alltoallv)(void *ctx, const void *dev_send, const uint64_t *send_bytes,
                   const uint64_t *send_offsets, void *dev_recv, const uint64_t *recv_bytes,
                   const uint64_t *recv_offsets, void *stream);
} qmcl_comm;
/* Built-in table on NCCL (resolved with dlopen at first use).  Rank 0 obtains
 * a 128-byte ncclUniqueId, the host hands it to every rank, each rank creates
 * its communicator on `device` (< 0: current device). */
qmcl_status qmcl_nccl_unique_id(uint8_t out[128]);
qmcl_status qmcl_nccl_comm_create(const uint8_t id[128], int rank, int size, int device,
                                  qmcl_comm *out);
qmcl_status qmcl_nccl_comm_destroy(qmcl_comm *comm);
/* comm == NULL (or size 1) returns the filter to the single-GPU path. */
qmcl_status qmcl_filter_set_comm(qmcl_filter *f, const qmcl_comm *comm);
/* After a sharded resampling, move particles so that every rank again owns an
 * equal contiguous block (default 1 = on). */
qmcl_status qmcl_filter_set_rebalance(qmcl_filter *f, int enabled);
/* ... but only when some rank holds more than (1 + t) or less than (1 - t)
 * times its equal share ("an optional all-to-all rebalances when the
 * imbalance exceeds a threshold").  Default 0: after every resampling. */
qmcl_status qmcl_filter_set_rebalance_threshold(qmcl_filter *f, double relative_imbalance);
/* This shard's block: first global index and local count. */
qmcl_status qmcl_filter_shard_range(const qmcl_filter *f, uint64_t *first, uint64_t *count,
                                    uint64_t *global_count);
/* Exchange counters of a sharded filter since it was created: out[0] = collectives
 * enqueued, out[1] = carry-chain rounds of the running sum that needed an exchange of
 * their own (a shard without a closed-form summary), out[2] = rebalancing all-to-alls
 * skipped under the threshold. */
qmcl_status qmcl_filter_shard_stats(const qmcl_filter *f, uint64_t out[3]);
/* Pure host helpers of the sharded path (no device needed; used by the CPU
 * tests of the multi-rank logic).
 * qmcl_shard_slice: block [first, first+count) of rank `rank` of `size` over n.
 * qmcl_lv_owner_range: output slots [m_lo, m_hi) of the low-variance resampler
 * (particle_filter.cpp:225-238) whose position U_m = r + m * minv selects a
 * particle of the shard whose running sum spans (c_lo, c_hi]. */
void qmcl_shard_slice(uint64_t n, int rank, int size, uint64_t *first, uint64_t *count);
void qmcl_lv_owner_range(uint64_t M, double r, double minv, double c_lo, double c_hi,
                         int is_first, int is_last, uint64_t *m_lo, uint64_t *m_hi);

/* IParticleFilter::set_parameters (particle_filter.cpp:177-184). */
qmcl_status qmcl_filter_set_params(qmcl_filter *f, const qmcl_pf_params *p);
/* IParticleFilter::initialise (particle_filter.cpp:41-69). cov row-major. */
qmcl_status qmcl_filter_initialise(qmcl_filter *f, const float pose[3],
                                   const float cov[9]);
/* Same, with the 3x3 factor V*sqrt(Lambda) (random.h:60-62) supplied. */
qmcl_status qmcl_filter_initialise_factor(qmcl_filter *f, const float pose[3],
                                          const float factor[9]);
/* IParticleFilter::global_localization (particle_filter.cpp:71-83). */
qmcl_status qmcl_filter_global_localization(qmcl_filter *f);
/* IParticleFilter::handle_odometry (particle_filter.cpp:85-93,
 * odometry.cpp:44-102). */
qmcl_status qmcl_filter_handle_odometry(qmcl_filter *f, const double odom_new[3],
                                        const double odom_old[3],
                                        const float alpha[4]);
/* IParticleFilter::update_importance_from_observations
 * (particle_filter.cpp:95-136). cloud_xy = n_points x {float x, y}.  The cloud
 * is copied before the call returns; in the default sync mode
 * (qmcl_filter_set_sync_mode) the kernels may still be running when it does. */
qmcl_status qmcl_filter_update_sensor(qmcl_filter *f, const float *cloud_xy,
                                      size_t n_points);
/* Device-resident variant for benchmarking the kernel alone: the cloud is
 * already in HBM (dev_cloud_xy), nothing is copied to the host.  dev_cloud_xy
 * is read by the first kernel enqueued: keep it valid until the next call that
 * returns a result (get_estimate, a getter, qmcl_filter_synchronize). */
qmcl_status qmcl_filter_update_sensor_dev(qmcl_filter *f, const float *dev_cloud_xy,
                                          size_t n_points);
/* The same update fed with a raw planar LaserScan: the projection
 * of laser_handler.cpp:108-121 (laser_geometry: a point per range with
 * range_min <= r < range_max at angle_min + i * angle_increment) and the rigid
 * laser->base mount transform (x, y, yaw) run on the device, so only the 4 B
 * ranges cross PCIe.  n_points_out (optional) receives the number of valid points. */
qmcl_status qmcl_filter_update_sensor_scan(qmcl_filter *f, const float *ranges, size_t n_ranges,
                                           float angle_min, float angle_increment,
                                           float range_min, float range_max,
                                           const double mount[3], size_t *n_points_out);
/* Test hook: the projected cloud of the last qmcl_filter_update_sensor_scan. */
qmcl_status qmcl_filter_copy_cloud(const qmcl_filter *f, float *xy_out, size_t cap, size_t *n_out);
/* IParticleFilter::resample_and_cluster (particle_filter.cpp:138-161). */
qmcl_status qmcl_filter_resample_and_cluster(qmcl_filter *f);
/* IParticleFilter::cluster (particle_filter.cpp:163-175). */
qmcl_status qmcl_filter_cluster(qmcl_filter *f);
/* IParticleFilter::get_estimated_pose (particle_filter.cpp:186-212).
 * *valid = 1 and pose/cov written when a cluster has weight, else 0. */
qmcl_status qmcl_filter_get_estimate(qmcl_filter *f, double pose[3],
                                     double cov[9], int *valid);
/* effective_particles (particle.cpp:32-39, test_particle.cpp:12-27):
 * 1 / sum of squared weights of the current set. */
qmcl_status qmcl_filter_effective_particles(const qmcl_filter *f, double *out);
/* IParticleFilter::num_particles / get_particles (AoS copy). */
qmcl_status qmcl_filter_num_particles(const qmcl_filter *f, size_t *out);
qmcl_status qmcl_filter_copy_particles(const qmcl_filter *f, qmcl_particle *out,
                                       size_t cap);

/* ---- multi-robot batch (BASELINE configs[4]) ------------------------------
 * One scan cycle (laser_handler.cpp:243-276: odometry, sensor, resample or
 * cluster, estimate) for n_filters independent filters, dealt round-robin to
 * `lanes` native host threads.  Filters handled by different lanes must sit on
 * different map handles (streams).  odom_*: n x 3 doubles; clouds_xy[i] holds
 * n_points[i] x {x, y}; poses_out n x 3, covs_out n x 9, valid_out n. */
/* The particle cloud as the node publishes it (publishing.cpp:45-67,114-154):
 * out5[5*i..] = x, y, sin(theta/2), cos(theta/2), float(weight / max_weight),
 * the maximum taken over the whole set.  Copies min(n, capacity) particles;
 * n_out / max_weight may be NULL. */
qmcl_status qmcl_filter_export_markers(const qmcl_filter *filter, float *out5, size_t capacity,
                                       size_t *n_out, double *max_weight);

qmcl_status qmcl_batch_update(qmcl_filter *const *filters, size_t n_filters, int lanes,
                              const double *odom_new, const double *odom_old,
                              const float alpha[4], const float *const *clouds_xy,
                              const size_t *n_points, int resample, double *poses_out,
                              double *covs_out, int *valid_out);

/* The same cycle for n filters that live on ONE map handle, in three launches
 * for the whole batch (motion + beam selection, sensor model, and the fused
 * tail with one block per filter) and one host wait.  The batch owns its
 * filters (seeds seed, seed + 1, ...); qmcl_batch_filter lends filter i for the
 * per-filter calls (initialise, set_particles, getters).  clouds_xy holds the
 * scans back to back, robot i's points at [cloud_offsets[i], cloud_offsets[i+1])
 * (offsets in points, n + 1 entries).  Falls back to a per-filter loop on the
 * same stream when a filter cannot take the fused path (sets above 65536
 * particles, eager sync mode, stage timing). */
typedef struct qmcl_batch qmcl_batch;
qmcl_status qmcl_batch_create(qmcl_map *map, size_t n_filters, uint64_t seed, qmcl_batch **out);
qmcl_status qmcl_batch_destroy(qmcl_batch *batch);
qmcl_status qmcl_batch_size(const qmcl_batch *batch, size_t *out);
qmcl_status qmcl_batch_filter(qmcl_batch *batch, size_t i, qmcl_filter **out);
qmcl_status qmcl_batch_set_params(qmcl_batch *batch, const qmcl_pf_params *p);
qmcl_status qmcl_batch_cycle(qmcl_batch *batch, const double *odom_new, const double *odom_old,
                             const float alpha[4], const float *clouds_xy,
                             const size_t *cloud_offsets, int resample, double *poses_out,
                             double *covs_out, int *valid_out);
/* Cycles that ran fused / as a per-filter loop so far. */
qmcl_status qmcl_batch_stats(const qmcl_batch *batch, uint64_t *fused_cycles, uint64_t *looped_cycles);

/* ---- test hooks (not part of the reference interface) ---- */
qmcl_status qmcl_filter_set_particles(qmcl_filter *f, const qmcl_particle *p, size_t n);
/* Sharded variant: p[0..n) is this rank's block, starting at global index
 * `global_first` of `global_count` particles. */
qmcl_status qmcl_filter_set_particles_shard(qmcl_filter *f, const qmcl_particle *p, size_t n,
                                            uint64_t global_first, uint64_t global_count);
qmcl_status qmcl_filter_set_rng(qmcl_filter *f, uint64_t seed, uint32_t step);
/* Pre-penalty total weight of the last sensor update. */
qmcl_status qmcl_filter_last_total_weight(const qmcl_filter *f, double *out);
/* Source index per output slot of the last resampling (-1 = injected). */
qmcl_status qmcl_filter_copy_resample_indices(const qmcl_filter *f, int64_t *out,
                                              size_t cap, size_t *n_out);
/* Occupied bins as (kx, ky, ktheta, count), lexicographically sorted, with a
 * canonical cluster label per bin (index of the first bin of its component). */
qmcl_status qmcl_filter_num_bins(const qmcl_filter *f, size_t *out);
qmcl_status qmcl_filter_copy_bins(const qmcl_filter *f, int32_t *keys4,
                                  int32_t *labels, size_t cap, size_t *n_out);
qmcl_status qmcl_filter_num_clusters(const qmcl_filter *f, size_t *out);
qmcl_status qmcl_filter_get_lowpass(const qmcl_filter *f, double out[2]);
qmcl_status qmcl_filter_set_lowpass(qmcl_filter *f, double w_slow, double w_fast);
/* Running-sum kernel selection: 0 = parallel, bit-identical to the sequential
 * loop (default); 1 = the literal one-warp sequential loop (cross-check);
 * 2 = the parallel path with the chunks in which the sum crosses a power of two
 * resolved lane-wise instead of by 256 literal additions (same bits; measured
 * slower, kept for A/B). */
qmcl_status qmcl_filter_set_prefix_mode(qmcl_filter *f, int mode);
/* out[i] = sequentially rounded running sum of carry_in, w[0..i] (host arrays). */
qmcl_status qmcl_filter_debug_prefix(qmcl_filter *f, const double *w, size_t n,
                                     double carry_in, double *out, double *carry_out);
/* Wait for everything enqueued by this filter. */
qmcl_status qmcl_filter_synchronize(qmcl_filter *f);
/* Host/device synchronisation policy of a single-GPU filter.  1 (default):
 * values the host needs only later (the sensor total feeding w_slow / w_fast,
 * the cluster count) stay in flight and are folded in at the next call that
 * synchronises anyway, so update_sensor returns before its kernels finish and
 * a KLD update waits on the device three times instead of seven.  0: one
 * synchronisation per value.  Results are identical in both modes; the
 * environment variable QMCL_EAGER_SYNC=1 selects 0 for new filters. */
qmcl_status qmcl_filter_set_sync_mode(qmcl_filter *f, int lazy);

/* ---- measurement hooks (replace the reference's CodeTimer scopes,
 *      timer.cpp:37-55; laser_handler.cpp:244,253,262) ---- */
enum {
  QMCL_STAGE_MOTION = 0,      /* K6 motion model */
  QMCL_STAGE_SENSOR = 1,      /* K7 sensor kernel alone (the roofline kernel) */
  QMCL_STAGE_NORMALISE = 2,   /* K8 */
  QMCL_STAGE_CDF = 3,         /* K9 sequentially rounded running sum (+ records) */
  QMCL_STAGE_SELECT = 4,      /* K10 index search / draws */
  QMCL_STAGE_GATHER = 5,      /* K12 */
  QMCL_STAGE_KLD = 6,         /* K11 first occurrence, prefix count, stop rule */
  QMCL_STAGE_BINS = 7,        /* bin table + occupied-bin list */
  QMCL_STAGE_LABEL = 8,       /* K13 connected components */
  QMCL_STAGE_POST_NORMALISE = 9,
  QMCL_STAGE_MOMENTS = 10,    /* K14 */
  QMCL_STAGE_TAIL = 11,       /* stages 2-10 as one fused launch (default on one GPU) */
  QMCL_N_STAGES = 12
};
const char *qmcl_stage_name(int stage);
/* level 0 = off, 1 = CUDA events around the sensor kernel and the fused tail,
 * 2 = around every stage (the stages then run as separate launches).  Events
 * are recorded on the filter's stream. */
qmcl_status qmcl_filter_set_timing(qmcl_filter *f, int level);
/* Synchronises, then returns accumulated device milliseconds and the number
 * of timed occurrences per stage since the last reset. */
qmcl_status qmcl_filter_get_timing(qmcl_filter *f, double ms_out[QMCL_N_STAGES],
                                   uint64_t count_out[QMCL_N_STAGES], int reset);
/* The fused tail (QMCL_STAGE_TAIL): device nanoseconds spent in each of its
 * phases (qmcl_tail_phase_name(k), k = 1..14), summed over the *calls fused
 * calls this filter completed since the last reset, and the number of fused
 * calls that had to be re-run on the multi-launch path so far. */
qmcl_status qmcl_filter_tail_profile(qmcl_filter *f, uint64_t out[16], uint64_t *calls,
                                     uint64_t *fallbacks, int reset);
const char *qmcl_tail_phase_name(int k);
/* Bytes this library has copied host->device (out[0]) and device->host
 * (out[1]) so far in this process. */
qmcl_status qmcl_transfer_bytes(uint64_t out[2]);

#ifdef __cplusplus
}
#endif
#endif /* QMCL_H */
