/*
 * qmcl_oracle.h — CPU restatement of QuickMCL's particle-filter hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library, and only as the checker or as the
 * timed CPU baseline.  The CUDA library (quickmcl_b200/csrc) never links it.
 *
 * The reference's own build (catkin / cmake, ROS, Eigen, Boost, PCL, gtest) cannot
 * run in this image, so this is a dependency-free restatement of QuickMCL 0.5.0
 * (/root/reference).  Every function cites the reference file:line it follows.
 * Parity pinning: (1) the helpers are checked against all golden vectors of the
 * reference's own gtest suite (tests/test_oracle_golden.py); (2) the reference's
 * translation units of the path - distance_calculator, map_base, map_likelihood,
 * laser, odometry, particle, particle_filter, pose_2d, pose_restore,
 * space_partitioning, statistics, timer and the node's laser_handler
 * .cpp and the header-only parts - are compiled unmodified from where they lie
 * into oracle/_ref/libqmcl_ref.so (oracle/Makefile target `ref`,
 * oracle/ref_shim.cpp, stand-in Eigen / ROS / PCL headers and a scripted
 * <random> in oracle/ref_stubs) and this oracle is held against them bit for
 * bit on random inputs: field and state maps, grid truncation, weights, running
 * sum, KLD bound, bins and clusters (tests/test_oracle_vs_ref.py), the sensor
 * model, the motion model, the statistics, the three resamplers, global
 * localisation and whole scan cycles (tests/test_oracle_vs_ref_filter.py).
 * "Parity unpinned" remains for: ParticleFilter::initialise beyond T T^T =
 * covariance (Eigen's eigen-solver and libstdc++'s normal_distribution decide
 * the individual draws), the laser_geometry projection (third-party, absent),
 * and the three pieces of Eigen's own arithmetic that the stand-in headers
 * restate instead of running (Isometry2f product, the statistics' outer
 * product, the packet exp() of the table build: DESIGN.md section 2).  The
 * reference draws from private, unseeded std::mt19937 engines, so both this
 * oracle and the CUDA path consume one shared counter-based Philox4x32-10
 * stream instead (contract in DESIGN.md section 6); the tests replay that
 * stream into the reference's engines.
 */
#ifndef QMCL_ORACLE_H
#define QMCL_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same memory layout as quickmcl::WeightedParticle (particle.h:32-47):
 * Pose2D<float> {x,y,theta} + 4 pad bytes + double weight = 24 bytes. */
typedef struct orc_particle {
  float x, y, theta;
  float pad_;
  double weight;
} orc_particle;

/* Parameters::LikelihoodMap (parameters.h:90-113). */
typedef struct orc_likelihood_params {
  float z_hit;
  float z_rand;
  int32_t num_beams;
  float max_laser_distance;
  float max_obstacle_distance;
  float sigma_hit;
} orc_likelihood_params;

/* Parameters::ParticleFilter (parameters.h:55-85). resample_type:
 * 0 low_variance, 1 adaptive, 2 kld (parameters.h:28-35). */
typedef struct orc_pf_params {
  int32_t resample_type;
  int32_t particle_count_min;
  int32_t particle_count_max;
  int32_t resample_count;
  double alpha_fast;
  double alpha_slow;
  double kld_epsilon;
  double kld_z;
  float res_xy;
  float res_theta;
} orc_pf_params;

enum { ORC_FIELD_BRUSHFIRE = 0, ORC_FIELD_EXACT_EDT = 1 };
/* How cos/sin of the particle heading are obtained in the sensor model:
 * LIBM = sinf/cosf of the host libm (what the reference does);
 * CR   = correctly rounded: (float)sin((double)theta) (what the kernels do). */
enum { ORC_TRIG_LIBM = 0, ORC_TRIG_CR = 1 };

typedef struct orc_map orc_map;
typedef struct orc_filter orc_filter;

/* ---- helpers pinned by the reference's gtest golden vectors ---- */
void orc_distance_cache(float resolution, float max_dist, int *size_out,
                        float *out, size_t cap);
void orc_world_to_map(const float *pos, const float *offset, const float *res,
                      int dims, int *out);
void orc_map_to_world(const int *cell, const float *offset, const float *res,
                      int dims, float *out);
int orc_is_in_map(const int *cell, const int *size, int dims);
double orc_normalise_angle_f64(double a);
float orc_normalise_angle_f32(float a);
double orc_angle_delta_f64(double a, double b);
float orc_angle_delta_f32(float a, float b);
void orc_transform_point_f32(const float pose[3], const float pt[2],
                             float out[2], int trig_mode);
void orc_to_decomposed(const float pose[3], double out[4]);
double orc_effective_particles(const orc_particle *p, size_t n);
void orc_normalise_weights(orc_particle *p, size_t n);
void orc_normalise_weights_total(orc_particle *p, size_t n, double total);
void orc_equalise_weights(orc_particle *p, size_t n);
/* ParticlesRunningSum: build over p[0..n), answer nq lookups. */
double orc_running_sum_lookup(const orc_particle *p, size_t n,
                              const double *queries, size_t nq,
                              uint64_t *idx_out);
/* ParticleCloudStatistics: accumulate then finalise.
 * out = {total_weight, sample_count, mean[0..3], cov[0..8] row-major}. */
void orc_statistics(const orc_particle *p, size_t n, int finalise,
                    double out[15]);
void orc_statistics_merge(const double a[15], const double b[15],
                          int finalise, double out[15]);
double orc_lowpass_run(double alpha, const double *xs, size_t n, int as_float);
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2],
                       uint32_t out[4]);
/* The shared RNG contract, exposed for tests that replay the oracle's draws
 * into the reference's own code (oracle/ref_shim.cpp): the four Philox words of
 * draws first_index .. first_index + count - 1 for one (seed, step, purpose),
 * and the Box-Muller pair the motion model makes of two words. */
void orc_rng_words(uint64_t seed, uint32_t step, uint32_t purpose,
                   uint64_t first_index, size_t count, uint32_t *out4);
void orc_normal_pairs(const uint32_t *words2, size_t count, double *out2);
uint64_t orc_kld_limit(uint64_t k, double eps, double z, int32_t nmin,
                       int32_t nmax);

/* ---- map / likelihood field (map_base.cpp, map_likelihood.cpp,
 *      distance_calculator.cpp) ---- */
orc_map *orc_map_create(void);
void orc_map_destroy(orc_map *m);
void orc_map_set_params(orc_map *m, const orc_likelihood_params *p);
void orc_map_load(orc_map *m, const int8_t *occ, uint32_t w, uint32_t h,
                  float res, double ox, double oy, int field_mode);
void orc_map_load_field(orc_map *m, const float *field, const int8_t *state,
                        uint32_t w, uint32_t h, float res, double ox,
                        double oy);
void orc_map_copy_distance(const orc_map *m, float *out);
void orc_map_copy_field(const orc_map *m, float *out);
void orc_map_copy_state(const orc_map *m, int8_t *out);
size_t orc_map_num_free(const orc_map *m);
void orc_map_copy_free_cells(const orc_map *m, int32_t *xy_out);
int8_t orc_map_state_at(const orc_map *m, float x, float y);
void orc_map_likelihood_grid(const orc_map *m, int8_t *out);
double orc_map_update_importance(const orc_map *m, const float *xy,
                                 size_t npoints, orc_particle *p, size_t n,
                                 int trig_mode);
/* Brute-force squared distance (cells^2) to the nearest obstacle within
 * radius cells; -1 if none.  Used to pin the exact EDT. */
void orc_bruteforce_d2(const int8_t *occ, uint32_t w, uint32_t h, int radius,
                       int32_t *out);

/* ---- filter (particle_filter.cpp, odometry.cpp, space_partitioning.cpp,
 *      statistics.cpp) ---- */
orc_filter *orc_filter_create(orc_map *m, uint64_t seed);
void orc_filter_destroy(orc_filter *f);
void orc_filter_set_params(orc_filter *f, const orc_pf_params *p);
void orc_filter_set_rng(orc_filter *f, uint64_t seed, uint32_t step);
void orc_filter_set_trig_mode(orc_filter *f, int trig_mode);
void orc_filter_initialise(orc_filter *f, const float pose[3],
                           const float cov[9]);
/* The 3x3 factor V*sqrt(Lambda) used by initialise (random.h:60-62). */
void orc_cov_factor(const float cov[9], float out[9]);
void orc_filter_initialise_factor(orc_filter *f, const float pose[3],
                                  const float factor[9]);
void orc_filter_global_localization(orc_filter *f);
void orc_filter_handle_odometry(orc_filter *f, const double odom_new[3],
                                const double odom_old[3],
                                const float alpha[4]);
void orc_filter_update_sensor(orc_filter *f, const float *xy, size_t npoints);
void orc_filter_resample_and_cluster(orc_filter *f);
void orc_filter_cluster(orc_filter *f);
int orc_filter_get_estimate(const orc_filter *f, double pose[3],
                            double cov[9]);
size_t orc_filter_num_particles(const orc_filter *f);
void orc_filter_copy_particles(const orc_filter *f, orc_particle *out,
                               size_t cap);
void orc_filter_set_particles(orc_filter *f, const orc_particle *p, size_t n);
size_t orc_filter_copy_resample_indices(const orc_filter *f, int64_t *out,
                                        size_t cap);
size_t orc_filter_num_bins(const orc_filter *f);
/* bins as (kx,ky,kt,count) sorted lexicographically; cluster label per bin
 * canonicalised as the smallest sorted-bin index in its component. */
size_t orc_filter_copy_bins(const orc_filter *f, int32_t *keys4,
                            int32_t *labels, size_t cap);
size_t orc_filter_num_clusters(const orc_filter *f);
void orc_filter_get_lowpass(const orc_filter *f, double out[2]);
void orc_filter_set_lowpass(orc_filter *f, double w_slow, double w_fast);
double orc_filter_last_total_weight(const orc_filter *f);

/* LaserScan -> base-frame point cloud: what
 * laser_handler.cpp:108-121 obtains from laser_geometry's
 * transformLaserScanToPointCloud(localised_frame, scan, cloud, tf, -1.0, 0)
 * followed by pcl::fromROSMsg (laser.cpp:25-34), for a laser mounted rigidly on
 * the base.  laser_geometry (1.6, ROS Melodic) is a third-party dependency that
 * is not in the reference tree; this restates its published algorithm: a point
 * per range with range_min <= r < range_max at angle_min + i * increment,
 * cos/sin in double, stored as float, then the rigid mount transform in double
 * (tf2), stored as float.  Parity unpinned (no reference test covers it).
 * out_xy holds up to n points (x, y); returns the number written. */
size_t orc_project_scan(const float *ranges, size_t n, float angle_min,
                        float angle_increment, float range_min,
                        float range_max, const double mount[3],
                        float *out_xy);

/* Stage timings of the last call, seconds (steady_clock). */
double orc_last_seconds(void);

#ifdef __cplusplus
}
#endif
#endif
