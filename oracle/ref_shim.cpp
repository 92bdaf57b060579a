/*
 * ref_shim.cpp — C entry points over the REFERENCE's own code (oracle/_ref).
 *
 * TEST INFRASTRUCTURE ONLY.  This file is compiled together with the reference's
 * translation units of the particle-filter path, taken from where they lie under
 * /root/reference (never copied): distance_calculator.cpp, map_base.cpp,
 * map_factory.cpp, map_likelihood.cpp, odometry.cpp, particle.cpp, particle_filter.cpp,
 * particle_filter_factory.cpp, pose_2d.cpp,
 * pose_restore.cpp, space_partitioning.cpp, statistics.cpp, timer.cpp, plus the header-only scaled_map.h,
 * low_pass_filter.h, random.h and utils.h.  Eigen, ROS and PCL are absent from this
 * image; the units build against the stand-in headers in oracle/ref_stubs (what each
 * stand-in restates is written at its top: Eigen/Core, Eigen/Geometry,
 * Eigen/Eigenvalues, random).  The result, oracle/_ref/libqmcl_ref.so, lets
 * tests/test_oracle_vs_ref.py hold the restated oracle against what the reference's
 * statements actually compute, on inputs far beyond the reference's own golden
 * vectors: brushfire distance and state maps, the likelihood table, the sensor
 * model, the motion model, the three resamplers, KLD, clustering, the estimate and
 * whole scan cycles.
 *
 * Randomness: the reference's engines are private and unseeded, so ref_stubs/random
 * replaces std::mt19937 by an engine that reads a script (ref_script below).  The
 * tests write into that script the very words the shared Philox contract (DESIGN.md
 * section 6) hands the oracle for each draw, in the order the reference consumes them.
 *
 * This file is compiled with -fno-access-control: it reads and sets private members
 * of the reference's classes (the likelihood table, the low-pass filters, the
 * current particle set) where the reference has no accessor.
 *
 * The same sources and stand-ins also build the reference's own gtest suite
 * (oracle/Makefile, target ref_gtests): 28 of 28 cases pass.
 *
 * laser.cpp and the node's laser_handler.cpp are built too: see ref_shim_node.cpp.
 * Not built: the rest of the node (main, publishing, tf_reader, commands, parameter_manager).
 */
#include "quickmcl/distance_calculator.h"
#include "quickmcl/low_pass_filter.h"
#include "quickmcl/map_factory.h"
#include "quickmcl/map_likelihood.h"
#include "quickmcl/odometry.h"
#include "quickmcl/parameters.h"
#include "quickmcl/particle.h"
#include "quickmcl/particle_filter.h"
#include "quickmcl/particle_filter_factory.h"
#include "quickmcl/pose_restore.h"
#include "quickmcl/scaled_map.h"
#include "quickmcl/space_partitioning.h"
#include "quickmcl/statistics.h"
#include "quickmcl/utils.h"

#include <ros/console.h>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

namespace {

// quickmcl::WeightedParticle as the oracle and the C ABI lay it out (24 bytes)
struct PodParticle {
  float x, y, theta;
  uint32_t pad;
  double weight;
};
static_assert(sizeof(PodParticle) == 24, "particle layout");

quickmcl::ParticleCollection to_collection(const PodParticle *p, size_t n) {
  quickmcl::ParticleCollection c(n);
  for (size_t i = 0; i < n; ++i) {
    c[i].data.x = p[i].x;
    c[i].data.y = p[i].y;
    c[i].data.theta = p[i].theta;
    c[i].weight = p[i].weight;
  }
  return c;
}

void from_collection(const quickmcl::ParticleCollection &c, PodParticle *p) {
  for (size_t i = 0; i < c.size(); ++i) p[i].weight = c[i].weight;
}

void from_collection_full(const quickmcl::ParticleCollection &c, PodParticle *p) {
  for (size_t i = 0; i < c.size(); ++i) {
    p[i].x = c[i].data.x;
    p[i].y = c[i].data.y;
    p[i].theta = c[i].data.theta;
    p[i].pad = 0;
    p[i].weight = c[i].weight;
  }
}

quickmcl::LaserPointCloud to_cloud(const float *xy, size_t n) {
  quickmcl::LaserPointCloud cloud;
  cloud.points.resize(n);
  for (size_t i = 0; i < n; ++i) {
    cloud.points[i].x = xy[2 * i];
    cloud.points[i].y = xy[2 * i + 1];
  }
  return cloud;
}

// the script the stand-in engines read (ref_stubs/random)
// (per thread: bench.py times one reference filter per host thread, each with its own script)
thread_local std::vector<uint64_t> g_words;
thread_local std::vector<double> g_normals;
thread_local size_t g_word_pos = 0, g_normal_pos = 0;
thread_local int g_underflow = 0;
thread_local double g_last_seconds = 0;

struct Stopwatch {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  ~Stopwatch() { g_last_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

nav_msgs::OccupancyGrid make_grid(const int8_t *occ, uint32_t w, uint32_t h, float resolution) {
  nav_msgs::OccupancyGrid g;
  g.info.width = w;
  g.info.height = h;
  g.info.resolution = resolution;
  g.data.assign(occ, occ + static_cast<size_t>(w) * h);
  return g;
}

// distance_calculator.cpp:45-57 keeps one distance cache in a function-local static and
// rebuilds it when the resolution changes or when the cached range is >= the requested
// one - so a process that asks for a LARGER range than before at the same resolution is
// handed the old, too small cache and reads past its end.  One process building maps with
// different parameters is what these tests do and the node never does; a dummy request
// at another resolution before every real one makes the next call rebuild.
void forget_distance_cache() {
  const int8_t cell = 0;
  const nav_msgs::OccupancyGrid g = make_grid(&cell, 1, 1, 12345.0f);
  quickmcl::MapContainer m;
  quickmcl::compute_distance_map(12345.0f, g, &m);
}

}  // namespace

extern "C" {

// distance_calculator.cpp:86-101; returns the cache's side length, values row by row
int ref_distance_cache(float resolution, float max_dist, float *out, size_t cap) {
  const quickmcl::DistanceCache c = quickmcl::create_distance_cache(resolution, max_dist);
  const Eigen::Index n = c.rows();
  if (out && cap >= static_cast<size_t>(n * n))
    for (Eigen::Index r = 0; r < n; ++r)
      for (Eigen::Index col = 0; col < n; ++col) out[r * n + col] = c(r, col);
  return static_cast<int>(n);
}

// distance_calculator.cpp:103-182 (the brushfire), metres, row-major [y][x]
void ref_distance_map(const int8_t *occ, uint32_t w, uint32_t h, float resolution, float max_dist, float *out) {
  const nav_msgs::OccupancyGrid g = make_grid(occ, w, h, resolution);
  quickmcl::MapContainer m;
  forget_distance_cache();
  quickmcl::compute_distance_map(max_dist, g, &m);
  for (Eigen::Index r = 0; r < m.rows(); ++r)
    for (Eigen::Index c = 0; c < m.cols(); ++c) out[r * m.cols() + c] = m(r, c);
}

// distance_calculator.cpp:184-209
void ref_state_map(const int8_t *occ, uint32_t w, uint32_t h, int8_t *out) {
  const nav_msgs::OccupancyGrid g = make_grid(occ, w, h, 1.0f);
  quickmcl::MapStateContainer m;
  quickmcl::compute_state_map(g, &m);
  for (Eigen::Index r = 0; r < m.rows(); ++r)
    for (Eigen::Index c = 0; c < m.cols(); ++c) out[r * m.cols() + c] = static_cast<int8_t>(m(r, c));
}

// scaled_map.h: ScaledMapLogic<2> configured like MapBase::new_map (map_base.cpp:38-41)
void ref_world_to_map2(const float *pos_xy, size_t n, const float offset[2], float resolution, int *out_xy) {
  quickmcl::ScaledMapLogic<2> logic;
  using L = quickmcl::ScaledMapLogic<2>;
  logic.reconfigure_from_map_size(L::MapCoordinate{1, 1}, L::WorldCoordinate{offset[0], offset[1]},
                                  L::WorldCoordinate{resolution, resolution});
  for (size_t i = 0; i < n; ++i) {
    const L::MapCoordinate c = logic.world_to_map(L::WorldCoordinate{pos_xy[2 * i], pos_xy[2 * i + 1]});
    out_xy[2 * i] = c(0);
    out_xy[2 * i + 1] = c(1);
  }
}

void ref_map_to_world2(const int *cell_xy, size_t n, const float offset[2], float resolution, float *out_xy) {
  quickmcl::ScaledMapLogic<2> logic;
  using L = quickmcl::ScaledMapLogic<2>;
  logic.reconfigure_from_map_size(L::MapCoordinate{1, 1}, L::WorldCoordinate{offset[0], offset[1]},
                                  L::WorldCoordinate{resolution, resolution});
  for (size_t i = 0; i < n; ++i) {
    const L::WorldCoordinate c = logic.map_to_world(L::MapCoordinate{cell_xy[2 * i], cell_xy[2 * i + 1]});
    out_xy[2 * i] = c(0);
    out_xy[2 * i + 1] = c(1);
  }
}

int ref_is_in_map2(const int cell[2], const int size[2]) {
  quickmcl::ScaledMapLogic<2> logic;
  using L = quickmcl::ScaledMapLogic<2>;
  logic.reconfigure_from_map_size(L::MapCoordinate{size[0], size[1]}, L::WorldCoordinate{0.0f, 0.0f},
                                  L::WorldCoordinate{1.0f, 1.0f});
  return logic.is_in_map(L::MapCoordinate{cell[0], cell[1]}) ? 1 : 0;
}

// particle.cpp:32-82
double ref_effective_particles(const PodParticle *p, size_t n) {
  return quickmcl::effective_particles(to_collection(p, n));
}
void ref_normalise_weights(PodParticle *p, size_t n) {
  auto c = to_collection(p, n);
  quickmcl::normalise_particle_weights(&c);
  from_collection(c, p);
}
void ref_normalise_weights_total(PodParticle *p, size_t n, double total) {
  auto c = to_collection(p, n);
  quickmcl::normalise_particle_weights(&c, total);
  from_collection(c, p);
}
void ref_equalise_weights(PodParticle *p, size_t n) {
  auto c = to_collection(p, n);
  quickmcl::equalise_particle_weights(&c);
  from_collection(c, p);
}
// ParticlesRunningSum (particle.cpp:66-82, particle.h:99-107): total weight, one index per query
double ref_running_sum_lookup(const PodParticle *p, size_t n, const double *queries, size_t nq, uint64_t *idx_out) {
  const quickmcl::ParticlesRunningSum rs(to_collection(p, n));
  for (size_t i = 0; i < nq; ++i) idx_out[i] = rs.lookup_index(queries[i]);
  return rs.get_total_weight();
}

// low_pass_filter.h:39-50
double ref_lowpass_run(double alpha, const double *xs, size_t n, int as_float) {
  if (as_float) {
    quickmcl::LowPassFilter<float> f;
    f.set_alpha(static_cast<float>(alpha));
    for (size_t i = 0; i < n; ++i) f.update(static_cast<float>(xs[i]));
    return f.get();
  }
  quickmcl::LowPassFilter<double> f;
  f.set_alpha(alpha);
  for (size_t i = 0; i < n; ++i) f.update(xs[i]);
  return f.get();
}

// utils.h:28-61
double ref_normalise_angle_f64(double a) { return quickmcl::normalise_angle(a); }
float ref_normalise_angle_f32(float a) { return quickmcl::normalise_angle(a); }
double ref_angle_delta_f64(double a, double b) { return quickmcl::angle_delta(a, b); }
float ref_angle_delta_f32(float a, float b) { return quickmcl::angle_delta(a, b); }

// pose_2d.h: Pose2D::to_decomposed
void ref_to_decomposed(const float pose[3], double out[4]) {
  const Eigen::Vector4d d = quickmcl::Pose2D<float>(pose[0], pose[1], pose[2]).to_decomposed();
  for (int i = 0; i < 4; ++i) out[i] = d(i);
}

// space_partitioning.cpp.  The particles are added one by one, as the KLD loop of
// particle_filter.cpp:313-335 does; after each add_point: limits_out[i] = limit().
// Then assign_clusters(): keys_out[3 i ..] = the particle's bin, cluster_out[i] = its
// cluster id (numbering follows the unordered_map's iteration order: compare partitions).
// Returns the number of clusters.
int ref_space_partitioning(const PodParticle *p, size_t n, int32_t particle_count_min, int32_t particle_count_max,
                           double kld_epsilon, double kld_z, float resolution_xy, float resolution_theta,
                           uint64_t *limits_out, int32_t *keys_out, int32_t *cluster_out) {
  quickmcl::Parameters::ParticleFilter pf;
  pf.particle_count_min = particle_count_min;
  pf.particle_count_max = particle_count_max;
  pf.kld_epsilon = kld_epsilon;
  pf.kld_z = kld_z;
  pf.space_partitioning_resolution_xy = resolution_xy;
  pf.space_partitioning_resolution_theta = resolution_theta;
  quickmcl::SpacePartitioning sp;
  sp.set_parameters(pf);
  sp.reset();
  for (size_t i = 0; i < n; ++i) {
    const quickmcl::WeightedParticle::ParticleT pose(p[i].x, p[i].y, p[i].theta);
    sp.add_point(pose);
    if (limits_out) limits_out[i] = sp.limit();
  }
  sp.assign_clusters();
  for (size_t i = 0; i < n; ++i) {
    const quickmcl::WeightedParticle::ParticleT pose(p[i].x, p[i].y, p[i].theta);
    if (keys_out) {
      const auto key = sp.world_to_map(quickmcl::SpacePartitioning::WorldCoordinate(pose.x, pose.y, pose.theta));
      keys_out[3 * i] = key(0);
      keys_out[3 * i + 1] = key(1);
      keys_out[3 * i + 2] = key(2);
    }
    if (cluster_out) cluster_out[i] = sp.get_cluster_id(pose);
  }
  return sp.get_cluster_count();
}

// ----------------------------------------------------------------- the script
uint64_t qmcl_ref_next_word(void) {
  if (g_word_pos < g_words.size()) return g_words[g_word_pos++];
  g_underflow = 1;
  return 0;
}
double qmcl_ref_next_normal(void) {
  if (g_normal_pos < g_normals.size()) return g_normals[g_normal_pos++];
  g_underflow = 1;
  return 0;
}
// replaces the script: `words` feed every uniform draw, `normals` every normal draw
void ref_script(const uint64_t *words, size_t n_words, const double *normals, size_t n_normals) {
  g_words.assign(words, words + n_words);
  g_normals.assign(normals, normals + n_normals);
  g_word_pos = g_normal_pos = 0;
  g_underflow = 0;
}
// consumed[0] = words read, consumed[1] = normals read; returns 1 if a draw found the script empty
int ref_script_state(uint64_t consumed[2]) {
  consumed[0] = g_word_pos;
  consumed[1] = g_normal_pos;
  return g_underflow;
}
// wall-clock seconds of the last timed entry point (sensor update, resampling, cycle)
double ref_last_seconds(void) { return g_last_seconds; }

// ------------------------------------------------- MapLikelihood (+ MapBase)
struct ref_map {
  std::shared_ptr<quickmcl::MapLikelihood> map;
};
struct RefLikelihoodParams {  // = orc_likelihood_params
  float z_hit, z_rand;
  int32_t num_beams;
  float max_laser_distance, max_obstacle_distance, sigma_hit;
};

// set_parameters (map_likelihood.cpp:39-48) then new_map (:50-61, map_base.cpp:29-56)
ref_map *ref_map_create(const int8_t *occ, uint32_t w, uint32_t h, float resolution, double origin_x,
                        double origin_y, const RefLikelihoodParams *p) {
  quickmcl::Parameters::LikelihoodMap lp;
  lp.z_hit = p->z_hit;
  lp.z_rand = p->z_rand;
  lp.num_beams = p->num_beams;
  lp.max_laser_distance = p->max_laser_distance;
  lp.max_obstacle_distance = p->max_obstacle_distance;
  lp.sigma_hit = p->sigma_hit;
  auto grid = std::make_shared<nav_msgs::OccupancyGrid>(make_grid(occ, w, h, resolution));
  grid->info.origin.position.x = origin_x;
  grid->info.origin.position.y = origin_y;
  auto *m = new ref_map;
  // through the reference's factory, as main.cpp does (map_factory.cpp:23-26)
  m->map = std::dynamic_pointer_cast<quickmcl::MapLikelihood>(quickmcl::create_likelihood_map());
  m->map->set_parameters(lp);
  forget_distance_cache();
  m->map->new_map(grid);
  return m;
}
void ref_map_destroy(ref_map *m) { delete m; }
void ref_map_copy_field(const ref_map *m, float *out) {
  const auto &f = m->map->likelihood_data;
  std::memcpy(out, f.data(), sizeof(float) * static_cast<size_t>(f.rows() * f.cols()));
}
// replaces the likelihood table (same shape) so that two implementations can be
// compared on one identical field
void ref_map_set_field(ref_map *m, const float *field) {
  auto &f = m->map->likelihood_data;
  std::memcpy(f.data(), field, sizeof(float) * static_cast<size_t>(f.rows() * f.cols()));
}
void ref_map_copy_state(const ref_map *m, int8_t *out) {
  const auto &s = m->map->map_state;
  for (Eigen::Index r = 0; r < s.rows(); ++r)
    for (Eigen::Index c = 0; c < s.cols(); ++c) out[r * s.cols() + c] = static_cast<int8_t>(s(r, c));
}
size_t ref_map_num_free(const ref_map *m) { return m->map->free_cells.size(); }
void ref_map_copy_free_cells(const ref_map *m, int32_t *xy_out) {
  const auto &fc = m->map->free_cells;
  for (size_t i = 0; i < fc.size(); ++i) {
    xy_out[2 * i] = fc[i](0);
    xy_out[2 * i + 1] = fc[i](1);
  }
}
int8_t ref_map_state_at(const ref_map *m, float x, float y) {
  return static_cast<int8_t>(m->map->get_map_state_at(Eigen::Vector2f(x, y)));
}
double ref_map_probability_at(const ref_map *m, float x, float y) {
  return m->map->get_probability_at(Eigen::Vector2f(x, y));
}
void ref_map_likelihood_grid(const ref_map *m, int8_t *out) {
  const nav_msgs::OccupancyGrid g = m->map->get_likelihood_as_gridmap();
  std::memcpy(out, g.data.data(), g.data.size());
}
// map_base.cpp:58-64; reads two words of the script (cell, heading)
void ref_map_random_pose(const ref_map *m, float out[3]) {
  const auto p = m->map->generate_random_pose();
  out[0] = p.x;
  out[1] = p.y;
  out[2] = p.theta;
}
// map_likelihood.cpp:63-130: weights written into p, returns the total
double ref_map_update_importance(const ref_map *m, const float *xy, size_t npoints, PodParticle *p, size_t n) {
  const quickmcl::LaserPointCloud cloud = to_cloud(xy, npoints);
  auto c = to_collection(p, n);
  double total;
  {
    Stopwatch sw;
    total = m->map->update_importance(cloud, &c);
  }
  from_collection(c, p);
  return total;
}

// statistics.cpp: add every particle in order, optionally finalise;
// out = {total_weight, sample_count, mean[4], covariance[9] row-major}
void ref_statistics(const PodParticle *p, size_t n, int finalise, double out[15]) {
  quickmcl::ParticleCloudStatistics st;
  st.add(to_collection(p, n));
  if (finalise) st.finalise();
  out[0] = st.total_weight;
  out[1] = static_cast<double>(st.sample_count);
  for (int k = 0; k < 4; ++k) out[2 + k] = st.mean(k);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) out[6 + 3 * i + j] = st.covariance(i, j);
}

// odometry.cpp:44-102; reads three normals of the script per particle
void ref_sample_odometry(const double odom_new[3], const double odom_old[3], PodParticle *p, size_t n,
                         const float alpha[4]) {
  auto c = to_collection(p, n);
  std::mt19937 rng;
  float a[4] = {alpha[0], alpha[1], alpha[2], alpha[3]};
  quickmcl::sample_odometry(quickmcl::Odometry(odom_new[0], odom_new[1], odom_new[2]),
                            quickmcl::Odometry(odom_old[0], odom_old[1], odom_old[2]), &c, &rng, a);
  from_collection_full(c, p);
}

// ------------------------------------------------------------ ParticleFilter
struct ref_filter {
  std::shared_ptr<quickmcl::MapLikelihood> map;
  std::shared_ptr<quickmcl::ParticleFilter> pf;
};
struct RefPfParams {  // = orc_pf_params
  int32_t resample_type, particle_count_min, particle_count_max, resample_count;
  double alpha_fast, alpha_slow, kld_epsilon, kld_z;
  float res_xy, res_theta;
};

ref_filter *ref_filter_create(ref_map *m) {
  auto *f = new ref_filter;
  f->map = m->map;
  // through the reference's factory (particle_filter_factory.cpp:23-27)
  f->pf = std::dynamic_pointer_cast<quickmcl::ParticleFilter>(quickmcl::create_particle_filter(m->map));
  return f;
}
void ref_filter_destroy(ref_filter *f) { delete f; }
void ref_filter_set_params(ref_filter *f, const RefPfParams *p) {
  quickmcl::Parameters::ParticleFilter pp;
  pp.resample_type = p->resample_type == 0   ? quickmcl::ResampleType::low_variance
                     : p->resample_type == 1 ? quickmcl::ResampleType::adaptive
                                             : quickmcl::ResampleType::kld;
  pp.particle_count_min = p->particle_count_min;
  pp.particle_count_max = p->particle_count_max;
  pp.resample_count = p->resample_count;
  pp.alpha_fast = p->alpha_fast;
  pp.alpha_slow = p->alpha_slow;
  pp.kld_epsilon = p->kld_epsilon;
  pp.kld_z = p->kld_z;
  pp.space_partitioning_resolution_xy = p->res_xy;
  pp.space_partitioning_resolution_theta = p->res_theta;
  f->pf->set_parameters(pp);
}
// particle_filter.cpp:39-68; reads three normals per particle
void ref_filter_initialise(ref_filter *f, const float pose[3], const float cov[9]) {
  quickmcl::WeightedParticle::ParticleT::EigenMatrix c;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) c(i, j) = cov[3 * i + j];
  f->pf->initialise(quickmcl::WeightedParticle::ParticleT(pose[0], pose[1], pose[2]), c);
}
// particle_filter.cpp:70-82; reads two words per particle
void ref_filter_global_localization(ref_filter *f) { f->pf->global_localization(); }
void ref_filter_handle_odometry(ref_filter *f, const double odom_new[3], const double odom_old[3],
                                const float alpha[4]) {
  float a[4] = {alpha[0], alpha[1], alpha[2], alpha[3]};
  Stopwatch sw;
  f->pf->handle_odometry(quickmcl::Odometry(odom_new[0], odom_new[1], odom_new[2]),
                         quickmcl::Odometry(odom_old[0], odom_old[1], odom_old[2]), a);
}
void ref_filter_update_sensor(ref_filter *f, const float *xy, size_t npoints) {
  const quickmcl::LaserPointCloud cloud = to_cloud(xy, npoints);
  Stopwatch sw;
  f->pf->update_importance_from_observations(cloud);
}
// returns 0, or 1 when the reference threw std::out_of_range: its low-variance walk
// ran off the end of the set (particle_filter.cpp:229-235, the assert is compiled out in
// Release and .at(i) throws)
int ref_filter_resample_and_cluster(ref_filter *f) {
  Stopwatch sw;
  try {
    f->pf->resample_and_cluster();
  } catch (const std::out_of_range &) {
    return 1;
  }
  return 0;
}
void ref_filter_cluster(ref_filter *f) { f->pf->cluster(); }
int ref_filter_get_estimate(const ref_filter *f, double pose[3], double cov[9]) {
  quickmcl::Pose2D<double> p(0, 0, 0);
  Eigen::Matrix3d c = Eigen::Matrix3d::Zero();
  bool ok;
  {
    Stopwatch sw;
    ok = f->pf->get_estimated_pose(&p, &c);
  }
  if (!ok) return 0;
  pose[0] = p.x;
  pose[1] = p.y;
  pose[2] = p.theta;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) cov[3 * i + j] = c(i, j);
  return 1;
}
size_t ref_filter_num_particles(const ref_filter *f) { return f->pf->num_particles(); }
void ref_filter_copy_particles(const ref_filter *f, PodParticle *out) {
  from_collection_full(f->pf->get_particles(), out);
}
void ref_filter_set_particles(ref_filter *f, const PodParticle *p, size_t n) {
  f->pf->data_sets[f->pf->current_data_set].particles = to_collection(p, n);
}
void ref_filter_get_lowpass(const ref_filter *f, double out[2]) {
  out[0] = f->pf->w_slow.get();
  out[1] = f->pf->w_fast.get();
}
void ref_filter_set_lowpass(ref_filter *f, double w_slow, double w_fast) {
  f->pf->w_slow.reset(w_slow);
  f->pf->w_fast.reset(w_fast);
}
int ref_filter_num_clusters(const ref_filter *f) { return f->pf->space_partitioning.get_cluster_count(); }
// cluster id of every current particle (numbering follows the unordered_map: compare partitions)
void ref_filter_cluster_ids(const ref_filter *f, int32_t *out) {
  const auto &ps = f->pf->get_particles();
  for (size_t i = 0; i < ps.size(); ++i) out[i] = f->pf->space_partitioning.get_cluster_id(ps[i].data);
}

// -------------------------------------------------------------- PoseRestorer
namespace {
// an IParticleFilter that records what PoseRestorer does with it
struct RecordingFilter final : quickmcl::IParticleFilter {
  bool has_estimate = false;
  quickmcl::Pose2D<double> estimate{0, 0, 0};
  Eigen::Matrix3d estimate_cov = Eigen::Matrix3d::Zero();
  int initialise_calls = 0;
  quickmcl::WeightedParticle::ParticleT init_pose{0, 0, 0};
  quickmcl::WeightedParticle::ParticleT::EigenMatrix init_cov;
  quickmcl::ParticleCollection none;

  void initialise(const quickmcl::WeightedParticle::ParticleT &p,
                  const quickmcl::WeightedParticle::ParticleT::EigenMatrix &c) override {
    ++initialise_calls;
    init_pose = p;
    init_cov = c;
  }
  void global_localization() override {}
  void handle_odometry(const quickmcl::Odometry &, const quickmcl::Odometry &, float[4]) override {}
  void update_importance_from_observations(const quickmcl::LaserPointCloud &) override {}
  void resample_and_cluster() override {}
  void cluster() override {}
  void set_parameters(const quickmcl::Parameters::ParticleFilter &) override {}
  bool get_estimated_pose(quickmcl::Pose2D<double> *pose, Eigen::Matrix3d *cov) const override {
    if (!has_estimate) return false;
    *pose = estimate;
    *cov = estimate_cov;
    return true;
  }
  const quickmcl::ParticleCollection &get_particles() const override { return none; }
  size_t num_particles() const override { return 0; }
};

size_t copy_warnings(char *out, size_t cap) {
  std::string all;
  for (const auto &line : ros::stand_in_warnings()) all += line + "\n";
  ros::stand_in_warnings().clear();
  if (out && cap) {
    const size_t n = all.size() < cap - 1 ? all.size() : cap - 1;
    std::memcpy(out, all.data(), n);
    out[n] = 0;
  }
  return all.size();
}
}  // namespace

// PoseRestorer::restore_pose (pose_restore.cpp:57-78).  n < 0: no ~initial_pose parameter.
// Returns how often the filter was initialised; pose / cov = what it was initialised with;
// warnings = the ROS_WARN lines, one per line.
int ref_pose_restore(const double *v, int n, float pose_out[3], float cov_out[9], char *warnings, size_t cap) {
  ros::stand_in_parameters().clear();
  ros::stand_in_warnings().clear();
  if (n >= 0) ros::stand_in_parameters()["initial_pose"] = std::vector<double>(v, v + n);
  auto filter = std::make_shared<RecordingFilter>();
  quickmcl::PoseRestorer restorer(ros::NodeHandle(), filter);
  restorer.restore_pose();
  pose_out[0] = filter->init_pose.x;
  pose_out[1] = filter->init_pose.y;
  pose_out[2] = filter->init_pose.theta;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) cov_out[3 * i + j] = filter->init_cov(i, j);
  copy_warnings(warnings, cap);
  return filter->initialise_calls;
}

// PoseRestorer::store_pose (pose_restore.cpp:37-55).  Returns the length of the stored
// ~initial_pose vector (copied to out6), or -1 when nothing was stored.
int ref_pose_store(int has_estimate, const double pose[3], const double cov[9], double out6[6], char *warnings,
                   size_t cap) {
  ros::stand_in_parameters().clear();
  ros::stand_in_warnings().clear();
  auto filter = std::make_shared<RecordingFilter>();
  filter->has_estimate = has_estimate != 0;
  filter->estimate = quickmcl::Pose2D<double>(pose[0], pose[1], pose[2]);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) filter->estimate_cov(i, j) = cov[3 * i + j];
  quickmcl::PoseRestorer restorer(ros::NodeHandle(), filter);
  restorer.store_pose();
  copy_warnings(warnings, cap);
  const auto it = ros::stand_in_parameters().find("initial_pose");
  if (it == ros::stand_in_parameters().end()) return -1;
  for (size_t i = 0; i < it->second.size() && i < 6; ++i) out6[i] = it->second[i];
  return static_cast<int>(it->second.size());
}

}  // extern "C"
