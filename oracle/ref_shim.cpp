/*
 * ref_shim.cpp — C entry points over the REFERENCE's own code (oracle/_ref).
 *
 * TEST INFRASTRUCTURE ONLY.  This file is compiled together with a handful of the
 * reference's translation units, taken from where they lie under /root/reference
 * (never copied): distance_calculator.cpp, particle.cpp, pose_2d.cpp,
 * space_partitioning.cpp, plus its header-only scaled_map.h, low_pass_filter.h and
 * utils.h.  Those units use Eigen only as a container / for element-wise float
 * operations and ROS only for two message structs, so they build against the
 * stand-in headers in oracle/ref_stubs (see the note in ref_stubs/Eigen/Core).  The
 * result, oracle/_ref/libqmcl_ref.so, lets tests/test_oracle_vs_ref.py hold the
 * restated oracle against what the reference's statements actually compute, on
 * inputs far beyond the reference's own golden vectors: brushfire distance and
 * state maps of random occupancy grids, grid truncation, weights / N_eff / running
 * sum, the KLD bound as bins fill up, the bin keys and the cluster partition.
 *
 * What is NOT here: the sensor model, the motion model, the resamplers, the
 * statistics - they need real Eigen numerics, PCL or the reference's private RNG.
 */
#include "quickmcl/distance_calculator.h"
#include "quickmcl/low_pass_filter.h"
#include "quickmcl/parameters.h"
#include "quickmcl/particle.h"
#include "quickmcl/scaled_map.h"
#include "quickmcl/space_partitioning.h"
#include "quickmcl/utils.h"

#include <cstdint>
#include <cstring>
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

nav_msgs::OccupancyGrid make_grid(const int8_t *occ, uint32_t w, uint32_t h, float resolution) {
  nav_msgs::OccupancyGrid g;
  g.info.width = w;
  g.info.height = h;
  g.info.resolution = resolution;
  g.data.assign(occ, occ + static_cast<size_t>(w) * h);
  return g;
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

}  // extern "C"
