// quickmcl_b200.cpp — the C++ host mirror: reference-shaped classes that
// forward to the C ABI of libqmcl_b200.so.  No computation happens here.
#include "quickmcl_b200.h"

#include "../../include/qmcl.h"

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>

namespace quickmcl_b200 {

namespace {

// Error convention.  The reference has no status returns: it logs through
// rosconsole and carries on, and abort()s on an invalid cluster id
// (particle_filter.cpp:375-379).  Here a failed device call is never silent:
// it throws, and the invalid-cluster case aborts like the reference.
void check(qmcl_status st, const char *where) {
  if (st == QMCL_OK) return;
  const std::string msg = std::string(where) + ": status " + std::to_string(st) + " " +
                          qmcl_last_error();
  if (st == QMCL_ERR_INVALID_CLUSTER) {
    std::fprintf(stderr, "[FATAL] %s\n", msg.c_str());
    std::abort();
  }
  throw std::runtime_error(msg);
}

static_assert(sizeof(WeightedParticle) == sizeof(qmcl_particle), "layout");

}  // namespace

// ---------------------------------------------------------------------- map
CudaMapLikelihood::CudaMapLikelihood(int device) {
  check(qmcl_map_create(device, &map_), "qmcl_map_create");
}

CudaMapLikelihood::~CudaMapLikelihood() {
  if (map_) qmcl_map_destroy(map_);
}

void CudaMapLikelihood::set_parameters(const Parameters::LikelihoodMap &p) {
  qmcl_likelihood_params q{p.z_hit,           p.z_rand,           p.num_beams, p.max_laser_distance,
                           p.max_obstacle_distance, p.sigma_hit};
  check(qmcl_map_set_params(map_, &q), "qmcl_map_set_params");
}

void CudaMapLikelihood::new_map(const OccupancyGrid::ConstPtr &msg) {
  info_ = msg->info;
  check(qmcl_map_load(map_, msg->data.data(), msg->info.width, msg->info.height,
                      msg->info.resolution, msg->info.origin.position.x,
                      msg->info.origin.position.y),
        "qmcl_map_load");
}

double CudaMapLikelihood::update_importance(const LaserPointCloud &cloud,
                                            ParticleCollection *particles) const {
  double total = 0.0;
  check(qmcl_map_update_importance(map_, reinterpret_cast<const float *>(cloud.data()),
                                   cloud.size(),
                                   reinterpret_cast<qmcl_particle *>(particles->data()),
                                   particles->size(), &total),
        "qmcl_map_update_importance");
  return total;
}

WeightedParticle::ParticleT CudaMapLikelihood::generate_random_pose() const {
  float p[3];
  // The reference draws from a private mt19937 (map_base.h:61-68); here every
  // call consumes the next index of the shared counter-based stream.
  check(qmcl_map_generate_random_pose(map_, 0, 0xffffffffu, random_pose_counter_++, p),
        "qmcl_map_generate_random_pose");
  return WeightedParticle::ParticleT(p[0], p[1], p[2]);
}

OccupancyGrid CudaMapLikelihood::get_likelihood_as_gridmap() const {
  OccupancyGrid out;
  out.info = info_;
  out.data.resize(static_cast<size_t>(info_.width) * info_.height);
  if (!out.data.empty()) check(qmcl_map_likelihood_grid(map_, out.data.data()), "qmcl_map_likelihood_grid");
  return out;
}

MapState CudaMapLikelihood::get_map_state_at(const Vector2f &w) const {
  int8_t s = 2;
  check(qmcl_map_state_at(map_, w.x, w.y, &s), "qmcl_map_state_at");
  return static_cast<MapState>(s);
}

// ------------------------------------------------------------------- filter
CudaParticleFilter::CudaParticleFilter(const std::shared_ptr<Map> &map, uint64_t seed) : map_(map) {
  auto *cuda_map = dynamic_cast<CudaMapLikelihood *>(map.get());
  if (!cuda_map)
    throw std::runtime_error("CudaParticleFilter needs a CudaMapLikelihood (device-resident field)");
  check(qmcl_filter_create(cuda_map->handle(), seed, &filter_), "qmcl_filter_create");
}

CudaParticleFilter::~CudaParticleFilter() {
  if (filter_) qmcl_filter_destroy(filter_);
}

void CudaParticleFilter::initialise(const WeightedParticle::ParticleT &p, const Matrix3f &cov) {
  const float pose[3] = {p.x, p.y, p.theta};
  check(qmcl_filter_initialise(filter_, pose, cov.m), "qmcl_filter_initialise");
  mirror_valid_ = false;
}

void CudaParticleFilter::global_localization() {
  check(qmcl_filter_global_localization(filter_), "qmcl_filter_global_localization");
  mirror_valid_ = false;
}

void CudaParticleFilter::handle_odometry(const Odometry &n, const Odometry &o, float alpha[4]) {
  const double a[3] = {n.x, n.y, n.theta}, b[3] = {o.x, o.y, o.theta};
  check(qmcl_filter_handle_odometry(filter_, a, b, alpha), "qmcl_filter_handle_odometry");
  mirror_valid_ = false;
}

void CudaParticleFilter::update_importance_from_observations(const LaserPointCloud &cloud) {
  check(qmcl_filter_update_sensor(filter_, reinterpret_cast<const float *>(cloud.data()),
                                  cloud.size()),
        "qmcl_filter_update_sensor");
  mirror_valid_ = false;
}

void CudaParticleFilter::resample_and_cluster() {
  check(qmcl_filter_resample_and_cluster(filter_), "qmcl_filter_resample_and_cluster");
  mirror_valid_ = false;
}

void CudaParticleFilter::cluster() { check(qmcl_filter_cluster(filter_), "qmcl_filter_cluster"); }

void CudaParticleFilter::set_parameters(const Parameters::ParticleFilter &p) {
  qmcl_pf_params q;
  q.resample_type = static_cast<int32_t>(p.resample_type);
  q.particle_count_min = p.particle_count_min;
  q.particle_count_max = p.particle_count_max;
  q.resample_count = p.resample_count;
  q.alpha_fast = p.alpha_fast;
  q.alpha_slow = p.alpha_slow;
  q.kld_epsilon = p.kld_epsilon;
  q.kld_z = p.kld_z;
  q.res_xy = p.space_partitioning_resolution_xy;
  q.res_theta = p.space_partitioning_resolution_theta;
  check(qmcl_filter_set_params(filter_, &q), "qmcl_filter_set_params");
}

bool CudaParticleFilter::get_estimated_pose(Pose2D<double> *pose, Matrix3d *covariance) const {
  double p[3], c[9];
  int valid = 0;
  check(qmcl_filter_get_estimate(filter_, p, c, &valid), "qmcl_filter_get_estimate");
  if (!valid) return false;  // "ALL clusters suck" (particle_filter.cpp:206-211): outputs untouched
  *pose = Pose2D<double>(p[0], p[1], p[2]);
  for (int k = 0; k < 9; ++k) covariance->m[k] = c[k];
  return true;
}

const ParticleCollection &CudaParticleFilter::get_particles() const {
  if (!mirror_valid_) {
    size_t n = 0;
    check(qmcl_filter_num_particles(filter_, &n), "qmcl_filter_num_particles");
    mirror_.resize(n);
    if (n)
      check(qmcl_filter_copy_particles(filter_, reinterpret_cast<qmcl_particle *>(mirror_.data()), n),
            "qmcl_filter_copy_particles");
    mirror_valid_ = true;
  }
  return mirror_;
}

std::vector<CudaParticleFilter::Marker> CudaParticleFilter::export_markers(double *max_weight) const {
  static_assert(sizeof(Marker) == 5 * sizeof(float), "Marker must be five packed floats");
  size_t n = 0;
  check(qmcl_filter_num_particles(filter_, &n), "qmcl_filter_num_particles");
  std::vector<Marker> out(n);
  size_t got = 0;
  check(qmcl_filter_export_markers(filter_, reinterpret_cast<float *>(out.data()), n, &got,
                                   max_weight),
        "qmcl_filter_export_markers");
  out.resize(got);
  return out;
}

ScanCycle::Result ScanCycle::on_scan(const Pose2D<double> &odom_new, const LaserPointCloud &cloud) {
  Result r;
  if (!have_odom_) {  // laser_handler.cpp:196-199
    odom_at_last_filter_run_ = odom_new;
    have_odom_ = true;
    r.recompute_transform = true;
  }
  // :205-210: diff.normalise() wraps the heading difference (utils.h:28-35)
  const double two_pi = 2 * 3.14159265358979323846;
  const double dx = odom_new.x - odom_at_last_filter_run_.x;
  const double dy = odom_new.y - odom_at_last_filter_run_.y;
  double dt = odom_new.theta - odom_at_last_filter_run_.theta;
  dt = dt - two_pi * std::floor((dt + two_pi / 2) / two_pi);
  const bool has_moved = std::fabs(dx) > motion.min_trans || std::fabs(dy) > motion.min_trans ||
                         std::fabs(dt) > motion.min_rot;
  auto estimate = [&]() {
    calls.push_back("get_estimated_pose");
    r.pose_ok = filter_->get_estimated_pose(&r.pose, &r.covariance);
  };
  if (!pose_reset_ && !has_moved) {  // :212-219
    calls.push_back("cluster");
    filter_->cluster();
    estimate();
    return r;
  }
  if (pose_reset_) r.recompute_transform = true;
  pose_reset_ = false;
  r.processed = true;
  if (filter_->num_particles() == 0) {  // get_particles().empty() without the copy (:236-241)
    calls.push_back("global_localization");
    filter_->global_localization();
    r.recompute_transform = true;
    r.global_localization = true;
  }
  calls.push_back("handle_odometry");
  filter_->handle_odometry(odom_new, odom_at_last_filter_run_, motion.alpha);  // :243-251
  odom_at_last_filter_run_ = odom_new;
  calls.push_back("update_importance_from_observations");
  filter_->update_importance_from_observations(cloud);  // :252-256
  if (resample_counter_++ % static_cast<uint32_t>(resample_count_)) {  // :259-270
    r.recompute_transform = true;
    r.resampled = true;
    calls.push_back("resample_and_cluster");
    filter_->resample_and_cluster();
  } else {
    calls.push_back("cluster");
    filter_->cluster();
  }
  estimate();
  return r;
}

std::vector<double> PoseCheckpoint::store() const {
  Pose2D<double> pose;
  Matrix3d cov;
  if (!filter_->get_estimated_pose(&pose, &cov)) return {};
  return {pose.x, pose.y, pose.theta, cov(0, 0), cov(1, 1), cov(2, 2)};
}

void PoseCheckpoint::restore(const std::vector<double> *v) const {
  const double defaults[6] = {0.0, 0.0, 0.0, 0.5 * 0.5, 0.5 * 0.5, 0.1 * 0.1};
  double val[6] = {0.0, 0.0, 0.0, 1.0, 1.0, 1.0};  // no parameter: identity covariance
  if (v)
    for (size_t i = 0; i < 6; ++i)
      val[i] = (i < v->size() && !std::isnan((*v)[i])) ? (*v)[i] : defaults[i];
  WeightedParticle::ParticleT start;
  start.x = static_cast<float>(val[0]);
  start.y = static_cast<float>(val[1]);
  start.theta = static_cast<float>(val[2]);
  Matrix3f cov{};
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) cov(r, c) = (r == c) ? static_cast<float>(val[3 + r]) : 0.f;
  filter_->initialise(start, cov);
}

size_t CudaParticleFilter::num_particles() const {
  size_t n = 0;
  check(qmcl_filter_num_particles(filter_, &n), "qmcl_filter_num_particles");
  return n;
}

// ----------------------------------------------------------------- factories
std::shared_ptr<Map> create_likelihood_map() { return std::make_shared<CudaMapLikelihood>(); }

std::shared_ptr<IParticleFilter> create_particle_filter(const std::shared_ptr<Map> &map) {
  return std::make_shared<CudaParticleFilter>(map);
}

}  // namespace quickmcl_b200
