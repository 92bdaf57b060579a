// quickmcl_b200.h — C++ host mirror of QuickMCL's plugin seam over the C ABI.
//
// The reference's node touches the filter and the map only through two
// abstract classes and two factories (src/quickmcl_node/main.cpp:18-23):
//   quickmcl::IParticleFilter   include/quickmcl/i_particle_filter.h:30-85
//   quickmcl::Map               include/quickmcl/map.h:31-62
//   create_particle_filter(map) include/quickmcl/particle_filter_factory.h:30-31
//   create_likelihood_map()     include/quickmcl/map_factory.h:29
// This header declares the same interface — same method names, argument
// meaning and ownership — and a CUDA-backed implementation of each on top of
// libqmcl_b200.so (include/qmcl.h).
//
// ROS, Eigen and PCL are not installed in this image, so the message / matrix
// / point-cloud types are the small stand-ins below (only the members the path
// reads).  INTEGRATION.md shows the same adapter written against the real
// types inside the reference tree.
#pragma once

#include <cstddef>
#include <cstdint>
#include <memory>
#include <string>
#include <vector>

struct qmcl_map;
struct qmcl_filter;

namespace quickmcl_b200 {

// ---- stand-ins for types of the reference's dependencies ------------------
// Eigen::Matrix<T,3,3> (row/col access only), Eigen::Vector2f.
template <typename T> struct Matrix3 {
  T m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  T &operator()(int r, int c) { return m[r * 3 + c]; }
  const T &operator()(int r, int c) const { return m[r * 3 + c]; }
};
using Matrix3f = Matrix3<float>;
using Matrix3d = Matrix3<double>;
struct Vector2f {
  float x = 0, y = 0;
};
// pcl::PointCloud<pcl::PointXY> (include/quickmcl/laser.h:28): size(), at(), {x, y}.
struct PointXY {
  float x, y;
};
using LaserPointCloud = std::vector<PointXY>;
// The members of nav_msgs::OccupancyGrid the map reads (map_base.cpp:30-40).
struct OccupancyGrid {
  struct Info {
    float resolution = 0;
    uint32_t width = 0, height = 0;
    struct { struct { double x = 0, y = 0; } position; } origin;
  } info;
  std::vector<int8_t> data;  // row-major [y][x]
  using ConstPtr = std::shared_ptr<const OccupancyGrid>;
};

// ---- value types of the reference (same layouts) -----------------------------
// pose_2d.h: Pose2D<T> {x, y, theta}
template <typename T = float> struct Pose2D {
  T x, y, theta;
  Pose2D() {}
  Pose2D(T x_, T y_, T theta_) : x(x_), y(y_), theta(theta_) {}
};
using Odometry = Pose2D<double>;  // odometry.h:29
// particle.h:32-47: 24 bytes, weight defaults to 1
struct WeightedParticle {
  using ParticleT = Pose2D<float>;
  ParticleT data;
  double weight = 1.0;
};
static_assert(sizeof(WeightedParticle) == 24, "must match qmcl_particle");
using ParticleCollection = std::vector<WeightedParticle>;
// map_state.h:26-33
enum class MapState : int8_t { Free = 0, Occupied = 1, Unknown = 2 };
// parameters.h:28-113 (same fields, same defaults)
enum class ResampleType { low_variance, adaptive, kld };
struct Parameters {
  struct ParticleFilter {
    ResampleType resample_type = ResampleType::kld;
    int32_t particle_count_min = 100;
    int32_t particle_count_max = 5000;
    int32_t resample_count = 2;
    double alpha_fast = 0.1;
    double alpha_slow = 0.001;
    double kld_epsilon = 0.01;
    double kld_z = 0.01;
    float space_partitioning_resolution_xy = 0.5f;
    float space_partitioning_resolution_theta = 0.17453292f;  // pi / 18
  };
  struct LikelihoodMap {
    float z_hit = 0.5f;
    float z_rand = 0.5f;
    int32_t num_beams = 30;
    float max_laser_distance = 14.0f;
    float max_obstacle_distance = 2.0f;
    float sigma_hit = 0.01f;
  };
};

// ---- the seam ---------------------------------------------------------------
// quickmcl::Map (map.h:31-62)
class Map {
public:
  virtual ~Map() = default;
  virtual void set_parameters(const Parameters::LikelihoodMap &parameters) = 0;
  virtual void new_map(const OccupancyGrid::ConstPtr &msg) = 0;
  virtual double update_importance(const LaserPointCloud &cloud,
                                   ParticleCollection *particles) const = 0;
  virtual WeightedParticle::ParticleT generate_random_pose() const = 0;
  virtual OccupancyGrid get_likelihood_as_gridmap() const = 0;
  virtual MapState get_map_state_at(const Vector2f &world_coord) const = 0;
};

// quickmcl::IParticleFilter (i_particle_filter.h:30-85)
class IParticleFilter {
public:
  virtual ~IParticleFilter() = default;
  virtual void initialise(const WeightedParticle::ParticleT &starting_point,
                          const Matrix3f &covariance) = 0;
  virtual void global_localization() = 0;
  virtual void handle_odometry(const Odometry &odom_new, const Odometry &odom_old,
                               float alpha[4]) = 0;
  virtual void update_importance_from_observations(const LaserPointCloud &cloud) = 0;
  virtual void resample_and_cluster() = 0;
  virtual void cluster() = 0;
  virtual void set_parameters(const Parameters::ParticleFilter &parameters) = 0;
  virtual bool get_estimated_pose(Pose2D<double> *pose, Matrix3d *covariance) const = 0;
  virtual const ParticleCollection &get_particles() const = 0;
  virtual size_t num_particles() const = 0;
};

// map_factory.h:29 / particle_filter_factory.h:30-31.  Throw std::runtime_error
// when no CUDA device is present: there is no CPU fallback.
std::shared_ptr<Map> create_likelihood_map();
std::shared_ptr<IParticleFilter> create_particle_filter(const std::shared_ptr<Map> &map);

// ---- concrete classes (named only here and in the factories) ------------------
// quickmcl::MapLikelihood (map_likelihood.h:35-93)
class CudaMapLikelihood : public Map {
public:
  explicit CudaMapLikelihood(int device = -1);
  ~CudaMapLikelihood() override;
  void set_parameters(const Parameters::LikelihoodMap &parameters) override;
  void new_map(const OccupancyGrid::ConstPtr &msg) override;
  double update_importance(const LaserPointCloud &cloud,
                           ParticleCollection *particles) const override;
  WeightedParticle::ParticleT generate_random_pose() const override;
  OccupancyGrid get_likelihood_as_gridmap() const override;
  MapState get_map_state_at(const Vector2f &world_coord) const override;
  // The device-side handle (C ABI); the filter computes on it directly instead
  // of calling update_importance on a host vector (SURVEY 8b "coupling").
  qmcl_map *handle() const { return map_; }

private:
  qmcl_map *map_ = nullptr;
  OccupancyGrid::Info info_;
  mutable uint64_t random_pose_counter_ = 0;
};

// quickmcl::ParticleFilter (particle_filter.h)
class CudaParticleFilter : public IParticleFilter {
public:
  explicit CudaParticleFilter(const std::shared_ptr<Map> &map, uint64_t seed = 0);
  ~CudaParticleFilter() override;
  void initialise(const WeightedParticle::ParticleT &starting_point,
                  const Matrix3f &covariance) override;
  void global_localization() override;
  void handle_odometry(const Odometry &odom_new, const Odometry &odom_old,
                       float alpha[4]) override;
  void update_importance_from_observations(const LaserPointCloud &cloud) override;
  void resample_and_cluster() override;
  void cluster() override;
  void set_parameters(const Parameters::ParticleFilter &parameters) override;
  bool get_estimated_pose(Pose2D<double> *pose, Matrix3d *covariance) const override;
  const ParticleCollection &get_particles() const override;
  size_t num_particles() const override;
  qmcl_filter *handle() const { return filter_; }
  // The cloud in the form Publishing::publish_cloud sends it (publishing.cpp:
  // 45-67, 114-154): 5 floats per particle — x, y, sin(theta/2), cos(theta/2),
  // weight / max_weight — without the 24-byte mirror and the host pass.
  struct Marker {
    float x, y, qz, qw, shade;
  };
  std::vector<Marker> export_markers(double *max_weight = nullptr) const;

private:
  std::shared_ptr<Map> map_;  // keeps the map alive (particle_filter.h:96)
  qmcl_filter *filter_ = nullptr;
  mutable ParticleCollection mirror_;  // lazy host copy for get_particles()
  mutable bool mirror_valid_ = false;
};

// quickmcl::PoseRestorer without the parameter server (pose_restore.cpp:37-100):
// store() yields the six numbers written to ~initial_pose (empty when there is
// no estimate), restore() re-initialises the filter from them with the
// reference's defaults for missing / NaN entries; restore(nullptr) is the "no
// parameter" case (pose 0, identity covariance).
// LaserHandler::Impl::cloud_callback (src/quickmcl_node/laser_handler.cpp:179-288)
// without ROS: the moved-enough gate, the fall-back to global
// localisation when the set is empty, motion -> sensor -> resample-or-cluster
// with the reference's cadence (it resamples when `counter++ % resample_count`
// is NON-zero, :259-260), and the estimate Publishing::publish_estimated_pose
// (publishing.cpp:157-227) turns into the map->odom transform.  Drives any
// IParticleFilter; `calls` records the interface calls made, in order.
class ScanCycle {
public:
  struct MotionModel {  // Parameters::MotionModel (parameters.h:40-50), the node's cfg defaults
    float alpha[4] = {0.05f, 0.1f, 0.02f, 0.05f};
    float min_trans = 0.2f;  // floats like the reference's: doubles are compared with them (:207-209)
    float min_rot = static_cast<float>(3.14159265358979323846 / 6);
  };
  struct Result {
    bool processed = false;            // false: "Skipping laser scan: haven't moved"
    bool resampled = false;
    bool global_localization = false;  // the set was empty and was re-seeded
    bool recompute_transform = false;  // what publish_estimated_pose is told
    bool pose_ok = false;              // get_estimated_pose's return value
    Pose2D<double> pose;
    Matrix3d covariance;
  };
  explicit ScanCycle(const std::shared_ptr<IParticleFilter> &filter, int resample_count = 2)
      : filter_(filter), resample_count_(resample_count) {}
  MotionModel motion;
  std::vector<std::string> calls;
  //! LaserHandler::force_pose_reset: the next scan is processed even without motion.
  void force_pose_reset() { pose_reset_ = true; }
  Result on_scan(const Pose2D<double> &odom_new, const LaserPointCloud &cloud);

private:
  std::shared_ptr<IParticleFilter> filter_;
  int resample_count_;
  bool have_odom_ = false;
  Pose2D<double> odom_at_last_filter_run_;
  bool pose_reset_ = false;
  uint32_t resample_counter_ = 0;
};

class PoseCheckpoint {
public:
  explicit PoseCheckpoint(const std::shared_ptr<IParticleFilter> &filter) : filter_(filter) {}
  std::vector<double> store() const;
  void restore(const std::vector<double> *v) const;

private:
  std::shared_ptr<IParticleFilter> filter_;
};

}  // namespace quickmcl_b200
