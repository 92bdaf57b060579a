/*
 * ref_shim_node.cpp — the REFERENCE'S OWN scan callback (LaserHandler::Impl::cloud_callback,
 * src/quickmcl_node/laser_handler.cpp:179-288) behind C entry points (oracle/_ref).
 *
 * TEST INFRASTRUCTURE ONLY.  laser_handler.cpp and laser.cpp are compiled from where they
 * lie; ROS, tf2, message_filters and laser_geometry are stand-in headers (oracle/ref_stubs).
 * The callback's collaborators are pimpl classes of the node whose headers are plain C++;
 * this file implements them as recorders:
 *   quickmcl_node::TFReader    answers get_odometry_pose with the odometry the test supplies;
 *   quickmcl_node::Publishing  writes "publish_cloud" / "publish_estimated_pose <recompute>"
 *                              into the event log;
 *   the IParticleFilter        writes every call with its arguments into the same log.
 * tests/test_oracle_vs_ref_filter.py replays the same scans through
 * quickmcl_b200.cycle.ScanCycle and compares the logs.
 */
#include "quickmcl/i_particle_filter.h"
#include "quickmcl/parameters.h"
#include "quickmcl/pose_restore.h"
#include "quickmcl_node/laser_handler.h"
#include "quickmcl_node/publishing.h"
#include "quickmcl_node/tf_reader.h"

#include <ros/console.h>
#include <ros/time.h>
#include <sensor_msgs/PointCloud2.h>
#include <tf2_ros/message_filter.h>

#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace {

std::vector<std::string> g_events;
quickmcl::Odometry g_odometry(0, 0, 0);
bool g_odometry_ok = true;
tf2_ros::Buffer g_buffer;

std::string hex(double v) {
  char buf[64];
  std::snprintf(buf, sizeof buf, " %a", v);
  return buf;
}

struct LoggingFilter final : quickmcl::IParticleFilter {
  quickmcl::ParticleCollection particles;  // empty, or one particle: only emptiness is asked
  void initialise(const quickmcl::WeightedParticle::ParticleT &,
                  const quickmcl::WeightedParticle::ParticleT::EigenMatrix &) override {
    g_events.push_back("initialise");
  }
  void global_localization() override {
    g_events.push_back("global_localization");
    particles.assign(1, quickmcl::WeightedParticle{});
  }
  void handle_odometry(const quickmcl::Odometry &n, const quickmcl::Odometry &o, float alpha[4]) override {
    std::string e = "handle_odometry";
    e += hex(n.x) + hex(n.y) + hex(n.theta) + hex(o.x) + hex(o.y) + hex(o.theta);
    for (int i = 0; i < 4; ++i) e += hex(alpha[i]);
    g_events.push_back(e);
  }
  void update_importance_from_observations(const quickmcl::LaserPointCloud &cloud) override {
    g_events.push_back("update_importance_from_observations " + std::to_string(cloud.size()));
  }
  void resample_and_cluster() override { g_events.push_back("resample_and_cluster"); }
  void cluster() override { g_events.push_back("cluster"); }
  void set_parameters(const quickmcl::Parameters::ParticleFilter &) override {}
  bool get_estimated_pose(quickmcl::Pose2D<double> *, Eigen::Matrix3d *) const override {
    g_events.push_back("get_estimated_pose");
    return false;
  }
  const quickmcl::ParticleCollection &get_particles() const override { return particles; }
  size_t num_particles() const override { return particles.size(); }
};

}  // namespace

// ---- the node's pimpl collaborators, implemented as recorders
namespace quickmcl_node {

class TFReader::Impl {};
TFReader::TFReader(const std::shared_ptr<quickmcl::Parameters> &) {}
TFReader::~TFReader() {}
bool TFReader::get_odometry_transform(const ros::Time &, Eigen::Affine3d *) { return false; }
bool TFReader::get_odometry_pose(const ros::Time &, quickmcl::Odometry *odom) {
  if (g_odometry_ok) *odom = g_odometry;
  return g_odometry_ok;
}
tf2_ros::Buffer *TFReader::get_buffer() { return &g_buffer; }

class Publishing::Impl {};
Publishing::Publishing(const std::shared_ptr<quickmcl::Parameters> &, const std::shared_ptr<quickmcl::Map> &,
                       const std::shared_ptr<quickmcl::IParticleFilter> &, const std::shared_ptr<TFReader> &) {}
Publishing::~Publishing() {}
void Publishing::setup() {}
void Publishing::publish_map() { g_events.push_back("publish_map"); }
void Publishing::publish_cloud() { g_events.push_back("publish_cloud"); }
void Publishing::publish_estimated_pose(const ros::Time &, bool recompute_transform) {
  g_events.push_back(recompute_transform ? "publish_estimated_pose recompute" : "publish_estimated_pose");
}

}  // namespace quickmcl_node

struct ref_cycle {
  std::shared_ptr<quickmcl::Parameters> parameters;
  std::shared_ptr<LoggingFilter> filter;
  std::shared_ptr<quickmcl_node::TFReader> tf_reader;
  std::shared_ptr<quickmcl_node::Publishing> publishing;
  std::shared_ptr<quickmcl::PoseRestorer> pose_restorer;
  std::unique_ptr<quickmcl_node::LaserHandler> handler;
};

extern "C" {

// a LaserHandler wired like main.cpp:76-93 does it, on the point-cloud subscriber
ref_cycle *ref_cycle_create(int resample_count, float min_trans, float min_rot, const float alpha[4],
                            double save_pose_period) {
  auto *c = new ref_cycle;
  c->parameters = std::make_shared<quickmcl::Parameters>();
  c->parameters->particle_filter.resample_count = resample_count;
  c->parameters->motion_model.min_trans = min_trans;
  c->parameters->motion_model.min_rot = min_rot;
  for (int i = 0; i < 4; ++i) c->parameters->motion_model.alpha[i] = alpha[i];
  c->parameters->ros.internal_laser_processing = false;
  c->parameters->ros.save_pose_period = save_pose_period;
  c->filter = std::make_shared<LoggingFilter>();
  c->tf_reader = std::make_shared<quickmcl_node::TFReader>(c->parameters);
  c->publishing = std::make_shared<quickmcl_node::Publishing>(c->parameters, nullptr, c->filter, c->tf_reader);
  c->pose_restorer = std::make_shared<quickmcl::PoseRestorer>(ros::NodeHandle(), c->filter);
  c->handler.reset(new quickmcl_node::LaserHandler(c->parameters, c->filter, c->tf_reader, c->publishing,
                                                   c->pose_restorer));
  c->handler->setup();
  return c;
}
void ref_cycle_destroy(ref_cycle *c) { delete c; }
void ref_cycle_force_pose_reset(ref_cycle *c) { c->handler->force_pose_reset(); }
void ref_cycle_set_particles_present(ref_cycle *c, int present) {
  c->filter->particles.assign(present ? 1 : 0, quickmcl::WeightedParticle{});
}

// Delivers one point cloud of n_points points to the callback the handler registered, at
// time `now`, with the odometry the TF reader will answer (odom_ok = 0: the lookup fails).
// events receives the log of this call, one event per line; returns its length.
size_t ref_cycle_on_scan(ref_cycle *c, const double odom[3], int odom_ok, size_t n_points, double now, char *events,
                         size_t cap) {
  (void)c;
  g_events.clear();
  ros::stand_in_warnings().clear();
  g_odometry = quickmcl::Odometry(odom[0], odom[1], odom[2]);
  g_odometry_ok = odom_ok != 0;
  ros::stand_in_clock() = now;
  auto msg = std::make_shared<sensor_msgs::PointCloud2>();
  msg->header.stamp = ros::Time(now);
  msg->stand_in_points.resize(n_points);
  tf2_ros::stand_in_callback<sensor_msgs::PointCloud2>()(msg);
  std::string all;
  for (const auto &w : ros::stand_in_warnings()) all += "log: " + w + "\n";
  for (const auto &e : g_events) all += e + "\n";
  ros::stand_in_warnings().clear();
  if (events && cap) {
    const size_t n = all.size() < cap - 1 ? all.size() : cap - 1;
    std::memcpy(events, all.data(), n);
    events[n] = 0;
  }
  return all.size();
}

}  // extern "C"
