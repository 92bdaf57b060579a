// CPU probe of the C++ host mirror's ScanCycle (quickmcl_b200/host): replays a scenario read
// from stdin on a filter that only logs, prints the log.  tests/test_host_cpp.py compares it
// with the reference's own LaserHandler::Impl::cloud_callback (oracle/_ref).  No device needed.
//
//   argv: resample_count
//   stdin, one scan per line: x y theta n_points lose_particles(0/1) pose_reset(0/1)
//   stdout per scan: the filter calls (handle_odometry with its ten arguments as hex floats),
//   then "result processed resampled global recompute"
#include "quickmcl_b200.h"

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <string>
#include <vector>

namespace qb = quickmcl_b200;

struct Logger final : qb::IParticleFilter {
  qb::ParticleCollection particles;
  void initialise(const qb::WeightedParticle::ParticleT &, const qb::Matrix3f &) override {}
  void global_localization() override {
    std::printf("global_localization\n");
    particles.assign(1, qb::WeightedParticle{});
  }
  void handle_odometry(const qb::Odometry &n, const qb::Odometry &o, float alpha[4]) override {
    std::printf("handle_odometry %a %a %a %a %a %a %a %a %a %a\n", n.x, n.y, n.theta, o.x, o.y, o.theta, alpha[0],
                alpha[1], alpha[2], alpha[3]);
  }
  void update_importance_from_observations(const qb::LaserPointCloud &cloud) override {
    std::printf("update_importance_from_observations %zu\n", cloud.size());
  }
  void resample_and_cluster() override { std::printf("resample_and_cluster\n"); }
  void cluster() override { std::printf("cluster\n"); }
  void set_parameters(const qb::Parameters::ParticleFilter &) override {}
  bool get_estimated_pose(qb::Pose2D<double> *, qb::Matrix3d *) const override { return false; }
  const qb::ParticleCollection &get_particles() const override { return particles; }
  size_t num_particles() const override { return particles.size(); }
};

int main(int argc, char **argv) {
  if (argc < 2) return 2;
  auto filter = std::make_shared<Logger>();
  qb::ScanCycle cycle(filter, std::atoi(argv[1]));
  double x, y, theta;
  unsigned long n_points;
  int lose, reset;
  while (std::scanf("%la %la %la %lu %d %d", &x, &y, &theta, &n_points, &lose, &reset) == 6) {
    if (reset) cycle.force_pose_reset();
    if (lose) filter->particles.clear();
    qb::LaserPointCloud cloud;
    cloud.resize(n_points);
    qb::Pose2D<double> odom;
    odom.x = x;
    odom.y = y;
    odom.theta = theta;
    const qb::ScanCycle::Result r = cycle.on_scan(odom, cloud);
    std::printf("result %d %d %d %d\n", r.processed, r.resampled, r.global_localization, r.recompute_transform);
  }
  return 0;
}
