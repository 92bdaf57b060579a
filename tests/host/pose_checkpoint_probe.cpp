// CPU probe of the C++ host mirror's PoseCheckpoint (quickmcl_b200/host): runs it on a filter
// that only records, prints what it did in hex floats.  tests/test_host_cpp.py compares the
// output with the reference's own PoseRestorer (oracle/_ref).  No device needed.
//
//   pose_checkpoint_probe restore none | restore [v0 v1 ...]   -> "pose a b c" / "cov 9 values"
//   pose_checkpoint_probe store 0 | store 1 x y theta c00 .. c22 (row-major) -> "stored n v..."
#include "quickmcl_b200.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <vector>

namespace qb = quickmcl_b200;

struct Recorder final : qb::IParticleFilter {
  bool has_estimate = false;
  qb::Pose2D<double> estimate;
  qb::Matrix3d estimate_cov{};
  mutable int calls = 0;
  qb::WeightedParticle::ParticleT init_pose;
  qb::Matrix3f init_cov{};
  qb::ParticleCollection none;

  void initialise(const qb::WeightedParticle::ParticleT &p, const qb::Matrix3f &c) override {
    ++calls;
    init_pose = p;
    init_cov = c;
  }
  void global_localization() override {}
  void handle_odometry(const qb::Odometry &, const qb::Odometry &, float[4]) override {}
  void update_importance_from_observations(const qb::LaserPointCloud &) override {}
  void resample_and_cluster() override {}
  void cluster() override {}
  void set_parameters(const qb::Parameters::ParticleFilter &) override {}
  bool get_estimated_pose(qb::Pose2D<double> *pose, qb::Matrix3d *cov) const override {
    if (!has_estimate) return false;
    *pose = estimate;
    *cov = estimate_cov;
    return true;
  }
  const qb::ParticleCollection &get_particles() const override { return none; }
  size_t num_particles() const override { return 0; }
};

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  auto rec = std::make_shared<Recorder>();
  qb::PoseCheckpoint checkpoint(rec);
  if (!std::strcmp(argv[1], "restore")) {
    std::vector<double> v;
    const bool absent = !std::strcmp(argv[2], "none");
    if (!absent)
      for (int i = 2; i < argc; ++i)
        if (std::strcmp(argv[i], "empty")) v.push_back(std::strtod(argv[i], nullptr));
    checkpoint.restore(absent ? nullptr : &v);
    std::printf("calls %d\n", rec->calls);
    std::printf("pose %a %a %a\n", rec->init_pose.x, rec->init_pose.y, rec->init_pose.theta);
    std::printf("cov");
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) std::printf(" %a", rec->init_cov(r, c));
    std::printf("\n");
    return 0;
  }
  if (!std::strcmp(argv[1], "store")) {
    rec->has_estimate = std::atoi(argv[2]) != 0;
    if (rec->has_estimate) {
      if (argc < 15) return 2;
      rec->estimate.x = std::strtod(argv[3], nullptr);
      rec->estimate.y = std::strtod(argv[4], nullptr);
      rec->estimate.theta = std::strtod(argv[5], nullptr);
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) rec->estimate_cov(r, c) = std::strtod(argv[6 + 3 * r + c], nullptr);
    }
    const std::vector<double> stored = checkpoint.store();
    std::printf("stored %zu", stored.size());
    for (double d : stored) std::printf(" %a", d);
    std::printf("\n");
    return 0;
  }
  return 2;
}
