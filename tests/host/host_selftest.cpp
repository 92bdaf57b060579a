// host_selftest.cpp — drives the C++ host mirror (quickmcl_b200/host) the way
// the reference's node drives IParticleFilter / Map (laser_handler.cpp:236-276,
// main.cpp:39-74) and checks the results against the CPU oracle.
//
//   host_selftest --expect-no-device   the factories must refuse loudly
//   host_selftest                      full cycle parity on cuda:0
//
// Test infrastructure: this is the only C++ translation unit that links the
// oracle; the host library itself does not.
#include "quickmcl_b200.h"
#include "qmcl.h"
#include "qmcl_oracle.h"

#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>

using namespace quickmcl_b200;

static int failures = 0;
#define EXPECT(cond)                                                             \
  do {                                                                           \
    if (!(cond)) {                                                               \
      std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);       \
      ++failures;                                                                \
    }                                                                            \
  } while (0)

// A 300 x 300 map at 5 cm: unknown border, wall ring, two inner walls.
static OccupancyGrid::ConstPtr make_map() {
  auto g = std::make_shared<OccupancyGrid>();
  const int W = 300;
  g->info.width = g->info.height = W;
  g->info.resolution = 0.05f;
  g->info.origin.position.x = g->info.origin.position.y = -7.5;
  g->data.assign(static_cast<size_t>(W) * W, 0);
  auto at = [&](int x, int y) -> int8_t & { return g->data[static_cast<size_t>(y) * W + x]; };
  for (int y = 0; y < W; ++y)
    for (int x = 0; x < W; ++x) {
      const int edge = std::min(std::min(x, y), std::min(W - 1 - x, W - 1 - y));
      if (edge < 10) at(x, y) = -1;
      else if (edge < 12) at(x, y) = 100;
    }
  for (int y = 12; y < 200; ++y) at(100, y) = at(101, y) = 100;
  for (int x = 150; x < 288; ++x) at(x, 180) = at(x, 181) = 100;
  return g;
}

// Ray-cast a 360-beam scan from `pose` (base frame points), hits only.
static LaserPointCloud make_scan(const OccupancyGrid &g, double px, double py, double pth) {
  LaserPointCloud cloud;
  const int W = static_cast<int>(g.info.width);
  for (int b = 0; b < 360; ++b) {
    const double a = -M_PI + b * (2 * M_PI / 360);
    for (double r = 0.025; r < 14.0; r += 0.025) {
      const double wx = px + r * std::cos(a + pth), wy = py + r * std::sin(a + pth);
      const int ix = static_cast<int>(std::floor((wx - g.info.origin.position.x) / g.info.resolution));
      const int iy = static_cast<int>(std::floor((wy - g.info.origin.position.y) / g.info.resolution));
      if (ix < 0 || iy < 0 || ix >= W || iy >= W) break;
      if (g.data[static_cast<size_t>(iy) * W + ix] > 65) {
        cloud.push_back(PointXY{static_cast<float>(r * std::cos(a)), static_cast<float>(r * std::sin(a))});
        break;
      }
    }
  }
  return cloud;
}

static int expect_no_device() {
  try {
    auto map = create_likelihood_map();
    std::fprintf(stderr, "FAIL: create_likelihood_map() succeeded without a device\n");
    return 1;
  } catch (const std::runtime_error &e) {
    if (std::strstr(e.what(), "no CPU fallback") == nullptr) {
      std::fprintf(stderr, "FAIL: unexpected message: %s\n", e.what());
      return 1;
    }
    std::printf("ok: refused without a CUDA device (%s)\n", e.what());
    return 0;
  }
}

int main(int argc, char **argv) {
  if (argc > 1 && std::string(argv[1]) == "--expect-no-device") return expect_no_device();

  // main.cpp:39-56: map, filter, parameters
  std::shared_ptr<Map> map = create_likelihood_map();
  std::shared_ptr<IParticleFilter> pf = create_particle_filter(map);
  Parameters::LikelihoodMap lp;
  lp.z_hit = 0.9f;
  lp.z_rand = 0.1f;
  lp.sigma_hit = 0.1f;
  lp.num_beams = 30;
  map->set_parameters(lp);
  Parameters::ParticleFilter pp;
  pp.resample_type = ResampleType::low_variance;
  pp.particle_count_min = pp.particle_count_max = 3000;
  pf->set_parameters(pp);
  auto grid = make_map();
  map->new_map(grid);  // main.cpp:68-74

  // the oracle, same inputs
  orc_map *om = orc_map_create();
  orc_likelihood_params olp{lp.z_hit, lp.z_rand, lp.num_beams, lp.max_laser_distance,
                            lp.max_obstacle_distance, lp.sigma_hit};
  orc_map_set_params(om, &olp);
  orc_map_load(om, grid->data.data(), grid->info.width, grid->info.height, grid->info.resolution,
               grid->info.origin.position.x, grid->info.origin.position.y, ORC_FIELD_EXACT_EDT);
  orc_filter *of = orc_filter_create(om, 7);
  orc_pf_params opp{0, 3000, 3000, 2, pp.alpha_fast, pp.alpha_slow, pp.kld_epsilon, pp.kld_z,
                    pp.space_partitioning_resolution_xy, pp.space_partitioning_resolution_theta};
  orc_filter_set_params(of, &opp);
  orc_filter_set_trig_mode(of, ORC_TRIG_CR);

  // Map interface
  EXPECT(map->get_map_state_at(Vector2f{0.f, 0.f}) == MapState::Free);
  EXPECT(map->get_map_state_at(Vector2f{-7.4f, 0.f}) == MapState::Unknown);
  EXPECT(map->get_map_state_at(Vector2f{100.f, 0.f}) == MapState::Unknown);
  EXPECT(static_cast<int>(map->get_map_state_at(Vector2f{-2.475f, -5.f})) ==
         orc_map_state_at(om, -2.475f, -5.f));
  {
    OccupancyGrid lg = map->get_likelihood_as_gridmap();
    std::vector<int8_t> want(lg.data.size());
    orc_map_likelihood_grid(om, want.data());
    EXPECT(lg.data == want);
    const auto rp = map->generate_random_pose();
    EXPECT(map->get_map_state_at(Vector2f{rp.x + 0.01f, rp.y + 0.01f}) == MapState::Free);
  }

  // laser_handler.cpp:236-276, three scans with the same shared RNG stream on both sides
  auto *cf = dynamic_cast<CudaParticleFilter *>(pf.get());
  EXPECT(cf != nullptr);
  qmcl_filter_set_rng(cf->handle(), 7, 0);
  orc_filter_set_rng(of, 7, 0);
  const float pose0[3] = {1.0f, -2.0f, 0.3f};
  const float factor[9] = {0.5f, 0, 0, 0, 0.5f, 0, 0, 0, 0.1f};
  // the covariance factor is passed explicitly: eigenvector signs are not canonical
  qmcl_filter_initialise_factor(cf->handle(), pose0, factor);
  orc_filter_initialise_factor(of, pose0, factor);
  EXPECT(pf->num_particles() == 3000);
  const LaserPointCloud cloud = make_scan(*grid, 1.0, -2.0, 0.3);
  EXPECT(cloud.size() > 300);
  float alpha[4] = {0.05f, 0.1f, 0.02f, 0.05f};
  for (int k = 0; k < 3; ++k) {
    const Odometry o_new(0.02 * (k + 1), 0.01 * (k + 1), 0.01 * (k + 1)), o_old(0.02 * k, 0.01 * k, 0.01 * k);
    pf->handle_odometry(o_new, o_old, alpha);
    pf->update_importance_from_observations(cloud);
    const double nw[3] = {o_new.x, o_new.y, o_new.theta}, od[3] = {o_old.x, o_old.y, o_old.theta};
    orc_filter_handle_odometry(of, nw, od, alpha);
    orc_filter_update_sensor(of, reinterpret_cast<const float *>(cloud.data()), cloud.size());
    if (k == 0) {
      // per-particle weights after the first update (1e-5 relative)
      const ParticleCollection &gp = pf->get_particles();
      std::vector<orc_particle> op(3000);
      orc_filter_copy_particles(of, op.data(), op.size());
      double worst = 0;
      for (size_t i = 0; i < gp.size(); ++i) {
        EXPECT(std::fabs(gp[i].data.x - op[i].x) <= 1e-6f);
        if (op[i].weight > 0) worst = std::max(worst, std::fabs(gp[i].weight - op[i].weight) / op[i].weight);
      }
      EXPECT(worst <= 1e-5);
      std::printf("weights: max relative error %.2e\n", worst);
    }
    pf->resample_and_cluster();
    orc_filter_resample_and_cluster(of);
    Pose2D<double> est;
    Matrix3d cov;
    const bool ok = pf->get_estimated_pose(&est, &cov);
    double opose[3], ocov[9];
    const int ook = orc_filter_get_estimate(of, opose, ocov);
    EXPECT(ok && ook == 1);
    EXPECT(std::fabs(est.x - opose[0]) <= 1e-4 && std::fabs(est.y - opose[1]) <= 1e-4);
    EXPECT(std::fabs(std::remainder(est.theta - opose[2], 2 * M_PI)) <= 1e-4);
    for (int i = 0; i < 9; ++i) EXPECT(std::fabs(cov.m[i] - ocov[i]) <= 1e-6);
    std::printf("scan %d: estimate (%.5f, %.5f, %.5f) vs oracle (%.5f, %.5f, %.5f)\n", k, est.x, est.y,
                est.theta, opose[0], opose[1], opose[2]);
  }
  // Map::update_importance on a host vector (the reference's own coupling)
  {
    ParticleCollection host(500);
    std::vector<orc_particle> oh(500);
    for (int i = 0; i < 500; ++i) {
      host[i].data = WeightedParticle::ParticleT(1.0f + 0.002f * i, -2.0f + 0.001f * i, 0.3f);
      oh[i] = orc_particle{host[i].data.x, host[i].data.y, host[i].data.theta, 0.f, 1.0};
    }
    const double total = map->update_importance(cloud, &host);
    const double ototal = orc_map_update_importance(om, reinterpret_cast<const float *>(cloud.data()),
                                                    cloud.size(), oh.data(), oh.size(), ORC_TRIG_CR);
    EXPECT(std::fabs(total - ototal) <= 1e-6 * ototal);
    for (int i = 0; i < 500; ++i)
      EXPECT(std::fabs(host[i].weight - oh[i].weight) <= 1e-5 * std::fabs(oh[i].weight) + 1e-300);
  }
  // publishing.cpp:45-67,114-154 and pose_restore.cpp:37-100 through the mirror
  {
    double max_w = 0.0;
    auto *cuda_pf = dynamic_cast<CudaParticleFilter *>(pf.get());
    EXPECT(cuda_pf != nullptr);
    const auto markers = cuda_pf->export_markers(&max_w);
    const auto &gp = pf->get_particles();
    EXPECT(markers.size() == gp.size());
    double mw = 0.0;
    for (const auto &p : gp) mw = std::max(mw, p.weight);
    EXPECT(max_w == mw);
    for (size_t i = 0; i < gp.size(); i += 97) {
      EXPECT(markers[i].x == gp[i].data.x && markers[i].y == gp[i].data.y);
      EXPECT(std::fabs(markers[i].qz - std::sin(gp[i].data.theta / 2.0f)) <= 2.5e-7f);
      EXPECT(std::fabs(markers[i].qw - std::cos(gp[i].data.theta / 2.0f)) <= 2.5e-7f);
      EXPECT(markers[i].shade == static_cast<float>(gp[i].weight / mw));
    }
    PoseCheckpoint checkpoint(pf);
    const std::vector<double> stored = checkpoint.store();
    EXPECT(stored.size() == 6);
    std::vector<double> damaged = stored;
    damaged[1] = std::nan("");
    damaged.resize(4);
    checkpoint.restore(&damaged);   // y -> 0, var(y) -> 0.25, var(theta) -> 0.01
    pf->cluster();
    Pose2D<double> est;
    Matrix3d cov;
    EXPECT(pf->get_estimated_pose(&est, &cov));
    EXPECT(std::fabs(est.x - stored[0]) <= 0.1 && std::fabs(est.y) <= 0.1);
    EXPECT(std::fabs(cov(1, 1) - 0.25) <= 0.05 && std::fabs(cov(0, 0) - stored[3]) <= 0.2 * stored[3] + 1e-3);
    std::printf("checkpoint: restored around (%.3f, %.3f, %.3f)\n", est.x, est.y, est.theta);
  }
  {
    // ScanCycle (laser_handler.cpp:179-288): gate, cadence, call order on the CUDA filter
    ScanCycle cycle(pf, 2);
    const auto cloud2 = make_scan(*grid, 1.0, -2.0, 0.3);
    auto r0 = cycle.on_scan(Pose2D<double>{0.0, 0.0, 0.0}, cloud2);       // first odometry: no motion yet
    EXPECT(!r0.processed && r0.recompute_transform && r0.pose_ok);
    auto r1 = cycle.on_scan(Pose2D<double>{0.05, 0.0, 0.0}, cloud2);      // below min_trans: skipped
    EXPECT(!r1.processed && !r1.recompute_transform);
    auto r2 = cycle.on_scan(Pose2D<double>{0.3, 0.0, 0.0}, cloud2);       // moved: counter 0 -> cluster only
    EXPECT(r2.processed && !r2.resampled && r2.pose_ok);
    auto r3 = cycle.on_scan(Pose2D<double>{0.6, 0.05, 0.0}, cloud2);      // counter 1 -> resample
    EXPECT(r3.processed && r3.resampled && r3.recompute_transform && r3.pose_ok);
    auto r4 = cycle.on_scan(Pose2D<double>{0.6, 0.05, 0.7}, cloud2);      // rotation alone passes the gate
    EXPECT(r4.processed && !r4.resampled);
    cycle.force_pose_reset();
    auto r5 = cycle.on_scan(Pose2D<double>{0.6, 0.05, 0.7}, cloud2);      // forced although nothing moved
    EXPECT(r5.processed && r5.resampled && r5.recompute_transform);
    const std::vector<std::string> want = {
        "cluster", "get_estimated_pose", "cluster", "get_estimated_pose",
        "handle_odometry", "update_importance_from_observations", "cluster", "get_estimated_pose",
        "handle_odometry", "update_importance_from_observations", "resample_and_cluster", "get_estimated_pose",
        "handle_odometry", "update_importance_from_observations", "cluster", "get_estimated_pose",
        "handle_odometry", "update_importance_from_observations", "resample_and_cluster", "get_estimated_pose"};
    EXPECT(cycle.calls == want);
    EXPECT(std::isfinite(r5.pose.x) && std::isfinite(r5.covariance(0, 0)));
    std::printf("scan cycle: %zu interface calls, last estimate (%.3f, %.3f, %.3f)\n", cycle.calls.size(),
                r5.pose.x, r5.pose.y, r5.pose.theta);
  }
  EXPECT(qmcl_kernel_launch_count() > 20);
  orc_filter_destroy(of);
  orc_map_destroy(om);
  if (failures) {
    std::fprintf(stderr, "%d check(s) failed\n", failures);
    return 1;
  }
  std::printf("host_selftest ok (%llu kernel launches)\n",
              static_cast<unsigned long long>(qmcl_kernel_launch_count()));
  return 0;
}
