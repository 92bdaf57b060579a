// batch.cu — many independent filters driven from native threads
// (BASELINE.json configs[4]: the multi-robot batch).
//
// One robot's update at QuickMCL's default size is a chain of ~40 short
// kernels and a handful of 8-byte read-backs: far too small to fill a B200
// and bounded by host-side launch latency.  The batch call deals the filters
// round-robin to `lanes` host threads; filters of one lane should share a map
// handle (hence a stream), filters of different lanes should not, so that the
// lanes overlap on the device.  No filter reads another's data; there is no
// exchange.  Per robot the sequence is laser_handler.cpp:243-276.

#include "state.cuh"

#include <algorithm>
#include <atomic>
#include <string>
#include <thread>
#include <vector>

using namespace qmcl;

extern "C" {

qmcl_status qmcl_batch_update(qmcl_filter *const *filters, size_t n_filters, int lanes,
                              const double *odom_new, const double *odom_old, const float alpha[4],
                              const float *const *clouds_xy, const size_t *n_points, int resample,
                              double *poses_out, double *covs_out, int *valid_out) {
  if (!filters || !odom_new || !odom_old || !alpha || !clouds_xy || !n_points || !poses_out ||
      !covs_out || !valid_out)
    return QMCL_ERR_INVALID_ARG;
  if (n_filters == 0) return QMCL_OK;
  const int L = std::max(1, std::min<int>(lanes, static_cast<int>(n_filters)));
  std::vector<qmcl_status> status(L, QMCL_OK);
  std::vector<std::string> errors(L);
  auto work = [&](int lane) {
    for (size_t i = lane; i < n_filters; i += L) {
      qmcl_filter *f = filters[i];
      qmcl_status st = qmcl_filter_handle_odometry(f, odom_new + 3 * i, odom_old + 3 * i, alpha);
      if (st == QMCL_OK) st = qmcl_filter_update_sensor(f, clouds_xy[i], n_points[i]);
      if (st == QMCL_OK) st = resample ? qmcl_filter_resample_and_cluster(f) : qmcl_filter_cluster(f);
      if (st == QMCL_OK)
        st = qmcl_filter_get_estimate(f, poses_out + 3 * i, covs_out + 9 * i, valid_out + i);
      if (st != QMCL_OK) {
        status[lane] = st;
        errors[lane] = qmcl_last_error();  // thread-local: carry it to the caller's thread
        return;
      }
    }
  };
  std::vector<std::thread> threads;
  threads.reserve(L - 1);
  for (int lane = 1; lane < L; ++lane) threads.emplace_back(work, lane);
  work(0);
  for (auto &t : threads) t.join();
  for (int lane = 0; lane < L; ++lane)
    if (status[lane] != QMCL_OK) {
      set_error("batch lane %d: %s", lane, errors[lane].c_str());
      return status[lane];
    }
  return QMCL_OK;
}

}  // extern "C"
