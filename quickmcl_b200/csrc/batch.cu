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

#include "filter.cuh"
#include "front.cuh"
#include "sensor.cuh"
#include "shard.cuh"
#include "tail.cuh"

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
  // Filters that share a map handle share its stream and its mutable sensor
  // state, so they always run on the same lane: lanes are dealt per map.
  std::vector<qmcl_map *> maps;
  std::vector<int> map_of(n_filters);
  for (size_t i = 0; i < n_filters; ++i) {
    if (!filters[i]) return QMCL_ERR_INVALID_ARG;
    size_t k = 0;
    while (k < maps.size() && maps[k] != filters[i]->map) ++k;
    if (k == maps.size()) maps.push_back(filters[i]->map);
    map_of[i] = static_cast<int>(k);
  }
  const int L = std::max(1, std::min<int>(lanes, static_cast<int>(maps.size())));
  std::vector<qmcl_status> status(L, QMCL_OK);
  std::vector<std::string> errors(L);
  auto work = [&](int lane) {
    for (size_t i = 0; i < n_filters; ++i) {
      if (map_of[i] % L != lane) continue;
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


// ------------------------------------------------------------- fused batch
// n independent filters on ONE map handle, updated by three launches in all
// (front, sensor, tail, one per stage with blockIdx.y / blockIdx.x = filter)
// instead of ~28 launches and two host waits per filter.  A filter's tail runs
// in a single block (BlockTeam, tail.cu), which is enough for the default-size
// sets of the multi-robot configuration (BASELINE.json configs[4]).
struct qmcl_batch {
  qmcl_map *map = nullptr;
  std::vector<qmcl_filter *> filters;
  qmcl::DevBuf<qmcl::FrontJob> d_front;
  qmcl::DevBuf<qmcl::SensorLaunch> d_sensor;
  qmcl::DevBuf<qmcl::TailJob> d_tail;
  qmcl::PinnedBuf<qmcl::FrontJob> h_front;
  qmcl::PinnedBuf<qmcl::SensorLaunch> h_sensor;
  qmcl::PinnedBuf<qmcl::TailJob> h_tail;
  qmcl::DevBuf<float> d_clouds;
  qmcl::PinnedBuf<float> h_clouds;
  cudaEvent_t staged = nullptr;  // the last H2D copies out of the pinned staging buffers
  uint64_t fused_cycles = 0, looped_cycles = 0;
};

qmcl_status qmcl_batch_create(qmcl_map *map, size_t n_filters, uint64_t seed, qmcl_batch **out) {
  if (!map || !out || n_filters == 0 || n_filters > 65535) return QMCL_ERR_INVALID_ARG;
  *out = nullptr;
  qmcl_batch *b = new qmcl_batch();
  b->map = map;
  map->refs++;
  for (size_t i = 0; i < n_filters; ++i) {
    qmcl_filter *f = nullptr;
    const qmcl_status st = qmcl_filter_create(map, seed + i, &f);
    if (st != QMCL_OK) {
      qmcl_batch_destroy(b);
      return st;
    }
    b->filters.push_back(f);
  }
  *out = b;
  return QMCL_OK;
}

qmcl_status qmcl_batch_destroy(qmcl_batch *b) {
  if (!b) return QMCL_ERR_INVALID_ARG;
  cudaSetDevice(b->map->device);
  cudaStreamSynchronize(b->map->stream);
  for (qmcl_filter *f : b->filters) qmcl_filter_destroy(f);
  b->d_front.release();
  b->d_sensor.release();
  b->d_tail.release();
  b->h_front.release();
  b->h_sensor.release();
  b->h_tail.release();
  b->d_clouds.release();
  b->h_clouds.release();
  if (b->staged) cudaEventDestroy(b->staged);
  qmcl_map *map = b->map;
  delete b;
  return qmcl_map_destroy(map);
}

qmcl_status qmcl_batch_size(const qmcl_batch *b, size_t *out) {
  if (!b || !out) return QMCL_ERR_INVALID_ARG;
  *out = b->filters.size();
  return QMCL_OK;
}

qmcl_status qmcl_batch_filter(qmcl_batch *b, size_t i, qmcl_filter **out) {
  if (!b || !out || i >= b->filters.size()) return QMCL_ERR_INVALID_ARG;
  *out = b->filters[i];
  return QMCL_OK;
}

qmcl_status qmcl_batch_set_params(qmcl_batch *b, const qmcl_pf_params *p) {
  if (!b || !p) return QMCL_ERR_INVALID_ARG;
  for (qmcl_filter *f : b->filters) QMCL_TRY(qmcl_filter_set_params(f, p));
  return QMCL_OK;
}

qmcl_status qmcl_batch_stats(const qmcl_batch *b, uint64_t *fused_cycles, uint64_t *looped_cycles) {
  if (!b) return QMCL_ERR_INVALID_ARG;
  if (fused_cycles) *fused_cycles = b->fused_cycles;
  if (looped_cycles) *looped_cycles = b->looped_cycles;
  return QMCL_OK;
}

qmcl_status qmcl_batch_cycle(qmcl_batch *b, const double *odom_new, const double *odom_old,
                             const float alpha[4], const float *clouds_xy, const size_t *cloud_offsets,
                             int resample, double *poses_out, double *covs_out, int *valid_out) {
  if (!b || !odom_new || !odom_old || !alpha || !clouds_xy || !cloud_offsets || !poses_out || !covs_out ||
      !valid_out)
    return QMCL_ERR_INVALID_ARG;
  qmcl_map *m = b->map;
  if (!m->has_map) return QMCL_ERR_NO_MAP;
  QMCL_TRY(use_device(m));
  cudaStream_t s = m->stream;
  const size_t n = b->filters.size();
  const size_t total_points = cloud_offsets[n];
  // the scans of all robots: one staged H2D copy (beams of robot i behind its own cloud + 2)
  if (!b->staged) QMCL_CUDA(cudaEventCreateWithFlags(&b->staged, cudaEventDisableTiming));
  else QMCL_CUDA(cudaEventSynchronize(b->staged));  // the previous cycle's copies have left the staging buffers
  QMCL_TRY(b->h_clouds.ensure(2 * total_points + 2));
  QMCL_TRY(b->d_clouds.ensure(2 * total_points + 2));
  std::memcpy(b->h_clouds.ptr, clouds_xy, 2 * total_points * sizeof(float));
  QMCL_CUDA(copy_h2d(b->d_clouds.ptr, b->h_clouds.ptr, 2 * total_points * sizeof(float), s));

  // odometry: the host part of handle_odometry per filter (the motion kernel is deferred)
  for (size_t i = 0; i < n; ++i)
    QMCL_TRY(qmcl_filter_handle_odometry(b->filters[i], odom_new + 3 * i, odom_old + 3 * i, alpha));

  // can every filter take the fused path for this call?
  bool fused = true;
  for (size_t i = 0; i < n && fused; ++i) {
    qmcl_filter *f = b->filters[i];
    const int mode = resample ? f->params.resample_type : qmcl::TAIL_CLUSTER;
    fused = f->map == m && mode >= 0 && mode <= 3 && qmcl::tail_eligible(f, mode) && f->motion_pending &&
            f->cur().n <= 65536 && !f->tail_pending && !f->normalise_pending && !f->total_pending &&
            !f->bins_pending && !f->clusters_pending;
  }
  if (!fused) {
    b->looped_cycles++;
    for (size_t i = 0; i < n; ++i) {
      qmcl_filter *f = b->filters[i];
      const size_t np = cloud_offsets[i + 1] - cloud_offsets[i];
      QMCL_TRY(qmcl_filter_update_sensor_dev(f, b->d_clouds.ptr + 2 * cloud_offsets[i], np));
      QMCL_TRY(resample ? qmcl_filter_resample_and_cluster(f) : qmcl_filter_cluster(f));
    }
    QMCL_CUDA(cudaEventRecord(b->staged, s));
    for (size_t i = 0; i < n; ++i)
      QMCL_TRY(qmcl_filter_get_estimate(b->filters[i], poses_out + 3 * i, covs_out + 9 * i, valid_out + i));
    return QMCL_OK;
  }
  b->fused_cycles++;
  QMCL_TRY(b->h_front.ensure(n));
  QMCL_TRY(b->h_sensor.ensure(n));
  QMCL_TRY(b->h_tail.ensure(n));
  QMCL_TRY(b->d_front.ensure(n));
  QMCL_TRY(b->d_sensor.ensure(n));
  QMCL_TRY(b->d_tail.ensure(n));
  size_t max_particles = 0;
  for (size_t i = 0; i < n; ++i) {
    qmcl_filter *f = b->filters[i];
    const size_t np = cloud_offsets[i + 1] - cloud_offsets[i];
    QMCL_TRY(f->beams.ensure(2 * np + 2));
    int n_used = 0;
    QMCL_TRY(qmcl::front_prepare(f, b->d_clouds.ptr + 2 * cloud_offsets[i], np, &b->h_front.ptr[i], &n_used));
    QMCL_TRY(qmcl::sensor_prepare_deferred(f, n_used, np, &b->h_sensor.ptr[i]));
    max_particles = std::max(max_particles, f->normalise_n);
  }
  QMCL_CUDA(copy_h2d(b->d_front.ptr, b->h_front.ptr, n * sizeof(qmcl::FrontJob), s));
  QMCL_CUDA(copy_h2d(b->d_sensor.ptr, b->h_sensor.ptr, n * sizeof(qmcl::SensorLaunch), s));
  QMCL_TRY(qmcl::launch_front_batch(m, b->d_front.ptr, static_cast<unsigned>(n), max_particles));
  QMCL_TRY(qmcl::launch_sensor_batch(m, b->d_sensor.ptr, static_cast<unsigned>(n), max_particles));
  // the tail jobs are filled while the sensor kernel runs
  for (size_t i = 0; i < n; ++i) {
    qmcl_filter *f = b->filters[i];
    const int mode = resample ? f->params.resample_type : qmcl::TAIL_CLUSTER;
    QMCL_TRY(qmcl::tail_prepare(f, mode, &b->h_tail.ptr[i]));
  }
  QMCL_CUDA(copy_h2d(b->d_tail.ptr, b->h_tail.ptr, n * sizeof(qmcl::TailJob), s));
  QMCL_CUDA(cudaEventRecord(b->staged, s));
  QMCL_TRY(qmcl::launch_tail_batch(m, b->d_tail.ptr, static_cast<unsigned>(n)));
  QMCL_CUDA(cudaStreamSynchronize(s));
  for (size_t i = 0; i < n; ++i) {
    qmcl_filter *f = b->filters[i];
    // folds the result block in (the stream is idle); a filter whose tail asked for the
    // multi-launch path re-runs there
    QMCL_TRY(qmcl::tail_complete(f));
    QMCL_TRY(qmcl_filter_get_estimate(f, poses_out + 3 * i, covs_out + 9 * i, valid_out + i));
  }
  return QMCL_OK;
}

}  // extern "C"
