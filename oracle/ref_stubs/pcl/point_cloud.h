// Stand-in for pcl/point_cloud.h: a point cloud as the sensor model reads it
// (size(), at(i); map_likelihood.cpp:66-94).  TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

#include <cstddef>
#include <vector>

namespace pcl {

template <class PointT> class PointCloud {
 public:
  std::vector<PointT> points;
  std::size_t size() const { return points.size(); }
  const PointT &at(std::size_t i) const { return points.at(i); }
  PointT &at(std::size_t i) { return points.at(i); }
  void push_back(const PointT &p) { points.push_back(p); }
  void clear() { points.clear(); }
};

}  // namespace pcl
