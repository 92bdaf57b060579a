// Stand-in for pcl_conversions: fromROSMsg hands over the points the stand-in PointCloud2
// carries (laser.cpp:26-35).  TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <sensor_msgs/PointCloud2.h>

namespace pcl {

template <class PointT> void fromROSMsg(const sensor_msgs::PointCloud2 &msg, PointCloud<PointT> &out) {
  out.points.assign(msg.stand_in_points.begin(), msg.stand_in_points.end());
}

}  // namespace pcl
