// Stand-in for laser_geometry: laser_handler.cpp's scan subscriber names the projector; the
// tests deliver point clouds, so it projects nothing.  TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

#include <sensor_msgs/LaserScan.h>
#include <sensor_msgs/PointCloud2.h>
#include <tf2_ros/buffer.h>

#include <string>

namespace laser_geometry {

class LaserProjection {
 public:
  void transformLaserScanToPointCloud(const std::string &, const sensor_msgs::LaserScan &, sensor_msgs::PointCloud2 &,
                                      tf2_ros::Buffer &, double, int) {}
};

}  // namespace laser_geometry
