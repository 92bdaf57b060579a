// Stand-in for the ROS message nav_msgs/OccupancyGrid (ROS is not installed): the
// fields distance_calculator.cpp and map_base.cpp read.  TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

#include <geometry_msgs/Pose.h>

#include <cstdint>
#include <memory>
#include <vector>

namespace nav_msgs {

struct MapMetaData {
  float resolution = 0;
  uint32_t width = 0;
  uint32_t height = 0;
  geometry_msgs::Pose origin;
};

struct OccupancyGrid {
  using ConstPtr = std::shared_ptr<const OccupancyGrid>;  // boost::shared_ptr in ROS
  MapMetaData info;
  std::vector<int8_t> data;  // row-major, -1 unknown, 0..100 occupancy
};

}  // namespace nav_msgs
