// Stand-in for the ROS message sensor_msgs/LaserScan (named by laser_handler.cpp's scan
// subscriber, which the tests do not use).  TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

#include <std_msgs/Header.h>

#include <memory>
#include <vector>

namespace sensor_msgs {

struct LaserScan {
  using ConstPtr = std::shared_ptr<const LaserScan>;
  std_msgs::Header header;
  float angle_min = 0, angle_max = 0, angle_increment = 0, range_min = 0, range_max = 0;
  std::vector<float> ranges;
};

}  // namespace sensor_msgs
