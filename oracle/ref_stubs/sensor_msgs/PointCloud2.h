// Stand-in for the ROS message sensor_msgs/PointCloud2 (only named by a declaration
// in laser.h; laser.cpp is not built).  TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

namespace sensor_msgs {

struct PointCloud2 {};

}  // namespace sensor_msgs
