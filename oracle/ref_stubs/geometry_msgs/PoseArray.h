// Stand-in for the ROS message geometry_msgs/PoseArray (included by
// map_likelihood.cpp for a debug publisher that is compiled out).
// TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

#include <geometry_msgs/Pose.h>

#include <vector>

namespace geometry_msgs {

struct PoseArray {
  std::vector<Pose> poses;
};

}  // namespace geometry_msgs
