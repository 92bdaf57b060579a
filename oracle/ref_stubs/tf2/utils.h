// Stand-in for tf2/utils.h: getYaw of a quaternion (only named by a Pose2D
// constructor that oracle/_ref never calls).  TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

#include <geometry_msgs/Pose.h>

#include <cmath>

namespace tf2 {

inline double getYaw(const geometry_msgs::Quaternion &q) {
  return std::atan2(2.0 * (q.w * q.z + q.x * q.y), 1.0 - 2.0 * (q.y * q.y + q.z * q.z));
}

}  // namespace tf2
