// Stand-in for the ROS message geometry_msgs/Pose.  TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

namespace geometry_msgs {

struct Point {
  double x = 0, y = 0, z = 0;
};
struct Quaternion {
  double x = 0, y = 0, z = 0, w = 1;
};
struct Pose {
  Point position;
  Quaternion orientation;
};

}  // namespace geometry_msgs
