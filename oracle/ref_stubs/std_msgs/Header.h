// Stand-in for the ROS message std_msgs/Header.  TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

#include <ros/time.h>

#include <string>

namespace std_msgs {

struct Header {
  ros::Time stamp;
  std::string frame_id;
};

}  // namespace std_msgs
