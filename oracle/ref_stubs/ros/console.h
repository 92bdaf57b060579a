// Stand-in for ros/console.h: the *_NAMED logging macros (particle_filter.cpp, timer.cpp)
// expand to nothing (their arguments are not evaluated; none has a side effect); the two
// unnamed warning macros of pose_restore.cpp keep their text so that a test can read it.
// TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

#include <sstream>
#include <string>
#include <vector>

namespace ros {
// what ROS_WARN / ROS_WARN_STREAM would have printed (pose_restore.cpp); read by ref_shim.cpp
inline std::vector<std::string> &stand_in_warnings() {
  static std::vector<std::string> lines;
  return lines;
}
}  // namespace ros

#define ROS_WARN(text) ros::stand_in_warnings().push_back(text)
#define ROS_WARN_STREAM(args)                      \
  do {                                             \
    std::ostringstream qmcl_stand_in_stream;       \
    qmcl_stand_in_stream << args;                  \
    ros::stand_in_warnings().push_back(qmcl_stand_in_stream.str()); \
  } while (0)
#define ROS_ERROR(text) ros::stand_in_warnings().push_back(std::string("ERROR: ") + text)
#define ROS_DEBUG(...) ((void)0)
#define ROS_DEBUG_STREAM_NAMED(name, args) ((void)0)
#define ROS_DEBUG_NAMED(...) ((void)0)
#define ROS_INFO_NAMED(...) ((void)0)
#define ROS_WARN_NAMED(...) ((void)0)
#define ROS_ERROR_NAMED(...) ((void)0)
#define ROS_FATAL_NAMED(...) ((void)0)
#define ROS_INFO_STREAM_NAMED(name, args) ((void)0)
