// Stand-in for ros/ros.h: map_likelihood.h includes it for a debug publisher that
// config.h compiles out.  TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

#include <ros/console.h>

namespace ros {
// test_main.cpp of the reference's gtest suite calls this; there is no ROS master to talk to
inline void init(int &, char **, const char *) {}
}  // namespace ros
