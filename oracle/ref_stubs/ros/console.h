// Stand-in for ros/console.h: the logging macros particle_filter.cpp uses expand to
// nothing (their arguments are not evaluated; none has a side effect).
// TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

#define ROS_DEBUG_NAMED(...) ((void)0)
#define ROS_INFO_NAMED(...) ((void)0)
#define ROS_WARN_NAMED(...) ((void)0)
#define ROS_ERROR_NAMED(...) ((void)0)
#define ROS_FATAL_NAMED(...) ((void)0)
#define ROS_INFO_STREAM_NAMED(name, args) ((void)0)
