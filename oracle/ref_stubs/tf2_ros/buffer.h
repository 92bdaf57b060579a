// Stand-in for tf2_ros/buffer.h: an empty buffer type (the tests' TFReader answers odometry
// queries itself).  Real ROS headers bring boost::optional along; so does this one.
// TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

#include <boost/optional.hpp>

namespace tf2_ros {

class Buffer {};

}  // namespace tf2_ros
