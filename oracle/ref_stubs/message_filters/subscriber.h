// Stand-in for message_filters/subscriber.h: a subscriber that subscribes to nothing.
// TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

#include <ros/node_handle.h>

#include <cstdint>
#include <string>

namespace message_filters {

template <class M> class Subscriber {
 public:
  Subscriber(const ros::NodeHandle &, const std::string &, uint32_t) {}
};

}  // namespace message_filters
