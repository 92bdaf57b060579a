// Stand-in for tf2_ros/message_filter.h: keeps the callback laser_handler.cpp registers so
// that a test can deliver a message to it.  TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

#include <ros/node_handle.h>
#include <tf2_ros/buffer.h>

#include <cstdint>
#include <functional>
#include <string>

namespace tf2_ros {

// the callback most recently registered for messages of type M
template <class M> std::function<void(const typename M::ConstPtr &)> &stand_in_callback() {
  static std::function<void(const typename M::ConstPtr &)> fn;
  return fn;
}

template <class M> class MessageFilter {
 public:
  template <class Sub> MessageFilter(Sub &, Buffer &, const std::string &, uint32_t, const ros::NodeHandle &) {}
  template <class C> void registerCallback(void (C::*fn)(const typename M::ConstPtr &), C *object) {
    stand_in_callback<M>() = [fn, object](const typename M::ConstPtr &msg) { (object->*fn)(msg); };
  }
};

}  // namespace tf2_ros
