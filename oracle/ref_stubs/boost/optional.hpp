// Stand-in for boost/optional.hpp: the handful of members laser_handler.cpp uses on its
// "odometry at the last filter run".  TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

namespace boost {

struct none_t {};
static const none_t none{};

template <class T> class optional {
 public:
  optional() {}
  optional(none_t) {}
  optional(const T &v) : has_(true), value_(v) {}
  optional &operator=(const T &v) {
    value_ = v;
    has_ = true;
    return *this;
  }
  bool operator!() const { return !has_; }
  explicit operator bool() const { return has_; }
  T &get() { return value_; }
  const T &get() const { return value_; }

 private:
  bool has_ = false;
  T value_;
};

}  // namespace boost
