// Stand-in for ros/time.h / ros/duration.h: seconds as a double, and a clock the tests set.
// TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

namespace ros {

inline double &stand_in_clock() {
  static double now = 0;
  return now;
}

class Duration {
 public:
  Duration() {}
  explicit Duration(double seconds) : s_(seconds) {}
  bool isZero() const { return s_ == 0; }
  double toSec() const { return s_; }
  friend bool operator>=(const Duration &a, const Duration &b) { return a.s_ >= b.s_; }

 private:
  double s_ = 0;
};

class Time {
 public:
  Time() {}
  explicit Time(double seconds) : s_(seconds) {}
  static Time now() { return Time(stand_in_clock()); }
  double toSec() const { return s_; }
  friend Duration operator-(const Time &a, const Time &b) { return Duration(a.s_ - b.s_); }

 private:
  double s_ = 0;
};

}  // namespace ros
