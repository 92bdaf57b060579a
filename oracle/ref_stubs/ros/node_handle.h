// Stand-in for ros/node_handle.h: the two parameter-server calls pose_restore.cpp makes,
// backed by a process-wide table instead of a ROS master.  TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

#include <ros/console.h>  // the real header brings the logging macros along

#include <map>
#include <memory>
#include <string>
#include <vector>

namespace ros {

inline std::map<std::string, std::vector<double>> &stand_in_parameters() {
  static std::map<std::string, std::vector<double>> table;
  return table;
}

class NodeHandle {
 public:
  bool getParam(const std::string &key, std::vector<double> &out) const {
    const auto it = stand_in_parameters().find(key);
    if (it == stand_in_parameters().end()) return false;
    out = it->second;
    return true;
  }
  void setParam(const std::string &key, const std::vector<double> &value) const { stand_in_parameters()[key] = value; }
};

}  // namespace ros
