// Stand-in for the ROS message sensor_msgs/PointCloud2: a header and, instead of the binary
// blob with field descriptions, the x / y points themselves (pcl_conversions' stand-in copies
// them out, which is all laser.cpp asks of PCL).  TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

#include <pcl/point_types.h>
#include <std_msgs/Header.h>

#include <memory>
#include <vector>

namespace sensor_msgs {

struct PointCloud2 {
  using ConstPtr = std::shared_ptr<const PointCloud2>;
  std_msgs::Header header;
  std::vector<pcl::PointXY> stand_in_points;
};

}  // namespace sensor_msgs
