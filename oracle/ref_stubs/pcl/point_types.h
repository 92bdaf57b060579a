// Stand-in for pcl/point_types.h: the one point type laser.h names.
// TEST INFRASTRUCTURE (oracle/_ref).
#pragma once

namespace pcl {

struct PointXY {
  float x = 0, y = 0;
};

}  // namespace pcl
