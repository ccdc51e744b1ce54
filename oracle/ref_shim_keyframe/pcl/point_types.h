// TEST INFRASTRUCTURE ONLY (oracle/_ref/libref_keyframe.so): the point types of oracle/ref_shim_fastlio plus the PointXYZRGB
// that slam/common/mapping_types.h names in a typedef.  Written from scratch.
#pragma once
#include "../../ref_shim_fastlio/pcl/point_types.h"
namespace pcl {
struct PointXYZRGB {
  float x = 0, y = 0, z = 0, _pad = 1.f;
  union { struct { unsigned char b, g, r, a; }; float rgb; };
  float _p1 = 0, _p2 = 0, _p3 = 0;
  PointXYZRGB() : rgb(0) {}
};
}  // namespace pcl
