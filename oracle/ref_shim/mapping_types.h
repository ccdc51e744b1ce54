// TEST INFRASTRUCTURE ONLY (oracle/_ref build): stand-in for slam/common/mapping_types.h, which
// drags in PCL/OpenCV.  common_lib.h only needs the two sensor record types to exist.
#pragma once
#include <deque>
#include <vector>
#include <string>
#include <iostream>
#include <Eigen/Core>
#include <Eigen/Geometry>
struct RTKType { double timestamp = 0; double Ve = 0, Vn = 0, Vu = 0; };
struct ImuType {
  double stamp = 0;
  Eigen::Vector3d acc = Eigen::Vector3d::Zero(), gyr = Eigen::Vector3d::Zero();
  Eigen::Quaterniond rot = Eigen::Quaterniond::Identity();
};
