// TEST INFRASTRUCTURE ONLY: stand-in for slam/common/mapping_types.h (which drags in OpenCV, g2o and a lock-free queue).
// The record types the LIO front-end consumes, with the members it reads; written from scratch.
#pragma once
#include <cstdint>
#include <deque>
#include <iostream>
#include <memory>
#include <string>
#include <vector>
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
typedef pcl::PointXYZI Point;
typedef pcl::PointCloud<Point> PointCloud;
struct PointAttr { int id; uint32_t stamp; };  // stamp: microseconds relative to the cloud's header stamp
struct PointCloudAttr {
  PointCloudAttr() { cloud = PointCloud::Ptr(new PointCloud()); }
  PointCloud::Ptr cloud;
  std::vector<PointAttr> attr;
  Eigen::Matrix4d T = Eigen::Matrix4d::Identity();
};
typedef std::shared_ptr<PointCloudAttr> PointCloudAttrPtr;
struct RTKType {
  uint64_t timestamp = 0;
  double heading = 0, pitch = 0, roll = 0, Ve = 0, Vn = 0, Vu = 0;
  std::string sensor, state;
};
struct ImuType {
  EIGEN_MAKE_ALIGNED_OPERATOR_NEW
  double stamp = 0;
  Eigen::Vector3d acc = Eigen::Vector3d::Zero(), gyr = Eigen::Vector3d::Zero();
  Eigen::Quaterniond rot = Eigen::Quaterniond::Identity();
};
