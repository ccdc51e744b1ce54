// TEST INFRASTRUCTURE ONLY (oracle/_ref build of the whole reference LIO, oracle/ref_fastlio.cpp): minimal stand-in for
// <pcl/point_types.h>.  PCL 1.9.1 is not installed; the reference LIO needs POD points with
// x/y/z/intensity/normal/curvature members and getVector3fMap().  Written from scratch.
#pragma once
#include <Eigen/Core>
namespace pcl {
struct PointXYZ {
  float x = 0, y = 0, z = 0, _pad = 1.f;
  Eigen::Map<Eigen::Vector3f> getVector3fMap() { return Eigen::Map<Eigen::Vector3f>(&x); }
  Eigen::Map<const Eigen::Vector3f> getVector3fMap() const { return Eigen::Map<const Eigen::Vector3f>(&x); }
};
struct PointXYZI {
  float x = 0, y = 0, z = 0, _pad = 1.f;
  float intensity = 0, _p1 = 0, _p2 = 0, _p3 = 0;
  Eigen::Map<Eigen::Vector3f> getVector3fMap() { return Eigen::Map<Eigen::Vector3f>(&x); }
  Eigen::Map<const Eigen::Vector3f> getVector3fMap() const { return Eigen::Map<const Eigen::Vector3f>(&x); }
};
struct PointXYZINormal {
  float x = 0, y = 0, z = 0, _pad = 1.f;
  float normal_x = 0, normal_y = 0, normal_z = 0, _padn = 0;
  float intensity = 0, curvature = 0, _p2 = 0, _p3 = 0;
  Eigen::Map<Eigen::Vector3f> getVector3fMap() { return Eigen::Map<Eigen::Vector3f>(&x); }
  Eigen::Map<const Eigen::Vector3f> getVector3fMap() const { return Eigen::Map<const Eigen::Vector3f>(&x); }
};
}  // namespace pcl
