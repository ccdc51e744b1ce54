// TEST INFRASTRUCTURE ONLY (oracle/_ref build of the reference's fast_gicp headers): stand-in for
// <pcl/point_types.h>.  PCL 1.9.1 is not installed in this image.  fast_gicp needs a
// 16-byte-aligned point with getVector4fMap() whose 4th component is 1 (PCL's data[3]).
#pragma once
#include <Eigen/Core>
#include <Eigen/Geometry>   // PCL's point_types.h brings Eigen::Isometry3d etc. with it (information_matrix_calculator.hpp relies on that)
#include <memory>
#define PCL_VERSION_CALC(a, b, c) ((a) * 100000 + (b) * 100 + (c))
#define PCL_VERSION PCL_VERSION_CALC(1, 10, 0)  /* selects the pcl::shared_ptr typedef branch */
namespace pcl {
template <typename T> using shared_ptr = std::shared_ptr<T>;
struct alignas(16) PointXYZI {
  float x = 0, y = 0, z = 0, w = 1.f;
  float intensity = 0, _p1 = 0, _p2 = 0, _p3 = 0;
  Eigen::Map<Eigen::Vector3f> getVector3fMap() { return Eigen::Map<Eigen::Vector3f>(&x); }
  Eigen::Map<const Eigen::Vector3f> getVector3fMap() const { return Eigen::Map<const Eigen::Vector3f>(&x); }
  Eigen::Map<Eigen::Vector4f, Eigen::Aligned16> getVector4fMap() { return Eigen::Map<Eigen::Vector4f, Eigen::Aligned16>(&x); }
  Eigen::Map<const Eigen::Vector4f, Eigen::Aligned16> getVector4fMap() const { return Eigen::Map<const Eigen::Vector4f, Eigen::Aligned16>(&x); }
};
}  // namespace pcl
