// TEST INFRASTRUCTURE ONLY: stand-in for <pcl/common/transforms.h> (PCL 1.9.1 is not in this image).  The one overload the
// reference's information_matrix_calculator.cpp uses: transformPointCloud(cloud_in, cloud_out, Eigen::Isometry3f) — PCL 1.9.1
// (common/impl/transforms.hpp) computes, per point, xyz_out = transform * xyz_in in the transform's scalar type and copies
// the other fields.
#pragma once
#include <Eigen/Geometry>
#include <pcl/point_cloud.h>
namespace pcl {
template <typename PointT, typename Scalar, int Mode>
void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Transform<Scalar, 3, Mode>& T) {
  out.points.resize(in.points.size());
  for (size_t i = 0; i < in.points.size(); i++) {
    out.points[i] = in.points[i];
    const Eigen::Matrix<Scalar, 3, 1> p(in.points[i].x, in.points[i].y, in.points[i].z);
    const Eigen::Matrix<Scalar, 3, 1> q = T * p;
    out.points[i].x = static_cast<float>(q[0]); out.points[i].y = static_cast<float>(q[1]); out.points[i].z = static_cast<float>(q[2]);
  }
}
}  // namespace pcl
