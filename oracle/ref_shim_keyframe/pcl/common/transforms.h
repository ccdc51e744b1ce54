// TEST INFRASTRUCTURE ONLY (oracle/_ref/libref_keyframe.so): pcl::transformPointCloud for the one call of keyframe.cpp:144.
#pragma once
#include <Eigen/Geometry>
#include <pcl/point_cloud.h>
namespace pcl {
template <typename PointT, typename M>
void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const M& T) {
  out = in;
  const Eigen::Matrix4f Tf = T.template cast<float>();
  for (auto& p : out.points) {
    const Eigen::Vector3f q = Tf.template block<3, 3>(0, 0) * Eigen::Vector3f(p.x, p.y, p.z) + Tf.template block<3, 1>(0, 3);
    p.x = q[0]; p.y = q[1]; p.z = q[2];
  }
}
}  // namespace pcl
