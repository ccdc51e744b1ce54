// TEST INFRASTRUCTURE ONLY (oracle/_ref/libref_keyframe.so): KeyFrame::downsample (keyframe.cpp:134-141) must link; nothing
// compiled into this library calls it, so the filter copies its input.
#pragma once
#include <pcl/point_cloud.h>
namespace pcl {
template <typename PointT>
class VoxelGrid {
 public:
  void setLeafSize(float, float, float) {}
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr& c) { in_ = c; }
  void filter(PointCloud<PointT>& out) { out = *in_; }
 private:
  typename PointCloud<PointT>::ConstPtr in_;
};
}  // namespace pcl
