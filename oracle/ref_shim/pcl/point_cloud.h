// TEST INFRASTRUCTURE ONLY (oracle/_ref build): stand-in for <pcl/point_cloud.h>.
#pragma once
#include <memory>
#include <vector>
#include <Eigen/Core>
#include <Eigen/StdVector>
namespace pcl {
template <typename PointT>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  std::vector<PointT, Eigen::aligned_allocator<PointT>> points;
  PointCloud() {}
  PointCloud(unsigned w, unsigned h) : points(size_t(w) * h) {}
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); }
  void resize(size_t n) { points.resize(n); }
};
}  // namespace pcl
