// TEST INFRASTRUCTURE ONLY: stand-in for <pcl/point_cloud.h> (see point_types.h).
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
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); }
  void resize(size_t n) { points.resize(n); }
  const PointT& at(size_t i) const { return points.at(i); }
  PointT& at(size_t i) { return points.at(i); }
  void push_back(const PointT& p) { points.push_back(p); }
  auto begin() const { return points.begin(); }
  auto end() const { return points.end(); }
  auto begin() { return points.begin(); }
  auto end() { return points.end(); }
};
template <typename PointT>
void transformPointCloud(const PointCloud<PointT>& in, PointCloud<PointT>& out, const Eigen::Matrix4f& T) {
  out.points.resize(in.points.size());
  for (size_t i = 0; i < in.points.size(); i++) {
    out.points[i] = in.points[i];
    out.points[i].getVector4fMap() = T * in.points[i].getVector4fMap();
  }
}
}  // namespace pcl
