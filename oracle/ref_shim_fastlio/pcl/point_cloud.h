// TEST INFRASTRUCTURE ONLY: stand-in for <pcl/point_cloud.h> with the members the reference LIO touches.
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
#include <Eigen/Core>
#include <Eigen/StdVector>
namespace pcl {
struct PCLHeader { uint32_t seq = 0; uint64_t stamp = 0; };
template <typename PointT>
struct PointCloud {
  typedef std::shared_ptr<PointCloud<PointT>> Ptr;
  typedef std::shared_ptr<const PointCloud<PointT>> ConstPtr;
  PCLHeader header;
  std::vector<PointT, Eigen::aligned_allocator<PointT>> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  PointCloud() {}
  PointCloud(uint32_t w, uint32_t h) : points(size_t(w) * h), width(w), height(h) {}
  size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void clear() { points.clear(); width = height = 0; }
  void resize(size_t n) { points.resize(n); width = (uint32_t)n; height = 1; }
  void reserve(size_t n) { points.reserve(n); }
  void push_back(const PointT& p) { points.push_back(p); width = (uint32_t)points.size(); height = 1; }
  auto begin() { return points.begin(); }
  auto end() { return points.end(); }
  auto begin() const { return points.begin(); }
  auto end() const { return points.end(); }
  PointT& operator[](size_t i) { return points[i]; }
  const PointT& operator[](size_t i) const { return points[i]; }
  PointCloud& operator+=(const PointCloud& o) { points.insert(points.end(), o.points.begin(), o.points.end()); width = (uint32_t)points.size(); height = 1; return *this; }
};
}  // namespace pcl
