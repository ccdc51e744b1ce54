// TEST INFRASTRUCTURE ONLY: stand-in for pcl::VoxelGrid (PCL 1.9.1 is external), restating its published
// algorithm (filters/impl/voxel_grid.hpp) for the one use the reference LIO makes of it (laserMapping.cpp:1206-1207):
// bounding box -> integer leaf coordinates floor(p * inv_leaf) - min_b -> sort by linear index -> one output point per
// leaf = the centroid of ALL fields (downsample_all_data_ = true), in ascending index order; if the grid has more than
// INT32_MAX leaves the input is returned unchanged.  This is the ONE stage of the compiled reference LIO that is ours.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <vector>
#include <pcl/point_cloud.h>
namespace pcl {
template <typename PointT>
class VoxelGrid {
 public:
  void setLeafSize(float lx, float ly, float lz) { leaf_[0] = lx; leaf_[1] = ly; leaf_[2] = lz; }
  void setInputCloud(const typename PointCloud<PointT>::ConstPtr& c) { input_ = c; }
  void filter(PointCloud<PointT>& out) {
    const auto& in = input_->points;
    PointCloud<PointT> res;
    res.header = input_->header;
    if (in.empty()) { out = res; return; }
    float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
    float mx[3] = {-mn[0], -mn[0], -mn[0]};
    for (const auto& p : in) {
      const float v[3] = {p.x, p.y, p.z};
      for (int a = 0; a < 3; a++) { mn[a] = std::min(mn[a], v[a]); mx[a] = std::max(mx[a], v[a]); }
    }
    const float inv[3] = {1.0f / leaf_[0], 1.0f / leaf_[1], 1.0f / leaf_[2]};
    long long minb[3], div[3];
    for (int a = 0; a < 3; a++) {
      minb[a] = (long long)std::floor(mn[a] * inv[a]);
      div[a] = (long long)std::floor(mx[a] * inv[a]) - minb[a] + 1;
    }
    if (div[0] * div[1] * div[2] > (long long)std::numeric_limits<int32_t>::max()) { out = *input_; return; }
    struct Key { int idx; unsigned pt; };
    std::vector<Key> keys(in.size());
    for (size_t i = 0; i < in.size(); i++) {
      const int i0 = (int)(std::floor(in[i].x * inv[0]) - (float)minb[0]);
      const int i1 = (int)(std::floor(in[i].y * inv[1]) - (float)minb[1]);
      const int i2 = (int)(std::floor(in[i].z * inv[2]) - (float)minb[2]);
      keys[i] = {i0 + i1 * (int)div[0] + i2 * (int)(div[0] * div[1]), (unsigned)i};
    }
    // input order inside a leaf (PCL's std::sort leaves that order unspecified; fixing it keeps the fp32 sums reproducible)
    std::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) { return a.idx != b.idx ? a.idx < b.idx : a.pt < b.pt; });
    for (size_t s = 0; s < keys.size();) {
      size_t e = s;
      PointT c = PointT();
      float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      while (e < keys.size() && keys[e].idx == keys[s].idx) {
        const PointT& p = in[keys[e].pt];
        acc[0] += p.x; acc[1] += p.y; acc[2] += p.z; acc[3] += p.intensity;
        acc[4] += p.normal_x; acc[5] += p.normal_y; acc[6] += p.normal_z; acc[7] += p.curvature;
        e++;
      }
      const float n = (float)(e - s);
      c.x = acc[0] / n; c.y = acc[1] / n; c.z = acc[2] / n; c.intensity = acc[3] / n;
      c.normal_x = acc[4] / n; c.normal_y = acc[5] / n; c.normal_z = acc[6] / n; c.curvature = acc[7] / n;
      res.points.push_back(c);
      s = e;
    }
    res.width = (uint32_t)res.points.size(); res.height = 1;
    out = res;
  }
 private:
  float leaf_[3] = {1.f, 1.f, 1.f};
  typename PointCloud<PointT>::ConstPtr input_;
};
}  // namespace pcl
