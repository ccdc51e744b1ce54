// TEST INFRASTRUCTURE ONLY: stand-in for pcl::search::Search / pcl::search::KdTree.  The reference
// delegates neighbour search to PCL/FLANN (absent here); any EXACT k-NN gives the same neighbours up
// to distance ties, so this shim is a plain median-split k-d tree written for the test build.  Ties
// are broken by the smaller point index.
#pragma once
#include <algorithm>
#include <memory>
#include <numeric>
#include <vector>
#include <pcl/point_cloud.h>
namespace pcl { namespace search {
template <typename PointT>
class Search {
 public:
  typedef typename PointCloud<PointT>::ConstPtr PointCloudConstPtr;
  virtual ~Search() {}
  virtual void setInputCloud(const PointCloudConstPtr& cloud) = 0;
  PointCloudConstPtr getInputCloud() const { return cloud_; }
  virtual int nearestKSearch(const PointT& q, int k, std::vector<int>& idx, std::vector<float>& sqd) const = 0;
 protected:
  PointCloudConstPtr cloud_;
};
template <typename PointT>
class KdTree : public Search<PointT> {
 public:
  typedef typename Search<PointT>::PointCloudConstPtr PointCloudConstPtr;
  typedef std::shared_ptr<KdTree<PointT>> Ptr;
  typedef std::shared_ptr<const KdTree<PointT>> ConstPtr;
  void setInputCloud(const PointCloudConstPtr& cloud) override {
    this->cloud_ = cloud;
    const int n = (int)cloud->size();
    order_.resize(n);
    std::iota(order_.begin(), order_.end(), 0);
    nodes_.clear();
    nodes_.reserve(2 * n / kLeaf + 4);
    if (n > 0) build(0, n);
  }
  int nearestKSearch(const PointT& q, int k, std::vector<int>& idx, std::vector<float>& sqd) const override {
    std::vector<std::pair<float, int>> heap;  // max-heap on (d2, index)
    heap.reserve(k + 1);
    if (!nodes_.empty()) search(0, q, k, heap);
    std::sort_heap(heap.begin(), heap.end());
    idx.resize(heap.size()); sqd.resize(heap.size());
    for (size_t i = 0; i < heap.size(); i++) { idx[i] = heap[i].second; sqd[i] = heap[i].first; }
    return (int)heap.size();
  }
 private:
  static constexpr int kLeaf = 8;
  struct Node { int lo, hi, axis, left, right; float split; };
  float coord(int i, int a) const { const PointT& p = this->cloud_->points[i]; return a == 0 ? p.x : a == 1 ? p.y : p.z; }
  int build(int lo, int hi) {
    const int id = (int)nodes_.size();
    nodes_.push_back(Node{lo, hi, -1, -1, -1, 0.f});
    if (hi - lo <= kLeaf) return id;
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = lo; i < hi; i++) for (int a = 0; a < 3; a++) { float c = coord(order_[i], a); mn[a] = std::min(mn[a], c); mx[a] = std::max(mx[a], c); }
    int ax = 0;
    for (int a = 1; a < 3; a++) if (mx[a] - mn[a] > mx[ax] - mn[ax]) ax = a;
    if (!(mx[ax] > mn[ax])) return id;  // all points identical: keep as a leaf
    const int mid = (lo + hi) / 2;
    std::nth_element(order_.begin() + lo, order_.begin() + mid, order_.begin() + hi, [&](int a, int b) { return coord(a, ax) < coord(b, ax); });
    const float split = coord(order_[mid], ax);
    const int l = build(lo, mid), r = build(mid, hi);
    nodes_[id].axis = ax; nodes_[id].split = split; nodes_[id].left = l; nodes_[id].right = r;
    return id;
  }
  void search(int id, const PointT& q, int k, std::vector<std::pair<float, int>>& heap) const {
    const Node& nd = nodes_[id];
    if (nd.axis < 0) {
      for (int i = nd.lo; i < nd.hi; i++) {
        const PointT& p = this->cloud_->points[order_[i]];
        const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
        const std::pair<float, int> c(dx * dx + dy * dy + dz * dz, order_[i]);
        if ((int)heap.size() < k) { heap.push_back(c); std::push_heap(heap.begin(), heap.end()); }
        else if (c < heap.front()) { std::pop_heap(heap.begin(), heap.end()); heap.back() = c; std::push_heap(heap.begin(), heap.end()); }
      }
      return;
    }
    const float qc = nd.axis == 0 ? q.x : nd.axis == 1 ? q.y : q.z;
    const float d = qc - nd.split;
    const int first = d < 0 ? nd.left : nd.right, second = d < 0 ? nd.right : nd.left;
    search(first, q, k, heap);
    if ((int)heap.size() < k || d * d <= heap.front().first) search(second, q, k, heap);
  }
  std::vector<int> order_;
  std::vector<Node> nodes_;
};
}}  // namespace pcl::search
