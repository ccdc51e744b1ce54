// TEST INFRASTRUCTURE ONLY: stand-in for pcl::Registration (PCL 1.9.1, absent here) carrying exactly
// the members and the align()/getFitnessScore() protocol that the reference's fast_gicp classes use.
#pragma once
#include <iostream>
#include <limits>
#include <memory>
#include <string>
#include <stdexcept>
#include <omp.h>
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <Eigen/SVD>
#include <Eigen/Cholesky>
#include <Eigen/LU>
#include <pcl/point_cloud.h>
#include <pcl/search/search.h>
namespace boost { template <typename T> using shared_ptr = std::shared_ptr<T>; }
namespace pcl {
template <typename PointSource, typename PointTarget, typename Scalar = float>
class Registration {
 public:
  typedef Eigen::Matrix<Scalar, 4, 4> Matrix4;
  typedef pcl::PointCloud<PointSource> PointCloudSource;
  typedef typename PointCloudSource::Ptr PointCloudSourcePtr;
  typedef typename PointCloudSource::ConstPtr PointCloudSourceConstPtr;
  typedef pcl::PointCloud<PointTarget> PointCloudTarget;
  typedef typename PointCloudTarget::Ptr PointCloudTargetPtr;
  typedef typename PointCloudTarget::ConstPtr PointCloudTargetConstPtr;
  Registration() : tree_(new pcl::search::KdTree<PointTarget>) { final_transformation_.setIdentity(); }
  virtual ~Registration() {}
  virtual void setInputSource(const PointCloudSourceConstPtr& c) { input_ = c; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& c) { target_ = c; tree_dirty_ = true; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  void setMaxCorrespondenceDistance(double d) { corr_dist_threshold_ = d; }
  bool hasConverged() const { return converged_; }
  Matrix4 getFinalTransformation() const { return final_transformation_; }
  void align(PointCloudSource& out, const Matrix4& guess = Matrix4::Identity()) {
    converged_ = false; nr_iterations_ = 0;
    computeTransformation(out, guess);
  }
  // mean of the squared 1-NN distances d2 <= max_range (PCL compares the squared distance with max_range)
  double getFitnessScore(double max_range = std::numeric_limits<double>::max()) {
    if (tree_dirty_) { tree_->setInputCloud(target_); tree_dirty_ = false; }
    PointCloudSource tr;
    transformPointCloud(*input_, tr, final_transformation_);
    std::vector<int> ni(1); std::vector<float> nd(1);
    double s = 0; int nr = 0;
    for (size_t i = 0; i < tr.points.size(); i++) {
      tree_->nearestKSearch(tr.points[i], 1, ni, nd);
      if (nd[0] <= max_range) { s += nd[0]; nr++; }
    }
    return nr > 0 ? s / nr : std::numeric_limits<double>::max();
  }
  const std::string& getClassName() const { return reg_name_; }
 protected:
  virtual void computeTransformation(PointCloudSource& out, const Matrix4& guess) = 0;
  std::string reg_name_;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  std::shared_ptr<pcl::search::KdTree<PointTarget>> tree_;
  bool tree_dirty_ = true;
  int nr_iterations_ = 0, max_iterations_ = 10;
  Matrix4 final_transformation_;
  double transformation_epsilon_ = 0.0;
  double corr_dist_threshold_ = std::sqrt(std::numeric_limits<double>::max());
  bool converged_ = false;
};
}  // namespace pcl
