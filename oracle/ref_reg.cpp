// TEST INFRASTRUCTURE ONLY — never linked into the product (liblsdreg.so).
//
// C-ABI wrapper around the UNMODIFIED reference matcher sources, compiled where they lie under
// /root/reference/slam/thirdparty/fast_gicp by oracle/Makefile into oracle/_ref/libref_reg.so:
//   * fast_gicp::LsqRegistration  include/fast_gicp/gicp/impl/lsq_registration_impl.hpp   (LM on SE(3))
//   * fast_gicp::FastGICP         include/fast_gicp/gicp/impl/fast_gicp_impl.hpp
//   * fast_gicp::FastVGICP        include/fast_gicp/gicp/impl/fast_vgicp_impl.hpp + fast_vgicp_voxel.hpp
//   * fast_gicp::se3_exp          include/fast_gicp/so3/so3.hpp
// PCL and Boost are absent from this image: oracle/ref_shim_reg supplies pcl::Registration,
// pcl::PointCloud, pcl::search::KdTree (an exact k-d tree) and two Boost stubs, all ours.  Only this
// wrapper and those shims are ours; every line of matcher arithmetic executed is the reference's.
#include <pcl/point_types.h>
#include <pcl/point_cloud.h>
#include <pcl/registration/registration.h>

#include <fast_gicp/gicp/fast_gicp.hpp>
#include <fast_gicp/gicp/fast_vgicp.hpp>
#include <fast_gicp/gicp/impl/lsq_registration_impl.hpp>
#include <fast_gicp/gicp/impl/fast_gicp_impl.hpp>
#include <fast_gicp/gicp/impl/fast_vgicp_impl.hpp>

#include <cstring>

#include <hdl_graph_slam/information_matrix_calculator.hpp>   // compiled from the reference's .cpp (oracle/Makefile)

using P = pcl::PointXYZI;
using Cloud = pcl::PointCloud<P>;

struct GicpX : fast_gicp::FastGICP<P, P> {
  using fast_gicp::FastGICP<P, P>::linearize;
  using fast_gicp::FastGICP<P, P>::compute_error;
  const std::vector<int>& corr() const { return correspondences_; }
};
struct VgicpX : fast_gicp::FastVGICP<P, P> {
  using fast_gicp::FastVGICP<P, P>::linearize;
  using fast_gicp::FastVGICP<P, P>::compute_error;
  fast_gicp::GaussianVoxelMap<P>* map() {
    if (!voxelmap_) {  // built lazily by linearize() in the reference (fast_vgicp_impl.hpp:121-124)
      voxelmap_.reset(new fast_gicp::GaussianVoxelMap<P>(voxel_resolution_, voxel_mode_));
      voxelmap_->create_voxelmap(*target_, target_covs_);
    }
    return voxelmap_.get();
  }
  size_t n_corr() const { return voxel_correspondences_.size(); }
};

struct RefReg {
  int kind;  // 1 = FastGICP, 2 = FastVGICP
  std::shared_ptr<GicpX> g;
  std::shared_ptr<VgicpX> v;
  Cloud::Ptr src, tgt;
  fast_gicp::LsqRegistration<P, P>* base() { return kind == 1 ? static_cast<fast_gicp::LsqRegistration<P, P>*>(g.get()) : v.get(); }
  fast_gicp::FastGICP<P, P>* gicp() { return kind == 1 ? static_cast<fast_gicp::FastGICP<P, P>*>(g.get()) : v.get(); }
};

static Cloud::Ptr mk_cloud(const float* xyz, int n, int stride) {
  Cloud::Ptr c(new Cloud);
  c->points.resize(n);
  for (int i = 0; i < n; i++) { P& p = c->points[i]; p.x = xyz[(size_t)stride * i]; p.y = xyz[(size_t)stride * i + 1]; p.z = xyz[(size_t)stride * i + 2]; p.w = 1.f; }
  return c;
}
static Eigen::Isometry3d iso(const double* T16) {
  Eigen::Isometry3d t = Eigen::Isometry3d::Identity();
  for (int a = 0; a < 3; a++) { for (int b = 0; b < 3; b++) t.linear()(a, b) = T16[4 * a + b]; t.translation()(a) = T16[4 * a + 3]; }
  return t;
}

extern "C" {

void* ref_reg_create(int kind, int threads) {
  RefReg* r = new RefReg;
  r->kind = kind;
  if (kind == 1) { r->g.reset(new GicpX); r->g->setNumThreads(threads); }
  else { r->v.reset(new VgicpX); r->v->setNumThreads(threads); }
  return r;
}
void ref_reg_destroy(void* h) { delete static_cast<RefReg*>(h); }

// registrations.cpp:33-41 / :56-66 set these per method; neighbor: 1 / 7 / 27 (VGICP only)
void ref_reg_config(void* h, int max_iterations, double trans_eps, double rot_eps_deg, double max_corr, int k_corr,
                    double resolution, int neighbor, long long max_process_time_us) {
  RefReg* r = static_cast<RefReg*>(h);
  r->base()->setMaximumIterations(max_iterations);
  r->base()->setTransformationEpsilon(trans_eps);
  r->base()->setRotationEpsilon(rot_eps_deg);
  if (max_corr > 0) r->base()->setMaxCorrespondenceDistance(max_corr);
  r->gicp()->setCorrespondenceRandomness(k_corr);
  if (max_process_time_us > 0) r->base()->setMaxProcessTime(max_process_time_us);
  if (r->kind == 2) {
    r->v->setResolution(resolution);
    r->v->setNeighborSearchMethod(neighbor == 27 ? fast_gicp::NeighborSearchMethod::DIRECT27
                                  : neighbor == 7 ? fast_gicp::NeighborSearchMethod::DIRECT7 : fast_gicp::NeighborSearchMethod::DIRECT1);
  }
}
void ref_reg_set_source(void* h, const float* xyz, int n, int stride) {
  RefReg* r = static_cast<RefReg*>(h);
  r->src = mk_cloud(xyz, n, stride);
  r->base()->setInputSource(r->src);
}
void ref_reg_set_target(void* h, const float* xyz, int n, int stride) {
  RefReg* r = static_cast<RefReg*>(h);
  r->tgt = mk_cloud(xyz, n, stride);
  r->base()->setInputTarget(r->tgt);
}
// align(guess) -> final transformation (row-major 4x4 float), hasConverged
int ref_reg_align(void* h, const float* guess16, float* out16) {
  RefReg* r = static_cast<RefReg*>(h);
  Eigen::Matrix4f g;
  for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) g(a, b) = guess16[4 * a + b];
  Cloud out;
  r->base()->align(out, g);
  Eigen::Matrix4f T = r->base()->getFinalTransformation();
  for (int a = 0; a < 4; a++) for (int b = 0; b < 4; b++) out16[4 * a + b] = T(a, b);
  return r->base()->hasConverged() ? 1 : 0;
}
double ref_reg_fitness(void* h, double max_range) { return static_cast<RefReg*>(h)->base()->getFitnessScore(max_range); }

// linearize(T) (update_correspondences + H, b) — T row-major 4x4 double; H36/b6 may be null
double ref_reg_linearize(void* h, const double* T16, double* H36, double* b6, int* n_corr) {
  RefReg* r = static_cast<RefReg*>(h);
  Eigen::Matrix<double, 6, 6> H; Eigen::Matrix<double, 6, 1> b;
  const Eigen::Isometry3d T = iso(T16);
  double e;
  if (r->kind == 1) e = H36 ? r->g->linearize(T, &H, &b) : r->g->linearize(T, nullptr, nullptr);
  else e = H36 ? r->v->linearize(T, &H, &b) : r->v->linearize(T, nullptr, nullptr);
  if (H36) { for (int a = 0; a < 6; a++) { for (int c = 0; c < 6; c++) H36[6 * a + c] = H(a, c); b6[a] = b(a); } }
  if (n_corr) {
    if (r->kind == 1) { int c = 0; for (int v : r->g->corr()) c += v >= 0; *n_corr = c; }
    else *n_corr = (int)r->v->n_corr();
  }
  return e;
}
double ref_reg_compute_error(void* h, const double* T16) {
  RefReg* r = static_cast<RefReg*>(h);
  return r->kind == 1 ? r->g->compute_error(iso(T16)) : r->v->compute_error(iso(T16));
}
void ref_reg_get_corr(void* h, int* out) {
  RefReg* r = static_cast<RefReg*>(h);
  if (r->kind == 1) memcpy(out, r->g->corr().data(), r->g->corr().size() * sizeof(int));
}
// covariances of the source (which = 0) / target (1) cloud: [n, 9] row-major 3x3 blocks
void ref_reg_get_covs(void* h, int which, double* out) {
  RefReg* r = static_cast<RefReg*>(h);
  const auto& cv = which ? r->gicp()->getTargetCovariances() : r->gicp()->getSourceCovariances();
  for (size_t i = 0; i < cv.size(); i++) for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) out[9 * i + 3 * a + b] = cv[i](a, b);
}
// FastVGICP target voxel at integer coordinate: returns num_points (0 = none); mean[3], cov[9]
int ref_vgicp_voxel(void* h, int x, int y, int z, double* mean, double* cov) {
  RefReg* r = static_cast<RefReg*>(h);
  if (r->kind != 2) return -1;
  auto vx = r->v->map()->lookup_voxel(Eigen::Vector3i(x, y, z));
  if (!vx) return 0;
  for (int a = 0; a < 3; a++) { mean[a] = vx->mean(a); for (int b = 0; b < 3; b++) cov[3 * a + b] = vx->cov(a, b); }
  return vx->num_points;
}
void ref_vgicp_coord(void* h, const double* xyz, int* c) {
  RefReg* r = static_cast<RefReg*>(h);
  Eigen::Vector3i v = r->v->map()->voxel_coord(Eigen::Vector4d(xyz[0], xyz[1], xyz[2], 1.0));
  c[0] = v(0); c[1] = v(1); c[2] = v(2);
}
void ref_se3_exp(const double* a6, double* T16) {
  Eigen::Matrix<double, 6, 1> a;
  for (int i = 0; i < 6; i++) a(i) = a6[i];
  const Eigen::Matrix4d T = fast_gicp::se3_exp(a).matrix();
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) T16[4 * i + j] = T(i, j);
}

}  // extern "C"


// InformationMatrixCalculator::calc_fitness_score (slam/backend/hdl_graph_slam/src/hdl_graph_slam/
// information_matrix_calculator.cpp:71-102), compiled unmodified: the in-tree twin of pcl::Registration::getFitnessScore
// (mean squared nearest-neighbour distance of the transformed source in the target, nn_dists[0] <= max_range compared
// UNSQUARED).  cloud1 = target, cloud2 = source, relpose maps source into target.
extern "C" double ref_calc_fitness_score(const float* tgt, int n_t, int t_stride, const float* src, int n_s, int s_stride, const double* T16, double max_range) {
  Cloud::Ptr c1(new Cloud), c2(new Cloud);
  c1->points.resize(n_t); c2->points.resize(n_s);
  for (int i = 0; i < n_t; i++) { P& p = c1->points[i]; p.x = tgt[(size_t)t_stride * i]; p.y = tgt[(size_t)t_stride * i + 1]; p.z = tgt[(size_t)t_stride * i + 2]; p.w = 1.f; }
  for (int i = 0; i < n_s; i++) { P& p = c2->points[i]; p.x = src[(size_t)s_stride * i]; p.y = src[(size_t)s_stride * i + 1]; p.z = src[(size_t)s_stride * i + 2]; p.w = 1.f; }
  Eigen::Isometry3d rel = Eigen::Isometry3d::Identity();
  for (int a = 0; a < 3; a++) { for (int b = 0; b < 3; b++) rel.linear()(a, b) = T16[4 * a + b]; rel.translation()(a) = T16[4 * a + 3]; }
  return hdl_graph_slam::InformationMatrixCalculator::calc_fitness_score(c1, c2, rel, max_range);
}
