// TEST INFRASTRUCTURE ONLY — never linked into the product (liblsdreg.so).
//
// The WHOLE reference LIO front-end, unmodified, as a CPU oracle: slam/mapping/fastlio/src/laserMapping.cpp (fastlio_init,
// fastlio_imu_enqueue, fastlio_pcl_enqueue, fastlio_main with sync_packages / ImuProcess / h_share_model_geometric /
// map_incremental / the IKFoM filter), src/preprocess.cpp, include/ikd-Tree/ikd_Tree.cpp, compiled where they lie by
// oracle/Makefile into oracle/_ref/libref_fastlio.so.  Ours: this wrapper and the shims in oracle/ref_shim_fastlio
// (PCL containers, pcl::VoxelGrid — the one algorithmic piece that is a restatement, PCL being external — record types,
// logging) and oracle/ref_shim_ikfom (Boost.Preprocessor).
#include <omp.h>

#include <cstring>

#include "laserMapping.cpp"  // resolved through -I<reference>/slam/mapping/fastlio/src

Eigen::Matrix4d getTransformFromRPYT(double, double, double, double, double, double) {
  std::cerr << "ref_fastlio: the INS path is not part of this oracle" << std::endl;
  abort();
}

extern "C" {

int ref_fastlio_init(const double* extT3, const double* extR9, int filter_num, int max_point_num, double scan_period, int undistort) {
  std::vector<double> t(extT3, extT3 + 3), r(extR9, extR9 + 9);
  return fastlio_init(t, r, filter_num, max_point_num, scan_period, undistort != 0);
}
// acc in m/s^2: fastlio_imu_enqueue divides by 9.81 (laserMapping.cpp:414)
void ref_fastlio_imu(double stamp, const double* gyr3, const double* acc3) {
  ImuType m;
  m.stamp = stamp;
  m.gyr = Eigen::Vector3d(gyr3[0], gyr3[1], gyr3[2]);
  m.acc = Eigen::Vector3d(acc3[0], acc3[1], acc3[2]);
  fastlio_imu_enqueue(m);
}
// xyzi [n,4]; stamp_us [n] = per-point time relative to header_stamp_us (PointAttr::stamp)
void ref_fastlio_scan(const float* xyzi, const uint32_t* stamp_us, int n, uint64_t header_stamp_us) {
  PointCloudAttrPtr p(new PointCloudAttr());
  p->cloud->header.stamp = header_stamp_us;
  p->cloud->points.resize(n);
  p->attr.resize(n);
  for (int i = 0; i < n; i++) {
    Point& q = p->cloud->points[i];
    q.x = xyzi[4 * i]; q.y = xyzi[4 * i + 1]; q.z = xyzi[4 * i + 2]; q.intensity = xyzi[4 * i + 3];
    p->attr[i].id = 0; p->attr[i].stamp = stamp_us[i];
  }
  fastlio_pcl_enqueue(p);
}
int ref_fastlio_main() { return fastlio_main() ? 1 : 0; }
int ref_fastlio_is_init() { return fastlio_is_init() ? 1 : 0; }
// the filter state after the last scan: double[26] in the layout of include/lsdreg.h, covariance 23x23
void ref_fastlio_get_state(double* x26, double* P529) {
  const state_ikfom s = kf.get_x();
  for (int i = 0; i < 3; i++) { x26[i] = s.pos[i]; x26[11 + i] = s.offset_T_L_I[i]; x26[14 + i] = s.vel[i]; x26[17 + i] = s.bg[i]; x26[20 + i] = s.ba[i]; x26[23 + i] = s.grav[i]; }
  const Eigen::Vector4d q = s.rot.coeffs(), ql = s.offset_R_L_I.coeffs();
  for (int i = 0; i < 4; i++) { x26[3 + i] = q[i]; x26[7 + i] = ql[i]; }
  if (P529) { const auto& P = kf.get_P(); for (int a = 0; a < 23; a++) for (int b = 0; b < 23; b++) P529[23 * a + b] = P(a, b); }
}
// taps of the last scan
int ref_fastlio_counts(int* feats_down, int* effective, int* map_cells, int* degenerate) {
  if (feats_down) *feats_down = feats_down_size;
  if (effective) *effective = effct_feat_num;
  if (map_cells) *map_cells = ivox ? (int)ivox->NumValidGrids() : 0;
  if (degenerate) *degenerate = is_degenerate ? 1 : 0;
  return 0;
}
int ref_fastlio_get_down(float* xyzi, int cap) {
  const int n = (int)feats_down_body->points.size();
  for (int i = 0; i < n && i < cap; i++) { const PointType& p = feats_down_body->points[i]; xyzi[4 * i] = p.x; xyzi[4 * i + 1] = p.y; xyzi[4 * i + 2] = p.z; xyzi[4 * i + 3] = p.intensity; }
  return n;
}

}  // extern "C"
