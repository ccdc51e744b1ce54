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

// The reference's build defines MP_EN and MP_PROC_NUM (= 8 on x86_64, fastlio/CMakeLists.txt:20-25); oracle/Makefile
// passes -DMP_EN -DMP_PROC_NUM=ref_mp_threads so that the bench arm can give it every host thread it can use.
int ref_mp_threads = 8;

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

// ------------------------------------------------------------------------------------------------------------------
// CPU baseline of bench.py (--impl reference / cpu_baseline): the per-scan hot path of fastlio_main on a prebuilt map.
// ------------------------------------------------------------------------------------------------------------------
void ref_fastlio_set_threads(int n) { ref_mp_threads = n > 0 ? n : 1; }

// fastlio_init, then an iVox that can hold the benchmark map (the reference caps it at 100 000 voxels with LRU eviction,
// laserMapping.cpp:1063; SURVEY.md section 8a row a2: the CPU baseline must raise capacity_) on NEARBY18, the steady state.
int ref_fastlio_bench_init(size_t capacity) {
  std::vector<double> t(3, 0.0), r = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  int rc = fastlio_init(t, r, 1, -1, 0.1, false);
  IVoxType::Options o;
  o.resolution_ = 0.5;
  o.nearby_type_ = IVoxType::NearbyType::NEARBY18;
  o.capacity_ = capacity;
  o.max_distance_ = 100.0;
  ivox = std::make_shared<IVoxType>(o);
  flg_EKF_inited = true;
  return rc;
}
void ref_fastlio_map_add(const float* xyz, int n, int stride) {
  PointVector pv(n);
  for (int i = 0; i < n; i++) { pv[i].x = xyz[(size_t)stride * i]; pv[i].y = xyz[(size_t)stride * i + 1]; pv[i].z = xyz[(size_t)stride * i + 2]; pv[i].intensity = 0; }
  ivox->AddPoints(pv, 0.0);
}
// One scan through the statements of fastlio_main that follow ImuProcess (laserMapping.cpp:1198-1302), in their order and
// with the reference's own objects: VoxelGrid -> kf.update_iterated_dyn_share_modified (h_share_model) -> map_incremental.
// xyzi: the undistorted scan; x26 / P529: prior in, posterior out.  Returns feats_down_size, or -1 when the scan is skipped.
int ref_fastlio_pass(const float* xyzi, int n, double* x26, double* P529) {
  state_ikfom s;
  for (int i = 0; i < 3; i++) { s.pos[i] = x26[i]; s.offset_T_L_I[i] = x26[11 + i]; s.vel[i] = x26[14 + i]; s.bg[i] = x26[17 + i]; s.ba[i] = x26[20 + i]; s.grav.vec[i] = x26[23 + i]; }
  s.rot.coeffs() = Eigen::Vector4d(x26[3], x26[4], x26[5], x26[6]);
  s.offset_R_L_I.coeffs() = Eigen::Vector4d(x26[7], x26[8], x26[9], x26[10]);
  kf.change_x(s);
  esekfom::esekf<state_ikfom, 12, input_ikfom>::cov P;
  for (int a = 0; a < 23; a++) for (int b = 0; b < 23; b++) P(a, b) = P529[23 * a + b];
  kf.change_P(P);
  feats_undistort->points.resize(n);
  for (int i = 0; i < n; i++) {
    PointType& q = feats_undistort->points[i];
    q.x = xyzi[4 * i]; q.y = xyzi[4 * i + 1]; q.z = xyzi[4 * i + 2]; q.intensity = xyzi[4 * i + 3];
    q.normal_x = q.normal_y = q.normal_z = 0; q.curvature = 0;
  }
  state_point = kf.get_x();
  downSizeFilterSurf.setInputCloud(feats_undistort);                  // :1206-1209
  downSizeFilterSurf.filter(*feats_down_body);
  feats_down_size = feats_down_body->points.size();
  if (!ivox->NumValidGrids() || feats_down_size < 5) return -1;       // :1227-1256
  normvec->resize(feats_down_size);                                   // :1258-1259
  feats_down_world->resize(feats_down_size);
  Nearest_Points.resize(feats_down_size);                             // :1273
  double solve_H_time = 0;
  kf.update_iterated_dyn_share_modified(LASER_POINT_COV, solve_H_time);   // :1283
  state_point = kf.get_x();                                           // :1288-1291 (travel_distance stays 0: no eviction)
  pos_lid = state_point.pos + state_point.rot * state_point.offset_T_L_I;
  map_incremental();                                                  // :1302
  ref_fastlio_get_state(x26, P529);
  return feats_down_size;
}

}  // extern "C"
