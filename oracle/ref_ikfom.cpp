// TEST INFRASTRUCTURE ONLY — never linked into the product (liblsdreg.so).
//
// C-ABI wrapper around the UNMODIFIED reference filter and IMU stage, compiled where they lie under
// /root/reference/slam/mapping/fastlio by oracle/Makefile into oracle/_ref/libref_ikfom.so:
//   * esekfom::esekf<state_ikfom, 12, input_ikfom>   include/IKFoM_toolkit/esekfom/esekfom.hpp
//       predict :279-383, update_iterated_dyn_share_modified :1619-1931
//   * state_ikfom / get_f / df_dx / df_dw            include/use-ikfom.hpp
//   * ImuProcess (IMU_init, UndistortPcl, Process)   src/IMU_Processing.hpp
// IKFoM needs Boost.Preprocessor, which is not installed: oracle/ref_shim_ikfom/boost/preprocessor/seq.hpp
// re-implements from scratch the dozen macros mtk/build_manifold.hpp uses; PCL point types / logging come from the
// same kind of shims as oracle/ref_lio.cpp.  Only this wrapper and those shims are ours.
// State vectors are double[26] = pos, rot (x,y,z,w), offset_R (x,y,z,w), offset_T, vel, bg, ba, grav
// (include/lsdreg.h); covariances double[23*23] row-major.
#include <omp.h>

#include <pcl/point_types.h>
typedef pcl::PointXYZINormal PointType_;

#define private public  // test build only: read ImuProcess::IMUpose (the reference keeps it private)
#include <IMU_Processing.hpp>
#undef private

#include <cstring>
#include <functional>
#include <vector>

typedef esekfom::esekf<state_ikfom, 12, input_ikfom> Kf;

static void to26(const state_ikfom& s, double* x) {
  for (int i = 0; i < 3; i++) { x[i] = s.pos[i]; x[11 + i] = s.offset_T_L_I[i]; x[14 + i] = s.vel[i]; x[17 + i] = s.bg[i]; x[20 + i] = s.ba[i]; x[23 + i] = s.grav[i]; }
  const Eigen::Vector4d q = s.rot.coeffs(), ql = s.offset_R_L_I.coeffs();
  for (int i = 0; i < 4; i++) { x[3 + i] = q[i]; x[7 + i] = ql[i]; }
}
static state_ikfom from26(const double* x) {
  state_ikfom s;
  for (int i = 0; i < 3; i++) { s.pos[i] = x[i]; s.offset_T_L_I[i] = x[11 + i]; s.vel[i] = x[14 + i]; s.bg[i] = x[17 + i]; s.ba[i] = x[20 + i]; s.grav.vec[i] = x[23 + i]; }
  s.rot.coeffs() = Eigen::Vector4d(x[3], x[4], x[5], x[6]);
  s.offset_R_L_I.coeffs() = Eigen::Vector4d(x[7], x[8], x[9], x[10]);
  return s;
}
static void set_kf(Kf& kf, const double* x26, const double* P529) {
  state_ikfom s = from26(x26);
  kf.change_x(s);
  Kf::cov P;
  for (int a = 0; a < 23; a++) for (int b = 0; b < 23; b++) P(a, b) = P529[23 * a + b];
  kf.change_P(P);
}
static void get_kf(const Kf& kf, double* x26, double* P529) {
  to26(kf.get_x(), x26);
  const Kf::cov& P = kf.get_P();
  for (int a = 0; a < 23; a++) for (int b = 0; b < 23; b++) P529[23 * a + b] = P(a, b);
}

// ---- tabulated measurement model for update_iterated_dyn_share_modified
struct Table { const double* rows; const double* h; const int* n_rows; int n_table, max_rows, calls; int* converge_log; };
static thread_local Table* g_table = nullptr;
static void h_table(state_ikfom&, esekfom::dyn_share_datastruct<double>& d) {
  Table* t = g_table;
  const int e = t->calls < t->n_table ? t->calls : t->n_table - 1;
  if (t->converge_log && t->calls < 16) t->converge_log[t->calls] = d.converge ? 1 : 0;  // what h_share_model reads to decide on a new neighbour search
  t->calls++;
  const int n = t->n_rows[e];
  if (n < 1) { d.valid = false; return; }
  d.h_x = Eigen::MatrixXd::Zero(n, 15);  // h_share_model_geometric: N x 15, laserMapping.cpp:900
  d.h.resize(n);
  for (int i = 0; i < n; i++) {
    for (int c = 0; c < 6; c++) d.h_x(i, c) = t->rows[((size_t)e * t->max_rows + i) * 6 + c];
    d.h(i) = t->h[(size_t)e * t->max_rows + i];
  }
}

extern "C" {

void ref_ikfom_predict(double* x26, double* P529, double dt, const double* Q144, const double* acc3, const double* gyro3) {
  Kf kf;
  double epsi[23];
  std::fill(epsi, epsi + 23, 0.001);
  kf.init_dyn_share(get_f, df_dx, df_dw, h_table, 4, epsi);
  set_kf(kf, x26, P529);
  Eigen::Matrix<double, 12, 12> Q;
  for (int a = 0; a < 12; a++) for (int b = 0; b < 12; b++) Q(a, b) = Q144[12 * a + b];
  input_ikfom in;
  for (int i = 0; i < 3; i++) { in.acc[i] = acc3[i]; in.gyro[i] = gyro3[i]; }
  kf.predict(dt, Q, in);
  get_kf(kf, x26, P529);
}

// rows: [n_table, max_rows, 6] (the 6 non-zero columns of h_x), h: [n_table, max_rows], n_rows: [n_table] (< 1 = invalid).
// Returns the number of measurement-model evaluations.
int ref_ikfom_update_rows(double* x26, double* P529, const double* rows, const double* h, const int* n_rows, int n_table, int max_rows,
                          double R, int max_iterations, double eps, int* converge_log16) {
  Kf kf;
  double epsi[23];
  std::fill(epsi, epsi + 23, eps);
  kf.init_dyn_share(get_f, df_dx, df_dw, h_table, max_iterations, epsi);
  set_kf(kf, x26, P529);
  Table t{rows, h, n_rows, n_table, max_rows, 0, converge_log16};
  g_table = &t;
  double solve = 0;
  kf.update_iterated_dyn_share_modified(R, solve);
  g_table = nullptr;
  get_kf(kf, x26, P529);
  return t.calls;
}

void ref_ikfom_boxplus(double* x26, const double* d23) {
  state_ikfom s = from26(x26);
  Eigen::Matrix<double, 23, 1> d;
  for (int i = 0; i < 23; i++) d[i] = d23[i];
  s.boxplus(d);
  to26(s, x26);
}
void ref_ikfom_boxminus(const double* a26, const double* b26, double* out23) {
  state_ikfom a = from26(a26), b = from26(b26);
  Eigen::Matrix<double, 23, 1> d;
  a.boxminus(d, b);
  for (int i = 0; i < 23; i++) out23[i] = d[i];
}

// ---- ImuProcess
struct RefImu { ImuProcess imu; Kf kf; };

void* ref_imu_create(const double* ext_R9, const double* ext_t3, double gyr_cov, double acc_cov, double b_gyr_cov, double b_acc_cov, int undistort) {
  RefImu* h = new RefImu;
  double epsi[23];
  std::fill(epsi, epsi + 23, 0.001);
  h->kf.init_dyn_share(get_f, df_dx, df_dw, h_table, 4, epsi);
  M3D R; V3D t;
  for (int a = 0; a < 3; a++) { for (int b = 0; b < 3; b++) R(a, b) = ext_R9[3 * a + b]; t[a] = ext_t3[a]; }
  h->imu.set_extrinsic(t, R);                                     // laserMapping.cpp:1101-1106
  h->imu.set_gyr_cov(V3D(gyr_cov, gyr_cov, gyr_cov));
  h->imu.set_acc_cov(V3D(acc_cov, acc_cov, acc_cov));
  h->imu.set_gyr_bias_cov(V3D(b_gyr_cov, b_gyr_cov, b_gyr_cov));
  h->imu.set_acc_bias_cov(V3D(b_acc_cov, b_acc_cov, b_acc_cov));
  h->imu.undistort = undistort != 0;
  return h;
}
void ref_imu_destroy(void* p) { delete static_cast<RefImu*>(p); }

// ImuProcess::Process.  Returns the number of points in the undistorted cloud (0 while initialising);
// out_xyzi [n,4] receives it in the reference's order (sorted by time), out_time_ms [n] the sorted times.
int ref_imu_process(void* p, const double* imu7, int n_imu, const double* ins_vel3_or_null, double beg, double end, const float* xyzi,
                    const float* time_ms, int n, double* x26, double* P529, float* out_xyzi, float* out_time_ms) {
  RefImu* h = static_cast<RefImu*>(p);
  set_kf(h->kf, x26, P529);
  MeasureGroup meas;
  meas.lidar_beg_time = beg; meas.lidar_end_time = end;
  meas.lidar->points.resize(n);
  for (int i = 0; i < n; i++) {
    PointType& q = meas.lidar->points[i];
    q.x = xyzi[4 * i]; q.y = xyzi[4 * i + 1]; q.z = xyzi[4 * i + 2]; q.intensity = xyzi[4 * i + 3]; q.curvature = time_ms[i];
  }
  for (int i = 0; i < n_imu; i++) {
    ImuType m;
    m.stamp = imu7[7 * i];
    m.gyr = Eigen::Vector3d(imu7[7 * i + 1], imu7[7 * i + 2], imu7[7 * i + 3]);
    m.acc = Eigen::Vector3d(imu7[7 * i + 4], imu7[7 * i + 5], imu7[7 * i + 6]);
    meas.imu.push_back(m);
  }
  if (ins_vel3_or_null) { RTKType r; r.Ve = ins_vel3_or_null[0]; r.Vn = ins_vel3_or_null[1]; r.Vu = ins_vel3_or_null[2]; meas.ins.push_back(r); }
  PointCloudXYZI::Ptr out(new PointCloudXYZI());
  h->imu.Process(meas, h->kf, out);
  get_kf(h->kf, x26, P529);
  const int m = (int)out->points.size();
  for (int i = 0; i < m; i++) {
    const PointType& q = out->points[i];
    out_xyzi[4 * i] = q.x; out_xyzi[4 * i + 1] = q.y; out_xyzi[4 * i + 2] = q.z; out_xyzi[4 * i + 3] = q.intensity;
    out_time_ms[i] = q.curvature;
  }
  return m;
}
// IMUpose of the last scan: [n, 22] = (offset_time, acc, gyr, vel, pos, rot)
int ref_imu_get_poses(void* p, double* poses22, int cap) {
  RefImu* h = static_cast<RefImu*>(p);
  const int n = (int)h->imu.IMUpose.size();
  for (int i = 0; i < n && i < cap; i++) {
    const Pose6D& q = h->imu.IMUpose[i];
    double* o = poses22 + 22 * i;
    o[0] = q.offset_time;
    for (int k = 0; k < 3; k++) { o[1 + k] = q.acc[k]; o[4 + k] = q.gyr[k]; o[7 + k] = q.vel[k]; o[10 + k] = q.pos[k]; }
    for (int k = 0; k < 9; k++) o[13 + k] = q.rot[k];
  }
  return n;
}
int ref_imu_is_init(void* p) { return static_cast<RefImu*>(p)->imu.IsInit() ? 1 : 0; }

}  // extern "C"
