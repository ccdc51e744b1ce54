// fastlio_seam.cu — the reference's LIO seam (SURVEY.md 8b "C++ seam 1"), host side, on top of lsd_imu_* and lsd_lio_*.
//
// Replaces the eight free functions slam/mapping/fastlio/src/fastlio.cpp:9-16 binds, defined in
// slam/mapping/fastlio/src/laserMapping.cpp (reference):
//   fastlio_init :1025-1124, fastlio_imu_enqueue :397-416, fastlio_ins_enqueue :418-443, fastlio_pcl_enqueue :311-330
//   (Preprocess::velodyne_handler, preprocess.cpp:280-427), sync_packages :445-520, fastlio_main :1126-1387,
//   fastlio_odometry :690-710, fastlio_state :712-738, fastlio_is_init :740-743.
// The reference keeps this state in file-scope globals guarded by mtx_buffer; here it is a handle with the same
// producer / consumer contract: any thread may enqueue, ONE thread calls lsd_fastlio_main.
// Host code only in this file: queues, package synchronisation, point decimation; the points go to the device once, in
// lsd_imu_process (undistortion), and the undistorted cloud never comes back (lsd_lio_scan_dev).
#include <deque>
#include <mutex>
#include <vector>

#include "eskf.hpp"
#include "lsd_common.cuh"

struct lsd_fastlio {
  // fastlio_init arguments
  double ext_t[3], ext_R[9];
  int filter_num = 1, max_point_num = -1, undistort = 1;
  double scan_period = 0.1;
  int map_log2_lines = 22, max_scan_points = 262144;
  // buffers (laserMapping.cpp: lidar_buffer / time_buffer / imu_buffer / ins_buffer under mtx_buffer)
  struct Scan { std::vector<float> xyzi, t_ms; double beg; };
  struct Imu { double v[7]; };          // stamp, gyr xyz, acc xyz (acc already / 9.81)
  struct Ins { double stamp, vel[3]; };
  std::mutex mtx;
  std::deque<Scan> scan_buf;
  std::deque<Imu> imu_buf;
  std::deque<Ins> ins_buf;
  // fastlio_main state
  bool first_scan = true;
  double first_lidar_time = 0.0;
  double x[26], P[529];                  // kf
  double state_point[26];
  double travel_distance = 0.0, last_pos_lid[3] = {0, 0, 0};
  int nearby = LSD_STENCIL_NEARBY74;
  lsd_imu_t* imu = nullptr;
  lsd_lio_t* lio = nullptr;
  lsd_lio_info_t info{};
  int last_status = 0;
  Scan cur;                              // the package being processed
  std::vector<double> cur_imu;
};

namespace lsd {

constexpr double kInitTime = 0.1;   // INIT_TIME, laserMapping.cpp:70
constexpr double kBlind = 0.1;      // p_pre->blind, laserMapping.cpp:1094

// a default-constructed esekf / state_ikfom: identity rotations, zeros, grav = S2() = length * e_x (S2.hpp:62-66), P = I
static void default_state(double* x, double* P) {
  memset(x, 0, 26 * sizeof(double));
  x[eskf::S_ROT + 3] = 1.0; x[eskf::S_OFFR + 3] = 1.0;
  x[eskf::S_GRAV] = eskf::kS2Len;
  if (P) { memset(P, 0, 529 * sizeof(double)); for (int i = 0; i < 23; i++) P[i * 24] = 1.0; }
}

// sync_packages (laserMapping.cpp:445-520): one scan + every IMU / INS sample up to its end time
static bool sync_packages(lsd_fastlio* f, lsd_fastlio::Scan* scan, std::vector<double>* imu7, double* end, bool* have_ins, double* ins_vel) {
  std::lock_guard<std::mutex> lk(f->mtx);
  if (f->scan_buf.empty() || f->imu_buf.empty()) return false;
  *scan = std::move(f->scan_buf.front());
  f->scan_buf.pop_front();
  *end = scan->beg + f->scan_period;     // lidar_mean_scantime = scan_period (:1121)
  imu7->clear();
  while (!f->imu_buf.empty()) {
    if (f->imu_buf.front().v[0] > *end) break;
    imu7->insert(imu7->end(), f->imu_buf.front().v, f->imu_buf.front().v + 7);
    f->imu_buf.pop_front();
  }
  *have_ins = false;
  while (!f->ins_buf.empty()) {
    if (f->ins_buf.front().stamp > *end) break;
    *have_ins = true;                     // ImuProcess reads meas.ins.back() only (IMU_Processing.hpp:200-203)
    for (int i = 0; i < 3; i++) ins_vel[i] = f->ins_buf.front().vel[i];
    f->ins_buf.pop_front();
  }
  return true;
}

static lsd_status_t ensure_engines(lsd_fastlio* f) {
  if (f->imu && f->lio) return LSD_OK;
  lsd_imu_params_t ip;
  lsd_imu_default_params(&ip);
  memcpy(ip.ext_R, f->ext_R, sizeof(ip.ext_R)); memcpy(ip.ext_t, f->ext_t, sizeof(ip.ext_t));
  ip.undistort = f->undistort;
  lsd_status_t s = lsd_imu_create(&f->imu, &ip);
  if (s) return s;
  lsd_lio_params_t lp;
  lsd_lio_default_params(&lp);
  lp.ivox_nearby = LSD_STENCIL_NEARBY74;          // "switch to NEARBY18 after 1.0s", laserMapping.cpp:1062
  lp.map_log2_lines = f->map_log2_lines;
  lp.max_scan_points = f->max_scan_points;
  s = lsd_lio_create(&f->lio, &lp);
  if (s) return s;
  // ivox_options.capacity_ = 100000, max_distance_ = 100.0 (laserMapping.cpp:1063-1064): the map forgets what the
  // reference forgets, and a long drive cannot fill the table (2^22 lines for <= ~100 k live voxels; retired lines are
  // rehashed away)
  s = lsd_map_enable_lru(lsd_lio_map(f->lio), 100000, 100.0);
  if (s) return s;
  return lsd_lio_set_stale_rows(f->lio, 1);        // Nearest_Points outlives the scan (laserMapping.cpp:1273, ivox3d.h:155-157)
}

}  // namespace lsd

using namespace lsd;

extern "C" {

lsd_status_t lsd_fastlio_create(lsd_fastlio_t** out, const double* extT3, const double* extR9, int filter_num, int max_point_num,
                                double scan_period, int undistort) {
  if (!out || !extT3 || !extR9 || filter_num < 1 || !(scan_period > 0.0)) return LSD_ERR_INVALID;
  lsd_fastlio* f = new lsd_fastlio();
  memcpy(f->ext_t, extT3, sizeof(f->ext_t)); memcpy(f->ext_R, extR9, sizeof(f->ext_R));
  f->filter_num = filter_num; f->max_point_num = max_point_num; f->scan_period = scan_period; f->undistort = undistort ? 1 : 0;
  default_state(f->x, f->P);
  default_state(f->state_point, nullptr);
  *out = f;
  return LSD_OK;
}

lsd_status_t lsd_fastlio_set_capacity(lsd_fastlio_t* f, int map_log2_lines, int max_scan_points) {
  if (!f || f->lio || map_log2_lines < 10 || map_log2_lines > 34 || max_scan_points < 1024) return LSD_ERR_INVALID;
  f->map_log2_lines = map_log2_lines; f->max_scan_points = max_scan_points;
  return LSD_OK;
}

lsd_status_t lsd_fastlio_destroy(lsd_fastlio_t* f) {
  if (!f) return LSD_OK;
  if (f->lio) lsd_lio_destroy(f->lio);
  if (f->imu) lsd_imu_destroy(f->imu);
  delete f;
  return LSD_OK;
}

// fastlio_imu_enqueue: acc in m/s^2, stored / 9.81
lsd_status_t lsd_fastlio_imu_enqueue(lsd_fastlio_t* f, double stamp_s, const double* gyr3, const double* acc3) {
  if (!f || !gyr3 || !acc3) return LSD_ERR_INVALID;
  lsd_fastlio::Imu m;
  m.v[0] = stamp_s;
  for (int i = 0; i < 3; i++) { m.v[1 + i] = gyr3[i]; m.v[4 + i] = acc3[i] / 9.81; }
  std::lock_guard<std::mutex> lk(f->mtx);
  f->imu_buf.push_back(m);
  return LSD_OK;
}

// fastlio_ins_enqueue: ENU velocity -> ego (INS) frame -> IMU frame; the up component is dropped
lsd_status_t lsd_fastlio_ins_enqueue(lsd_fastlio_t* f, int rtk_valid, int is_wheel, uint64_t timestamp_us, const double* vel_enu3,
                                     double heading_deg, double pitch_deg, double roll_deg) {
  if (!f || !vel_enu3) return LSD_ERR_INVALID;
  if (!rtk_valid && !is_wheel) return LSD_OK;   // "check LLA or Wheel is available"
  // getTransformFromRPYT(0, 0, 0, -heading, pitch, roll) = Rz(yaw) Rx(pitch) Ry(roll) (slam_utils.cpp:89-96); its inverse rotates by R^T
  const double k = M_PI / 180.0;
  const double cy = cos(-heading_deg * k), sy = sin(-heading_deg * k), cp = cos(pitch_deg * k), sp = sin(pitch_deg * k),
               cr = cos(roll_deg * k), sr = sin(roll_deg * k);
  const double Rz[9] = {cy, -sy, 0, sy, cy, 0, 0, 0, 1}, Rx[9] = {1, 0, 0, 0, cp, -sp, 0, sp, cp}, Ry[9] = {cr, 0, sr, 0, 1, 0, -sr, 0, cr};
  double A[9], R[9];
  eskf::mm3(Rz, Rx, A); eskf::mm3(A, Ry, R);
  double v[3], w[3];
  for (int i = 0; i < 3; i++) v[i] = R[0 + i] * vel_enu3[0] + R[3 + i] * vel_enu3[1] + R[6 + i] * vel_enu3[2];   // R^T vel
  eskf::mv3(f->ext_R, v, w);
  lsd_fastlio::Ins m;
  m.stamp = (double)timestamp_us / 1000000.0;
  m.vel[0] = w[0]; m.vel[1] = w[1]; m.vel[2] = 0.0;
  std::lock_guard<std::mutex> lk(f->mtx);
  f->ins_buf.push_back(m);
  return LSD_OK;
}

// fastlio_pcl_enqueue -> Preprocess::process -> velodyne_handler: every point_filter_num-th point outside the blind zone,
// curvature = attr.stamp / 1000.0f (ms, float); stamp_us [n] = PointAttr::stamp relative to header_stamp_us
lsd_status_t lsd_fastlio_pcl_enqueue(lsd_fastlio_t* f, const float* xyzi, const uint32_t* stamp_us, int n, uint64_t header_stamp_us) {
  if (!f || n < 0 || (n > 0 && (!xyzi || !stamp_us))) return LSD_ERR_INVALID;
  lsd_fastlio::Scan s;
  s.beg = (double)header_stamp_us / 1000000.0;
  int step = f->filter_num;
  if (f->max_point_num > 0) step = std::max(1, n / f->max_point_num);
  s.xyzi.reserve((size_t)(n / step + 1) * 4); s.t_ms.reserve((size_t)n / step + 1);
  for (int i = 0; i < n; i++) {
    const float x = xyzi[4 * i], y = xyzi[4 * i + 1], z = xyzi[4 * i + 2];
    if (i % step != 0) continue;
    const float r2 = x * x + y * y + z * z;
    if (!((double)r2 > kBlind * kBlind)) continue;
    s.xyzi.insert(s.xyzi.end(), {x, y, z, xyzi[4 * i + 3]});
    s.t_ms.push_back((float)stamp_us[i] / 1000.0f);
  }
  std::lock_guard<std::mutex> lk(f->mtx);
  f->scan_buf.push_back(std::move(s));
  return LSD_OK;
}

// Parity tap (host only, no device): runs sync_packages and hands the package out instead of processing it.
// Returns 1 with the sizes filled (arrays copied up to the caps), 0 when no package is ready.
int lsd_fastlio_pop_package(lsd_fastlio_t* f, double* lidar_beg_time, double* lidar_end_time, int* n_points, float* xyzi, float* time_ms,
                            int cap_points, int* n_imu, double* imu7, int cap_imu, int* have_ins, double* ins_vel3) {
  if (!f || !lidar_beg_time || !lidar_end_time || !n_points || !n_imu) return LSD_ERR_INVALID;
  lsd_fastlio::Scan s; std::vector<double> imu; double end = 0.0, iv[3] = {0, 0, 0}; bool hi = false;
  if (!sync_packages(f, &s, &imu, &end, &hi, iv)) return 0;
  *lidar_beg_time = s.beg; *lidar_end_time = end;
  *n_points = (int)s.t_ms.size(); *n_imu = (int)(imu.size() / 7);
  const int cp = std::min(*n_points, cap_points), ci = std::min(*n_imu, cap_imu);
  if (xyzi && cp > 0) memcpy(xyzi, s.xyzi.data(), (size_t)cp * 16);
  if (time_ms && cp > 0) memcpy(time_ms, s.t_ms.data(), (size_t)cp * 4);
  if (imu7 && ci > 0) memcpy(imu7, imu.data(), (size_t)ci * 56);
  if (have_ins) *have_ins = hi ? 1 : 0;
  if (ins_vel3) for (int i = 0; i < 3; i++) ins_vel3[i] = iv[i];
  return 1;
}

// fastlio_main: 1 = a package was consumed (the reference's `true`), 0 = nothing to do, < 0 = error
int lsd_fastlio_main(lsd_fastlio_t* f) {
  if (!f) return LSD_ERR_INVALID;
  double end = 0.0, ins_vel[3] = {0, 0, 0};
  bool have_ins = false;
  if (!sync_packages(f, &f->cur, &f->cur_imu, &end, &have_ins, ins_vel)) return 0;
  f->last_status = LSD_OK;
  if (f->first_scan) {                                   // :1171-1177
    f->first_lidar_time = f->cur.beg;
    f->first_scan = false;
    return 1;
  }
  lsd_status_t s = ensure_engines(f);
  if (s < 0) return s;
  const int n = (int)f->cur.t_ms.size(), n_imu = (int)(f->cur_imu.size() / 7);
  int n_und = 0;
  s = lsd_imu_process(f->imu, f->cur_imu.data(), n_imu, have_ins ? ins_vel : nullptr, f->cur.beg, end, f->cur.xyzi.data(), f->cur.t_ms.data(), n,
                      f->x, f->P, &n_und);              // p_imu->Process(Measures, kf, feats_undistort), :1188
  if (s < 0) return s;
  memcpy(f->state_point, f->x, sizeof(f->x));           // state_point = kf.get_x()
  if (s == LSD_IMU_INITIALIZING || n_und == 0) { f->last_status = s; return 1; }   // "undistort points is empty", :1192-1196
  const double since = f->cur.beg - f->first_lidar_time;
  lsd_lio_set_ekf_inited(f->lio, since < kInitTime ? 0 : 1);                      // flg_EKF_inited, :1198
  if (f->nearby != LSD_STENCIL_NEARBY18 && since > 10.0 * kInitTime) {            // :1241-1243 (time-based: the seeding scan does not search)
    f->nearby = LSD_STENCIL_NEARBY18;
    lsd_lio_set_nearby(f->lio, f->nearby);
  }
  const float* d_cloud = nullptr; void* stream = nullptr;
  s = lsd_imu_get_cloud_dev(f->imu, &d_cloud, &n_und, &stream);
  if (s < 0) return s;
  if (cudaStreamSynchronize(static_cast<cudaStream_t>(stream)) != cudaSuccess) { set_error("lsd_fastlio_main: undistortion stream failed"); return LSD_ERR_CUDA; }
  s = lsd_lio_scan_dev(f->lio, d_cloud, n_und, f->x, f->P, &f->info);             // VoxelGrid .. map_incremental, :1204-1292
  if (s < 0) return s;
  f->last_status = s;                                   // LSD_MAP_SEEDED / LSD_SCAN_TOO_SMALL / LSD_NO_EFFECTIVE_POINTS / LSD_OK
  if (s == LSD_OK || s == LSD_NO_EFFECTIVE_POINTS) {
    memcpy(f->state_point, f->x, sizeof(f->x));         // state_point = kf.get_x(), :1284
    double R[9], o[3], pl[3];
    eskf::q2R(f->x + eskf::S_ROT, R);
    eskf::mv3(R, f->x + eskf::S_OFFT, o);
    for (int i = 0; i < 3; i++) pl[i] = f->x[eskf::S_POS + i] + o[i];             // pos_lid, :1285
    const double d0 = pl[0] - f->last_pos_lid[0], d1 = pl[1] - f->last_pos_lid[1], d2 = pl[2] - f->last_pos_lid[2];
    f->travel_distance += sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    for (int i = 0; i < 3; i++) f->last_pos_lid[i] = pl[i];
  }
  return 1;
}

// what the last lsd_fastlio_main did: the lsd_lio_scan status (or LSD_IMU_INITIALIZING) and its counters
lsd_status_t lsd_fastlio_last(lsd_fastlio_t* f, int* status, lsd_lio_info_t* info) {
  if (!f) return LSD_ERR_INVALID;
  if (status) *status = f->last_status;
  if (info) *info = f->info;
  return LSD_OK;
}

// kf.get_x() / kf.get_P()
lsd_status_t lsd_fastlio_get_filter(lsd_fastlio_t* f, double* state26, double* P529) {
  if (!f) return LSD_ERR_INVALID;
  if (state26) memcpy(state26, f->x, sizeof(f->x));
  if (P529) memcpy(P529, f->P, sizeof(f->P));
  return LSD_OK;
}

static void pose16(const double* x, double* T) {   // Quaterniond(w, x, y, z).normalized().toRotationMatrix(), row-major 4x4
  double q[4] = {x[eskf::S_ROT], x[eskf::S_ROT + 1], x[eskf::S_ROT + 2], x[eskf::S_ROT + 3]};
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (n > 0.0) for (int i = 0; i < 4; i++) q[i] /= n;
  double R[9];
  eskf::q2R(q, R);
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) T[4 * r + c] = R[3 * r + c]; T[4 * r + 3] = x[eskf::S_POS + r]; }
  T[12] = T[13] = T[14] = 0.0; T[15] = 1.0;
}

// fastlio_odometry: the pose at the scan start (ImuProcess::start_state_point) and after the update (state_point)
lsd_status_t lsd_fastlio_odometry(lsd_fastlio_t* f, double* odom_start16, double* odom_end16) {
  if (!f || !odom_start16 || !odom_end16) return LSD_ERR_INVALID;
  double xs[26], man = 0.0;
  default_state(xs, nullptr);
  if (f->imu) lsd_imu_get_start_state(f->imu, xs, &man);
  pose16(xs, odom_start16);
  pose16(f->state_point, odom_end16);
  return LSD_OK;
}

// fastlio_state: 20 doubles = start_state_point pos, rot (x y z w), vel, ba, bg, grav, then mean_acc_norm
lsd_status_t lsd_fastlio_state(lsd_fastlio_t* f, double* out20) {
  if (!f || !out20) return LSD_ERR_INVALID;
  double xs[26], man = 0.0;
  default_state(xs, nullptr);
  if (f->imu) lsd_imu_get_start_state(f->imu, xs, &man);
  for (int i = 0; i < 3; i++) out20[i] = xs[eskf::S_POS + i];
  for (int i = 0; i < 4; i++) out20[3 + i] = xs[eskf::S_ROT + i];
  for (int i = 0; i < 3; i++) { out20[7 + i] = xs[eskf::S_VEL + i]; out20[10 + i] = xs[eskf::S_BA + i]; out20[13 + i] = xs[eskf::S_BG + i]; out20[16 + i] = xs[eskf::S_GRAV + i]; }
  out20[19] = man;
  return LSD_OK;
}

int lsd_fastlio_is_init(lsd_fastlio_t* f) {
  int flag = 0;
  if (f && f->imu) lsd_imu_is_init(f->imu, &flag);
  return flag;
}

lsd_lio_t* lsd_fastlio_lio(lsd_fastlio_t* f) { return f ? f->lio : nullptr; }

}  // extern "C"
