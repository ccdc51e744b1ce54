// slam_wrapper — the hot-path subset of the reference's pybind11 module of the same name
// (slam/src/slam_wrapper.cpp:192-324), built on liblsdreg's C ABI (include/lsdreg.h).
//
//   pointcloud_align(source_point f32[N,4], target_point f32[M,4], guess f32[4,4]) -> f32[4,4]
//       slam_wrapper.cpp:241-242 / graph_utils.cpp:20-46; the one matcher entry Python calls directly
//       (map_manager.py:190-192 `keyframe_align`).  Same argument names, same 50 m guess guard, same
//       solver settings (eps 1e-2, 64 iterations, 20 neighbours, PCL GICP's 5 m correspondence gate);
//       the GICP cost is minimised by fast_gicp's LM instead of PCL's BFGS.
//   registration_align(method, source, target, guess, max_corr, max_process_time_us)
//       -> (T f32[4,4], converged, fitness): what every C++ caller does with the object handed out by
//       select_registration_method (registrations.hpp:15-16): setInputTarget/Source, align,
//       hasConverged, getFinalTransformation, getFitnessScore — exposed for Python tests and tools.
//   dump_keyframe(directory, stamp, id, points_input f32[N,4], pose_input f32[4,4])
//       slam_wrapper.cpp:273-276 / graph_utils.cpp:123-131 -> KeyFrame::save: cloud.pcd + data in `directory`
//       (map_manager.py:286-288 calls it once per key frame when a map is saved).
//   init_slam / setup_slam / deinit_slam / set_ins_external_param / set_imu_external_param / process / update_odom
//       slam_wrapper.cpp:6-33,62-135,194-224 — the per-frame entries slam/slam.py calls (slam.py:49-87,234-244), for
//       mode "mapping" with method "FastLIO" on top of the C++ seam 1 (lsd_fastlio_*, fastlio.cpp:9-16).  What
//       SLAM::run (slam.cpp:273-366) and HDL_FastLIO (fastlio.cpp:153-277) do between the dicts and fastlio_*: IMU rows
//       (t us, gyr deg/s, acc g) -> fastlio_imu_enqueue units, the lidar cloud moved to the INS frame by the static
//       transform, fastlio_pcl_enqueue, fastlio_main, fastlio_odometry conjugated with the IMU-INS extrinsic, the pose
//       dict with Eigen's eulerAngles(2, 0, 1) heading convention.  The pose graph behind enqueue_graph (g2o, loop
//       closure, floor / GNSS edges) is out of scope: key frames are chosen by hdl_graph_slam's
//       translation / rotation thresholds on the odometry, filtered like SLAM::runMappingThread (slam.cpp:398-410), and
//       update_odom()'s "odoms" are the key frames' odometry poses (what graph_update_odom returns before any loop
//       closure moved them).  RTKM / localization modes raise.
// The remaining ~40 functions of the reference module (graph editing / export / camera / destination) drive
// subsystems that are out of scope and are not provided here.
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>

#include <pybind11/stl.h>

#include <cmath>
#include <cstring>
#include <deque>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "lsdreg.h"

namespace py = pybind11;
using farray = py::array_t<float, py::array::c_style | py::array::forcecast>;

namespace {

struct Reg {
  lsd_reg_t* h = nullptr;
  Reg(int kind, const lsd_reg_params_t& p) { (void)kind; if (lsd_reg_create(&h, &p) != LSD_OK) throw std::runtime_error(lsd_last_error()); }
  ~Reg() { lsd_reg_destroy(h); }
};

void check_cloud(const farray& a, const char* name) {
  if (a.ndim() != 2 || a.shape(1) != 4 || a.shape(0) < 1) throw std::invalid_argument(std::string(name) + " must be float32 [N,4]");
}
void check(lsd_status_t s) { if (s < 0) throw std::runtime_error(lsd_last_error()); }

farray to_numpy(const float* T) {
  farray out({4, 4});
  std::copy(T, T + 16, out.mutable_data());
  return out;
}

farray pointcloud_align(farray source_point, farray target_point, farray guess) {
  check_cloud(source_point, "source_point"); check_cloud(target_point, "target_point");
  if (guess.ndim() != 2 || guess.shape(0) != 4 || guess.shape(1) != 4) throw std::invalid_argument("guess must be [4,4]");
  float g[16];
  std::copy(guess.data(), guess.data() + 16, g);
  const double dist = std::sqrt((double)g[3] * g[3] + (double)g[7] * g[7] + (double)g[11] * g[11]);
  if (dist >= 50.0) g[3] = g[7] = g[11] = 0.f;                          // graph_utils.cpp:24-32
  lsd_reg_params_t p;
  lsd_reg_default_params(&p, LSD_REG_GICP);
  p.transformation_epsilon = 1e-2; p.max_iterations = 64; p.k_correspondences = 20;   // graph_utils.cpp:35-37
  p.max_corr_dist = 5.0;                                                // pcl::GeneralizedIterativeClosestPoint default gate
  Reg r(LSD_REG_GICP, p);
  float T[16];
  int conv = 0, it = 0;
  {
    py::gil_scoped_release nogil;
    check(lsd_reg_set_source(r.h, source_point.data(), (int)source_point.shape(0)));
    check(lsd_reg_set_target(r.h, target_point.data(), (int)target_point.shape(0)));
    check(lsd_reg_align(r.h, g, T, &conv, &it));
  }
  return to_numpy(T);
}

py::tuple registration_align(const std::string& method, farray source, farray target, farray guess, double max_corr,
                             long long max_process_time_us) {
  check_cloud(source, "source"); check_cloud(target, "target");
  if (guess.ndim() != 2 || guess.shape(0) != 4 || guess.shape(1) != 4) throw std::invalid_argument("guess must be [4,4]");
  int kind;
  if (method == "NDT_CUDA") kind = LSD_REG_NDT_P2D;                      // registrations.cpp:107-118
  else if (method == "FAST_GICP") kind = LSD_REG_GICP;                   // registrations.cpp:33-41
  else if (method == "FAST_VGICP" || method == "FAST_VGICP_CUDA") kind = LSD_REG_VGICP;  // registrations.cpp:44-66
  else throw std::invalid_argument("unknown registration method " + method);
  lsd_reg_params_t p;
  lsd_reg_default_params(&p, kind);
  if (max_corr > 0) p.max_corr_dist = max_corr;
  p.max_process_time_us = max_process_time_us;
  Reg r(kind, p);
  float T[16];
  int conv = 0, it = 0;
  double fit = 0;
  {
    py::gil_scoped_release nogil;
    check(lsd_reg_set_target(r.h, target.data(), (int)target.shape(0)));
    check(lsd_reg_set_source(r.h, source.data(), (int)source.shape(0)));
    check(lsd_reg_align(r.h, guess.data(), T, &conv, &it));
    check(lsd_reg_fitness(r.h, nullptr, 25.0, &fit));
  }
  return py::make_tuple(to_numpy(T), conv != 0, fit);
}

void dump_keyframe(const std::string& directory, uint64_t stamp, int id, farray points_input, farray pose_input) {
  if (points_input.ndim() != 2 || points_input.shape(1) != 4) throw std::invalid_argument("points_input must be float32 [N,4]");
  if (pose_input.ndim() != 2 || pose_input.shape(0) != 4 || pose_input.shape(1) != 4) throw std::invalid_argument("pose_input must be [4,4]");
  double pose[16];
  for (int i = 0; i < 16; i++) pose[i] = pose_input.data()[i];   // numpy_to_eigen: float -> double, py_utils.cpp
  const float* pts = points_input.data();
  const int n = (int)points_input.shape(0);
  lsd_status_t s;
  { py::gil_scoped_release release; s = lsd_keyframe_save(directory.c_str(), stamp, id, pts, n, pose); }
  check(s);
}

// ------------------------------------------------------------------ the per-frame entries (FastLIO mapping mode)
using darray = py::array_t<double, py::array::c_style | py::array::forcecast>;
struct M4 { double m[16]; };
M4 eye4() { M4 r; std::memset(r.m, 0, sizeof(r.m)); r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0; return r; }
M4 mul4(const M4& a, const M4& b) {
  M4 r;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) { double v = 0; for (int k = 0; k < 4; k++) v += a.m[4 * i + k] * b.m[4 * k + j]; r.m[4 * i + j] = v; }
  return r;
}
M4 inv_rigid(const M4& a) {   // [R t; 0 1]^-1 = [R^T  -R^T t; 0 1]
  M4 r = eye4();
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) r.m[4 * i + j] = a.m[4 * j + i];
  for (int i = 0; i < 3; i++) r.m[4 * i + 3] = -(r.m[4 * i] * a.m[3] + r.m[4 * i + 1] * a.m[7] + r.m[4 * i + 2] * a.m[11]);
  return r;
}
// getTransformFromRPYT (slam/common/slam_utils.cpp:89-96): T = trans * Rz(yaw) * Rx(pitch) * Ry(roll), degrees
M4 from_rpyt(double x, double y, double z, double yaw, double pitch, double roll) {
  const double d = M_PI / 180.0, cy = std::cos(yaw * d), sy = std::sin(yaw * d), cp = std::cos(pitch * d), sp = std::sin(pitch * d),
               cr = std::cos(roll * d), sr = std::sin(roll * d);
  M4 Rz = eye4(), Rx = eye4(), Ry = eye4();
  Rz.m[0] = cy; Rz.m[1] = -sy; Rz.m[4] = sy; Rz.m[5] = cy;
  Rx.m[5] = cp; Rx.m[6] = -sp; Rx.m[9] = sp; Rx.m[10] = cp;
  Ry.m[0] = cr; Ry.m[2] = sr; Ry.m[8] = -sr; Ry.m[10] = cr;
  M4 T = mul4(Rz, mul4(Rx, Ry));
  T.m[3] = x; T.m[7] = y; T.m[11] = z;
  return T;
}
// getRPYTfromTransformFrom (slam_utils.cpp:98-110): Eigen's eulerAngles(2, 0, 1) (Geometry/EulerAngles.h:36-110 of the vendored
// Eigen: a0 = 2, a1 = 0 -> odd = 0, i = 2, j = 0, k = 1; first angle in [0, pi]), in degrees
void to_rpyt(const M4& T, double* yaw, double* pitch, double* roll) {
  auto c = [&](int r, int col) { return T.m[4 * r + col]; };
  double r0 = std::atan2(c(0, 1), c(1, 1)), r1;
  const double c2 = std::sqrt(c(2, 2) * c(2, 2) + c(2, 0) * c(2, 0));
  if (r0 > 0.0) { r0 -= M_PI; r1 = std::atan2(-c(2, 1), -c2); } else r1 = std::atan2(-c(2, 1), c2);
  const double s1 = std::sin(r0), c1 = std::cos(r0);
  const double r2 = std::atan2(s1 * c(1, 2) - c1 * c(0, 2), c1 * c(0, 0) - s1 * c(1, 0));
  const double k = 180.0 / M_PI;
  *yaw = -r0 * k; *pitch = -r1 * k; *roll = -r2 * k;      // (!odd) res = -res
}
darray m4_to_numpy(const M4& T) { darray out({4, 4}); std::copy(T.m, T.m + 16, out.mutable_data()); return out; }

struct KeyFrameOut { farray points; M4 pose; uint64_t stamp; };
struct SlamState {
  std::string mode, method, lidar;
  bool use_imu = false, use_gps = false;
  double resolution = 0.2, key_dist = 4.0, key_deg = 20.0, frame_range = 1e9, scan_period = 0.1;   // SLAM::setParams, slam.cpp:93-103
  M4 static_T = eye4(), imu_T = eye4(), imu_ins = eye4();
  lsd_fastlio_t* flio = nullptr;
  bool key_first = true; M4 prev_key = eye4();
  std::deque<KeyFrameOut> keyframes;
  std::map<int, M4> odoms;
  ~SlamState() { if (flio) lsd_fastlio_destroy(flio); }
};
std::unique_ptr<SlamState> g_slam;

SlamState& slam() { if (!g_slam) throw std::runtime_error("init_slam was not called"); return *g_slam; }

// init_slam (slam_wrapper.cpp:6-14): SLAM(mode, method) + setSensors (HDL_FastLIO::setSensors, fastlio.cpp:118-151: RTK and IMU
// kept, cameras kept, the FIRST "n-Name" lidar is the one registered; without an IMU only RTK / IMU survive) + setParams
py::list init_slam(const std::string mode, const std::string map_path, const std::string method, py::list& sensor_input, double resolution,
                   float dist_threshold, float degree_threshold, float frame_range) {
  (void)map_path;
  if (method != "FastLIO") throw std::invalid_argument("slam_wrapper (lsdreg): method \"" + method + "\" is outside the hot path; only \"FastLIO\" is provided");
  if (mode != "mapping" && mode != "online" && mode != "offline") throw std::invalid_argument("slam_wrapper (lsdreg): only the mapping run mode is provided (localization is out of scope)");
  g_slam.reset(new SlamState());
  SlamState& s = *g_slam;
  s.mode = mode; s.method = method; s.resolution = resolution; s.key_dist = dist_threshold; s.key_deg = degree_threshold; s.frame_range = frame_range;
  std::vector<std::string> in, out;
  for (auto h : sensor_input) in.push_back(py::cast<std::string>(h));
  for (auto& n : in) { if (n == "RTK") { out.push_back(n); s.use_gps = true; } else if (n == "IMU") { out.push_back(n); s.use_imu = true; } }
  if (s.use_imu) {
    for (auto& n : in) {
      if (n == "RTK" || n == "IMU") continue;
      out.push_back(n);
      if (!(n.length() < 2 || n[1] != '-') && s.lidar.empty()) s.lidar = n;
    }
  }
  return py::cast(out);
}
void set_ins_external_param(double x, double y, double z, double yaw, double pitch, double roll) { slam().static_T = from_rpyt(x, y, z, yaw, pitch, roll); }
void set_imu_external_param(double x, double y, double z, double yaw, double pitch, double roll) { slam().imu_T = from_rpyt(x, y, z, yaw, pitch, roll); }

// setup_slam (slam_wrapper.cpp:16-18) -> SLAM::setup -> HDL_FastLIO::init (fastlio.cpp:153-171): fastlio_init with the IMU-INS extrinsic
bool setup_slam() {
  SlamState& s = slam();
  if (s.lidar.empty()) throw std::runtime_error("setup_slam: FastLIO needs \"IMU\" and one \"n-Name\" lidar among the sensors (fastlio.cpp:118-151)");
  s.imu_ins = mul4(s.imu_T, inv_rigid(s.static_T));
  const double extT[3] = {s.imu_ins.m[3], s.imu_ins.m[7], s.imu_ins.m[11]};
  const double extR[9] = {s.imu_ins.m[0], s.imu_ins.m[1], s.imu_ins.m[2], s.imu_ins.m[4], s.imu_ins.m[5], s.imu_ins.m[6], s.imu_ins.m[8], s.imu_ins.m[9], s.imu_ins.m[10]};
  if (s.flio) { lsd_fastlio_destroy(s.flio); s.flio = nullptr; }
  lsd_status_t st;
  { py::gil_scoped_release nogil; st = lsd_fastlio_create(&s.flio, extT, extR, 1, -1, s.scan_period, 1); }
  check(st);
  return true;
}
void deinit_slam() { g_slam.reset(nullptr); }

// process (slam_wrapper.cpp:62-112) -> SLAM::run (slam.cpp:273-366) -> HDL_FastLIO::feedImuData / feedPointData / getPose
py::dict process(py::dict& points, py::dict& points_attr, py::dict& image_dict, py::dict& image_stream_dict, py::dict& image_param,
                 py::dict& rtk_dict, darray imu_list, uint64_t timestamp) {
  (void)image_dict; (void)image_stream_dict; (void)image_param;
  SlamState& s = slam();
  if (!s.flio) throw std::runtime_error("process: setup_slam was not called");
  if (!points.contains(s.lidar.c_str())) throw std::invalid_argument("process: points has no entry for lidar \"" + s.lidar + "\"");
  farray cloud = py::cast<farray>(points[s.lidar.c_str()]);
  py::dict attr = py::cast<py::dict>(points_attr[s.lidar.c_str()]);
  farray pattr = py::cast<farray>(attr["points_attr"]);
  const uint64_t header_stamp = py::cast<uint64_t>(attr["timestamp"]);
  if (cloud.ndim() != 2 || cloud.shape(1) < 4) throw std::invalid_argument("points must be float32 [N,4]");
  if (pattr.ndim() != 2 || pattr.shape(0) != cloud.shape(0) || pattr.shape(1) < 2) throw std::invalid_argument("points_attr must be float32 [N,2] = (stamp us, id)");
  const int n = (int)cloud.shape(0), cs = (int)cloud.shape(1), as = (int)pattr.shape(1);
  const int n_imu = imu_list.ndim() == 2 ? (int)imu_list.shape(0) : 0;
  if (n_imu && imu_list.shape(1) < 7) throw std::invalid_argument("imu_list must be float64 [M,7] = (t us, gyr deg/s, acc g)");
  // pcl::transformPointCloud(cloud, cloud, mStaticTrans) (slam_base.h:83-85): Matrix4d applied to float points
  std::vector<float> xyzi((size_t)n * 4);
  std::vector<uint32_t> stamps((size_t)n);
  const float* cp = cloud.data(); const float* ap = pattr.data();
  for (int i = 0; i < n; i++) {
    const double x = cp[(size_t)cs * i], y = cp[(size_t)cs * i + 1], z = cp[(size_t)cs * i + 2];
    const double* T = s.static_T.m;
    xyzi[4 * (size_t)i + 0] = (float)(T[0] * x + T[1] * y + T[2] * z + T[3]);
    xyzi[4 * (size_t)i + 1] = (float)(T[4] * x + T[5] * y + T[6] * z + T[7]);
    xyzi[4 * (size_t)i + 2] = (float)(T[8] * x + T[9] * y + T[10] * z + T[11]);
    xyzi[4 * (size_t)i + 3] = cp[(size_t)cs * i + 3];
    stamps[i] = (uint32_t)ap[(size_t)as * i];          // PointAttr::stamp: us relative to the header stamp (py_utils.cpp:165)
  }
  M4 odom_s = eye4(), odom_e = eye4();
  int consumed = 0;
  lsd_status_t st = LSD_OK;
  {
    py::gil_scoped_release nogil;                      // slam_wrapper.cpp:84-90
    if (s.use_imu) {
      const double* ip = imu_list.data();
      const int is = n_imu ? (int)imu_list.shape(1) : 7;
      for (int i = 0; i < n_imu && st >= 0; i++) {     // numpy_to_imu (py_utils.cpp:244-258): deg/s -> rad/s, g -> m/s^2, us -> s
        const double* r = ip + (size_t)is * i;
        const double gyr[3] = {r[1] / 180.0 * M_PI, r[2] / 180.0 * M_PI, r[3] / 180.0 * M_PI};
        const double acc[3] = {r[4] * 9.81, r[5] * 9.81, r[6] * 9.81};
        st = lsd_fastlio_imu_enqueue(s.flio, r[0] / 1000000.0, gyr, acc);
      }
    }
    if (st >= 0) st = lsd_fastlio_pcl_enqueue(s.flio, xyzi.data(), stamps.data(), n, header_stamp);
    if (st >= 0) {
      consumed = lsd_fastlio_main(s.flio);             // the reference polls fastlio_main on its own thread (fastlio.cpp:263-277); getPose waits for it
      if (consumed > 0) {
        double a[16], b[16];
        st = lsd_fastlio_odometry(s.flio, a, b);
        M4 A, B; std::copy(a, a + 16, A.m); std::copy(b, b + 16, B.m);
        const M4 inv = inv_rigid(s.imu_ins);
        odom_s = mul4(inv, mul4(A, s.imu_ins));        // fastlio.cpp:268-269
        odom_e = mul4(inv, mul4(B, s.imu_ins));
      }
    }
  }
  if (consumed < 0) check((lsd_status_t)consumed);
  check(st);
  // key frames: translation / rotation thresholds of hdl_graph_slam's KeyframeUpdater on the odometry, then the filters of
  // SLAM::runMappingThread (slam.cpp:398-410): RadiusOutlierRemoval(1.0 m, 3) + pointsDistanceFilter(0, key_frame_range)
  if (consumed > 0) {
    bool is_key = s.key_first;
    if (!s.key_first) {
      const M4 d = mul4(inv_rigid(s.prev_key), odom_s);
      const double dx = std::sqrt(d.m[3] * d.m[3] + d.m[7] * d.m[7] + d.m[11] * d.m[11]);
      const double ct = std::max(-1.0, std::min(1.0, (d.m[0] + d.m[5] + d.m[10] - 1.0) * 0.5));
      const double da = std::acos(ct) / M_PI * 180.0;
      is_key = !(dx < s.key_dist && da < s.key_deg);
    }
    if (is_key) {
      s.key_first = false; s.prev_key = odom_s;
      farray kept({(py::ssize_t)n, (py::ssize_t)4});
      int n_out = 0;
      { py::gil_scoped_release nogil; st = lsd_keyframe_filter(xyzi.data(), n, 1.0f, 3, 0.0f, (float)s.frame_range, kept.mutable_data(), &n_out); }
      check(st);
      farray out({(py::ssize_t)n_out, (py::ssize_t)4});
      std::copy(kept.data(), kept.data() + (size_t)n_out * 4, out.mutable_data());
      const int id = (int)s.odoms.size();
      s.odoms[id] = odom_s;
      s.keyframes.push_back({out, odom_s, header_stamp});
    }
  }
  // the pose dict (slam_wrapper.cpp:92-112; SLAM::run's mapping branch, slam.cpp:339-364)
  double yaw, pitch, roll;
  M4 T = odom_s;
  to_rpyt(T, &yaw, &pitch, &roll);
  double heading;
  if (std::abs(roll) >= 90.0 || std::abs(pitch) >= 90.0) {
    T = from_rpyt(T.m[3], T.m[7], T.m[11], -yaw, pitch, roll);
    to_rpyt(T, &yaw, &pitch, &roll);
    heading = yaw;
  } else {
    heading = -yaw;
  }
  if (heading < 0) heading += 360.0;
  auto num = [&](const char* k) { return rtk_dict.contains(k) ? py::cast<double>(rtk_dict[k]) : 0.0; };
  py::dict pose;
  pose["latitude"] = num("latitude"); pose["longitude"] = num("longitude"); pose["altitude"] = num("altitude");
  pose["heading"] = heading; pose["pitch"] = pitch; pose["roll"] = roll;
  pose["Ve"] = 0; pose["Vn"] = 0; pose["Vu"] = 0;
  pose["Status"] = rtk_dict.contains("Status") ? py::cast<int>(rtk_dict["Status"]) : 0;
  pose["state"] = "Mapping";
  pose["timestamp"] = header_stamp;
  pose["odom_matrix"] = m4_to_numpy(odom_s);
  py::dict data;
  data["frame_start_timestamp"] = timestamp;
  data["pose"] = pose;
  data["slam_valid"] = true;
  return data;
}

// update_odom (slam_wrapper.cpp:114-135): flush the key-frame queue, return the key frames' poses
py::dict update_odom() {
  SlamState& s = slam();
  py::list l;
  while (!s.keyframes.empty()) {
    KeyFrameOut& k = s.keyframes.front();
    py::dict d;
    d["points"] = k.points;
    d["image"] = py::dict();
    d["pose"] = m4_to_numpy(k.pose);
    d["stamp"] = k.stamp;
    l.append(d);
    s.keyframes.pop_front();
  }
  py::dict od;
  for (auto& kv : s.odoms) od[py::int_(kv.first)] = m4_to_numpy(kv.second);
  py::dict out;
  out["odoms"] = od;
  out["keyframes"] = l;
  return out;
}

}  // namespace

PYBIND11_MODULE(slam_wrapper, m) {
  m.doc() = "mapping python interface (lsdreg hot-path subset)";
  m.def("pointcloud_align", &pointcloud_align, "pointcloud align", py::arg("source_point"), py::arg("target_point"), py::arg("guess"));
  m.def("registration_align", &registration_align, "select_registration_method(method) + align + fitness", py::arg("method"),
        py::arg("source"), py::arg("target"), py::arg("guess"), py::arg("max_corr") = 0.0, py::arg("max_process_time_us") = 0LL);
  m.def("dump_keyframe", &dump_keyframe, "dump keyframe", py::arg("directory"), py::arg("stamp"), py::arg("id"), py::arg("points_input"),
        py::arg("pose_input"));
  m.def("init_slam", &init_slam, "init slam", py::arg("mode"), py::arg("map_path"), py::arg("method"), py::arg("sensor_input"), py::arg("resolution"),
        py::arg("dist_threshold"), py::arg("degree_threshold"), py::arg("frame_range"));
  m.def("setup_slam", &setup_slam, "setup slam");
  m.def("deinit_slam", &deinit_slam, "deinit slam");
  m.def("set_ins_external_param", &set_ins_external_param, "set ins external param", py::arg("x"), py::arg("y"), py::arg("z"), py::arg("yaw"),
        py::arg("pitch"), py::arg("roll"));
  m.def("set_imu_external_param", &set_imu_external_param, "set imu external param", py::arg("x"), py::arg("y"), py::arg("z"), py::arg("yaw"),
        py::arg("pitch"), py::arg("roll"));
  m.def("process", &process, "process", py::arg("points"), py::arg("points_attr"), py::arg("image_dict"), py::arg("image_stream_dict"),
        py::arg("image_param"), py::arg("rtk_dict"), py::arg("imu_list"), py::arg("timestamp"));
  m.def("update_odom", &update_odom, "update odom");
  m.def("lsd_version", []() { return std::string(lsd_version()); });
}
