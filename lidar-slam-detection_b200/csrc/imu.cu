// imu.cu — IMU forward propagation (host, double) and per-point undistortion (device).  Row N1.
//
// Product code, written from the reference's algorithm — NOT derived from oracle/.
// Replaces (paths relative to /root/reference/slam/mapping/fastlio):
//   ImuProcess::IMU_init / UndistortPcl / Process   src/IMU_Processing.hpp:167-450
//   esekf::predict                                  include/IKFoM_toolkit/esekfom/esekfom.hpp:279-383
//   get_f / df_dx / df_dw / process_noise_cov       include/use-ikfom.hpp:36-88
//   Exp(ang_vel, dt)                                include/so3_math.h:36-58
// The <= ~40 filter predictions per scan stay on the host (23x23 algebra, like the update); the
// backward propagation of ~100 k points is one kernel: every thread finds its IMU segment and applies
//   P = R_LI^T ( R_end^T ( R_i (R_LI p + t_LI) + T_ei ) - t_LI )            (IMU_Processing.hpp:390)
// in double.  The reference sorts the cloud by time first; nothing downstream of this library depends
// on the order (the voxel grid sums exactly), so points stay in input order.
#include <math.h>

#include <algorithm>
#include <vector>

#include "eskf.hpp"
#include "lsd_common.cuh"

namespace lsd {

// ------------------------------------------------------------------ esekf::predict (host)
namespace eskf {

// x (+)= f * dt on the flattened 24-vector (MTK_BUILD_MANIFOLD oplus; SOn.hpp:242-245, S2.hpp:129-134)
inline void state_oplus(double* x, const double* f, double dt) {
  double q[4], R[9];
  for (int i = 0; i < 3; i++) x[S_POS + i] += f[0 + i] * dt;
  mtk_exp(f + 3, dt / 2, q); qmul(x + S_ROT, q, x + S_ROT);
  mtk_exp(f + 6, dt / 2, q); qmul(x + S_OFFR, q, x + S_OFFR);
  for (int i = 0; i < 3; i++) x[S_OFFT + i] += f[9 + i] * dt;
  for (int i = 0; i < 3; i++) x[S_VEL + i] += f[12 + i] * dt;
  for (int i = 0; i < 3; i++) x[S_BG + i] += f[15 + i] * dt;
  for (int i = 0; i < 3; i++) x[S_BA + i] += f[18 + i] * dt;
  mtk_exp(f + 21, dt / 2, q); q2R(q, R); mv3(R, x + S_GRAV, x + S_GRAV);
}

// One prediction step: state and covariance, Q is 12x12 row-major (ng, na, nbg, nba).
inline void predict(double* x, double* P, double dt, const double* Q, const double* acc, const double* gyro) {
  double R[9];
  q2R(x + S_ROT, R);
  // get_f (use-ikfom.hpp:49-61), flattened DIM layout: pos 0, rot 3, offR 6, offT 9, vel 12, bg 15, ba 18, grav 21
  double f[24] = {0};
  double am[3] = {acc[0] - x[S_BA], acc[1] - x[S_BA + 1], acc[2] - x[S_BA + 2]}, Ra[3];
  mv3(R, am, Ra);
  for (int i = 0; i < 3; i++) { f[i] = x[S_VEL + i]; f[3 + i] = gyro[i] - x[S_BG + i]; f[12 + i] = Ra[i] + x[S_GRAV + i]; }
  // df_dx (24 x 23, :63-80) and df_dw (24 x 12, :83-91): only a handful of non-zero blocks
  static thread_local std::vector<double> fx(24 * 23), fw(24 * 12), fxf(23 * 23), fwf(23 * 12), F1(23 * 23), T(23 * 23), W(23 * 12), WQ(23 * 12);
  std::fill(fx.begin(), fx.end(), 0.0); std::fill(fw.begin(), fw.end(), 0.0);
  for (int i = 0; i < 3; i++) fx[(0 + i) * 23 + 12 + i] = 1.0;
  double Ha[9], RH[9], zero2[2] = {0, 0}, Mx[6];
  hat(am, Ha); mm3(R, Ha, RH);
  s2_Mx(x + S_GRAV, zero2, Mx);
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) { fx[(12 + i) * 23 + 3 + j] = -RH[3 * i + j]; fx[(12 + i) * 23 + 18 + j] = -R[3 * i + j]; }
    for (int j = 0; j < 2; j++) fx[(12 + i) * 23 + 21 + j] = Mx[2 * i + j];
    fx[(3 + i) * 23 + 15 + i] = -1.0;
    for (int j = 0; j < 3; j++) fw[(12 + i) * 12 + 3 + j] = -R[3 * i + j];
    fw[(3 + i) * 12 + i] = -1.0; fw[(15 + i) * 12 + 6 + i] = 1.0; fw[(18 + i) * 12 + 9 + i] = 1.0;
  }
  double xb[S_DIM];
  memcpy(xb, x, sizeof(xb));
  state_oplus(x, f, dt);
  // F_x1 = I with the SO3 / S2 "exp" blocks; the reference evaluates them with scalar_type(1/2) == 0
  // (integer division, esekfom.hpp:312,344), which makes every exp the identity.
  for (int i = 0; i < 23 * 23; i++) F1[i] = (i % 24 == 0) ? 1.0 : 0.0;
  const int vect_idx[5] = {0, 9, 12, 15, 18};
  for (int v = 0; v < 5; v++) for (int j = 0; j < 3; j++) {
    memcpy(&fxf[(vect_idx[v] + j) * 23], &fx[(vect_idx[v] + j) * 23], 23 * 8);
    memcpy(&fwf[(vect_idx[v] + j) * 12], &fw[(vect_idx[v] + j) * 12], 12 * 8);
  }
  const int so3_idx[2] = {3, 6};
  for (int s = 0; s < 2; s++) {
    const int idx = so3_idx[s];
    double seg[3] = {-f[idx] * dt, -f[idx + 1] * dt, -f[idx + 2] * dt}, A[9];
    A_matrix(seg, A);
    for (int c = 0; c < 23; c++) for (int i = 0; i < 3; i++)
      fxf[(idx + i) * 23 + c] = A[3 * i] * fx[idx * 23 + c] + A[3 * i + 1] * fx[(idx + 1) * 23 + c] + A[3 * i + 2] * fx[(idx + 2) * 23 + c];
    for (int c = 0; c < 12; c++) for (int i = 0; i < 3; i++)
      fwf[(idx + i) * 12 + c] = A[3 * i] * fw[idx * 12 + c] + A[3 * i + 1] * fw[(idx + 1) * 12 + c] + A[3 * i + 2] * fw[(idx + 2) * 12 + c];
  }
  {
    const int idx = 21;
    double seg[3] = {f[21] * dt, f[22] * dt, f[23] * dt}, Nx[6], Mb[6], Hb[9], A[9], At[9], NH[6], tmp[6];
    s2_Nx_yy(x + S_GRAV, Nx);
    s2_Mx(xb + S_GRAV, zero2, Mb);
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++)
      F1[(idx + i) * 23 + idx + j] = Nx[3 * i] * Mb[j] + Nx[3 * i + 1] * Mb[2 + j] + Nx[3 * i + 2] * Mb[4 + j];
    hat(xb + S_GRAV, Hb); A_matrix(seg, A); tr3(A, At);
    for (int i = 0; i < 2; i++) for (int j = 0; j < 3; j++) NH[3 * i + j] = Nx[3 * i] * Hb[j] + Nx[3 * i + 1] * Hb[3 + j] + Nx[3 * i + 2] * Hb[6 + j];
    for (int i = 0; i < 2; i++) for (int j = 0; j < 3; j++) tmp[3 * i + j] = -(NH[3 * i] * At[j] + NH[3 * i + 1] * At[3 + j] + NH[3 * i + 2] * At[6 + j]);
    for (int c = 0; c < 23; c++) for (int i = 0; i < 2; i++)
      fxf[(idx + i) * 23 + c] = tmp[3 * i] * fx[21 * 23 + c] + tmp[3 * i + 1] * fx[22 * 23 + c] + tmp[3 * i + 2] * fx[23 * 23 + c];
    for (int c = 0; c < 12; c++) for (int i = 0; i < 2; i++)
      fwf[(idx + i) * 12 + c] = tmp[3 * i] * fw[21 * 12 + c] + tmp[3 * i + 1] * fw[22 * 12 + c] + tmp[3 * i + 2] * fw[23 * 12 + c];
  }
  for (int i = 0; i < 23 * 23; i++) F1[i] += fxf[i] * dt;
  // P = F P F^T + (dt W) Q (dt W)^T
  for (int i = 0; i < 23; i++) for (int j = 0; j < 23; j++) { double s = 0; for (int k = 0; k < 23; k++) s += F1[i * 23 + k] * P[k * 23 + j]; T[i * 23 + j] = s; }
  for (int i = 0; i < 23 * 12; i++) W[i] = dt * fwf[i];
  for (int i = 0; i < 23; i++) for (int j = 0; j < 12; j++) { double s = 0; for (int k = 0; k < 12; k++) s += W[i * 12 + k] * Q[k * 12 + j]; WQ[i * 12 + j] = s; }
  for (int i = 0; i < 23; i++) for (int j = 0; j < 23; j++) {
    double s = 0;
    for (int k = 0; k < 23; k++) s += T[i * 23 + k] * F1[j * 23 + k];
    double q = 0;
    for (int k = 0; k < 12; k++) q += WQ[i * 12 + k] * W[j * 12 + k];
    P[i * 23 + j] = s + q;
  }
}

}  // namespace eskf

// ------------------------------------------------------------------ device: backward propagation
struct ImuPoseDev { double t, acc[3], gyr[3], vel[3], pos[3], rot[9]; };  // Pose6D, common_lib.h:37-44
struct ImuEndDev { double Rend[9], pend[3], RL[9], tL[3]; };
constexpr int kMaxImuPoses = 256;

// so3_math.h:36-58
__device__ __forceinline__ void exp_rodrigues(const double* w, double dt, double* E) {
  const double n = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  E[0] = E[4] = E[8] = 1.0; E[1] = E[2] = E[3] = E[5] = E[6] = E[7] = 0.0;
  if (!(n > 0.0000001)) return;
  const double ax = w[0] / n, ay = w[1] / n, az = w[2] / n, a = n * dt, s = sin(a), c1 = 1.0 - cos(a);
  const double K[9] = {0, -az, ay, az, 0, -ax, -ay, ax, 0};
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const double kk = K[3 * i] * K[j] + K[3 * i + 1] * K[3 + j] + K[3 * i + 2] * K[6 + j];
      E[3 * i + j] += s * K[3 * i + j] + c1 * kk;
    }
}

// one application of IMU_Processing.hpp:381-395 with segment head k; p is updated in place (float fields)
__device__ __forceinline__ void undistort_segment(const ImuPoseDev* poses, int k, const ImuEndDev& e, double t, float* px, float* py, float* pz) {
  const ImuPoseDev& hd = poses[k];
  const ImuPoseDev& tl = poses[k + 1];
  const double dt = t - hd.t;
  double E[9], Ri[9];
  exp_rodrigues(tl.gyr, dt, E);
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) Ri[3 * i + j] = hd.rot[3 * i] * E[j] + hd.rot[3 * i + 1] * E[3 + j] + hd.rot[3 * i + 2] * E[6 + j];
  const double p[3] = {(double)*px, (double)*py, (double)*pz};
  double a[3], b[3], c[3];
#pragma unroll
  for (int i = 0; i < 3; i++) a[i] = e.RL[3 * i] * p[0] + e.RL[3 * i + 1] * p[1] + e.RL[3 * i + 2] * p[2] + e.tL[i];
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const double Tei = hd.pos[i] + hd.vel[i] * dt + 0.5 * tl.acc[i] * dt * dt - e.pend[i];
    b[i] = Ri[3 * i] * a[0] + Ri[3 * i + 1] * a[1] + Ri[3 * i + 2] * a[2] + Tei;
  }
#pragma unroll
  for (int i = 0; i < 3; i++) c[i] = e.Rend[i] * b[0] + e.Rend[3 + i] * b[1] + e.Rend[6 + i] * b[2] - e.tL[i];   // R_end^T b - t_LI
  *px = (float)(e.RL[0] * c[0] + e.RL[3] * c[1] + e.RL[6] * c[2]);
  *py = (float)(e.RL[1] * c[0] + e.RL[4] * c[1] + e.RL[7] * c[2]);
  *pz = (float)(e.RL[2] * c[0] + e.RL[5] * c[1] + e.RL[8] * c[2]);
}

// A point with time t belongs to the LAST segment (scanning heads from the end) whose head time is < t
// (the `for (; curvature/1000 > head->offset_time; it_pcl--)` walk over the time-sorted cloud).
__global__ void __launch_bounds__(256) imu_undistort_kernel(const float4* __restrict__ in, const float* __restrict__ time_ms, int n,
                                                            const ImuPoseDev* __restrict__ poses, int n_poses, ImuEndDev e,
                                                            float4* __restrict__ out, unsigned long long* __restrict__ first_key) {
  extern __shared__ double s_raw[];
  ImuPoseDev* sp = reinterpret_cast<ImuPoseDev*>(s_raw);
  {
    const double* src = reinterpret_cast<const double*>(poses);
    const int nd = n_poses * (int)(sizeof(ImuPoseDev) / 8);
    for (int i = threadIdx.x; i < nd; i += blockDim.x) s_raw[i] = src[i];
  }
  __syncthreads();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = __ldg(in + i);
  const float tf = __ldg(time_ms + i);
  const double t = (double)tf / 1000.0;
  int k = -1;
  for (int h = n_poses - 2; h >= 0; h--) if (t > sp[h].t) { k = h; break; }
  if (k >= 0) undistort_segment(sp, k, e, t, &p.x, &p.y, &p.z);
  out[i] = p;
  // the earliest point (ties: lowest index) is revisited by every later segment in the reference; see the fix-up kernel
  atomicMin(first_key, ((unsigned long long)__float_as_uint(fmaxf(tf, 0.0f)) << 32) | (unsigned)i);
}

// IMU_Processing.hpp:399: `if (it_pcl == begin) break;` leaves the iterator ON the first point, so each
// remaining segment whose head time is below its time compensates it again, starting from the already
// compensated coordinates.  One thread redoes that chain for the one point concerned.
__global__ void imu_first_point_kernel(const float4* __restrict__ in, const float* __restrict__ time_ms, int n,
                                       const ImuPoseDev* __restrict__ poses, int n_poses, ImuEndDev e, float4* __restrict__ out,
                                       const unsigned long long* __restrict__ first_key) {
  if (threadIdx.x != 0 || blockIdx.x != 0 || n <= 0) return;
  const unsigned i = (unsigned)(*first_key & 0xffffffffull);
  if (i >= (unsigned)n) return;
  float4 p = in[i];
  const double t = (double)time_ms[i] / 1000.0;
  for (int h = n_poses - 2; h >= 0; h--)
    if (t > poses[h].t) undistort_segment(poses, h, e, t, &p.x, &p.y, &p.z);
  out[i] = p;
}

}  // namespace lsd

// ==================================================================== host side: ImuProcess
struct lsd_imu {
  lsd_imu_params_t p{};
  int device = 0;
  cudaStream_t stream = nullptr;
  // ImuProcess members (IMU_Processing.hpp:53-85)
  double Q[144];
  double cov_acc[3], cov_gyr[3], mean_acc[3], mean_gyr[3], vel_last[3], angvel_last[3], acc_s_last[3];
  double last_imu[7];
  double last_lidar_end_time = 0.0, first_lidar_time = 0.0;
  int init_iter_num = 1;
  bool b_first_frame = true, imu_need_init = true, state_init_done = false;
  double start_state[26];        // start_state_point (IMU_Processing.hpp:58): a default state_ikfom until the first assignment
  double mean_acc_norm = 0.0;    // IMU_Processing.hpp:59,206
  std::vector<lsd::ImuPoseDev> poses;
  // device staging
  float4 *d_in = nullptr, *d_out = nullptr;
  float* d_time = nullptr;
  int cap = 0, n_out = 0;
  lsd::ImuPoseDev *h_poses = nullptr, *d_poses = nullptr;  // mapped pinned memory
  unsigned long long* d_first = nullptr;
  long long launches = 0;
};

namespace lsd {

constexpr double kGms2 = 9.81;      // common_lib.h:21
constexpr int kMaxIniCount = 100;   // IMU_Processing.hpp:26

static void imu_reset(lsd_imu* m) {  // ImuProcess::Reset, :107-120 (+ constructor defaults :87-103)
  for (int i = 0; i < 3; i++) { m->cov_acc[i] = 0.1; m->cov_gyr[i] = 0.1; m->mean_gyr[i] = 0; m->vel_last[i] = 0; m->angvel_last[i] = 0; m->acc_s_last[i] = 0; }
  m->mean_acc[0] = 0; m->mean_acc[1] = 0; m->mean_acc[2] = -1.0;
  m->imu_need_init = true; m->state_init_done = false; m->init_iter_num = 1;
  memset(m->last_imu, 0, sizeof(m->last_imu));
  m->last_lidar_end_time = 0.0;
  m->poses.clear();
}

static void set_Q(lsd_imu* m) {
  for (int i = 0; i < 3; i++) {
    m->Q[(0 + i) * 13] = m->cov_gyr[i]; m->Q[(3 + i) * 13] = m->cov_acc[i];
    m->Q[(6 + i) * 13] = m->p.b_gyr_cov; m->Q[(9 + i) * 13] = m->p.b_acc_cov;
  }
}

static double norm3(const double* v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

static void after_predict(lsd_imu* m, const double* x, const double* acc, const double* gyr) {
  double R[9], am[3], Ra[3];
  eskf::q2R(x + eskf::S_ROT, R);
  for (int i = 0; i < 3; i++) { m->angvel_last[i] = gyr[i] - x[eskf::S_BG + i]; am[i] = acc[i] - x[eskf::S_BA + i]; }
  eskf::mv3(R, am, Ra);
  for (int i = 0; i < 3; i++) m->acc_s_last[i] = Ra[i] + x[eskf::S_GRAV + i];
}

static void push_pose(lsd_imu* m, double t, const double* x) {
  ImuPoseDev ps;
  ps.t = t;
  for (int i = 0; i < 3; i++) { ps.acc[i] = m->acc_s_last[i]; ps.gyr[i] = m->angvel_last[i]; ps.vel[i] = x[eskf::S_VEL + i]; ps.pos[i] = x[eskf::S_POS + i]; }
  eskf::q2R(x + eskf::S_ROT, ps.rot);
  m->poses.push_back(ps);
}

// IMU_init, :167-235.  imu7 rows = (stamp, gyr xyz, acc xyz)
static void imu_init(lsd_imu* m, const double* imu, int n_imu, const double* ins_vel, double beg, double end, double* x, double* P) {
  if (m->b_first_frame) {
    imu_reset(m);
    m->init_iter_num = 1;
    m->b_first_frame = false;
    for (int i = 0; i < 3; i++) { m->mean_acc[i] = imu[4 + i]; m->mean_gyr[i] = imu[1 + i]; }
    m->first_lidar_time = beg;
  }
  int N = m->init_iter_num;
  for (int s = 0; s < n_imu; s++) {
    const double* a = imu + 7 * s + 4; const double* g = imu + 7 * s + 1;
    const double Nd = (double)N;
    for (int i = 0; i < 3; i++) {
      m->mean_acc[i] += (a[i] - m->mean_acc[i]) / Nd;
      m->mean_gyr[i] += (g[i] - m->mean_gyr[i]) / Nd;
      m->cov_acc[i] = m->cov_acc[i] * (Nd - 1.0) / Nd + (a[i] - m->mean_acc[i]) * (a[i] - m->mean_acc[i]) * (Nd - 1.0) / (Nd * Nd);
      m->cov_gyr[i] = m->cov_gyr[i] * (Nd - 1.0) / Nd + (g[i] - m->mean_gyr[i]) * (g[i] - m->mean_gyr[i]) * (Nd - 1.0) / (Nd * Nd);
    }
    N++;
  }
  m->init_iter_num = N;
  if (ins_vel) for (int i = 0; i < 3; i++) m->vel_last[i] = ins_vel[i];
  const double na = norm3(m->mean_acc);
  m->mean_acc_norm = na;
  if (fabs(na - 1.0) > 0.1 || norm3(m->mean_gyr) > 10.0 / 180.0 * M_PI) { m->b_first_frame = true; return; }  // "init is not stable, reset"
  double g[3] = {-m->mean_acc[0] / na * kGms2, -m->mean_acc[1] / na * kGms2, -m->mean_acc[2] / na * kGms2};
  const double gn = norm3(g);
  for (int i = 0; i < 3; i++) x[eskf::S_GRAV + i] = g[i] / gn * eskf::kS2Len;  // S2(vec): normalised to the manifold length
  for (int i = 0; i < 3; i++) { x[eskf::S_VEL + i] = m->vel_last[i]; x[eskf::S_BG + i] = 0; x[eskf::S_BA + i] = 0; x[eskf::S_OFFT + i] = m->p.ext_t[i]; }
  eskf::R2q(m->p.ext_R, x + eskf::S_OFFR);
  eskf::init_cov(P);
  memcpy(m->last_imu, imu + 7 * (n_imu - 1), sizeof(m->last_imu));
  m->last_lidar_end_time = end;
  memcpy(m->start_state, x, sizeof(m->start_state));   // start_state_point = kf_state.get_x(), :234
}

// forward propagation of UndistortPcl, :237-361: fills m->poses, leaves x / P at the scan end
static void imu_forward(lsd_imu* m, const double* imu, int n_imu, double beg, double end, double* x, double* P) {
  const double scale = kGms2 / norm3(m->mean_acc);
  if (beg > m->last_lidar_end_time) {  // predict the state at the scan start time
    double gyr[3] = {m->last_imu[1], m->last_imu[2], m->last_imu[3]};
    double acc[3] = {m->last_imu[4] * scale, m->last_imu[5] * scale, m->last_imu[6] * scale};
    set_Q(m);
    eskf::predict(x, P, beg - m->last_lidar_end_time, m->Q, acc, gyr);
    after_predict(m, x, acc, gyr);
    m->last_lidar_end_time = beg;
  }
  memcpy(m->start_state, x, sizeof(m->start_state));   // start_state_point = kf_state.get_x(), :266
  m->poses.clear();
  push_pose(m, 0.0, x);
  const double imu_end_time = imu[7 * (n_imu - 1)];
  for (int i = -1; i < n_imu - 1; i++) {  // v_imu = [last_imu, imu...]
    const double* head = i < 0 ? m->last_imu : imu + 7 * i;
    const double* tail = imu + 7 * (i + 1);
    if (tail[0] < m->last_lidar_end_time) continue;
    double gyr[3], acc[3];
    for (int k = 0; k < 3; k++) { gyr[k] = 0.5 * (head[1 + k] + tail[1 + k]); acc[k] = 0.5 * (head[4 + k] + tail[4 + k]) * scale; }
    double dt = head[0] < m->last_lidar_end_time ? tail[0] - m->last_lidar_end_time : tail[0] - head[0];
    dt = std::min(1.0, dt);
    set_Q(m);
    eskf::predict(x, P, dt, m->Q, acc, gyr);
    after_predict(m, x, acc, gyr);
    push_pose(m, tail[0] - beg, x);
  }
  const double* lastm = imu + 7 * (n_imu - 1);
  double gyr[3] = {lastm[1], lastm[2], lastm[3]}, acc[3] = {lastm[4] * scale, lastm[5] * scale, lastm[6] * scale};
  const double note = end > imu_end_time ? 1.0 : -1.0;
  const double dt = std::min(1.0, note * (end - imu_end_time));
  eskf::predict(x, P, dt, m->Q, acc, gyr);
  after_predict(m, x, acc, gyr);
  push_pose(m, end - beg, x);
  memcpy(m->last_imu, lastm, sizeof(m->last_imu));
  m->last_lidar_end_time = end;
}

static lsd_status_t imu_alloc(lsd_imu* m, int n) {
  if (n <= m->cap) return LSD_OK;
  cudaFree(m->d_in); cudaFree(m->d_out); cudaFree(m->d_time);
  m->d_in = m->d_out = nullptr; m->d_time = nullptr; m->cap = 0;
  const size_t c = (size_t)n + (size_t)n / 4 + 1024;
  LSD_CUDA(cudaMalloc((void**)&m->d_in, c * 16));
  LSD_CUDA(cudaMalloc((void**)&m->d_out, c * 16));
  LSD_CUDA(cudaMalloc((void**)&m->d_time, c * 4));
  m->cap = (int)c;
  return LSD_OK;
}

// backward propagation on the device; d_pts / d_time hold n points (device memory)
static lsd_status_t imu_backward(lsd_imu* m, const float4* d_pts, const float* d_time, int n, const double* x) {
  cudaStream_t st = m->stream;
  m->n_out = n;
  if (n <= 0) return LSD_OK;
  if (!m->p.undistort) {
    LSD_CUDA(cudaMemcpyAsync(m->d_out, d_pts, (size_t)n * 16, cudaMemcpyDeviceToDevice, st));
    return LSD_OK;
  }
  const int np = (int)m->poses.size();
  if (np < 2 || np > kMaxImuPoses) { set_error("IMU poses per scan must be in [2, %d] (got %d)", kMaxImuPoses, np); return LSD_ERR_CAPACITY; }
  LSD_CUDA(cudaStreamSynchronize(st));  // the previous scan's kernels may still read the mapped pose buffer
  memcpy(m->h_poses, m->poses.data(), (size_t)np * sizeof(ImuPoseDev));
  ImuEndDev e;
  eskf::q2R(x + eskf::S_ROT, e.Rend);
  eskf::q2R(x + eskf::S_OFFR, e.RL);
  for (int i = 0; i < 3; i++) { e.pend[i] = x[eskf::S_POS + i]; e.tL[i] = x[eskf::S_OFFT + i]; }
  LSD_CUDA(cudaMemsetAsync(m->d_first, 0xff, 8, st));
  imu_undistort_kernel<<<(n + 255) / 256, 256, (size_t)np * sizeof(ImuPoseDev), st>>>(d_pts, d_time, n, m->d_poses, np, e, m->d_out, m->d_first);
  imu_first_point_kernel<<<1, 32, 0, st>>>(d_pts, d_time, n, m->d_poses, np, e, m->d_out, m->d_first);
  LSD_CUDA(cudaGetLastError());
  m->launches += 2;
  return LSD_OK;
}

// ImuProcess::Process, :408-450
static lsd_status_t imu_process(lsd_imu* m, const double* imu, int n_imu, const double* ins_vel, double beg, double end,
                                const float4* d_pts, const float* d_time, int n, double* x, double* P, int* n_out) {
  if (n_out) *n_out = 0;
  m->n_out = 0;
  if (n_imu <= 0) return LSD_IMU_INITIALIZING;  // `if (meas.imu.empty()) return;`
  if (m->imu_need_init) {
    imu_init(m, imu, n_imu, ins_vel, beg, end, x, P);
    m->imu_need_init = true;
    memcpy(m->last_imu, imu + 7 * (n_imu - 1), sizeof(m->last_imu));
    if (m->init_iter_num > kMaxIniCount) {
      m->imu_need_init = false;
      for (int i = 0; i < 3; i++) { m->cov_acc[i] = m->p.acc_cov; m->cov_gyr[i] = m->p.gyr_cov; }
    }
    return LSD_IMU_INITIALIZING;
  }
  m->state_init_done = true;
  imu_forward(m, imu, n_imu, beg, end, x, P);
  lsd_status_t s = imu_backward(m, d_pts, d_time, n, x);
  if (s) return s;
  if (n_out) *n_out = n;
  return LSD_OK;
}

}  // namespace lsd

using namespace lsd;

extern "C" {

void lsd_imu_default_params(lsd_imu_params_t* p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->ext_R[0] = p->ext_R[4] = p->ext_R[8] = 1.0;
  p->gyr_cov = 0.1; p->acc_cov = 0.1; p->b_gyr_cov = 0.0001; p->b_acc_cov = 0.0001;  // laserMapping.cpp:1076-1079
  p->undistort = 1;
}

lsd_status_t lsd_imu_create(lsd_imu_t** out, const lsd_imu_params_t* p) {
  if (!out || !p) return LSD_ERR_INVALID;
  lsd_status_t s = ensure_device();
  if (s) return s;
  lsd_imu* m = new lsd_imu();
  m->p = *p;
  cudaGetDevice(&m->device);
  memset(m->Q, 0, sizeof(m->Q));
  for (int i = 0; i < 6; i++) m->Q[i * 13] = 0.0001;    // process_noise_cov(), use-ikfom.hpp:36-44
  for (int i = 6; i < 12; i++) m->Q[i * 13] = 0.00001;
  imu_reset(m);
  memset(m->start_state, 0, sizeof(m->start_state));   // state_ikfom(): identity rotations, grav = S2() = length * e_x
  m->start_state[eskf::S_ROT + 3] = 1.0; m->start_state[eskf::S_OFFR + 3] = 1.0; m->start_state[eskf::S_GRAV] = eskf::kS2Len;
  cudaError_t e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaHostAlloc((void**)&m->h_poses, kMaxImuPoses * sizeof(ImuPoseDev), cudaHostAllocMapped);
  if (e == cudaSuccess) e = cudaHostGetDevicePointer((void**)&m->d_poses, m->h_poses, 0);
  if (e == cudaSuccess) e = cudaMalloc((void**)&m->d_first, 8);
  if (e != cudaSuccess) { lsd_status_t r = cuda_fail(e, "lsd_imu_create", __FILE__, __LINE__); lsd_imu_destroy(m); return r; }
  *out = m;
  return LSD_OK;
}

lsd_status_t lsd_imu_destroy(lsd_imu_t* m) {
  if (!m) return LSD_OK;
  cudaSetDevice(m->device);
  if (m->stream) cudaStreamSynchronize(m->stream);
  cudaFree(m->d_in); cudaFree(m->d_out); cudaFree(m->d_time); cudaFree(m->d_first);
  cudaFreeHost(m->h_poses);
  if (m->stream) cudaStreamDestroy(m->stream);
  delete m;
  return LSD_OK;
}

lsd_status_t lsd_imu_reset(lsd_imu_t* m) {
  if (!m) return LSD_ERR_INVALID;
  imu_reset(m);
  m->b_first_frame = true;
  return LSD_OK;
}

lsd_status_t lsd_imu_is_init(lsd_imu_t* m, int* flag) { if (!m || !flag) return LSD_ERR_INVALID; *flag = m->state_init_done ? 1 : 0; return LSD_OK; }

lsd_status_t lsd_eskf_predict(double* state26_inout, double* P529_inout, double dt, const double* Q144, const double* acc3, const double* gyro3) {
  if (!state26_inout || !P529_inout || !Q144 || !acc3 || !gyro3) return LSD_ERR_INVALID;
  eskf::predict(state26_inout, P529_inout, dt, Q144, acc3, gyro3);
  return LSD_OK;
}

lsd_status_t lsd_imu_process_dev(lsd_imu_t* m, const double* imu7, int n_imu, const double* ins_vel3_or_null, double lidar_beg_time,
                                 double lidar_end_time, const float* xyzi_dev, const float* time_ms_dev, int n, double* state26_inout,
                                 double* P529_inout, int* n_out) {
  if (!m || n < 0 || n_imu < 0 || (n_imu > 0 && !imu7) || (n > 0 && (!xyzi_dev || !time_ms_dev)) || !state26_inout || !P529_inout) return LSD_ERR_INVALID;
  LSD_CUDA(cudaSetDevice(m->device));
  lsd_status_t s = imu_alloc(m, n);
  if (s) return s;
  return imu_process(m, imu7, n_imu, ins_vel3_or_null, lidar_beg_time, lidar_end_time, reinterpret_cast<const float4*>(xyzi_dev), time_ms_dev, n,
                     state26_inout, P529_inout, n_out);
}

lsd_status_t lsd_imu_process(lsd_imu_t* m, const double* imu7, int n_imu, const double* ins_vel3_or_null, double lidar_beg_time,
                             double lidar_end_time, const float* xyzi_host, const float* time_ms_host, int n, double* state26_inout,
                             double* P529_inout, int* n_out) {
  if (!m || n < 0 || n_imu < 0 || (n_imu > 0 && !imu7) || (n > 0 && (!xyzi_host || !time_ms_host)) || !state26_inout || !P529_inout) return LSD_ERR_INVALID;
  LSD_CUDA(cudaSetDevice(m->device));
  lsd_status_t s = imu_alloc(m, n);
  if (s) return s;
  if (n > 0 && !m->imu_need_init && n_imu > 0) {  // the copies overlap the host-side forward propagation
    LSD_CUDA(cudaMemcpyAsync(m->d_in, xyzi_host, (size_t)n * 16, cudaMemcpyHostToDevice, m->stream));
    LSD_CUDA(cudaMemcpyAsync(m->d_time, time_ms_host, (size_t)n * 4, cudaMemcpyHostToDevice, m->stream));
  }
  return imu_process(m, imu7, n_imu, ins_vel3_or_null, lidar_beg_time, lidar_end_time, m->d_in, m->d_time, n, state26_inout, P529_inout, n_out);
}

lsd_status_t lsd_imu_get_cloud_dev(lsd_imu_t* m, const float** xyzi_dev, int* n, void** cuda_stream_out) {
  if (!m || !xyzi_dev || !n) return LSD_ERR_INVALID;
  *xyzi_dev = reinterpret_cast<const float*>(m->d_out);
  *n = m->n_out;
  if (cuda_stream_out) *cuda_stream_out = static_cast<void*>(m->stream);
  return LSD_OK;
}

lsd_status_t lsd_imu_get_cloud(lsd_imu_t* m, float* xyzi_host, int cap, int* n) {
  if (!m || !n || (cap > 0 && !xyzi_host)) return LSD_ERR_INVALID;
  LSD_CUDA(cudaSetDevice(m->device));
  *n = m->n_out;
  const int c = std::min(cap, m->n_out);
  if (c > 0) LSD_CUDA(cudaMemcpyAsync(xyzi_host, m->d_out, (size_t)c * 16, cudaMemcpyDeviceToHost, m->stream));
  LSD_CUDA(cudaStreamSynchronize(m->stream));
  return LSD_OK;
}

// ImuProcess::start_state_point and mean_acc_norm: what fastlio_state / fastlio_odometry read (laserMapping.cpp:690-738)
lsd_status_t lsd_imu_get_start_state(lsd_imu_t* m, double* state26, double* mean_acc_norm) {
  if (!m || !state26) return LSD_ERR_INVALID;
  memcpy(state26, m->start_state, sizeof(m->start_state));
  if (mean_acc_norm) *mean_acc_norm = m->mean_acc_norm;
  return LSD_OK;
}

lsd_status_t lsd_imu_get_poses(lsd_imu_t* m, double* poses22, int cap, int* n) {
  if (!m || !n || (cap > 0 && !poses22)) return LSD_ERR_INVALID;
  *n = (int)m->poses.size();
  const int c = std::min(cap, *n);
  static_assert(sizeof(ImuPoseDev) == 22 * 8, "pose tap layout");
  if (c > 0) memcpy(poses22, m->poses.data(), (size_t)c * sizeof(ImuPoseDev));
  return LSD_OK;
}

}  // extern "C"
