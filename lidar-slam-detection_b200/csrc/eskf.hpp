// eskf.hpp — host-side (double) iterated error-state Kalman filter on the 23-DOF LIO manifold.
//
// Product code, written from the reference's published algorithm — NOT derived from oracle/.
// Follows (paths relative to /root/reference/slam/mapping/fastlio/include):
//   esekf::update_iterated_dyn_share_modified   IKFoM_toolkit/esekfom/esekfom.hpp:1619-1931
//   state_ikfom                                 use-ikfom.hpp:12-21
//   SO3 / S2 / vect manifold operators          IKFoM_toolkit/mtk/types/{SOn,S2,vect}.hpp
//   exp / log / A_matrix / cos_sinc_sqrt        IKFoM_toolkit/mtk/src/mtkmath.hpp
// The 23x23 algebra is negligible next to the per-point work and stays on the host
// in double, exactly like the reference; only h_x^T h_x and h_x^T h come from the GPU.
#pragma once
#include <math.h>
#include <string.h>

#include <functional>
#include <vector>

namespace lsd {
namespace eskf {

constexpr int N = 23;
constexpr double kTol = 1e-11;             // MTK::tolerance<double>()
constexpr double kS2Len = 98090.0 / 10000.0;  // S2<double, 98090, 10000, 1>

// state vector (26 doubles): pos[3] rot[4] offR[4] offT[3] vel[3] bg[3] ba[3] grav[3]; quats (x,y,z,w)
enum { S_POS = 0, S_ROT = 3, S_OFFR = 7, S_OFFT = 11, S_VEL = 14, S_BG = 17, S_BA = 20, S_GRAV = 23, S_DIM = 26 };
// DOF indices
enum { D_POS = 0, D_ROT = 3, D_OFFR = 6, D_OFFT = 9, D_VEL = 12, D_BG = 15, D_BA = 18, D_GRAV = 21 };

struct Mat3 { double m[9]; };

inline void hat(const double* v, double* H) {
  H[0] = 0; H[1] = -v[2]; H[2] = v[1]; H[3] = v[2]; H[4] = 0; H[5] = -v[0]; H[6] = -v[1]; H[7] = v[0]; H[8] = 0;
}
inline void mm3(const double* A, const double* B, double* C) {
  double t[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  memcpy(C, t, sizeof(t));
}
inline void mv3(const double* A, const double* v, double* o) {
  double t[3];
  for (int i = 0; i < 3; i++) t[i] = A[3 * i] * v[0] + A[3 * i + 1] * v[1] + A[3 * i + 2] * v[2];
  o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
}
inline void tr3(const double* A, double* T) {
  double t[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * j + i];
  memcpy(T, t, sizeof(t));
}

// mtkmath.hpp:117-150
inline void cos_sinc_sqrt(double x2, double* c, double* s) {
  const double eps = 2.220446049250313e-16;
  const double t2 = sqrt(eps), tn = sqrt(t2);
  if (x2 >= tn) { double x = sqrt(x2); *c = cos(x); *s = sin(x) / x; return; }
  static const double inv[] = {1 / 3., 1 / 4., 1 / 5., 1 / 6., 1 / 7., 1 / 8., 1 / 9.};
  double cosi = 1., sinc = 1., term = -1 / 2. * x2;
  for (int i = 0; i < 3; ++i) { cosi += term; term *= inv[2 * i]; sinc += term; term *= -inv[2 * i + 1] * x2; }
  *c = cosi; *s = sinc;
}
// mtkmath.hpp:236-243: quaternion (x,y,z,w) of exp(scale * vec) as MTK::exp lays it out
inline void mtk_exp(const double* vec, double scale, double* q) {
  double n2 = vec[0] * vec[0] + vec[1] * vec[1] + vec[2] * vec[2], c, s;
  cos_sinc_sqrt(scale * scale * n2, &c, &s);
  double mult = s * scale;
  q[0] = mult * vec[0]; q[1] = mult * vec[1]; q[2] = mult * vec[2]; q[3] = c;
}
inline void qmul(const double* a, const double* b, double* o) {
  double t[4];
  t[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  t[1] = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
  t[2] = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
  t[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  memcpy(o, t, sizeof(t));
}
inline void qconj(const double* a, double* o) { o[0] = -a[0]; o[1] = -a[1]; o[2] = -a[2]; o[3] = a[3]; }
// Eigen::QuaternionBase::toRotationMatrix
inline void q2R(const double* q, double* R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
// Eigen's rotation-matrix -> quaternion (Quaternion.h, "Quaternion Calculus and Fast Animation"): SO3(matrix), SOn.hpp:222
inline void R2q(const double* R, double* q) {
  double t = R[0] + R[4] + R[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[4 * i] - R[4 * j] - R[4 * k] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[3 * k + j] - R[3 * j + k]) * t;
    q[j] = (R[3 * j + i] + R[3 * i + j]) * t;
    q[k] = (R[3 * k + i] + R[3 * i + k]) * t;
  }
}
// SOn.hpp:284-288 + mtkmath.hpp:254-275 (scale 2, plus_minus_periodicity)
inline void so3_log(const double* q, double* r) {
  double nv = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
  if (nv < kTol) nv = kTol;
  double s = 2.0 / nv * atan(nv / q[3]);
  r[0] = s * q[0]; r[1] = s * q[1]; r[2] = s * q[2];
}
// mtkmath.hpp:222-234
inline void A_matrix(const double* v, double* A) {
  double sq = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], n = sqrt(sq);
  for (int i = 0; i < 9; i++) A[i] = (i % 4 == 0) ? 1.0 : 0.0;
  if (n < kTol) return;
  double H[9], HH[9];
  hat(v, H); mm3(H, H, HH);
  double a = (1 - cos(n)) / sq, b = (1 - sin(n) / n) / sq;
  for (int i = 0; i < 9; i++) A[i] += a * H[i] + b * HH[i];
}
// S2.hpp:160-215, S2_typ == 1; Bx is 3x2 row-major
inline void s2_Bx(const double* v, double* B) {
  const double L = kS2Len;
  if (v[0] + L > kTol) {
    const double d = L + v[0];
    B[0] = -v[1]; B[1] = -v[2];
    B[2] = L - v[1] * v[1] / d; B[3] = -v[2] * v[1] / d;
    B[4] = -v[2] * v[1] / d; B[5] = L - v[2] * v[2] / d;
    for (int i = 0; i < 6; i++) B[i] /= L;
  } else {
    for (int i = 0; i < 6; i++) B[i] = 0;
    B[3] = -1; B[4] = 1;
  }
}
inline void s2_boxplus(double* v, const double* d2) {  // S2.hpp:124-130
  double B[6], Bu[3], q[4], R[9];
  s2_Bx(v, B);
  for (int i = 0; i < 3; i++) Bu[i] = B[2 * i] * d2[0] + B[2 * i + 1] * d2[1];
  mtk_exp(Bu, 0.5, q);
  q2R(q, R);
  mv3(R, v, v);
}
inline void s2_boxminus(const double* v, const double* o, double* r) {  // S2.hpp:132-157
  double H[9], c[3];
  hat(v, H); mv3(H, o, c);
  double v_sin = sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
  double v_cos = v[0] * o[0] + v[1] * o[1] + v[2] * o[2];
  double theta = atan2(v_sin, v_cos);
  if (v_sin < kTol) {
    if (fabs(theta) > kTol) { r[0] = 3.1415926; r[1] = 0; } else { r[0] = r[1] = 0; }
    return;
  }
  double B[6], Ho[9], t[3];
  s2_Bx(o, B); hat(o, Ho); mv3(Ho, v, t);
  for (int j = 0; j < 2; j++) r[j] = theta / v_sin * (B[j] * t[0] + B[2 + j] * t[1] + B[4 + j] * t[2]);
}
inline void s2_Nx_yy(const double* v, double* Nx /*2x3*/) {  // S2.hpp:251-256
  double B[6], H[9];
  s2_Bx(v, B); hat(v, H);
  for (int i = 0; i < 2; i++) for (int j = 0; j < 3; j++)
    Nx[3 * i + j] = (1 / kS2Len / kS2Len) * (B[i] * H[j] + B[2 + i] * H[3 + j] + B[4 + i] * H[6 + j]);
}
inline void s2_Mx(const double* v, const double* d2, double* Mx /*3x2*/) {  // S2.hpp:258-272
  double B[6], H[9];
  s2_Bx(v, B); hat(v, H);
  double T[9];
  if (sqrt(d2[0] * d2[0] + d2[1] * d2[1]) < kTol) {
    for (int i = 0; i < 9; i++) T[i] = -H[i];
  } else {
    // the reference evaluates exp(Bu, scalar(1/2)) with an INTEGER 1/2 == 0, i.e. the identity
    double Bu[3], A[9], At[9];
    for (int i = 0; i < 3; i++) Bu[i] = B[2 * i] * d2[0] + B[2 * i + 1] * d2[1];
    A_matrix(Bu, A); tr3(A, At);
    mm3(H, At, T);
    for (int i = 0; i < 9; i++) T[i] = -T[i];
  }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 2; j++) Mx[2 * i + j] = T[3 * i] * B[j] + T[3 * i + 1] * B[2 + j] + T[3 * i + 2] * B[4 + j];
}

inline void boxplus(double* x, const double* d) {
  double q[4];
  for (int i = 0; i < 3; i++) x[S_POS + i] += d[D_POS + i];
  mtk_exp(d + D_ROT, 0.5, q); qmul(x + S_ROT, q, x + S_ROT);
  mtk_exp(d + D_OFFR, 0.5, q); qmul(x + S_OFFR, q, x + S_OFFR);
  for (int i = 0; i < 3; i++) x[S_OFFT + i] += d[D_OFFT + i];
  for (int i = 0; i < 3; i++) x[S_VEL + i] += d[D_VEL + i];
  for (int i = 0; i < 3; i++) x[S_BG + i] += d[D_BG + i];
  for (int i = 0; i < 3; i++) x[S_BA + i] += d[D_BA + i];
  s2_boxplus(x + S_GRAV, d + D_GRAV);
}
inline void boxminus(const double* a, const double* b, double* r) {  // a (-) b
  double qc[4], q[4];
  for (int i = 0; i < 3; i++) r[D_POS + i] = a[S_POS + i] - b[S_POS + i];
  qconj(b + S_ROT, qc); qmul(qc, a + S_ROT, q); so3_log(q, r + D_ROT);
  qconj(b + S_OFFR, qc); qmul(qc, a + S_OFFR, q); so3_log(q, r + D_OFFR);
  for (int i = 0; i < 3; i++) r[D_OFFT + i] = a[S_OFFT + i] - b[S_OFFT + i];
  for (int i = 0; i < 3; i++) r[D_VEL + i] = a[S_VEL + i] - b[S_VEL + i];
  for (int i = 0; i < 3; i++) r[D_BG + i] = a[S_BG + i] - b[S_BG + i];
  for (int i = 0; i < 3; i++) r[D_BA + i] = a[S_BA + i] - b[S_BA + i];
  s2_boxminus(a + S_GRAV, b + S_GRAV, r + D_GRAV);
}

inline void init_cov(double* P) {  // IMU_Processing.hpp:224-230
  for (int i = 0; i < N * N; i++) P[i] = 0;
  for (int i = 0; i < N; i++) P[i * N + i] = 1.0;
  for (int i = 6; i < 12; i++) P[i * N + i] = 0.00001;
  for (int i = 15; i < 18; i++) P[i * N + i] = 0.0001;
  for (int i = 18; i < 21; i++) P[i * N + i] = 0.001;
  P[21 * N + 21] = P[22 * N + 22] = 0.00001;
}

// Dense inverse by LU with partial pivoting (what Eigen's inverse() does for sizes > 4).
inline bool invert(const double* A, double* Ainv, int n) {
  std::vector<double> a(A, A + n * n);
  std::vector<int> piv(n);
  for (int i = 0; i < n * n; i++) Ainv[i] = 0;
  for (int i = 0; i < n; i++) Ainv[i * n + i] = 1;
  for (int k = 0; k < n; k++) {
    int p = k; double best = fabs(a[k * n + k]);
    for (int i = k + 1; i < n; i++) if (fabs(a[i * n + k]) > best) { best = fabs(a[i * n + k]); p = i; }
    if (best == 0.0) return false;
    if (p != k) for (int j = 0; j < n; j++) { std::swap(a[k * n + j], a[p * n + j]); std::swap(Ainv[k * n + j], Ainv[p * n + j]); }
    const double d = a[k * n + k];
    for (int i = k + 1; i < n; i++) {
      const double f = a[i * n + k] / d;
      if (f == 0.0) continue;
      for (int j = k; j < n; j++) a[i * n + j] -= f * a[k * n + j];
      for (int j = 0; j < n; j++) Ainv[i * n + j] -= f * Ainv[k * n + j];
    }
  }
  for (int k = n - 1; k >= 0; k--) {
    const double d = a[k * n + k];
    for (int j = 0; j < n; j++) Ainv[k * n + j] /= d;
    for (int i = 0; i < k; i++) {
      const double f = a[i * n + k];
      if (f == 0.0) continue;
      for (int j = 0; j < n; j++) Ainv[i * n + j] -= f * Ainv[k * n + j];
    }
  }
  return true;
}

// rows [r0, r0+nr) of P (23x23) <- M (nr x nr) * rows ; cols likewise with M^T
inline void left_rows(double* P, int r0, int nr, const double* M) {
  double t[3 * N];
  for (int i = 0; i < nr; i++) for (int c = 0; c < N; c++) { double s = 0; for (int k = 0; k < nr; k++) s += M[i * nr + k] * P[(r0 + k) * N + c]; t[i * N + c] = s; }
  for (int i = 0; i < nr; i++) for (int c = 0; c < N; c++) P[(r0 + i) * N + c] = t[i * N + c];
}
inline void right_cols(double* P, int c0, int nc, const double* M) {  // P[:, c0:c0+nc] = P[:, ...] * M^T
  for (int r = 0; r < N; r++) {
    double t[3];
    for (int j = 0; j < nc; j++) { double s = 0; for (int k = 0; k < nc; k++) s += P[r * N + c0 + k] * M[j * nc + k]; t[j] = s; }
    for (int j = 0; j < nc; j++) P[r * N + c0 + j] = t[j];
  }
}

// Output of one measurement-model evaluation (h_share_model, laserMapping.cpp:984-1023)
struct HModel {
  bool valid = false;
  int n = 0;            // dof_Measurement (rows of h_x)
  double HTH[15 * 15];  // h_x^T h_x
  double HTh[15];       // h_x^T h
  std::vector<double> h_x, h;  // only filled when n < 23 (esekfom.hpp:1727 branch)
};
using HFunc = std::function<void(const double* x26, bool converge, HModel* out)>;

struct UpdateResult { int evaluations = 0; bool returned_converged = false; };

// esekfom.hpp:1619-1931
inline UpdateResult update_iterated(double* x, double* P, const HFunc& h_model, double R, int maximum_iter, double limit,
                                    bool literal = false) {
  UpdateResult res;
  double x_prop[S_DIM], P_prop[N * N];
  memcpy(x_prop, x, sizeof(x_prop)); memcpy(P_prop, P, sizeof(P_prop));
  bool converge = true;
  int t = 0;
  static thread_local HModel m;
  double K_x[N * N];
  for (int i = 0; i < N * N; i++) K_x[i] = 0.0;
  double K_h[N];
  int kx_cols = 15;  // non-zero columns of K_x
  const int so3_idx[2] = {D_ROT, D_OFFR};
  for (int it = -1; it < maximum_iter; it++) {
    m.valid = true;
    h_model(x, converge, &m);
    res.evaluations++;
    if (!m.valid) continue;
    const int dof = m.n;
    double dx[N], dx_new[N];
    boxminus(x, x_prop, dx);
    memcpy(dx_new, dx, sizeof(dx));
    memcpy(P, P_prop, sizeof(P_prop));
    for (int s = 0; s < 2; s++) {
      const int idx = so3_idx[s];
      double A[9], At[9], v[3];
      A_matrix(dx + idx, A); tr3(A, At);
      mv3(At, dx_new + idx, v); dx_new[idx] = v[0]; dx_new[idx + 1] = v[1]; dx_new[idx + 2] = v[2];
      left_rows(P, idx, 3, At);
      right_cols(P, idx, 3, At);
    }
    {
      double Nx[6], Mx[6], r2[4];
      s2_Nx_yy(x + S_GRAV, Nx); s2_Mx(x_prop + S_GRAV, dx + D_GRAV, Mx);
      for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) r2[2 * i + j] = Nx[3 * i] * Mx[j] + Nx[3 * i + 1] * Mx[2 + j] + Nx[3 * i + 2] * Mx[4 + j];
      double a = dx_new[D_GRAV], b = dx_new[D_GRAV + 1];
      dx_new[D_GRAV] = r2[0] * a + r2[1] * b; dx_new[D_GRAV + 1] = r2[2] * a + r2[3] * b;
      left_rows(P, D_GRAV, 2, r2);
      right_cols(P, D_GRAV, 2, r2);
    }
    if (N > dof) {
      // K = P H^T (H P H^T / R + I)^-1 / R   (esekfom.hpp:1727-1753)
      std::vector<double> H(dof * N, 0.0), PHt(N * dof), S(dof * dof), Sinv(dof * dof), K(N * dof);
      for (int r = 0; r < dof; r++) for (int c = 0; c < 15; c++) H[r * N + c] = m.h_x[r * 15 + c];
      for (int i = 0; i < N; i++) for (int r = 0; r < dof; r++) { double s = 0; for (int k = 0; k < N; k++) s += P[i * N + k] * H[r * N + k]; PHt[i * dof + r] = s; }
      for (int a = 0; a < dof; a++) for (int b = 0; b < dof; b++) { double s = 0; for (int k = 0; k < N; k++) s += H[a * N + k] * PHt[k * dof + b]; S[a * dof + b] = s / R + (a == b ? 1.0 : 0.0); }
      invert(S.data(), Sinv.data(), dof);
      for (int i = 0; i < N; i++) for (int b = 0; b < dof; b++) { double s = 0; for (int k = 0; k < dof; k++) s += PHt[i * dof + k] * Sinv[k * dof + b]; K[i * dof + b] = s / R; }
      for (int i = 0; i < N; i++) { double s = 0; for (int k = 0; k < dof; k++) s += K[i * dof + k] * m.h[k]; K_h[i] = s; }
      for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) { double s = 0; for (int k = 0; k < dof; k++) s += K[i * dof + k] * H[k * N + j]; K_x[i * N + j] = s; }
      kx_cols = N;
    } else if (literal) {
      kx_cols = 15;
      // the reference's literal evaluation: two dense 23x23 inversions (esekfom.hpp:1756-1789)
      double Pt[N * N], Pinv[N * N];
      for (int i = 0; i < N * N; i++) Pt[i] = P[i] / R;
      invert(Pt, Pinv, N);  // P_temp = (P/R)^-1
      for (int i = 0; i < 15; i++) for (int j = 0; j < 15; j++) Pinv[i * N + j] += m.HTH[i * 15 + j];
      invert(Pinv, Pt, N);  // P_inv
      for (int i = 0; i < N; i++) { double s = 0; for (int k = 0; k < 15; k++) s += Pt[i * N + k] * m.HTh[k]; K_h[i] = s; }
      for (int i = 0; i < N * N; i++) K_x[i] = 0;
      for (int i = 0; i < N; i++) for (int j = 0; j < 15; j++) { double s = 0; for (int k = 0; k < 15; k++) s += Pt[i * N + k] * m.HTH[k * 15 + j]; K_x[i * N + j] = s; }
    } else {
      // Same quantities without the two dense 23x23 inversions.  h_x^T h_x is non-zero only in its
      // leading 6x6 block H (extrinsic_est_en == false, laserMapping.cpp:82,926-930).  With A = P/R
      // partitioned at 6, the first 6 columns of P_inv = (A^-1 + U H U^T)^-1 — the only ones K_h and
      // K_x use — follow from the Schur complement of the trailing block:
      //   P_inv[:, 0:6] = [I; A21 A11^-1] (A11^-1 + H)^-1 = P[:, 0:6] P66^-1 (R P66^-1 + H)^-1
      // Two 6x6 inverses of positive-definite sums (no cancellation), 23x6 products.
      kx_cols = 6;
      double H6[36], P66[36], P66i[36], S[36], Sinv[36], T1[36], C6[N * 6];
      for (int a = 0; a < 6; a++) for (int c = 0; c < 6; c++) { H6[6 * a + c] = m.HTH[15 * a + c]; P66[6 * a + c] = P[a * N + c]; }
      invert(P66, P66i, 6);
      for (int i = 0; i < 36; i++) S[i] = R * P66i[i] + H6[i];
      invert(S, Sinv, 6);
      for (int a = 0; a < 6; a++) for (int c = 0; c < 6; c++) { double s = 0; for (int k = 0; k < 6; k++) s += P66i[6 * a + k] * Sinv[6 * k + c]; T1[6 * a + c] = s; }
      for (int i = 0; i < N; i++) for (int c = 0; c < 6; c++) { double s = 0; for (int k = 0; k < 6; k++) s += P[i * N + k] * T1[6 * k + c]; C6[i * 6 + c] = s; }
      for (int i = 0; i < N; i++) { double s = 0; for (int k = 0; k < 6; k++) s += C6[i * 6 + k] * m.HTh[k]; K_h[i] = s; }
      for (int i = 0; i < N * N; i++) K_x[i] = 0;
      for (int i = 0; i < N; i++) for (int j = 0; j < 6; j++) { double s = 0; for (int k = 0; k < 6; k++) s += C6[i * 6 + k] * H6[6 * k + j]; K_x[i * N + j] = s; }
    }
    double dx_[N];
    for (int i = 0; i < N; i++) { double s = K_h[i] - dx_new[i]; for (int j = 0; j < kx_cols; j++) s += K_x[i * N + j] * dx_new[j]; dx_[i] = s; }
    boxplus(x, dx_);
    converge = true;
    for (int i = 0; i < N; i++) if (fabs(dx_[i]) > limit) { converge = false; break; }
    if (converge) t++;
    if (!t && it == maximum_iter - 2) converge = true;
    if (t > 1 || it == maximum_iter - 1) {
      double L[N * N];
      memcpy(L, P, sizeof(L));
      for (int s = 0; s < 2; s++) {
        const int idx = so3_idx[s];
        double A[9], At[9];
        A_matrix(dx_ + idx, A); tr3(A, At);
        // L rows <- At * P rows (P as it stands now), K_x rows (first 15 cols) <- At * K_x rows
        for (int c = 0; c < N; c++) { double v[3] = {P[idx * N + c], P[(idx + 1) * N + c], P[(idx + 2) * N + c]}, o[3]; mv3(At, v, o); L[idx * N + c] = o[0]; L[(idx + 1) * N + c] = o[1]; L[(idx + 2) * N + c] = o[2]; }
        for (int c = 0; c < 15; c++) { double v[3] = {K_x[idx * N + c], K_x[(idx + 1) * N + c], K_x[(idx + 2) * N + c]}, o[3]; mv3(At, v, o); K_x[idx * N + c] = o[0]; K_x[(idx + 1) * N + c] = o[1]; K_x[(idx + 2) * N + c] = o[2]; }
        right_cols(L, idx, 3, At);
        right_cols(P, idx, 3, At);
      }
      {
        double Nx[6], Mx[6], r2[4];
        s2_Nx_yy(x + S_GRAV, Nx); s2_Mx(x_prop + S_GRAV, dx_ + D_GRAV, Mx);
        for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) r2[2 * i + j] = Nx[3 * i] * Mx[j] + Nx[3 * i + 1] * Mx[2 + j] + Nx[3 * i + 2] * Mx[4 + j];
        const int idx = D_GRAV;
        for (int c = 0; c < N; c++) { double a = P[idx * N + c], b = P[(idx + 1) * N + c]; L[idx * N + c] = r2[0] * a + r2[1] * b; L[(idx + 1) * N + c] = r2[2] * a + r2[3] * b; }
        for (int c = 0; c < 15; c++) { double a = K_x[idx * N + c], b = K_x[(idx + 1) * N + c]; K_x[idx * N + c] = r2[0] * a + r2[1] * b; K_x[(idx + 1) * N + c] = r2[2] * a + r2[3] * b; }
        right_cols(L, idx, 2, r2);
        right_cols(P, idx, 2, r2);
      }
      double Pn[N * N];
      for (int i = 0; i < N; i++) for (int j = 0; j < N; j++) { double s = 0; for (int k = 0; k < 15; k++) s += K_x[i * N + k] * P[k * N + j]; Pn[i * N + j] = L[i * N + j] - s; }
      memcpy(P, Pn, sizeof(Pn));
      res.returned_converged = t > 1;
      return res;
    }
  }
  return res;
}

}  // namespace eskf
}  // namespace lsd
