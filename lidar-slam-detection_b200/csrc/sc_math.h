// sc_math.h — the arithmetic of the ScanContext kernels (row N4), as __host__ __device__ functions so that the
// same lines are pinned on the CPU (tests/sc_host_harness.cpp runs them sequentially against the compiled reference)
// and run in csrc/scancontext.cu's kernels.  No CUDA headers needed when compiled by g++.
//
// Follows slam/common/Scancontext/Scancontext.cpp (reference): xy2theta :21-36, makeScancontext :160-203, the keys
// :206-236, distDirectSC :77-99, fastAlignUsingVkey :102-123; nanoflann's L2_Adaptor::evalMetric for the ring keys.
// Reductions reproduce Eigen 3.3's vectorised redux of the reference's x86-64 (SSE2) build: four interleaved partial
// sums s_k over i = k mod 4, result (s0 + s2) + (s1 + s3).  Build without FMA contraction (-fmad=false / -ffp-contract=off).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define LSD_HD __host__ __device__ __forceinline__
#else
#define LSD_HD inline
#endif

namespace lsd {
namespace sc {

constexpr int kRing = 20, kSector = 60, kDesc = kRing * kSector;
constexpr int kCand = 10;            // NUM_CANDIDATES_FROM_TREE (Scancontext.h:78)
constexpr int kSearchRadius = 3;     // round(0.5 * SEARCH_RATIO * PC_NUM_SECTOR), Scancontext.cpp:133
constexpr double kMaxRadius = 80.0, kLidarHeight = 0.5;
constexpr float kNoPoint = -1000.0f;
constexpr double kBig = 10000000.0;  // "init with something large"

// monotone float -> int map for atomicMax on z
LSD_HD int enc_z(float z) {
  union { float f; int i; } u;
  u.f = z;
  return u.i >= 0 ? u.i : (u.i ^ 0x7fffffff);
}
LSD_HD float dec_z(int e) {
  union { float f; int i; } u;
  u.i = e >= 0 ? e : (e ^ 0x7fffffff);
  return u.f;
}

// xy2theta (Scancontext.cpp:21-36): float in, `atan` is ::atan(double), float out
LSD_HD float xy2theta(float x, float y) {
  const double k = 180.0 / M_PI;
  if ((x >= 0) & (y >= 0)) return (float)(k * atan((double)(y / x)));
  if ((x < 0) & (y >= 0)) return (float)(180.0 - k * atan((double)(y / (-x))));
  if ((x < 0) & (y < 0)) return (float)(180.0 + k * atan((double)(y / x)));
  if ((x >= 0) & (y < 0)) return (float)(360.0 - k * atan((double)((-y) / x)));
  return 0.0f;
}

// int(ceil(v)) with the x86-64 result for NaN / out-of-range (cvttsd2si -> INT_MIN), so that NaN angles land in bin 1
LSD_HD int ceil_to_int(double v) {
  const double c = ceil(v);
  if (!(c > -2147483649.0 && c < 2147483648.0)) return (int)0x80000000;
  return (int)c;
}

// One point of makeScancontext (:171-192): false = outside the 80 m disc; else the bin (sector * 20 + ring, Eigen's
// column-major element order) and the height that competes for its maximum.
LSD_HD bool point_bin(float px, float py, float pz, double dx, double dy, int* bin, float* z) {
  const float x = (float)((double)px + dx), y = (float)((double)py + dy);
  *z = (float)((double)pz + kLidarHeight);
  const float xx = x * x, yy = y * y;
  const float range = (float)sqrt((double)(xx + yy));
  const float angle = xy2theta(x, y);
  if ((double)range > kMaxRadius) return false;
  int ring = ceil_to_int(((double)range / kMaxRadius) * kRing);
  int sector = ceil_to_int(((double)angle / 360.0) * kSector);
  ring = ring < kRing ? ring : kRing; ring = ring > 1 ? ring : 1;
  sector = sector < kSector ? sector : kSector; sector = sector > 1 ? sector : 1;
  *bin = (sector - 1) * kRing + (ring - 1);
  return true;
}

// Eigen's redux order over n (multiple of 4) elements f(0..n-1)
template <class F>
LSD_HD double eig_redux(int n, F f) {
  double s0 = f(0), s1 = f(1), s2 = f(2), s3 = f(3);
  for (int i = 4; i < n; i += 4) { s0 = s0 + f(i); s1 = s1 + f(i + 1); s2 = s2 + f(i + 2); s3 = s3 + f(i + 3); }
  return (s0 + s2) + (s1 + s3);
}

// keys of a descriptor d[1200] (column-major: d[sector * 20 + ring])
LSD_HD double ring_mean(const double* d, int ring) { return eig_redux(kSector, [&](int s) { return d[s * kRing + ring]; }) / (double)kSector; }
LSD_HD double sector_mean(const double* d, int sector) { return eig_redux(kRing, [&](int r) { return d[sector * kRing + r]; }) / (double)kRing; }
LSD_HD double sector_norm(const double* d, int sector) {
  return sqrt(eig_redux(kRing, [&](int r) { const double v = d[sector * kRing + r]; return v * v; }));
}

// nanoflann L2_Adaptor::evalMetric on 20 floats: groups of four, ((d0^2 + d1^2) + d2^2) + d3^2 added to the running sum
LSD_HD float ring_d2(const float* a, const float* b) {
  float res = 0.f;
  for (int g = 0; g < kRing; g += 4) {
    const float d0 = a[g] - b[g], d1 = a[g + 1] - b[g + 1], d2 = a[g + 2] - b[g + 2], d3 = a[g + 3] - b[g + 3];
    const float p0 = d0 * d0, p1 = d1 * d1, p2 = d2 * d2, p3 = d3 * d3;
    res = res + (((p0 + p1) + p2) + p3);
  }
  return res;
}

LSD_HD int wrap(int j) { return j < 0 ? j + kSector : (j >= kSector ? j - kSector : j); }

// fastAlignUsingVkey's |vkey1 - circshift(vkey2, shift)| (:106-111)
LSD_HD double vkey_diff_norm(const double* vk1, const double* vk2, int shift) {
  return sqrt(eig_redux(kSector, [&](int j) { const double v = vk1[j] - vk2[wrap(j - shift)]; return v * v; }));
}

// one sector pair of distDirectSC(sc1, circshift(sc2, shift)) (:81-92): false = not counted
LSD_HD bool sector_similarity(const double* a, const double* an, const double* b, const double* bn, int j, int shift, double* sim) {
  const int jb = wrap(j - shift);
  const double n1 = an[j], n2 = bn[jb];
  if ((n1 == 0) | (n2 == 0)) return false;
  const double dot = eig_redux(kRing, [&](int r) { return a[j * kRing + r] * b[jb * kRing + r]; });
  *sim = dot / (n1 * n2);
  return true;
}

// the sorted shift search space around the vkey alignment (:133-140)
LSD_HD void search_space(int align, int (&space)[2 * kSearchRadius + 1]) {
  int n = 0;
  space[n++] = align;
  for (int ii = 1; ii <= kSearchRadius; ii++) { space[n++] = (align + ii + kSector) % kSector; space[n++] = (align - ii + kSector) % kSector; }
  for (int i = 1; i < n; i++) {  // insertion sort, 7 elements
    const int v = space[i];
    int k = i - 1;
    while (k >= 0 && space[k] > v) { space[k + 1] = space[k]; k--; }
    space[k + 1] = v;
  }
}

// float deg2rad(float) (:15-18) of nn_align * PC_UNIT_SECTORANGLE
LSD_HD float shift_to_yaw(int shift) {
  const float deg = (float)((double)shift * (360.0 / (double)kSector));
  return (float)((double)deg * M_PI / 180.0);
}

}  // namespace sc
}  // namespace lsd
