// TEST INFRASTRUCTURE.  Runs the arithmetic of the ScanContext kernels (lidar-slam-detection_b200/csrc/sc_math.h — the very
// functions csrc/scancontext.cu's kernels call) sequentially on the host, in the kernels' own decomposition (per point,
// per lane, per round), so that it can be pinned against the compiled reference without a GPU
// (tests/test_oracle_scancontext.py::test_kernel_arithmetic_on_the_host).  Built by the test with
// g++ -O2 -ffp-contract=off.  Not part of the product: liblsdreg.so has no host path for these entry points.
#include <vector>

#include "../lidar-slam-detection_b200/csrc/sc_math.h"

using namespace lsd::sc;

extern "C" {

// sc_fill_kernel + sc_bin_kernel + sc_finish_kernel
void h_sc_make(const float* xyzi, int n, double dx, double dy, double* desc, double* ringkey, float* ringkey_f, double* vkey, double* norm) {
  std::vector<int> enc(kDesc, enc_z(kNoPoint));
  for (int i = 0; i < n; i++) {
    int bin; float z;
    if (point_bin(xyzi[4 * i], xyzi[4 * i + 1], xyzi[4 * i + 2], dx, dy, &bin, &z)) { const int e = enc_z(z); if (e > enc[bin]) enc[bin] = e; }
  }
  for (int i = 0; i < kDesc; i++) { const float z = dec_z(enc[i]); desc[i] = z == kNoPoint ? 0.0 : (double)z; }
  for (int r = 0; r < kRing; r++) { ringkey[r] = ring_mean(desc, r); ringkey_f[r] = (float)ringkey[r]; }
  for (int s = 0; s < kSector; s++) { vkey[s] = sector_mean(desc, s); norm[s] = sector_norm(desc, s); }
}

// sc_keys_kernel x 2 + sc_pair_kernel for one pair, lanes emulated
void h_sc_distance(const double* a, const double* b, double* dist, int* shift) {
  double avk[kSector], an[kSector], bvk[kSector], bn[kSector];
  for (int s = 0; s < kSector; s++) { avk[s] = sector_mean(a, s); an[s] = sector_norm(a, s); bvk[s] = sector_mean(b, s); bn[s] = sector_norm(b, s); }
  double bd[32]; int bs[32];
  for (int lane = 0; lane < 32; lane++) {
    bd[lane] = vkey_diff_norm(avk, bvk, lane); bs[lane] = lane;
    if (lane + 32 < kSector) { const double d = vkey_diff_norm(avk, bvk, lane + 32); if (d < bd[lane]) { bd[lane] = d; bs[lane] = lane + 32; } }
  }
  for (int o = 16; o > 0; o >>= 1) {   // the xor butterfly, all lanes
    double nd[32]; int ns[32];
    for (int lane = 0; lane < 32; lane++) {
      const double od = bd[lane ^ o]; const int os = bs[lane ^ o];
      nd[lane] = bd[lane]; ns[lane] = bs[lane];
      if (od < bd[lane] || (od == bd[lane] && os < bs[lane])) { nd[lane] = od; ns[lane] = os; }
    }
    for (int lane = 0; lane < 32; lane++) { bd[lane] = nd[lane]; bs[lane] = ns[lane]; }
  }
  int space[2 * kSearchRadius + 1];
  search_space(bs[0], space);
  double best = kBig; int arg = 0;
  for (int t = 0; t < 2 * kSearchRadius + 1; t++) {
    double sim[kSector]; bool ok[kSector];
    for (int j = 0; j < kSector; j++) { sim[j] = 0.0; ok[j] = sector_similarity(a, an, b, bn, j, space[t], &sim[j]); }
    double sum = 0.0; int num = 0;
    for (int j = 0; j < kSector; j++) if (ok[j]) { sum = sum + sim[j]; num = num + 1; }
    const double d = 1.0 - sum / (double)num;
    if (d < best) { best = d; arg = space[t]; }
  }
  *dist = best; *shift = arg;
}

// sc_ring_knn_kernel: distances, then rounds of argmin by (d2, index)
int h_sc_ring_knn(const float* keys, int n, const float* q, int* idx, float* d2out) {
  std::vector<float> row(n);
  for (int i = 0; i < n; i++) row[i] = ring_d2(q, keys + (size_t)i * kRing);
  const int k = n < kCand ? n : kCand;
  for (int r = 0; r < k; r++) {
    float bd = 3.0e38f; int bi = 0x7fffffff;
    for (int i = 0; i < n; i++) if (row[i] < bd || (row[i] == bd && i < bi)) { bd = row[i]; bi = i; }
    idx[r] = bi == 0x7fffffff ? -1 : bi; d2out[r] = bd;
    if (idx[r] >= 0) row[idx[r]] = __builtin_inff();
  }
  return k;
}

float h_sc_yaw(int shift) { return shift_to_yaw(shift); }

}  // extern "C"
