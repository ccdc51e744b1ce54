// filters.cu — key-frame cloud filters (row N3): radius-outlier removal + range box, order preserving.
//
// Replaces (reference: slam/src/slam.cpp:104-108,398-410, slam/common/slam_utils.cpp:236-247):
//   pcl::RadiusOutlierRemoval (setRadiusSearch(1.0), setMinNeighborsInRadius(3)): a point stays iff the radius
//     search around it returns MORE than min_neighbors points — the search result includes the point itself
//     (PCL 1.9.1 filters/impl/radius_outlier_removal.hpp: `if (k <= min_pts_radius_) -> outlier`);
//   pointsDistanceFilter(cloud, out, 0, key_frame_range): keeps min < |x| < max and min < |y| < max (a box, strict).
// Both preserve the input order.  The radius count is the hash-voxel stencil search with a counter instead of a
// top-K: the cloud is indexed at voxel size = radius, a thread walks the 27 cells around its point
// and stops as soon as the count exceeds the threshold.
#include "knn.cuh"
#include "map.h"

namespace lsd {


constexpr int kFiltBlock = 256;

// keep[i] decision + order-preserving compaction in ONE kernel: blocks take a ticket, evaluate their 256 points in
// parallel, then publish their running total along a chain (block t waits for block t-1's prefix).
__global__ void __launch_bounds__(kFiltBlock) keyframe_filter_kernel(MapView mv, const float4* __restrict__ pts, int n, float radius_sq,
                                                                     int min_neighbors, float min_range, float max_range, int use_radius,
                                                                     float4* __restrict__ out, unsigned* __restrict__ ticket,
                                                                     unsigned long long* __restrict__ prefix, int* __restrict__ n_out) {
  __shared__ unsigned s_ticket, s_base, s_warp[kFiltBlock / 32];
  if (threadIdx.x == 0) s_ticket = atomicAdd(ticket, 1u);
  __syncthreads();
  const unsigned b = s_ticket;
  const int i = (int)(b * kFiltBlock + threadIdx.x);
  bool keep = false;
  float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n) {
    p = __ldg(pts + i);
    const float ax = fabsf(p.x), ay = fabsf(p.y);
    keep = ax > min_range && ax < max_range && ay > min_range && ay < max_range;
    if (keep && use_radius) {
      const int3 c = pos2grid(p.x, p.y, p.z, mv.inv_res);
      int count = 0;
      for (int o = 0; o < 27 && count <= min_neighbors; o++) {
        const int x = c.x + o / 9 - 1, y = c.y + (o / 3) % 3 - 1, z = c.z + o % 3 - 1;
        if (!coord_ok(x, y, z)) continue;
        const unsigned long long key = pack_key(x, y, z, 0);
        uint4 h;
        const CellLine* ln = tag_find(mv, key, &h);
        if (!ln) continue;
        const unsigned cnt = h.z;
        const int levels = cnt > (unsigned)kPtsPerLine ? min((int)((cnt - 1) / kPtsPerLine), kMaxLevel) : 0;
        for (int L = 0; L <= levels && count <= min_neighbors; L++) {
          const CellLine* ll = ln;
          if (L > 0) { uint4 hl; ll = tag_find(mv, key | ((unsigned long long)L << 57), &hl); if (!ll) continue; }
          const int m = (int)min(cnt - (unsigned)(L * kPtsPerLine), (unsigned)kPtsPerLine);
          for (int k = 0; k < m; k++) {
            const float4 a = ldg_f4(&ll->pts[k]);
            if (dist2(p.x, p.y, p.z, a.x, a.y, a.z) < radius_sq) count++;
          }
        }
      }
      keep = count > min_neighbors;
    }
  }
  const unsigned bal = __ballot_sync(kFull, keep);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned tot = 0;
    for (int w = 0; w < kFiltBlock / 32; w++) { const unsigned c = s_warp[w]; s_warp[w] = tot; tot += c; }
    unsigned long long base = 0ull;
    if (b > 0) {
      volatile unsigned long long* pp = prefix + (b - 1);
      unsigned long long v;
      while (((v = *pp) >> 63) == 0ull) {}
      base = v & 0x7fffffffffffffffull;
    }
    const unsigned long long incl = base + tot;
    __threadfence();
    prefix[b] = incl | (1ull << 63);
    s_base = (unsigned)base;
    if ((long long)(b + 1) * kFiltBlock >= (long long)n) *n_out = (int)incl;   // the last block
  }
  __syncthreads();
  if (keep) out[s_base + s_warp[warp] + __popc(bal & ((1u << lane) - 1u))] = p;
}

}  // namespace lsd

using namespace lsd;

extern "C" {

// dev pointers in/out; *n_out (host) is written after a stream synchronise
lsd_status_t lsd_keyframe_filter_dev(const float* xyzi_dev, int n, float radius, int min_neighbors, float min_range, float max_range,
                                     float* out_dev, int* n_out) {
  if (n < 0 || !n_out || (n > 0 && (!xyzi_dev || !out_dev)) || radius < 0.f) return LSD_ERR_INVALID;
  *n_out = 0;
  if (n == 0) return LSD_OK;
  lsd_status_t s = ensure_device();
  if (s) return s;
  const bool use_radius = radius > 0.f;
  lsd_map* m = nullptr;
  int l2 = 12;
  while (l2 < 26 && (1ull << l2) < (unsigned long long)n * 2ull) l2++;
  s = lsd_map_create(&m, use_radius ? radius : 1.0f, use_radius ? l2 : 10);
  if (s) return s;
  cudaStream_t st = m->stream;
  if (use_radius) { s = launch_insert(m, reinterpret_cast<const float4*>(xyzi_dev), n, 0, st); if (s) { lsd_map_destroy(m); return s; } }
  const int nb = (n + kFiltBlock - 1) / kFiltBlock;
  void* scratch = nullptr;
  cudaError_t e = cudaMalloc(&scratch, (size_t)nb * 8 + 64);
  if (e == cudaSuccess) e = cudaMemsetAsync(scratch, 0, (size_t)nb * 8 + 64, st);
  if (e != cudaSuccess) { lsd_map_destroy(m); return cuda_fail(e, "lsd_keyframe_filter scratch", __FILE__, __LINE__); }
  unsigned* ticket = reinterpret_cast<unsigned*>(scratch);
  int* d_nout = reinterpret_cast<int*>(scratch) + 2;
  unsigned long long* prefix = reinterpret_cast<unsigned long long*>(static_cast<char*>(scratch) + 64);
  keyframe_filter_kernel<<<nb, kFiltBlock, 0, st>>>(m->view, reinterpret_cast<const float4*>(xyzi_dev), n, radius * radius, min_neighbors, min_range,
                                                    max_range, use_radius ? 1 : 0, reinterpret_cast<float4*>(out_dev), ticket, prefix, d_nout);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaMemcpyAsync(n_out, d_nout, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(scratch);
  lsd_map_destroy(m);
  if (e != cudaSuccess) return cuda_fail(e, "lsd_keyframe_filter", __FILE__, __LINE__);
  return LSD_OK;
}

lsd_status_t lsd_keyframe_filter(const float* xyzi_host, int n, float radius, int min_neighbors, float min_range, float max_range,
                                 float* out_host, int* n_out) {
  if (n < 0 || !n_out || (n > 0 && (!xyzi_host || !out_host))) return LSD_ERR_INVALID;
  *n_out = 0;
  if (n == 0) return LSD_OK;
  lsd_status_t s = ensure_device();
  if (s) return s;
  float4 *d_in = nullptr, *d_out = nullptr;
  LSD_CUDA(cudaMalloc((void**)&d_in, (size_t)n * 16));
  cudaError_t e = cudaMalloc((void**)&d_out, (size_t)n * 16);
  if (e != cudaSuccess) { cudaFree(d_in); return cuda_fail(e, "lsd_keyframe_filter alloc", __FILE__, __LINE__); }
  e = cudaMemcpy(d_in, xyzi_host, (size_t)n * 16, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    s = lsd_keyframe_filter_dev(reinterpret_cast<const float*>(d_in), n, radius, min_neighbors, min_range, max_range, reinterpret_cast<float*>(d_out), n_out);
    if (s == LSD_OK && *n_out > 0) e = cudaMemcpy(out_host, d_out, (size_t)*n_out * 16, cudaMemcpyDeviceToHost);
  }
  cudaFree(d_in); cudaFree(d_out);
  if (e != cudaSuccess) return cuda_fail(e, "lsd_keyframe_filter copy", __FILE__, __LINE__);
  return s;
}

}  // extern "C"
