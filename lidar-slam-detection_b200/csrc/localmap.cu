// localmap.cu — local-map assembly for localisation (row N2): key frames near the pose -> one cloud -> voxel grid.
//
// Replaces Localization::runUpdateLocalMap (reference: slam/localization/src/localization.cpp:303-373):
//   radius search (30 m) over key-frame positions, nearest first; a key frame is appended unless the map already has
//   points and it is less than key_frame_distance farther than the last appended one; stop at 200 000 points;
//   VoxelGrid at max(resolution, 0.1); nearest key frame >= 20 m away -> no local map.
// Key-frame clouds (mTransfromPoints: already in the map frame) stay resident on the device; the selection walks a
// few hundred positions on the host, the concatenation is device-to-device copies and the downsample is the K1
// voxel grid — the result never leaves HBM and goes straight into lsd_reg_set_target_dev.
#include <math.h>

#include <algorithm>
#include <vector>

#include "map.h"
#include "voxelgrid.h"

struct lsd_localmap {
  int device = 0;
  double resolution = 0.5, key_frame_distance = 1.0;
  double radius = 30.0, far_sq = 400.0;      // radius_distance_threshold; `pointDistance[0] >= 400` (20 m)
  int max_points = 200000, min_points = 1000;
  struct KeyFrame { float4* d_pts; int n; double pos[3]; };
  std::vector<KeyFrame> frames;
  lsd_voxelgrid* vg = nullptr;
  float4 *d_concat = nullptr, *d_out = nullptr;
  size_t concat_cap = 0;
  int* d_m = nullptr;
  int n_out = 0;
  bool valid = false;
};

using namespace lsd;

extern "C" {

lsd_status_t lsd_localmap_create(lsd_localmap_t** out, double resolution, double key_frame_distance) {
  if (!out || resolution <= 0 || key_frame_distance < 0) return LSD_ERR_INVALID;
  lsd_status_t s = ensure_device();
  if (s) return s;
  lsd_localmap* h = new lsd_localmap();
  cudaGetDevice(&h->device);
  h->resolution = std::max(resolution, 0.1);  // "< 0.1m will cause the filter index overflow", localization.cpp:311
  h->key_frame_distance = key_frame_distance;
  *out = h;
  return LSD_OK;
}

lsd_status_t lsd_localmap_destroy(lsd_localmap_t* h) {
  if (!h) return LSD_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  for (auto& k : h->frames) cudaFree(k.d_pts);
  cudaFree(h->d_concat); cudaFree(h->d_out); cudaFree(h->d_m);
  if (h->vg) lsd_voxelgrid_destroy(h->vg);
  delete h;
  return LSD_OK;
}

// mKeyFrames.push_back + one more point in the graph k-d tree: cloud in the MAP frame, key-frame position
lsd_status_t lsd_localmap_add_keyframe(lsd_localmap_t* h, const float* xyzi_map_host, int n, const double* position3) {
  if (!h || n < 0 || (n > 0 && !xyzi_map_host) || !position3) return LSD_ERR_INVALID;
  LSD_CUDA(cudaSetDevice(h->device));
  lsd_localmap::KeyFrame k;
  k.d_pts = nullptr; k.n = n;
  for (int i = 0; i < 3; i++) k.pos[i] = position3[i];
  if (n > 0) {
    LSD_CUDA(cudaMalloc((void**)&k.d_pts, (size_t)n * 16));
    LSD_CUDA(cudaMemcpy(k.d_pts, xyzi_map_host, (size_t)n * 16, cudaMemcpyHostToDevice));
  }
  h->frames.push_back(k);
  return LSD_OK;
}

// One pass of the update loop body for a pose that moved far enough (the caller keeps the 10 m hysteresis).
// Returns LSD_OK with the local map resident on the device, LSD_LOCALMAP_NONE when the reference would set
// mLocalMap = nullptr (no key frame within 30 m, or the nearest one >= 20 m away).  *n_points may be below the
// reference's 1000-point warning threshold; that is reported, not an error, as in the reference.
lsd_status_t lsd_localmap_update(lsd_localmap_t* h, const double* pose_xyz, int* n_points, int* n_keyframes_in_radius, double* nearest_dist) {
  if (!h || !pose_xyz) return LSD_ERR_INVALID;
  LSD_CUDA(cudaSetDevice(h->device));
  h->valid = false; h->n_out = 0;
  if (n_points) *n_points = 0;
  // radiusSearch(searchPoint, 30 m): float squared distances, sorted ascending (ties: lower index first)
  std::vector<std::pair<float, int>> hits;
  float best = 3.0e38f;
  for (size_t i = 0; i < h->frames.size(); i++) {
    const float dx = (float)h->frames[i].pos[0] - (float)pose_xyz[0], dy = (float)h->frames[i].pos[1] - (float)pose_xyz[1],
                dz = (float)h->frames[i].pos[2] - (float)pose_xyz[2];
    const float d2 = dx * dx + dy * dy + dz * dz;
    best = std::min(best, d2);
    if (d2 < (float)(h->radius * h->radius)) hits.push_back({d2, (int)i});
  }
  std::sort(hits.begin(), hits.end());
  if (n_keyframes_in_radius) *n_keyframes_in_radius = (int)hits.size();
  if (nearest_dist) *nearest_dist = h->frames.empty() ? -1.0 : sqrt((double)best);
  if (hits.empty()) return LSD_LOCALMAP_NONE;  // "out of map"
  std::vector<int> chosen;
  size_t total = 0;
  float accum = 0.f;
  for (auto& hit : hits) {
    const float distance = sqrtf(hit.first);
    if (total > 0 && (distance - accum) < (float)h->key_frame_distance) continue;
    accum = distance;
    chosen.push_back(hit.second);
    total += (size_t)h->frames[hit.second].n;
    if (total >= (size_t)h->max_points) break;
  }
  if (total == 0) return hits[0].first >= (float)h->far_sq ? LSD_LOCALMAP_NONE : LSD_OK;
  if (total > h->concat_cap) {
    cudaFree(h->d_concat); cudaFree(h->d_out); h->d_concat = h->d_out = nullptr; h->concat_cap = 0;
    if (h->vg) { lsd_voxelgrid_destroy(h->vg); h->vg = nullptr; }
    const size_t cap = total + total / 2 + 1024;
    LSD_CUDA(cudaMalloc((void**)&h->d_concat, cap * 16));
    LSD_CUDA(cudaMalloc((void**)&h->d_out, cap * 16));
    lsd_status_t s = lsd_voxelgrid_create(&h->vg, (int)cap, 28);
    if (s) return s;
    h->concat_cap = cap;
  }
  if (!h->d_m) LSD_CUDA(cudaMalloc((void**)&h->d_m, 64));
  cudaStream_t st = h->vg->stream;
  size_t off = 0;
  for (int id : chosen) {  // *mLocalMap += *(mKeyFrames[i]->mTransfromPoints)
    const auto& k = h->frames[id];
    if (k.n > 0) LSD_CUDA(cudaMemcpyAsync(h->d_concat + off, k.d_pts, (size_t)k.n * 16, cudaMemcpyDeviceToDevice, st));
    off += (size_t)k.n;
  }
  lsd_status_t s = vg_run(h->vg, h->d_concat, (int)total, (float)h->resolution, h->d_out, h->d_m, st);
  if (s) return s;
  int hm[2] = {0, 0};
  LSD_CUDA(cudaMemcpyAsync(&hm[0], h->d_m, sizeof(int), cudaMemcpyDeviceToHost, st));
  LSD_CUDA(cudaMemcpyAsync(&hm[1], &h->vg->grid->status, sizeof(int), cudaMemcpyDeviceToHost, st));
  LSD_CUDA(cudaStreamSynchronize(st));
  if (hm[1] == LSD_ERR_CAPACITY) { set_error("local map: voxel grid exceeds the handle's capacity"); return LSD_ERR_CAPACITY; }
  h->n_out = hm[0];
  if (n_points) *n_points = hm[0];
  if (hits[0].first >= (float)h->far_sq) { h->n_out = 0; if (n_points) *n_points = 0; return LSD_LOCALMAP_NONE; }  // nearest key frame >= 20 m
  h->valid = true;
  return LSD_OK;
}

lsd_status_t lsd_localmap_get_dev(lsd_localmap_t* h, const float** xyzi_dev, int* n) {
  if (!h || !xyzi_dev || !n) return LSD_ERR_INVALID;
  *xyzi_dev = reinterpret_cast<const float*>(h->d_out);
  *n = h->valid ? h->n_out : 0;
  return LSD_OK;
}

lsd_status_t lsd_localmap_get(lsd_localmap_t* h, float* xyzi_host, int cap, int* n) {
  if (!h || !n || (cap > 0 && !xyzi_host)) return LSD_ERR_INVALID;
  LSD_CUDA(cudaSetDevice(h->device));
  *n = h->valid ? h->n_out : 0;
  const int c = std::min(cap, *n);
  if (c > 0) LSD_CUDA(cudaMemcpy(xyzi_host, h->d_out, (size_t)c * 16, cudaMemcpyDeviceToHost));
  return LSD_OK;
}

}  // extern "C"
