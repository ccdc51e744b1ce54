// map.h — host-side handle of the hash-voxel map.
#pragma once
#include <algorithm>

#include "lsd_common.cuh"

struct lsd_map {
  lsd::MapView view{};
  unsigned long long n_lines = 0;
  int device = 0;
  cudaStream_t stream = nullptr;
  void* scratch = nullptr;  // staging for the host-pointer entry points
  size_t scratch_bytes = 0;
  long long launches = 0;   // kernels launched on behalf of this handle
  int knn_shape = 0;        // lsd_knn_set_shape: 0 auto, 1 warp/query, 2 thread/query, 3 brick pages (TMA)
  // brick layout (lsd_map_enable_bricks, brick.cuh): directory size and the per-batch binning scratch
  unsigned long long n_bricks = 0;
  unsigned* bin_count = nullptr;   // [n_bricks], zero between batches
  int* bin_base = nullptr;         // [n_bricks]
  void* bscratch = nullptr;        // q_slot, q_rank, sorted, work items, counters of the batch in flight
  size_t bscratch_bytes = 0;
};

namespace lsd {
lsd_status_t upload_stencils();
lsd_status_t launch_insert(lsd_map* m, const float4* d_pts, int n, int id0, cudaStream_t st);
lsd_status_t launch_knn(lsd_map* m, const float4* d_q, int nq, int k, float max_sq, int stencil, int* d_idx, float* d_d2,
                        int* d_cnt, cudaStream_t st);
__device__ void map_insert_point(const MapView& mv, float x, float y, float z, int id);
}  // namespace lsd
