// voxelgrid.h — host-side handle of the voxel-grid downsampler (K1).
#pragma once
#include <algorithm>

#include "lsd_common.cuh"

namespace lsd {
struct VgGrid {  // device-resident grid description, written by vg_setup_kernel
  float inv;
  int minb[3], divb[3], mul[3];
  int n_words;
  int status;  // 0, LSD_ERR_GRID_OVERFLOW (PCL: output = input) or LSD_ERR_CAPACITY
};
}  // namespace lsd

struct lsd_voxelgrid {
  int device = 0, max_points = 0, scan_blocks = 0, coop_blocks = 148;
  long long max_cells = 0;
  cudaStream_t stream = nullptr;
  int* bbox = nullptr;            // ordered-int min[3], max[3]
  lsd::VgGrid* grid = nullptr;
  unsigned* bitmap = nullptr;     // occupancy, 1 bit per leaf; all-zero between calls
  int* word_prefix = nullptr;     // exclusive popcount prefix inside a scan chunk
  int* chunk_sum = nullptr;       // chunk totals -> exclusive chunk offsets
  int* vidx = nullptr;            // leaf index per input point
  int* out_vidx = nullptr;        // leaf index per output point
  long long* sums = nullptr;      // [max_points,4] fixed-point channel sums; all-zero between calls
  int* cnt = nullptr;
  int *off = nullptr, *cur = nullptr, *seg = nullptr, *ord = nullptr;   // input-order sums: leaf offsets, scatter cursors (zero between calls), members, members in input order
  int input_order_sums = 1;       // centroids as pcl::VoxelGrid's sequential fp32 sums in input order (LSD_VG_SUMS=fixed: exact fixed-point sums)
  float4 *io_in = nullptr, *io_out = nullptr;  // staging for the host-pointer entry point
  int* d_m = nullptr;
  unsigned* done = nullptr;       // last-block ticket of vg_scan_kernel
  long long launches = 0;
  int pdl = 0;                    // launch the five kernels with programmatic dependent launch (lsd_lio_set_pdl)
};

namespace lsd {
lsd_status_t vg_run(lsd_voxelgrid* g, const float4* d_in, int n, float leaf, float4* d_out, int* d_m, cudaStream_t st);
}
