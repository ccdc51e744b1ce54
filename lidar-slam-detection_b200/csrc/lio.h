// lio.h — host-side handle of the LIO front-end.
#pragma once
#include <vector>

#include "lsd_common.cuh"

struct lsd_map;
struct lsd_voxelgrid;
namespace lsd {
// Tile-sharded mode: peers' device buffers, reachable with plain stores/loads over
// NVLink (cudaIpc-opened, or the same process' own pointers).  inbox[p] / flagbox[p] belong to rank p.
constexpr int kMaxRanks = 8;
constexpr int kInboxSlot = 32;                                  // doubles per (parity, source rank)
constexpr int kInboxRegion = 2 * kMaxRanks * kInboxSlot;        // h-model region, then the degeneracy region
struct ShardComm {
  int rank, world;
  double* inbox[kMaxRanks];
  unsigned char* flagbox[kMaxRanks];   // [cap] map_incremental decisions, then kMaxRanks u64 "done" sequence numbers
};
}  // namespace lsd

struct lsd_lio {
  lsd_lio_params_t p{};
  int device = 0;
  lsd_map* map = nullptr;       // hash-voxel map (iVox replacement), owned
  lsd_voxelgrid* vg = nullptr;  // scan downsampler, owned
  cudaStream_t stream = nullptr;
  // raw scan staging (host-pointer entry points): two slots, so lsd_lio_prefetch can upload scan k+1
  // while scan k (staged in the other slot) is being registered
  struct Stage {
    float4* buf = nullptr; const float* host = nullptr; int n = 0; bool valid = false; cudaEvent_t ev = nullptr; long long age = 0;
    // pipelined voxel grid (lsd_lio_set_pipeline): the staged scan, already downsampled on the copy stream
    float4* body = nullptr; int* dn = nullptr; bool down = false; bool is_dev = false;
  };
  Stage stage[2];
  long long stage_clock = 0;
  cudaStream_t copy_stream = nullptr;
  int busy_slot = -1;                 // slot the scan being registered reads (-1: a device-pointer scan)
  const float* defer_host = nullptr;  // prefetch request whose copy is issued from inside the next lsd_lio_scan,
  int defer_n = 0;                    // after that scan's first kernels are in flight
  bool defer_pending = false;
  bool defer_is_dev = false;          // the request names a device-resident scan (lsd_lio_prefetch_dev): no copy, voxel grid only
  // pipelined voxel grid: the downsample of scan k+1 depends on nothing scan k computes, so with the flag on it runs on
  // the copy stream (after the H2D copy) while scan k iterates, and lsd_lio_scan(k+1) starts at its first search
  int pipeline_vg = 0;                // lsd_lio_set_pipeline
  long long side_vg_issued = 0, side_vg_adopted = 0;   // lsd_lio_pipeline_stats
  Stage* pre = nullptr;               // stage whose downsampled scan the lio_scan in progress adopts (buffers are swapped)
  bool issue_in_linearize = false;    // deferred prefetch to be issued behind the first evaluation's kernels
  bool main_ev_fresh = false;         // ev_main_vg was recorded by the lio_scan in progress, right behind its load step
  bool side_vg_inflight = false;      // a voxel grid is queued on the copy stream (its scratch is shared with the main stream's)
  cudaEvent_t ev_main_vg = nullptr, ev_side_vg = nullptr;
  int* d_rows = nullptr;              // rows alive in Nearest_Points, double-buffered by scan parity (lio_resize_rows_kernel)
  float4* d_body = nullptr;     // feats_down_body
  int* d_n = nullptr;           // feats_down_size, device resident
  float4* d_near = nullptr;     // Nearest_Points: [n,5] (x, y, z, id)
  int* d_near_cnt = nullptr;
  int knn_shape = 0;            // lsd_lio_set_knn_shape: 0/1 warp per scan point (lio_knn_kernel)
  double xchg_cycles = 0.0; long long xchg_count = 0;   // tile-sharded: SM cycles spent in the in-kernel exchange (lsd_lio_shard_exchange_stats)
  double travel = 0.0, last_pos_lid[3] = {0, 0, 0};   // travel_distance / last_pos_lid (laserMapping.cpp:96,145,1289-1291)
  int rc_ctas = 8, rc_threads = 256;   // its cluster shape (LSD_REUSE_CLUSTER=CxT)
  int reuse_cluster = 0;        // reuse evaluations as ONE thread-block cluster with a DSMEM reduction (lio_hmodel_reuse_cluster_kernel)
  int pdl = 0;                  // lsd_lio_set_pdl: launch the scan's kernels with programmatic dependent launch
  int rows_parity = 0;          // which of d_rows[0..1] the next Nearest_Points.resize publishes (the other one is read)
  bool rows_resize_pending = false;   // the loaded scan's first search still has to do Nearest_Points.resize (lio_knn_kernel)
  int reference_order = 0;      // lsd_lio_set_reference_order: neighbours in the order IVox::GetClosestPoint returns them (lio.cu RefCand)
  unsigned* d_ref_fallbacks = nullptr;   // its counter of queries answered in (d2, id) order (the search's list overflowed)
  bool stale_rows = true;       // lsd_lio_set_stale_rows: keep Nearest_Points[i] when a search finds nothing, as the reference does
  double wait_timeout_s = 20.0; // wall-clock bound on the host's wait for a published reduction (wait_seq)
  unsigned char* d_selected = nullptr;  // point_selected_surf
  unsigned char* d_flags = nullptr;     // map_incremental decision per point
  float4* d_plane = nullptr;    // normvec: (normal, pd2)
  float4* d_pabcd = nullptr;    // cached plane (a, b, c, d) of the last neighbour search
  unsigned char* d_plane_ok = nullptr;  // esti_plane accepted the 5 neighbours
  float4* d_world = nullptr;    // feats_down_world
  double* d_partials = nullptr;
  unsigned* d_done = nullptr;
  unsigned* d_added = nullptr;
  double* h_result = nullptr;   // pinned + mapped: the reduction result the kernels publish
  double* d_result = nullptr;   // device alias of h_result
  unsigned* h_added = nullptr;  // pinned: insert count of the last map_incremental
  bool pending = false;         // an async lio_scan left its end event / insert count uncollected
  double last_gpu_ms = 0.0;
  int last_added = 0;
  long long seq = 0;            // sequence number of the last published result
  long long mi_seq = 0;         // sequence number of the last map_incremental (halo exchange)
  // tile-sharded mode (lsd_lio_shard_*): peers' inboxes / flag boxes
  int shard_rank = 0, shard_world = 1;
  double* d_inbox = nullptr;
  unsigned char* d_flagbox = nullptr;
  std::vector<void*> ipc_opened;
  lsd::ShardComm sc;
  int max_search_blocks = 888;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int n_bound = 1;   // launch bound for per-point kernels (>= true feats_down_size)
  int n_down = -1;   // feats_down_size once known on the host
  int next_id = 0;   // id of the next inserted map point
  int ekf_inited = 1;
  long long launches = 0;
  uint64_t map_cells_known = 0;
  double last_Pm[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  // optional per-kernel timing (bench roofline): 0 hmodel<search>, 1 hmodel<reuse>, 2 voxel grid, 3 map_incremental
  int profile = 0;
  cudaEvent_t pev[2] = {nullptr, nullptr};
  double prof_ms[4] = {0, 0, 0, 0};
  long long prof_cnt[4] = {0, 0, 0, 0};
};
