// knn.cuh — k-NN over the hash-voxel map (K3).  One thread per query: neighbouring threads hold
// neighbouring queries (the voxel-grid output is ordered by leaf index), so a warp's stencil
// probes hit the same 128-byte cell lines and coalesce in L1.
//
// Replaces IVox::GetClosestPoint (ivox3d.h:139-171) + IVoxNode::KNNPointByCondition
// (ivox3d_node.hpp:107-127), and — with the EXACT shell search — KD_TREE::Nearest_Search
// (ikd_Tree.cpp:367-397).  Result order is canonical: ascending (d2, id).
#pragma once
#include "lsd_common.cuh"

namespace lsd {

// Unity build (lsdreg.cu includes every .cu): one definition, visible to all kernels.
__constant__ Stencil c_stencils[5];  // CENTER, NEARBY6, NEARBY18, NEARBY26, NEARBY74
__host__ __device__ __forceinline__ int stencil_slot(int type) {
  return type == LSD_STENCIL_CENTER ? 0 : type == LSD_STENCIL_NEARBY6 ? 1 : type == LSD_STENCIL_NEARBY18 ? 2
       : type == LSD_STENCIL_NEARBY26 ? 3 : type == LSD_STENCIL_NEARBY74 ? 4 : -1;
}

template <int K>
struct TopK {
  float d[K];
  int id[K];
  unsigned loc[K];  // line * 8 + slot (1..7): where the point lives, to re-read its coordinates
  int n;
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < K; j++) { d[j] = __int_as_float(0x7f800000); id[j] = 0x7fffffff; loc[j] = 0; }
    n = 0;
  }
  // branch-free sorted insert: bubble the candidate through the K slots
  __device__ __forceinline__ void insert(float cd, int cid, unsigned cloc) {
    if (!(cd < d[K - 1] || (cd == d[K - 1] && cid < id[K - 1]))) return;
#pragma unroll
    for (int j = 0; j < K; j++) {
      bool lt = cd < d[j] || (cd == d[j] && cid < id[j]);
      float td = lt ? d[j] : cd; int ti = lt ? id[j] : cid; unsigned tl = lt ? loc[j] : cloc;
      d[j] = lt ? cd : d[j]; id[j] = lt ? cid : id[j]; loc[j] = lt ? cloc : loc[j];
      cd = td; cid = ti; cloc = tl;
    }
    n = n < K ? n + 1 : K;
  }
};

template <int K>
__device__ __forceinline__ void consider(TopK<K>& tk, const float4& p, unsigned loc, float qx, float qy, float qz,
                                         float max_sq, bool inclusive) {
  float d2 = dist2(qx, qy, qz, p.x, p.y, p.z);
  bool ok = inclusive ? (d2 <= max_sq) : (d2 < max_sq);
  if (ok) tk.insert(d2, __float_as_int(p.w), loc);
}

// Scan points [first, min(count,7)) of a matched level-0 line, then any overflow levels
// (lines keyed (voxel, L) holding points 7L..7L+6).  Runtime loops: one inlined copy of insert().
template <int K>
__device__ __forceinline__ void scan_line(const MapView& mv, const CellLine* ln, unsigned long long s,
                                          unsigned long long key, unsigned count, int first, TopK<K>& tk, float qx,
                                          float qy, float qz, float max_sq, bool inclusive) {
  const int n0 = (int)min(count, (unsigned)kPtsPerLine);
#pragma unroll 1
  for (int j = first; j < n0; j++)
    consider(tk, ldg_f4(&ln->pts[j]), (unsigned)(s * 8 + j + 1), qx, qy, qz, max_sq, inclusive);
  if (count <= (unsigned)kPtsPerLine) return;
  const int levels = min((int)((count - 1) / kPtsPerLine), kMaxLevel);
#pragma unroll 1
  for (int L = 1; L <= levels; L++) {
    const unsigned long long kl = key | ((unsigned long long)L << 57);
    unsigned long long sl = hash_key(kl) & mv.mask;
#pragma unroll 1
    for (unsigned probe = 0; probe < kMaxProbe; probe++) {
      const CellLine* l2 = mv.lines + sl;
      const uint4 h = ldg_u4(l2);
      const unsigned long long k = (unsigned long long)h.x | ((unsigned long long)h.y << 32);
      if (k == kl) {
        const int n = (int)min(count - (unsigned)(L * kPtsPerLine), (unsigned)kPtsPerLine);
#pragma unroll 1
        for (int j = 0; j < n; j++)
          consider(tk, ldg_f4(&l2->pts[j]), (unsigned)(sl * 8 + j + 1), qx, qy, qz, max_sq, inclusive);
        break;
      }
      if (k == 0) break;
      sl = (sl + 1) & mv.mask;
    }
  }
}

// Generic (slow-path) voxel lookup: linear probe from `s`, scan every point of the voxel.
template <int K>
__device__ __forceinline__ void scan_voxel(const MapView& mv, unsigned long long key, unsigned long long s, TopK<K>& tk,
                                           float qx, float qy, float qz, float max_sq, bool inclusive) {
#pragma unroll 1
  for (unsigned probe = 0; probe < kMaxProbe; probe++) {
    const CellLine* ln = mv.lines + s;
    const uint4 h = ldg_u4(ln);
    const unsigned long long k = (unsigned long long)h.x | ((unsigned long long)h.y << 32);
    if (k == key) { scan_line(mv, ln, s, key, h.z, 0, tk, qx, qy, qz, max_sq, inclusive); return; }
    if (k == 0) return;
    s = (s + 1) & mv.mask;
  }
}

// Fixed-stencil search.  Chunks of CH cells: all CH header+first-point sectors are requested
// before any is consumed (CH independent 32-byte sector loads in flight per thread).  Voxels whose
// home slot is taken by another voxel, or that hold more than one point, are finished in runtime
// loops afterwards so the unrolled fast path stays small.
template <int K, int CH = 8>
__device__ __forceinline__ void knn_stencil(const MapView& mv, int st_slot, float qx, float qy, float qz, float max_sq,
                                            TopK<K>& tk) {
  const Stencil& st = c_stencils[st_slot];
  const int3 c = pos2grid(qx, qy, qz, mv.inv_res);
#pragma unroll 1
  for (int c0 = 0; c0 < st.n; c0 += CH) {
    uint4 h[CH]; float4 p0[CH];
    unsigned valid = 0;
#pragma unroll
    for (int u = 0; u < CH; u++) {
      if (c0 + u < st.n) {
        const int x = c.x + st.off[c0 + u][0], y = c.y + st.off[c0 + u][1], z = c.z + st.off[c0 + u][2];
        if (coord_ok(x, y, z)) {
          const unsigned long long key = pack_key(x, y, z, 0);
          const CellLine* ln = mv.lines + (hash_key(key) & mv.mask);
          h[u] = ldg_u4(ln);
          p0[u] = ldg_f4(&ln->pts[0]);
          valid |= 1u << u;
        }
      }
    }
    unsigned more = 0;  // bit u: voxel needs the runtime path (collision chain or > 1 point)
#pragma unroll
    for (int u = 0; u < CH; u++) {
      if (!(valid >> u & 1)) continue;
      const int x = c.x + st.off[c0 + u][0], y = c.y + st.off[c0 + u][1], z = c.z + st.off[c0 + u][2];
      const unsigned long long key = pack_key(x, y, z, 0);
      const unsigned long long k = (unsigned long long)h[u].x | ((unsigned long long)h[u].y << 32);
      if (k == key) {
        if (h[u].z > 0) consider(tk, p0[u], (unsigned)((hash_key(key) & mv.mask) * 8 + 1), qx, qy, qz, max_sq, false);
        if (h[u].z > 1) more |= 1u << u;
      } else if (k != 0) {
        more |= 1u << (u + 16);
      }
    }
#pragma unroll 1
    while (more) {
      const int b = __ffs(more) - 1;
      more &= more - 1;
      const int u = b & 15;
      const int x = c.x + st.off[c0 + u][0], y = c.y + st.off[c0 + u][1], z = c.z + st.off[c0 + u][2];
      const unsigned long long key = pack_key(x, y, z, 0);
      const unsigned long long s = hash_key(key) & mv.mask;
      if (b < 16) {  // matched at home slot: points 1.. (header re-read hits L1)
        const CellLine* ln = mv.lines + s;
        scan_line(mv, ln, s, key, ldg_u4(ln).z, 1, tk, qx, qy, qz, max_sq, false);
      } else {
        scan_voxel(mv, key, (s + 1) & mv.mask, tk, qx, qy, qz, max_sq, false);
      }
    }
  }
}

// Exact k-NN with d2 <= max_sq by Chebyshev shells around the query's voxel; stops as soon as the
// K-th best distance is below the lower bound of every unseen shell.
template <int K>
__device__ __forceinline__ void knn_exact(const MapView& mv, float qx, float qy, float qz, float max_sq, TopK<K>& tk) {
  const int3 c = pos2grid(qx, qy, qz, mv.inv_res);
  const int rmax = (int)ceilf(sqrtf(max_sq) * mv.inv_res) + 1;
#pragma unroll 1
  for (int r = 0; r <= rmax; r++) {
    if (r >= 1 && tk.n == K) {
      // a point in shell >= r is at least (r-1)*res away along one axis (voxels are centred on
      // integer multiples of res, the query is anywhere inside its own voxel)
      const float lo = (float)(r - 1) * mv.res * (1.0f - 1e-5f);
      if (tk.d[K - 1] < lo * lo) break;
    }
    const int side = 2 * r + 1;
    const int ncell = side * side * side;
#pragma unroll 1
    for (int t = 0; t < ncell; t++) {
      const int i = t / (side * side) - r, j = (t / side) % side - r, l = t % side - r;
      if (max(max(abs(i), abs(j)), abs(l)) != r) continue;
      const int x = c.x + i, y = c.y + j, z = c.z + l;
      if (!coord_ok(x, y, z)) continue;
      const unsigned long long key = pack_key(x, y, z, 0);
      scan_voxel(mv, key, hash_key(key) & mv.mask, tk, qx, qy, qz, max_sq, true);
    }
  }
}

template <int K>
__device__ __forceinline__ void knn_search(const MapView& mv, int stencil_type, float qx, float qy, float qz, float max_sq,
                                           TopK<K>& tk) {
  tk.init();
  if (stencil_type == LSD_STENCIL_EXACT) knn_exact<K>(mv, qx, qy, qz, max_sq, tk);
  else knn_stencil<K>(mv, stencil_slot(stencil_type), qx, qy, qz, max_sq, tk);
}

// coordinates of a stored neighbour from its location code
__device__ __forceinline__ float4 load_loc(const MapView& mv, unsigned loc) {
  const CellLine* ln = mv.lines + (loc >> 3);
  return ldg_f4(&ln->pts[(loc & 7) - 1]);
}

}  // namespace lsd
