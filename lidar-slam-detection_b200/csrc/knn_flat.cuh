// knn_flat.cuh — work-flattened k-NN over the hash-voxel map (K3, "flat" shape).
//
// Same search as knn.cuh (IVox::GetClosestPoint, ivox3d.h:139-171; KNNPointByCondition, ivox3d_node.hpp:107-127),
// same candidates, same canonical (d2, id) order — only the mapping of work to lanes differs.
//
// Why a third shape.  The thread-per-query kernel (map.cu) spends 235 warp instructions per query with 14 of 32
// lanes active on average (profiles/r01_knn_batch.txt): after each tag probe only the lanes whose voxel exists
// walk its points, the others wait.  Here a warp owns 32 queries and runs three convergent phases per pass:
//   A  every lane probes the TAG array (L2-resident) for the same stencil cell of its own query; voxels that
//      exist are ballot-compacted into a per-warp list of (query lane, cell offset, line) entries in shared memory;
//   B  the warp walks that list 32 entries at a time — every lane fetches one EXISTING voxel's line (header +
//      points 0..2 requested together, two entries in flight per lane), verifies the key, computes distances and
//      appends in-radius points to the owning query's candidate list (shared-memory atomicAdd slot claim);
//   C  every lane folds its own query's candidate list into its register top-K (compare-exchange chain).
// The expensive parts (DRAM line fetch, distance, list append) run with all lanes busy; divergence is left only in
// the tails (entries mod 32, list lengths).  Capacity misses are exact, not approximate: a pass stops taking cells when
// the entry list is nearly full, and a query whose candidate list overflows in a pass re-walks that pass's cells
// serially (the thread-per-query code path).
#pragma once
#include "knn.cuh"

namespace lsd {

constexpr int kFlatEnt = 384;      // per-warp entries staged per pass (a NEARBY18 query has ~7.5 existing voxels)
constexpr int kFlatCand = 24;      // candidate slots per query per pass
constexpr int kFlatCandStride = 25;  // odd stride: lane L reading slot p of its own list hits bank (25 L + p) mod 32

// register top-K in canonical (d2, id) order; LOC also carries the location code of each neighbour
template <int K, bool LOC>
struct FlatTopK {
  float d[K]; int id[K]; unsigned loc[LOC ? K : 1];
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int i = 0; i < K; i++) { d[i] = 3.0e38f; id[i] = 0x7fffffff; if (LOC) loc[i] = 0u; }
  }
  __device__ __forceinline__ void push(float d2, int pid, unsigned l) {
    if (!(d2 < d[K - 1] || (d2 == d[K - 1] && pid < id[K - 1]))) return;
    d[K - 1] = d2; id[K - 1] = pid;
    if (LOC) loc[K - 1] = l;
#pragma unroll
    for (int i = K - 1; i > 0; i--) {
      const bool sw = d[i] < d[i - 1] || (d[i] == d[i - 1] && id[i] < id[i - 1]);
      const float td = d[i]; const int ti = id[i];
      d[i] = sw ? d[i - 1] : d[i]; id[i] = sw ? id[i - 1] : id[i];
      d[i - 1] = sw ? td : d[i - 1]; id[i - 1] = sw ? ti : id[i - 1];
      if (LOC) { const unsigned tl = loc[i]; loc[i] = sw ? loc[i - 1] : loc[i]; loc[i - 1] = sw ? tl : loc[i - 1]; }
    }
  }
};

template <bool LOC>
struct FlatSmem {  // per warp
  float4 q[32];
  unsigned ent_line[kFlatEnt];
  unsigned short ent_who[kFlatEnt];  // lane | (dx+2) << 5 | (dy+2) << 8 | (dz+2) << 11
  unsigned cand_d[32 * kFlatCandStride];
  int cand_id[32 * kFlatCandStride];
  unsigned cand_loc[LOC ? 32 * kFlatCandStride : 1];
  unsigned cand_n[32];
};

// First slot on `key`'s probe sequence whose TAG matches, before the first empty slot (no line touched).
// The key itself is verified by whoever fetches the line (1/255 false positives per occupied slot walked).
__device__ __forceinline__ bool tag_first(const MapView& mv, unsigned long long key, unsigned* line) {
  constexpr unsigned long long k7f = 0x7f7f7f7f7f7f7f7full;
  const unsigned long long hh = hash_key(key);
  const unsigned t4 = slot_tag(hh) * 0x01010101u;                 // the tag in every byte: one 32-bit multiply, then doubled
  const unsigned long long tagv = ((unsigned long long)t4 << 32) | t4;
  unsigned long long s = hh & mv.mask;
  for (unsigned walked = 0; walked < kMaxProbe;) {
    const unsigned pos = (unsigned)(s & 7ull), nv = 8u - pos;
    const unsigned long long v = __ldg(reinterpret_cast<const unsigned long long*>(mv.tags + (s & ~7ull))) >> (8u * pos);
    const unsigned long long ze = ~(((v & k7f) + k7f) | v | k7f);  // bit 7 of every zero byte (bytes shifted in are zero)
    const unsigned long long x = v ^ tagv;
    const unsigned long long zm = ~(((x & k7f) + k7f) | x | k7f);
    const unsigned fe = ze ? (unsigned)(__ffsll((long long)ze) - 1) >> 3 : 8u;
    if (zm) {
      const unsigned fm = (unsigned)(__ffsll((long long)zm) - 1) >> 3;
      if (fm < fe) { *line = (unsigned)((s + fm) & mv.mask); return true; }
    }
    if (fe < nv) return false;
    walked += nv;
    s = (s + nv) & mv.mask;
  }
  return false;
}

// serial walk of one voxel for one query (the thread-per-query body, map.cu::knn_query_thread_kernel)
template <int K, bool LOC>
__device__ __forceinline__ void flat_walk_voxel(const MapView& mv, unsigned long long key, float qx, float qy, float qz,
                                                float max_sq, FlatTopK<K, LOC>& best, int& found) {
  uint4 h;
  const CellLine* ln = tag_find(mv, key, &h);
  if (!ln) return;
  const unsigned cnt = h.z;
  const int levels = cnt > (unsigned)kPtsPerLine ? min((int)((cnt - 1) / kPtsPerLine), kMaxLevel) : 0;
  for (int L = 0; L <= levels; L++) {
    const CellLine* ll = ln;
    if (L > 0) {
      uint4 hl;
      ll = tag_find(mv, key | ((unsigned long long)L << 57), &hl);
      if (!ll) continue;
    }
    const int n = (int)min(cnt - (unsigned)(L * kPtsPerLine), (unsigned)kPtsPerLine);
    const unsigned base = (unsigned)(ll - mv.lines) * 8u + 1u;
    for (int j = 0; j < n; j++) {
      const float4 a = ldg_f4(&ll->pts[j]);
      const float d2 = dist2(qx, qy, qz, a.x, a.y, a.z);
      if (d2 < max_sq) { found++; best.push(d2, __float_as_int(a.w), base + (unsigned)j); }
    }
  }
}

// One warp, 32 queries (lane = query; `active` false for padding lanes, which still take part in the warp-wide steps).
// On return `best` holds the lane's K nearest in canonical order and `found` the number of in-radius points seen.
template <int K, bool LOC>
__device__ __forceinline__ void flat_search(const MapView& mv, const Stencil& st, float qx, float qy, float qz, bool active,
                                            float max_sq, FlatSmem<LOC>& sm, FlatTopK<K, LOC>& best, int& found) {
  const int lane = threadIdx.x & 31;
  const int3 c = pos2grid(qx, qy, qz, mv.inv_res);
  // A stencil cell's key is the home voxel's key plus a per-cell constant as long as no coordinate field can leave its 19
  // bits: true when the home voxel is at least 3 voxels inside the representable cube (stencil offsets are within +-2).
  constexpr int kInner = kCoordBias - 3;
  const bool inner = c.x > -kInner && c.x < kInner && c.y > -kInner && c.y < kInner && c.z > -kInner && c.z < kInner;
  const unsigned long long key_c = inner ? pack_key(c.x, c.y, c.z, 0) : 0ull;
  sm.q[lane] = make_float4(qx, qy, qz, 0.f);
  sm.cand_n[lane] = 0u;
  __syncwarp();
  int o = 0;
  while (o < st.n) {  // one pass per iteration; o and n_ent are warp-uniform
    const int o_begin = o;
    // ---------------- phase A: tag probes, ballot compaction of the voxels that exist
    int n_ent = 0;
    for (; o < st.n && n_ent <= kFlatEnt - 32; o++) {
      const int dx = st.off[o][0], dy = st.off[o][1], dz = st.off[o][2];
      const long long dkey = ((long long)dx << 38) + ((long long)dy << 19) + (long long)dz;   // warp-uniform
      unsigned line = 0u;
      bool hit = false;
      if (active) {
        if (inner) hit = tag_first(mv, key_c + (unsigned long long)dkey, &line);
        else {   // at the rim of the representable cube: per-coordinate range check (queries this far out are rare)
          const int x = c.x + dx, y = c.y + dy, z = c.z + dz;
          if (coord_ok(x, y, z)) hit = tag_first(mv, pack_key(x, y, z, 0), &line);
        }
      }
      const unsigned m = __ballot_sync(kFull, hit);
      if (hit) {
        const int pos = n_ent + __popc(m & lanemask_lt());
        sm.ent_line[pos] = line;
        sm.ent_who[pos] = (unsigned short)(lane | ((dx + 2) << 5) | ((dy + 2) << 8) | ((dz + 2) << 11));
      }
      n_ent += __popc(m);
    }
    __syncwarp();
#ifdef LSD_SIMT_EMU  // occupancy statistics for tuning kFlatEnt / kFlatCand (tests/simt only)
    if (lane == 0) { SIMT_STAT_ADD(1, 1); SIMT_STAT_ADD(2, n_ent); SIMT_STAT_MAX(6, n_ent); }
#endif
    // ---------------- phase B: one existing voxel per lane, two in flight
    struct Ent { const CellLine* cl; unsigned long long key; float4 qp; uint4 h; float4 a0, a1, a2; int ql; };
    auto fetch = [&](int e, Ent& t) {
      const unsigned who = sm.ent_who[e];
      t.ql = (int)(who & 31u);
      t.qp = sm.q[t.ql];
      const int3 qc = pos2grid(t.qp.x, t.qp.y, t.qp.z, mv.inv_res);
      t.key = pack_key(qc.x + (int)((who >> 5) & 7u) - 2, qc.y + (int)((who >> 8) & 7u) - 2, qc.z + (int)((who >> 11) & 7u) - 2, 0);
      t.cl = mv.lines + sm.ent_line[e];
      t.h = ldg_u4(t.cl);
      t.a0 = ldg_f4(&t.cl->pts[0]); t.a1 = ldg_f4(&t.cl->pts[1]); t.a2 = ldg_f4(&t.cl->pts[2]);
    };
    auto offer = [&](const Ent& t, const float4& a, unsigned loc) {
      const float d2 = dist2(t.qp.x, t.qp.y, t.qp.z, a.x, a.y, a.z);
      if (d2 < max_sq) {
        const unsigned pos = atomicAdd(&sm.cand_n[t.ql], 1u);
        if (pos < (unsigned)kFlatCand) {
          const int at = t.ql * kFlatCandStride + (int)pos;
          sm.cand_d[at] = __float_as_uint(d2);
          sm.cand_id[at] = __float_as_int(a.w);
          if (LOC) sm.cand_loc[at] = loc;
        }
      }
    };
    auto consume = [&](Ent& t) {
      if (((unsigned long long)t.h.x | ((unsigned long long)t.h.y << 32)) != t.key) {  // tag collision: resolve properly
#ifdef LSD_SIMT_EMU
        SIMT_STAT_ADD(5, 1);
#endif
        t.cl = tag_find(mv, t.key, &t.h);
        if (!t.cl) return;
        t.a0 = ldg_f4(&t.cl->pts[0]); t.a1 = ldg_f4(&t.cl->pts[1]); t.a2 = ldg_f4(&t.cl->pts[2]);
      }
      const unsigned cnt = t.h.z;
      const unsigned base = (unsigned)(t.cl - mv.lines) * 8u + 1u;
      if (cnt > 0u) offer(t, t.a0, base);
      if (cnt > 1u) offer(t, t.a1, base + 1u);
      if (cnt > 2u) offer(t, t.a2, base + 2u);
      if (cnt > 3u) {
        const unsigned n0 = min(cnt, (unsigned)kPtsPerLine);
        for (unsigned j = 3; j < n0; j++) offer(t, ldg_f4(&t.cl->pts[j]), base + j);
        if (cnt > (unsigned)kPtsPerLine) {  // overflow levels (rare in a 0.5 m-thinned map)
          const int levels = min((int)((cnt - 1) / kPtsPerLine), kMaxLevel);
          for (int L = 1; L <= levels; L++) {
            uint4 hl;
            const CellLine* ll = tag_find(mv, t.key | ((unsigned long long)L << 57), &hl);
            if (!ll) continue;
            const int n = (int)min(cnt - (unsigned)(L * kPtsPerLine), (unsigned)kPtsPerLine);
            const unsigned lb = (unsigned)(ll - mv.lines) * 8u + 1u;
            for (int j = 0; j < n; j++) offer(t, ldg_f4(&ll->pts[j]), lb + (unsigned)j);
          }
        }
      }
    };
#pragma unroll 1
    for (int e = lane; e < n_ent; e += 64) {
      Ent t0, t1;
      const bool two = e + 32 < n_ent;
      fetch(e, t0);
      if (two) fetch(e + 32, t1);
      consume(t0);
      if (two) consume(t1);
    }
    __syncwarp();
    // ---------------- phase C: fold the own query's candidates into the register top-K
    const unsigned nc = sm.cand_n[lane];
#ifdef LSD_SIMT_EMU
    if (active && o_begin == 0) SIMT_STAT_ADD(0, 1);
    SIMT_STAT_ADD(3, nc); SIMT_STAT_MAX(7, nc);
    if (nc > (unsigned)kFlatCand) SIMT_STAT_ADD(4, 1);
#endif
    if (nc > (unsigned)kFlatCand) {
      // the list lost points: walk this pass's cells serially instead (exact; the list is ignored)
      for (int oo = o_begin; oo < o; oo++) {
        const int x = c.x + st.off[oo][0], y = c.y + st.off[oo][1], z = c.z + st.off[oo][2];
        if (coord_ok(x, y, z)) flat_walk_voxel<K, LOC>(mv, pack_key(x, y, z, 0), qx, qy, qz, max_sq, best, found);
      }
    } else {
      found += (int)nc;
      for (unsigned p = 0; p < nc; p++) {
        const int at = lane * kFlatCandStride + (int)p;
        best.push(__uint_as_float(sm.cand_d[at]), sm.cand_id[at], LOC ? sm.cand_loc[at] : 0u);
      }
    }
    sm.cand_n[lane] = 0u;
    __syncwarp();
  }
}

}  // namespace lsd
