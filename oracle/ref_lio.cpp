// TEST INFRASTRUCTURE ONLY — never linked into the product (liblsdreg.so).
//
// C-ABI wrapper around the UNMODIFIED reference sources, compiled where they lie under
// /root/reference by oracle/Makefile into oracle/_ref/libref_lio.so:
//   * faster_lio::IVox      slam/mapping/fastlio/include/ivox3d/ivox3d.h:31-112   (live LIO map)
//   * KD_TREE (ikd-Tree)    slam/mapping/fastlio/include/ikd-Tree/ikd_Tree.{h,cpp} (named oracle)
//   * esti_plane<float>     slam/mapping/fastlio/include/common_lib.h:236-268
// Only this wrapper and the PCL point-type shims (oracle/ref_shim) are ours.  Map point ids are
// carried through the reference code in the normal_x / normal_y fields (16 bits each, exactly
// representable as float) so neighbour *indices* can be compared, not just coordinates.
#include <pcl/point_types.h>
typedef pcl::PointXYZINormal PointType;  // ivox3d_node.hpp:59 expects a global PointType

#include <common_lib.h>  // reference: esti_plane, calc_dist, PointVector
#include <ivox3d/ivox3d.h>
#include <ikd-Tree/ikd_Tree.h>

#include <omp.h>
#include <cstdint>
#include <cstring>

using IVoxType = faster_lio::IVox<3, faster_lio::IVoxNodeType::DEFAULT, PointType>;

static inline PointType mk(const float* p, int id) {
  PointType q;
  q.x = p[0]; q.y = p[1]; q.z = p[2];
  q.normal_x = float(id & 0xFFFF);
  q.normal_y = float((id >> 16) & 0xFFFF);
  return q;
}
static inline int id_of(const PointType& p) { return int(p.normal_x) | (int(p.normal_y) << 16); }

extern "C" {

// ---------------------------------------------------------------- iVox
void* ref_ivox_create(float res, int nearby, size_t capacity) {
  IVoxType::Options o;
  o.resolution_ = res;
  o.capacity_ = capacity;
  o.max_distance_ = 100.0;
  switch (nearby) {
    case 0: o.nearby_type_ = IVoxType::NearbyType::CENTER; break;
    case 6: o.nearby_type_ = IVoxType::NearbyType::NEARBY6; break;
    case 18: o.nearby_type_ = IVoxType::NearbyType::NEARBY18; break;
    case 26: o.nearby_type_ = IVoxType::NearbyType::NEARBY26; break;
    default: o.nearby_type_ = IVoxType::NearbyType::NEARBY74; break;
  }
  return new IVoxType(o);
}
void ref_ivox_destroy(void* h) { delete static_cast<IVoxType*>(h); }

void ref_ivox_add(void* h, const float* xyz, int n, int id0, double distance) {
  PointVector pv;
  pv.reserve(n);
  for (int i = 0; i < n; i++) pv.push_back(mk(xyz + 3 * i, id0 + i));
  static_cast<IVoxType*>(h)->AddPoints(pv, distance);
}
size_t ref_ivox_num_cells(void* h) { return static_cast<IVoxType*>(h)->NumValidGrids(); }

// k-NN exactly as laserMapping.cpp:849 calls it.  out_ids [nq,k] (-1 padded), out_xyz [nq,k,3],
// out_cnt [nq].  Order within a row is the reference's (nearest first, rest nth_element order).
void ref_ivox_knn(void* h, const float* q, int nq, int k, double max_sq, int* out_ids, float* out_xyz,
                  int* out_cnt, int nthreads) {
  IVoxType* iv = static_cast<IVoxType*>(h);
  omp_set_num_threads(nthreads > 0 ? nthreads : 1);
#pragma omp parallel for schedule(dynamic, 256)
  for (int i = 0; i < nq; i++) {
    PointType p = mk(q + 3 * i, 0);
    PointVector near;
    iv->GetClosestPoint(p, near, k, max_sq);
    int c = int(near.size());
    out_cnt[i] = c;
    for (int j = 0; j < k; j++) {
      out_ids[i * k + j] = j < c ? id_of(near[j]) : -1;
      for (int d = 0; d < 3; d++) out_xyz[(i * k + j) * 3 + d] = j < c ? (&near[j].x)[d] : 0.f;
    }
  }
}

// ---------------------------------------------------------------- ikd-Tree
void* ref_ikd_create() { return new KD_TREE<PointType>(); }  // heap: 1M-entry op log inside
void ref_ikd_destroy(void* h) { delete static_cast<KD_TREE<PointType>*>(h); }
void ref_ikd_build(void* h, const float* xyz, int n, int id0) {
  PointVector pv;
  pv.reserve(n);
  for (int i = 0; i < n; i++) pv.push_back(mk(xyz + 3 * i, id0 + i));
  static_cast<KD_TREE<PointType>*>(h)->Build(pv);
}
// Nearest_Search as laserMapping.cpp:846; out_d2 [nq,k] are the tree's own float distances.
void ref_ikd_knn(void* h, const float* q, int nq, int k, int* out_ids, float* out_d2, int* out_cnt,
                 int nthreads) {
  KD_TREE<PointType>* t = static_cast<KD_TREE<PointType>*>(h);
  omp_set_num_threads(nthreads > 0 ? nthreads : 1);
#pragma omp parallel for schedule(dynamic, 256)
  for (int i = 0; i < nq; i++) {
    PointType p = mk(q + 3 * i, 0);
    PointVector near;
    std::vector<float> d2(k);
    t->Nearest_Search(p, k, near, d2);
    int c = int(near.size());
    out_cnt[i] = c;
    for (int j = 0; j < k; j++) {
      out_ids[i * k + j] = j < c ? id_of(near[j]) : -1;
      out_d2[i * k + j] = j < c ? d2[j] : -1.f;
    }
  }
}

// ---------------------------------------------------------------- esti_plane
// pts5: [n,5,3]; pabcd: [n,4]; ok: [n]
void ref_esti_plane(const float* pts5, int n, float thr, float* pabcd, int* ok) {
  for (int i = 0; i < n; i++) {
    PointVector pv(5);
    for (int j = 0; j < 5; j++) {
      pv[j].x = pts5[(i * 5 + j) * 3 + 0];
      pv[j].y = pts5[(i * 5 + j) * 3 + 1];
      pv[j].z = pts5[(i * 5 + j) * 3 + 2];
    }
    Eigen::Matrix<float, 4, 1> r;
    ok[i] = esti_plane(r, pv, thr) ? 1 : 0;
    for (int d = 0; d < 4; d++) pabcd[i * 4 + d] = r(d);
  }
}

}  // extern "C"
