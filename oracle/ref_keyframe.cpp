// TEST INFRASTRUCTURE ONLY — never linked into the product (liblsdreg.so).
//
// The reference's key-frame record, unmodified: slam/common/keyframe.cpp (KeyFrame::save / loadOdom / loadPcd /
// computeDescriptor), slam/common/pcd_writer.cpp and slam/common/Scancontext/Scancontext.cpp, compiled where they lie by
// oracle/Makefile into oracle/_ref/libref_keyframe.so.  Ours: this wrapper and the shims in oracle/ref_shim_keyframe
// (cv::Mat as a member type, pcl::io::savePCDFileBinary / PCDReader restating PCL's published binary PCD layout, two PCL
// helpers nothing here calls) plus the PCL containers of oracle/ref_shim_fastlio.  What this pins: the `data` text file
// (Eigen's operator<< formatting included) byte for byte, the x255 / /255 intensity convention, the stamp split, and — for
// row N4 of SURVEY.md section 8f, next round — the ScanContext descriptor and its ring / sector keys.
#include <cstring>

#include "keyframe.h"
#include "Scancontext/Scancontext.h"

extern "C" {

// dump_keyframe (slam/src/graph_utils.cpp:123-131): points with intensity x 255 -> KeyFrame(stamp, id, pose, points).save(dir)
void ref_keyframe_save(const char* dir, uint64_t stamp, long id, const float* xyzi, int n, const double* pose16) {
  PointCloud::Ptr cloud(new PointCloud());
  cloud->width = n; cloud->height = 1; cloud->points.resize(n);
  for (int i = 0; i < n; i++) {
    cloud->points[i].x = xyzi[4 * i]; cloud->points[i].y = xyzi[4 * i + 1]; cloud->points[i].z = xyzi[4 * i + 2];
    cloud->points[i].intensity = xyzi[4 * i + 3] * 255.0f;
  }
  Eigen::Matrix4d T;
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) T(r, c) = pose16[4 * r + c];
  KeyFrame kf(stamp, id, Eigen::Isometry3d(T), cloud);
  kf.save(dir);
}
// KeyFrame(id, dir, true): loadOdom + loadPcd.  Returns the point count, -1 if the odometry file or -2 if the cloud is missing.
int ref_keyframe_load(const char* dir, uint64_t* stamp, long* id, double* pose16, float* xyzi, int cap) {
  KeyFrame kf(-1, dir, true);
  if (!kf.loadOdom()) return -1;
  if (!kf.loadPcd()) return -2;
  *stamp = kf.mTimestamp; *id = kf.mId;
  for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) pose16[4 * r + c] = kf.mOdom(r, c);
  const int n = (int)kf.mPoints->points.size();
  for (int i = 0; i < n && i < cap; i++) {
    const Point& p = kf.mPoints->points[i];
    xyzi[4 * i] = p.x; xyzi[4 * i + 1] = p.y; xyzi[4 * i + 2] = p.z; xyzi[4 * i + 3] = p.intensity;
  }
  return n;
}
// SCManager::makeScancontext + ring / sector keys (Scancontext.cpp) of one cloud: sc [rows*cols] row-major.
int ref_scancontext(const float* xyzi, int n, double* sc, int cap, int* rows, int* cols, double* ringkey, double* sectorkey) {
  pcl::PointCloud<SCPointType> cloud;
  cloud.points.resize(n);
  for (int i = 0; i < n; i++) { cloud.points[i].x = xyzi[4 * i]; cloud.points[i].y = xyzi[4 * i + 1]; cloud.points[i].z = xyzi[4 * i + 2]; cloud.points[i].intensity = xyzi[4 * i + 3]; }
  SCManager m;
  Eigen::MatrixXd d = m.makeScancontext(cloud);
  *rows = (int)d.rows(); *cols = (int)d.cols();
  if ((long)d.size() > cap) return -1;
  for (int r = 0; r < d.rows(); r++) for (int c = 0; c < d.cols(); c++) sc[(size_t)r * d.cols() + c] = d(r, c);
  Eigen::MatrixXd rk = m.makeRingkeyFromScancontext(d), sk = m.makeSectorkeyFromScancontext(d);
  for (int i = 0; i < rk.size(); i++) ringkey[i] = rk(i);
  for (int i = 0; i < sk.size(); i++) sectorkey[i] = sk(i);
  return 0;
}


// ---- row N4: the rest of SCManager, so the restatement (oracle/scancontext.py) and the product are pinned on the whole
// retrieval, not only on the descriptor.  Descriptors cross this boundary in Eigen's own storage order (column-major
// 20 x 60, MatrixXd::data()).
static Eigen::MatrixXd sc_from(const double* d) { return Eigen::Map<const Eigen::MatrixXd>(d, 20, 60); }

// makeScancontext(cloud, dx, dy) + ring / sector keys
void ref_sc_make(const float* xyzi, int n, double dx, double dy, double* sc1200, double* ringkey20, double* sectorkey60) {
  pcl::PointCloud<SCPointType> cloud;
  cloud.points.resize(n);
  for (int i = 0; i < n; i++) { cloud.points[i].x = xyzi[4 * i]; cloud.points[i].y = xyzi[4 * i + 1]; cloud.points[i].z = xyzi[4 * i + 2]; cloud.points[i].intensity = xyzi[4 * i + 3]; }
  SCManager m;
  Eigen::MatrixXd d = m.makeScancontext(cloud, dx, dy);
  std::memcpy(sc1200, d.data(), sizeof(double) * 1200);
  Eigen::MatrixXd rk = m.makeRingkeyFromScancontext(d), sk = m.makeSectorkeyFromScancontext(d);
  for (int i = 0; i < 20; i++) ringkey20[i] = rk(i);
  for (int i = 0; i < 60; i++) sectorkey60[i] = sk(i);
}
// distanceBtnScanContext
void ref_sc_distance(const double* a1200, const double* b1200, double* dist, int* shift) {
  SCManager m;
  Eigen::MatrixXd a = sc_from(a1200), b = sc_from(b1200);
  std::pair<double, int> r = m.distanceBtnScanContext(a, b);
  *dist = r.first; *shift = r.second;
}
// buildRingKeyKDTree on a database, then detectClosestMatch / detectCandidateMatch for one query
void* ref_sc_db_create(const double* descs, int n, double dist_thres) {
  SCManager* m = new SCManager();
  m->SC_DIST_THRES = dist_thres;
  KeyMat keys;
  std::vector<Eigen::MatrixXd> pcs;
  for (int i = 0; i < n; i++) {
    Eigen::MatrixXd d = sc_from(descs + (size_t)i * 1200);
    keys.push_back(eig2stdvec(m->makeRingkeyFromScancontext(d)));
    pcs.push_back(d);
  }
  m->buildRingKeyKDTree(keys, pcs);
  return m;
}
void ref_sc_db_destroy(void* h) { delete static_cast<SCManager*>(h); }
int ref_sc_detect_closest(void* h, const double* q1200, float* yaw, double* score) {
  SCManager* m = static_cast<SCManager*>(h);
  Eigen::MatrixXd sc = sc_from(q1200);
  std::vector<float> rk = eig2stdvec(m->makeRingkeyFromScancontext(sc));
  Eigen::MatrixXd sk = m->makeSectorkeyFromScancontext(sc);
  *score = 1.0;   // what globalSearch passes in (global_localization.cpp:397)
  std::pair<int, float> r = m->detectClosestMatch(sc, rk, sk, *score);
  *yaw = r.second;
  return r.first;
}
int ref_sc_detect_candidates(void* h, const double* q1200, int* idx, float* yaw, float* dist, int cap) {
  SCManager* m = static_cast<SCManager*>(h);
  Eigen::MatrixXd sc = sc_from(q1200);
  std::vector<float> rk = eig2stdvec(m->makeRingkeyFromScancontext(sc));
  Eigen::MatrixXd sk = m->makeSectorkeyFromScancontext(sc);
  std::vector<SCMatch> r = m->detectCandidateMatch(sc, rk, sk);
  for (size_t i = 0; i < r.size() && (int)i < cap; i++) { idx[i] = r[i].sc_idx; yaw[i] = r[i].sc_yaw; dist[i] = r[i].sc_dist; }
  return (int)r.size();
}
// the ring-key tree alone: the NUM_CANDIDATES_FROM_TREE nearest database entries, in nanoflann's order
int ref_sc_ring_knn(void* h, const float* ringkey20, int* idx, float* d2, int cap) {
  SCManager* m = static_cast<SCManager*>(h);
  const int k = std::min(std::min(m->NUM_CANDIDATES_FROM_TREE, (int)m->polarcontexts_.size()), cap);
  if (k <= 0) return 0;
  std::vector<size_t> ci(k);
  std::vector<float> cd(k);
  nanoflann::KNNResultSet<float> rs(k);
  rs.init(&ci[0], &cd[0]);
  m->polarcontext_tree_->index->findNeighbors(rs, ringkey20, nanoflann::SearchParams(10));
  for (int i = 0; i < k; i++) { idx[i] = (int)ci[i]; d2[i] = cd[i]; }
  return k;
}

}  // extern "C"
