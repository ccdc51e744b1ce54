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

}  // extern "C"
