// TEST INFRASTRUCTURE ONLY (oracle/_ref/libref_keyframe.so): stand-in for <pcl/io/pcd_io.h> (PCL 1.9.1 is external, SURVEY.md
// F4) with the two entry points slam/common/keyframe.cpp and pcd_writer.h use, restating PCL's published PCD v0.7 layout
// for pcl::PointXYZI: a text header (FIELDS x y z intensity / SIZE 4 4 4 4 / TYPE F F F F / COUNT 1 1 1 1 / WIDTH / HEIGHT /
// VIEWPOINT 0 0 0 1 0 0 0 / POINTS / DATA binary) followed by the points packed field by field (16 bytes each; the padding
// of the in-memory struct is not a field).  The same layout is read back by the reference's own vendored reader,
// third_party/pypcd.py (tests/test_keyframe_io.py).  Written from scratch.
#pragma once
#include <fstream>
#include <sstream>
#include <string>
#include <vector>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
namespace pcl {
namespace io {
inline int savePCDFileBinary(const std::string& file, const PointCloud<PointXYZI>& cloud) {
  std::ofstream os(file, std::ios::binary | std::ios::trunc);
  if (!os) return -1;
  const size_t n = cloud.points.size();
  std::ostringstream h;
  h.imbue(std::locale::classic());
  h << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\nWIDTH "
    << n << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA binary\n";
  os << h.str();
  for (const auto& p : cloud.points) { const float v[4] = {p.x, p.y, p.z, p.intensity}; os.write(reinterpret_cast<const char*>(v), 16); }
  return os ? 0 : -1;
}
}  // namespace io
class PCDReader {
 public:
  int read(const std::string& file, PointCloud<PointXYZI>& cloud) {
    std::ifstream is(file, std::ios::binary);
    if (!is) return -1;
    std::string line, data;
    long points = -1;
    while (std::getline(is, line)) {
      if (line.rfind("POINTS", 0) == 0) points = atol(line.c_str() + 6);
      else if (line.rfind("DATA", 0) == 0) { data = line.substr(5); break; }
    }
    if (points < 0) return -1;
    cloud.points.resize(points);
    cloud.width = (uint32_t)points; cloud.height = 1;
    for (long i = 0; i < points; i++) {
      float v[4];
      if (data == "binary") is.read(reinterpret_cast<char*>(v), 16);
      else is >> v[0] >> v[1] >> v[2] >> v[3];
      if (!is) return -1;
      cloud.points[i].x = v[0]; cloud.points[i].y = v[1]; cloud.points[i].z = v[2]; cloud.points[i].intensity = v[3];
    }
    return 0;
  }
};
}  // namespace pcl
