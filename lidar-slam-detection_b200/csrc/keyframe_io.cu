// keyframe_io.cu — key-frame files, host only.
//
// Writes and reads what the reference's KeyFrame::save / loadOdom / loadPcd do (slam/common/keyframe.cpp:41-132), so that a
// map produced through this library opens in the reference's map editor and loader:
//   cloud.pcd   PCL 1.9 binary PCD of pcl::PointXYZI as pcl::io::savePCDFileBinary lays it out: the text header below and
//               the points packed x, y, z, intensity (padding members of the in-memory struct are not fields)
//   data        "stamp s ns / estimate / 4x4 / odom  / 4x4 / id n", matrices through Eigen's operator<<, whose default
//               IOFormat is restated in print_matrix4 (Eigen/src/Core/IO.h print_matrix: stream precision, one common
//               column width, " " between coefficients, "\n" between rows)
#include <cstdio>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "lsd_common.cuh"

namespace lsd {

static void print_matrix4(std::ostream& s, const double* m) {
  std::streamsize width = 0;
  for (int j = 0; j < 4; j++)
    for (int i = 0; i < 4; i++) {
      std::stringstream sstr;
      sstr.copyfmt(s);
      sstr << m[4 * i + j];
      width = std::max<std::streamsize>(width, (std::streamsize)sstr.str().length());
    }
  for (int i = 0; i < 4; i++) {
    if (width) s.width(width);
    s << m[4 * i];
    for (int j = 1; j < 4; j++) {
      s << " ";
      if (width) s.width(width);
      s << m[4 * i + j];
    }
    if (i < 3) s << "\n";
  }
}

// header of an empty cloud: the reference writes its own ASCII stub (slam/common/pcd_writer.cpp:9-25,47-59)
static const char* kEmptyHeader =
    "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
    "WIDTH 0\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS 0\nDATA ascii\n";

static std::string pcd_header_binary(int n) {
  std::ostringstream o;
  o.imbue(std::locale::classic());
  o << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
    << "WIDTH " << n << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA binary\n";
  return o.str();
}

}  // namespace lsd

using namespace lsd;

extern "C" {

lsd_status_t lsd_keyframe_save(const char* directory, uint64_t stamp_us, int64_t id, const float* xyzi, int n, const double* pose16) {
  if (!directory || !pose16 || n < 0 || (n > 0 && !xyzi)) { set_error("lsd_keyframe_save: bad arguments"); return LSD_ERR_INVALID; }
  const std::string dir(directory);
  {
    std::ofstream os(dir + "/cloud.pcd", std::ios::binary | std::ios::trunc);
    if (!os) { set_error("lsd_keyframe_save: cannot open %s/cloud.pcd", directory); return LSD_ERR_IO; }
    if (n == 0) {
      os << kEmptyHeader;
    } else {
      const std::string h = pcd_header_binary(n);
      os.write(h.data(), (std::streamsize)h.size());
      std::vector<float> buf((size_t)n * 4);
      for (int i = 0; i < n; i++) {
        buf[4 * (size_t)i] = xyzi[4 * (size_t)i]; buf[4 * (size_t)i + 1] = xyzi[4 * (size_t)i + 1]; buf[4 * (size_t)i + 2] = xyzi[4 * (size_t)i + 2];
        buf[4 * (size_t)i + 3] = xyzi[4 * (size_t)i + 3] * 255.0f;  // numpy_to_pointcloud(points, 255.0), graph_utils.cpp:124
      }
      os.write(reinterpret_cast<const char*>(buf.data()), (std::streamsize)(buf.size() * sizeof(float)));
    }
    if (!os) { set_error("lsd_keyframe_save: write to %s/cloud.pcd failed", directory); return LSD_ERR_IO; }
  }
  std::ofstream ofs(dir + "/data");
  if (!ofs) { set_error("lsd_keyframe_save: cannot open %s/data", directory); return LSD_ERR_IO; }
  const uint64_t sec = stamp_us / 1000000ULL, nsec = stamp_us % 1000000ULL * 1000;  // keyframe.cpp:127-128
  ofs << "stamp " << sec << " " << nsec << std::endl;
  ofs << "estimate" << std::endl; print_matrix4(ofs, pose16); ofs << std::endl;
  ofs << "odom " << std::endl; print_matrix4(ofs, pose16); ofs << std::endl;
  ofs << "id " << (long)id << std::endl;
  if (!ofs) { set_error("lsd_keyframe_save: write to %s/data failed", directory); return LSD_ERR_IO; }
  return LSD_OK;
}

lsd_status_t lsd_keyframe_load(const char* directory, uint64_t* stamp_us, int64_t* id, double* pose16, float* xyzi_out, int cap, int* n_out) {
  if (!directory || !n_out) { set_error("lsd_keyframe_load: bad arguments"); return LSD_ERR_INVALID; }
  const std::string dir(directory);
  {  // KeyFrame::loadOdom, keyframe.cpp:41-72: whitespace-separated tokens, "odom" is skipped
    std::ifstream ifs(dir + "/data");
    if (!ifs) { set_error("lsd_keyframe_load: cannot open %s/data", directory); return LSD_ERR_IO; }
    uint64_t ts = 0; long kid = -1; double T[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    while (!ifs.eof()) {
      std::string token;
      ifs >> token;
      if (token == "stamp") { uint64_t s = 0, ns = 0; ifs >> s; ifs >> ns; ts = s * 1000000ULL + ns / 1000ULL; }
      else if (token == "estimate") { for (int i = 0; i < 16; i++) ifs >> T[i]; }
      else if (token == "id") { ifs >> kid; }
    }
    if (stamp_us) *stamp_us = ts;
    if (id) *id = kid;
    if (pose16) memcpy(pose16, T, sizeof(T));
  }
  // cloud.pcd: the fields this library and the reference write (x y z intensity as 4-byte floats), binary or ascii
  std::ifstream is(dir + "/cloud.pcd", std::ios::binary);
  if (!is) { set_error("lsd_keyframe_load: cannot open %s/cloud.pcd", directory); return LSD_ERR_IO; }
  std::string line, fields, sizes, types, data;
  long points = -1;
  while (std::getline(is, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (line.rfind("FIELDS", 0) == 0) fields = line.substr(6);
    else if (line.rfind("SIZE", 0) == 0) sizes = line.substr(4);
    else if (line.rfind("TYPE", 0) == 0) types = line.substr(4);
    else if (line.rfind("POINTS", 0) == 0) points = atol(line.c_str() + 6);
    else if (line.rfind("DATA", 0) == 0) { data = line.substr(4); break; }
  }
  auto trim = [](std::string v) { size_t a = v.find_first_not_of(" \t"), b = v.find_last_not_of(" \t"); return a == std::string::npos ? std::string() : v.substr(a, b - a + 1); };
  fields = trim(fields); sizes = trim(sizes); types = trim(types); data = trim(data);
  if (points < 0 || fields != "x y z intensity" || sizes != "4 4 4 4" || types != "F F F F" || (data != "binary" && data != "ascii")) {
    set_error("lsd_keyframe_load: %s/cloud.pcd is not an x y z intensity float PCD (FIELDS '%s', DATA '%s')", directory, fields.c_str(), data.c_str());
    return LSD_ERR_IO;
  }
  *n_out = (int)points;
  if (!xyzi_out) return LSD_OK;
  if (points > cap) { set_error("lsd_keyframe_load: %ld points, room for %d", points, cap); return LSD_ERR_CAPACITY; }
  if (data == "binary") {
    is.read(reinterpret_cast<char*>(xyzi_out), (std::streamsize)((size_t)points * 16));
    if (is.gcount() != (std::streamsize)((size_t)points * 16)) { set_error("lsd_keyframe_load: %s/cloud.pcd is truncated", directory); return LSD_ERR_IO; }
  } else {
    for (long i = 0; i < points * 4; i++) if (!(is >> xyzi_out[i])) { set_error("lsd_keyframe_load: %s/cloud.pcd is truncated", directory); return LSD_ERR_IO; }
  }
  for (long i = 0; i < points; i++) xyzi_out[4 * i + 3] = xyzi_out[4 * i + 3] / 255.0f;  // KeyFrame::loadPcd, keyframe.cpp:101-103
  return LSD_OK;
}

}  // extern "C"
