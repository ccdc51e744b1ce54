// TEST INFRASTRUCTURE ONLY (oracle/_ref/libref_keyframe.so): OpenCV is absent and nothing compiled here decodes an image;
// the reference's headers (slam/common/mapping_types.h) only need cv::Mat to exist as a member type with the handful of
// members their inline constructors touch.  Written from scratch.
#pragma once
#include <cstddef>
#include <vector>
#define CV_8UC1 0
#define CV_32F 5
namespace cv {
enum { IMREAD_COLOR = 1 };
template <typename T> struct DataType { enum { type = CV_32F }; };
class Mat {
 public:
  Mat() {}
  Mat(int rows, int cols, int type) : buf_((size_t)rows * cols * (type == CV_32F ? 4 : 1)) { data = buf_.data(); }
  template <typename T> T& at(int i) { return reinterpret_cast<T*>(buf_.data())[i]; }
  unsigned char* data = nullptr;
 private:
  std::vector<unsigned char> buf_;
};
inline Mat imdecode(const Mat&, int) { return Mat(); }
}  // namespace cv
