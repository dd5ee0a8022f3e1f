// cv_compat.h -- minimal stand-ins for the OpenCV types that cross the reference's front-end
// interface (cv::Mat for 8-bit images / descriptor matrices, cv::KeyPoint, cv::Point2f,
// cv::InputArray / cv::OutputArray), used ONLY when the real OpenCV headers are absent (this build
// container has no OpenCV C++), so that include/orbfe_shim.hpp can be compile- and run-tested.
// With OpenCV installed, orbfe_shim.hpp includes <opencv2/core/core.hpp> instead and this file is
// not used.  Layouts match OpenCV's (cv::KeyPoint is the 28-byte POD the C ABI exchanges).
#ifndef ORBFE_CV_COMPAT_H_
#define ORBFE_CV_COMPAT_H_

#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

#define CV_8U 0
#define CV_8UC1 0

namespace cv {

struct Point2f {
  float x, y;
  Point2f() : x(0), y(0) {}
  Point2f(float x_, float y_) : x(x_), y(y_) {}
};

struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
};
static_assert(sizeof(KeyPoint) == 28, "cv::KeyPoint layout");

// reference-counted 2-D u8 matrix (continuous rows with a byte step), enough of cv::Mat for this path
class Mat {
 public:
  int rows, cols;
  size_t step;
  uint8_t* data;
  Mat() : rows(0), cols(0), step(0), data(nullptr) {}
  Mat(int r, int c, int /*type*/) : rows(0), cols(0), step(0), data(nullptr) { create(r, c, CV_8U); }
  Mat(int r, int c, int /*type*/, void* ext, size_t stp = 0) : rows(r), cols(c), step(stp ? stp : (size_t)c), data((uint8_t*)ext) {}
  void create(int r, int c, int /*type*/) {
    if (r == rows && c == cols && owner_ && step == (size_t)c) return;
    rows = r; cols = c; step = (size_t)c;
    owner_.reset(new uint8_t[(size_t)r * c + 1], std::default_delete<uint8_t[]>());
    data = owner_.get();
  }
  void release() { rows = cols = 0; step = 0; data = nullptr; owner_.reset(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return CV_8UC1; }
  uint8_t* ptr(int r = 0) { return data + (size_t)r * step; }
  const uint8_t* ptr(int r = 0) const { return data + (size_t)r * step; }
  Mat clone() const {
    Mat m(rows, cols, CV_8U);
    for (int r = 0; r < rows; ++r) std::memcpy(m.ptr(r), ptr(r), (size_t)cols);
    return m;
  }
  Mat row(int r) const { Mat m = *this; m.rows = 1; m.data = data + (size_t)r * step; return m; }

 private:
  std::shared_ptr<uint8_t> owner_;
};

// the proxies only need to hand a Mat through
class _InputArray {
 public:
  _InputArray() : m_(nullptr) {}
  _InputArray(const Mat& m) : m_(&m) {}
  bool empty() const { return !m_ || m_->empty(); }
  Mat getMat() const { return m_ ? *m_ : Mat(); }
 private:
  const Mat* m_;
};
class _OutputArray {
 public:
  _OutputArray(Mat& m) : m_(&m) {}
  void create(int r, int c, int t) const { m_->create(r, c, t); }
  void release() const { m_->release(); }
  Mat getMat() const { return *m_; }
 private:
  Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

}  // namespace cv
#endif
