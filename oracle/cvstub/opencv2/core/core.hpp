// oracle/cvstub -- TEST INFRASTRUCTURE ONLY.  A stand-in for the small OpenCV subset the reference's front-end sources use,
// so that those sources can be compiled UNMODIFIED, where they lie under /root/reference, into oracle/_ref (see
// oracle/Makefile.ref).  The container has no OpenCV C++.  Container types (Mat, KeyPoint, Point_,
// Size, Rect, InputArray/OutputArray) are re-declared here; the image primitives (resize, copyMakeBorder, FAST, GaussianBlur,
// fastAtan2) forward to the oracle's implementations, which are pinned bit-exactly against cv2 (tests/golden).  What the
// resulting library adds is the reference's OWN control logic (grid loop, quad-tree, orientation / descriptor loops, DBoW2
// tree walk) run as written, against which the oracle's line-by-line restatement is checked.
#ifndef ORACLE_CVSTUB_CORE_HPP_
#define ORACLE_CVSTUB_CORE_HPP_

#include <algorithm>
#include <array>
#include <cfloat>
#include <climits>
#include <limits>
#include <numeric>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <deque>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_8UC1 0
#define CV_32F 5
#define CV_PI 3.1415926535897932384626433832795

inline int cvRound(double v) { return (int)lrint(v); }
inline int cvRound(float v) { return (int)lrintf(v); }
inline int cvRound(int v) { return v; }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvFloor(float v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
inline int cvCeil(float v) { int i = (int)v; return i + (i < v); }

namespace cv {

template <class T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T x_, T y_) : x(x_), y(y_) {}
  template <class U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
  Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
  Point_& operator*=(double s) { x = (T)(x * s); y = (T)(y * s); return *this; }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

template <class T>
struct Size_ {
  T width, height;
  Size_() : width(0), height(0) {}
  Size_(T w, T h) : width(w), height(h) {}
};
typedef Size_<int> Size;

struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int x_, int y_, int w, int h) : x(x_), y(y_), width(w), height(h) {}
};

struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float size_, float angle_ = -1, float response_ = 0, int octave_ = 0, int class_id_ = -1)
      : pt(x, y), size(size_), angle(angle_), response(response_), octave(octave_), class_id(class_id_) {}
};

template <class T>
struct Point3_ {
  T x, y, z;
  Point3_() : x(0), y(0), z(0) {}
  Point3_(T x_, T y_, T z_) : x(x_), y(y_), z(z_) {}
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;

enum { NORM_INF = 1, NORM_L1 = 2, NORM_L2 = 4 };

class Mat;
class _OutputArray;

struct MatStep {
  size_t v;
  MatStep(size_t s = 0) : v(s) {}
  operator size_t() const { return v; }
};

// cv::Mat::zeros returns a MatExpr; ASSIGNING it to an existing Mat of the same size and type zero-fills that matrix in
// place (MatOp_Initializer::assign -> Mat::create keeps the buffer).  computeDescriptors relies on it: it assigns
// Mat::zeros to a row range of the caller's descriptor matrix (orb_extractor.cpp, "descriptors = Mat::zeros(...)").
struct MatZerosExpr { int rows, cols, type; };

// 2-D matrix with byte step and ROI views (shared, reference-counted buffer); element size from the type code
class Mat {
 public:
  int rows, cols;
  MatStep step;
  uchar* data;
  Mat() : rows(0), cols(0), step(0), data(nullptr), type_(CV_8U) {}
  Mat(int r, int c, int type) : rows(0), cols(0), step(0), data(nullptr), type_(type) { create(r, c, type); }
  Mat(Size sz, int type) : rows(0), cols(0), step(0), data(nullptr), type_(type) { create(sz.height, sz.width, type); }
  Mat(int r, int c, int type, void* ext, size_t stp = 0)
      : rows(r), cols(c), step(stp ? stp : (size_t)c * esz(type)), data((uchar*)ext), type_(type) {}
  static size_t esz(int type) { return type == CV_32F ? 4 : 1; }
  size_t elemSize() const { return esz(type_); }
  void create(int r, int c, int type) {
    if (r == rows && c == cols && type == type_ && data) return;  // cv::Mat::create keeps a matching matrix (also an ROI view)
    rows = r; cols = c; type_ = type; step = (size_t)c * esz(type);
    // zero guard bands of 19 rows + 64 bytes on both sides: computeOrbDescriptor samples up to 18 px around keypoints that
    // may sit 16 px from the edge of the un-padded blurred clone (orb_extractor.cpp:59-61, 1029) -- an out-of-allocation
    // read in the reference; here (and in the oracle) it reads 0
    const size_t guard = 19 * (size_t)step + 64, body = (size_t)r * c * esz(type);
    owner_.reset(new uchar[body + 2 * guard], std::default_delete<uchar[]>());
    std::memset(owner_.get(), 0, guard);
    std::memset(owner_.get() + guard + body, 0, guard);
    data = owner_.get() + guard;
  }
  void release() { rows = cols = 0; step = 0; data = nullptr; owner_.reset(); }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  int type() const { return type_; }
  size_t step1() const { return (size_t)step / esz(type_); }
  bool isContinuous() const { return (size_t)step == (size_t)cols * esz(type_); }
  Size size() const { return Size(cols, rows); }
  template <class T> T* ptr(int r = 0) { return reinterpret_cast<T*>(data + (size_t)r * step); }
  template <class T> const T* ptr(int r = 0) const { return reinterpret_cast<const T*>(data + (size_t)r * step); }
  uchar* ptr(int r = 0) { return data + (size_t)r * step; }
  const uchar* ptr(int r = 0) const { return data + (size_t)r * step; }
  template <class T> T& at(int r, int c) { return reinterpret_cast<T*>(data + (size_t)r * step)[c]; }
  template <class T> const T& at(int r, int c) const { return reinterpret_cast<const T*>(data + (size_t)r * step)[c]; }
  Mat rowRange(int a, int b) const { Mat m = *this; m.rows = b - a; m.data = data + (size_t)a * step; return m; }
  Mat colRange(int a, int b) const { Mat m = *this; m.cols = b - a; m.data = data + (size_t)a * esz(type_); return m; }
  Mat row(int r) const { return rowRange(r, r + 1); }
  Mat col(int c) const { return colRange(c, c + 1); }
  // single-index access of a vector (n x 1 or 1 x n)
  template <class T> T& at(int i) { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
  template <class T> const T& at(int i) const { return cols == 1 ? at<T>(i, 0) : at<T>(0, i); }
  // channels are not modelled: an N x 2 CV_32F matrix IS the N x 1 two-channel view undistortPoints takes (frame.cpp:629-631)
  Mat reshape(int) const { return *this; }
  Mat t() const;                                   // CV_32F
  double dot(const Mat& m) const;                  // CV_32F, dotProd_: double accumulator over (double)a*b
  void convertTo(Mat& dst, int type) const;        // CV_8U -> CV_32F
  void copyTo(const _OutputArray& dst) const;
  static Mat ones(int r, int c, int type);
  static Mat eye(int r, int c, int type);
  Mat operator()(const Rect& r) const { return rowRange(r.y, r.y + r.height).colRange(r.x, r.x + r.width); }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int r = 0; r < rows; ++r) std::memcpy(m.ptr(r), ptr(r), (size_t)cols * esz(type_));
    return m;
  }
  static MatZerosExpr zeros(int r, int c, int type) { MatZerosExpr e = {r, c, type}; return e; }
  Mat(const MatZerosExpr& e) : rows(0), cols(0), step(0), data(nullptr), type_(e.type) { *this = e; }
  Mat& operator=(const MatZerosExpr& e) {
    create(e.rows, e.cols, e.type);
    for (int r = 0; r < rows; ++r) std::memset(ptr(r), 0, (size_t)cols * esz(type_));
    return *this;
  }

 private:
  int type_;
  std::shared_ptr<uchar> owner_;
};

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
  _OutputArray(const Mat& roi) : tmp_(roi), m_(&tmp_) {}  // a temporary ROI header: writes land in the parent matrix
  void create(int r, int c, int t) const { m_->create(r, c, t); }
  void release() const { m_->release(); }
  Mat getMat() const { return *m_; }
  Mat& ref() const { return *m_; }
 private:
  mutable Mat tmp_;
  Mat* m_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

// ---- CV_32F algebra used by src/data/*.cpp and orb_matcher.cpp.  Evaluated eagerly (no MatExpr).  Arithmetic follows what cv2
// does for the shapes that matter here (tests/golden/cv2_frame_tail.npz): matrix product = float products and float sums in k
// order; norm / dot = double accumulators.  Transposed products are formed on the explicit transpose.
inline void Mat::copyTo(const _OutputArray& o) const {
  Mat& dst = o.ref();
  dst.create(rows, cols, type_);
  for (int r = 0; r < rows; ++r) std::memcpy(dst.ptr(r), ptr(r), (size_t)cols * esz(type_));
}
inline Mat Mat::t() const {
  Mat m(cols, rows, CV_32F);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) m.at<float>(c, r) = at<float>(r, c);
  return m;
}
inline double Mat::dot(const Mat& m) const {
  double s = 0;
  const int n = rows * cols;
  for (int i = 0; i < n; ++i) s += (double)at<float>(i / cols, i % cols) * (double)m.at<float>(i / m.cols, i % m.cols);
  return s;
}
inline void Mat::convertTo(Mat& dst, int type) const {
  Mat out(rows, cols, type);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      const float v = type_ == CV_32F ? at<float>(r, c) : (float)at<uchar>(r, c);
      if (type == CV_32F) out.at<float>(r, c) = v; else out.at<uchar>(r, c) = (uchar)v;
    }
  dst = out;
}
inline Mat Mat::ones(int r, int c, int type) {
  Mat m(r, c, type);
  for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) m.at<float>(i, j) = 1.0f;
  return m;
}
inline Mat Mat::eye(int r, int c, int type) {
  Mat m(r, c, type);
  for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) m.at<float>(i, j) = i == j ? 1.0f : 0.0f;
  return m;
}
template <class F> inline Mat mat_map2(const Mat& a, const Mat& b, F f) {
  Mat m(a.rows, a.cols, CV_32F);
  for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) m.at<float>(r, c) = f(a.at<float>(r, c), b.at<float>(r, c));
  return m;
}
template <class F> inline Mat mat_map1(const Mat& a, F f) {
  Mat m(a.rows, a.cols, CV_32F);
  for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) m.at<float>(r, c) = f(a.at<float>(r, c));
  return m;
}
inline Mat operator+(const Mat& a, const Mat& b) { return mat_map2(a, b, [](float x, float y) { return x + y; }); }
inline Mat operator-(const Mat& a, const Mat& b) { return mat_map2(a, b, [](float x, float y) { return x - y; }); }
inline Mat operator-(const Mat& a) { return mat_map1(a, [](float x) { return -x; }); }
inline Mat operator*(const Mat& a, double s) { return mat_map1(a, [s](float x) { return (float)(x * s); }); }
inline Mat operator*(double s, const Mat& a) { return a * s; }
inline Mat operator/(const Mat& a, double s) { return mat_map1(a, [s](float x) { return (float)(x * (1.0 / s)); }); }  // scale 1/s, as MatExpr does
inline Mat operator*(const Mat& a, const Mat& b) {  // cv::gemm, small-matrix float path
  Mat m(a.rows, b.cols, CV_32F);
  for (int i = 0; i < a.rows; ++i)
    for (int j = 0; j < b.cols; ++j) {
      float s = a.at<float>(i, 0) * b.at<float>(0, j);
      for (int k = 1; k < a.cols; ++k) s = s + a.at<float>(i, k) * b.at<float>(k, j);
      m.at<float>(i, j) = s;
    }
  return m;
}
inline double norm(const Mat& a, int type = NORM_L2) {
  double s = 0;
  for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) { const double v = a.at<float>(r, c); s += type == NORM_L1 ? std::fabs(v) : v * v; }
  return type == NORM_L1 ? s : std::sqrt(s);
}
inline double norm(const Mat& a, const Mat& b, int type = NORM_L2) {
  double s = 0;
  for (int r = 0; r < a.rows; ++r)
    for (int c = 0; c < a.cols; ++c) {
      const double v = (double)(a.at<float>(r, c) - b.at<float>(r, c));  // the difference is formed in float (normDiffL1_32f)
      s += type == NORM_L1 ? std::fabs(v) : v * v;
    }
  return type == NORM_L1 ? s : std::sqrt(s);
}

// cv::Mat_<float>(r, c) << a, b, c  (frame.cpp:606, keyframe.cpp:93,485)
template <class T> class Mat_;
template <class T>
class MatCommaInitializer_ {
 public:
  MatCommaInitializer_(Mat_<T>* m, T v);
  MatCommaInitializer_& operator,(T v);
  operator Mat_<T>() const;
  operator Mat() const;
 private:
  Mat_<T>* m_;
  int i_;
};
template <class T>
class Mat_ : public Mat {
 public:
  Mat_() {}
  Mat_(int r, int c) : Mat(r, c, CV_32F) {}
  Mat_(const Mat& m) : Mat(m) {}
  T& operator()(int r, int c) { return this->template at<T>(r, c); }
  const T& operator()(int r, int c) const { return this->template at<T>(r, c); }
  T& operator()(int i) { return this->template at<T>(i); }
};
template <class T> inline MatCommaInitializer_<T>::MatCommaInitializer_(Mat_<T>* m, T v) : m_(m), i_(0) { m_->template at<T>(0, 0) = v; i_ = 1; }
template <class T> inline MatCommaInitializer_<T>& MatCommaInitializer_<T>::operator,(T v) {
  m_->template at<T>(i_ / m_->cols, i_ % m_->cols) = v; ++i_; return *this;
}
template <class T> inline MatCommaInitializer_<T>::operator Mat_<T>() const { return *m_; }
template <class T> inline MatCommaInitializer_<T>::operator Mat() const { return *m_; }
template <class T, class V> inline MatCommaInitializer_<T> operator<<(const Mat_<T>& m, V v) {
  return MatCommaInitializer_<T>(const_cast<Mat_<T>*>(&m), (T)v);
}

void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& dist, const Mat& R, const Mat& P);

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_DEFAULT = 4,
       BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1 };

// ---- primitives: forwarded to the oracle's cv2-pinned implementations (oracle/orb_oracle.h) ----
float fastAtan2(float y, float x);
void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType);
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_DEFAULT);
void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);

// persistence API referenced by DBoW2's YAML save/load (TemplatedVocabulary.h:1452-1640); never called by the oracle's
// checks (the text loader :1335-1422 is), so these only have to compile
class FileNode {
 public:
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode operator[](const char*) const { return FileNode(); }
  FileNode operator[](int) const { return FileNode(); }
  size_t size() const { return 0; }
  operator int() const { return 0; }
  operator float() const { return 0.f; }
  operator double() const { return 0.0; }
  operator std::string() const { return std::string(); }
};
class FileStorage {
 public:
  enum { READ = 0, WRITE = 1 };
  FileStorage(const std::string&, int) {}
  bool isOpened() const { return false; }
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode operator[](const char*) const { return FileNode(); }
  void release() {}
};
template <class T> inline FileStorage& operator<<(FileStorage& fs, const T&) { return fs; }

struct KeyPointsFilter {  // only referenced by the dead ComputeKeyPointsOld path (orb_extractor.cpp:796-983)
  static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};

}  // namespace cv
#endif
