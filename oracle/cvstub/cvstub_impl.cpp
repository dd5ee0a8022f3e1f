// oracle/cvstub/cvstub_impl.cpp -- TEST INFRASTRUCTURE ONLY: the image primitives of the OpenCV stand-in, forwarded to the
// oracle's cv2-pinned implementations (oracle/orb_oracle.h, tests/golden/cv2_primitives.npz).
#include "opencv2/core/core.hpp"

#include "../orb_oracle.h"

#include <cstdio>
#include <cstdlib>

namespace cv {

float fastAtan2(float y, float x) { return orc_fast_atan2(y, x); }

void resize(const Mat& src, Mat& dst, Size dsize, double, double, int interpolation) {
  if (interpolation != INTER_LINEAR || src.type() != CV_8UC1) { fprintf(stderr, "cvstub: unsupported resize\n"); abort(); }
  dst.create(dsize.height, dsize.width, src.type());
  orc_resize_linear(src.data, src.cols, src.rows, (int)(size_t)src.step, dst.data, dst.cols, dst.rows, (int)(size_t)dst.step);
}

static inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}

// BORDER_REFLECT_101 with the source treated as isolated (the two call sites, orb_extractor.cpp:1066 and :1071, are: an ROI
// with BORDER_ISOLATED, and the caller's whole input image)
void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType) {
  if ((borderType & ~BORDER_ISOLATED) != BORDER_REFLECT_101 || src.type() != CV_8UC1) { fprintf(stderr, "cvstub: unsupported border\n"); abort(); }
  const Mat s = src.clone();  // dst may alias src (temp holds the ROI it is built from)
  dst.create(s.rows + top + bottom, s.cols + left + right, s.type());
  for (int y = 0; y < dst.rows; ++y) {
    const uchar* srow = s.ptr(reflect101(y - top, s.rows));
    uchar* drow = dst.ptr(y);
    for (int x = 0; x < dst.cols; ++x) drow[x] = srow[reflect101(x - left, s.cols)];
  }
}

void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY, int borderType) {
  if (ksize.width != 7 || ksize.height != 7 || sigmaX != 2 || sigmaY != 2 || borderType != BORDER_REFLECT_101 || src.type() != CV_8UC1) {
    fprintf(stderr, "cvstub: unsupported GaussianBlur\n");
    abort();
  }
  const Mat s = src.clone();  // in-place call (orb_extractor.cpp:1030)
  dst.create(s.rows, s.cols, s.type());
  orc_gaussian7x7(s.data, s.cols, s.rows, (int)(size_t)s.step, dst.data, (int)(size_t)dst.step);
}

void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
  static_assert(sizeof(KeyPoint) == sizeof(orc_keypoint), "cv::KeyPoint layout");
  const int cap = std::max(image.rows * image.cols, 1);
  keypoints.resize((size_t)cap);
  const int n = orc_fast9(image.data, image.cols, image.rows, (int)(size_t)image.step, threshold, nonmaxSuppression ? 1 : 0,
                          reinterpret_cast<orc_keypoint*>(keypoints.data()), cap);
  keypoints.resize((size_t)std::max(n, 0));
}

void KeyPointsFilter::retainBest(std::vector<KeyPoint>&, int) {
  fprintf(stderr, "cvstub: KeyPointsFilter::retainBest is only reachable from the dead ComputeKeyPointsOld path\n");
  abort();
}

}  // namespace cv

namespace cv {
// cv::undistortPoints(src, dst, K, dist, Mat(), K) on an N x 2 CV_32F matrix (frame.cpp:629-631): the oracle's cv2-pinned
// implementation (tests/golden/cv2_frame_tail.npz)
void undistortPoints(const Mat& src, Mat& dst, const Mat& K, const Mat& dist, const Mat& R, const Mat& P) {
  if (!R.empty() || P.empty() || src.type() != CV_32F || src.cols != 2) { fprintf(stderr, "cvstub: unsupported undistortPoints\n"); abort(); }
  std::vector<float> in((size_t)src.rows * 2), out((size_t)src.rows * 2), d;
  for (int i = 0; i < src.rows; ++i) { in[2 * i] = src.at<float>(i, 0); in[2 * i + 1] = src.at<float>(i, 1); }
  const int nd = dist.rows * dist.cols;
  for (int i = 0; i < nd; ++i) d.push_back(dist.at<float>(i));
  orc_undistort_points(src.rows, in.data(), K.at<float>(0, 0), K.at<float>(1, 1), K.at<float>(0, 2), K.at<float>(1, 2), d.data(), nd, out.data());
  dst.create(src.rows, 2, CV_32F);
  for (int i = 0; i < src.rows; ++i) { dst.at<float>(i, 0) = out[2 * i]; dst.at<float>(i, 1) = out[2 * i + 1]; }
}
}  // namespace cv
