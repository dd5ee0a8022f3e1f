// orbfe_shim.hpp -- header-only C++ drop-in shims over the C ABI of include/orbfe.h.
//
// The reference's Tracking / LocalMapping code calls the front-end through three C++ interfaces; each
// shim below keeps the reference's own signature so those callers link unchanged (INTEGRATION.md shows
// the reference-side edit):
//
//   class ORBextractor                  src/orb_features/orb_extractor.h:25-93
//       ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
//       void Compute(cv::InputArray image, cv::InputArray mask,
//                    std::vector<cv::KeyPoint>& keypoints, cv::OutputArray descriptors)   (+ operator())
//       GetLevels / GetScaleFactor / GetScaleFactors / GetInverseScaleFactors /
//       GetScaleSigmaSquares / GetInverseScaleSigmaSquares / GetImagePyramid
//   orbfe::ComputeStereoMatches(...)    body of Frame::ComputeStereoMatches, src/data/frame.cpp:406-577
//   orbfe::SearchForInitialization / SearchByProjection(Frame&, vector<MapPoint*>&, th) /
//   orbfe::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono)
//                                       bodies of the OrbMatcher routines, orb_matcher.cpp:264-382, 13-111,
//                                       1312-1453 (templates over the reference's Frame / MapPoint types,
//                                       using only their public accessors, src/data/frame.h:104-190)
//
// There is no CPU fallback: a failing C-ABI call throws std::runtime_error with orbfe_last_error().
#ifndef ORBFE_SHIM_HPP_
#define ORBFE_SHIM_HPP_

#include "orbfe.h"

#if defined(ORBFE_USE_OPENCV) || (defined(__has_include) && __has_include(<opencv2/core/core.hpp>))
#include <opencv2/core/core.hpp>
#else
#include "cv_compat.h"
#endif

#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

namespace orbfe {

inline void check(int rc, const char* what) {
  if (rc != ORBFE_OK) throw std::runtime_error(std::string(what) + ": " + orbfe_last_error());
}
static_assert(sizeof(cv::KeyPoint) == sizeof(orbfe_keypoint), "cv::KeyPoint must be the 28-byte POD");

}  // namespace orbfe

class ORBextractor {
 public:
  enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

  ORBextractor(int nfeatures, float scaleFactor, int nlevels, int iniThFAST, int minThFAST, int device = 0)
      : nlevels_(nlevels), scaleFactor_(scaleFactor), handle_(nullptr) {
    orbfe_params p;
    p.nfeatures = nfeatures; p.scale_factor = scaleFactor; p.nlevels = nlevels; p.ini_th_fast = iniThFAST;
    p.min_th_fast = minThFAST; p.max_width = 0; p.max_height = 0; p.max_images = 1;
    orbfe::check(orbfe_extractor_create(&p, device, &handle_), "orbfe_extractor_create");
    mvScaleFactor.resize(nlevels); mvInvScaleFactor.resize(nlevels); mvLevelSigma2.resize(nlevels);
    mvInvLevelSigma2.resize(nlevels); mnFeaturesPerLevel.resize(nlevels);
    orbfe::check(orbfe_extractor_tables(handle_, nullptr, mvScaleFactor.data(), mvInvScaleFactor.data(), mvLevelSigma2.data(),
                                        mvInvLevelSigma2.data(), mnFeaturesPerLevel.data()), "orbfe_extractor_tables");
    mvImagePyramid.resize(nlevels);
  }
  ~ORBextractor() { orbfe_extractor_destroy(handle_); }
  ORBextractor(const ORBextractor&) = delete;
  ORBextractor& operator=(const ORBextractor&) = delete;

  // orb_extractor.cpp:985-1049.  Mask is ignored, as in the reference (orb_extractor.h:40).
  void Compute(cv::InputArray _image, cv::InputArray /*_mask*/, std::vector<cv::KeyPoint>& _keypoints,
               cv::OutputArray _descriptors) {
    if (_image.empty()) return;
    cv::Mat image = _image.getMat();
    if (image.type() != CV_8UC1) throw std::runtime_error("ORBextractor::Compute: image must be CV_8UC1");
    const int cap = orbfe_extractor_max_keypoints(handle_);
    kps_.resize((size_t)cap);
    desc_.resize((size_t)cap * 32);
    int n = 0;
    int rc = orbfe_extract(handle_, image.data, image.cols, image.rows, (size_t)image.step, kps_.data(), desc_.data(), cap, &n);
    if (rc == ORBFE_ERR_CAPACITY) {  // aspect ratio changed the bound: retry with the reported count
      kps_.resize((size_t)n); desc_.resize((size_t)n * 32);
      rc = orbfe_extract(handle_, image.data, image.cols, image.rows, (size_t)image.step, kps_.data(), desc_.data(), n, &n);
    }
    orbfe::check(rc, "orbfe_extract");
    pyramid_stale_ = true;
    if (n == 0) _descriptors.release();
    else {
      _descriptors.create(n, 32, CV_8U);
      cv::Mat d = _descriptors.getMat();
      for (int i = 0; i < n; ++i) std::memcpy(d.ptr(i), desc_.data() + (size_t)i * 32, 32);
    }
    _keypoints.clear();
    _keypoints.resize((size_t)n);
    if (n) std::memcpy(static_cast<void*>(_keypoints.data()), kps_.data(), (size_t)n * sizeof(orbfe_keypoint));
  }
  void operator()(cv::InputArray image, cv::InputArray mask, std::vector<cv::KeyPoint>& keypoints,
                  cv::OutputArray descriptors) { Compute(image, mask, keypoints, descriptors); }

  int GetLevels() const { return nlevels_; }
  float GetScaleFactor() const { return scaleFactor_; }
  std::vector<float> GetScaleFactors() const { return mvScaleFactor; }
  std::vector<float> GetInverseScaleFactors() const { return mvInvScaleFactor; }
  std::vector<float> GetScaleSigmaSquares() const { return mvLevelSigma2; }
  std::vector<float> GetInverseScaleSigmaSquares() const { return mvInvLevelSigma2; }

  // The pyramid lives on the device (ComputeStereoMatches reads it there); host copies are made
  // lazily for callers that really want the pixels (orb_extractor.h:58).
  const std::vector<cv::Mat>& GetImagePyramid() {
    if (pyramid_stale_) {
      for (int l = 0; l < nlevels_; ++l) {
        int w = 0, h = 0;
        orbfe::check(orbfe_pyramid_level(handle_, 0, l, nullptr, 0, &w, &h), "orbfe_pyramid_level");
        mvImagePyramid[l].create(h, w, CV_8UC1);
        orbfe::check(orbfe_pyramid_level(handle_, 0, l, mvImagePyramid[l].data, (size_t)mvImagePyramid[l].step, &w, &h),
                     "orbfe_pyramid_level");
      }
      pyramid_stale_ = false;
    }
    return mvImagePyramid;
  }

  orbfe_extractor* handle() const { return handle_; }
  std::vector<cv::Mat> mvImagePyramid;

 protected:
  int nlevels_;
  float scaleFactor_;
  orbfe_extractor* handle_;
  bool pyramid_stale_ = true;
  std::vector<float> mvScaleFactor, mvInvScaleFactor, mvLevelSigma2, mvInvLevelSigma2;
  std::vector<int32_t> mnFeaturesPerLevel;
  std::vector<orbfe_keypoint> kps_;
  std::vector<uint8_t> desc_;
};

namespace orbfe {

inline const orbfe_keypoint* as_pod(const std::vector<cv::KeyPoint>& v) {
  return reinterpret_cast<const orbfe_keypoint*>(v.data());
}
// descriptors_ is N x 32 CV_8U; continuous in the reference (created by Compute), copied if it is not
inline const uint8_t* desc_rows(const cv::Mat& m, std::vector<uint8_t>& tmp) {
  if (m.rows == 0) return nullptr;
  if ((size_t)m.step == 32) return m.data;
  tmp.resize((size_t)m.rows * 32);
  for (int r = 0; r < m.rows; ++r) std::memcpy(tmp.data() + (size_t)r * 32, m.ptr(r), 32);
  return tmp.data();
}

// Body of Frame::ComputeStereoMatches (frame.cpp:406-577): fills stereo_coords_ / depths_.
// `baseline` replaces the uninitialised baseline_ read at frame.cpp:436 (pass baseline_fx_ / fx_).
inline void ComputeStereoMatches(ORBextractor& left, ORBextractor& right, const std::vector<cv::KeyPoint>& keypoints,
                                 const std::vector<cv::KeyPoint>& right_keypoints, const cv::Mat& descriptors,
                                 const cv::Mat& right_descriptors, float baseline_fx, float baseline,
                                 std::vector<float>& stereo_coords, std::vector<float>& depths) {
  const int n = (int)keypoints.size();
  stereo_coords.assign((size_t)n, -1.0f);
  depths.assign((size_t)n, -1.0f);
  if (n == 0) return;
  std::vector<uint8_t> tl, tr;
  check(orbfe_stereo_match(left.handle(), right.handle(), n, as_pod(keypoints), desc_rows(descriptors, tl),
                           (int)right_keypoints.size(), as_pod(right_keypoints), desc_rows(right_descriptors, tr), baseline_fx,
                           baseline, stereo_coords.data(), depths.data(), nullptr),
        "orbfe_stereo_match");
}

// RAII device view of a Frame for the matchers (64x48 grid built on the GPU, frame.cpp:234-248).
template <class FrameT>
class DeviceFrame {
 public:
  explicit DeviceFrame(const FrameT& F, int device = 0) : h_(nullptr) {
    std::vector<uint8_t> tmp;
    const std::vector<float>& ur = F.StereoCoordRight();
    check(orbfe_frame_create(device, (int)F.GetUndistortedKeys().size(), as_pod(F.GetUndistortedKeys()),
                             desc_rows(F.GetDescriptors(), tmp), ur.empty() ? nullptr : ur.data(), F.GetMinX(), F.GetMaxX(),
                             F.GetMinY(), F.GetMaxY(), (int)F.ScaleFactors().size(), F.ScaleFactors().data(), &h_),
          "orbfe_frame_create");
  }
  ~DeviceFrame() { orbfe_frame_destroy(h_); }
  DeviceFrame(const DeviceFrame&) = delete;
  DeviceFrame& operator=(const DeviceFrame&) = delete;
  orbfe_frame* get() const { return h_; }
 private:
  orbfe_frame* h_;
};

// Body of OrbMatcher::SearchForInitialization (orb_matcher.cpp:264-382).
template <class FrameT>
int SearchForInitialization(FrameT& F1, FrameT& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                            int windowSize, float mfNNratio, bool mbCheckOrientation) {
  DeviceFrame<FrameT> d1(F1), d2(F2);
  vnMatches12.assign(F1.GetUndistortedKeys().size(), -1);
  int n = 0;
  static_assert(sizeof(cv::Point2f) == 2 * sizeof(float), "cv::Point2f layout");
  check(orbfe_search_for_initialization(d1.get(), d2.get(), reinterpret_cast<float*>(vbPrevMatched.data()), vnMatches12.data(),
                                        windowSize, mfNNratio, mbCheckOrientation ? 1 : 0, &n),
        "orbfe_search_for_initialization");
  return n;
}

// Body of OrbMatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (orb_matcher.cpp:13-111).
template <class FrameT, class MapPointT>
int SearchByProjection(FrameT& F, const std::vector<MapPointT*>& vpMapPoints, int th, float mfNNratio) {
  const int nmp = (int)vpMapPoints.size(), nkp = (int)F.GetUndistortedKeys().size();
  std::vector<uint8_t> valid(nmp, 0), has_obs(nmp, 0), desc((size_t)nmp * 32, 0), occupied(nkp, 0);
  std::vector<float> px(nmp, 0.f), py(nmp, 0.f), pxr(nmp, 0.f), vc(nmp, 0.f);
  std::vector<int32_t> lvl(nmp, 0), assigned(nkp, -1);
  for (int i = 0; i < nmp; ++i) {
    MapPointT* p = vpMapPoints[i];
    if (!p->track_is_in_view || p->isBad()) continue;  // :23-27
    valid[i] = 1;
    px[i] = p->track_projected_x; py[i] = p->track_projected_y; pxr[i] = p->track_projected_x_right;
    lvl[i] = p->track_scale_level; vc[i] = p->track_view_cos;
    has_obs[i] = p->NumObservations() > 0;
    const cv::Mat d = p->GetDescriptor();
    std::memcpy(desc.data() + (size_t)i * 32, d.data, 32);
  }
  for (int k = 0; k < nkp; ++k) occupied[k] = F.GetMapPoint(k) && F.GetMapPoint(k)->NumObservations() > 0;  // :59-63
  DeviceFrame<FrameT> dF(F);
  int n = 0;
  check(orbfe_search_by_projection_mappoints(dF.get(), nmp, valid.data(), px.data(), py.data(), pxr.data(), lvl.data(), vc.data(),
                                             desc.data(), has_obs.data(), occupied.data(), th, mfNNratio, assigned.data(), &n),
        "orbfe_search_by_projection_mappoints");
  for (int k = 0; k < nkp; ++k)
    if (assigned[k] >= 0) F.SetMapPoint(k, vpMapPoints[assigned[k]]);  // :97
  return n;
}

// Body of OrbMatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono) (orb_matcher.cpp:1312-1453).
// `project(i, u, v, invzc)` performs the reference's own cv::Mat projection of LastFrame's map point i
// (:1346-1357: x3Dc = Rcw*x3Dw+tcw; invzc = 1/z; u = fx*xc*invzc+cx; v = fy*yc*invzc+cy) and returns false
// when the keypoint has no map point or is an outlier; forward/backward are the booleans of :1334-1335.
// That OpenCV arithmetic stays on the reference's side of the boundary so that it is bit-identical.
template <class FrameT, class ProjectFn>
int SearchByProjectionLastFrame(FrameT& CurrentFrame, const FrameT& LastFrame, float th, bool bForward, bool bBackward,
                                bool mbCheckOrientation, ProjectFn project) {
  const int nl = LastFrame.NumKeypoints(), nkp = (int)CurrentFrame.GetUndistortedKeys().size();
  std::vector<uint8_t> valid(nl, 0), has_obs(nl, 0), desc((size_t)nl * 32, 0), occupied(nkp, 0);
  std::vector<float> u(nl, 0.f), v(nl, 0.f), iz(nl, 0.f), ang(nl, 0.f);
  std::vector<int32_t> oct(nl, 0), assigned(nkp, -1);
  for (int i = 0; i < nl; ++i) {
    if (!project(i, u[i], v[i], iz[i])) continue;
    valid[i] = 1;
    oct[i] = LastFrame.GetKeys()[i].octave;
    ang[i] = LastFrame.GetUndistortedKeys()[i].angle;
    has_obs[i] = LastFrame.GetMapPoint(i)->NumObservations() > 0;
    const cv::Mat d = LastFrame.GetMapPoint(i)->GetDescriptor();
    std::memcpy(desc.data() + (size_t)i * 32, d.data, 32);
  }
  for (int k = 0; k < nkp; ++k) occupied[k] = CurrentFrame.GetMapPoint(k) && CurrentFrame.GetMapPoint(k)->NumObservations() > 0;
  DeviceFrame<FrameT> dC(CurrentFrame);
  int n = 0;
  check(orbfe_search_by_projection_lastframe(dC.get(), nl, valid.data(), u.data(), v.data(), iz.data(), oct.data(), ang.data(),
                                             desc.data(), has_obs.data(), CurrentFrame.GetBaselineFx(), bForward ? 1 : 0,
                                             bBackward ? 1 : 0, occupied.data(), th, mbCheckOrientation ? 1 : 0, assigned.data(), &n),
        "orbfe_search_by_projection_lastframe");
  for (int k = 0; k < nkp; ++k)
    if (assigned[k] >= 0) CurrentFrame.SetMapPoint(k, LastFrame.GetMapPoint(assigned[k]));
  return n;
}

// static OrbMatcher::DescriptorDistance (orb_matcher.cpp:1630-1646), batched on the device
inline std::vector<int32_t> DescriptorDistances(const cv::Mat& a, const cv::Mat& b, int device = 0) {
  std::vector<uint8_t> ta, tb;
  std::vector<int32_t> d((size_t)a.rows);
  check(orbfe_descriptor_distance(device, desc_rows(a, ta), desc_rows(b, tb), a.rows, d.data()), "orbfe_descriptor_distance");
  return d;
}

}  // namespace orbfe
#endif  // ORBFE_SHIM_HPP_
