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
//   orbfe::SearchByBoW x2 / SearchForTriangulation / SearchByProjection(KeyFrame*, Scw, ...) /
//   orbfe::SearchByProjectionKeyFrame / Fuse x2 / SearchBySim3
//                                       the other OrbMatcher routines, orb_matcher.cpp:133-262, 499-632,
//                                       634-802, 384-497, 1455-1582, 804-954, 956-1079, 1081-1310
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
#include <type_traits>
#include <utility>
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

// The same view built device to device from the extractor that just processed the frame's image (rectified cameras only:
// UndistortKeyPoints is then the identity, frame.cpp:616-619).  with_stereo: take the stereo coordinates ComputeStereoMatches left
// on the device for this (left) extractor.
class DeviceFrameFromExtractor {
 public:
  DeviceFrameFromExtractor(ORBextractor& ex, float minX, float maxX, float minY, float maxY, bool with_stereo) : h_(nullptr) {
    check(orbfe_frame_from_extractor(ex.handle(), 0, with_stereo ? 1 : 0, minX, maxX, minY, maxY, &h_), "orbfe_frame_from_extractor");
  }
  ~DeviceFrameFromExtractor() { orbfe_frame_destroy(h_); }
  DeviceFrameFromExtractor(const DeviceFrameFromExtractor&) = delete;
  DeviceFrameFromExtractor& operator=(const DeviceFrameFromExtractor&) = delete;
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

// ---- the remaining OrbMatcher routines -------------------------------------------------------------
// Pattern: the adapter gathers what the routine reads from the reference's objects, the caller supplies a `gate`
// functor that runs the reference's OWN cv::Mat geometry for map point i (projection, depth sign, IsInImage,
// distance invariance, viewing angle, PredictScale -- e.g. orb_matcher.cpp:414-451) and returns false where the
// reference `continue`s before GetFeaturesInArea; the search runs on the GPU; the adapter applies the side effects in
// the reference's order.  KeyFrame's image bounds are private statics (keyframe.h:181-184): pass them in.

template <class KeyFrameT>
class DeviceKeyFrame {
 public:
  DeviceKeyFrame(const KeyFrameT& kf, float minX, float maxX, float minY, float maxY, int device = 0) : h_(nullptr) {
    std::vector<uint8_t> tmp;
    check(orbfe_frame_create(device, (int)kf.undistorted_keypoints.size(), as_pod(kf.undistorted_keypoints),
                             desc_rows(kf.descriptors, tmp), kf.right_coords.empty() ? nullptr : kf.right_coords.data(), minX,
                             maxX, minY, maxY, (int)kf.scale_factors.size(), kf.scale_factors.data(), &h_),
          "orbfe_frame_create");
  }
  ~DeviceKeyFrame() { orbfe_frame_destroy(h_); }
  DeviceKeyFrame(const DeviceKeyFrame&) = delete;
  DeviceKeyFrame& operator=(const DeviceKeyFrame&) = delete;
  orbfe_frame* get() const { return h_; }
 private:
  orbfe_frame* h_;
};
struct ImageBounds { float minX, maxX, minY, maxY; };

// DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned>>) flattened for the C ABI
struct FlatFeatureVector {
  std::vector<uint32_t> ids, idx;
  std::vector<int32_t> start;
  template <class FeatureVectorT>
  explicit FlatFeatureVector(const FeatureVectorT& fv) {
    start.push_back(0);
    for (typename FeatureVectorT::const_iterator it = fv.begin(); it != fv.end(); ++it) {
      ids.push_back((uint32_t)it->first);
      for (size_t k = 0; k < it->second.size(); ++k) idx.push_back((uint32_t)it->second[k]);
      start.push_back((int32_t)idx.size());
    }
  }
  int n() const { return (int)ids.size(); }
};

// per map point: gate result, projection, predicted level, representative descriptor
struct ProjectedPoints {
  std::vector<uint8_t> valid, desc;
  std::vector<float> u, v, ur;
  std::vector<int32_t> level;
  explicit ProjectedPoints(size_t n) : valid(n, 0), desc(n * 32, 0), u(n, 0.f), v(n, 0.f), ur(n, 0.f), level(n, 0) {}
  template <class MapPointT>
  void take_descriptor(size_t i, MapPointT* p) {
    const cv::Mat d = p->GetDescriptor();
    std::memcpy(desc.data() + i * 32, d.data, 32);
  }
};

// Body of OrbMatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (orb_matcher.cpp:133-262)
template <class KeyFrameT, class FrameT, class MapPointT>
int SearchByBoW(KeyFrameT* pKF, FrameT& F, std::vector<MapPointT*>& vpMapPointMatches, float mfNNratio, bool mbCheckOrientation) {
  const std::vector<MapPointT*> vpMapPointsKF = pKF->GetMapPointMatches();
  const int nkf = (int)vpMapPointsKF.size(), nf = F.NumKeypoints();
  vpMapPointMatches.assign((size_t)nf, static_cast<MapPointT*>(nullptr));
  std::vector<uint8_t> valid(nkf, 0), tmp;
  std::vector<float> ang(nkf, 0.f);
  for (int i = 0; i < nkf; ++i) {
    valid[i] = vpMapPointsKF[i] && !vpMapPointsKF[i]->isBad();  // :162-168
    ang[i] = pKF->undistorted_keypoints[i].angle;
  }
  const FlatFeatureVector fk(pKF->feature_vec), ff(F.GetFeatureVector());
  DeviceFrame<FrameT> dF(F);
  std::vector<int32_t> m((size_t)nf, -1);
  int n = 0;
  check(orbfe_search_by_bow(dF.get(), nkf, desc_rows(pKF->descriptors, tmp), ang.data(), valid.data(), fk.n(), fk.ids.data(),
                            fk.start.data(), fk.idx.data(), ff.n(), ff.ids.data(), ff.start.data(), ff.idx.data(), mfNNratio,
                            mbCheckOrientation ? 1 : 0, m.data(), &n),
        "orbfe_search_by_bow");
  for (int k = 0; k < nf; ++k)
    if (m[k] >= 0) vpMapPointMatches[k] = vpMapPointsKF[m[k]];
  return n;
}

// Body of OrbMatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&) (orb_matcher.cpp:499-632)
template <class KeyFrameT, class MapPointT>
int SearchByBoW(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12, const ImageBounds& b, float mfNNratio,
                bool mbCheckOrientation) {
  const std::vector<MapPointT*> mp1 = pKF1->GetMapPointMatches(), mp2 = pKF2->GetMapPointMatches();
  const int n1 = (int)mp1.size(), n2 = (int)mp2.size();
  vpMatches12.assign((size_t)n1, static_cast<MapPointT*>(nullptr));
  std::vector<uint8_t> v1(n1, 0), v2(n2, 0), tmp;
  std::vector<float> ang(n1, 0.f);
  for (int i = 0; i < n1; ++i) { v1[i] = mp1[i] && !mp1[i]->isBad(); ang[i] = pKF1->undistorted_keypoints[i].angle; }
  for (int i = 0; i < n2; ++i) v2[i] = mp2[i] && !mp2[i]->isBad();
  const FlatFeatureVector f1(pKF1->feature_vec), f2(pKF2->feature_vec);
  DeviceKeyFrame<KeyFrameT> d2(*pKF2, b.minX, b.maxX, b.minY, b.maxY);
  std::vector<int32_t> m((size_t)n1, -1);
  int n = 0;
  check(orbfe_search_by_bow_keyframes(d2.get(), n1, desc_rows(pKF1->descriptors, tmp), ang.data(), v1.data(), v2.data(), f1.n(),
                                      f1.ids.data(), f1.start.data(), f1.idx.data(), f2.n(), f2.ids.data(), f2.start.data(),
                                      f2.idx.data(), mfNNratio, mbCheckOrientation ? 1 : 0, m.data(), &n),
        "orbfe_search_by_bow_keyframes");
  for (int i = 0; i < n1; ++i)
    if (m[i] >= 0) vpMatches12[i] = mp2[m[i]];
  return n;
}

// Body of OrbMatcher::SearchForTriangulation (orb_matcher.cpp:634-802).  F12 = the 9 floats of the fundamental matrix
// (row-major), (ex, ey) = the epipole computed at :643-649 with the reference's cv::Mat arithmetic.
template <class KeyFrameT>
int SearchForTriangulation(KeyFrameT* pKF1, KeyFrameT* pKF2, const float F12[9], float ex, float ey,
                           std::vector<std::pair<size_t, size_t> >& vMatchedPairs, bool bOnlyStereo, const ImageBounds& b,
                           bool mbCheckOrientation) {
  const int n1 = (int)pKF1->undistorted_keypoints.size(), n2 = (int)pKF2->undistorted_keypoints.size();
  std::vector<uint8_t> v1(n1, 0), s1(n1, 0), v2(n2, 0), tmp;
  for (int i = 0; i < n1; ++i) { v1[i] = !pKF1->GetMapPoint(i); s1[i] = pKF1->right_coords[i] >= 0; }  // :678-684
  for (int i = 0; i < n2; ++i) v2[i] = !pKF2->GetMapPoint(i);                                           // :701-705
  const FlatFeatureVector f1(pKF1->feature_vec), f2(pKF2->feature_vec);
  DeviceKeyFrame<KeyFrameT> d2(*pKF2, b.minX, b.maxX, b.minY, b.maxY);
  std::vector<int32_t> m((size_t)n1, -1);
  int n = 0;
  check(orbfe_search_for_triangulation(d2.get(), n1, as_pod(pKF1->undistorted_keypoints), desc_rows(pKF1->descriptors, tmp),
                                       v1.data(), s1.data(), v2.data(), f1.n(), f1.ids.data(), f1.start.data(), f1.idx.data(),
                                       f2.n(), f2.ids.data(), f2.start.data(), f2.idx.data(), F12, ex, ey, bOnlyStereo ? 1 : 0,
                                       mbCheckOrientation ? 1 : 0, m.data(), &n),
        "orbfe_search_for_triangulation");
  vMatchedPairs.clear();
  vMatchedPairs.reserve((size_t)(n > 0 ? n : 0));
  for (int i = 0; i < n1; ++i)
    if (m[i] >= 0) vMatchedPairs.push_back(std::make_pair((size_t)i, (size_t)m[i]));  // :794-799
  return n;
}

// Body of OrbMatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpMatched, th) (orb_matcher.cpp:384-497).
// gate(i, u, v, level): the reference's :408-451 for vpPoints[i] (incl. isBad / spAlreadyFound).
template <class KeyFrameT, class MapPointT, class GateFn>
int SearchByProjection(KeyFrameT* pKF, const std::vector<MapPointT*>& vpPoints, std::vector<MapPointT*>& vpMatched, int th,
                       const ImageBounds& b, GateFn gate) {
  const size_t n = vpPoints.size(), nkp = vpMatched.size();
  ProjectedPoints P(n);
  for (size_t i = 0; i < n; ++i)
    if (gate(i, P.u[i], P.v[i], P.level[i])) { P.valid[i] = 1; P.take_descriptor(i, vpPoints[i]); }
  std::vector<uint8_t> in(nkp, 0);
  for (size_t k = 0; k < nkp; ++k) in[k] = vpMatched[k] != nullptr;
  DeviceKeyFrame<KeyFrameT> d(*pKF, b.minX, b.maxX, b.minY, b.maxY);
  std::vector<int32_t> m(nkp, -1);
  int nm = 0;
  check(orbfe_search_by_projection_sim3(d.get(), (int)n, P.valid.data(), P.u.data(), P.v.data(), P.level.data(), P.desc.data(),
                                        in.data(), th, m.data(), &nm),
        "orbfe_search_by_projection_sim3");
  for (size_t k = 0; k < nkp; ++k)
    if (m[k] >= 0) vpMatched[k] = vpPoints[m[k]];  // :490
  return nm;
}

// Body of OrbMatcher::SearchByProjection(Frame& Cur, KeyFrame*, sAlreadyFound, th, ORBdist) (orb_matcher.cpp:1455-1582).
// gate(i, u, v, level): :1473-1508 for pKF's map point i except the image-bound test (done by the library).
template <class FrameT, class KeyFrameT, class GateFn>
int SearchByProjectionKeyFrame(FrameT& CurrentFrame, KeyFrameT* pKF, float th, int ORBdist, bool mbCheckOrientation, GateFn gate) {
  typedef typename std::remove_reference<decltype(*pKF->GetMapPointMatches()[0])>::type MapPointT;
  const std::vector<MapPointT*> vpMPs = pKF->GetMapPointMatches();
  const size_t n = vpMPs.size();
  const int nkp = CurrentFrame.NumKeypoints();
  ProjectedPoints P(n);
  std::vector<float> ang(n, 0.f);
  for (size_t i = 0; i < n; ++i)
    if (gate(i, P.u[i], P.v[i], P.level[i])) {
      P.valid[i] = 1; P.take_descriptor(i, vpMPs[i]); ang[i] = pKF->undistorted_keypoints[i].angle;
    }
  std::vector<uint8_t> occ((size_t)nkp, 0);
  for (int k = 0; k < nkp; ++k) occ[k] = CurrentFrame.GetMapPoint(k) != nullptr;  // :1526
  DeviceFrame<FrameT> d(CurrentFrame);
  std::vector<int32_t> a((size_t)nkp, -1);
  int nm = 0;
  check(orbfe_search_by_projection_keyframe(d.get(), (int)n, P.valid.data(), P.u.data(), P.v.data(), P.level.data(), ang.data(),
                                            P.desc.data(), occ.data(), th, ORBdist, mbCheckOrientation ? 1 : 0, a.data(), &nm),
        "orbfe_search_by_projection_keyframe");
  for (int k = 0; k < nkp; ++k)
    if (a[k] >= 0) CurrentFrame.SetMapPoint(k, vpMPs[a[k]]);  // :1543 (rotation-rejected ones come back as -1, :1574)
  return nm;
}

// Body of OrbMatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) (orb_matcher.cpp:804-954).
// gate(i, u, v, ur, level): :823-866 for vpMapPoints[i].  The graph edits of :933-949 are applied here in the
// reference's order; a point an earlier edit made bad or put into pKF is skipped exactly as :828 would.
template <class KeyFrameT, class MapPointT, class GateFn>
int Fuse(KeyFrameT* pKF, const std::vector<MapPointT*>& vpMapPoints, float th, const ImageBounds& b, GateFn gate) {
  const size_t n = vpMapPoints.size();
  ProjectedPoints P(n);
  for (size_t i = 0; i < n; ++i)
    if (vpMapPoints[i] && gate(i, P.u[i], P.v[i], P.ur[i], P.level[i])) { P.valid[i] = 1; P.take_descriptor(i, vpMapPoints[i]); }
  DeviceKeyFrame<KeyFrameT> d(*pKF, b.minX, b.maxX, b.minY, b.maxY);
  std::vector<int32_t> best(n, -1);
  check(orbfe_fuse(d.get(), (int)n, P.valid.data(), P.u.data(), P.v.data(), P.ur.data(), P.level.data(), P.desc.data(), th,
                   best.data(), nullptr),
        "orbfe_fuse");
  int nFused = 0;
  for (size_t i = 0; i < n; ++i) {
    if (best[i] < 0) continue;
    MapPointT* pMP = vpMapPoints[i];
    if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) continue;  // :828, re-evaluated after the earlier edits
    MapPointT* pMPinKF = pKF->GetMapPoint(best[i]);
    if (pMPinKF) {
      if (!pMPinKF->isBad()) {
        if (pMPinKF->NumObservations() > pMP->NumObservations()) pMP->Replace(pMPinKF);
        else pMPinKF->Replace(pMP);
      }
    } else {
      pMP->AddObservation(pKF, best[i]);
      pKF->AddMapPoint(pMP, best[i]);
    }
    nFused++;
  }
  return nFused;
}

// Body of OrbMatcher::Fuse(KeyFrame*, cv::Mat Scw, vpPoints, th, vpReplacePoint) (orb_matcher.cpp:956-1079).
// gate(i, u, v, level): :981-1025 for vpPoints[i] (incl. isBad / spAlreadyFound).
template <class KeyFrameT, class MapPointT, class GateFn>
int Fuse(KeyFrameT* pKF, const std::vector<MapPointT*>& vpPoints, float th, std::vector<MapPointT*>& vpReplacePoint,
         const ImageBounds& b, GateFn gate) {
  const size_t n = vpPoints.size();
  ProjectedPoints P(n);
  for (size_t i = 0; i < n; ++i)
    if (gate(i, P.u[i], P.v[i], P.level[i])) { P.valid[i] = 1; P.take_descriptor(i, vpPoints[i]); }
  DeviceKeyFrame<KeyFrameT> d(*pKF, b.minX, b.maxX, b.minY, b.maxY);
  std::vector<int32_t> best(n, -1);
  check(orbfe_fuse_sim3(d.get(), (int)n, P.valid.data(), P.u.data(), P.v.data(), P.level.data(), P.desc.data(), th, best.data(),
                        nullptr),
        "orbfe_fuse_sim3");
  int nFused = 0;
  for (size_t i = 0; i < n; ++i) {
    if (best[i] < 0) continue;
    MapPointT* pMP = vpPoints[i];
    MapPointT* pMPinKF = pKF->GetMapPoint(best[i]);
    if (pMPinKF) {
      if (!pMPinKF->isBad()) vpReplacePoint[i] = pMPinKF;
    } else {
      pMP->AddObservation(pKF, best[i]);
      pKF->AddMapPoint(pMP, best[i]);
    }
    nFused++;
  }
  return nFused;
}

// Body of OrbMatcher::SearchBySim3 (orb_matcher.cpp:1081-1310).  gate1(i1, u, v, level): :1134-1170 (KF1's point
// into KF2, incl. vbAlreadyMatched1); gate2(i2, ...): :1214-1250.
template <class KeyFrameT, class MapPointT, class Gate1, class Gate2>
int SearchBySim3(KeyFrameT* pKF1, KeyFrameT* pKF2, std::vector<MapPointT*>& vpMatches12, float th, const ImageBounds& b,
                 Gate1 gate1, Gate2 gate2) {
  const std::vector<MapPointT*> mp1 = pKF1->GetMapPointMatches(), mp2 = pKF2->GetMapPointMatches();
  const size_t N1 = mp1.size(), N2 = mp2.size();
  ProjectedPoints P1(N1), P2(N2);
  for (size_t i = 0; i < N1; ++i)
    if (mp1[i] && gate1(i, P1.u[i], P1.v[i], P1.level[i])) { P1.valid[i] = 1; P1.take_descriptor(i, mp1[i]); }
  for (size_t i = 0; i < N2; ++i)
    if (mp2[i] && gate2(i, P2.u[i], P2.v[i], P2.level[i])) { P2.valid[i] = 1; P2.take_descriptor(i, mp2[i]); }
  DeviceKeyFrame<KeyFrameT> d1(*pKF1, b.minX, b.maxX, b.minY, b.maxY), d2(*pKF2, b.minX, b.maxX, b.minY, b.maxY);
  std::vector<int32_t> m(N1, -1);
  int nFound = 0;
  check(orbfe_search_by_sim3(d1.get(), d2.get(), P1.valid.data(), P1.u.data(), P1.v.data(), P1.level.data(), P1.desc.data(),
                             P2.valid.data(), P2.u.data(), P2.v.data(), P2.level.data(), P2.desc.data(), th, m.data(), &nFound),
        "orbfe_search_by_sim3");
  for (size_t i = 0; i < N1; ++i)
    if (m[i] >= 0) vpMatches12[i] = mp2[m[i]];  // :1303
  return nFound;
}

// ---- the Frame tail -------------------------------------------------------------------------------
// Body of Frame::UndistortKeyPoints (frame.cpp:614-641): K = calib_mat_ (fx, fy, cx, cy), dist = dist_coeff_ as floats.
inline void UndistortKeyPoints(const std::vector<cv::KeyPoint>& keypoints, float fx, float fy, float cx, float cy,
                               const std::vector<float>& dist_coeff, std::vector<cv::KeyPoint>& undistorted_keypoints,
                               int device = 0) {
  undistorted_keypoints.resize(keypoints.size());
  check(orbfe_undistort_keypoints(device, (int)keypoints.size(), as_pod(keypoints), fx, fy, cx, cy, dist_coeff.data(),
                                  (int)dist_coeff.size(), reinterpret_cast<orbfe_keypoint*>(undistorted_keypoints.data())),
        "orbfe_undistort_keypoints");
}

// The visibility loop of Tracker::SearchLocalPoints (core/tracker.cpp:1196-1211): Frame::IsInFrustum (frame.cpp:277-337)
// for every candidate local map point in one launch.  fetch(i, P[3], Pn[3], minDist, maxDist, maxDistRaw) copies GetWorldPos(),
// GetNormal(), Get{Min,Max}DistanceInvariance() and max_dist_ (what PredictScale divides, map_point.cpp:386; needs a one-line
// getter on MapPoint) of vpMapPoints[i] and returns false for points the loop skips
// (last_frame_id_seen == frame id || isBad(), :1202-1205).  Rcw (row-major) / tcw / Ow = Frame::Rcw_, tcw_, Ow_.
// Writes the track_* fields exactly as IsInFrustum does and returns the number of visible points (nToMatch).
template <class MapPointT, class FetchFn>
int IsInFrustumBatch(const std::vector<MapPointT*>& vpMapPoints, const float Rcw[9], const float tcw[3], const float Ow[3],
                     float fx, float fy, float cx, float cy, float baseline_fx, const ImageBounds& b, float log_scale_factor,
                     int scale_levels, float viewingCosLimit, FetchFn fetch, int device = 0) {
  const size_t n = vpMapPoints.size();
  std::vector<float> P, Pn, mn, mx, raw;
  std::vector<size_t> idx;
  for (size_t i = 0; i < n; ++i) {
    float p[3], q[3], a = 0, c = 0, r = 0;
    if (!fetch(i, p, q, a, c, r)) continue;
    vpMapPoints[i]->track_is_in_view = false;  // frame.cpp:278
    idx.push_back(i);
    P.insert(P.end(), p, p + 3); Pn.insert(Pn.end(), q, q + 3); mn.push_back(a); mx.push_back(c); raw.push_back(r);
  }
  const int m = (int)idx.size();
  std::vector<uint8_t> in((size_t)m);
  std::vector<float> u((size_t)m), v((size_t)m), ur((size_t)m), vc((size_t)m);
  std::vector<int32_t> lvl((size_t)m);
  int cnt = 0;
  check(orbfe_is_in_frustum(device, m, P.data(), Pn.data(), mn.data(), mx.data(), raw.data(), Rcw, tcw, Ow, fx, fy, cx, cy, baseline_fx, b.minX,
                            b.maxX, b.minY, b.maxY, log_scale_factor, scale_levels, viewingCosLimit, in.data(), u.data(), v.data(),
                            ur.data(), lvl.data(), vc.data(), &cnt),
        "orbfe_is_in_frustum");
  for (int j = 0; j < m; ++j) {
    if (!in[j]) continue;
    MapPointT* pMP = vpMapPoints[idx[j]];  // frame.cpp:328-334
    pMP->track_is_in_view = true;
    pMP->track_projected_x = u[j];
    pMP->track_projected_x_right = ur[j];
    pMP->track_projected_y = v[j];
    pMP->track_scale_level = lvl[j];
    pMP->track_view_cos = vc[j];
  }
  return cnt;
}

// Body of Tracker::SearchLocalPoints' two loops (core/tracker.cpp:1196-1226) in one device-resident call: the visibility test
// of IsInFrustumBatch chained into SearchByProjection(Frame&, vpMapPoints, th).  fetch as in IsInFrustumBatch.  Writes
// track_is_in_view / track_scale_level and F.SetMapPoint exactly as the two reference routines do; returns nToMatch and the
// number of matches through n_matches.
template <class FrameT, class MapPointT, class FetchFn>
int SearchLocalPoints(FrameT& F, const std::vector<MapPointT*>& vpMapPoints, const float Rcw[9], const float tcw[3], const float Ow[3],
                      int th, float mfNNratio, float viewingCosLimit, FetchFn fetch, int* n_matches = nullptr) {
  const size_t n = vpMapPoints.size();
  std::vector<float> P, Pn, mn, mx, raw;
  std::vector<uint8_t> desc, has_obs;
  std::vector<size_t> idx;
  for (size_t i = 0; i < n; ++i) {
    float p[3], q[3], a = 0, c = 0, r = 0;
    if (!fetch(i, p, q, a, c, r)) continue;
    vpMapPoints[i]->track_is_in_view = false;
    idx.push_back(i);
    P.insert(P.end(), p, p + 3); Pn.insert(Pn.end(), q, q + 3); mn.push_back(a); mx.push_back(c); raw.push_back(r);
    const cv::Mat d = vpMapPoints[i]->GetDescriptor();
    desc.insert(desc.end(), d.data, d.data + 32);
    has_obs.push_back(vpMapPoints[i]->NumObservations() > 0);
  }
  const int m = (int)idx.size(), nkp = (int)F.GetUndistortedKeys().size();
  std::vector<uint8_t> occupied((size_t)nkp, 0), in((size_t)m);
  for (int k = 0; k < nkp; ++k) occupied[k] = F.GetMapPoint(k) && F.GetMapPoint(k)->NumObservations() > 0;  // orb_matcher.cpp:59-63
  std::vector<int32_t> lvl((size_t)m), assigned((size_t)nkp, -1);
  DeviceFrame<FrameT> dF(F);
  int nv = 0, nm = 0;
  check(orbfe_search_local_points(dF.get(), m, P.data(), Pn.data(), mn.data(), mx.data(), raw.data(), Rcw, tcw, Ow, F.GetFx(), F.GetFy(),
                                  F.GetCx(), F.GetCy(), F.GetBaselineFx(), F.GetLogScaleFactor(), viewingCosLimit, desc.data(),
                                  has_obs.data(), occupied.data(), th, mfNNratio, in.data(), lvl.data(), assigned.data(), &nv, &nm),
        "orbfe_search_local_points");
  for (int j = 0; j < m; ++j)
    if (in[j]) { vpMapPoints[idx[j]]->track_is_in_view = true; vpMapPoints[idx[j]]->track_scale_level = lvl[j]; }
  for (int k = 0; k < nkp; ++k)
    if (assigned[k] >= 0) F.SetMapPoint(k, vpMapPoints[idx[assigned[k]]]);
  if (n_matches) *n_matches = nm;
  return nv;
}

// ---- OrbVocabulary: the transform path of DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>
// (src/orb_features/orb_vocabulary.h; used by Frame::ComputeBoW frame.cpp:258-263 and KeyFrame::ComputeBoW
// keyframe.cpp:127-137).  BowVectorT / FeatureVectorT are DBoW2::BowVector / DBoW2::FeatureVector (std::map subclasses).
class OrbVocabulary {
 public:
  explicit OrbVocabulary(int device = 0) : device_(device), h_(nullptr) {}
  ~OrbVocabulary() { orbfe_vocabulary_destroy(h_); }
  OrbVocabulary(const OrbVocabulary&) = delete;
  OrbVocabulary& operator=(const OrbVocabulary&) = delete;
  bool loadFromTextFile(const std::string& filename) {  // TemplatedVocabulary.h:1335
    orbfe_vocabulary_destroy(h_);
    h_ = nullptr;
    return orbfe_vocabulary_load_text(filename.c_str(), device_, &h_) == ORBFE_OK;
  }
  bool empty() const {
    int words = 0;
    return !h_ || orbfe_vocabulary_info(h_, nullptr, nullptr, nullptr, nullptr, nullptr, &words) != ORBFE_OK || words == 0;
  }
  // transform(features, v, fv, levelsup) (TemplatedVocabulary.h:1124); features = Converter::toDescriptorVector(descriptors_)
  template <class BowVectorT, class FeatureVectorT>
  void transform(const std::vector<cv::Mat>& features, BowVectorT& v, FeatureVectorT& fv, int levelsup) const {
    v.clear();
    fv.clear();
    const int n = (int)features.size();
    if (empty() || n == 0) return;
    std::vector<uint8_t> d((size_t)n * 32);
    for (int i = 0; i < n; ++i) std::memcpy(d.data() + (size_t)i * 32, features[i].data, 32);
    std::vector<uint32_t> bw((size_t)n), fn((size_t)n), fi((size_t)n);
    std::vector<double> bv((size_t)n);
    std::vector<int32_t> fs((size_t)n + 1);
    int nb = 0, nf = 0;
    check(orbfe_bow_transform(h_, n, d.data(), levelsup, nullptr, nullptr, bw.data(), bv.data(), &nb, fn.data(), fs.data(), fi.data(), &nf),
          "orbfe_bow_transform");
    for (int u = 0; u < nb; ++u) v.insert(v.end(), typename BowVectorT::value_type(bw[u], bv[u]));
    for (int f = 0; f < nf; ++f)
      fv.insert(fv.end(), typename FeatureVectorT::value_type(fn[f], std::vector<unsigned int>(fi.begin() + fs[f], fi.begin() + fs[f + 1])));
  }
  orbfe_vocabulary* handle() const { return h_; }
 private:
  int device_;
  orbfe_vocabulary* h_;
};

}  // namespace orbfe
#endif  // ORBFE_SHIM_HPP_
