// oracle/ref_wrap_frame.cpp -- TEST INFRASTRUCTURE ONLY: C entry points around the reference's OWN Frame / MapPoint / KeyFrame /
// OrbMatcher classes (src/data/*.cpp, src/orb_features/orb_matcher.cpp), compiled unmodified from /root/reference into
// oracle/_ref/libslam_ref.so by oracle/Makefile.ref.  Nothing here restates reference logic: the wrappers build the reference's
// objects from test inputs, call the reference's functions and copy the results out.
#include <opencv2/core/core.hpp>
#include <deque>
#include <list>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <vector>
// Frame::mbInitialComputations (frame.h:218, private static) gates the once-per-process MakeInitialComputations (image bounds,
// grid cell sizes, intrinsics); ref_set_test_distortion must re-arm it when it changes the calibration.  Access only: the
// reference sources themselves are compiled unmodified and the data layout does not depend on access specifiers.
#define private public
#include "data/frame.h"
#undef private
#include "data/keyframe.h"
#include "data/map.h"
#include "data/map_point.h"
#include "orb_features/orb_matcher.h"
#include "util/converter.h"
#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"

#include <atomic>
#include <chrono>
#include <cstring>
#include <map>
#include <thread>
#include <memory>
#include <new>
#include <set>

// util/converter.cpp:3-10 needs Eigen/g2o for its other members; this is the one the compiled files use
std::vector<cv::Mat> Converter::toDescriptorVector(const cv::Mat& Descriptors) {
  std::vector<cv::Mat> vDesc;
  vDesc.reserve(Descriptors.rows);
  for (int j = 0; j < Descriptors.rows; ++j) vDesc.push_back(Descriptors.row(j));
  return vDesc;
}

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> RefVoc;
namespace {
RefVoc* g_voc = nullptr;  // vocabulary handed to the Frames constructed next (ref_set_vocabulary); not owned
std::shared_ptr<OrbVocabulary> voc_ptr() { return std::shared_ptr<OrbVocabulary>(g_voc, [](OrbVocabulary*) {}); }
struct RefFrame {
  alignas(Frame) unsigned char storage[sizeof(Frame)];
  bool live = false;
  std::shared_ptr<ORBextractor> exL, exR;
  std::shared_ptr<Map> map;
  std::vector<std::unique_ptr<MapPoint>> owned;
  std::unique_ptr<KeyFrame> kf;  // observer used to give map points observations
  Frame* f() { return reinterpret_cast<Frame*>(storage); }
  ~RefFrame() { kf.reset(); owned.clear(); if (live) f()->~Frame(); }
};
cv::Mat make_K(float fx, float fy, float cx, float cy) {
  cv::Mat K = cv::Mat::eye(3, 3, CV_32F);
  K.at<float>(0, 0) = fx; K.at<float>(1, 1) = fy; K.at<float>(0, 2) = cx; K.at<float>(1, 2) = cy;
  return K;
}
// rotation / Sim3 scale given to every pose that is built from a translation (ref_set_test_transform; identity and 1 by
// default).  make_pose(nullptr) stays the identity: "seen from the origin".
float g_dist[4] = {0, 0, 0, 0};   // k1 k2 p1 p2 given to every Frame constructed below (ref_set_test_distortion)
float g_rot[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
float g_sim3_scale = 1.f;
cv::Mat make_pose(const float* t) {  // rotation g_rot (identity unless set), translation t
  cv::Mat T = cv::Mat::eye(4, 4, CV_32F);
  if (t)
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) T.at<float>(r, c) = g_rot[3 * r + c];
  for (int i = 0; i < 3; ++i) T.at<float>(i, 3) = t ? t[i] : 0.f;
  return T;
}
cv::Mat make_vec3(const float* p) {
  cv::Mat m(3, 1, CV_32F);
  for (int i = 0; i < 3; ++i) m.at<float>(i) = p[i];
  return m;
}
KeyFrame* observer(RefFrame* R) {
  if (!R->kf) R->kf.reset(new KeyFrame(*R->f(), R->map, std::shared_ptr<KeyframeDatabase>()));
  return R->kf.get();
}
}  // namespace

extern "C" {

// Frame's stereo constructor (frame.cpp:61-111): two ORBextractor::Compute on two std::threads, UndistortKeyPoints,
// ComputeStereoMatches, AssignFeaturesToGrid.  ComputeStereoMatches reads baseline_ before the constructor sets it
// (frame.cpp:436 vs :108): the Frame is constructed twice in the same storage, so that the second construction finds the value
// the first one left there (= bf / fx) -- which is what happens to Tracker's frame objects from the second frame on.
void* ref_frame_stereo(const unsigned char* left, const unsigned char* right, int w, int h, int nfeatures, float sf, int nl, int ini,
                       int mn, float fx, float fy, float cx, float cy, float bf, float th_depth) {
  RefFrame* R = new RefFrame();
  std::memset(R->storage, 0, sizeof(R->storage));
  R->exL.reset(new ORBextractor(nfeatures, sf, nl, ini, mn));
  R->exR.reset(new ORBextractor(nfeatures, sf, nl, ini, mn));
  R->map = std::make_shared<Map>();
  cv::Mat imL(h, w, CV_8UC1, const_cast<unsigned char*>(left)), imR(h, w, CV_8UC1, const_cast<unsigned char*>(right));
  cv::Mat K = make_K(fx, fy, cx, cy), D = cv::Mat::zeros(4, 1, CV_32F);
  for (int i = 0; i < 4; ++i) D.at<float>(i) = g_dist[i];
  for (int pass = 0; pass < 2; ++pass) {
    if (pass) R->f()->~Frame();
    new (R->storage) Frame(imL, imR, 0.0, R->exL, R->exR, voc_ptr(), K, D, bf, th_depth);
  }
  R->live = true;
  R->f()->SetPose(make_pose(nullptr));
  if (g_voc) R->f()->ComputeBoW();  // frame.cpp:258-263
  return R;
}

// Frame's monocular constructor (frame.cpp:163-205)
void* ref_frame_mono(const unsigned char* img, int w, int h, int nfeatures, float sf, int nl, int ini, int mn, float fx, float fy,
                     float cx, float cy, float bf, float th_depth) {
  RefFrame* R = new RefFrame();
  std::memset(R->storage, 0, sizeof(R->storage));
  R->exL.reset(new ORBextractor(nfeatures, sf, nl, ini, mn));
  R->map = std::make_shared<Map>();
  cv::Mat im(h, w, CV_8UC1, const_cast<unsigned char*>(img));
  cv::Mat K = make_K(fx, fy, cx, cy), D = cv::Mat::zeros(4, 1, CV_32F);
  for (int i = 0; i < 4; ++i) D.at<float>(i) = g_dist[i];
  new (R->storage) Frame(im, 0.0, R->exL, voc_ptr(), K, D, bf, th_depth);
  R->live = true;
  R->f()->SetPose(make_pose(nullptr));
  if (g_voc) R->f()->ComputeBoW();
  return R;
}
void ref_frame_destroy(void* p) { delete static_cast<RefFrame*>(p); }
// every pose / Sim3 the entry points below build from a translation gets this rotation (row-major 3x3) and Sim3 scale
void ref_set_test_transform(const float* R9, float sim3_scale) {
  for (int i = 0; i < 9; ++i) g_rot[i] = R9 ? R9[i] : (i % 4 == 0 ? 1.f : 0.f);
  g_sim3_scale = sim3_scale > 0 ? sim3_scale : 1.f;
}
// distortion coefficients (k1 k2 p1 p2; null = none) of every Frame built from now on: the k1 != 0 paths of
// Frame::UndistortKeyPoints (frame.cpp:614-641) and Frame::ComputeImageBounds (:644-673)
void ref_set_test_distortion(const float* d4) {
  for (int i = 0; i < 4; ++i) g_dist[i] = d4 ? d4[i] : 0.f;
  Frame::mbInitialComputations = true;   // the next Frame recomputes bounds / grid / intrinsics (frame.cpp:100-103)
}
void ref_set_vocabulary(void* voc) { g_voc = static_cast<RefVoc*>(voc); }
void ref_frame_set_translation(void* p, const float* t) { static_cast<RefFrame*>(p)->f()->SetPose(make_pose(t)); }
int ref_frame_n(void* p) { return static_cast<RefFrame*>(p)->f()->NumKeypoints(); }
int ref_frame_n_right(void* p) { return (int)static_cast<RefFrame*>(p)->f()->GetRightKeys().size(); }
int ref_frame_levels(void* p) { return static_cast<RefFrame*>(p)->f()->GetScaleLevel(); }

// keypoints_ / undistorted_keypoints_ / descriptors_ / stereo_coords_ / depths_ / bounds / scale factors
void ref_frame_get(void* p, void* kps, void* kps_un, unsigned char* desc, float* u_right, float* depth, float* bounds, float* scale,
                   float* misc /* baseline, log_scale_factor */) {
  Frame* F = static_cast<RefFrame*>(p)->f();
  const int n = F->NumKeypoints();
  if (kps && n) std::memcpy(kps, F->GetKeys().data(), (size_t)n * sizeof(cv::KeyPoint));
  if (kps_un && n) std::memcpy(kps_un, F->GetUndistortedKeys().data(), (size_t)n * sizeof(cv::KeyPoint));
  if (desc) for (int i = 0; i < n; ++i) std::memcpy(desc + (size_t)i * 32, F->GetDescriptors().ptr(i), 32);
  if (u_right && !F->StereoCoordRight().empty()) std::memcpy(u_right, F->StereoCoordRight().data(), (size_t)n * sizeof(float));
  if (depth && !F->StereoDepth().empty()) std::memcpy(depth, F->StereoDepth().data(), (size_t)n * sizeof(float));
  if (bounds) { bounds[0] = F->GetMinX(); bounds[1] = F->GetMaxX(); bounds[2] = F->GetMinY(); bounds[3] = F->GetMaxY(); }
  if (scale) for (int l = 0; l < F->GetScaleLevel(); ++l) scale[l] = F->ScaleFactors()[l];
  if (misc) { misc[0] = F->GetBaseline(); misc[1] = F->GetLogScaleFactor(); }
}
void ref_frame_get_right(void* p, void* kps, unsigned char* desc) {
  Frame* F = static_cast<RefFrame*>(p)->f();
  const int n = (int)F->GetRightKeys().size();
  if (n) std::memcpy(kps, F->GetRightKeys().data(), (size_t)n * sizeof(cv::KeyPoint));
  for (int i = 0; i < n; ++i) std::memcpy(desc + (size_t)i * 32, F->GetRightDescriptors().ptr(i), 32);
}

// Frame::GetFeaturesInArea (frame.cpp:348-403)
int ref_features_in_area(void* p, float x, float y, float r, int min_level, int max_level, int* out, int cap) {
  const std::vector<size_t> v = static_cast<RefFrame*>(p)->f()->GetFeaturesInArea(x, y, r, min_level, max_level);
  for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int)v[i];
  return (int)v.size();
}

// OrbMatcher::SearchForInitialization (orb_matcher.cpp:264-382)
int ref_search_for_initialization(void* p1, void* p2, float* prev_matched_xy, int* matches12, int window, float nnratio, int check_ori) {
  Frame *F1 = static_cast<RefFrame*>(p1)->f(), *F2 = static_cast<RefFrame*>(p2)->f();
  const int n1 = F1->NumKeypoints();
  std::vector<cv::Point2f> prev((size_t)n1);
  for (int i = 0; i < n1; ++i) prev[i] = cv::Point2f(prev_matched_xy[2 * i], prev_matched_xy[2 * i + 1]);
  std::vector<int> m12;
  OrbMatcher matcher(nnratio, check_ori != 0);
  const int n = matcher.SearchForInitialization(*F1, *F2, prev, m12, window);
  for (int i = 0; i < n1; ++i) { matches12[i] = m12[i]; prev_matched_xy[2 * i] = prev[i].x; prev_matched_xy[2 * i + 1] = prev[i].y; }
  return n;
}

// builds n map points whose descriptor_ is desc[i] (through the MapPoint(pos, map, frame, idx) constructor on a scratch copy of
// the frame whose descriptor row 0 is overwritten) at world position pos[i]; has_obs[i] adds one observation
static void make_points(RefFrame* R, int n, const float* pos, const unsigned char* desc, const unsigned char* has_obs,
                        std::vector<MapPoint*>& out, const int* idx = nullptr) {
  Frame scratch(*R->f());
  scratch.SetPose(make_pose(nullptr));  // the points are created as seen from the origin (normal, scale range: map_point.cpp:61-73)
  KeyFrame* kf = observer(R);
  out.assign((size_t)n, nullptr);
  for (int i = 0; i < n; ++i) {
    const int row = idx ? idx[i] : 0;   // the keypoint whose octave sets max_dist_; its descriptor row is overwritten first
    std::memcpy(scratch.GetDescriptors().data + (size_t)row * (size_t)scratch.GetDescriptors().step, desc + (size_t)i * 32, 32);
    const float far[3] = {0.f, 0.f, 10.f};
    MapPoint* mp = new MapPoint(make_vec3(pos ? pos + 3 * i : far), R->map, &scratch, row);
    if (has_obs && has_obs[i]) mp->AddObservation(kf, 0);
    R->owned.emplace_back(mp);
    out[(size_t)i] = mp;
  }
}
static void set_occupied(RefFrame* R, const unsigned char* occupied) {
  Frame* F = R->f();
  const int n = F->NumKeypoints();
  std::vector<MapPoint*> one;
  const unsigned char zero[32] = {0}, yes = 1;
  make_points(R, 1, nullptr, zero, &yes, one);
  for (int k = 0; k < n; ++k) F->SetMapPoint(k, occupied && occupied[k] ? one[0] : nullptr);
}

// OrbMatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (orb_matcher.cpp:13-111): the track_* fields are set
// directly (as Frame::IsInFrustum would)
int ref_search_by_projection_mappoints(void* p, int n_mp, const unsigned char* valid, const float* px, const float* py,
                                       const float* pxr, const int* level, const float* view_cos, const unsigned char* desc,
                                       const unsigned char* has_obs, const unsigned char* occupied, int th, float nnratio,
                                       int* assigned) {
  RefFrame* R = static_cast<RefFrame*>(p);
  Frame* F = R->f();
  std::vector<MapPoint*> pts;
  make_points(R, n_mp, nullptr, desc, has_obs, pts);
  std::map<MapPoint*, int> index;
  for (int i = 0; i < n_mp; ++i) {
    MapPoint* m = pts[i];
    m->track_is_in_view = valid[i] != 0;
    m->track_projected_x = px[i]; m->track_projected_y = py[i]; m->track_projected_x_right = pxr[i];
    m->track_scale_level = level[i]; m->track_view_cos = view_cos[i];
    index[m] = i;
  }
  set_occupied(R, occupied);
  OrbMatcher matcher(nnratio, true);
  const int n = matcher.SearchByProjection(*F, pts, th);
  for (int k = 0; k < F->NumKeypoints(); ++k) {
    std::map<MapPoint*, int>::const_iterator it = index.find(F->GetMapPoint(k));
    assigned[k] = it == index.end() ? -1 : it->second;
  }
  return n;
}

// OrbMatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono) (orb_matcher.cpp:1312-1453).  Cur has the identity
// pose, so the projection of :1346-1356 is (fx*X/Z+cx, fy*Y/Z+cy) of the given world positions; Last's translation last_t
// selects bForward / bBackward (:1330-1335).  valid[i] = Last has a (non-outlier) map point at keypoint i.
int ref_search_by_projection_lastframe(void* pcur, void* plast, const unsigned char* valid, const float* world_pos,
                                       const unsigned char* desc, const unsigned char* has_obs, const float* last_t,
                                       const unsigned char* occupied, float th, int mono, int check_ori, int* assigned) {
  RefFrame *RC = static_cast<RefFrame*>(pcur), *RL = static_cast<RefFrame*>(plast);
  Frame *C = RC->f(), *Lf = RL->f();
  const int nl = Lf->NumKeypoints();
  std::vector<MapPoint*> pts;
  make_points(RC, nl, world_pos, desc, has_obs, pts);
  std::map<MapPoint*, int> index;
  for (int i = 0; i < nl; ++i) {
    Lf->SetMapPoint(i, valid[i] ? pts[i] : nullptr);
    Lf->SetOutlier(i, false);
    index[pts[i]] = i;
  }
  Lf->SetPose(make_pose(last_t));
  C->SetPose(make_pose(nullptr));
  set_occupied(RC, occupied);
  OrbMatcher matcher(0.9f, check_ori != 0);
  const int n = matcher.SearchByProjection(*C, *Lf, th, mono != 0);
  for (int k = 0; k < C->NumKeypoints(); ++k) {
    std::map<MapPoint*, int>::const_iterator it = index.find(C->GetMapPoint(k));
    assigned[k] = it == index.end() ? -1 : it->second;
  }
  return n;
}


// ---- vocabulary-node searches on the reference's own KeyFrame objects -----------------------------------

// Frame::ComputeBoW (frame.cpp:258-263) with the vocabulary loaded by ref_voc_load_text; outputs flattened like
// ref_voc_transform
// what Frame::ComputeBoW (frame.cpp:258-263) left in bow_vec_ / feature_vec_, flattened like ref_voc_transform
void ref_frame_get_bow(void* p, unsigned* bow_words, double* bow_values, int* n_bow, unsigned* fv_nodes, int* fv_start,
                       unsigned* fv_idx, int* n_fv) {
  Frame* F = static_cast<RefFrame*>(p)->f();
  const DBoW2::BowVector& bv = F->GetBowVector();
  const DBoW2::FeatureVector& fv = F->GetFeatureVector();
  int u = 0;
  for (DBoW2::BowVector::const_iterator it = bv.begin(); it != bv.end(); ++it, ++u) { bow_words[u] = it->first; bow_values[u] = it->second; }
  *n_bow = u;
  int f = 0, q = 0;
  for (DBoW2::FeatureVector::const_iterator it = fv.begin(); it != fv.end(); ++it, ++f) {
    fv_nodes[f] = it->first;
    fv_start[f] = q;
    for (size_t k = 0; k < it->second.size(); ++k) fv_idx[q++] = it->second[k];
  }
  fv_start[f] = q;
  *n_fv = f;
}

struct RefKeyFrame {
  RefFrame* owner;
  std::unique_ptr<KeyFrame> kf;
  std::vector<MapPoint*> pts;  // the map point created for keypoint i (or null)
};

// KeyFrame(frame, map, kfdb) (keyframe.cpp:20-80) + KeyFrame::ComputeBoW (:127-137); valid[i] gives
// keypoint i a (good) map point, bad[i] makes it isBad()
void* ref_keyframe_create(void* pframe, const unsigned char* valid, const unsigned char* bad, const float* translation) {
  RefFrame* R = static_cast<RefFrame*>(pframe);
  RefKeyFrame* K = new RefKeyFrame();
  K->owner = R;
  K->kf.reset(new KeyFrame(*R->f(), R->map, std::shared_ptr<KeyframeDatabase>()));  // copies the frame's bow_vec_ / feature_vec_
  if (R->f()->GetVocabulary()) K->kf->ComputeBoW();                                  // keyframe.cpp:127-137 (no-op when filled)
  K->kf->SetPose(make_pose(translation));
  const int n = R->f()->NumKeypoints();
  K->pts.assign((size_t)n, nullptr);
  if (valid) {
    std::vector<unsigned char> d((size_t)n * 32);
    for (int i = 0; i < n; ++i) std::memcpy(d.data() + (size_t)i * 32, K->kf->descriptors.ptr(i), 32);
    std::vector<MapPoint*> pts;
    make_points(R, n, nullptr, d.data(), nullptr, pts);
    for (int i = 0; i < n; ++i)
      if (valid[i]) {
        K->pts[i] = pts[i];
        K->kf->AddMapPoint(pts[i], i);
        pts[i]->AddObservation(K->kf.get(), i);
        if (bad && bad[i]) pts[i]->SetBadFlag();
      }
  }
  return K;
}
void ref_keyframe_destroy(void* p) { delete static_cast<RefKeyFrame*>(p); }

// OrbMatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (orb_matcher.cpp:133-262).  The Frame's feature vector is
// private state filled by Frame::ComputeBoW, which needs the vocabulary passed to the constructor: ref_frame_*_voc below.
int ref_search_by_bow_kf_f(void* pkf, void* pframe, float nnratio, int check_ori, int* matched_kf_idx) {
  RefKeyFrame* K = static_cast<RefKeyFrame*>(pkf);
  Frame* F = static_cast<RefFrame*>(pframe)->f();
  std::map<MapPoint*, int> index;
  for (size_t i = 0; i < K->pts.size(); ++i) if (K->pts[i]) index[K->pts[i]] = (int)i;
  std::vector<MapPoint*> matches;
  OrbMatcher matcher(nnratio, check_ori != 0);
  const int n = matcher.SearchByBoW(K->kf.get(), *F, matches);
  for (int k = 0; k < F->NumKeypoints(); ++k) {
    std::map<MapPoint*, int>::const_iterator it = index.find(matches[k]);
    matched_kf_idx[k] = it == index.end() ? -1 : it->second;
  }
  return n;
}

// OrbMatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&) (orb_matcher.cpp:499-632)
int ref_search_by_bow_kf_kf(void* pkf1, void* pkf2, float nnratio, int check_ori, int* matches12) {
  RefKeyFrame *K1 = static_cast<RefKeyFrame*>(pkf1), *K2 = static_cast<RefKeyFrame*>(pkf2);
  std::map<MapPoint*, int> index2;
  for (size_t i = 0; i < K2->pts.size(); ++i) if (K2->pts[i]) index2[K2->pts[i]] = (int)i;
  std::vector<MapPoint*> m12;
  OrbMatcher matcher(nnratio, check_ori != 0);
  const int n = matcher.SearchByBoW(K1->kf.get(), K2->kf.get(), m12);
  for (size_t i = 0; i < m12.size(); ++i) {
    std::map<MapPoint*, int>::const_iterator it = index2.find(m12[i]);
    matches12[i] = it == index2.end() ? -1 : it->second;
  }
  return n;
}

// OrbMatcher::SearchForTriangulation (orb_matcher.cpp:634-802); the epipole is computed inside from the two poses
int ref_search_for_triangulation(void* pkf1, void* pkf2, const float* F12, int only_stereo, float nnratio, int check_ori, int* matches12) {
  RefKeyFrame *K1 = static_cast<RefKeyFrame*>(pkf1), *K2 = static_cast<RefKeyFrame*>(pkf2);
  cv::Mat F(3, 3, CV_32F);
  for (int i = 0; i < 9; ++i) F.at<float>(i / 3, i % 3) = F12[i];
  std::vector<std::pair<size_t, size_t> > pairs;
  OrbMatcher matcher(nnratio, check_ori != 0);
  const int n = matcher.SearchForTriangulation(K1->kf.get(), K2->kf.get(), F, pairs, only_stereo != 0);
  const int n1 = K1->owner->f()->NumKeypoints();
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  for (size_t k = 0; k < pairs.size(); ++k) matches12[pairs[k].first] = (int)pairs[k].second;
  return n;
}

// Frame::IsInFrustum (frame.cpp:277-337) for n map points created (MapPoint(pos, map, frame, idx), map_point.cpp:40-80) while the
// frame sits at the origin, then tested from the pose `translation`.  idx[i] = the keypoint whose octave sets the point's
// scale range.  Also returns what the point holds (normal, min / max distance) so that the oracle can be fed the same.
int ref_is_in_frustum(void* pframe, int n, const float* world_pos, const int* idx, const float* translation, float viewing_cos_limit,
                      float* normal_out, float* min_dist_out, float* max_dist_out, float* ow_out, unsigned char* in_view, float* proj_x,
                      float* proj_y, float* proj_xr, int* level, float* view_cos) {
  RefFrame* R = static_cast<RefFrame*>(pframe);
  Frame* F = R->f();
  F->SetPose(make_pose(nullptr));
  std::vector<MapPoint*> pts((size_t)n);
  for (int i = 0; i < n; ++i) {
    pts[i] = new MapPoint(make_vec3(world_pos + 3 * i), R->map, F, idx[i]);
    R->owned.emplace_back(pts[i]);
  }
  F->SetPose(make_pose(translation));
  const cv::Mat Ow = F->GetCameraCenter();
  for (int c = 0; c < 3; ++c) ow_out[c] = Ow.at<float>(c);
  int count = 0;
  for (int i = 0; i < n; ++i) {
    MapPoint* m = pts[i];
    const cv::Mat nrm = m->GetNormal();
    for (int c = 0; c < 3; ++c) normal_out[3 * i + c] = nrm.at<float>(c);
    min_dist_out[i] = m->GetMinDistanceInvariance();
    max_dist_out[i] = m->GetMaxDistanceInvariance();
    const bool ok = F->IsInFrustum(m, viewing_cos_limit);
    in_view[i] = ok ? 1 : 0;
    proj_x[i] = ok ? m->track_projected_x : 0.f;
    proj_y[i] = ok ? m->track_projected_y : 0.f;
    proj_xr[i] = ok ? m->track_projected_x_right : 0.f;
    level[i] = ok ? m->track_scale_level : 0;
    view_cos[i] = ok ? m->track_view_cos : 0.f;
    count += ok;
  }
  return count;
}


// ---- the projection searches of N1 on the reference's own KeyFrame / MapPoint objects ---------------------------------------------
// free map points (not attached to the KeyFrame) at world_pos[i], created from the origin with the octave of keypoint idx[i]
// and descriptor desc[i]; normal / distance bounds they hold are returned for the oracle's gate replication
static void free_points(RefFrame* R, int n, const float* world_pos, const int* idx, const unsigned char* desc, std::vector<MapPoint*>& pts,
                        float* normal_out, float* min_out, float* max_out) {
  make_points(R, n, world_pos, desc, nullptr, pts, idx);
  for (int i = 0; i < n; ++i) {
    const cv::Mat nrm = pts[i]->GetNormal();
    for (int c = 0; c < 3; ++c) normal_out[3 * i + c] = nrm.at<float>(c);
    min_out[i] = pts[i]->GetMinDistanceInvariance();
    max_out[i] = pts[i]->GetMaxDistanceInvariance();
  }
}
static cv::Mat make_sim3(const float* t) {  // [s R | t], s = 1 and R = I unless ref_set_test_transform changed them
  cv::Mat S = make_pose(t);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) S.at<float>(r, c) = g_sim3_scale * g_rot[3 * r + c];
  return S;
}

// OrbMatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpMatched, th) (orb_matcher.cpp:384-497).  matched_in[k]
// pre-fills vpMatched[k] with a foreign point; matched[k] (out) = index of the point this call wrote there, else -1.
int ref_search_by_projection_sim3(void* pkf, int n, const float* world_pos, const int* idx, const unsigned char* desc,
                                  const unsigned char* bad, const unsigned char* matched_in, const float* t, int th, float* normal_out,
                                  float* min_out, float* max_out, int* matched) {
  RefKeyFrame* K = static_cast<RefKeyFrame*>(pkf);
  std::vector<MapPoint*> pts, foreign;
  free_points(K->owner, n, world_pos, idx, desc, pts, normal_out, min_out, max_out);
  const unsigned char zero[32] = {0};
  make_points(K->owner, 1, nullptr, zero, nullptr, foreign);
  std::map<MapPoint*, int> index;
  for (int i = 0; i < n; ++i) { index[pts[i]] = i; if (bad && bad[i]) pts[i]->SetBadFlag(); }
  const int nk = K->owner->f()->NumKeypoints();
  std::vector<MapPoint*> vpMatched((size_t)nk, nullptr);
  for (int k = 0; k < nk; ++k) if (matched_in[k]) vpMatched[k] = foreign[0];
  OrbMatcher matcher(0.75f, true);
  const int nm = matcher.SearchByProjection(K->kf.get(), make_sim3(t), pts, vpMatched, th);
  for (int k = 0; k < nk; ++k) {
    std::map<MapPoint*, int>::const_iterator it = index.find(vpMatched[k]);
    matched[k] = it == index.end() ? -1 : it->second;
  }
  return nm;
}

// OrbMatcher::Fuse(KeyFrame*, cv::Mat Scw, vpPoints, th, vpReplacePoint) (orb_matcher.cpp:956-1079): best_idx[i] = the keypoint
// point i was fused into (read back from vpReplacePoint / the new observation), else -1
int ref_fuse_sim3(void* pkf, int n, const float* world_pos, const int* idx, const unsigned char* desc, const unsigned char* bad,
                  const float* t, float th, float* normal_out, float* min_out, float* max_out, int* best_idx) {
  RefKeyFrame* K = static_cast<RefKeyFrame*>(pkf);
  std::vector<MapPoint*> pts;
  free_points(K->owner, n, world_pos, idx, desc, pts, normal_out, min_out, max_out);
  for (int i = 0; i < n; ++i) if (bad && bad[i]) pts[i]->SetBadFlag();
  std::vector<MapPoint*> vpReplace((size_t)n, nullptr);
  OrbMatcher matcher(0.75f, true);
  const int nf = matcher.Fuse(K->kf.get(), make_sim3(t), pts, th, vpReplace);
  for (int i = 0; i < n; ++i) {
    if (vpReplace[i]) best_idx[i] = vpReplace[i]->GetIndexInKeyFrame(K->kf.get());
    else best_idx[i] = pts[i]->GetIndexInKeyFrame(K->kf.get());
  }
  return nf;
}

// OrbMatcher::SearchByProjection(Frame& Cur, KeyFrame*, sAlreadyFound, th, ORBdist) (orb_matcher.cpp:1455-1582).  The KeyFrame's
// map points (ref_keyframe_create_at) are projected with the current frame's pose (translation cur_t).
int ref_search_by_projection_keyframe(void* pcur, void* pkf, const unsigned char* already_found, const float* cur_t,
                                      const unsigned char* occupied, float th, int orb_dist, int check_ori, int* assigned) {
  RefFrame* RC = static_cast<RefFrame*>(pcur);
  RefKeyFrame* K = static_cast<RefKeyFrame*>(pkf);
  Frame* C = RC->f();
  C->SetPose(make_pose(cur_t));
  set_occupied(RC, occupied);
  std::set<MapPoint*> found;
  std::map<MapPoint*, int> index;
  for (size_t i = 0; i < K->pts.size(); ++i)
    if (K->pts[i]) { index[K->pts[i]] = (int)i; if (already_found && already_found[i]) found.insert(K->pts[i]); }
  OrbMatcher matcher(0.9f, check_ori != 0);
  const int nm = matcher.SearchByProjection(*C, K->kf.get(), found, th, orb_dist);
  for (int k = 0; k < C->NumKeypoints(); ++k) {
    std::map<MapPoint*, int>::const_iterator it = index.find(C->GetMapPoint(k));
    assigned[k] = it == index.end() ? -1 : it->second;
  }
  return nm;
}

// a KeyFrame whose keypoint i holds a map point at world_pos[i] (valid[i]), created from the origin with the keypoint's own
// octave and descriptor desc[i]; returns what the points hold
void* ref_keyframe_create_at(void* pframe, const unsigned char* valid, const float* world_pos, const unsigned char* desc,
                             const float* translation, float* normal_out, float* min_out, float* max_out) {
  RefFrame* R = static_cast<RefFrame*>(pframe);
  RefKeyFrame* K = new RefKeyFrame();
  K->owner = R;
  K->kf.reset(new KeyFrame(*R->f(), R->map, std::shared_ptr<KeyframeDatabase>()));
  K->kf->SetPose(make_pose(translation));
  const int n = R->f()->NumKeypoints();
  std::vector<int> idx((size_t)n);
  for (int i = 0; i < n; ++i) idx[i] = i;
  std::vector<MapPoint*> pts;
  free_points(R, n, world_pos, idx.data(), desc, pts, normal_out, min_out, max_out);
  K->pts.assign((size_t)n, nullptr);
  for (int i = 0; i < n; ++i)
    if (valid[i]) { K->pts[i] = pts[i]; K->kf->AddMapPoint(pts[i], i); pts[i]->AddObservation(K->kf.get(), i); }
  return K;
}

// OrbMatcher::SearchBySim3 (orb_matcher.cpp:1081-1310) between two KeyFrames built by ref_keyframe_create_at; s12 = 1, R12 = I,
// translation t12; pre_matched1[i] pre-fills vpMatches12[i] (vbAlreadyMatched, :1116-1126) with KF2's point pre_idx2[i]
int ref_search_by_sim3(void* pkf1, void* pkf2, const float* t12, float th, const int* pre_idx2, int* match12) {
  RefKeyFrame *K1 = static_cast<RefKeyFrame*>(pkf1), *K2 = static_cast<RefKeyFrame*>(pkf2);
  const int n1 = K1->owner->f()->NumKeypoints();
  std::map<MapPoint*, int> index2;
  for (size_t i = 0; i < K2->pts.size(); ++i) if (K2->pts[i]) index2[K2->pts[i]] = (int)i;
  std::vector<MapPoint*> m12((size_t)n1, nullptr);
  for (int i = 0; i < n1; ++i) if (pre_idx2 && pre_idx2[i] >= 0 && K2->pts[pre_idx2[i]]) m12[i] = K2->pts[pre_idx2[i]];
  cv::Mat R12 = cv::Mat::eye(3, 3, CV_32F);
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) R12.at<float>(r, c) = g_rot[3 * r + c];
  OrbMatcher matcher(0.75f, true);
  const int nf = matcher.SearchBySim3(K1->kf.get(), K2->kf.get(), m12, g_sim3_scale, R12, make_vec3(t12), th);
  for (int i = 0; i < n1; ++i) {
    std::map<MapPoint*, int>::const_iterator it = index2.find(m12[i]);
    match12[i] = it == index2.end() ? -1 : it->second;
  }
  return nf;
}


// OrbMatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) (orb_matcher.cpp:804-954) on a KeyFrame placed at translation t.
// best_idx[i] = the keypoint point i was fused into: read back from its new observation, or -- when an earlier point had
// taken that keypoint and the two were merged (:933-943) -- from the point that replaced it.
int ref_fuse(void* pkf, int n, const float* world_pos, const int* idx, const unsigned char* desc, const unsigned char* bad,
             const float* t, float th, float* normal_out, float* min_out, float* max_out, int* best_idx) {
  RefKeyFrame* K = static_cast<RefKeyFrame*>(pkf);
  std::vector<MapPoint*> pts;
  free_points(K->owner, n, world_pos, idx, desc, pts, normal_out, min_out, max_out);
  for (int i = 0; i < n; ++i) if (bad && bad[i]) pts[i]->SetBadFlag();
  K->kf->SetPose(make_pose(t));
  OrbMatcher matcher(0.75f, true);
  const int nf = matcher.Fuse(K->kf.get(), pts, th);
  for (int i = 0; i < n; ++i) {
    MapPoint* p = pts[i];
    if (p->isBad() && p->GetReplaced()) p = p->GetReplaced();
    best_idx[i] = (bad && bad[i]) ? -1 : p->GetIndexInKeyFrame(K->kf.get());
  }
  return nf;
}


// Throughput of the reference's own stereo front-end (Frame's stereo constructor: 2 extraction threads + ComputeStereoMatches) over
// n_pairs pairs with n_workers concurrent constructions; wall seconds.  bench.py --impl reference only.  Runs on malloc/free
// (ref_set_malloc_mode): timing does not depend on the quad-tree tie-break, and the bump arena never reclaims memory.
extern "C" void ref_set_malloc_mode(int on);
double ref_bench_stereo_batch(const unsigned char* const* left, const unsigned char* const* right, int n_pairs, int w, int h,
                              int nfeatures, float sf, int nl, int ini, int mn, float fx, float fy, float cx, float cy, float bf,
                              int n_workers, long* total_kps, long* total_matches) {
  ref_set_malloc_mode(1);
  std::atomic<int> next(0);
  std::atomic<long> kps(0), matches(0);
  const auto t0 = std::chrono::steady_clock::now();
  auto worker = [&]() {
    std::shared_ptr<ORBextractor> exL(new ORBextractor(nfeatures, sf, nl, ini, mn)), exR(new ORBextractor(nfeatures, sf, nl, ini, mn));
    cv::Mat K = make_K(fx, fy, cx, cy), D = cv::Mat::zeros(4, 1, CV_32F);
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n_pairs) break;
      cv::Mat imL(h, w, CV_8UC1, const_cast<unsigned char*>(left[i])), imR(h, w, CV_8UC1, const_cast<unsigned char*>(right[i]));
      Frame F(imL, imR, 0.0, exL, exR, std::shared_ptr<OrbVocabulary>(), K, D, bf, 35.0f);
      kps += F.NumKeypoints() + (long)F.GetRightKeys().size();
      long m = 0;
      for (size_t k = 0; k < F.StereoCoordRight().size(); ++k) m += F.StereoCoordRight()[k] >= 0;
      matches += m;
    }
  };
  {
    std::vector<std::thread> pool;
    for (int t = 0; t < n_workers; ++t) pool.emplace_back(worker);
    for (size_t t = 0; t < pool.size(); ++t) pool[t].join();
  }
  const auto t1 = std::chrono::steady_clock::now();
  ref_set_malloc_mode(0);
  if (total_kps) *total_kps = kps.load();
  if (total_matches) *total_matches = matches.load();
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
