// dropin/orb_features/orb_matcher.cpp -- drop-in for the reference's src/orb_features/orb_matcher.cpp: class OrbMatcher exactly as
// declared in the reference's own header (src/orb_features/orb_matcher.h:14-119, included unmodified), every search routine
// running on the GPU through include/orbfe_shim.hpp (C ABI: include/orbfe.h).  Tracking, LocalMapping and LoopClosing call these
// members as before and link unchanged.
//
// Division of labour, per routine: what the reference computes with cv::Mat BEFORE it calls GetFeaturesInArea -- the projection
// of a map point with the pose / Sim3, the depth-sign, image-bound, distance-invariance and viewing-angle gates,
// MapPoint::PredictScale -- is evaluated here with the same cv::Mat expressions in the same order (so it is bit-identical to
// the reference on the same OpenCV); the window gather, the Hamming distances, the best / second-best selection with the
// reference's serial feedback semantics and the rotation-histogram consistency check run on the device; the side effects
// (SetMapPoint, Replace, AddObservation, the output vectors) are applied here in the reference's order.
#include "orb_features/orb_matcher.h"

#include "orbfe_shim.hpp"

#include <cmath>
#include <deque>

const int OrbMatcher::TH_HIGH = 100;    // orb_matcher.cpp:5-7
const int OrbMatcher::TH_LOW = 50;
const int OrbMatcher::HISTO_LENGTH = 30;

namespace {

// KeyFrame keeps its image bounds in protected statics (src/data/keyframe.h:181-184); a derived class may read them
struct KeyFrameBounds : public KeyFrame {
  static orbfe::ImageBounds get() {
    orbfe::ImageBounds b;
    b.minX = (float)min_x_; b.maxX = (float)max_x_; b.minY = (float)min_y_; b.maxY = (float)max_y_;
    return b;
  }
};

// [R | t] of a 4x4 pose or similarity as the reference slices it
struct Pose {
  cv::Mat Rcw, tcw, Ow;
};
// Scw -> Rcw, tcw, Ow as orb_matcher.cpp:392-397 / 964-969
Pose decompose_sim3(const cv::Mat& Scw) {
  Pose p;
  cv::Mat sRcw = Scw.rowRange(0, 3).colRange(0, 3);
  const float scw = std::sqrt(sRcw.row(0).dot(sRcw.row(0)));
  p.Rcw = sRcw / scw;
  p.tcw = Scw.rowRange(0, 3).col(3) / scw;
  p.Ow = -p.Rcw.t() * p.tcw;
  return p;
}

}  // namespace

OrbMatcher::OrbMatcher(float nnratio, bool checkOri) : mfNNratio(nnratio), mbCheckOrientation(checkOri) {}

// orb_matcher.cpp:1630-1646
int OrbMatcher::DescriptorDistance(const cv::Mat& a, const cv::Mat& b) {
  return orbfe::DescriptorDistances(a, b)[0];
}

float OrbMatcher::RadiusByViewingCos(const float& viewCos) { return viewCos > 0.998 ? 2.5f : 4.0f; }  // :105-111 (done on the device)

// :113-131; the device applies the same test inside SearchForTriangulation, this member is kept for link compatibility
bool OrbMatcher::CheckDistEpipolarLine(const cv::KeyPoint& kp1, const cv::KeyPoint& kp2, const cv::Mat& F12, const KeyFrame* pKF2) {
  const float a = kp1.pt.x * F12.at<float>(0, 0) + kp1.pt.y * F12.at<float>(1, 0) + F12.at<float>(2, 0);
  const float b = kp1.pt.x * F12.at<float>(0, 1) + kp1.pt.y * F12.at<float>(1, 1) + F12.at<float>(2, 1);
  const float c = kp1.pt.x * F12.at<float>(0, 2) + kp1.pt.y * F12.at<float>(1, 2) + F12.at<float>(2, 2);
  const float num = a * kp2.pt.x + b * kp2.pt.y + c, den = a * a + b * b;
  if (den == 0) return false;
  return num * num / den < 3.84 * pKF2->level_sigma_sq[kp2.octave];
}

// :1584-1625; the rotation histograms live on the device, this member is kept for link compatibility
void OrbMatcher::ComputeThreeMaxima(std::vector<int>* histo, const int L, int& ind1, int& ind2, int& ind3) {
  int max1 = 0, max2 = 0, max3 = 0;
  for (int i = 0; i < L; i++) {
    const int s = (int)histo[i].size();
    if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
    else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
    else if (s > max3) { max3 = s; ind3 = i; }
  }
  if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
  else if (max3 < 0.1f * (float)max1) ind3 = -1;
}

// ---- Tracking ------------------------------------------------------------------------------------------------------------

// :13-103
int OrbMatcher::SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const int th) {
  return orbfe::SearchByProjection(F, vpMapPoints, th, mfNNratio);
}

// :1312-1453
int OrbMatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono) {
  const cv::Mat Rcw = CurrentFrame.GetPose().rowRange(0, 3).colRange(0, 3);
  const cv::Mat tcw = CurrentFrame.GetPose().rowRange(0, 3).col(3);
  const cv::Mat twc = -Rcw.t() * tcw;
  const cv::Mat Rlw = LastFrame.GetPose().rowRange(0, 3).colRange(0, 3);
  const cv::Mat tlw = LastFrame.GetPose().rowRange(0, 3).col(3);
  const cv::Mat tlc = Rlw * twc + tlw;
  const bool bForward = tlc.at<float>(2) > CurrentFrame.GetBaseline() && !bMono;    // :1334-1335
  const bool bBackward = -tlc.at<float>(2) > CurrentFrame.GetBaseline() && !bMono;
  const float fx = CurrentFrame.GetFx(), fy = CurrentFrame.GetFy(), cx = CurrentFrame.GetCx(), cy = CurrentFrame.GetCy();
  return orbfe::SearchByProjectionLastFrame(
      CurrentFrame, LastFrame, th, bForward, bBackward, mbCheckOrientation, [&](int i, float& u, float& v, float& invzc) -> bool {
        MapPoint* pMP = LastFrame.GetMapPoint(i);
        if (!pMP || LastFrame.IsOutlier(i)) return false;      // :1340-1342
        cv::Mat x3Dw = pMP->GetWorldPos();                      // :1344-1356; invzc < 0 and the image bounds are tested by the library
        cv::Mat x3Dc = Rcw * x3Dw + tcw;
        const float xc = x3Dc.at<float>(0), yc = x3Dc.at<float>(1);
        invzc = 1.0 / x3Dc.at<float>(2);
        u = fx * xc * invzc + cx;
        v = fy * yc * invzc + cy;
        return true;
      });
}

// :1455-1582 (relocalisation)
int OrbMatcher::SearchByProjection(Frame& CurrentFrame, KeyFrame* pKF, const std::set<MapPoint*>& sAlreadyFound, const float th,
                                   const int ORBdist) {
  const cv::Mat Rcw = CurrentFrame.GetPose().rowRange(0, 3).colRange(0, 3);
  const cv::Mat tcw = CurrentFrame.GetPose().rowRange(0, 3).col(3);
  const cv::Mat Ow = -Rcw.t() * tcw;
  const std::vector<MapPoint*> vpMPs = pKF->GetMapPointMatches();
  const float fx = CurrentFrame.GetFx(), fy = CurrentFrame.GetFy(), cx = CurrentFrame.GetCx(), cy = CurrentFrame.GetCy();
  return orbfe::SearchByProjectionKeyFrame(
      CurrentFrame, pKF, th, ORBdist, mbCheckOrientation, [&](size_t i, float& u, float& v, int32_t& level) -> bool {
        MapPoint* pMP = vpMPs[i];
        if (!pMP || pMP->isBad() || sAlreadyFound.count(pMP)) return false;   // :1474-1478
        cv::Mat x3Dw = pMP->GetWorldPos();                                     // :1480-1488
        cv::Mat x3Dc = Rcw * x3Dw + tcw;
        const float xc = x3Dc.at<float>(0), yc = x3Dc.at<float>(1);
        const float invzc = 1.0 / x3Dc.at<float>(2);
        u = fx * xc * invzc + cx;
        v = fy * yc * invzc + cy;
        // the image-bound test of :1490-1495 is applied by the library; the reference evaluates it before the distance gate,
        // which has no side effect, so the order is immaterial -- but PredictScale must not see a point outside the bounds
        if (u < CurrentFrame.GetMinX() || u > CurrentFrame.GetMaxX() || v < CurrentFrame.GetMinY() || v > CurrentFrame.GetMaxY()) return false;
        cv::Mat PO = x3Dw - Ow;                                                // :1498-1508
        const float dist3D = cv::norm(PO);
        const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
        if (dist3D < minDistance || dist3D > maxDistance) return false;
        level = pMP->PredictScale(dist3D, &CurrentFrame);
        return true;
      });
}

// :264-382 (monocular initialisation)
int OrbMatcher::SearchForInitialization(Frame& F1, Frame& F2, std::vector<cv::Point2f>& vbPrevMatched, std::vector<int>& vnMatches12,
                                        int windowSize) {
  return orbfe::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize, mfNNratio, mbCheckOrientation);
}

// :133-262
int OrbMatcher::SearchByBoW(KeyFrame* pKF, Frame& F, std::vector<MapPoint*>& vpMapPointMatches) {
  return orbfe::SearchByBoW(pKF, F, vpMapPointMatches, mfNNratio, mbCheckOrientation);
}

// ---- Loop closing / local mapping -------------------------------------------------------------------------------------------

// :499-632
int OrbMatcher::SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12) {
  return orbfe::SearchByBoW(pKF1, pKF2, vpMatches12, KeyFrameBounds::get(), mfNNratio, mbCheckOrientation);
}

// :634-802
int OrbMatcher::SearchForTriangulation(KeyFrame* pKF1, KeyFrame* pKF2, const cv::Mat& F12,
                                       std::vector<std::pair<size_t, size_t>>& vMatchedPairs, const bool bOnlyStereo) {
  // epipole in the second image, :642-649
  cv::Mat Cw = pKF1->GetCameraCenter();
  cv::Mat R2w = pKF2->GetRotation();
  cv::Mat t2w = pKF2->GetTranslation();
  cv::Mat C2 = R2w * Cw + t2w;
  const float invz = 1.0f / C2.at<float>(2);
  const float ex = pKF2->fx * C2.at<float>(0) * invz + pKF2->cx;
  const float ey = pKF2->fy * C2.at<float>(1) * invz + pKF2->cy;
  float f[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) f[3 * r + c] = F12.at<float>(r, c);
  return orbfe::SearchForTriangulation(pKF1, pKF2, f, ex, ey, vMatchedPairs, bOnlyStereo, KeyFrameBounds::get(), mbCheckOrientation);
}

namespace {
// the gate shared by the projection searches into a KeyFrame (:408-451, :823-866, :981-1025): camera coordinates with
// (Rcw, tcw), positive depth, projection, KeyFrame::IsInImage, distance invariance, viewing angle below 60 degrees, predicted level.
// `invz_double`: the routine forms 1/z as `1.0/z` (double divide, then narrowed) instead of `1/z` (float divide).
bool project_into_keyframe(MapPoint* pMP, KeyFrame* pKF, const Pose& T, bool invz_double, float& u, float& v, float& invz, int32_t& level) {
  cv::Mat p3Dw = pMP->GetWorldPos();
  cv::Mat p3Dc = T.Rcw * p3Dw + T.tcw;
  if (p3Dc.at<float>(2) < 0.0f) return false;
  if (invz_double) invz = 1.0 / p3Dc.at<float>(2); else invz = 1 / p3Dc.at<float>(2);
  const float x = p3Dc.at<float>(0) * invz, y = p3Dc.at<float>(1) * invz;
  u = pKF->fx * x + pKF->cx;
  v = pKF->fy * y + pKF->cy;
  if (!pKF->IsInImage(u, v)) return false;
  const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
  cv::Mat PO = p3Dw - T.Ow;
  const float dist = cv::norm(PO);
  if (dist < minDistance || dist > maxDistance) return false;
  cv::Mat Pn = pMP->GetNormal();
  if (PO.dot(Pn) < 0.5 * dist) return false;
  level = pMP->PredictScale(dist, pKF);
  return true;
}
}  // namespace

// :384-497 (loop detection)
int OrbMatcher::SearchByProjection(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, std::vector<MapPoint*>& vpMatched,
                                   int th) {
  const Pose T = decompose_sim3(Scw);
  std::set<MapPoint*> spAlreadyFound(vpMatched.begin(), vpMatched.end());   // :399-401
  spAlreadyFound.erase(static_cast<MapPoint*>(nullptr));
  return orbfe::SearchByProjection(pKF, vpPoints, vpMatched, th, KeyFrameBounds::get(),
                                   [&](size_t i, float& u, float& v, int32_t& level) -> bool {
                                     MapPoint* pMP = vpPoints[i];
                                     if (pMP->isBad() || spAlreadyFound.count(pMP)) return false;   // :411-412
                                     float invz;
                                     return project_into_keyframe(pMP, pKF, T, false, u, v, invz, level);
                                   });
}

// :804-954 (local mapping)
int OrbMatcher::Fuse(KeyFrame* pKF, const std::vector<MapPoint*>& vpMapPoints, const float th) {
  Pose T;
  T.Rcw = pKF->GetRotation();
  T.tcw = pKF->GetTranslation();
  T.Ow = pKF->GetCameraCenter();
  const float bf = pKF->mbf;
  return orbfe::Fuse(pKF, vpMapPoints, th, KeyFrameBounds::get(),
                     [&](size_t i, float& u, float& v, float& ur, int32_t& level) -> bool {
                       MapPoint* pMP = vpMapPoints[i];                        // null pointers are skipped by the adapter (:824-825)
                       if (pMP->isBad() || pMP->IsInKeyFrame(pKF)) return false;   // :827-828
                       float invz;
                       if (!project_into_keyframe(pMP, pKF, T, false, u, v, invz, level)) return false;
                       ur = u - bf * invz;                                     // :849
                       return true;
                     });
}

// :956-1079 (loop closing)
int OrbMatcher::Fuse(KeyFrame* pKF, cv::Mat Scw, const std::vector<MapPoint*>& vpPoints, float th, std::vector<MapPoint*>& vpReplacePoint) {
  const Pose T = decompose_sim3(Scw);
  const std::set<MapPoint*> spAlreadyFound = pKF->GetMapPoints();            // :971
  return orbfe::Fuse(pKF, vpPoints, th, vpReplacePoint, KeyFrameBounds::get(),
                     [&](size_t i, float& u, float& v, int32_t& level) -> bool {
                       MapPoint* pMP = vpPoints[i];
                       if (pMP->isBad() || spAlreadyFound.count(pMP)) return false;   // :984-985
                       float invz;
                       return project_into_keyframe(pMP, pKF, T, true, u, v, invz, level);
                     });
}

// :1081-1310 (loop closing)
int OrbMatcher::SearchBySim3(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, const float s12, const cv::Mat& R12,
                             const cv::Mat& t12, const float th) {
  const float fx = pKF1->fx, fy = pKF1->fy, cx = pKF1->cx, cy = pKF1->cy;   // :1089-1092
  cv::Mat R1w = pKF1->GetRotation(), t1w = pKF1->GetTranslation();           // :1094-1100
  cv::Mat R2w = pKF2->GetRotation(), t2w = pKF2->GetTranslation();
  cv::Mat sR12 = s12 * R12;                                                  // :1102-1105
  cv::Mat sR21 = (1.0 / s12) * R12.t();
  cv::Mat t21 = -sR21 * t12;
  const std::vector<MapPoint*> vpMapPoints1 = pKF1->GetMapPointMatches(), vpMapPoints2 = pKF2->GetMapPointMatches();
  const int N1 = (int)vpMapPoints1.size(), N2 = (int)vpMapPoints2.size();
  std::deque<bool> vbAlreadyMatched1(N1, false), vbAlreadyMatched2(N2, false);   // :1113-1126
  for (int i = 0; i < N1; i++) {
    MapPoint* pMP = vpMatches12[i];
    if (pMP) {
      vbAlreadyMatched1[i] = true;
      const int idx2 = pMP->GetIndexInKeyFrame(pKF2);
      if (idx2 >= 0 && idx2 < N2) vbAlreadyMatched2[idx2] = true;
    }
  }
  // one direction: the point seen in `from` (Rw, tw) carried into the other camera by (sR, t), projected into `into`
  auto gate = [&](MapPoint* pMP, const cv::Mat& Rw, const cv::Mat& tw, const cv::Mat& sR, const cv::Mat& t, KeyFrame* into, float& u,
                  float& v, int32_t& level) -> bool {
    if (pMP->isBad()) return false;
    cv::Mat p3Dw = pMP->GetWorldPos();
    cv::Mat p3Da = Rw * p3Dw + tw;
    cv::Mat p3Db = sR * p3Da + t;
    if (p3Db.at<float>(2) < 0.0) return false;
    const float invz = 1.0 / p3Db.at<float>(2);
    const float x = p3Db.at<float>(0) * invz, y = p3Db.at<float>(1) * invz;
    u = fx * x + cx;
    v = fy * y + cy;
    if (!into->IsInImage(u, v)) return false;
    const float maxDistance = pMP->GetMaxDistanceInvariance(), minDistance = pMP->GetMinDistanceInvariance();
    const float dist3D = cv::norm(p3Db);
    if (dist3D < minDistance || dist3D > maxDistance) return false;
    level = pMP->PredictScale(dist3D, into);
    return true;
  };
  return orbfe::SearchBySim3(
      pKF1, pKF2, vpMatches12, th, KeyFrameBounds::get(),
      [&](size_t i1, float& u, float& v, int32_t& level) -> bool {   // :1131-1170, KF1's points into KF2
        return !vbAlreadyMatched1[i1] && gate(vpMapPoints1[i1], R1w, t1w, sR21, t21, pKF2, u, v, level);
      },
      [&](size_t i2, float& u, float& v, int32_t& level) -> bool {   // :1211-1250, KF2's points into KF1
        return !vbAlreadyMatched2[i2] && gate(vpMapPoints2[i2], R2w, t2w, sR12, t12, pKF1, u, v, level);
      });
}
