// shim_demo.cpp -- exercises include/orbfe_shim.hpp the way the reference's Frame does
// (src/data/frame.cpp:61-111): two ORBextractor objects run Compute on two std::threads, then
// ComputeStereoMatches.  Reads two raw u8 images, writes keypoints / descriptors / stereo results as raw
// binary so that the Python test can compare them with the oracle.  usage:
//   shim_demo W H left.raw right.raw out_prefix      |      shim_demo --link-check
#include "orbfe_shim.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>

static std::vector<uint8_t> slurp(const char* path, size_t n) {
  std::vector<uint8_t> v(n);
  FILE* f = fopen(path, "rb");
  if (!f || fread(v.data(), 1, n, f) != n) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
  fclose(f);
  return v;
}

// ---- stand-ins for the reference's MapPoint / KeyFrame / Frame (only what the N1 adapters touch:
// src/data/map_point.h, keyframe.h:64-71,147-170, frame.h:104-190) so that every adapter template is instantiated --------
#include <map>
#include <set>
struct MockKeyFrame;
struct MockMapPoint {
  cv::Mat desc;
  bool bad = false;
  int nobs = 1;
  bool track_is_in_view = false;  // map_point.h tracking fields written by Frame::IsInFrustum
  float track_projected_x = 0, track_projected_y = 0, track_projected_x_right = 0, track_view_cos = 0;
  int track_scale_level = 0;
  std::set<MockKeyFrame*> in;
  bool isBad() const { return bad; }
  cv::Mat GetDescriptor() const { return desc; }
  int NumObservations() const { return nobs; }
  bool IsInKeyFrame(MockKeyFrame* kf) const { return in.count(kf) != 0; }
  void AddObservation(MockKeyFrame* kf, size_t) { in.insert(kf); ++nobs; }
  void Replace(MockMapPoint* by) { bad = true; by->in.insert(in.begin(), in.end()); }
};
struct MockKeyFrame {
  std::vector<cv::KeyPoint> undistorted_keypoints;
  cv::Mat descriptors;
  std::vector<float> right_coords, scale_factors;
  std::map<unsigned, std::vector<unsigned> > feature_vec;
  std::vector<MockMapPoint*> mps;
  std::vector<MockMapPoint*> GetMapPointMatches() const { return mps; }
  MockMapPoint* GetMapPoint(size_t i) const { return mps[i]; }
  void AddMapPoint(MockMapPoint* p, size_t i) { mps[i] = p; }
};
struct MockFrame {
  std::vector<cv::KeyPoint> kps;
  cv::Mat desc;
  std::vector<float> ur, scale;
  std::vector<MockMapPoint*> mps;
  std::map<unsigned, std::vector<unsigned> > fv;
  float w, h;
  const std::vector<cv::KeyPoint>& GetUndistortedKeys() const { return kps; }
  const cv::Mat& GetDescriptors() const { return desc; }
  const std::vector<float>& StereoCoordRight() const { return ur; }
  const std::vector<float>& ScaleFactors() const { return scale; }
  const std::map<unsigned, std::vector<unsigned> >& GetFeatureVector() const { return fv; }
  float GetFx() const { return 718.856f; }
  float GetFy() const { return 718.856f; }
  float GetCx() const { return 607.1928f; }
  float GetCy() const { return 185.2157f; }
  float GetBaselineFx() const { return 386.1448f; }
  float GetLogScaleFactor() const { return std::log(1.2f); }
  float GetMinX() const { return 0; }
  float GetMaxX() const { return w; }
  float GetMinY() const { return 0; }
  float GetMaxY() const { return h; }
  int NumKeypoints() const { return (int)kps.size(); }
  MockMapPoint* GetMapPoint(int i) const { return mps[i]; }
  void SetMapPoint(int i, MockMapPoint* p) { mps[i] = p; }
};

// KeyFrame = the extracted frame, one map point per keypoint; every routine is asked to find the frame in itself
static int n1_selfcheck(const std::vector<cv::KeyPoint>& kps, const cv::Mat& desc, const ORBextractor& ex, int W, int H) {
  const size_t n = kps.size();
  std::vector<MockMapPoint> pts(n);
  MockKeyFrame kf;
  kf.undistorted_keypoints = kps; kf.descriptors = desc; kf.right_coords.assign(n, -1.0f); kf.scale_factors = ex.GetScaleFactors();
  kf.mps.resize(n);
  for (size_t i = 0; i < n; ++i) {
    pts[i].desc = desc.row((int)i).clone();
    kf.mps[i] = &pts[i];
    kf.feature_vec[desc.ptr((int)i)[0] % 50u].push_back((unsigned)i);
  }
  MockFrame F;
  F.kps = kps; F.desc = desc; F.scale = kf.scale_factors; F.mps.assign(n, nullptr); F.fv = kf.feature_vec; F.w = (float)W; F.h = (float)H;
  const orbfe::ImageBounds b = {0.f, (float)W, 0.f, (float)H};
  auto gate = [&](size_t i, float& u, float& v, int32_t& lvl) { u = kps[i].pt.x; v = kps[i].pt.y; lvl = kps[i].octave; return true; };
  auto gate_ur = [&](size_t i, float& u, float& v, float& ur, int32_t& lvl) { ur = kps[i].pt.x - 10.f; return gate(i, u, v, lvl); };
  std::vector<MockMapPoint*> m1, m2(n, nullptr), m3(n, nullptr), rep(n, nullptr), mpv(kf.mps);
  const int a = orbfe::SearchByBoW(&kf, F, m1, 0.9f, true);
  std::vector<MockMapPoint*> m12;
  const int c = orbfe::SearchByBoW(&kf, &kf, m12, b, 0.9f, true);
  MockKeyFrame bare = kf;                       // no map points: everything is a triangulation candidate
  bare.mps.assign(n, nullptr);
  const float F12[9] = {0, 0, 0, 0, 0, 1, 0, -1, 0};  // pure x-translation: epipolar lines y2 = y1
  std::vector<std::pair<size_t, size_t> > pairs;
  const int d = orbfe::SearchForTriangulation(&bare, &bare, F12, 1e6f, 0.f, pairs, false, b, true);
  const int e = orbfe::SearchByProjection(&kf, mpv, m2, 10, b, gate);
  const int f = orbfe::SearchByProjectionKeyFrame(F, &kf, 10.f, 100, true, gate);
  const int g = orbfe::SearchBySim3(&kf, &kf, m3, 7.5f, b, gate, gate);
  const int h2 = orbfe::Fuse(&kf, mpv, 3.f, rep, b, gate);
  MockKeyFrame other = kf;                      // Fuse(KF, points): the points are not yet in `other`
  const int h1 = orbfe::Fuse(&other, mpv, 3.f, b, gate_ur);
  // N2: undistortion with a real distortion model, and the SearchLocalPoints visibility loop (identity pose, points
  // back-projected 10 m in front of the camera: every one must be visible at its own pixel)
  std::vector<cv::KeyPoint> und;
  orbfe::UndistortKeyPoints(kps, 718.856f, 718.856f, 607.1928f, 185.2157f, std::vector<float>{-0.2834f, 0.0739f, 0.0002f, 0.00002f}, und);
  int moved = 0;
  for (size_t i = 0; i < n; ++i) moved += und[i].pt.x != kps[i].pt.x || und[i].pt.y != kps[i].pt.y;
  const float I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Z3[3] = {0, 0, 0};
  const float fx = 718.856f, cx0 = 607.1928f, cy0 = 185.2157f;
  auto fetch = [&](size_t i, float* Pw, float* Pn, float& mn, float& mx, float& raw) {
    Pw[0] = (kps[i].pt.x - cx0) / fx * 10.f; Pw[1] = (kps[i].pt.y - cy0) / fx * 10.f; Pw[2] = 10.f;
    const float len = std::sqrt(Pw[0] * Pw[0] + Pw[1] * Pw[1] + Pw[2] * Pw[2]);
    for (int c2 = 0; c2 < 3; ++c2) Pn[c2] = Pw[c2] / len;
    raw = len * std::pow(1.2f, (float)kps[i].octave);  // PredictScale then lands on the keypoint's own octave (or one above)
    mn = 1.f; mx = 1.2f * raw;
    return true;
  };
  const int vis = orbfe::IsInFrustumBatch(mpv, I3, Z3, Z3, fx, fx, cx0, cy0, 386.1448f, b, std::log(1.2f), 8, 0.5f, fetch);
  int near_px = 0;
  for (size_t i = 0; i < n; ++i)
    near_px += pts[i].track_is_in_view && std::fabs(pts[i].track_projected_x - kps[i].pt.x) < 0.01f && pts[i].track_view_cos > 0.99f;
  // SearchLocalPoints in one device-resident call: the same points must be visible and (nearly) all matched to their keypoints
  MockFrame F2 = F;
  F2.mps.assign(n, nullptr);
  int nm = 0;
  const int vis2 = orbfe::SearchLocalPoints(F2, mpv, I3, Z3, Z3, 1, 0.8f, 0.5f, fetch, &nm);
  if (vis2 != vis || nm < (int)(0.5 * n)) { fprintf(stderr, "SearchLocalPoints: %d visible, %d matched of %zu\n", vis2, nm, n); return 4; }
  printf("%zu %d %d %d %d %d %d %d %d %d %d %d\n", n, a, c, d, e, f, g, h1, h2, moved, vis, near_px);
  return 0;
}

// Frame::ComputeBoW through the shim's OrbVocabulary: prints |BowVector|, |FeatureVector|, sum of the BoW values
static int bow_check(const char* vocfile, const cv::Mat& desc) {
  orbfe::OrbVocabulary voc;
  if (!voc.loadFromTextFile(vocfile)) { fprintf(stderr, "%s\n", orbfe_last_error()); return 3; }
  std::vector<cv::Mat> rows;
  for (int i = 0; i < desc.rows; ++i) rows.push_back(desc.row(i));
  std::map<unsigned, double> bow;
  std::map<unsigned, std::vector<unsigned> > fv;
  voc.transform(rows, bow, fv, 1);
  double sum = 0;
  size_t nfeat = 0;
  for (std::map<unsigned, double>::const_iterator it = bow.begin(); it != bow.end(); ++it) sum += it->second;
  for (std::map<unsigned, std::vector<unsigned> >::const_iterator it = fv.begin(); it != fv.end(); ++it) nfeat += it->second.size();
  printf("%zu %zu %zu %.17g\n", bow.size(), fv.size(), nfeat, sum);
  return 0;
}

template <class T>
static void dump(const std::string& path, const T* p, size_t n) {
  FILE* f = fopen(path.c_str(), "wb");
  fwrite(p, sizeof(T), n, f);
  fclose(f);
}

int main(int argc, char** argv) {
  if (argc == 2 && std::string(argv[1]) == "--link-check") {
    printf("%s devices=%d\n", orbfe_version(), orbfe_device_count());
    return 0;
  }
  if (argc == 5 && std::string(argv[1]) == "--n1") {  // shim_demo --n1 W H image.raw
    const int W = atoi(argv[2]), H = atoi(argv[3]);
    std::vector<uint8_t> im = slurp(argv[4], (size_t)W * H);
    cv::Mat img(H, W, CV_8UC1, im.data());
    ORBextractor ex(1000, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> k;
    cv::Mat dsc;
    ex(img, cv::Mat(), k, dsc);
    return n1_selfcheck(k, dsc, ex, W, H);
  }
  if (argc == 6 && std::string(argv[1]) == "--bow") {  // shim_demo --bow W H image.raw voc.txt
    const int W = atoi(argv[2]), H = atoi(argv[3]);
    std::vector<uint8_t> im = slurp(argv[4], (size_t)W * H);
    cv::Mat img(H, W, CV_8UC1, im.data());
    ORBextractor ex(1000, 1.2f, 8, 20, 7);
    std::vector<cv::KeyPoint> k;
    cv::Mat dsc;
    ex(img, cv::Mat(), k, dsc);
    return bow_check(argv[5], dsc);
  }
  if (argc == 7 && std::string(argv[1]) == "--latency") {  // shim_demo --latency W H left.raw right.raw N
    // the drop-in path exactly as Frame's stereo constructor drives it (frame.cpp:86-99): two std::threads per frame running
    // ORBextractor::Compute, then ComputeStereoMatches; prints p50 / min / max milliseconds per stereo frame over N frames
    const int W = atoi(argv[2]), H = atoi(argv[3]), N = atoi(argv[6]);
    std::vector<uint8_t> l = slurp(argv[4], (size_t)W * H), r = slurp(argv[5], (size_t)W * H);
    cv::Mat left(H, W, CV_8UC1, l.data()), right(H, W, CV_8UC1, r.data());
    ORBextractor exL(2000, 1.2f, 8, 20, 7), exR(2000, 1.2f, 8, 20, 7);
    std::vector<double> ms;
    size_t nkp = 0, nmatch = 0;
    for (int it = 0; it < N + 10; ++it) {
      std::vector<cv::KeyPoint> kl, kr;
      cv::Mat dl, dr;
      std::vector<float> ur, depth;
      const auto t0 = std::chrono::steady_clock::now();
      std::thread tl([&]() { exL.Compute(left, cv::Mat(), kl, dl); });
      std::thread tr([&]() { exR.Compute(right, cv::Mat(), kr, dr); });
      tl.join();
      tr.join();
      orbfe::ComputeStereoMatches(exL, exR, kl, kr, dl, dr, 386.1448f, 386.1448f / 718.856f, ur, depth);
      const auto t1 = std::chrono::steady_clock::now();
      if (it >= 10) ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
      nkp = kl.size() + kr.size();
      nmatch = 0;
      for (size_t i = 0; i < ur.size(); ++i) nmatch += ur[i] >= 0;
    }
    std::sort(ms.begin(), ms.end());
    printf("%.4f %.4f %.4f %zu %zu\n", ms[ms.size() / 2], ms.front(), ms.back(), nkp, nmatch);
    return 0;
  }
  if (argc != 6) return 1;
  const int W = atoi(argv[1]), H = atoi(argv[2]);
  std::vector<uint8_t> l = slurp(argv[3], (size_t)W * H), r = slurp(argv[4], (size_t)W * H);
  cv::Mat left(H, W, CV_8UC1, l.data()), right(H, W, CV_8UC1, r.data());
  ORBextractor exL(2000, 1.2f, 8, 20, 7), exR(2000, 1.2f, 8, 20, 7);
  std::vector<cv::KeyPoint> kl, kr;
  cv::Mat dl, dr;
  std::thread tl([&]() { exL.Compute(left, cv::Mat(), kl, dl); });
  std::thread tr([&]() { exR(right, cv::Mat(), kr, dr); });
  tl.join();
  tr.join();
  std::vector<float> ur, depth;
  const float bf = 386.1448f, fx = 718.856f;
  orbfe::ComputeStereoMatches(exL, exR, kl, kr, dl, dr, bf, bf / fx, ur, depth);
  const std::string out(argv[5]);
  dump(out + ".kl", kl.data(), kl.size());
  dump(out + ".kr", kr.data(), kr.size());
  dump(out + ".dl", dl.data, (size_t)dl.rows * 32);
  dump(out + ".dr", dr.data, (size_t)dr.rows * 32);
  dump(out + ".ur", ur.data(), ur.size());
  dump(out + ".depth", depth.data(), depth.size());
  const std::vector<cv::Mat>& pyr = exL.GetImagePyramid();
  dump(out + ".pyr3", pyr[3].data, (size_t)pyr[3].rows * pyr[3].cols);
  printf("%zu %zu %d %d\n", kl.size(), kr.size(), pyr[3].cols, pyr[3].rows);
  return 0;
}
