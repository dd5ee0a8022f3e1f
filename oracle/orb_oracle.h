/*
 * orb_oracle.h -- CPU ORACLE for the ORB front-end hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This is a plain C++17 restatement (no OpenCV, no CUDA) of the reference's
 *   src/orb_features/orb_extractor.cpp   (all but the dead ComputeKeyPointsOld :796-974)
 *   src/data/frame.cpp:406-577           (Frame::ComputeStereoMatches)
 *   src/data/frame.cpp:234-248,339-403   (feature grid, GetFeaturesInArea)
 *   src/orb_features/orb_matcher.cpp:13-111, 264-382, 1312-1453, 1584-1646
 * and of the OpenCV primitives those files call (resize, copyMakeBorder, FAST,
 * GaussianBlur, fastAtan2, cvRound), whose arithmetic lives in the un-vendored
 * third-party dependency OpenCV 3.x (find_package(OpenCV 3.0 REQUIRED), CMakeLists.txt:45).
 *
 * PARITY PIN (two layers): the OpenCV primitives are pinned bit-exactly
 * against the Python cv2 4.13.0 build of the same library (tests/test_oracle_primitives.py +
 * tests/golden/ fixtures and their generator scripts).  The control logic above them (grid
 * loop, quad-tree, orientation / descriptor loops, stereo search, grid lookup, every OrbMatcher
 * routine, IsInFrustum / PredictScale, DBoW2 transform) is pinned against the reference's OWN
 * sources, compiled unmodified from /root/reference into oracle/_ref by oracle/Makefile.ref
 * against the OpenCV stand-in oracle/cvstub (tests/test_reference_pin.py, bit-exact).
 * Caveat: cv::Mat::dot (IsInFrustum, viewing-angle gates) is the stand-in's restatement.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may load this library.  The product (slam_framework_b200/) never links or calls it.
 *
 * Oracle-defined behaviour where the reference is indeterminate:
 *   1. quad-tree sort ties (orb_extractor.cpp:625 sorts pair<int,ExtractorNode*> => ties
 *      ordered by heap address): ties are ordered by node creation sequence.
 *   2. baseline_ is read uninitialised in ComputeStereoMatches (frame.cpp:436 vs :108):
 *      the caller passes baseline explicitly (KITTI: bf/fx).
 *   3. median of an empty match list (frame.cpp:565-566) is UB: the cut is skipped.
 *   4. FP contraction: built with -ffp-contract=off (IEEE, no fused x*b+y*a).
 *   5. GaussianBlur: OpenCV >= 3.4.1 / 4.x fixed-point path (taps 18,34,48,56,48,34,18).
 *   6. rBRIEF samples reach 18 px from the keypoint but keypoints may sit 16 px from the level
 *      edge (orb_extractor.cpp:59-61 on the un-padded blurred clone, :1029): a column overshoot
 *      reads the adjacent row of the continuous clone (kept, it is deterministic); a sample
 *      before/after the whole buffer is an out-of-allocation read (UB) and is defined as 0.
 */
#ifndef ORB_ORACLE_H_
#define ORB_ORACLE_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cv::KeyPoint memory layout (28 bytes). */
typedef struct {
  float x, y, size, angle, response;
  int octave, class_id;
} orc_keypoint;

typedef struct orc_extractor orc_extractor;
typedef struct orc_frame orc_frame;
typedef struct orc_vocabulary orc_vocabulary;

/* ---- OpenCV primitives (Appendix A of SURVEY.md) ---------------------------------- */
void orc_resize_linear(const uint8_t* src, int sw, int sh, int sstride,
                       uint8_t* dst, int dw, int dh, int dstride);
/* fills the `border`-px frame of a (w+2b)x(h+2b) buffer whose interior is already set */
void orc_border_reflect101(uint8_t* buf, int w, int h, int stride, int border);
void orc_gaussian7x7(const uint8_t* src, int w, int h, int sstride, uint8_t* dst, int dstride);
int orc_fast9(const uint8_t* img, int w, int h, int stride, int threshold, int nms,
              orc_keypoint* out, int cap);
float orc_fast_atan2(float y, float x);
/* cv::cvtColor(..2GRAY), 8-bit, channels 3 or 4, rgb_order 1 = R first (tracker.cpp:110-127) */
void orc_cvt_gray(const uint8_t* src, size_t n_px, int channels, int rgb_order, uint8_t* dst);
int orc_descriptor_distance(const uint8_t* a, const uint8_t* b);

/* ---- extractor (orb_extractor.cpp) ------------------------------------------------ */
orc_extractor* orc_extractor_create(int nfeatures, float scale_factor, int nlevels,
                                    int ini_th_fast, int min_th_fast);
void orc_extractor_destroy(orc_extractor*);
/* ORBextractor::Compute; returns the keypoint count, or -(count) if cap is too small */
int orc_extract(orc_extractor*, const uint8_t* img, int w, int h, int stride,
                orc_keypoint* kps, uint8_t* desc, int cap);
int orc_levels(const orc_extractor*);
void orc_scale_factors(const orc_extractor*, float* scale, float* inv_scale,
                       float* sigma2, float* inv_sigma2);
void orc_features_per_level(const orc_extractor*, int* out);
void orc_umax(const orc_extractor*, int* out16);
/* ROI view of pyramid level (valid until the next orc_extract on this handle) */
const uint8_t* orc_pyramid_level(const orc_extractor*, int level, int* w, int* h, int* stride);
/* padded plane (w+38)x(h+38) of the level */
const uint8_t* orc_pyramid_padded(const orc_extractor*, int level, int* w, int* h, int* stride);
/* stage outputs of the last orc_extract, for per-stage parity tests */
int orc_stage_candidates(const orc_extractor*, int level, orc_keypoint* out, int cap);
int orc_stage_level_keypoints(const orc_extractor*, int level, orc_keypoint* out, int cap);
const uint8_t* orc_stage_blurred(const orc_extractor*, int level, int* w, int* h, int* stride);

/* DistributeOctTree on an explicit candidate list (coordinates relative to minX/minY) */
int orc_distribute_octree(const orc_keypoint* cand, int n, int minX, int maxX, int minY,
                          int maxY, int N, orc_keypoint* out, int cap);

/* ---- Frame::ComputeStereoMatches (frame.cpp:406-577) ------------------------------- */
/* uses the pyramids held by the two extractor handles (last orc_extract on each). */
int orc_stereo_match(const orc_extractor* left, const orc_extractor* right,
                     int nl, const orc_keypoint* kps_l, const uint8_t* desc_l,
                     int nr, const orc_keypoint* kps_r, const uint8_t* desc_r,
                     float bf, float baseline, float* u_right, float* depth);

/* ---- Frame grid + matchers (frame.cpp:234-248,339-403; orb_matcher.cpp) ------------ */
orc_frame* orc_frame_create(int n, const orc_keypoint* kps_un, const uint8_t* desc,
                            const float* u_right /* may be NULL */,
                            float min_x, float max_x, float min_y, float max_y,
                            int nlevels, const float* scale_factors);
void orc_frame_destroy(orc_frame*);
int orc_features_in_area(const orc_frame*, float x, float y, float r, int min_level,
                         int max_level, int* out, int cap);

/* OrbMatcher::SearchForInitialization (orb_matcher.cpp:264-382) */
int orc_search_for_initialization(const orc_frame* f1, const orc_frame* f2,
                                  float* prev_matched_xy /* n1*2, in/out */,
                                  int* matches12 /* n1 */, int window_size,
                                  float nnratio, int check_orientation);

/* OrbMatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th) (orb_matcher.cpp:13-103)
 * per map point: valid = track_is_in_view && !isBad(); has_obs = NumObservations()>0.
 * occupied[idx] (in) = F.GetMapPoint(idx) && its NumObservations()>0.
 * assigned[idx] (out) = index of the map point set on keypoint idx, else -1 (untouched). */
int orc_search_by_projection_mappoints(const orc_frame* f, int n_mp,
                                       const uint8_t* valid, const float* proj_x,
                                       const float* proj_y, const float* proj_xr,
                                       const int* pred_level, const float* view_cos,
                                       const uint8_t* mp_desc, const uint8_t* has_obs,
                                       const uint8_t* occupied, int th, float nnratio,
                                       int* assigned);

/* OrbMatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono)
 * (orb_matcher.cpp:1312-1453).  Per last-frame keypoint i: valid = has map point && !outlier.
 * (u, v, invzc) are the projection of :1346-1356, computed by the caller with the same
 * cv::Mat arithmetic the reference uses (it is OpenCV gemm, kept on the host side of the
 * boundary); the invzc<0 and image-bound rejections (:1353-1366) are applied here.
 * forward/backward are the two booleans of :1334-1335 (computed by the caller). */
int orc_search_by_projection_lastframe(const orc_frame* cur, int n_last,
                                       const uint8_t* valid, const float* u, const float* v,
                                       const float* invzc, const int* last_octave,
                                       const float* last_angle, const uint8_t* mp_desc,
                                       const uint8_t* has_obs, float bf, int forward,
                                       int backward, const uint8_t* occupied, float th,
                                       int check_orientation, int* assigned);

/* OrbMatcher::SearchByBoW(KeyFrame*, Frame&, ...) (orb_matcher.cpp:133-262); feature vectors flattened as in
 * include/orbfe.h orbfe_search_by_bow.  matched_kf[k] = KeyFrame feature index matched to Frame keypoint k, or -1. */
int orc_search_by_bow(const orc_frame* F, int n_kf, const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_valid,
                      int kf_nodes, const uint32_t* kf_ids, const int* kf_start, const uint32_t* kf_idx, int f_nodes,
                      const uint32_t* f_ids, const int* f_start, const uint32_t* f_idx, float nnratio, int check_ori,
                      int* matched_kf);

/* ---- the remaining OrbMatcher searches; argument conventions of include/orbfe.h ---- */
int orc_search_by_projection_sim3(const orc_frame* kf, int n_mp, const uint8_t* valid, const float* u, const float* v,
                                  const int* pred_level, const uint8_t* mp_desc, const uint8_t* matched_in, int th, int* matched);
int orc_search_by_projection_keyframe(const orc_frame* cur, int n_kf, const uint8_t* valid, const float* u, const float* v,
                                      const int* pred_level, const float* kf_angle, const uint8_t* mp_desc,
                                      const uint8_t* occupied, float th, int orb_dist, int check_ori, int* assigned);
/* both Fuse overloads: ur != NULL = Fuse(KF, vpMapPoints, th) (:804-954), ur == NULL = Fuse(KF, Scw, ...) (:956-1079) */
int orc_fuse(const orc_frame* kf, int n_mp, const uint8_t* valid, const float* u, const float* v, const float* ur,
             const int* pred_level, const uint8_t* mp_desc, float th, int* best_idx);
int orc_search_by_sim3(const orc_frame* kf1, const orc_frame* kf2, const uint8_t* valid1, const float* u1, const float* v1,
                       const int* lvl1, const uint8_t* desc1, const uint8_t* valid2, const float* u2, const float* v2,
                       const int* lvl2, const uint8_t* desc2, float th, int* match12);
int orc_search_by_bow_keyframes(const orc_frame* kf2, int n1, const uint8_t* desc1, const float* angle1, const uint8_t* valid1,
                                const uint8_t* valid2, int nodes1, const uint32_t* ids1, const int* start1, const uint32_t* idx1,
                                int nodes2, const uint32_t* ids2, const int* start2, const uint32_t* idx2, float nnratio,
                                int check_ori, int* matches12);
int orc_search_for_triangulation(const orc_frame* kf2, int n1, const orc_keypoint* kps1, const uint8_t* desc1,
                                 const uint8_t* valid1, const uint8_t* stereo1, const uint8_t* valid2, int nodes1,
                                 const uint32_t* ids1, const int* start1, const uint32_t* idx1, int nodes2,
                                 const uint32_t* ids2, const int* start2, const uint32_t* idx2, const float* F12, float ex,
                                 float ey, int only_stereo, int check_ori, int* matches12);

/* ---- N3: DBoW2 vocabulary transform (TemplatedVocabulary.h:1124-1250); arrays as include/orbfe.h ---- */
orc_vocabulary* orc_vocabulary_create(int k, int L, int scoring, int weighting, int n_nodes, const int* parent,
                                      const uint8_t* is_leaf, const uint8_t* desc, const double* weight);
void orc_vocabulary_destroy(orc_vocabulary*);
int orc_bow_transform(const orc_vocabulary* v, int n, const uint8_t* desc, int levelsup, unsigned* word_id, unsigned* node_id,
                      unsigned* bow_words, double* bow_values, int* n_bow, unsigned* fv_nodes, int* fv_start, unsigned* fv_idx,
                      int* n_fv);

/* ---- N2: Frame tail (frame.cpp:614-641, 277-337; map_point.cpp:382-396) ---- */
void orc_undistort_points(int n, const float* xy_in, float fx, float fy, float cx, float cy, const float* dist, int n_dist,
                          float* xy_out);
int orc_is_in_frustum(int n, const float* world, const float* normal, const float* min_dist, const float* max_dist,
                      const float* max_dist_raw,
                      const float* Rcw, const float* tcw, const float* Ow, float fx, float fy, float cx, float cy, float bf,
                      float min_x, float max_x, float min_y, float max_y, float log_scale_factor, int n_levels,
                      float viewing_cos_limit, uint8_t* in_view, float* proj_x, float* proj_y, float* proj_xr, int* level,
                      float* view_cos);

#ifdef __cplusplus
}
#endif
#endif
