/*
 * orbfe.h -- C ABI of the B200-native ORB front-end (liborbfe.so).
 *
 * Drop-in boundary for the frame front-end hot path of ThorsteinnJonsson/SLAM_framework.
 * Each entry point names the reference interface it replaces (file:line relative to the
 * reference tree).  The C++ shim classes with the reference's own signatures
 * (ORBextractor::Compute, Frame::ComputeStereoMatches, OrbMatcher::*) live in
 * include/orbfe_shim.hpp and are thin wrappers over these functions; INTEGRATION.md shows the
 * reference-side edit.
 *
 * Conventions: plain pointers and sizes, no C++/OpenCV/torch types; every function returns an
 * orbfe_status (0 = ok, <0 = error) and never throws; orbfe_last_error() gives a thread-local
 * message.  A handle is bound to (device, private stream); calls on ONE handle must be
 * serialised by the caller (an ORBextractor instance is not re-entrant either,
 * orb_extractor.h:25-93), calls on DIFFERENT handles may run concurrently from different
 * threads (frame.cpp:86-89 runs the left and right extractor on two std::threads).
 * There is no CPU fallback: without a CUDA device every compute call returns ORBFE_ERR_CUDA.
 */
#ifndef ORBFE_H_
#define ORBFE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  ORBFE_OK = 0,
  ORBFE_ERR_INVALID = -1,   /* bad argument / unsupported geometry */
  ORBFE_ERR_CUDA = -2,      /* CUDA runtime error (message in orbfe_last_error) */
  ORBFE_ERR_CAPACITY = -3,  /* caller buffer too small; *n_out holds the needed count */
  ORBFE_ERR_NOMEM = -4
} orbfe_status;

/* cv::KeyPoint memory layout (28 bytes): pt.x, pt.y, size, angle, response, octave, class_id */
typedef struct {
  float x, y, size, angle, response;
  int32_t octave, class_id;
} orbfe_keypoint;

/* ORBextractor ctor arguments (orb_extractor.cpp:351-355) + capacity of the device arena. */
typedef struct {
  int32_t nfeatures;
  float scale_factor;
  int32_t nlevels;
  int32_t ini_th_fast;
  int32_t min_th_fast;
  int32_t max_width;   /* largest image this handle will see */
  int32_t max_height;
  int32_t max_images;  /* image slots resident on the device (1 for the drop-in shim) */
} orbfe_params;

typedef struct orbfe_extractor orbfe_extractor;
typedef struct orbfe_frame orbfe_frame;

const char* orbfe_last_error(void);
const char* orbfe_version(void);
/* number of CUDA devices visible (0 => every compute call will fail with ORBFE_ERR_CUDA) */
int orbfe_device_count(void);

/* ---- ORBextractor (src/orb_features/orb_extractor.h:25-93) ------------------------------ */
/* replaces ORBextractor::ORBextractor (orb_extractor.cpp:351-411) */
int orbfe_extractor_create(const orbfe_params* p, int device, orbfe_extractor** out);
int orbfe_extractor_destroy(orbfe_extractor* ex);
/* replaces GetLevels/GetScaleFactors/GetInverseScaleFactors/GetScaleSigmaSquares/
 * GetInverseScaleSigmaSquares (orb_extractor.h:46-56); any output pointer may be NULL */
int orbfe_extractor_tables(const orbfe_extractor* ex, int* nlevels, float* scale, float* inv_scale,
                           float* sigma2, float* inv_sigma2, int32_t* features_per_level);
/* upper bound of keypoints one image can produce (sum over levels of max(N_l+3, 4*nIni)) */
int orbfe_extractor_max_keypoints(const orbfe_extractor* ex);

/* replaces ORBextractor::Compute(image, mask, keypoints, descriptors)
 * (orb_extractor.cpp:985-1049; mask is ignored there too, orb_extractor.h:40).
 * img: CV_8UC1, `stride` bytes per row, host memory.  kps/desc: caller-allocated, `capacity`
 * entries (desc is capacity x 32 bytes).  Synchronous.  Empty image (w<=0||h<=0||!img) => ok,
 * *n_out = 0 (orb_extractor.cpp:990-991). */
int orbfe_extract(orbfe_extractor* ex, const uint8_t* img, int w, int h, size_t stride,
                  orbfe_keypoint* kps, uint8_t* desc, int capacity, int* n_out);

/* Batched form of the same call over n_imgs same-sized images (image i -> slot i);
 * kps/desc hold n_imgs*capacity entries, n_out has n_imgs entries.  This is the offline
 * loop of examples/main_stereo.cpp:102-143 with the frames of one batch in flight at once. */
int orbfe_extract_batch(orbfe_extractor* ex, const uint8_t* const* imgs, int n_imgs, int w, int h,
                        size_t stride, orbfe_keypoint* kps, uint8_t* desc, int capacity, int* n_out);

/* replaces GetImagePyramid()[level] (orb_extractor.h:58): copies the ROI view (un-padded
 * level) of image slot `slot` to host memory.  dst may be NULL to query the size only. */
int orbfe_pyramid_level(orbfe_extractor* ex, int slot, int level, uint8_t* dst, size_t dst_stride,
                        int* w, int* h);

/* ---- device-resident batch path (bench / streaming; same kernels) ----------------------- */
/* async H2D of n_imgs images into slots [first_slot, first_slot+n_imgs) (pinned host memory
 * recommended). */
int orbfe_upload(orbfe_extractor* ex, int first_slot, const uint8_t* const* imgs, int n_imgs, int w,
                 int h, size_t stride);
/* same for 3- or 4-channel 8-bit colour frames: converts to gray on the device as cv::cvtColor
 * (CV_RGB2GRAY / CV_BGR2GRAY / CV_RGBA2GRAY / CV_BGRA2GRAY) does in Tracker::GrabImageStereo before the Frame
 * is built (src/core/tracker.cpp:110-127); rgb_order = the config's camera.rgb flag (1: R first). */
int orbfe_upload_color(orbfe_extractor* ex, int first_slot, const uint8_t* const* imgs, int n_imgs, int w, int h,
                       size_t stride, int channels, int rgb_order);
/* enqueue the full extractor on slots [0, n_imgs); results stay on the device */
int orbfe_run(orbfe_extractor* ex, int n_imgs);
/* enqueue Frame::ComputeStereoMatches for pairs p = (slot 2p, slot 2p+1), p < n_pairs */
int orbfe_run_stereo(orbfe_extractor* ex, int n_pairs, float bf, float baseline);
/* async D2H of the results of slots [0, n_imgs); kps/desc/n_out as in orbfe_extract_batch;
 * u_right/depth (n_imgs*capacity floats, may be NULL) are filled for even (left) slots. */
int orbfe_download(orbfe_extractor* ex, int n_imgs, orbfe_keypoint* kps, uint8_t* desc, int capacity,
                   int* n_out, float* u_right, float* depth);
/* same, without the staging copy and without waiting: capacity must equal
 * orbfe_extractor_max_keypoints() so that the host arrays mirror the device layout; the D2H copies
 * are enqueued straight into the caller's (ideally pinned) arrays, which are valid after orbfe_sync.
 * Lets a caller keep two handles in flight (upload of batch k+1 overlapping compute of batch k). */
int orbfe_download_async(orbfe_extractor* ex, int n_imgs, orbfe_keypoint* kps, uint8_t* desc, int capacity,
                         int* n_out, float* u_right, float* depth);
int orbfe_sync(orbfe_extractor* ex);
/* Page-locked host memory for frames (orbfe_upload) and results (orbfe_download_async), allocated PORTABLE (cudaHostAlloc with
 * cudaHostAllocPortable): pinned for every device of the process, so one process can feed several GPUs from it without
 * staging copies.  The reference reads its frames with cv::imread into pageable cv::Mat (examples/main_stereo.cpp:102-143);
 * decoding into a buffer from here removes the driver's staging copy.  orbfe_pinned_free(NULL) is a no-op. */
int orbfe_pinned_alloc(size_t bytes, void** out);
int orbfe_pinned_free(void* p);
/* CUDA-event timing on the handle's own stream: record event `slot` (0..63) now; elapsed ms
 * between two recorded slots (after orbfe_sync). */
int orbfe_event_record(orbfe_extractor* ex, int slot);
int orbfe_event_elapsed_ms(orbfe_extractor* ex, int slot_a, int slot_b, float* ms);
/* when enabled, every orbfe_run (+ the orbfe_run_stereo that follows it) brackets its stages
 * with CUDA events on the handle's stream, kept in a ring of the last 64 runs. */
int orbfe_set_stage_timing(orbfe_extractor* ex, int enabled);
/* sums the per-stage device time over the runs recorded since the last summary (at most 64) and
 * clears the ring.  ms_sum[7] = pyramid, FAST, quad-tree, blur, orientation+descriptor,
 * stereo search, stereo median.  Synchronises the handle's stream. */
#define ORBFE_NUM_STAGES 7
int orbfe_stage_summary(orbfe_extractor* ex, float* ms_sum, int* n_runs);
/* number of kernel launches issued by this handle since creation */
long long orbfe_launch_count(const orbfe_extractor* ex);
/* stage outputs of slot `slot` for per-stage parity tests: FAST candidates of a level
 * (x, y relative to the (16,16) detection origin, response) and the level's distributed
 * keypoints before orientation; returns count via *n_out */
int orbfe_debug_candidates(orbfe_extractor* ex, int slot, int level, orbfe_keypoint* out, int capacity, int* n_out);
int orbfe_debug_level_keypoints(orbfe_extractor* ex, int slot, int level, orbfe_keypoint* out, int capacity, int* n_out);
int orbfe_debug_blurred(orbfe_extractor* ex, int slot, int level, uint8_t* dst, size_t dst_stride, int* w, int* h);

/* ---- Frame::ComputeStereoMatches (src/data/frame.cpp:406-577; frame.h:95) ---------------- */
/* left/right: the two extractor handles whose last orbfe_extract produced the keypoints (their
 * device-resident pyramids are read, as the reference reads GetImagePyramid() at
 * frame.cpp:412,501,514,520).  kps/desc are the host arrays the Frame holds (keypoints_,
 * right_keypoints_, descriptors_, right_descriptors_).  baseline replaces the uninitialised
 * baseline_ read at frame.cpp:436 (pass bf/fx).  u_right/depth: n_left floats (stereo_coords_,
 * depths_; -1 = no match).  Both handles must be on the same device with identical params
 * and image size. */
int orbfe_stereo_match(orbfe_extractor* left, orbfe_extractor* right, int n_left,
                       const orbfe_keypoint* kps_left, const uint8_t* desc_left, int n_right,
                       const orbfe_keypoint* kps_right, const uint8_t* desc_right, float bf,
                       float baseline, float* u_right, float* depth, int* n_matched);

/* ---- OrbMatcher (src/orb_features/orb_matcher.h:14-119) ---------------------------------- */
/* static OrbMatcher::DescriptorDistance (orb_matcher.cpp:1630-1646), batched: d[i] =
 * hamming(a[i], b[i]) over n pairs of 32-byte rows (host memory). */
int orbfe_descriptor_distance(int device, const uint8_t* a, const uint8_t* b, int n, int32_t* d);

/* A Frame as the matchers see it: undistorted keypoints, descriptors, stereo right
 * coordinates (NULL = monocular), image bounds and scale table; builds the 64x48 feature
 * grid of Frame::AssignFeaturesToGrid / PosInGrid (frame.cpp:234-248, 339-346) on the GPU. */
int orbfe_frame_create(int device, int n, const orbfe_keypoint* kps_un, const uint8_t* desc,
                       const float* u_right, float min_x, float max_x, float min_y, float max_y,
                       int nlevels, const float* scale_factors, orbfe_frame** out);
int orbfe_frame_destroy(orbfe_frame* f);
/* Destroyed frame handles keep their stream, device arrays and pinned staging and are recycled by the next
 * orbfe_frame_create / orbfe_frame_from_extractor on the same device (Tracking builds one view per frame: frame.cpp:61-111 runs
 * per frame; a steady-state frame then allocates nothing).  At most 16 idle handles are kept; this call frees them and
 * returns how many it freed.  Live handles are not touched. */
int orbfe_frame_pool_trim(void);
/* The same handle built from the DEVICE-resident results of extractor slot `slot` (after orbfe_run / orbfe_extract; with
 * use_stereo the slot's stereo coordinates from orbfe_run_stereo / orbfe_stereo_match are taken as StereoCoordRight()):
 * keypoints and descriptors go device to device.  Only for cameras whose undistortion is the identity (Frame::
 * UndistortKeyPoints returns the keypoints unchanged when dist_coeff[0] == 0, frame.cpp:616-619). */
int orbfe_frame_from_extractor(orbfe_extractor* ex, int slot, int use_stereo, float min_x, float max_x, float min_y, float max_y,
                               orbfe_frame** out);
/* per-frame form: refresh a handle made by orbfe_frame_from_extractor with the next frame's results; its device arrays only grow,
 * so steady-state tracking allocates nothing */
int orbfe_frame_refresh_from_extractor(orbfe_frame* f, orbfe_extractor* ex, int slot, int use_stereo, float min_x, float max_x,
                                       float min_y, float max_y);
int orbfe_frame_num_keypoints(const orbfe_frame* f);
/* Frame::GetFeaturesInArea (frame.cpp:348-403): indices in the reference's order */
int orbfe_features_in_area(orbfe_frame* f, float x, float y, float r, int min_level, int max_level,
                           int32_t* out, int capacity, int* n_out);

/* OrbMatcher::SearchForInitialization (orb_matcher.cpp:264-382).  prev_matched: n1 (x,y)
 * pairs, in/out (vbPrevMatched); matches12: n1 ints (vnMatches12). */
int orbfe_search_for_initialization(orbfe_frame* f1, orbfe_frame* f2, float* prev_matched_xy,
                                    int32_t* matches12, int window_size, float nnratio,
                                    int check_orientation, int* n_matches);

/* OrbMatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th) (orb_matcher.cpp:13-111).
 * Per map point i: valid = track_is_in_view && !isBad(); proj_x/proj_y/proj_xr/pred_level/
 * view_cos = the track_* fields written by Frame::IsInFrustum (frame.cpp:328-334);
 * mp_desc = GetDescriptor() (32 bytes); has_obs = NumObservations()>0.
 * occupied[k] = keypoint k already holds a map point with observations (orb_matcher.cpp:59-63).
 * assigned[k] (out) = index of the map point F.SetMapPoint(k, .) received, else -1. */
int orbfe_search_by_projection_mappoints(orbfe_frame* f, int n_mp, const uint8_t* valid,
                                         const float* proj_x, const float* proj_y,
                                         const float* proj_xr, const int32_t* pred_level,
                                         const float* view_cos, const uint8_t* mp_desc,
                                         const uint8_t* has_obs, const uint8_t* occupied, int th,
                                         float nnratio, int32_t* assigned, int* n_matches);

/* OrbMatcher::SearchByProjection(Frame& Cur, const Frame& Last, th, bMono)
 * (orb_matcher.cpp:1312-1453).  Per last-frame keypoint i: valid = has map point && !outlier;
 * (u, v, invzc) = the projection of :1346-1356 (cv::Mat arithmetic, done by the shim);
 * forward/backward = the booleans of :1334-1335. */
int orbfe_search_by_projection_lastframe(orbfe_frame* cur, int n_last, const uint8_t* valid,
                                         const float* u, const float* v, const float* invzc,
                                         const int32_t* last_octave, const float* last_angle,
                                         const uint8_t* mp_desc, const uint8_t* has_obs, float bf,
                                         int forward, int backward, const uint8_t* occupied,
                                         float th, int check_orientation, int32_t* assigned,
                                         int* n_matches);

/* OrbMatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (orb_matcher.cpp:133-262).
 * f = the Frame (descriptors + keypoint angles).  KeyFrame side: n_kf descriptors / undistorted keypoint angles,
 * kf_valid[i] = GetMapPointMatches()[i] != NULL && !isBad().  The two DBoW2 FeatureVectors (std::map<NodeId,
 * vector<unsigned>>) are passed flattened: node ids ascending, start offsets (n_nodes+1) into the feature indices.
 * matched_kf_idx[k] (out, NumKeypoints entries) = index of the KeyFrame feature whose map point
 * vpMapPointMatches[k] receives, else -1. */
int orbfe_search_by_bow(orbfe_frame* f, int n_kf, const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_valid,
                        int kf_nnodes, const uint32_t* kf_node_ids, const int32_t* kf_node_start,
                        const uint32_t* kf_feat_idx, int f_nnodes, const uint32_t* f_node_ids,
                        const int32_t* f_node_start, const uint32_t* f_feat_idx, float nnratio,
                        int check_orientation, int32_t* matched_kf_idx, int* n_matches);

/* ---- the remaining OrbMatcher searches.  Convention as above: the cv::Mat geometry of each routine
 * (projection with Rcw/tcw or the Sim3, depth / distance-invariance / viewing-angle gates, MapPoint::PredictScale)
 * stays in the caller (include/orbfe_shim.hpp); valid[i] = "map point i passed every gate that precedes
 * GetFeaturesInArea", (u, v) its projection and pred_level its nPredictedLevel. ------------------------------- */

/* OrbMatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpMatched, th) (orb_matcher.cpp:384-497).
 * matched_in[k] = vpMatched[k] != NULL on entry; matched[k] (out) = index iMP of the point written to vpMatched[k]
 * by this call, else -1. */
int orbfe_search_by_projection_sim3(orbfe_frame* kf, int n_mp, const uint8_t* valid, const float* u, const float* v,
                                    const int32_t* pred_level, const uint8_t* mp_desc, const uint8_t* matched_in, int th,
                                    int32_t* matched, int* n_matches);

/* OrbMatcher::SearchByProjection(Frame& Cur, KeyFrame*, sAlreadyFound, th, ORBdist) (orb_matcher.cpp:1455-1582).
 * Per KeyFrame feature i: valid = map point && !isBad() && !sAlreadyFound.count() && distance-invariance gate
 * (:1505); the image-bound gate (:1490-1495) is applied here.  kf_angle = pKF->undistorted_keypoints[i].angle.
 * occupied[k] = CurrentFrame.GetMapPoint(k) != NULL.  assigned[k] (out) = i or -1. */
int orbfe_search_by_projection_keyframe(orbfe_frame* cur, int n_kf, const uint8_t* valid, const float* u, const float* v,
                                        const int32_t* pred_level, const float* kf_angle, const uint8_t* mp_desc,
                                        const uint8_t* occupied, float th, int orb_dist, int check_orientation,
                                        int32_t* assigned, int* n_matches);

/* OrbMatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) (orb_matcher.cpp:804-954), the search part: ur = u - bf*invz
 * (:849); best_idx[i] (out) = keypoint with bestDist <= TH_LOW for map point i, else -1.  The caller then walks i in
 * order and applies Replace / AddObservation / AddMapPoint (:933-949), re-checking isBad()/IsInKeyFrame (:828) first
 * (the graph edits never feed back into the search, which reads only keypoints and descriptors). */
int orbfe_fuse(orbfe_frame* kf, int n_mp, const uint8_t* valid, const float* u, const float* v, const float* ur,
               const int32_t* pred_level, const uint8_t* mp_desc, float th, int32_t* best_idx, int* n_fused);
/* OrbMatcher::Fuse(KeyFrame*, cv::Mat Scw, vpPoints, th, vpReplacePoint) (orb_matcher.cpp:956-1079), the search part */
int orbfe_fuse_sim3(orbfe_frame* kf, int n_mp, const uint8_t* valid, const float* u, const float* v,
                    const int32_t* pred_level, const uint8_t* mp_desc, float th, int32_t* best_idx, int* n_fused);

/* OrbMatcher::SearchBySim3 (orb_matcher.cpp:1081-1310).  Side 1 = the NumKeypoints(kf1) map-point slots of KF1
 * projected into KF2 (valid1 = map point && !vbAlreadyMatched1 && !isBad() && gates :1147-1167), side 2 likewise
 * into KF1.  match12[i1] (out) = idx2 when the two searches agree (:1291-1307), else -1. */
int orbfe_search_by_sim3(orbfe_frame* kf1, orbfe_frame* kf2, const uint8_t* valid1, const float* u1, const float* v1,
                         const int32_t* pred_level1, const uint8_t* mp_desc1, const uint8_t* valid2, const float* u2,
                         const float* v2, const int32_t* pred_level2, const uint8_t* mp_desc2, float th, int32_t* match12,
                         int* n_found);

/* OrbMatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&) (orb_matcher.cpp:499-632).  kf2 = the second
 * KeyFrame's handle; side 1 arrays as in orbfe_search_by_bow; valid1/valid2 = map point && !isBad().
 * matches12[i1] (out, n1 entries) = idx2 whose map point vpMatches12[i1] receives, else -1. */
int orbfe_search_by_bow_keyframes(orbfe_frame* kf2, int n1, const uint8_t* desc1, const float* angle1, const uint8_t* valid1,
                                  const uint8_t* valid2, int nnodes1, const uint32_t* node_ids1, const int32_t* node_start1,
                                  const uint32_t* feat_idx1, int nnodes2, const uint32_t* node_ids2, const int32_t* node_start2,
                                  const uint32_t* feat_idx2, float nnratio, int check_orientation, int32_t* matches12,
                                  int* n_matches);

/* OrbMatcher::SearchForTriangulation (orb_matcher.cpp:634-802).  valid1/valid2 = the feature has NO map point;
 * stereo1[i] = pKF1->right_coords[i] >= 0 (side 2 reads the handle's u_right); F12 = 9 floats row-major; (ex, ey) =
 * the epipole of :643-649.  matches12[i1] (out) = idx2 or -1; vMatchedPairs = the pairs (i1, matches12[i1] >= 0) in
 * ascending i1 (:794-799). */
int orbfe_search_for_triangulation(orbfe_frame* kf2, int n1, const orbfe_keypoint* kps1_un, const uint8_t* desc1,
                                   const uint8_t* valid1, const uint8_t* stereo1, const uint8_t* valid2, int nnodes1,
                                   const uint32_t* node_ids1, const int32_t* node_start1, const uint32_t* feat_idx1,
                                   int nnodes2, const uint32_t* node_ids2, const int32_t* node_start2,
                                   const uint32_t* feat_idx2, const float* F12, float ex, float ey, int only_stereo,
                                   int check_orientation, int32_t* matches12, int* n_matches);

/* ---- ORB vocabulary / bag of words: Frame::ComputeBoW (src/data/frame.cpp:258-263), KeyFrame::ComputeBoW
 * (src/data/keyframe.cpp:127-137) -> DBoW2 TemplatedVocabulary<FORB>::transform(features, BowVector&, FeatureVector&,
 * levelsup) (third_party/DBoW2/DBoW2/TemplatedVocabulary.h:1124-1250, FORB.cpp:81-101). -------------------------- */
typedef struct orbfe_vocabulary orbfe_vocabulary;

/* The tree as TemplatedVocabulary::loadFromTextFile builds it (TemplatedVocabulary.h:1335-1422): header k L scoring
 * weighting (ScoringType / WeightingType of BowVector.h:36-53); n_nodes entries, node 0 = root (its array entries are
 * ignored); node i >= 1: parent[i] < i, is_leaf[i] (takes the next word id), 32 descriptor bytes, weight. */
int orbfe_vocabulary_create(int device, int k, int L, int scoring, int weighting, int n_nodes, const int32_t* parent,
                            const uint8_t* is_leaf, const uint8_t* desc, const double* weight, orbfe_vocabulary** out);
/* the ORBvoc.txt format read by system.cpp / loadFromTextFile */
int orbfe_vocabulary_load_text(const char* path, int device, orbfe_vocabulary** out);
int orbfe_vocabulary_destroy(orbfe_vocabulary* v);
int orbfe_vocabulary_info(const orbfe_vocabulary* v, int* k, int* L, int* scoring, int* weighting, int* n_nodes, int* n_words);

/* transform(): desc = n x 32 descriptor rows (host).  Per feature (optional, may be NULL): word_id[n], node_id[n] (the node
 * `levelsup` levels above the leaves).  BowVector = bow_words[*n_bow] ascending with bow_values (after the weighting and
 * normalisation the vocabulary's scoring asks for).  FeatureVector = fv_nodes[*n_fv] ascending, fv_start[*n_fv + 1]
 * offsets into fv_idx (feature indices ascending inside a node) -- the flattened form orbfe_search_by_bow* take.
 * Output capacities: n entries each (fv_start: n + 1).  Features whose word weight is 0 (stopped words) are skipped. */
int orbfe_bow_transform(orbfe_vocabulary* v, int n, const uint8_t* desc, int levelsup, uint32_t* word_id, uint32_t* node_id,
                        uint32_t* bow_words, double* bow_values, int* n_bow, uint32_t* fv_nodes, int32_t* fv_start,
                        uint32_t* fv_idx, int* n_fv);

/* ---- the Frame tail ------------------------------------------------------------------------------- */
/* Frame::UndistortKeyPoints (src/data/frame.cpp:614-641): cv::undistortPoints(mat, mat, K, dist, cv::Mat(), K) on the
 * keypoint coordinates, every other cv::KeyPoint field copied.  (fx, fy, cx, cy) = calib_mat_ (CV_32F); dist_coeffs =
 * dist_coeff_ (k1 k2 p1 p2 [k3 ...], up to 12 used; a tilted-sensor model is rejected).  dist_coeffs[0] == 0 copies the
 * keypoints unchanged, as the reference does (:616-619). */
int orbfe_undistort_keypoints(int device, int n, const orbfe_keypoint* kps, float fx, float fy, float cx, float cy,
                              const float* dist_coeffs, int n_dist, orbfe_keypoint* kps_un);

/* Frame::IsInFrustum (src/data/frame.cpp:277-337) + MapPoint::PredictScale (src/data/map_point.cpp:382-396) for n map
 * points in one launch = the loop of Tracker::SearchLocalPoints (src/core/tracker.cpp:1196-1211).  world_pos / normal:
 * n x 3 floats (GetWorldPos / GetNormal); min_dist / max_dist = Get{Min,Max}DistanceInvariance() (the range gate, :306-314);
 * max_dist_raw = MapPoint::max_dist_ itself, which PredictScale divides by the distance (map_point.cpp:386; the class has
 * no getter for it: INTEGRATION.md section 4); Rcw (row-major 3x3),
 * tcw, Ow = the frame pose (frame.cpp:270-275).  Outputs = the track_* fields (:328-334), zero where in_view[i] == 0;
 * they are the arrays orbfe_search_by_projection_mappoints takes.  AssignFeaturesToGrid (frame.cpp:234-248), the third
 * piece of this row, is the grid orbfe_frame_create builds on the device. */
int orbfe_is_in_frustum(int device, int n, const float* world_pos, const float* normal, const float* min_dist,
                        const float* max_dist, const float* max_dist_raw, const float* Rcw, const float* tcw, const float* Ow,
                        float fx, float fy, float cx, float cy, float bf, float min_x, float max_x, float min_y, float max_y,
                        float log_scale_factor, int n_levels, float viewing_cos_limit, uint8_t* in_view, float* proj_x,
                        float* proj_y, float* proj_xr, int32_t* scale_level, float* view_cos, int* n_in_view);
/* Tracker::SearchLocalPoints (src/core/tracker.cpp:1196-1226) in one call: IsInFrustum over the n candidate local map points
 * (arguments as orbfe_is_in_frustum; the image bounds, level count and scale factors are the frame handle's) chained on the
 * device into OrbMatcher::SearchByProjection(Frame&, vpMapPoints, th) (arguments as orbfe_search_by_projection_mappoints).
 * in_view[n] / scale_level[n] (optional outputs) = track_is_in_view / track_scale_level; assigned[k] = map point given to
 * keypoint k, else -1. */
int orbfe_search_local_points(orbfe_frame* f, int n, const float* world_pos, const float* normal, const float* min_dist,
                              const float* max_dist, const float* max_dist_raw, const float* Rcw, const float* tcw, const float* Ow,
                              float fx, float fy, float cx, float cy, float bf, float log_scale_factor, float viewing_cos_limit,
                              const uint8_t* mp_desc, const uint8_t* has_obs, const uint8_t* occupied, int th, float nnratio,
                              uint8_t* in_view, int32_t* scale_level, int32_t* assigned, int* n_in_view, int* n_matches);

/* parity tap: std::log(float) as PredictScale evaluates it (glibc logf restated on the device), y[i] = logf(x[i]) */
int orbfe_debug_logf(int device, int n, const float* x, float* y);

#ifdef __cplusplus
}
#endif
#endif /* ORBFE_H_ */
