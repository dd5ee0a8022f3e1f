// orbfe_match.cu -- host side of the OrbMatcher / Frame-grid entry points of include/orbfe.h
// (kernels in k_match.cuh).  A frame handle owns the device copy of one Frame's matcher view
// (undistorted keypoints, descriptors, stereo coordinates, 64x48 grid) plus query scratch and a
// private stream; calls on one handle are serialised by the caller.
#include "../../include/orbfe.h"

#include "k_match.cuh"
#include "orbfe_host.h"

#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#define CUDA_TRY(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return orbfe_fail(ORBFE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#ifdef ORBFE_EMU
#define MATCH_LAUNCH(f, kernel, grid, block, smem, ...) emu::launch(grid, block, smem, [&]() { kernel(__VA_ARGS__); })
#else
#define MATCH_LAUNCH(f, kernel, grid, block, smem, ...) kernel<<<grid, block, smem, (f)->stream>>>(__VA_ARGS__)
#endif

struct orbfe_frame {
  int device = 0;
  cudaStream_t stream = nullptr;
  int n = 0, nlevels = 0;
  float minX = 0, maxX = 0, minY = 0, maxY = 0, gw = 1, gh = 1;
  std::vector<float> scale;
  std::vector<MatchKp> hkp;  // host copy (prevMatched update, validation)
  MatchKp* d_kp = nullptr;
  uint8_t* d_desc = nullptr;
  float* d_uR = nullptr;
  int* d_cellStart = nullptr;
  int* d_cellItems = nullptr;
  // query scratch (grown on demand)
  int qCap = 0;
  float *d_qx = nullptr, *d_qy = nullptr, *d_qr = nullptr, *d_qxr = nullptr, *d_qAngle = nullptr;
  int *d_qMinL = nullptr, *d_qMaxL = nullptr, *d_qOff = nullptr, *d_qCnt = nullptr, *d_evBin = nullptr, *d_evIdx = nullptr;
  uint8_t *d_qValid = nullptr, *d_qDesc = nullptr, *d_qHasObs = nullptr;
  uint8_t* d_occ = nullptr;
  int* d_out = nullptr;
  int outCap = 0;
  uint2* d_cand = nullptr;
  int candCap = 0;
  int* d_jbest = nullptr;   // Jacobi resolve state (k_match_iterate)
  int* d_jown = nullptr;
  int* d_jchanged = nullptr;
  int jCap = 0;
  int* d_cursor = nullptr;  // [0] cursor [1] overflow [2] nmatches
  int* h_res = nullptr;     // pinned, 4 ints
  FrameGrid grid() const {
    FrameGrid G;
    G.kp = d_kp; G.desc = d_desc; G.uR = d_uR; G.cellStart = d_cellStart; G.cellItems = d_cellItems; G.n = n;
    G.minX = minX; G.minY = minY; G.gw = gw; G.gh = gh;
    return G;
  }
};

template <class T>
static cudaError_t regrow(T** p, size_t count) {
  if (*p) cudaFree(*p);
  *p = nullptr;
  return cudaMalloc(p, std::max<size_t>(count, 1) * sizeof(T));
}

static int ensure_queries(orbfe_frame* f, int nq) {
  if (nq <= f->qCap) return ORBFE_OK;
  const size_t c = (size_t)nq + 256;
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  CUDA_TRY(regrow(&f->d_qx, c)); CUDA_TRY(regrow(&f->d_qy, c)); CUDA_TRY(regrow(&f->d_qr, c)); CUDA_TRY(regrow(&f->d_qxr, c));
  CUDA_TRY(regrow(&f->d_qAngle, c)); CUDA_TRY(regrow(&f->d_qMinL, c)); CUDA_TRY(regrow(&f->d_qMaxL, c));
  CUDA_TRY(regrow(&f->d_qOff, c)); CUDA_TRY(regrow(&f->d_qCnt, c)); CUDA_TRY(regrow(&f->d_evBin, c)); CUDA_TRY(regrow(&f->d_evIdx, c));
  CUDA_TRY(regrow(&f->d_qValid, c)); CUDA_TRY(regrow(&f->d_qDesc, c * 32)); CUDA_TRY(regrow(&f->d_qHasObs, c));
  f->qCap = (int)c;
  return ORBFE_OK;
}
static int ensure_out(orbfe_frame* f, int n) {
  if (n <= f->outCap) return ORBFE_OK;
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  CUDA_TRY(regrow(&f->d_out, (size_t)n + 256));
  f->outCap = n + 256;
  return ORBFE_OK;
}
static int ensure_cand(orbfe_frame* f, int n) {
  if (n <= f->candCap) return ORBFE_OK;
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  CUDA_TRY(regrow(&f->d_cand, (size_t)n));
  f->candCap = n;
  return ORBFE_OK;
}

struct HostQueries {
  std::vector<float> x, y, r, xr, angle;
  std::vector<int> minL, maxL;
  std::vector<uint8_t> valid;
  const uint8_t* desc = nullptr;    // nq x 32 (host)
  const uint8_t* hasObs = nullptr;  // nq (host) or null
  int n = 0;
  bool checkUR = false;
  void resize(int nq) {
    n = nq; x.assign(nq, 0.f); y.assign(nq, 0.f); r.assign(nq, 0.f); xr.assign(nq, 0.f); angle.assign(nq, 0.f);
    minL.assign(nq, -1); maxL.assign(nq, -1); valid.assign(nq, 0);
  }
};

// uploads the queries, runs phase A + phase B on `f` (the searched frame); out_n entries of d_out come
// back in `out`; *nmatches gets the count
static int run_search(orbfe_frame* f, const HostQueries& Q, int mode, float nnratio, int checkOri, const uint8_t* occupied,
                      int32_t* out, int out_n, int* nmatches) {
  CUDA_TRY(cudaSetDevice(f->device));
  int rc;
  const int nq = Q.n;
  if ((rc = ensure_queries(f, nq))) return rc;
  if ((rc = ensure_out(f, out_n))) return rc;
  if (f->candCap == 0 && (rc = ensure_cand(f, std::max(nq * 48 + 4096, 1 << 16)))) return rc;
  const size_t stateInts = mode == ORBFE_MODE_INIT ? 2 * (size_t)f->n : (size_t)f->n;
  const size_t smem = std::max<size_t>(stateInts * sizeof(int), 16);
  if (mode == ORBFE_MODE_INIT && smem > 200 * 1024) return orbfe_fail(ORBFE_ERR_INVALID, "frame has too many keypoints (%d) for the resolve kernel", f->n);
  cudaStream_t st = f->stream;
  const size_t q4 = (size_t)nq * sizeof(float);
  CUDA_TRY(cudaMemcpyAsync(f->d_qx, Q.x.data(), q4, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(f->d_qy, Q.y.data(), q4, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(f->d_qr, Q.r.data(), q4, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(f->d_qxr, Q.xr.data(), q4, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(f->d_qAngle, Q.angle.data(), q4, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(f->d_qMinL, Q.minL.data(), q4, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(f->d_qMaxL, Q.maxL.data(), q4, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(f->d_qValid, Q.valid.data(), (size_t)nq, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(f->d_qDesc, Q.desc, (size_t)nq * 32, cudaMemcpyHostToDevice, st));
  if (Q.hasObs) CUDA_TRY(cudaMemcpyAsync(f->d_qHasObs, Q.hasObs, (size_t)nq, cudaMemcpyHostToDevice, st));
  if (occupied) CUDA_TRY(cudaMemcpyAsync(f->d_occ, occupied, (size_t)f->n, cudaMemcpyHostToDevice, st));
#ifndef ORBFE_EMU
  CUDA_TRY(cudaFuncSetAttribute(k_match_resolve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
#endif
  const bool jacobi = mode != ORBFE_MODE_INIT;  // SearchForInitialization keeps the serial one-warp resolve
  if (jacobi && nq > f->jCap) {
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(regrow(&f->d_jbest, (size_t)nq + 256));
    CUDA_TRY(regrow(&f->d_jchanged, (size_t)nq + 256 + 16));
    f->jCap = nq + 256;
  }
  if (jacobi && !f->d_jown) CUDA_TRY(regrow(&f->d_jown, 3 * (size_t)std::max(f->n, 1)));
  for (int attempt = 0; attempt < 8; ++attempt) {
    CUDA_TRY(cudaMemsetAsync(f->d_cursor, 0, 4 * sizeof(int), st));
    MatchQueries MQ;
    MQ.x = f->d_qx; MQ.y = f->d_qy; MQ.r = f->d_qr; MQ.xr = f->d_qxr; MQ.minLevel = f->d_qMinL; MQ.maxLevel = f->d_qMaxL;
    MQ.valid = f->d_qValid; MQ.desc = f->d_qDesc; MQ.n = nq; MQ.checkUR = Q.checkUR ? 1 : 0;
    MatchScratch S;
    S.cand = f->d_cand; S.qOff = f->d_qOff; S.qCnt = f->d_qCnt; S.cursor = f->d_cursor; S.capacity = f->candCap;
    ResolveArgs A;
    A.mode = mode; A.nQ = nq; A.nKp = f->n; A.nnratio = nnratio; A.checkOri = checkOri; A.hasObs = f->d_qHasObs;
    A.occupiedIn = f->d_occ; A.qAngle = f->d_qAngle; A.kp = f->d_kp; A.out = f->d_out; A.evBin = f->d_evBin;
    A.evIdx = f->d_evIdx; A.result = f->d_cursor + 2;
    if (nq > 0)
      MATCH_LAUNCH(f, k_match_candidates, dim3((nq + ORBFE_MATCH_THREADS / 32 - 1) / (ORBFE_MATCH_THREADS / 32)),
                   dim3(ORBFE_MATCH_THREADS), 0, f->grid(), MQ, S);
    if (!jacobi) {
      MATCH_LAUNCH(f, k_match_resolve, dim3(1), dim3(32), smem, A, S);
    } else {
      JacobiState J;
      J.best = f->d_jbest; J.own = f->d_jown; J.changed = f->d_jchanged;
      CUDA_TRY(cudaMemsetAsync(f->d_jown, 0x7f, 3 * (size_t)std::max(f->n, 1) * sizeof(int), st));
      CUDA_TRY(cudaMemsetAsync(f->d_jchanged, 0, ((size_t)nq + 16) * sizeof(int), st));
      const int grid = std::max(1, (nq + ORBFE_MATCH_THREADS / 32 - 1) / (ORBFE_MATCH_THREADS / 32));
      const int chunk = 8;
      int t = 0;
      for (;;) {  // iterations are launched in chunks; converged iterations return immediately
        for (int k = 0; k < chunk && t <= nq; ++k, ++t) MATCH_LAUNCH(f, k_match_iterate, dim3(grid), dim3(ORBFE_MATCH_THREADS), 0, A, S, J, t);
        CUDA_TRY(cudaMemcpyAsync(f->h_res + 3, f->d_jchanged + (t - 1), sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (f->h_res[3] == 0 || t > nq) break;  // a fixed point is reached after at most nq iterations
      }
      MATCH_LAUNCH(f, k_match_finalize, dim3(1), dim3(1024), 0, A, S, J);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(f->h_res, f->d_cursor, 3 * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (!f->h_res[1]) {
      if (out_n > 0) CUDA_TRY(cudaMemcpy(out, f->d_out, (size_t)out_n * sizeof(int), cudaMemcpyDeviceToHost));
      if (nmatches) *nmatches = f->h_res[2];
      return ORBFE_OK;
    }
    // candidate buffer too small: h_res[0] is the total that was requested
    if ((rc = ensure_cand(f, f->h_res[0] + 4096))) return rc;
  }
  return orbfe_fail(ORBFE_ERR_CUDA, "candidate buffer did not converge");
}

extern "C" {

int orbfe_descriptor_distance(int device, const uint8_t* a, const uint8_t* b, int n, int32_t* d) {
  if (n < 0 || (n && (!a || !b || !d))) return orbfe_fail(ORBFE_ERR_INVALID, "bad arguments");
  if (n == 0) return ORBFE_OK;
  CUDA_TRY(cudaSetDevice(device));
  uint8_t *da = nullptr, *db = nullptr;
  int* dd = nullptr;
  CUDA_TRY(cudaMalloc(&da, (size_t)n * 32));
  cudaError_t e = cudaMalloc(&db, (size_t)n * 32);
  if (e == cudaSuccess) e = cudaMalloc(&dd, (size_t)n * sizeof(int));
  if (e == cudaSuccess) e = cudaMemcpy(da, a, (size_t)n * 32, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(db, b, (size_t)n * 32, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
#ifdef ORBFE_EMU
    emu::launch(dim3((n + 255) / 256), dim3(256), 0, [&]() { k_descriptor_distance(da, db, n, dd); });
#else
    k_descriptor_distance<<<(n + 255) / 256, 256>>>(da, db, n, dd);
    e = cudaGetLastError();
#endif
  }
  if (e == cudaSuccess) e = cudaMemcpy(d, dd, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(da); cudaFree(db); cudaFree(dd);
  if (e != cudaSuccess) return orbfe_fail(ORBFE_ERR_CUDA, "descriptor distance failed: %s", cudaGetErrorString(e));
  return ORBFE_OK;
}

int orbfe_frame_destroy(orbfe_frame* f) {
  if (!f) return ORBFE_OK;
  cudaSetDevice(f->device);
  if (f->stream) cudaStreamSynchronize(f->stream);
  cudaFree(f->d_kp); cudaFree(f->d_desc); cudaFree(f->d_uR); cudaFree(f->d_cellStart); cudaFree(f->d_cellItems);
  cudaFree(f->d_qx); cudaFree(f->d_qy); cudaFree(f->d_qr); cudaFree(f->d_qxr); cudaFree(f->d_qAngle); cudaFree(f->d_qMinL);
  cudaFree(f->d_qMaxL); cudaFree(f->d_qOff); cudaFree(f->d_qCnt); cudaFree(f->d_evBin); cudaFree(f->d_evIdx);
  cudaFree(f->d_qValid); cudaFree(f->d_qDesc); cudaFree(f->d_qHasObs); cudaFree(f->d_occ); cudaFree(f->d_out);
  cudaFree(f->d_cand); cudaFree(f->d_cursor); cudaFree(f->d_jbest); cudaFree(f->d_jown); cudaFree(f->d_jchanged);
  cudaFreeHost(f->h_res);
  if (f->stream) cudaStreamDestroy(f->stream);
  delete f;
  return ORBFE_OK;
}

int orbfe_frame_create(int device, int n, const orbfe_keypoint* kps_un, const uint8_t* desc, const float* u_right,
                       float min_x, float max_x, float min_y, float max_y, int nlevels, const float* scale_factors,
                       orbfe_frame** out) {
  if (!out) return orbfe_fail(ORBFE_ERR_INVALID, "null argument");
  *out = nullptr;
  if (n < 0 || (n && (!kps_un || !desc)) || nlevels < 1 || !scale_factors || !(max_x > min_x) || !(max_y > min_y))
    return orbfe_fail(ORBFE_ERR_INVALID, "bad frame arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); ndev = 0; }
  if (device < 0 || device >= ndev)
    return orbfe_fail(ORBFE_ERR_CUDA, "CUDA device %d not available (%d visible); this library has no CPU path", device, ndev);
  orbfe_frame* f = new (std::nothrow) orbfe_frame();
  if (!f) return orbfe_fail(ORBFE_ERR_NOMEM, "out of host memory");
  f->device = device; f->n = n; f->nlevels = nlevels;
  f->minX = min_x; f->maxX = max_x; f->minY = min_y; f->maxY = max_y;
  f->gw = static_cast<float>(max_x - min_x) / ORBFE_GRID_COLS;  // frame.cpp:223-224
  f->gh = static_cast<float>(max_y - min_y) / ORBFE_GRID_ROWS;
  f->scale.assign(scale_factors, scale_factors + nlevels);
  f->hkp.resize(n);
  std::vector<float> ur(n, -1.0f);
  for (int i = 0; i < n; ++i) {
    f->hkp[i] = MatchKp{kps_un[i].x, kps_un[i].y, kps_un[i].angle, kps_un[i].octave};
    if (kps_un[i].octave < 0 || kps_un[i].octave >= nlevels) {
      delete f;
      return orbfe_fail(ORBFE_ERR_INVALID, "keypoint %d has octave %d outside [0,%d)", i, kps_un[i].octave, nlevels);
    }
    if (u_right) ur[i] = u_right[i];
  }
  auto fail = [&](cudaError_t e) {
    orbfe_frame_destroy(f);
    return orbfe_fail(ORBFE_ERR_CUDA, "frame setup failed: %s", cudaGetErrorString(e));
  };
  cudaError_t e = cudaSetDevice(device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
  const size_t n1 = std::max(n, 1);
  if (e == cudaSuccess) e = cudaMalloc(&f->d_kp, n1 * sizeof(MatchKp));
  if (e == cudaSuccess) e = cudaMalloc(&f->d_desc, n1 * 32);
  if (e == cudaSuccess) e = cudaMalloc(&f->d_uR, n1 * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&f->d_cellStart, (ORBFE_GRID_CELLS + 1) * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&f->d_cellItems, n1 * sizeof(int));
  if (e == cudaSuccess) e = cudaMalloc(&f->d_occ, n1);
  if (e == cudaSuccess) e = cudaMalloc(&f->d_cursor, 4 * sizeof(int));
  if (e == cudaSuccess) e = cudaMallocHost(&f->h_res, 4 * sizeof(int));
  if (e == cudaSuccess && n) e = cudaMemcpyAsync(f->d_kp, f->hkp.data(), (size_t)n * sizeof(MatchKp), cudaMemcpyHostToDevice, f->stream);
  if (e == cudaSuccess && n) e = cudaMemcpyAsync(f->d_desc, desc, (size_t)n * 32, cudaMemcpyHostToDevice, f->stream);
  if (e == cudaSuccess && n) e = cudaMemcpyAsync(f->d_uR, ur.data(), (size_t)n * sizeof(float), cudaMemcpyHostToDevice, f->stream);
  if (e == cudaSuccess) e = cudaMemsetAsync(f->d_occ, 0, n1, f->stream);
  if (e != cudaSuccess) return fail(e);
  MATCH_LAUNCH(f, k_grid_build, dim3(1), dim3(1024), 0, f->d_kp, n, f->minX, f->minY, f->gw, f->gh, f->d_cellStart, f->d_cellItems);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(f->stream);  // `ur` goes out of scope
  if (e != cudaSuccess) return fail(e);
  *out = f;
  return ORBFE_OK;
}

int orbfe_features_in_area(orbfe_frame* f, float x, float y, float r, int min_level, int max_level, int32_t* out,
                           int capacity, int* n_out) {
  if (!f || !n_out || capacity < 0 || (capacity && !out)) return orbfe_fail(ORBFE_ERR_INVALID, "bad arguments");
  CUDA_TRY(cudaSetDevice(f->device));
  int rc;
  if ((rc = ensure_out(f, capacity + 1))) return rc;
  MATCH_LAUNCH(f, k_features_in_area, dim3(1), dim3(32), 0, f->grid(), x, y, r, min_level, max_level, f->d_out, capacity,
               f->d_cursor + 3);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(f->h_res, f->d_cursor, 4 * sizeof(int), cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  const int n = f->h_res[3];
  *n_out = n;
  if (std::min(n, capacity) > 0) CUDA_TRY(cudaMemcpy(out, f->d_out, (size_t)std::min(n, capacity) * sizeof(int), cudaMemcpyDeviceToHost));
  if (n > capacity) return orbfe_fail(ORBFE_ERR_CAPACITY, "capacity %d too small for %d indices", capacity, n);
  return ORBFE_OK;
}

int orbfe_search_for_initialization(orbfe_frame* f1, orbfe_frame* f2, float* prev_matched_xy, int32_t* matches12,
                                    int window_size, float nnratio, int check_orientation, int* n_matches) {
  if (!f1 || !f2 || !prev_matched_xy || !matches12) return orbfe_fail(ORBFE_ERR_INVALID, "null argument");
  if (f1->device != f2->device) return orbfe_fail(ORBFE_ERR_INVALID, "frames live on different devices");
  if (n_matches) *n_matches = 0;
  const int n1 = f1->n;
  if (n1 == 0) return ORBFE_OK;
  // queries = F1 keypoints of octave 0, window centred on vbPrevMatched (orb_matcher.cpp:283-297)
  HostQueries Q;
  Q.resize(n1);
  std::vector<uint8_t> desc1((size_t)n1 * 32);
  CUDA_TRY(cudaSetDevice(f1->device));
  CUDA_TRY(cudaStreamSynchronize(f1->stream));
  CUDA_TRY(cudaMemcpy(desc1.data(), f1->d_desc, (size_t)n1 * 32, cudaMemcpyDeviceToHost));
  for (int i = 0; i < n1; ++i) {
    const int level1 = f1->hkp[i].octave;
    Q.valid[i] = level1 > 0 ? 0 : 1;
    Q.x[i] = prev_matched_xy[2 * i]; Q.y[i] = prev_matched_xy[2 * i + 1];
    Q.r[i] = (float)window_size;
    Q.minL[i] = level1; Q.maxL[i] = level1;
    Q.angle[i] = f1->hkp[i].angle;
  }
  Q.desc = desc1.data();
  int nm = 0;
  const int rc = run_search(f2, Q, ORBFE_MODE_INIT, nnratio, check_orientation, nullptr, matches12, n1, &nm);
  if (rc) return rc;
  for (int i = 0; i < n1; ++i)  // :377-379
    if (matches12[i] >= 0) {
      prev_matched_xy[2 * i] = f2->hkp[matches12[i]].x;
      prev_matched_xy[2 * i + 1] = f2->hkp[matches12[i]].y;
    }
  if (n_matches) *n_matches = nm;
  return ORBFE_OK;
}

int orbfe_search_by_projection_mappoints(orbfe_frame* f, int n_mp, const uint8_t* valid, const float* proj_x,
                                         const float* proj_y, const float* proj_xr, const int32_t* pred_level,
                                         const float* view_cos, const uint8_t* mp_desc, const uint8_t* has_obs,
                                         const uint8_t* occupied, int th, float nnratio, int32_t* assigned, int* n_matches) {
  if (!f || !assigned || n_mp < 0) return orbfe_fail(ORBFE_ERR_INVALID, "bad arguments");
  if (n_mp && (!valid || !proj_x || !proj_y || !proj_xr || !pred_level || !view_cos || !mp_desc || !has_obs))
    return orbfe_fail(ORBFE_ERR_INVALID, "null map-point array");
  if (f->n && !occupied) return orbfe_fail(ORBFE_ERR_INVALID, "null occupied array");
  if (n_matches) *n_matches = 0;
  HostQueries Q;
  Q.resize(n_mp);
  const bool bFactor = th != 1;
  for (int i = 0; i < n_mp; ++i) {
    if (!valid[i]) continue;
    const int lvl = pred_level[i];
    if (lvl < 0 || lvl >= f->nlevels) return orbfe_fail(ORBFE_ERR_INVALID, "map point %d: predicted level %d out of range", i, lvl);
    float r = (view_cos[i] > 0.998) ? 2.5f : 4.0f;  // RadiusByViewingCos (orb_matcher.cpp:105-111)
    if (bFactor) r *= th;
    Q.valid[i] = 1;
    Q.x[i] = proj_x[i]; Q.y[i] = proj_y[i];
    Q.r[i] = r * f->scale[lvl];
    Q.xr[i] = proj_xr[i];
    Q.minL[i] = lvl - 1; Q.maxL[i] = lvl;
  }
  Q.desc = mp_desc; Q.hasObs = has_obs; Q.checkUR = true;
  return run_search(f, Q, ORBFE_MODE_MAPPOINTS, nnratio, 0, occupied, assigned, f->n, n_matches);
}

int orbfe_search_by_projection_lastframe(orbfe_frame* cur, int n_last, const uint8_t* valid, const float* u, const float* v,
                                         const float* invzc, const int32_t* last_octave, const float* last_angle,
                                         const uint8_t* mp_desc, const uint8_t* has_obs, float bf, int forward,
                                         int backward, const uint8_t* occupied, float th, int check_orientation,
                                         int32_t* assigned, int* n_matches) {
  if (!cur || !assigned || n_last < 0) return orbfe_fail(ORBFE_ERR_INVALID, "bad arguments");
  if (n_last && (!valid || !u || !v || !invzc || !last_octave || !last_angle || !mp_desc || !has_obs))
    return orbfe_fail(ORBFE_ERR_INVALID, "null last-frame array");
  if (cur->n && !occupied) return orbfe_fail(ORBFE_ERR_INVALID, "null occupied array");
  if (n_matches) *n_matches = 0;
  HostQueries Q;
  Q.resize(n_last);
  for (int i = 0; i < n_last; ++i) {
    if (!valid[i]) continue;
    if (invzc[i] < 0) continue;                                   // :1353-1354
    if (u[i] < cur->minX || u[i] > cur->maxX) continue;           // :1359-1366
    if (v[i] < cur->minY || v[i] > cur->maxY) continue;
    const int oct = last_octave[i];
    if (oct < 0 || oct >= cur->nlevels) return orbfe_fail(ORBFE_ERR_INVALID, "last-frame point %d: octave %d out of range", i, oct);
    Q.valid[i] = 1;
    Q.x[i] = u[i]; Q.y[i] = v[i];
    Q.r[i] = th * cur->scale[oct];                                // :1371
    if (forward) { Q.minL[i] = oct; Q.maxL[i] = -1; }             // :1375-1380
    else if (backward) { Q.minL[i] = 0; Q.maxL[i] = oct; }
    else { Q.minL[i] = oct - 1; Q.maxL[i] = oct + 1; }
    const float prod = bf * invzc[i];
    Q.xr[i] = u[i] - prod;                                        // :1406 (no FMA)
    Q.angle[i] = last_angle[i];
  }
  Q.desc = mp_desc; Q.hasObs = has_obs; Q.checkUR = true;
  return run_search(cur, Q, ORBFE_MODE_LASTFRAME, 0.f, check_orientation, occupied, assigned, cur->n, n_matches);
}

// OrbMatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (orb_matcher.cpp:133-262)
int orbfe_search_by_bow(orbfe_frame* f, int n_kf, const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_valid,
                        int kf_nnodes, const uint32_t* kf_node_ids, const int32_t* kf_node_start, const uint32_t* kf_feat_idx,
                        int f_nnodes, const uint32_t* f_node_ids, const int32_t* f_node_start, const uint32_t* f_feat_idx,
                        float nnratio, int check_orientation, int32_t* matched_kf_idx, int* n_matches) {
  if (!f || !matched_kf_idx || n_kf < 0 || kf_nnodes < 0 || f_nnodes < 0) return orbfe_fail(ORBFE_ERR_INVALID, "bad arguments");
  if (n_kf && (!kf_desc || !kf_angle || !kf_valid)) return orbfe_fail(ORBFE_ERR_INVALID, "null keyframe array");
  if ((kf_nnodes && (!kf_node_ids || !kf_node_start || !kf_feat_idx)) || (f_nnodes && (!f_node_ids || !f_node_start || !f_feat_idx)))
    return orbfe_fail(ORBFE_ERR_INVALID, "null feature-vector array");
  if (n_matches) *n_matches = 0;
  for (int i = 0; i < f->n; ++i) matched_kf_idx[i] = -1;
  // queries in the reference's visiting order: common nodes ascending (the two-iterator walk with
  // lower_bound, :152-236, is a sorted-set intersection), KeyFrame features in node order
  std::vector<int> qKf, qSrc, qCnt, qOff;
  std::vector<float> qAng;
  int a = 0, b = 0, total = 0;
  while (a < kf_nnodes && b < f_nnodes) {
    if (kf_node_ids[a] == f_node_ids[b]) {
      const int fo = f_node_start[b], fc = f_node_start[b + 1] - fo;
      for (int k = kf_node_start[a]; k < kf_node_start[a + 1]; ++k) {
        const int real = (int)kf_feat_idx[k];
        if (real < 0 || real >= n_kf) return orbfe_fail(ORBFE_ERR_INVALID, "keyframe feature index %d out of range", real);
        if (!kf_valid[real]) continue;  // no map point / bad map point (:162-168)
        qKf.push_back(real); qSrc.push_back(fo); qCnt.push_back(fc); qOff.push_back(total); qAng.push_back(kf_angle[real]);
        total += fc;
      }
      ++a; ++b;
    } else if (kf_node_ids[a] < f_node_ids[b]) ++a;
    else ++b;
  }
  const int nq = (int)qKf.size();
  if (nq == 0 || f->n == 0) return ORBFE_OK;
  const int nfi = f_node_start[f_nnodes];
  for (int k = 0; k < nfi; ++k)
    if ((int)f_feat_idx[k] < 0 || (int)f_feat_idx[k] >= f->n) return orbfe_fail(ORBFE_ERR_INVALID, "frame feature index out of range");
  CUDA_TRY(cudaSetDevice(f->device));
  int rc;
  if ((rc = ensure_queries(f, std::max(nq, n_kf)))) return rc;
  if ((rc = ensure_out(f, std::max(f->n, nfi)))) return rc;
  if ((rc = ensure_cand(f, total + 16))) return rc;
  cudaStream_t st = f->stream;
  if (nq > f->jCap) {
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(regrow(&f->d_jbest, (size_t)nq + 256));
    CUDA_TRY(regrow(&f->d_jchanged, (size_t)nq + 256 + 16));
    f->jCap = nq + 256;
  }
  if (!f->d_jown) CUDA_TRY(regrow(&f->d_jown, 3 * (size_t)std::max(f->n, 1)));
  // scratch re-use: d_qDesc <- all keyframe descriptors, d_qMinL <- descriptor index, d_qMaxL <- source offset,
  // d_out (as unsigned) <- the Frame's feature-vector indices until the finalize kernel overwrites it
  unsigned* d_featIdx = nullptr;
  CUDA_TRY(cudaMalloc(&d_featIdx, (size_t)std::max(nfi, 1) * sizeof(unsigned)));
  auto done = [&](int code) { cudaFree(d_featIdx); return code; };
  cudaError_t e = cudaMemcpyAsync(f->d_qDesc, kf_desc, (size_t)n_kf * 32, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(f->d_qMinL, qKf.data(), (size_t)nq * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(f->d_qMaxL, qSrc.data(), (size_t)nq * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(f->d_qOff, qOff.data(), (size_t)nq * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(f->d_qCnt, qCnt.data(), (size_t)nq * sizeof(int), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(f->d_qAngle, qAng.data(), (size_t)nq * sizeof(float), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_featIdx, f_feat_idx, (size_t)nfi * sizeof(unsigned), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(f->d_cursor, 0, 4 * sizeof(int), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(f->d_jown, 0x7f, 3 * (size_t)std::max(f->n, 1) * sizeof(int), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(f->d_jchanged, 0, ((size_t)nq + 16) * sizeof(int), st);
  if (e != cudaSuccess) return done(orbfe_fail(ORBFE_ERR_CUDA, "SearchByBoW upload failed: %s", cudaGetErrorString(e)));
  MatchScratch S;
  S.cand = f->d_cand; S.qOff = f->d_qOff; S.qCnt = f->d_qCnt; S.cursor = f->d_cursor; S.capacity = f->candCap;
  ResolveArgs A;
  A.mode = ORBFE_MODE_BOW; A.nQ = nq; A.nKp = f->n; A.nnratio = nnratio; A.checkOri = check_orientation; A.hasObs = nullptr;
  A.occupiedIn = nullptr; A.qAngle = f->d_qAngle; A.kp = f->d_kp; A.out = f->d_out; A.evBin = f->d_evBin; A.evIdx = f->d_evIdx;
  A.result = f->d_cursor + 2;
  JacobiState J;
  J.best = f->d_jbest; J.own = f->d_jown; J.changed = f->d_jchanged;
  const int grid = (nq + ORBFE_MATCH_THREADS / 32 - 1) / (ORBFE_MATCH_THREADS / 32);
  MATCH_LAUNCH(f, k_match_candidates_bow, dim3(grid), dim3(ORBFE_MATCH_THREADS), 0, f->grid(), f->d_qDesc, f->d_qMinL, d_featIdx,
               f->d_qMaxL, nq, S);
  int t = 0;
  for (;;) {
    for (int k = 0; k < 8 && t <= nq; ++k, ++t) MATCH_LAUNCH(f, k_match_iterate, dim3(grid), dim3(ORBFE_MATCH_THREADS), 0, A, S, J, t);
    e = cudaMemcpyAsync(f->h_res + 3, f->d_jchanged + (t - 1), sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return done(orbfe_fail(ORBFE_ERR_CUDA, "SearchByBoW resolve failed: %s", cudaGetErrorString(e)));
    if (f->h_res[3] == 0 || t > nq) break;
  }
  MATCH_LAUNCH(f, k_match_finalize, dim3(1), dim3(1024), 0, A, S, J);
  e = cudaGetLastError();
  std::vector<int> assigned((size_t)f->n);
  if (e == cudaSuccess) e = cudaMemcpyAsync(f->h_res, f->d_cursor, 3 * sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(assigned.data(), f->d_out, (size_t)f->n * sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) return done(orbfe_fail(ORBFE_ERR_CUDA, "SearchByBoW failed: %s", cudaGetErrorString(e)));
  for (int i = 0; i < f->n; ++i) matched_kf_idx[i] = assigned[i] >= 0 ? qKf[assigned[i]] : -1;  // query -> keyframe feature
  if (n_matches) *n_matches = f->h_res[2];
  return done(ORBFE_OK);
}

}  // extern "C"
