// orbfe_match.cu -- host side of the OrbMatcher / Frame-grid entry points of include/orbfe.h
// (kernels in k_match.cuh).  A frame handle owns the device copy of one Frame's matcher view
// (undistorted keypoints, descriptors, stereo coordinates, 64x48 grid) plus query scratch and a
// private stream; calls on one handle are serialised by the caller.
#include "../../include/orbfe.h"

#include "k_match.cuh"
#include "k_frame.cuh"
#include "orbfe_host.h"

#include <cmath>
#include <cstring>
#include <mutex>
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

#define ORBFE_CURSOR_INTS 40   // [0] cursor [1] overflow [2] nmatches [3] init overflow [4..6] solve flags [7] misc counter [8..38) histogram

struct orbfe_frame {
  int device = 0;
  cudaStream_t stream = nullptr;
  int n = 0, nlevels = 0;
  int kpCap = 0;  // capacity of the per-keypoint device arrays (handles refreshed from an extractor grow in place)
  float minX = 0, maxX = 0, minY = 0, maxY = 0, gw = 1, gh = 1;
  std::vector<float> scale;
  std::vector<MatchKp> hkp;  // host copy (prevMatched update, validation)
  MatchKp* d_kp = nullptr;
  uint8_t* d_desc = nullptr;
  float* d_uR = nullptr;
  float* d_lvl = nullptr;   // scale_factors | level_sigma_sq | inv_level_sigma_sq
  int* d_cellStart = nullptr;
  int* d_cellItems = nullptr;
  // per-query device scratch (grown on demand); the query inputs themselves live in d_stage
  int qCap = 0;
  int *d_qOff = nullptr, *d_qCnt = nullptr, *d_evBin = nullptr, *d_evIdx = nullptr;
  int* d_out = nullptr;
  int outCap = 0;
  uint2* d_cand = nullptr;
  int candCap = 0;
  int* d_jbest = nullptr;   // Jacobi resolve state (k_match_iterate)
  int* d_jown = nullptr;
  int* d_jchanged = nullptr;
  int jCap = 0;
  char* d_lp = nullptr;           // orbfe_search_local_points: frustum inputs / outputs
  size_t lpCap = 0;
  unsigned* d_featIdx = nullptr;  // vocabulary-node searches: the searched frame's flattened feature-vector indices
  int featIdxCap = 0;
  uint2* d_islots = nullptr;  // SearchForInitialization Jacobi: acceptor slots (3 x n x ORBFE_INIT_SLOTS)
  int* d_iowner = nullptr;
  int* d_cursor = nullptr;  // [0] cursor [1] overflow [2] nmatches [3] misc counter / init overflow [4..6] solve flags
  int* h_res = nullptr;     // pinned, 4 ints
  // one search = one H2D copy: the queries are packed into a pinned host block whose device mirror the kernels read
  char* h_stage = nullptr;
  char* d_stage = nullptr;
  size_t stageCap = 0;
  char* h_outStage = nullptr;   // pinned landing zone of the results ([0,16): d_cursor, then d_out, then routine-specific extras)
  size_t outStageCap = 0;
  int coopBlocks = 0;           // co-resident CTAs of k_match_solve on this device (0 = not queried yet)
  bool complete = false;        // stream and fixed-size blocks exist (frame_acquire succeeded): the handle may be recycled
  template <class T> T* dev(const T* hostPtr) const {
    return reinterpret_cast<T*>(d_stage + (reinterpret_cast<const char*>(hostPtr) - h_stage));
  }
  FrameGrid grid() const {
    FrameGrid G;
    G.kp = d_kp; G.desc = d_desc; G.uR = d_uR; G.cellStart = d_cellStart; G.cellItems = d_cellItems; G.n = n;
    G.lvl = d_lvl; G.nlevels = nlevels;
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

static inline size_t al16(size_t v) { return (v + 15) & ~(size_t)15; }

// device-only per-query scratch (candidate ranges, histogram bins)
static int ensure_queries(orbfe_frame* f, int nq) {
  if (nq <= f->qCap) return ORBFE_OK;
  const size_t c = (size_t)nq + 256;
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  CUDA_TRY(regrow(&f->d_qOff, c)); CUDA_TRY(regrow(&f->d_qCnt, c)); CUDA_TRY(regrow(&f->d_evBin, c)); CUDA_TRY(regrow(&f->d_evIdx, c));
  f->qCap = (int)c;
  return ORBFE_OK;
}
// the packed query block (pinned host + device mirror), grow-only
static int ensure_stage(orbfe_frame* f, size_t bytes) {
  if (bytes <= f->stageCap) return ORBFE_OK;
  const size_t c = al16(bytes + bytes / 4 + 4096);
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  if (f->h_stage) cudaFreeHost(f->h_stage);
  if (f->d_stage) cudaFree(f->d_stage);
  f->h_stage = nullptr; f->d_stage = nullptr; f->stageCap = 0;
  CUDA_TRY(cudaMallocHost(&f->h_stage, c));
  CUDA_TRY(cudaMalloc(&f->d_stage, c));
  f->stageCap = c;
  return ORBFE_OK;
}
static int ensure_out_stage(orbfe_frame* f, size_t bytes) {
  if (bytes <= f->outStageCap) return ORBFE_OK;
  const size_t c = al16(bytes + bytes / 4 + 4096);
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  if (f->h_outStage) cudaFreeHost(f->h_outStage);
  f->h_outStage = nullptr; f->outStageCap = 0;
  CUDA_TRY(cudaMallocHost(&f->h_outStage, c));
  f->outStageCap = c;
  return ORBFE_OK;
}
template <class T>
static T* stage_take(orbfe_frame* f, size_t& off, size_t count) {
  off = al16(off);
  T* p = reinterpret_cast<T*>(f->h_stage + off);
  off += count * sizeof(T);
  return p;
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

// the window queries of one search, laid out in the frame's pinned staging block:
//   [front: routine-specific inputs][desc nq x 32][hasObs nq][occupied nKp]  <- always uploaded
//   [x y r xr angle minL maxL : nq x 4 each][valid nq]                       <- uploaded unless produced on the device
struct HostQueries {
  float *x = nullptr, *y = nullptr, *r = nullptr, *xr = nullptr, *angle = nullptr;
  int *minL = nullptr, *maxL = nullptr;
  uint8_t *valid = nullptr, *desc = nullptr, *hasObs = nullptr, *occ = nullptr;
  const uint8_t* descSrc = nullptr;     // nq x 32 (host), copied into the block ...
  const uint8_t* descDev = nullptr;     // ... or descriptors that already live on the device (same device as the searched frame)
  const uint8_t* hasObsSrc = nullptr;   // nq (host) or null
  int n = 0;
  int filter = ORBFE_FILTER_NONE;
  bool onDevice = false;                // x, y, r, xr, minL, maxL, valid are written by a kernel, not uploaded
  size_t front = 0, headBytes = 0, allBytes = 0;
  // carves the block for nq queries on frame f (the searched frame); `frontBytes` are reserved at the start for the caller
  int bind(orbfe_frame* f, int nq, size_t frontBytes = 0, bool producedOnDevice = false) {
    n = nq; onDevice = producedOnDevice; front = al16(frontBytes);
    const size_t N = (size_t)std::max(nq, 1), K = (size_t)std::max(f->n, 1);
    const size_t need = front + al16(N * 32) + al16(N) + al16(K) + 7 * al16(N * 4) + al16(N) + 64;
    cudaError_t ce = cudaSetDevice(f->device);
    if (ce != cudaSuccess) return orbfe_fail(ORBFE_ERR_CUDA, "cudaSetDevice failed: %s", cudaGetErrorString(ce));
    int rc;
    if ((rc = ensure_stage(f, need))) return rc;
    size_t off = front;
    desc = stage_take<uint8_t>(f, off, N * 32);
    hasObs = stage_take<uint8_t>(f, off, N);
    occ = stage_take<uint8_t>(f, off, K);
    headBytes = al16(off);
    x = stage_take<float>(f, off, N); y = stage_take<float>(f, off, N); r = stage_take<float>(f, off, N);
    xr = stage_take<float>(f, off, N); angle = stage_take<float>(f, off, N);
    minL = stage_take<int>(f, off, N); maxL = stage_take<int>(f, off, N);
    valid = stage_take<uint8_t>(f, off, N);
    allBytes = al16(off);
    if (!onDevice) {  // defaults of a query the routine skips: invalid, level gate off
      memset(x, 0, (size_t)((char*)minL - (char*)x));
      memset(minL, 0xff, (size_t)((char*)valid - (char*)minL));
      memset(valid, 0, N);
    }
    return ORBFE_OK;
  }
};

// what couples / accepts the queries of one search routine (ResolveArgs fields, k_match.cuh)
struct SearchSpec {
  int mode = ORBFE_MODE_GENERIC;
  float nnratio = 0.f;
  int checkOri = 0;
  int thAccept = 100;
  int ratio = ORBFE_RATIO_NONE;
  int feedback = ORBFE_FEEDBACK_NONE;
  int tieLast = 0;
  int perQuery = 0;
};
static void fill_args(ResolveArgs& A, const SearchSpec& sp) {
  A.mode = sp.mode; A.nnratio = sp.nnratio; A.checkOri = sp.checkOri; A.thAccept = sp.thAccept; A.ratio = sp.ratio;
  A.feedback = sp.feedback; A.tieLast = sp.tieLast; A.perQuery = sp.perQuery;
}

// where the candidates of a search come from (device pointers)
struct SolveInput {
  int bow = 0;
  int nq = 0;
  MatchQueries MQ;            // bow == 0
  BowQueries BQ;              // bow == 1 (S.qOff / S.qCnt are inputs then)
  BowFilter B;
  const int* qOff = nullptr;  // bow == 1
  const int* qCnt = nullptr;
  const uint8_t* hasObs = nullptr;
  const uint8_t* occupied = nullptr;
  const float* qAngle = nullptr;
};

// candidates -> fixed point -> finalize on `f` (the searched frame), then the D2H copy of the result: d_out[0, out_n) lands in
// `out` (may be null: the result stays in f->d_out), the match count in *nmatches.  CUDA build: ONE cooperative launch
// (k_match_solve) + ONE stream synchronisation per search; the emulated test build runs the same device functions as
// separate launches (its blocks run one after the other, so it has no grid barrier).
static int solve(orbfe_frame* f, SolveInput& in, const SearchSpec& sp, int32_t* out, int out_n, int* nmatches) {
  const int mode = sp.mode;
  const int nq = in.nq;
  int rc;
  if ((rc = ensure_queries(f, nq))) return rc;
  if ((rc = ensure_out(f, out_n))) return rc;
  if ((rc = ensure_out_stage(f, 16 + (size_t)std::max(out_n, 0) * sizeof(int)))) return rc;
  if (!in.bow && f->candCap == 0 && (rc = ensure_cand(f, std::max(nq * 48 + 4096, 1 << 16)))) return rc;
  const size_t stateInts = mode == ORBFE_MODE_INIT ? 2 * (size_t)f->n : (size_t)f->n;
  const size_t smem = std::max<size_t>(stateInts * sizeof(int), 16);
  if (mode == ORBFE_MODE_INIT && smem > 200 * 1024) return orbfe_fail(ORBFE_ERR_INVALID, "frame has too many keypoints (%d) for the resolve kernel", f->n);
  cudaStream_t st = f->stream;
  const bool init = mode == ORBFE_MODE_INIT;
  const bool initJacobi = init && f->n < (1 << 22) && nq > 0;  // parallel SearchForInitialization (k_init_iterate)
  if (nq > f->jCap || !f->d_jbest) {
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(regrow(&f->d_jbest, (size_t)nq + 256));
    CUDA_TRY(regrow(&f->d_jchanged, (size_t)nq + 256 + 16));
    f->jCap = nq + 256;
  }
  const size_t kp1 = (size_t)std::max(std::max(f->n, f->kpCap), 1);
  if (!f->d_jown) CUDA_TRY(regrow(&f->d_jown, 3 * kp1));
  if (initJacobi && !f->d_islots) {
    CUDA_TRY(regrow(&f->d_islots, 3 * kp1 * ORBFE_INIT_SLOTS));
    CUDA_TRY(regrow(&f->d_iowner, kp1));
  }
  int* hres = reinterpret_cast<int*>(f->h_outStage);
  int* hout = hres + 4;
  for (int attempt = 0; attempt < 8; ++attempt) {
    CUDA_TRY(cudaMemsetAsync(f->d_cursor, 0, ORBFE_CURSOR_INTS * sizeof(int), st));
    MatchScratch S;
    S.cand = f->d_cand; S.cursor = f->d_cursor; S.capacity = f->candCap;
    S.qOff = in.bow ? const_cast<int*>(in.qOff) : f->d_qOff;
    S.qCnt = in.bow ? const_cast<int*>(in.qCnt) : f->d_qCnt;
    ResolveArgs A;
    fill_args(A, sp);
    A.nQ = nq; A.nKp = f->n; A.hasObs = in.hasObs; A.occupiedIn = in.occupied; A.qAngle = in.qAngle; A.kp = f->d_kp; A.out = f->d_out;
    A.evBin = f->d_evBin; A.evIdx = f->d_evIdx; A.result = f->d_cursor + 2;
    JacobiState J;
    J.best = f->d_jbest; J.own = f->d_jown; J.changed = f->d_jchanged;
    InitJacobi IJ;
    IJ.acc = f->d_jbest; IJ.cnt = f->d_jown; IJ.slots = f->d_islots; IJ.changed = f->d_jchanged; IJ.overflow = f->d_cursor + 3;
    const FrameGrid G = f->grid();
    bool serialInit = init && !initJacobi;
#ifndef ORBFE_EMU
    if (nq > 0 && !serialInit) {
      if (!f->coopBlocks) {
        int perSm = 0, sms = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, k_match_solve, ORBFE_SOLVE_THREADS, 0));
        CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, f->device));
        f->coopBlocks = std::max(1, perSm * sms);
      }
      SolveCfg C;
      C.bow = in.bow; C.init = init ? 1 : 0; C.flags = f->d_cursor + 4; C.iowner = f->d_iowner; C.hist = f->d_cursor + 8;
      const int wpb = ORBFE_SOLVE_THREADS / 32;
      const int blocks = std::max(1, std::min(f->coopBlocks, (nq + wpb - 1) / wpb));
      void* args[] = {(void*)&G, (void*)&in.MQ, (void*)&in.BQ, (void*)&in.B, (void*)&S, (void*)&A, (void*)&J, (void*)&IJ, (void*)&C};
      CUDA_TRY(cudaLaunchCooperativeKernel((const void*)k_match_solve, dim3(blocks), dim3(ORBFE_SOLVE_THREADS), args, 0, st));
    }
#else
    const int grid = std::max(1, (nq + ORBFE_MATCH_THREADS / 32 - 1) / (ORBFE_MATCH_THREADS / 32));
    if (nq > 0) {
      if (in.bow) MATCH_LAUNCH(f, k_match_candidates_bow, dim3(grid), dim3(ORBFE_MATCH_THREADS), 0, G, in.BQ, S, in.B);
      else MATCH_LAUNCH(f, k_match_candidates, dim3(grid), dim3(ORBFE_MATCH_THREADS), 0, G, in.MQ, S);
    }
    if (nq > 0 && !serialInit) {
      CUDA_TRY(cudaMemsetAsync(f->d_jown, init ? 0 : 0x7f, 3 * (size_t)std::max(f->n, 1) * sizeof(int), st));
      CUDA_TRY(cudaMemsetAsync(f->d_jchanged, 0, ((size_t)nq + 16) * sizeof(int), st));
      const bool coupled = init || sp.feedback != ORBFE_FEEDBACK_NONE;
      for (int t = 0; t <= nq; ++t) {
        if (init) MATCH_LAUNCH(f, k_init_iterate, dim3(grid), dim3(ORBFE_MATCH_THREADS), 0, A, S, IJ, t);
        else MATCH_LAUNCH(f, k_match_iterate, dim3(grid), dim3(ORBFE_MATCH_THREADS), 0, A, S, J, t);
        if (!coupled) break;
        CUDA_TRY(cudaMemcpyAsync(f->h_res + 3, f->d_jchanged + t, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        if (f->h_res[3] == 0) break;
      }
      if (init) MATCH_LAUNCH(f, k_init_finalize, dim3(1), dim3(1024), 0, A, S, IJ, f->d_iowner);
      else MATCH_LAUNCH(f, k_match_finalize, dim3(1), dim3(1024), 0, A, S, J);
    }
#endif
    if (nq == 0 && !init) {  // no queries: an all -1 result (perQuery results have no entries)
      if (out_n > 0) CUDA_TRY(cudaMemsetAsync(f->d_out, 0xff, (size_t)out_n * sizeof(int), st));
    }
    if (serialInit) {
#ifndef ORBFE_EMU
      CUDA_TRY(orbfe_raise_dynamic_smem(k_match_resolve, f->device, smem));
      if (nq > 0) k_match_candidates<<<std::max(1, (nq + ORBFE_MATCH_THREADS / 32 - 1) / (ORBFE_MATCH_THREADS / 32)), ORBFE_MATCH_THREADS, 0, st>>>(G, in.MQ, S);
#endif
      MATCH_LAUNCH(f, k_match_resolve, dim3(1), dim3(32), smem, A, S);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(hres, f->d_cursor, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
    if (out_n > 0 && out) CUDA_TRY(cudaMemcpyAsync(hout, f->d_out, (size_t)out_n * sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (hres[1]) {  // candidate buffer too small: hres[0] is the total that was requested
      if ((rc = ensure_cand(f, hres[0] + 4096))) return rc;
      continue;
    }
    if (init && !serialInit && hres[3]) {
      // a keypoint had more than ORBFE_INIT_SLOTS acceptors: exact serial resolve on the candidates already built
#ifndef ORBFE_EMU
      CUDA_TRY(orbfe_raise_dynamic_smem(k_match_resolve, f->device, smem));
#endif
      MATCH_LAUNCH(f, k_match_resolve, dim3(1), dim3(32), smem, A, S);
      CUDA_TRY(cudaGetLastError());
      CUDA_TRY(cudaMemcpyAsync(hres, f->d_cursor, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
      if (out_n > 0 && out) CUDA_TRY(cudaMemcpyAsync(hout, f->d_out, (size_t)out_n * sizeof(int), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(cudaStreamSynchronize(st));
    }
    if (out_n > 0 && out) memcpy(out, hout, (size_t)out_n * sizeof(int));
    if (nmatches) *nmatches = hres[2];
    return ORBFE_OK;
  }
  return orbfe_fail(ORBFE_ERR_CUDA, "candidate buffer did not converge");
}

// fills the source arrays of a bound HostQueries, uploads the block with ONE copy and runs the search on `f`
static int run_search(orbfe_frame* f, HostQueries& Q, const SearchSpec& sp, const uint8_t* occupied, int32_t* out, int out_n,
                      int* nmatches, bool uploadDone = false) {
  CUDA_TRY(cudaSetDevice(f->device));
  const int nq = Q.n;
  if (!uploadDone) {
    if (Q.descSrc && nq) memcpy(Q.desc, Q.descSrc, (size_t)nq * 32);
    if (Q.hasObsSrc && nq) memcpy(Q.hasObs, Q.hasObsSrc, (size_t)nq);
    if (occupied && f->n) memcpy(Q.occ, occupied, (size_t)f->n);
    const size_t bytes = (Q.onDevice ? Q.headBytes : Q.allBytes) - Q.front;
    CUDA_TRY(cudaMemcpyAsync(f->d_stage + Q.front, f->h_stage + Q.front, bytes, cudaMemcpyHostToDevice, f->stream));
  }
  SolveInput in;
  in.nq = nq;
  MatchQueries& MQ = in.MQ;
  MQ.x = f->dev(Q.x); MQ.y = f->dev(Q.y); MQ.r = f->dev(Q.r); MQ.xr = f->dev(Q.xr); MQ.minLevel = f->dev(Q.minL); MQ.maxLevel = f->dev(Q.maxL);
  MQ.valid = f->dev(Q.valid); MQ.desc = Q.descDev ? Q.descDev : f->dev(Q.desc); MQ.n = nq; MQ.filter = Q.filter;
  in.BQ = BowQueries{nullptr, nullptr, nullptr, nullptr, 0};
  memset(&in.B, 0, sizeof(in.B));
  in.hasObs = Q.hasObsSrc ? f->dev(Q.hasObs) : nullptr;
  in.occupied = (occupied || sp.mode == ORBFE_MODE_INIT) ? f->dev(Q.occ) : nullptr;
  in.qAngle = f->dev(Q.angle);
  return solve(f, in, sp, out, out_n, nmatches);
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

// ---- handle recycling --------------------------------------------------------------------------------------------------
// Tracking builds one matcher view per frame and drops it a few frames later.  A destroyed handle keeps its stream, its device
// arrays and its pinned staging blocks and waits in a small per-process pool; the next orbfe_frame_create on that device takes
// it over (arrays only ever grow), so a steady-state frame costs no cudaMalloc / cudaMallocHost / stream creation.
static std::mutex g_poolMu;
static std::vector<orbfe_frame*> g_pool;
static const size_t kPoolMax = 16;

static void frame_free(orbfe_frame* f) {
  cudaSetDevice(f->device);
  cudaFree(f->d_kp); cudaFree(f->d_desc); cudaFree(f->d_uR); cudaFree(f->d_lvl); cudaFree(f->d_cellStart); cudaFree(f->d_cellItems);
  cudaFree(f->d_qOff); cudaFree(f->d_qCnt); cudaFree(f->d_evBin); cudaFree(f->d_evIdx);
  cudaFree(f->d_out);
  cudaFree(f->d_cand); cudaFree(f->d_cursor); cudaFree(f->d_jbest); cudaFree(f->d_jown); cudaFree(f->d_jchanged); cudaFree(f->d_islots); cudaFree(f->d_iowner); cudaFree(f->d_featIdx); cudaFree(f->d_lp);
  cudaFreeHost(f->h_res);
  if (f->h_stage) cudaFreeHost(f->h_stage);
  if (f->h_outStage) cudaFreeHost(f->h_outStage);
  cudaFree(f->d_stage);
  if (f->stream) cudaStreamDestroy(f->stream);
  delete f;
}

// an idle handle of `device` (the one with the largest keypoint capacity), or null
static orbfe_frame* pool_take(int device) {
  std::lock_guard<std::mutex> lock(g_poolMu);
  int best = -1;
  for (size_t i = 0; i < g_pool.size(); ++i)
    if (g_pool[i]->device == device && (best < 0 || g_pool[i]->kpCap > g_pool[best]->kpCap)) best = (int)i;
  if (best < 0) return nullptr;
  orbfe_frame* f = g_pool[best];
  g_pool.erase(g_pool.begin() + best);
  return f;
}

int orbfe_frame_destroy(orbfe_frame* f) {
  if (!f) return ORBFE_OK;
  cudaSetDevice(f->device);
  if (f->stream) cudaStreamSynchronize(f->stream);
  if (f->complete && !getenv("ORBFE_NO_HANDLE_POOL")) {
    std::lock_guard<std::mutex> lock(g_poolMu);
    if (g_pool.size() < kPoolMax) {
      f->n = 0;
      g_pool.push_back(f);
      return ORBFE_OK;
    }
  }
  frame_free(f);
  return ORBFE_OK;
}

// frees the idle handles kept for recycling (tests, orderly shutdown); live handles are not touched
int orbfe_frame_pool_trim(void) {
  std::vector<orbfe_frame*> idle;
  {
    std::lock_guard<std::mutex> lock(g_poolMu);
    idle.swap(g_pool);
  }
  for (orbfe_frame* f : idle) frame_free(f);
  return (int)idle.size();
}

// a handle on `device` with its stream and fixed-size blocks: recycled or new
static int frame_acquire(int device, orbfe_frame** out) {
  *out = nullptr;
  CUDA_TRY(cudaSetDevice(device));
  orbfe_frame* f = getenv("ORBFE_NO_HANDLE_POOL") ? nullptr : pool_take(device);
  if (!f) {
    f = new (std::nothrow) orbfe_frame();
    if (!f) return orbfe_fail(ORBFE_ERR_NOMEM, "out of host memory");
    f->device = device;
    cudaError_t e = cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaMalloc(&f->d_cellStart, (ORBFE_GRID_CELLS + 1) * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&f->d_cursor, ORBFE_CURSOR_INTS * sizeof(int));
    if (e == cudaSuccess) e = cudaMalloc(&f->d_lvl, 3 * ORBFE_MAX_LEVELS * sizeof(float));
    if (e == cudaSuccess) e = cudaMallocHost(&f->h_res, 4 * sizeof(int));
    if (e != cudaSuccess) {
      frame_free(f);
      return orbfe_fail(ORBFE_ERR_CUDA, "frame setup failed: %s", cudaGetErrorString(e));
    }
  }
  f->complete = true;
  *out = f;
  return ORBFE_OK;
}

// per-keypoint device arrays for n keypoints (grow-only; the lazily sized resolve buffers follow a growth)
static int frame_reserve_keypoints(orbfe_frame* f, int n, int hint) {
  if (n <= f->kpCap && f->d_kp) return ORBFE_OK;
  const size_t c = (size_t)std::max(std::max(n, hint), 1);
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  CUDA_TRY(regrow(&f->d_kp, c)); CUDA_TRY(regrow(&f->d_desc, c * 32)); CUDA_TRY(regrow(&f->d_uR, c));
  CUDA_TRY(regrow(&f->d_cellItems, c));
  cudaFree(f->d_jown); f->d_jown = nullptr;
  cudaFree(f->d_islots); f->d_islots = nullptr;
  cudaFree(f->d_iowner); f->d_iowner = nullptr;
  f->kpCap = (int)c;
  return ORBFE_OK;
}

int orbfe_frame_create(int device, int n, const orbfe_keypoint* kps_un, const uint8_t* desc, const float* u_right,
                       float min_x, float max_x, float min_y, float max_y, int nlevels, const float* scale_factors,
                       orbfe_frame** out) {
  if (!out) return orbfe_fail(ORBFE_ERR_INVALID, "null argument");
  *out = nullptr;
  if (n < 0 || (n && (!kps_un || !desc)) || nlevels < 1 || nlevels > ORBFE_MAX_LEVELS || !scale_factors || !(max_x > min_x) || !(max_y > min_y))
    return orbfe_fail(ORBFE_ERR_INVALID, "bad frame arguments");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); ndev = 0; }
  if (device < 0 || device >= ndev)
    return orbfe_fail(ORBFE_ERR_CUDA, "CUDA device %d not available (%d visible); this library has no CPU path", device, ndev);
  for (int i = 0; i < n; ++i)
    if (kps_un[i].octave < 0 || kps_un[i].octave >= nlevels)
      return orbfe_fail(ORBFE_ERR_INVALID, "keypoint %d has octave %d outside [0,%d)", i, kps_un[i].octave, nlevels);
  orbfe_frame* f = nullptr;
  int rc;
  if ((rc = frame_acquire(device, &f))) return rc;
  auto fail = [&](int code) { frame_free(f); return code; };
  f->n = n; f->nlevels = nlevels;
  f->minX = min_x; f->maxX = max_x; f->minY = min_y; f->maxY = max_y;
  f->gw = static_cast<float>(max_x - min_x) / ORBFE_GRID_COLS;  // frame.cpp:223-224
  f->gh = static_cast<float>(max_y - min_y) / ORBFE_GRID_ROWS;
  f->scale.assign(scale_factors, scale_factors + nlevels);
  f->hkp.resize(n);
  if ((rc = frame_reserve_keypoints(f, n, 0))) return fail(rc);
  // keypoints, descriptors, stereo coordinates and the level tables travel in ONE pinned block: one H2D copy
  const size_t N = (size_t)std::max(n, 1);
  const size_t need = al16(N * sizeof(MatchKp)) + al16(N * 32) + al16(N * 4) + al16(3 * ORBFE_MAX_LEVELS * 4) + 64;
  if ((rc = ensure_stage(f, need))) return fail(rc);
  size_t off = 0;
  MatchKp* hK = stage_take<MatchKp>(f, off, N);
  uint8_t* hD = stage_take<uint8_t>(f, off, N * 32);
  float* hU = stage_take<float>(f, off, N);
  float* hL = stage_take<float>(f, off, 3 * (size_t)nlevels);  // ORBextractor tables (orb_extractor.cpp:356-372), float arithmetic
  for (int l = 0; l < nlevels; ++l) {
    hL[l] = scale_factors[l];
    hL[nlevels + l] = l == 0 ? 1.0f : scale_factors[l] * scale_factors[l];
    hL[2 * nlevels + l] = 1.0f / hL[nlevels + l];
  }
  for (int i = 0; i < n; ++i) {
    f->hkp[i] = MatchKp{kps_un[i].x, kps_un[i].y, kps_un[i].angle, kps_un[i].octave};
    hU[i] = u_right ? u_right[i] : -1.0f;
  }
  if (n) { memcpy(hK, f->hkp.data(), (size_t)n * sizeof(MatchKp)); memcpy(hD, desc, (size_t)n * 32); }
  cudaStream_t st = f->stream;
  cudaError_t e = cudaMemcpyAsync(f->d_stage, f->h_stage, al16(off), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess && n) e = cudaMemcpyAsync(f->d_kp, f->dev(hK), (size_t)n * sizeof(MatchKp), cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess && n) e = cudaMemcpyAsync(f->d_desc, f->dev(hD), (size_t)n * 32, cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess && n) e = cudaMemcpyAsync(f->d_uR, f->dev(hU), (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(f->d_lvl, f->dev(hL), 3 * (size_t)nlevels * sizeof(float), cudaMemcpyDeviceToDevice, st);
  if (e != cudaSuccess) return fail(orbfe_fail(ORBFE_ERR_CUDA, "frame setup failed: %s", cudaGetErrorString(e)));
  MATCH_LAUNCH(f, k_grid_build, dim3(1), dim3(1024), 0, f->d_kp, n, f->minX, f->minY, f->gw, f->gh, f->d_cellStart, f->d_cellItems);
  e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // the staging block is reused by the first search
  if (e != cudaSuccess) return fail(orbfe_fail(ORBFE_ERR_CUDA, "frame setup failed: %s", cudaGetErrorString(e)));
  *out = f;
  return ORBFE_OK;
}

// A Frame's matcher view built from the DEVICE-resident results of an extractor slot (after orbfe_run / orbfe_extract, and
// orbfe_run_stereo when use_stereo): keypoints, descriptors and stereo coordinates go device to device, so tracking a frame
// needs no D2H -> H2D round trip of its own features.  Undistortion must be the identity (rectified input, dist_coeff[0] == 0,
// frame.cpp:616-619); otherwise build the handle from the undistorted host keypoints with orbfe_frame_create.
// (re)fills `f` from the slot; device arrays only ever grow, so a handle refreshed every frame allocates nothing in steady state
static int frame_fill_from_slot(orbfe_frame* f, orbfe_extractor* ex, int slot, int use_stereo, float min_x, float max_x, float min_y,
                                float max_y) {
  if (!(max_x > min_x) || !(max_y > min_y)) return orbfe_fail(ORBFE_ERR_INVALID, "bad image bounds");
  OrbfeSlotView V;
  int rc;
  if ((rc = orbfe_internal_slot_view(ex, slot, &V))) return rc;
  if (f->stream && f->device != V.device) return orbfe_fail(ORBFE_ERR_INVALID, "frame handle and extractor live on different devices");
  CUDA_TRY(cudaSetDevice(V.device));
  CUDA_TRY(cudaStreamSynchronize(static_cast<cudaStream_t>(V.stream)));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  CUDA_TRY(cudaMemcpyAsync(f->h_res, V.nKp, sizeof(int), cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  const int n = std::max(0, std::min(f->h_res[0], V.capacity));
  int rcg;
  if ((rcg = frame_reserve_keypoints(f, n, V.capacity > 0 ? std::min(V.capacity, n + n / 4 + 256) : 1))) return rcg;
  if (V.nlevels > ORBFE_MAX_LEVELS) return orbfe_fail(ORBFE_ERR_INVALID, "too many pyramid levels");
  {  // level tables (a recycled handle may hold another extractor's): 3 x nlevels floats, synchronous and tiny
    std::vector<float> lvl(3 * (size_t)V.nlevels);
    for (int l = 0; l < V.nlevels; ++l) {
      lvl[l] = V.scale[l];
      lvl[V.nlevels + l] = l == 0 ? 1.0f : V.scale[l] * V.scale[l];
      lvl[2 * V.nlevels + l] = 1.0f / lvl[V.nlevels + l];
    }
    if (f->nlevels != V.nlevels || f->scale.size() != (size_t)V.nlevels || !std::equal(f->scale.begin(), f->scale.end(), V.scale))
      CUDA_TRY(cudaMemcpy(f->d_lvl, lvl.data(), lvl.size() * sizeof(float), cudaMemcpyHostToDevice));
    f->scale.assign(V.scale, V.scale + V.nlevels);
    f->nlevels = V.nlevels;
  }
  f->n = n;
  f->minX = min_x; f->maxX = max_x; f->minY = min_y; f->maxY = max_y;
  f->gw = static_cast<float>(max_x - min_x) / ORBFE_GRID_COLS;
  f->gh = static_cast<float>(max_y - min_y) / ORBFE_GRID_ROWS;
  f->hkp.resize(n);
  if (n) CUDA_TRY(cudaMemcpyAsync(f->d_desc, V.desc, (size_t)n * 32, cudaMemcpyDeviceToDevice, f->stream));
  if (n) MATCH_LAUNCH(f, k_kp_to_match, dim3((n + 255) / 256), dim3(256), 0, static_cast<const float*>(V.kps), n, f->d_kp,
                      use_stereo ? V.uR : nullptr, f->d_uR);
  MATCH_LAUNCH(f, k_grid_build, dim3(1), dim3(1024), 0, f->d_kp, n, f->minX, f->minY, f->gw, f->gh, f->d_cellStart, f->d_cellItems);
  CUDA_TRY(cudaGetLastError());
  if (n) CUDA_TRY(cudaMemcpyAsync(f->hkp.data(), f->d_kp, (size_t)n * sizeof(MatchKp), cudaMemcpyDeviceToHost, f->stream));
  CUDA_TRY(cudaStreamSynchronize(f->stream));
  return ORBFE_OK;
}

int orbfe_frame_from_extractor(orbfe_extractor* ex, int slot, int use_stereo, float min_x, float max_x, float min_y, float max_y,
                               orbfe_frame** out) {
  if (!out) return orbfe_fail(ORBFE_ERR_INVALID, "null argument");
  *out = nullptr;
  OrbfeSlotView V0;
  int rc;
  if ((rc = orbfe_internal_slot_view(ex, slot, &V0))) return rc;
  orbfe_frame* f = nullptr;
  if ((rc = frame_acquire(V0.device, &f))) return rc;
  f->scale.clear(); f->nlevels = 0;   // a recycled handle: force the level tables of THIS extractor
  rc = frame_fill_from_slot(f, ex, slot, use_stereo, min_x, max_x, min_y, max_y);
  if (rc) { frame_free(f); return rc; }
  *out = f;
  return ORBFE_OK;
}

// the per-frame form: refresh an existing handle (from orbfe_frame_from_extractor) with the next frame's results
int orbfe_frame_refresh_from_extractor(orbfe_frame* f, orbfe_extractor* ex, int slot, int use_stereo, float min_x, float max_x,
                                       float min_y, float max_y) {
  if (!f) return orbfe_fail(ORBFE_ERR_INVALID, "null frame handle");
  return frame_fill_from_slot(f, ex, slot, use_stereo, min_x, max_x, min_y, max_y);
}

int orbfe_frame_num_keypoints(const orbfe_frame* f) { return f ? f->n : 0; }

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
  int rc;
  if ((rc = Q.bind(f2, n1))) return rc;
  CUDA_TRY(cudaStreamSynchronize(f1->stream));   // F1's descriptors are read where they are (same device)
  for (int i = 0; i < n1; ++i) {
    const int level1 = f1->hkp[i].octave;
    Q.valid[i] = level1 > 0 ? 0 : 1;
    Q.x[i] = prev_matched_xy[2 * i]; Q.y[i] = prev_matched_xy[2 * i + 1];
    Q.r[i] = (float)window_size;
    Q.minL[i] = level1; Q.maxL[i] = level1;
    Q.angle[i] = f1->hkp[i].angle;
  }
  Q.descDev = f1->d_desc;
  int nm = 0;
  SearchSpec sp;
  sp.mode = ORBFE_MODE_INIT; sp.nnratio = nnratio; sp.checkOri = check_orientation;
  if ((rc = run_search(f2, Q, sp, nullptr, matches12, n1, &nm))) return rc;
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
  int rc;
  if ((rc = Q.bind(f, n_mp))) return rc;
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
  Q.descSrc = mp_desc; Q.hasObsSrc = has_obs; Q.filter = ORBFE_FILTER_UR;
  SearchSpec sp;  // TH_HIGH, ratio test against a second best of the same level (:87-97), SetMapPoint feedback (:59-63)
  sp.mode = ORBFE_MODE_MAPPOINTS; sp.nnratio = nnratio; sp.thAccept = 100; sp.ratio = ORBFE_RATIO_SAMELEVEL;
  sp.feedback = ORBFE_FEEDBACK_HASOBS;
  return run_search(f, Q, sp, occupied, assigned, f->n, n_matches);
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
  int rc;
  if ((rc = Q.bind(cur, n_last))) return rc;
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
  Q.descSrc = mp_desc; Q.hasObsSrc = has_obs; Q.filter = ORBFE_FILTER_UR;
  SearchSpec sp;
  sp.mode = ORBFE_MODE_LASTFRAME; sp.checkOri = check_orientation; sp.thAccept = 100; sp.feedback = ORBFE_FEEDBACK_HASOBS;
  return run_search(cur, Q, sp, occupied, assigned, cur->n, n_matches);
}

// ---- the other projection searches of OrbMatcher: same window-Hamming kernels, different gates ----

// OrbMatcher::SearchByProjection(KeyFrame*, cv::Mat Scw, vpPoints, vpMatched, th) (orb_matcher.cpp:384-497)
int orbfe_search_by_projection_sim3(orbfe_frame* kf, int n_mp, const uint8_t* valid, const float* u, const float* v,
                                    const int32_t* pred_level, const uint8_t* mp_desc, const uint8_t* matched_in, int th,
                                    int32_t* matched, int* n_matches) {
  if (!kf || !matched || n_mp < 0) return orbfe_fail(ORBFE_ERR_INVALID, "bad arguments");
  if (n_mp && (!valid || !u || !v || !pred_level || !mp_desc)) return orbfe_fail(ORBFE_ERR_INVALID, "null map-point array");
  if (kf->n && !matched_in) return orbfe_fail(ORBFE_ERR_INVALID, "null matched array");
  if (n_matches) *n_matches = 0;
  HostQueries Q;
  int rc;
  if ((rc = Q.bind(kf, n_mp))) return rc;
  for (int i = 0; i < n_mp; ++i) {
    if (!valid[i]) continue;
    const int lvl = pred_level[i];
    if (lvl < 0 || lvl >= kf->nlevels) return orbfe_fail(ORBFE_ERR_INVALID, "map point %d: predicted level %d out of range", i, lvl);
    Q.valid[i] = 1;
    Q.x[i] = u[i]; Q.y[i] = v[i];
    Q.r[i] = th * kf->scale[lvl];             // :454
    Q.minL[i] = lvl - 1; Q.maxL[i] = lvl;     // :474 (maxL >= 0, so the grid's level gate is exactly this test)
  }
  Q.descSrc = mp_desc;
  SearchSpec sp;  // bestDist<=TH_LOW (:488); vpMatched[bestIdx]=pMP occupies the keypoint for later points (:469, :490)
  sp.thAccept = 50; sp.feedback = ORBFE_FEEDBACK_ALL;
  return run_search(kf, Q, sp, matched_in, matched, kf->n, n_matches);
}

// OrbMatcher::SearchByProjection(Frame& Cur, KeyFrame*, sAlreadyFound, th, ORBdist) (orb_matcher.cpp:1455-1582)
int orbfe_search_by_projection_keyframe(orbfe_frame* cur, int n_kf, const uint8_t* valid, const float* u, const float* v,
                                        const int32_t* pred_level, const float* kf_angle, const uint8_t* mp_desc,
                                        const uint8_t* occupied, float th, int orb_dist, int check_orientation,
                                        int32_t* assigned, int* n_matches) {
  if (!cur || !assigned || n_kf < 0) return orbfe_fail(ORBFE_ERR_INVALID, "bad arguments");
  if (n_kf && (!valid || !u || !v || !pred_level || !kf_angle || !mp_desc)) return orbfe_fail(ORBFE_ERR_INVALID, "null keyframe array");
  if (cur->n && !occupied) return orbfe_fail(ORBFE_ERR_INVALID, "null occupied array");
  if (n_matches) *n_matches = 0;
  HostQueries Q;
  int rc;
  if ((rc = Q.bind(cur, n_kf))) return rc;
  for (int i = 0; i < n_kf; ++i) {
    if (!valid[i]) continue;
    if (u[i] < cur->minX || u[i] > cur->maxX) continue;   // :1490-1495
    if (v[i] < cur->minY || v[i] > cur->maxY) continue;
    const int lvl = pred_level[i];
    if (lvl < 0 || lvl >= cur->nlevels) return orbfe_fail(ORBFE_ERR_INVALID, "keyframe point %d: predicted level %d out of range", i, lvl);
    Q.valid[i] = 1;
    Q.x[i] = u[i]; Q.y[i] = v[i];
    Q.r[i] = th * cur->scale[lvl];                        // :1511
    Q.minL[i] = lvl - 1; Q.maxL[i] = lvl + 1;             // :1513
    Q.angle[i] = kf_angle[i];
  }
  Q.descSrc = mp_desc;
  SearchSpec sp;  // bestDist<=ORBdist (:1541); SetMapPoint occupies the keypoint (:1526, :1543)
  sp.thAccept = orb_dist; sp.feedback = ORBFE_FEEDBACK_ALL; sp.checkOri = check_orientation;
  return run_search(cur, Q, sp, occupied, assigned, cur->n, n_matches);
}

// the search part of both Fuse overloads: per map point the keypoint it would be fused into, else -1
static int fuse_core(orbfe_frame* kf, int n_mp, const uint8_t* valid, const float* u, const float* v, const float* ur,
                     const int32_t* pred_level, const uint8_t* mp_desc, float th, int32_t* best_idx, int* n_fused) {
  if (!kf || n_mp < 0 || (n_mp && !best_idx)) return orbfe_fail(ORBFE_ERR_INVALID, "bad arguments");
  if (n_mp && (!valid || !u || !v || !pred_level || !mp_desc)) return orbfe_fail(ORBFE_ERR_INVALID, "null map-point array");
  if (n_fused) *n_fused = 0;
  HostQueries Q;
  int rc;
  if ((rc = Q.bind(kf, n_mp))) return rc;
  for (int i = 0; i < n_mp; ++i) {
    if (!valid[i]) continue;
    const int lvl = pred_level[i];
    if (lvl < 0 || lvl >= kf->nlevels) return orbfe_fail(ORBFE_ERR_INVALID, "map point %d: predicted level %d out of range", i, lvl);
    Q.valid[i] = 1;
    Q.x[i] = u[i]; Q.y[i] = v[i];
    if (ur) Q.xr[i] = ur[i];
    Q.r[i] = th * kf->scale[lvl];             // :869, :1028
    Q.minL[i] = lvl - 1; Q.maxL[i] = lvl;     // :890, :1046
  }
  Q.descSrc = mp_desc;
  Q.filter = ur ? ORBFE_FILTER_FUSE : ORBFE_FILTER_NONE;
  SearchSpec sp;  // bestDist<=TH_LOW (:931, :1061); the scan reads no state another map point writes
  sp.thAccept = 50; sp.perQuery = 1;
  return run_search(kf, Q, sp, nullptr, best_idx, n_mp, n_fused);
}

// OrbMatcher::Fuse(KeyFrame*, const vector<MapPoint*>&, th) (orb_matcher.cpp:804-954)
int orbfe_fuse(orbfe_frame* kf, int n_mp, const uint8_t* valid, const float* u, const float* v, const float* ur,
               const int32_t* pred_level, const uint8_t* mp_desc, float th, int32_t* best_idx, int* n_fused) {
  if (n_mp && !ur) return orbfe_fail(ORBFE_ERR_INVALID, "null ur array");
  return fuse_core(kf, n_mp, valid, u, v, ur, pred_level, mp_desc, th, best_idx, n_fused);
}

// OrbMatcher::Fuse(KeyFrame*, cv::Mat Scw, vpPoints, th, vpReplacePoint) (orb_matcher.cpp:956-1079)
int orbfe_fuse_sim3(orbfe_frame* kf, int n_mp, const uint8_t* valid, const float* u, const float* v,
                    const int32_t* pred_level, const uint8_t* mp_desc, float th, int32_t* best_idx, int* n_fused) {
  return fuse_core(kf, n_mp, valid, u, v, nullptr, pred_level, mp_desc, th, best_idx, n_fused);
}

// OrbMatcher::SearchBySim3 (orb_matcher.cpp:1081-1310)
int orbfe_search_by_sim3(orbfe_frame* kf1, orbfe_frame* kf2, const uint8_t* valid1, const float* u1, const float* v1,
                         const int32_t* pred_level1, const uint8_t* mp_desc1, const uint8_t* valid2, const float* u2,
                         const float* v2, const int32_t* pred_level2, const uint8_t* mp_desc2, float th, int32_t* match12,
                         int* n_found) {
  if (!kf1 || !kf2) return orbfe_fail(ORBFE_ERR_INVALID, "null frame");
  if (kf1->device != kf2->device) return orbfe_fail(ORBFE_ERR_INVALID, "keyframes live on different devices");
  const int n1 = kf1->n, n2 = kf2->n;
  if ((n1 && (!valid1 || !u1 || !v1 || !pred_level1 || !mp_desc1 || !match12)) || (n2 && (!valid2 || !u2 || !v2 || !pred_level2 || !mp_desc2)))
    return orbfe_fail(ORBFE_ERR_INVALID, "null array");
  if (n_found) *n_found = 0;
  if (n1 == 0) return ORBFE_OK;
  SearchSpec sp;  // bestDist<=TH_HIGH (:1205, :1285), no coupling between map points
  sp.thAccept = 100; sp.perQuery = 1;
  // map points of KF1 searched in KF2 (:1132-1209), then map points of KF2 searched in KF1 (:1212-1289)
  for (int side = 0; side < 2; ++side) {
    orbfe_frame* dst = side == 0 ? kf2 : kf1;
    const int n = side == 0 ? n1 : n2;
    const uint8_t* valid = side == 0 ? valid1 : valid2;
    const float *u = side == 0 ? u1 : u2, *v = side == 0 ? v1 : v2;
    const int32_t* pl = side == 0 ? pred_level1 : pred_level2;
    HostQueries Q;
    int rc;
    if ((rc = Q.bind(dst, n))) return rc;
    for (int i = 0; i < n; ++i) {
      if (!valid[i]) continue;
      const int lvl = pl[i];
      if (lvl < 0 || lvl >= dst->nlevels) return orbfe_fail(ORBFE_ERR_INVALID, "map point %d: predicted level %d out of range", i, lvl);
      Q.valid[i] = 1;
      Q.x[i] = u[i]; Q.y[i] = v[i];
      Q.r[i] = th * dst->scale[lvl];          // :1173, :1253
      Q.minL[i] = lvl - 1; Q.maxL[i] = lvl;   // :1191, :1271
    }
    Q.descSrc = side == 0 ? mp_desc1 : mp_desc2;
    if ((rc = run_search(dst, Q, sp, nullptr, nullptr, n, nullptr))) return rc;
  }
  // agreement (:1291-1307): vnMatch1 = kf2->d_out (n1), vnMatch2 = kf1->d_out (n2); both streams are idle here
  CUDA_TRY(cudaSetDevice(kf1->device));
  CUDA_TRY(cudaMemsetAsync(kf1->d_cursor + 3, 0, sizeof(int), kf1->stream));
  MATCH_LAUNCH(kf1, k_sim3_agree, dim3((n1 + 255) / 256), dim3(256), 0, kf2->d_out, n1, kf1->d_out, n2, kf2->d_evIdx, kf1->d_cursor + 3);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(kf1->h_res + 3, kf1->d_cursor + 3, sizeof(int), cudaMemcpyDeviceToHost, kf1->stream));
  CUDA_TRY(cudaMemcpyAsync(match12, kf2->d_evIdx, (size_t)n1 * sizeof(int), cudaMemcpyDeviceToHost, kf1->stream));
  CUDA_TRY(cudaStreamSynchronize(kf1->stream));
  if (n_found) *n_found = kf1->h_res[3];
  return ORBFE_OK;
}

// ---- the vocabulary-node searches -----------------------------------------------------------------------------
// flattened DBoW2::FeatureVector (std::map<NodeId, std::vector<unsigned>>): node ids ascending, n_nodes+1 offsets
struct FeatVec {
  int nnodes;
  const uint32_t* ids;
  const int32_t* start;
  const uint32_t* idx;
};
struct BowHost {            // optional gates, see BowFilter (k_match.cuh)
  const uint8_t* valid2 = nullptr;
  int tri = 0, onlyStereo = 0;
  const orbfe_keypoint* kps1 = nullptr;
  const uint8_t* stereo1 = nullptr;
  const float* F12 = nullptr;
  float ex = 0, ey = 0;
};

// queries = the valid features of side 1 in the reference's visiting order (common nodes ascending -- the two-iterator
// walk with lower_bound is a sorted-set intersection -- then the node's own feature order); candidates of a query = the
// features of the same node on the searched frame `f`.  out: per searched keypoint (perQuery = 0) or per query.
static int bow_core(orbfe_frame* f, int n1, const uint8_t* desc1, const float* angle1, const uint8_t* valid1, const FeatVec& fv1,
                    const FeatVec& fv2, const SearchSpec& sp, const BowHost& bh, std::vector<int>& qFeat, std::vector<int>& result,
                    int* n_matches) {
  if (n_matches) *n_matches = 0;
  qFeat.clear();
  std::vector<int> qSrc, qCnt, qOff;
  std::vector<float> qAng, qX, qY;
  std::vector<uint8_t> qSt;
  int a = 0, b = 0, total = 0;
  while (a < fv1.nnodes && b < fv2.nnodes) {
    if (fv1.ids[a] == fv2.ids[b]) {
      const int fo = fv2.start[b], fc = fv2.start[b + 1] - fo;
      for (int k = fv1.start[a]; k < fv1.start[a + 1]; ++k) {
        const int real = (int)fv1.idx[k];
        if (real < 0 || real >= n1) return orbfe_fail(ORBFE_ERR_INVALID, "feature index %d out of range", real);
        if (!valid1[real]) continue;
        qFeat.push_back(real); qSrc.push_back(fo); qCnt.push_back(fc); qOff.push_back(total); qAng.push_back(angle1[real]);
        if (bh.tri) { qX.push_back(bh.kps1[real].x); qY.push_back(bh.kps1[real].y); qSt.push_back(bh.stereo1[real] ? 1 : 0); }
        total += fc;
      }
      ++a; ++b;
    } else if (fv1.ids[a] < fv2.ids[b]) ++a;
    else ++b;
  }
  const int nq = (int)qFeat.size();
  result.assign(sp.perQuery ? nq : f->n, -1);
  if (nq == 0 || f->n == 0) return ORBFE_OK;
  const int nfi = fv2.start[fv2.nnodes];
  for (int k = 0; k < nfi; ++k)
    if ((int)fv2.idx[k] < 0 || (int)fv2.idx[k] >= f->n) return orbfe_fail(ORBFE_ERR_INVALID, "searched-frame feature index out of range");
  CUDA_TRY(cudaSetDevice(f->device));
  int rc;
  // the packed block of this search: all side-1 descriptors, per query (descriptor row, source offset, list offset, list
  // length, angle[, kp1 x, y, stereo flag]), the searched frame's feature indices[, its validity flags]: one H2D copy
  const size_t Nq = (size_t)nq, N1 = (size_t)std::max(n1, 1), Nf = (size_t)std::max(nfi, 1), K = (size_t)f->n;
  const size_t need = al16(N1 * 32) + 5 * al16(Nq * 4) + al16(Nf * 4) + 2 * al16(Nq * 4) + al16(Nq) + al16(K) + 64;
  if ((rc = ensure_stage(f, need))) return rc;
  if ((rc = ensure_cand(f, total + 16))) return rc;
  size_t off = 0;
  uint8_t* hDesc = stage_take<uint8_t>(f, off, N1 * 32);
  int* hFeat = stage_take<int>(f, off, Nq);
  int* hSrc = stage_take<int>(f, off, Nq);
  int* hOff = stage_take<int>(f, off, Nq);
  int* hCnt = stage_take<int>(f, off, Nq);
  float* hAng = stage_take<float>(f, off, Nq);
  unsigned* hFi = stage_take<unsigned>(f, off, Nf);
  float* hX = stage_take<float>(f, off, Nq);
  float* hY = stage_take<float>(f, off, Nq);
  uint8_t* hSt = stage_take<uint8_t>(f, off, Nq);
  uint8_t* hV2 = stage_take<uint8_t>(f, off, K);
  memcpy(hDesc, desc1, (size_t)n1 * 32);
  memcpy(hFeat, qFeat.data(), Nq * 4); memcpy(hSrc, qSrc.data(), Nq * 4); memcpy(hOff, qOff.data(), Nq * 4);
  memcpy(hCnt, qCnt.data(), Nq * 4); memcpy(hAng, qAng.data(), Nq * 4);
  memcpy(hFi, fv2.idx, (size_t)nfi * 4);
  if (bh.tri) { memcpy(hX, qX.data(), Nq * 4); memcpy(hY, qY.data(), Nq * 4); memcpy(hSt, qSt.data(), Nq); }
  if (bh.valid2) memcpy(hV2, bh.valid2, K);
  CUDA_TRY(cudaMemcpyAsync(f->d_stage, f->h_stage, al16(off), cudaMemcpyHostToDevice, f->stream));
  SolveInput in;
  in.bow = 1; in.nq = nq;
  memset(&in.MQ, 0, sizeof(in.MQ));
  in.BQ.qDescAll = f->dev(hDesc); in.BQ.qDescIdx = f->dev(hFeat); in.BQ.featIdx = f->dev(hFi); in.BQ.qSrcOff = f->dev(hSrc); in.BQ.nQ = nq;
  in.qOff = f->dev(hOff); in.qCnt = f->dev(hCnt);
  BowFilter& B = in.B;
  B.valid2 = bh.valid2 ? f->dev(hV2) : nullptr; B.tri = bh.tri; B.onlyStereo = bh.onlyStereo; B.qx = f->dev(hX); B.qy = f->dev(hY);
  B.qStereo = f->dev(hSt); B.ex = bh.ex; B.ey = bh.ey;
  for (int k = 0; k < 9; ++k) B.F12[k] = bh.F12 ? bh.F12[k] : 0.f;
  in.hasObs = nullptr; in.occupied = nullptr; in.qAngle = f->dev(hAng);
  return solve(f, in, sp, result.data(), (int)result.size(), n_matches);
}

static int check_featvec(const FeatVec& fv) {
  if (fv.nnodes < 0 || (fv.nnodes && (!fv.ids || !fv.start || !fv.idx))) return orbfe_fail(ORBFE_ERR_INVALID, "bad feature-vector arrays");
  return ORBFE_OK;
}

// OrbMatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) (orb_matcher.cpp:133-262)
int orbfe_search_by_bow(orbfe_frame* f, int n_kf, const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_valid,
                        int kf_nnodes, const uint32_t* kf_node_ids, const int32_t* kf_node_start, const uint32_t* kf_feat_idx,
                        int f_nnodes, const uint32_t* f_node_ids, const int32_t* f_node_start, const uint32_t* f_feat_idx,
                        float nnratio, int check_orientation, int32_t* matched_kf_idx, int* n_matches) {
  if (!f || !matched_kf_idx || n_kf < 0) return orbfe_fail(ORBFE_ERR_INVALID, "bad arguments");
  if (n_kf && (!kf_desc || !kf_angle || !kf_valid)) return orbfe_fail(ORBFE_ERR_INVALID, "null keyframe array");
  const FeatVec fv1{kf_nnodes, kf_node_ids, kf_node_start, kf_feat_idx}, fv2{f_nnodes, f_node_ids, f_node_start, f_feat_idx};
  int rc;
  if ((rc = check_featvec(fv1)) || (rc = check_featvec(fv2))) return rc;
  SearchSpec sp;  // bestDist1<=TH_LOW && bestDist1 < mfNNratio*bestDist2 (:187-189); vpMapPointMatches[realIdxF] occupies (:173, :191)
  sp.mode = ORBFE_MODE_BOW; sp.nnratio = nnratio; sp.checkOri = check_orientation; sp.thAccept = 50; sp.ratio = ORBFE_RATIO_BOW;
  sp.feedback = ORBFE_FEEDBACK_ALL;
  std::vector<int> qFeat, assigned;
  if ((rc = bow_core(f, n_kf, kf_desc, kf_angle, kf_valid, fv1, fv2, sp, BowHost(), qFeat, assigned, n_matches))) return rc;
  for (int i = 0; i < f->n; ++i) matched_kf_idx[i] = assigned[i] >= 0 ? qFeat[assigned[i]] : -1;  // query -> keyframe feature
  return ORBFE_OK;
}

// OrbMatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&) (orb_matcher.cpp:499-632)
int orbfe_search_by_bow_keyframes(orbfe_frame* kf2, int n1, const uint8_t* desc1, const float* angle1, const uint8_t* valid1,
                                  const uint8_t* valid2, int nnodes1, const uint32_t* node_ids1, const int32_t* node_start1,
                                  const uint32_t* feat_idx1, int nnodes2, const uint32_t* node_ids2, const int32_t* node_start2,
                                  const uint32_t* feat_idx2, float nnratio, int check_orientation, int32_t* matches12,
                                  int* n_matches) {
  if (!kf2 || n1 < 0 || (n1 && !matches12)) return orbfe_fail(ORBFE_ERR_INVALID, "bad arguments");
  if (n1 && (!desc1 || !angle1 || !valid1)) return orbfe_fail(ORBFE_ERR_INVALID, "null keyframe-1 array");
  if (kf2->n && !valid2) return orbfe_fail(ORBFE_ERR_INVALID, "null keyframe-2 validity array");
  const FeatVec fv1{nnodes1, node_ids1, node_start1, feat_idx1}, fv2{nnodes2, node_ids2, node_start2, feat_idx2};
  int rc;
  if ((rc = check_featvec(fv1)) || (rc = check_featvec(fv2))) return rc;
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  SearchSpec sp;  // bestDist1<TH_LOW (strict, :575) && ratio (:577); vbMatched2 occupies (:553, :580)
  sp.mode = ORBFE_MODE_BOW; sp.nnratio = nnratio; sp.checkOri = check_orientation; sp.thAccept = 49; sp.ratio = ORBFE_RATIO_BOW;
  sp.feedback = ORBFE_FEEDBACK_ALL;
  BowHost bh;
  bh.valid2 = valid2;
  std::vector<int> qFeat, assigned;
  if ((rc = bow_core(kf2, n1, desc1, angle1, valid1, fv1, fv2, sp, bh, qFeat, assigned, n_matches))) return rc;
  // every KeyFrame-2 feature is taken at most once and every KeyFrame-1 feature is one query: invert
  for (int i2 = 0; i2 < kf2->n && i2 < (int)assigned.size(); ++i2)
    if (assigned[i2] >= 0) matches12[qFeat[assigned[i2]]] = i2;
  return ORBFE_OK;
}

// OrbMatcher::SearchForTriangulation (orb_matcher.cpp:634-802)
int orbfe_search_for_triangulation(orbfe_frame* kf2, int n1, const orbfe_keypoint* kps1_un, const uint8_t* desc1,
                                   const uint8_t* valid1, const uint8_t* stereo1, const uint8_t* valid2, int nnodes1,
                                   const uint32_t* node_ids1, const int32_t* node_start1, const uint32_t* feat_idx1,
                                   int nnodes2, const uint32_t* node_ids2, const int32_t* node_start2,
                                   const uint32_t* feat_idx2, const float* F12, float ex, float ey, int only_stereo,
                                   int check_orientation, int32_t* matches12, int* n_matches) {
  if (!kf2 || n1 < 0 || !F12 || (n1 && !matches12)) return orbfe_fail(ORBFE_ERR_INVALID, "bad arguments");
  if (n1 && (!kps1_un || !desc1 || !valid1 || !stereo1)) return orbfe_fail(ORBFE_ERR_INVALID, "null keyframe-1 array");
  if (kf2->n && !valid2) return orbfe_fail(ORBFE_ERR_INVALID, "null keyframe-2 validity array");
  const FeatVec fv1{nnodes1, node_ids1, node_start1, feat_idx1}, fv2{nnodes2, node_ids2, node_start2, feat_idx2};
  int rc;
  if ((rc = check_featvec(fv1)) || (rc = check_featvec(fv2))) return rc;
  for (int i = 0; i < n1; ++i) matches12[i] = -1;
  std::vector<uint8_t> v1(n1);
  std::vector<float> ang(n1);
  for (int i = 0; i < n1; ++i) {
    v1[i] = valid1[i] && !(only_stereo && !stereo1[i]);  // :681-688
    ang[i] = kps1_un[i].angle;
  }
  SearchSpec sp;  // dist<=TH_LOW, ties go to the LAST candidate ('dist>bestDist' skip, :717); this reference never sets
  sp.thAccept = 50; sp.tieLast = 1; sp.perQuery = 1; sp.checkOri = check_orientation;  // vbMatched2, so no coupling
  BowHost bh;
  bh.valid2 = valid2; bh.tri = 1; bh.onlyStereo = only_stereo; bh.kps1 = kps1_un; bh.stereo1 = stereo1; bh.F12 = F12;
  bh.ex = ex; bh.ey = ey;
  std::vector<int> qFeat, res;
  if ((rc = bow_core(kf2, n1, desc1, ang.data(), v1.data(), fv1, fv2, sp, bh, qFeat, res, n_matches))) return rc;
  for (size_t q = 0; q < qFeat.size() && q < res.size(); ++q) matches12[qFeat[q]] = res[q];
  return ORBFE_OK;
}


// Tracker::SearchLocalPoints (core/tracker.cpp:1196-1226): Frame::IsInFrustum over the candidate local map points
// (frame.cpp:277-337) chained into OrbMatcher::SearchByProjection(Frame&, vpMapPoints, th) (orb_matcher.cpp:13-111) on the
// device: the track_* arrays are produced and consumed in HBM.
int orbfe_search_local_points(orbfe_frame* f, int n, const float* world_pos, const float* normal, const float* min_dist,
                              const float* max_dist, const float* max_dist_raw, const float* Rcw, const float* tcw, const float* Ow,
                              float fx, float fy, float cx, float cy, float bf, float log_scale_factor, float viewing_cos_limit,
                              const uint8_t* mp_desc, const uint8_t* has_obs, const uint8_t* occupied, int th, float nnratio,
                              uint8_t* in_view, int32_t* scale_level, int32_t* assigned, int* n_in_view, int* n_matches) {
  if (!f || n < 0 || !assigned || !Rcw || !tcw || !Ow) return orbfe_fail(ORBFE_ERR_INVALID, "bad arguments");
  if (n && (!world_pos || !normal || !min_dist || !max_dist || !max_dist_raw || !mp_desc || !has_obs))
    return orbfe_fail(ORBFE_ERR_INVALID, "null map-point array");
  if (f->n && !occupied) return orbfe_fail(ORBFE_ERR_INVALID, "null occupied array");
  if (n_in_view) *n_in_view = 0;
  if (n_matches) *n_matches = 0;
  CUDA_TRY(cudaSetDevice(f->device));
  int rc;
  cudaStream_t st = f->stream;
  const size_t N = (size_t)std::max(n, 1);
  // inputs of the frustum test ride at the front of the query block (one H2D copy with the descriptors and the flags);
  // its outputs live in a device scratch block (grow-only, owned by the frame handle)
  const size_t front = 2 * al16(3 * N * 4) + 3 * al16(N * 4);
  HostQueries Q;
  if ((rc = Q.bind(f, n, front, true))) return rc;
  size_t off = 0;
  float* hW = stage_take<float>(f, off, 3 * N);
  float* hN = stage_take<float>(f, off, 3 * N);
  float* hMin = stage_take<float>(f, off, N);
  float* hMax = stage_take<float>(f, off, N);
  float* hRaw = stage_take<float>(f, off, N);
  const size_t need = 3 * N * sizeof(float) + N * sizeof(int) + N + 64;
  if (need > f->lpCap) {
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(regrow(&f->d_lp, need + need / 2));
    f->lpCap = need + need / 2;
  }
  float* d_px = reinterpret_cast<float*>(f->d_lp); float* d_py = d_px + N; float* d_vc = d_py + N;
  int* d_lvl = reinterpret_cast<int*>(d_vc + N);
  uint8_t* d_in = reinterpret_cast<uint8_t*>(d_lvl + N);
  FrustumArgs FA;
  for (int i = 0; i < 9; ++i) FA.R[i] = Rcw[i];
  for (int i = 0; i < 3; ++i) { FA.t[i] = tcw[i]; FA.Ow[i] = Ow[i]; }
  FA.fx = fx; FA.fy = fy; FA.cx = cx; FA.cy = cy; FA.bf = bf; FA.minX = f->minX; FA.maxX = f->maxX; FA.minY = f->minY; FA.maxY = f->maxY;
  FA.logScaleFactor = log_scale_factor; FA.viewingCosLimit = viewing_cos_limit; FA.nLevels = f->nlevels;
  // results of the frustum test land behind the search's own results in the pinned output block
  const size_t outBase = al16(16 + (size_t)std::max(f->n, 0) * sizeof(int));
  if ((rc = ensure_out_stage(f, outBase + al16(N * 4) + al16(N) + 16))) return rc;
  int* hLvl = reinterpret_cast<int*>(f->h_outStage + outBase);
  uint8_t* hIn = reinterpret_cast<uint8_t*>(f->h_outStage + outBase + al16(N * 4));
  int* hCnt = reinterpret_cast<int*>(f->h_outStage + outBase + al16(N * 4) + al16(N));
  if (n) {
    memcpy(hW, world_pos, (size_t)n * 12); memcpy(hN, normal, (size_t)n * 12);
    memcpy(hMin, min_dist, (size_t)n * 4); memcpy(hMax, max_dist, (size_t)n * 4); memcpy(hRaw, max_dist_raw, (size_t)n * 4);
    memcpy(Q.desc, mp_desc, (size_t)n * 32); memcpy(Q.hasObs, has_obs, (size_t)n);
  }
  if (f->n) memcpy(Q.occ, occupied, (size_t)f->n);
  CUDA_TRY(cudaMemcpyAsync(f->d_stage, f->h_stage, Q.headBytes, cudaMemcpyHostToDevice, st));
  *hCnt = 0;
  if (n) {
    CUDA_TRY(cudaMemsetAsync(f->d_cursor + 7, 0, sizeof(int), st));
    const dim3 grid((n + 255) / 256), block(256);
    // projected x_right lands directly in the query array the candidate kernel reads
    MATCH_LAUNCH(f, k_is_in_frustum, grid, block, 0, FA, n, f->dev(hW), f->dev(hN), f->dev(hMin), f->dev(hMax), f->dev(hRaw), d_in, d_px,
                 d_py, f->dev(Q.xr), d_lvl, d_vc, f->d_cursor + 7);
    MATCH_LAUNCH(f, k_frustum_to_queries, grid, block, 0, n, d_in, d_px, d_py, f->dev(Q.xr), d_lvl, d_vc, f->d_lvl, th, f->dev(Q.valid),
                 f->dev(Q.x), f->dev(Q.y), f->dev(Q.r), f->dev(Q.xr), f->dev(Q.minL), f->dev(Q.maxL));
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(hCnt, f->d_cursor + 7, sizeof(int), cudaMemcpyDeviceToHost, st));
    if (in_view) CUDA_TRY(cudaMemcpyAsync(hIn, d_in, (size_t)n, cudaMemcpyDeviceToHost, st));
    if (scale_level) CUDA_TRY(cudaMemcpyAsync(hLvl, d_lvl, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, st));
  }
  Q.hasObsSrc = has_obs; Q.filter = ORBFE_FILTER_UR;
  SearchSpec sp;
  sp.mode = ORBFE_MODE_MAPPOINTS; sp.nnratio = nnratio; sp.thAccept = 100; sp.ratio = ORBFE_RATIO_SAMELEVEL;
  sp.feedback = ORBFE_FEEDBACK_HASOBS;
  if ((rc = run_search(f, Q, sp, occupied, assigned, f->n, n_matches, true))) return rc;  // synchronises the stream
  if (n_in_view) *n_in_view = *hCnt;
  if (n && in_view) memcpy(in_view, hIn, (size_t)n);
  if (n && scale_level) memcpy(scale_level, hLvl, (size_t)n * sizeof(int));
  return ORBFE_OK;
}

}  // extern "C"
