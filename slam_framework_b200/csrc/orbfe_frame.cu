// orbfe_frame.cu -- host side of the Frame-tail entry points of include/orbfe.h (kernels in k_frame.cuh): N2 of SURVEY 8f.
#include "../../include/orbfe.h"

#include "k_frame.cuh"
#include "orbfe_host.h"

#include <vector>

#define CUDA_TRY(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return orbfe_fail(ORBFE_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

#ifdef ORBFE_EMU
#define FRAME_LAUNCH(kernel, grid, block, ...) emu::launch(grid, block, 0, [&]() { kernel(__VA_ARGS__); })
#else
#define FRAME_LAUNCH(kernel, grid, block, ...) kernel<<<grid, block>>>(__VA_ARGS__)
#endif

static int check_device(int device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess) { cudaGetLastError(); ndev = 0; }
  if (device < 0 || device >= ndev)
    return orbfe_fail(ORBFE_ERR_CUDA, "CUDA device %d not available (%d visible); this library has no CPU path", device, ndev);
  return ORBFE_OK;
}

// Per-thread, per-device scratch block carved into aligned pieces.  These entry points are stateless calls made once per frame:
// a cudaMalloc / cudaFree pair per call would cost more than the kernels, so the block is kept and only ever grows (a thread that
// switches devices re-allocates).  Calls are synchronous, so the block is free again when they return.
struct ScratchCache {
  int device = -1;
  char* base = nullptr;
  size_t cap = 0;
  char* pinned = nullptr;   // host staging: a call packs its inputs here (ONE H2D copy) and lands its outputs here (ONE D2H copy)
  size_t pinnedCap = 0;
  ~ScratchCache() { /* the CUDA context may be gone at thread exit: the block is left to process teardown */ }
};
static thread_local ScratchCache t_scratch;

struct DeviceArena {
  char* base = nullptr;
  size_t used = 0, cap = 0;
  static size_t pad(size_t b) { return (b + 255) & ~(size_t)255; }
  template <class T> T* take(size_t count) { T* p = reinterpret_cast<T*>(base + used); used += pad(count * sizeof(T)); return p; }
  // binds the arena to the calling thread's cached block of `device`, growing it to `bytes`
  cudaError_t acquire(int device, size_t bytes) {
    ScratchCache& c = t_scratch;
    if (c.device != device || c.cap < bytes) {
      if (c.base && c.device == device) cudaFree(c.base);
      c.base = nullptr; c.cap = 0; c.device = device;
      const size_t want = bytes + bytes / 2 + 4096;
      const cudaError_t e = cudaMalloc(&c.base, want);
      if (e != cudaSuccess) { c.base = nullptr; return e; }
      c.cap = want;
    }
    base = c.base; cap = c.cap; used = 0;
    return cudaSuccess;
  }
  // the calling thread's pinned staging block, at least `bytes` long
  static cudaError_t pinned(size_t bytes, char** out) {
    ScratchCache& c = t_scratch;
    if (c.pinnedCap < bytes) {
      if (c.pinned) cudaFreeHost(c.pinned);
      c.pinned = nullptr; c.pinnedCap = 0;
      const size_t want = bytes + bytes / 2 + 4096;
      const cudaError_t e = cudaMallocHost(&c.pinned, want);
      if (e != cudaSuccess) { c.pinned = nullptr; return e; }
      c.pinnedCap = want;
    }
    *out = c.pinned;
    return cudaSuccess;
  }
};

extern "C" {

int orbfe_undistort_keypoints(int device, int n, const orbfe_keypoint* kps, float fx, float fy, float cx, float cy,
                              const float* dist_coeffs, int n_dist, orbfe_keypoint* kps_un) {
  if (n < 0 || (n && (!kps || !kps_un)) || n_dist < 0 || n_dist > 14 || (n_dist && !dist_coeffs))
    return orbfe_fail(ORBFE_ERR_INVALID, "bad arguments");
  if (n_dist > 12 && (dist_coeffs[12] != 0.0f || (n_dist > 13 && dist_coeffs[13] != 0.0f)))
    return orbfe_fail(ORBFE_ERR_INVALID, "tilted sensor model (tauX, tauY) is not supported");
  int rc;
  if ((rc = check_device(device))) return rc;
  if (n == 0) return ORBFE_OK;
  CUDA_TRY(cudaSetDevice(device));
  if (n_dist == 0 || dist_coeffs[0] == 0.0f) {  // frame.cpp:616-619: undistorted_keypoints_ = keypoints_
    if (kps_un != kps) memcpy(kps_un, kps, (size_t)n * sizeof(orbfe_keypoint));
    return ORBFE_OK;
  }
  UndistortArgs U;
  U.fx = fx; U.fy = fy; U.cx = cx; U.cy = cy; U.ifx = 1. / U.fx; U.ify = 1. / U.fy;
  for (int i = 0; i < 14; ++i) U.k[i] = i < n_dist ? (double)dist_coeffs[i] : 0.0;
  static_assert(sizeof(orbfe_keypoint) == 28, "keypoint layout");
  DeviceArena A;
  CUDA_TRY(A.acquire(device, 2 * DeviceArena::pad((size_t)n * sizeof(orbfe_keypoint))));
  float* d_in = A.take<float>((size_t)n * 7);
  float* d_out = A.take<float>((size_t)n * 7);
  CUDA_TRY(cudaMemcpy(d_in, kps, (size_t)n * sizeof(orbfe_keypoint), cudaMemcpyHostToDevice));
  FRAME_LAUNCH(k_undistort_points, dim3((n + 255) / 256), dim3(256), U, n, d_in, d_out, 7);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(kps_un, d_out, (size_t)n * sizeof(orbfe_keypoint), cudaMemcpyDeviceToHost));
  return ORBFE_OK;
}

int orbfe_is_in_frustum(int device, int n, const float* world_pos, const float* normal, const float* min_dist,
                        const float* max_dist, const float* max_dist_raw, const float* Rcw, const float* tcw, const float* Ow,
                        float fx, float fy, float cx, float cy, float bf, float min_x, float max_x, float min_y, float max_y,
                        float log_scale_factor, int n_levels, float viewing_cos_limit, uint8_t* in_view, float* proj_x,
                        float* proj_y, float* proj_xr, int32_t* scale_level, float* view_cos, int* n_in_view) {
  if (n < 0 || !Rcw || !tcw || !Ow || n_levels < 1) return orbfe_fail(ORBFE_ERR_INVALID, "bad arguments");
  if (n && (!world_pos || !normal || !min_dist || !max_dist || !max_dist_raw || !in_view || !proj_x || !proj_y || !proj_xr || !scale_level || !view_cos))
    return orbfe_fail(ORBFE_ERR_INVALID, "null array");
  if (n_in_view) *n_in_view = 0;
  int rc;
  if ((rc = check_device(device))) return rc;
  if (n == 0) return ORBFE_OK;
  CUDA_TRY(cudaSetDevice(device));
  FrustumArgs F;
  for (int i = 0; i < 9; ++i) F.R[i] = Rcw[i];
  for (int i = 0; i < 3; ++i) { F.t[i] = tcw[i]; F.Ow[i] = Ow[i]; }
  F.fx = fx; F.fy = fy; F.cx = cx; F.cy = cy; F.bf = bf; F.minX = min_x; F.maxX = max_x; F.minY = min_y; F.maxY = max_y;
  F.logScaleFactor = log_scale_factor; F.viewingCosLimit = viewing_cos_limit; F.nLevels = n_levels;
  const size_t N = (size_t)n;
  // inputs: world (3N) | normal (3N) | min | max | raw floats; outputs: px | py | pxr | view_cos floats | level ints | count | in_view bytes.
  // Both blocks are contiguous on the device and mirrored in pinned host memory: one copy each way instead of twelve.
  const size_t inBytes = 9 * N * 4, outBytes = 5 * N * 4 + 16 + N;
  DeviceArena A;
  CUDA_TRY(A.acquire(device, DeviceArena::pad(inBytes) + DeviceArena::pad(outBytes) + 256));
  char* h = nullptr;
  CUDA_TRY(DeviceArena::pinned(DeviceArena::pad(inBytes) + outBytes, &h));
  float* d_w = A.take<float>(9 * N);
  float* d_n = d_w + 3 * N; float* d_min = d_n + 3 * N; float* d_max = d_min + N; float* d_raw = d_max + N;
  char* d_outBlock = A.take<char>(outBytes);
  float* d_px = reinterpret_cast<float*>(d_outBlock); float* d_py = d_px + N; float* d_pxr = d_py + N; float* d_vc = d_pxr + N;
  int* d_lvl = reinterpret_cast<int*>(d_vc + N);
  int* d_cnt = d_lvl + N;
  uint8_t* d_in = reinterpret_cast<uint8_t*>(d_cnt + 4);
  float* hIn = reinterpret_cast<float*>(h);
  memcpy(hIn, world_pos, N * 12); memcpy(hIn + 3 * N, normal, N * 12);
  memcpy(hIn + 6 * N, min_dist, N * 4); memcpy(hIn + 7 * N, max_dist, N * 4); memcpy(hIn + 8 * N, max_dist_raw, N * 4);
  CUDA_TRY(cudaMemcpy(d_w, hIn, inBytes, cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMemset(d_cnt, 0, sizeof(int)));
  FRAME_LAUNCH(k_is_in_frustum, dim3((n + 255) / 256), dim3(256), F, n, d_w, d_n, d_min, d_max, d_raw, d_in, d_px, d_py, d_pxr, d_lvl,
               d_vc, d_cnt);
  CUDA_TRY(cudaGetLastError());
  char* hOut = h + DeviceArena::pad(inBytes);
  CUDA_TRY(cudaMemcpy(hOut, d_outBlock, outBytes, cudaMemcpyDeviceToHost));
  memcpy(proj_x, hOut, N * 4); memcpy(proj_y, hOut + N * 4, N * 4); memcpy(proj_xr, hOut + 2 * N * 4, N * 4);
  memcpy(view_cos, hOut + 3 * N * 4, N * 4); memcpy(scale_level, hOut + 4 * N * 4, N * 4);
  int cnt = 0;
  memcpy(&cnt, hOut + 5 * N * 4, sizeof(int));
  memcpy(in_view, hOut + 5 * N * 4 + 16, N);
  if (n_in_view) *n_in_view = cnt;
  return ORBFE_OK;
}

int orbfe_debug_logf(int device, int n, const float* x, float* y) {
  if (n < 0 || (n && (!x || !y))) return orbfe_fail(ORBFE_ERR_INVALID, "bad arguments");
  int rc;
  if ((rc = check_device(device))) return rc;
  if (n == 0) return ORBFE_OK;
  CUDA_TRY(cudaSetDevice(device));
  DeviceArena A;
  CUDA_TRY(A.acquire(device, 2 * DeviceArena::pad((size_t)n * 4)));
  float* d_x = A.take<float>(n);
  float* d_y = A.take<float>(n);
  CUDA_TRY(cudaMemcpy(d_x, x, (size_t)n * 4, cudaMemcpyHostToDevice));
  FRAME_LAUNCH(k_debug_logf, dim3((n + 255) / 256), dim3(256), d_x, n, d_y);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpy(y, d_y, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return ORBFE_OK;
}

}  // extern "C"
