// orbfe_host.h -- host-side helpers shared by the translation units of liborbfe.so.
#pragma once
#include <cstdarg>

// records a thread-local message for orbfe_last_error() and returns `code`
int orbfe_fail(int code, const char* fmt, ...);

// device-resident view of one slot of an extractor handle (orbfe_api.cu), for entry points of other translation units that
// consume extraction results without a host round trip (orbfe_frame_from_extractor)
struct orbfe_extractor;
struct OrbfeSlotView {
  int device;
  void* stream;          // cudaStream_t of the extractor handle
  const void* kps;       // orbfe_kp_dev[capacity] (28-byte cv::KeyPoint records)
  const unsigned char* desc;
  const float* uR;       // stereo right coordinates of the slot (valid after orbfe_run_stereo on a left slot), may be null
  const int* nKp;        // device: keypoint count of the slot
  int capacity, nlevels, w, h;
  float scale[16];
};
int orbfe_internal_slot_view(orbfe_extractor* ex, int slot, OrbfeSlotView* out);

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device property of a kernel, shared by every handle and thread of the
// process: it is only ever RAISED (a handle with a smaller geometry must not lower the limit a live handle launches with), under
// a lock, and the call is skipped when the recorded maximum already covers the request.
#ifndef ORBFE_EMU
#include <cuda_runtime.h>
#include <map>
#include <mutex>
#include <utility>
template <class K>
static inline cudaError_t orbfe_raise_dynamic_smem(K kernel, int device, size_t bytes) {
  static std::mutex m;
  static std::map<std::pair<const void*, int>, size_t> have;
  std::lock_guard<std::mutex> lock(m);
  size_t& cur = have[std::make_pair((const void*)kernel, device)];
  if (bytes <= cur) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) cur = bytes;
  return e;
}
#endif
