// k_gray.cuh -- N4 (the step before the path): colour -> gray of the input frames, cv::cvtColor
// (CV_RGB2GRAY / CV_BGR2GRAY / CV_RGBA2GRAY / CV_BGRA2GRAY) as called by Tracker::GrabImageStereo /
// GrabImageMonocular (src/core/tracker.cpp:110-127, 179-196) before the Frame is built.
// OpenCV 4.x 8-bit fixed point: gray = (R*9798 + G*19235 + B*3735 + 16384) >> 15.
// The colour frames are staged tightly packed, so both arrays are flat: one thread = 4 consecutive gray
// pixels = one aligned output word, read from 3 (RGB) or 4 (RGBA) aligned input words.
#pragma once
#include "orbfe_common.cuh"

#define ORBFE_GRAY_THREADS 256

__device__ __forceinline__ unsigned orbfe_gray1(unsigned r, unsigned g, unsigned b) {
  return (r * 9798u + g * 19235u + b * 3735u + 16384u) >> 15;
}

// src: n frames of nWords*4 pixels * channels bytes each (frame stride srcStride bytes, a multiple of 4);
// dst: gray frames (frame stride dstStride bytes, a multiple of 4).  nWords = ceil(w*h / 4).
__global__ void __launch_bounds__(ORBFE_GRAY_THREADS)
k_gray(const uint8_t* __restrict__ src, const size_t srcStride, uint8_t* __restrict__ dst, const size_t dstStride,
       const int nWords, const int channels, const int rgbOrder) {
  const int t = blockIdx.x * ORBFE_GRAY_THREADS + threadIdx.x;
  if (t >= nWords) return;
  const unsigned* s = reinterpret_cast<const unsigned*>(src + (size_t)blockIdx.y * srcStride);
  unsigned out = 0;
  if (channels == 4) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned p = __ldg(s + 4 * (size_t)t + j);
      const unsigned c0 = p & 0xffu, c1 = (p >> 8) & 0xffu, c2 = (p >> 16) & 0xffu;
      out |= orbfe_gray1(rgbOrder ? c0 : c2, c1, rgbOrder ? c2 : c0) << (8 * j);
    }
  } else {
    const unsigned w0 = __ldg(s + 3 * (size_t)t), w1 = __ldg(s + 3 * (size_t)t + 1), w2 = __ldg(s + 3 * (size_t)t + 2);
    // 12 bytes = 4 pixels: (w0.b0 w0.b1 w0.b2) (w0.b3 w1.b0 w1.b1) (w1.b2 w1.b3 w2.b0) (w2.b1 w2.b2 w2.b3)
    const unsigned px[4] = {w0 & 0xffffffu, __funnelshift_r(w0, w1, 24) & 0xffffffu, __funnelshift_r(w1, w2, 16) & 0xffffffu, w2 >> 8};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const unsigned c0 = px[j] & 0xffu, c1 = (px[j] >> 8) & 0xffu, c2 = px[j] >> 16;
      out |= orbfe_gray1(rgbOrder ? c0 : c2, c1, rgbOrder ? c2 : c0) << (8 * j);
    }
  }
  reinterpret_cast<unsigned*>(dst + (size_t)blockIdx.y * dstStride)[t] = out;
}
