// k_pyramid.cuh -- E1: scale pyramid (ORBextractor::ComputePyramid, orb_extractor.cpp:1051-1076).
// One launch per level (level l is resized from level l-1: the chain is the reference's), each
// launch covering every image slot of the batch.  A thread produces one aligned 4-pixel word of
// the PADDED plane; border pixels are produced by evaluating the interior formula at the
// BORDER_REFLECT_101 source coordinate, so there is no second border pass and no intra-kernel
// dependency (cv::copyMakeBorder(.., REFLECT_101 [+ISOLATED]), orb_extractor.cpp:1066,1071).
// Resize arithmetic = cv::resize INTER_LINEAR u8: 11-bit coefficients from a host-built LUT,
// vertical blend ((b0*(T0>>4))>>16 + (b1*(T1>>4))>>16 + 2)>>2.
#pragma once
#include "orbfe_common.cuh"

#define ORBFE_PYR_THREADS 256

__device__ __forceinline__ int orbfe_resize_px(const uint8_t* __restrict__ src, int spitch, int sw, int sh,
                                               const ResizeLut lx, const ResizeLut ly) {
  const int sx0 = lx.ofs, sx1 = min(lx.ofs + 1, sw - 1);
  const int sy0 = ly.ofs, sy1 = min(ly.ofs + 1, sh - 1);
  const uint8_t* r0 = src + (size_t)sy0 * spitch;
  const uint8_t* r1 = src + (size_t)sy1 * spitch;
  const int t0 = (int)__ldg(r0 + sx0) * lx.c0 + (int)__ldg(r0 + sx1) * lx.c1;
  const int t1 = (int)__ldg(r1 + sx0) * lx.c0 + (int)__ldg(r1 + sx1) * lx.c1;
  return ((((int)ly.c0 * (t0 >> 4)) >> 16) + (((int)ly.c1 * (t1 >> 4)) >> 16) + 2) >> 2;
}

__global__ void __launch_bounds__(ORBFE_PYR_THREADS)
k_pyramid_level(const __grid_constant__ Geom g, const int level, const uint8_t* __restrict__ img,
                uint8_t* __restrict__ pyr, const ResizeLut* __restrict__ lut) {
  const LevelGeom& L = g.lv[level];
  const int slot = blockIdx.y;
  const int word = blockIdx.x * ORBFE_PYR_THREADS + threadIdx.x;
  const int pw = L.w + 2 * ORBFE_EDGE, ph = L.h + 2 * ORBFE_EDGE;
  const int py = word / L.pyrWords;
  const int wx = word - py * L.pyrWords;
  if (py >= ph) return;
  uint8_t* plane = pyr + (size_t)slot * g.pyrStride + L.planeOff;
  const int y = orbfe_reflect101(py - ORBFE_EDGE, L.h);
  unsigned out = 0;
  if (level == 0) {
    const uint8_t* src = img + (size_t)slot * g.imgStride + (size_t)y * g.imgPitch;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int px = 4 * wx + b;
      const int x = orbfe_reflect101(min(px, pw - 1) - ORBFE_EDGE, L.w);
      out |= (unsigned)__ldg(src + x) << (8 * b);
    }
  } else {
    const LevelGeom& P = g.lv[level - 1];
    const uint8_t* src = pyr + (size_t)slot * g.pyrStride + P.planeOff + (size_t)ORBFE_EDGE * P.pitch + ORBFE_EDGE;
    if (L.area2) {
      const uint8_t* r0 = src + (size_t)(2 * y) * P.pitch;
      const uint8_t* r1 = r0 + P.pitch;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int px = 4 * wx + b;
        const int x = orbfe_reflect101(min(px, pw - 1) - ORBFE_EDGE, L.w);
        const int v = ((int)__ldg(r0 + 2 * x) + (int)__ldg(r0 + 2 * x + 1) + (int)__ldg(r1 + 2 * x) +
                       (int)__ldg(r1 + 2 * x + 1) + 2) >> 2;
        out |= (unsigned)v << (8 * b);
      }
    } else {
      const ResizeLut ly = lut[L.lutYOff + y];
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int px = 4 * wx + b;
        const int x = orbfe_reflect101(min(px, pw - 1) - ORBFE_EDGE, L.w);
        const ResizeLut lx = lut[L.lutXOff + x];
        out |= (unsigned)orbfe_resize_px(src, P.pitch, P.w, P.h, lx, ly) << (8 * b);
      }
    }
  }
  *reinterpret_cast<unsigned*>(plane + (size_t)py * L.pitch + 4 * wx) = out;
}
