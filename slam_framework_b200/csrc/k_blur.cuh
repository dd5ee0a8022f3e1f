// k_blur.cuh -- E6: 7x7 Gaussian blur, sigma 2, BORDER_REFLECT_101, on the un-padded level
// (cv::GaussianBlur on a clone of the level, orb_extractor.cpp:1029-1030).
// OpenCV >= 3.4.1 fixed-point path: taps [18,34,48,56,48,34,18]/256 per
// axis, u16 horizontal sums (<= 65280), one final rounding (V + 32768) >> 16.
// One CTA = one 64x32 output tile of one level of one image slot; the (64+6)x(32+6) source
// window is staged in shared memory with the reflection applied at the level edges, the
// horizontal pass writes u16 rows to shared memory, the vertical pass writes 4-pixel words.
#pragma once
#include "orbfe_common.cuh"

#define ORBFE_BLUR_THREADS 256
#define ORBFE_BLUR_TW 64
#define ORBFE_BLUR_TH 32

__global__ void __launch_bounds__(ORBFE_BLUR_THREADS)
k_blur(const __grid_constant__ Geom g, const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur) {
  __shared__ uint8_t s_src[ORBFE_BLUR_TH + 6][ORBFE_BLUR_TW + 8];
  __shared__ uint16_t s_h[ORBFE_BLUR_TH + 6][ORBFE_BLUR_TW];
  const int slot = blockIdx.y;
  const int tile = blockIdx.x;
  int level = 0;
  for (int l = 1; l < g.nlevels; ++l)
    if (tile >= g.lv[l].tileBase) level = l;
  const LevelGeom& L = g.lv[level];
  const int ti = tile - L.tileBase;
  const int ty = ti / L.tilesX, tx = ti - ty * L.tilesX;
  const int x0 = tx * ORBFE_BLUR_TW, y0 = ty * ORBFE_BLUR_TH;
  const uint8_t* src = pyr + (size_t)slot * g.pyrStride + L.planeOff + (size_t)ORBFE_EDGE * L.pitch + ORBFE_EDGE;
  uint8_t* dst = blur + (size_t)slot * g.blurStride + L.blurOff;
  for (int t = threadIdx.x; t < (ORBFE_BLUR_TH + 6) * (ORBFE_BLUR_TW + 6); t += ORBFE_BLUR_THREADS) {
    const int r = t / (ORBFE_BLUR_TW + 6), c = t - r * (ORBFE_BLUR_TW + 6);
    const int sy = orbfe_reflect101(min(y0 + r - 3, L.h + 2), L.h);
    const int sx = orbfe_reflect101(min(x0 + c - 3, L.w + 2), L.w);
    s_src[r][c] = __ldg(src + (size_t)sy * L.pitch + sx);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < (ORBFE_BLUR_TH + 6) * ORBFE_BLUR_TW; t += ORBFE_BLUR_THREADS) {
    const int r = t / ORBFE_BLUR_TW, c = t - r * ORBFE_BLUR_TW;
    const uint8_t* p = &s_src[r][c];
    s_h[r][c] = (uint16_t)(18 * (p[0] + p[6]) + 34 * (p[1] + p[5]) + 48 * (p[2] + p[4]) + 56 * p[3]);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < ORBFE_BLUR_TH * (ORBFE_BLUR_TW / 4); t += ORBFE_BLUR_THREADS) {
    const int r = t / (ORBFE_BLUR_TW / 4), c4 = (t - r * (ORBFE_BLUR_TW / 4)) * 4;
    const int y = y0 + r, x = x0 + c4;
    if (y >= L.h || x >= L.w) continue;
    unsigned w = 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int c = c4 + b;
      const unsigned acc = 18u * ((unsigned)s_h[r][c] + s_h[r + 6][c]) + 34u * ((unsigned)s_h[r + 1][c] + s_h[r + 5][c]) +
                           48u * ((unsigned)s_h[r + 2][c] + s_h[r + 4][c]) + 56u * (unsigned)s_h[r + 3][c];
      w |= ((acc + 32768u) >> 16) << (8 * b);
    }
    // bpitch is a multiple of 16, so the word store is aligned; bytes past w land in row padding
    *reinterpret_cast<unsigned*>(dst + (size_t)y * L.bpitch + x) = w;
  }
}
