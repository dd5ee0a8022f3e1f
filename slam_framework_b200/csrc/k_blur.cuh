// k_blur.cuh -- E6: 7x7 Gaussian blur, sigma 2, BORDER_REFLECT_101, on the un-padded level
// (cv::GaussianBlur on a clone of the level, orb_extractor.cpp:1029-1030).
// OpenCV >= 3.4.1 fixed-point path: taps [18,34,48,56,48,34,18]/256 per
// axis, u16 horizontal sums (<= 65280), one final rounding (V + 32768) >> 16.
//
// The source is the PADDED pyramid plane: its 19-px border already holds the level's
// BORDER_REFLECT_101 extension (orb_extractor.cpp:1066-1071), which is exactly the extension
// cv::GaussianBlur applies to the clone, so no reflection logic is needed; and because the ROI
// starts at column 19, output word k (pixels 4k..4k+3) reads input bytes 16+4k..25+4k = the three
// ALIGNED words k+4..k+6 of the padded row.
//
// One warp = a strip of 30 output words (120 px) x ORBFE_BLUR_ROWS rows.  Lanes load one aligned
// word per input row (coalesced 128 B), fetch the two following words with shuffles, form the 7-tap
// horizontal sums with byte funnel-shifts + IDP.4A (2 dp4a per pixel), keep the last 7 rows of sums
// in registers (paired row-wise, so that the column sum is 3 IDP.2A + 1 IMAD) and emit one 4-pixel word per row: no shared
// memory, every byte read once per strip.
#pragma once
#include "orbfe_common.cuh"

#define ORBFE_BLUR_THREADS 128
#define ORBFE_BLUR_WORDS 30   // output words per warp strip (32 loaded - 2 for the right neighbours)
#ifndef ORBFE_BLUR_ROWS
#define ORBFE_BLUR_ROWS 64    // output rows per warp strip (6 halo rows re-read per strip)
#endif
// legacy tile macros (geometry fields tilesX/tilesY/tileBase now count warp strips)
#define ORBFE_BLUR_TW (4 * ORBFE_BLUR_WORDS)
#define ORBFE_BLUR_TH ORBFE_BLUR_ROWS

__device__ __forceinline__ unsigned orbfe_dp4a_u8(unsigned a, unsigned b, unsigned c) { return __dp4a(a, b, c); }

#ifndef ORBFE_BLUR_MINB
#define ORBFE_BLUR_MINB 1
#endif
__global__ void __launch_bounds__(ORBFE_BLUR_THREADS, ORBFE_BLUR_MINB)
k_blur(const __grid_constant__ Geom g, const uint8_t* __restrict__ pyr, uint8_t* __restrict__ blur) {
  const int slot = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int task = blockIdx.x * (ORBFE_BLUR_THREADS / 32) + (threadIdx.x >> 5);
  if (task >= g.totalTiles) return;
  int level = 0;
  for (int l = 1; l < g.nlevels; ++l)
    if (task >= g.lv[l].tileBase) level = l;
  const LevelGeom& L = g.lv[level];
  const int ti = task - L.tileBase;
  const int ty = ti / L.tilesX, tx = ti - ty * L.tilesX;
  const int k = tx * ORBFE_BLUR_WORDS + lane;  // output word of this lane (lanes 30,31 only feed neighbours)
  const int y0 = ty * ORBFE_BLUR_ROWS;
  const int nrows = min(ORBFE_BLUR_ROWS, L.h - y0);
  const int pitchW = L.pitch >> 2;
  // input word k+4 of padded row (19 + y - 3); clamp the column so that every lane reads inside the plane
  const int inW = min(k + 4, pitchW - 1);
  const unsigned* src = reinterpret_cast<const unsigned*>(pyr + (size_t)slot * g.pyrStride + L.planeOff) +
                        (size_t)(ORBFE_EDGE - 3 + y0) * pitchW + inW;
  uint8_t* dst = blur + (size_t)slot * g.blurStride + L.blurOff + (size_t)y0 * L.bpitch + 4 * k;
  const bool writer = lane < ORBFE_BLUR_WORDS && 4 * k < L.w;
  const unsigned KA = 0x38302212u;  // taps 18,34,48,56 (bytes 0..3)
  const unsigned KB = 0x00122230u;  // taps 48,34,18,0
  // Vertical pass: consecutive rows' horizontal sums (<= 65280, 16 bits) are paired in one register, P[slot of row a] =
  // H[a] | H[a+1] << 16, so that the 7-tap column sum is 3 IDP.2A + 1 IMAD instead of 3 adds + 4 IMADs per pixel.
  const unsigned W01 = 18u | (34u << 8), W23 = 48u | (56u << 8), W45 = 48u | (34u << 8);
  unsigned P[7][4], Hp[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) Hp[j] = 0u;
#pragma unroll
  for (int s = 0; s < 7; ++s)
#pragma unroll
    for (int j = 0; j < 4; ++j) P[s][j] = 0u;
  const int total = nrows + 6;
  for (int r0 = 0; r0 < total; r0 += 7) {
    unsigned wl[7];  // the 7 row loads of this group are issued back to back (latency overlap)
#pragma unroll
    for (int s = 0; s < 7; ++s) wl[s] = r0 + s < total ? __ldg(src + (size_t)(r0 + s) * pitchW) : 0u;
#pragma unroll
    for (int s = 0; s < 7; ++s) {
      const int r = r0 + s;
      if (r < total) {  // warp-uniform
        const unsigned w0 = wl[s];
        const unsigned w1 = __shfl_down_sync(0xffffffffu, w0, 1);
        const unsigned w2 = __shfl_down_sync(0xffffffffu, w0, 2);
        // horizontal sums of the 4 pixels of this word: bytes j..j+6 of (w0,w1,w2)
        unsigned Hn[4];
        Hn[0] = orbfe_dp4a_u8(w1, KB, orbfe_dp4a_u8(w0, KA, 0u));
        Hn[1] = orbfe_dp4a_u8(__funnelshift_r(w1, w2, 8), KB, orbfe_dp4a_u8(__funnelshift_r(w0, w1, 8), KA, 0u));
        Hn[2] = orbfe_dp4a_u8(__funnelshift_r(w1, w2, 16), KB, orbfe_dp4a_u8(__funnelshift_r(w0, w1, 16), KA, 0u));
        Hn[3] = orbfe_dp4a_u8(__funnelshift_r(w1, w2, 24), KB, orbfe_dp4a_u8(__funnelshift_r(w0, w1, 24), KA, 0u));
#pragma unroll
        for (int j = 0; j < 4; ++j) { P[(s + 6) % 7][j] = Hn[j] * 65536u + Hp[j]; Hp[j] = Hn[j]; }  // pair (r-1, r)
        if (r >= 6 && writer) {
          // pairs (r-6,r-5), (r-4,r-3), (r-2,r-1) live in slots s+1, s+3, s+5 (mod 7); row r itself is Hn
          unsigned acc[4];  // V + 32768 < 2^24: the rounded result is byte 2 of the accumulator
#pragma unroll
          for (int j = 0; j < 4; ++j)
            acc[j] = __dp2a_lo(P[(s + 1) % 7][j], W01, __dp2a_lo(P[(s + 3) % 7][j], W23,
                     __dp2a_lo(P[(s + 5) % 7][j], W45, 18u * Hn[j] + 32768u)));
          const unsigned out = __byte_perm(__byte_perm(acc[0], acc[1], 0x0062), __byte_perm(acc[2], acc[3], 0x0062), 0x5410);
          // bpitch is a multiple of 16: the word store is aligned; bytes past w land in row padding
          *reinterpret_cast<unsigned*>(dst + (size_t)(r - 6) * L.bpitch) = out;
        }
      }
    }
  }
}
