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
#include "orbfe_tma.cuh"

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
    const size_t rowByte = (size_t)slot * g.imgStride + (size_t)y * g.imgPitch;
    const uint8_t* src = img + rowByte;
    const int x0 = 4 * wx - ORBFE_EDGE;  // image column of byte 0
    if (x0 >= 0 && x0 + 3 < L.w) {
      // interior: the image block is tightly packed (imgPitch == width, rows start at any byte), so
      // the 4 pixels come from the two aligned words around them (the block has 16 B of slack)
      const size_t a = rowByte + (size_t)x0;
      const unsigned* s4 = reinterpret_cast<const unsigned*>(img) + (a >> 2);
      out = __funnelshift_r(__ldg(s4), __ldg(s4 + 1), 8 * (int)(a & 3));
    } else {
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const int px = 4 * wx + b;
        const int x = orbfe_reflect101(min(px, pw - 1) - ORBFE_EDGE, L.w);
        out |= (unsigned)__ldg(src + x) << (8 * b);
      }
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

// ---- level 0: copy of the input frame into its padded plane (copyMakeBorder REFLECT_101, :1071) -------
// One CTA = one padded row of one image.  Interior 16-byte chunks: 5 aligned words of the (tightly packed,
// arbitrarily aligned) image row + 4 funnel shifts, one uint4 store.  The few chunks that touch the 19-px
// border are then filled byte by byte by the whole CTA (one byte per thread), so that no warp runs a
// 16-byte serial border path next to its interior lanes.
#define ORBFE_PYR0_THREADS 128
__global__ void __launch_bounds__(ORBFE_PYR0_THREADS)
k_pyramid_level0(const __grid_constant__ Geom g, const uint8_t* __restrict__ img, uint8_t* __restrict__ pyr) {
  const LevelGeom& L = g.lv[0];
  const int py = blockIdx.x, slot = blockIdx.y;
  const int y = orbfe_reflect101(py - ORBFE_EDGE, L.h);
  const size_t rowByte = (size_t)slot * g.imgStride + (size_t)y * g.imgPitch;
  const uint8_t* src = img + rowByte;
  uint8_t* drow = pyr + (size_t)slot * g.pyrStride + L.planeOff + (size_t)py * L.pitch;
  uint4* dst = reinterpret_cast<uint4*>(drow);
  const int pw = L.w + 2 * ORBFE_EDGE;
  // interior chunks: c in [cLo, cHi): 16c-19 >= 0 and 16c-19+15 < w
  const int cLo = (ORBFE_EDGE + 15) >> 4, cHi = max(cLo, (L.w + ORBFE_EDGE - 16) / 16 + 1);
  for (int c = cLo + threadIdx.x; c < cHi; c += ORBFE_PYR0_THREADS) {
    const size_t a = rowByte + (size_t)(16 * c - ORBFE_EDGE);
    const unsigned* s4 = reinterpret_cast<const unsigned*>(img) + (a >> 2);
    const int sh = 8 * (int)(a & 3);
    const unsigned w0 = __ldg(s4), w1 = __ldg(s4 + 1), w2 = __ldg(s4 + 2), w3 = __ldg(s4 + 3), w4 = __ldg(s4 + 4);
    dst[c] = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                        __funnelshift_r(w3, w4, sh));
  }
  // border bytes: padded columns [0, 16*cLo) and [16*cHi, pitch)
  const int nLeft = 16 * cLo, nRight = L.pitch - 16 * cHi;
  for (int t = threadIdx.x; t < nLeft + nRight; t += ORBFE_PYR0_THREADS) {
    const int px = t < nLeft ? t : 16 * cHi + (t - nLeft);
    drow[px] = __ldg(src + orbfe_reflect101(min(px, pw - 1) - ORBFE_EDGE, L.w));
  }
}

// ---- fast resize path (levels >= 1, scale factor <= 2, not the exact-2x INTER_AREA case) ------------
// One warp = 32 consecutive words of the PADDED destination plane x stripRows rows (8 for batches; 2 when
// only a frame or two is in flight: a strip is a chain of dependent loads, short chains = low latency).  Everything that
// depends only on the column lives in registers for the whole strip (host-built PyrWordLut: source word,
// byte shift, PRMT selectors and the packed 11-bit coefficient pairs); everything that depends only on
// the row comes from one 16-byte PyrRowLut entry (source rows of the REFLECTED destination row, vertical
// coefficients pre-shifted for IMAD.HI).  Per SOURCE row a lane loads 3 aligned words, aligns them with 2
// funnel shifts and forms the 4 horizontal interpolations with one PRMT + one IDP.2A each.  The two
// register sets of horizontally interpolated rows swap roles every destination row (loop unrolled by 2),
// so the common "+1 source row" step re-uses the previous bottom row as the new top row without moves.
#define ORBFE_PYR_ROWS 8
#ifndef ORBFE_PYR_PREFETCH
#define ORBFE_PYR_PREFETCH 1  // A/B on B200 (64 pairs): 0 -> 0.263 ms, 1 -> 0.255 ms, 2 (both rows of every destination row) -> 0.277 ms
#endif

struct PyrWordLut {
  int srcW;          // first source word of the padded source row
  int sh;            // 8 * (first source byte & 3)
  unsigned sel;      // byte j = PRMT selector picking pixel j's two source bytes from the aligned window
  unsigned cpack[4]; // c0 | c1 << 16 (cv::resize 11-bit coefficients)
};
struct PyrRowLut {
  int s0, s1;        // source rows (level l-1 coordinates) of the vertical blend
  unsigned b0, b1;   // vertical coefficients << 16
};

template <bool NC>
__device__ __forceinline__ unsigned orbfe_ldw(const unsigned* p) {
  if (NC) return __ldg(p);
  return *p;  // the tail kernel reads planes it wrote itself: coherent path
}

template <bool NC>
__device__ __forceinline__ void orbfe_hrow(const unsigned* __restrict__ srow, const int srcW, const int sh,
                                           const unsigned (&sel)[4], const unsigned (&cp)[4], unsigned (&T)[4]) {
  const unsigned w0 = orbfe_ldw<NC>(srow + srcW), w1 = orbfe_ldw<NC>(srow + srcW + 1), w2 = orbfe_ldw<NC>(srow + srcW + 2);
  const unsigned lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
#pragma unroll
  for (int j = 0; j < 4; ++j) T[j] = __dp2a_lo(cp[j], __byte_perm(lo, hi, sel[j]), 0u) >> 4;  // (S0*c0 + S1*c1) >> 4
}

__device__ __forceinline__ unsigned orbfe_vblend(const PyrRowLut R, const unsigned (&Tt)[4], const unsigned (&Tb)[4]) {
  unsigned out = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const unsigned v = (__umulhi(R.b0, Tt[j]) + __umulhi(R.b1, Tb[j]) + 2u) >> 2;  // ((b0*(T0>>4))>>16 + (b1*(T1>>4))>>16 + 2) >> 2
    out |= v << (8 * j);
  }
  return out;
}

// one warp strip (32 words x ORBFE_PYR_ROWS rows) of level `level` of image `slot`
template <bool NC>
__device__ __forceinline__ void orbfe_resize_strip(const Geom& g, const int level, const int slot, const int task, const int lane,
                                                   const int stripRows, uint8_t* __restrict__ pyr,
                                                   const PyrRowLut* __restrict__ rlut, const PyrWordLut* __restrict__ wlut) {
  const LevelGeom& L = g.lv[level];
  const LevelGeom& P = g.lv[level - 1];
  const int strips = (L.pyrWords + 31) >> 5;
  const int ph = L.h + 2 * ORBFE_EDGE;
  const int ty = task / strips, tx = task - ty * strips;
  const int py0 = ty * stripRows;
  if (py0 >= ph) return;
  const int wx = min(tx * 32 + lane, L.pyrWords - 1);  // duplicate lanes rewrite the last word with the same value
  const PyrWordLut W = wlut[L.wlutOff + wx];
  unsigned sel[4], cp[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { sel[j] = (W.sel >> (8 * j)) & 0xffu; cp[j] = W.cpack[j]; }
  const int spitchW = P.pitch >> 2;
  const unsigned* src = reinterpret_cast<const unsigned*>(pyr + (size_t)slot * g.pyrStride + P.planeOff) + (size_t)ORBFE_EDGE * spitchW;
  const int dpitchW = L.pitch >> 2;
  unsigned* dst = reinterpret_cast<unsigned*>(pyr + (size_t)slot * g.pyrStride + L.planeOff) + (size_t)py0 * dpitchW + wx;
  const PyrRowLut* rl = rlut + L.rlutOff + py0;
  unsigned TA[4] = {0, 0, 0, 0}, TB[4] = {0, 0, 0, 0};
  int ra = -1, rb = -1;
  const int nrows = min(stripRows, ph - py0);
#if ORBFE_PYR_PREFETCH && !defined(ORBFE_EMU)
  // the strip is a chain of dependent row loads (each source row is fetched when its destination row is reached):
  // touch every source row of the strip first so that the chain runs out of L1 instead of paying one L2 round
  // trip per row
  for (int r = 0; r < nrows; ++r) {
    const PyrRowLut R = rl[r];
    asm volatile("prefetch.global.L1 [%0];" ::"l"(src + (size_t)R.s1 * spitchW + W.srcW + 1));
    if (r == 0 || ORBFE_PYR_PREFETCH > 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(src + (size_t)R.s0 * spitchW + W.srcW + 1));
  }
#endif
  for (int r = 0; r < nrows; r += 2) {
    {  // even row: top = TA, bottom = TB
      const PyrRowLut R = rl[r];
      if (ra != R.s0) { orbfe_hrow<NC>(src + (size_t)R.s0 * spitchW, W.srcW, W.sh, sel, cp, TA); ra = R.s0; }
      if (rb != R.s1) { orbfe_hrow<NC>(src + (size_t)R.s1 * spitchW, W.srcW, W.sh, sel, cp, TB); rb = R.s1; }
      dst[(size_t)r * dpitchW] = orbfe_vblend(R, TA, TB);
    }
    if (r + 1 < nrows) {  // odd row: roles swapped, so a "+1 source row" step re-uses TB as the top row
      const PyrRowLut R = rl[r + 1];
      if (rb != R.s0) { orbfe_hrow<NC>(src + (size_t)R.s0 * spitchW, W.srcW, W.sh, sel, cp, TB); rb = R.s0; }
      if (ra != R.s1) { orbfe_hrow<NC>(src + (size_t)R.s1 * spitchW, W.srcW, W.sh, sel, cp, TA); ra = R.s1; }
      dst[(size_t)(r + 1) * dpitchW] = orbfe_vblend(R, TB, TA);
    }
  }
}

#ifndef ORBFE_PYR_MINB
#define ORBFE_PYR_MINB 1
#endif
__global__ void __launch_bounds__(ORBFE_PYR_THREADS, ORBFE_PYR_MINB)
k_pyramid_resize(const __grid_constant__ Geom g, const int level, const int stripRows, uint8_t* __restrict__ pyr,
                 const PyrRowLut* __restrict__ rlut, const PyrWordLut* __restrict__ wlut) {
  const int task = blockIdx.x * (ORBFE_PYR_THREADS / 32) + (threadIdx.x >> 5);
  orbfe_resize_strip<true>(g, level, blockIdx.y, task, threadIdx.x & 31, stripRows, pyr, rlut, wlut);
}


// ---- streaming resize (levels >= 1, the k_pyramid_resize arithmetic fed from shared memory) ----------------------
// One WARP = a column strip of 32 words (128 px) of the PADDED destination plane x one vertical segment of the level's
// interior rows; the ORBFE_PYRS_WPC warps of a CTA work on consecutive strips and share nothing.  The rows of the source level arrive as a stream of TMA boxes (pyrBoxW bytes x ORBFE_PYRS_RB
// rows) through a ring of ORBFE_PYRS_NB shared-memory buffers with one mbarrier each (orbfe_tma.cuh): the warp marches
// down its destination rows, waits for the box that holds the lower source row of the current row, and re-arms a buffer
// as soon as the march has left its rows.  No load sits on the dependent chain of a row any more (the previous form was
// bound by it: ncu long-scoreboard 7.8 per issue), and every source byte is fetched from L2 once per strip.
// Columns: the per-word LUT of k_pyramid_resize (source word, byte shift, PRMT selectors, 11-bit coefficient pairs); the
// box starts at the 16-byte boundary at or below the strip's first source byte (host table pyrBoxX).  Rows: the
// interior rows of the per-row LUT.  The 19-px top / bottom borders are BORDER_REFLECT_101 images of interior rows
// 1..19 and h-20..h-2 of the SAME level (orb_extractor.cpp:1066,1071), so the warp that produces such a row stores it
// twice; the left / right borders are ordinary words of the strip whose LUT entries point at the reflected columns.
#ifndef ORBFE_PYRS_RB
#define ORBFE_PYRS_RB 8      // source rows per box (power of two)
#endif
#ifndef ORBFE_PYRS_NB
#define ORBFE_PYRS_NB 4      // boxes in the ring (power of two): the ring holds RB * NB source rows
#endif
#ifndef ORBFE_PYRS_SEG
#define ORBFE_PYRS_SEG 56    // target destination rows per segment (A/B on B200: 112 -> 0.253 ms, 56 -> 0.246, 32 -> 0.263, 20 -> 0.285)
#endif
#ifndef ORBFE_PYRS_WPC
#define ORBFE_PYRS_WPC 1     // independent warps (strips) per CTA.  A/B on B200 (128 frames): 1 -> 0.246 ms, 2 -> 0.266, 4 -> 0.271
#endif
#define ORBFE_PYRS_THREADS (32 * ORBFE_PYRS_WPC)
#define ORBFE_PYRS_RING (ORBFE_PYRS_RB * ORBFE_PYRS_NB)

// horizontal interpolation of the lane's 4 pixels from one source row held in shared memory
__device__ __forceinline__ void orbfe_hrow_smem(const unsigned* srow, const int sh, const unsigned (&sel)[4], const unsigned (&cp)[4],
                                                unsigned (&T)[4]) {
  const unsigned w0 = srow[0], w1 = srow[1], w2 = srow[2];
  const unsigned lo = __funnelshift_r(w0, w1, sh), hi = __funnelshift_r(w1, w2, sh);
#pragma unroll
  for (int j = 0; j < 4; ++j) T[j] = __dp2a_lo(cp[j], __byte_perm(lo, hi, sel[j]), 0u) >> 4;  // (S0*c0 + S1*c1) >> 4
}

__global__ void __launch_bounds__(ORBFE_PYRS_THREADS, 32 / ORBFE_PYRS_WPC)
k_pyramid_strip(const __grid_constant__ Geom g, const int level, uint8_t* __restrict__ pyr, const CUtensorMap* __restrict__ tmaps,
                const PyrRowLut* __restrict__ rlut, const PyrWordLut* __restrict__ wlut, const int* __restrict__ boxX,
                const int nSegs) {
  // dynamic shared memory: the ring (RING rows of pyrBoxW bytes: source row r of the segment lives at row r % RING; a box
  // of RB rows starts on a 128-byte boundary because pyrBoxW is a multiple of 16 and RB of 8), then the segment's row LUT
  // (a global load per row would sit on the march's critical path)
  ORBFE_DYN_SMEM(smem);
  __shared__ __align__(8) unsigned long long s_bars[ORBFE_PYRS_WPC][ORBFE_PYRS_NB];
  static_assert((ORBFE_PYRS_RB & (ORBFE_PYRS_RB - 1)) == 0 && (ORBFE_PYRS_NB & (ORBFE_PYRS_NB - 1)) == 0 && ORBFE_PYRS_RB % 8 == 0,
                "ring geometry");
  const LevelGeom& L = g.lv[level];
  const LevelGeom& S = g.lv[level - 1];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, slot = blockIdx.y;
  const int nStrips = (L.pyrWords + 31) >> 5;
  const int task = blockIdx.x * ORBFE_PYRS_WPC + wid;
  const int sg = task / nStrips, tx = task - sg * nStrips;
  const int segH = (L.h + nSegs - 1) / nSegs;
  const int y0 = sg * segH, y1 = min(y0 + segH, L.h);
  if (sg >= nSegs || y0 >= y1) return;
  const int pitchW = L.pyrBoxW >> 2;              // box row pitch in words
  const int warpBytes = (ORBFE_PYRS_RING * L.pyrBoxW + segH * (int)sizeof(PyrRowLut) + 127) & ~127;  // this warp's slice
  unsigned* s_ring = reinterpret_cast<unsigned*>(smem + wid * warpBytes);
  PyrRowLut* s_rl = reinterpret_cast<PyrRowLut*>(smem + wid * warpBytes + ORBFE_PYRS_RING * L.pyrBoxW);
  unsigned long long* s_bar = s_bars[wid];
  const int wx = min(tx * 32 + lane, L.pyrWords - 1);  // duplicate lanes rewrite the last word with the same value
  const PyrWordLut W = wlut[L.wlutOff + wx];
  unsigned sel[4], cp[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) { sel[j] = (W.sel >> (8 * j)) & 0xffu; cp[j] = W.cpack[j]; }
  const int bx = boxX[L.pyrBoxOff + tx];          // first byte of the boxes (padded source column, multiple of 16)
  const unsigned* ringLane = s_ring + (W.srcW - (bx >> 2));  // the lane's first source word inside a ring row
  const int sh = W.sh;
  const PyrRowLut* rl = rlut + L.rlutOff + ORBFE_EDGE;  // entry of interior row y at [y]
  const int sBase = rl[y0].s0;                    // first source row of the segment (interior coordinates)
  for (int r = lane; r < y1 - y0; r += 32) {      // source rows relative to the segment's first
    PyrRowLut R = rl[y0 + r];
    R.s0 -= sBase; R.s1 -= sBase;
    s_rl[r] = R;
  }
  const int nBoxes = (rl[y1 - 1].s1 - sBase) / ORBFE_PYRS_RB + 1;
  OrbfeTmaPlane P;
  P.base = pyr + S.planeOff; P.sliceStride = g.pyrStride; P.pitch = S.pitch; P.rows = S.h + 2 * ORBFE_EDGE;
  P.slices = gridDim.y; P.boxW = L.pyrBoxW; P.boxH = ORBFE_PYRS_RB;
  const int by = ORBFE_EDGE + sBase;              // padded source row of box 0
  const int boxWords = ORBFE_PYRS_RB * pitchW;
  if (lane == 0)
    for (int b = 0; b < ORBFE_PYRS_NB; ++b) orbfe_tile_barrier_init(&s_bar[b]);
  __syncwarp();  // the barriers, the ring and the row LUT belong to this warp alone
  int nextIssue = min(ORBFE_PYRS_NB, nBoxes), waited = 0;
  if (lane == 0) {
    orbfe_tmap_acquire(tmaps + level);
    for (int b = 0; b < nextIssue; ++b)
      orbfe_tile_issue(s_ring + b * boxWords, &s_bar[b], tmaps + level, P, bx, by + b * ORBFE_PYRS_RB, slot);
  }
  const int dpitchW = L.pitch >> 2;
  unsigned* plane = reinterpret_cast<unsigned*>(pyr + (size_t)slot * g.pyrStride + L.planeOff) + wx;
  unsigned* drow = plane + (size_t)(ORBFE_EDGE + y0) * dpitchW;
  unsigned TA[4] = {0, 0, 0, 0}, TB[4] = {0, 0, 0, 0};
  int ra = -1, rb = -1;
  int landed = 0;  // source rows [0, landed) of the segment are in the ring
  // one destination row: upper source row in Tt (cached as rt), lower in Tb (rbm)
  auto row = [&](const int y, unsigned (&Tt)[4], int& rt, unsigned (&Tb)[4], int& rbm, const bool border) {
    const PyrRowLut R = s_rl[y - y0];
    if (R.s1 >= landed) {  // warp-uniform, once per box: the lower source row enters the next box
      orbfe_tile_wait_warp(&s_bar[waited & (ORBFE_PYRS_NB - 1)], (unsigned)(waited / ORBFE_PYRS_NB) & 1u);
      ++waited;
      landed += ORBFE_PYRS_RB;
      // re-arm the boxes the march has left (every lane has read them: __syncwarp): box nextIssue re-uses the buffer
      // of box nextIssue - NB, whose last row is (nextIssue - NB + 1) * RB - 1
      if (nextIssue < nBoxes && (nextIssue - ORBFE_PYRS_NB + 1) * ORBFE_PYRS_RB <= R.s0) {
        __syncwarp();
        if (lane == 0)
          orbfe_tile_issue(s_ring + (nextIssue & (ORBFE_PYRS_NB - 1)) * boxWords, &s_bar[nextIssue & (ORBFE_PYRS_NB - 1)],
                           tmaps + level, P, bx, by + nextIssue * ORBFE_PYRS_RB, slot);
        ++nextIssue;
      }
    }
    if (rt != R.s0) { orbfe_hrow_smem(ringLane + (R.s0 & (ORBFE_PYRS_RING - 1)) * pitchW, sh, sel, cp, Tt); rt = R.s0; }
    if (rbm != R.s1) { orbfe_hrow_smem(ringLane + (R.s1 & (ORBFE_PYRS_RING - 1)) * pitchW, sh, sel, cp, Tb); rbm = R.s1; }
    const unsigned out = orbfe_vblend(R, Tt, Tb);
    *drow = out;
    drow += dpitchW;
    if (border) {
      if (y <= ORBFE_EDGE && y >= 1) plane[(size_t)(ORBFE_EDGE - y) * dpitchW] = out;  // top border: row -y = row y
      if (y >= L.h - 1 - ORBFE_EDGE && y <= L.h - 2) plane[(size_t)(2 * (L.h - 1) - y + ORBFE_EDGE) * dpitchW] = out;  // bottom
    }
  };
  // the two register sets swap roles from row to row, so that the common "+1 source row" step re-uses the previous
  // lower row as the new upper row without moves.  Rows that also feed a border row: y <= 19 or y >= h - 20.
  const int yA = min(y1, max(y0, ORBFE_EDGE + 1)), yB = max(yA, min(y1, L.h - 1 - ORBFE_EDGE));  // [yA, yB): no border copies
  int y = y0;
  bool flip = false;
  for (; y < yA; ++y, flip = !flip) { if (!flip) row(y, TA, ra, TB, rb, true); else row(y, TB, rb, TA, ra, true); }
  if (flip && y < yB) { row(y, TB, rb, TA, ra, false); ++y; flip = false; }
  for (; y + 1 < yB; y += 2) {
    row(y, TA, ra, TB, rb, false);
    row(y + 1, TB, rb, TA, ra, false);
  }
  for (; y < y1; ++y, flip = !flip) { if (!flip) row(y, TA, ra, TB, rb, y >= yB); else row(y, TB, rb, TA, ra, y >= yB); }
}
