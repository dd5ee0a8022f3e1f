// k_fast.cuh -- E2+E3: grid FAST-9/16 with adaptive threshold, per-cell NMS and ordered select.
// Follows the grid loop of ComputeKeyPointsOctTree (orb_extractor.cpp:706-770) and
// cv::FAST(cell, kps, th, true):
//   * one CTA per 30-px grid cell; the cell sub-image [iniX,maxX)x[iniY,maxY) is staged in
//     shared memory; FAST runs on its 3-px inset;
//   * the corner score (max threshold at which the pixel is still a corner) is computed once at
//     minThFAST; NMS is strict (>) over the 8 neighbours INSIDE the cell (outside counts 0);
//   * survivors with score >= iniThFAST are emitted; if there are none, survivors with
//     score >= minThFAST are emitted instead (the post-NMS emptiness fallback, :753-757);
//   * emission order is row-major inside the cell (warp ballots + popc prefix), which is the
//     order cv::FAST returns; cells are consumed in row-major cell order by the quad-tree kernel.
#pragma once
#include "orbfe_common.cuh"

#define ORBFE_FAST_THREADS 256

// 16-bit mask has a circular run of >= 9 set bits
__device__ __forceinline__ bool orbfe_has_run9(unsigned m) {
  unsigned t = m | (m << 16);
  unsigned a = t & (t >> 1);   // runs of 2
  a &= a >> 2;                 // runs of 4
  a &= a >> 4;                 // runs of 8
  a &= t >> 8;                 // runs of 9
  return (a & 0xffffu) != 0;
}

// max over the 16 circular 9-arcs of the minimum of d[] over the arc
__device__ __forceinline__ int orbfe_arc_maxmin(const int (&d)[16]) {
  int m2[16], m4[16], best = -100000;
#pragma unroll
  for (int i = 0; i < 16; ++i) m2[i] = min(d[i], d[(i + 1) & 15]);
#pragma unroll
  for (int i = 0; i < 16; ++i) m4[i] = min(m2[i], m2[(i + 2) & 15]);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int m8 = min(m4[i], m4[(i + 4) & 15]);
    best = max(best, min(m8, d[(i + 8) & 15]));
  }
  return best;
}

// FAST corner score of the pixel at p (shared-memory tile, pitch tp); 0 if not a corner at th.
__device__ __forceinline__ int orbfe_fast_score(const uint8_t* p, int tp, int th) {
  const int v = p[0];
  int r[16];
  r[0] = p[3 * tp];      r[1] = p[3 * tp + 1];  r[2] = p[2 * tp + 2];   r[3] = p[tp + 3];
  r[4] = p[3];           r[5] = p[-tp + 3];     r[6] = p[-2 * tp + 2];  r[7] = p[-3 * tp + 1];
  r[8] = p[-3 * tp];     r[9] = p[-3 * tp - 1]; r[10] = p[-2 * tp - 2]; r[11] = p[-tp - 3];
  r[12] = p[-3];         r[13] = p[tp - 3];     r[14] = p[2 * tp - 2];  r[15] = p[3 * tp - 1];
  unsigned bm = 0, dm = 0;
  const int hi = v + th, lo = v - th;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    bm |= (unsigned)(r[k] > hi) << k;
    dm |= (unsigned)(r[k] < lo) << k;
  }
  const bool bright = orbfe_has_run9(bm), dark = orbfe_has_run9(dm);
  if (!bright && !dark) return 0;
  int d[16], s = 0;
  if (bright) {
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = r[k] - v;
    s = orbfe_arc_maxmin(d);
  }
  if (dark) {
#pragma unroll
    for (int k = 0; k < 16; ++k) d[k] = v - r[k];
    s = max(s, orbfe_arc_maxmin(d));
  }
  return s - 1;  // >= th by construction
}

__global__ void __launch_bounds__(ORBFE_FAST_THREADS)
k_fast_cells(const __grid_constant__ Geom g, const uint8_t* __restrict__ pyr, int* __restrict__ cellCnt,
             unsigned* __restrict__ cellList, const int tilePitch, const int maxInnerH) {
  ORBFE_DYN_SMEM(smem);
  __shared__ int s_warpCnt[ORBFE_FAST_THREADS / 32];
  const int slot = blockIdx.y;
  const int cell = blockIdx.x;
  int level = 0;
  for (int l = 1; l < g.nlevels; ++l)
    if (cell >= g.lv[l].cellBase) level = l;
  const LevelGeom& L = g.lv[level];
  const int ci = cell - L.cellBase;
  const int i = ci / L.nCols, j = ci - i * L.nCols;
  int* cnt = cellCnt + (size_t)slot * g.totalCells + cell;
  unsigned* list = cellList + (size_t)slot * g.cellListStride + L.cellListOff + (size_t)ci * L.cellCap;
  const int iniY = ORBFE_MINB + i * L.hCell, iniX = ORBFE_MINB + j * L.wCell;
  const int maxY = min(iniY + L.hCell + 6, L.maxBY), maxX = min(iniX + L.wCell + 6, L.maxBX);
  const int cw = maxX - iniX, ch = maxY - iniY;
  // skipped cells (orb_extractor.cpp:735,744) and sub-images too small for FAST
  if (iniY >= L.maxBY - 3 || iniX >= L.maxBX - 6 || cw < 7 || ch < 7) {
    if (threadIdx.x == 0) *cnt = 0;
    return;
  }
  // shared layout: tile [ch][tilePitch] | score [(ih+2)][sp] with a zero apron
  const int iw = cw - 6, ih = ch - 6;
  const int sp = tilePitch;  // >= iw + 2
  uint8_t* tile = smem;
  uint8_t* score = smem + (size_t)(maxInnerH + 6) * tilePitch;
  const uint8_t* src = pyr + (size_t)slot * g.pyrStride + L.planeOff + (size_t)(iniY + ORBFE_EDGE) * L.pitch + iniX + ORBFE_EDGE;
  for (int t = threadIdx.x; t < ch * cw; t += ORBFE_FAST_THREADS) {
    const int y = t / cw, x = t - y * cw;
    tile[y * tilePitch + x] = __ldg(src + (size_t)y * L.pitch + x);
  }
  for (int t = threadIdx.x; t < (ih + 2) * sp; t += ORBFE_FAST_THREADS) score[t] = 0;
  __syncthreads();
  for (int t = threadIdx.x; t < ih * iw; t += ORBFE_FAST_THREADS) {
    const int y = t / iw, x = t - y * iw;
    const int s = orbfe_fast_score(tile + (y + 3) * tilePitch + x + 3, tilePitch, g.minTh);
    score[(y + 1) * sp + x + 1] = (uint8_t)s;
  }
  __syncthreads();
  // pass 1: NMS, count survivors at iniTh
  int nIni = 0;
  for (int t = threadIdx.x; t < ih * iw; t += ORBFE_FAST_THREADS) {
    const int y = t / iw, x = t - y * iw;
    const uint8_t* c = score + (y + 1) * sp + x + 1;
    const int s = c[0];
    const bool keep = s > 0 && s > c[-1] && s > c[1] && s > c[-sp - 1] && s > c[-sp] && s > c[-sp + 1] &&
                      s > c[sp - 1] && s > c[sp] && s > c[sp + 1];
    // reuse the tile buffer to hold the NMS result (tile is no longer needed)
    tile[t] = keep ? (uint8_t)s : (uint8_t)0;
    nIni += (keep && s >= g.iniTh) ? 1 : 0;
  }
  const int anyIni = __syncthreads_or(nIni);
  const int th = anyIni ? g.iniTh : g.minTh;
  // pass 2: ordered compaction, row-major
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int base = 0;
  const int xoff = 3 + j * L.wCell, yoff = 3 + i * L.hCell;
  for (int t0 = 0; t0 < ih * iw; t0 += ORBFE_FAST_THREADS) {
    const int t = t0 + threadIdx.x;
    const int s = t < ih * iw ? tile[t] : 0;
    const bool emit = s >= th && s > 0;
    const unsigned bal = __ballot_sync(0xffffffffu, emit);
    if (lane == 0) s_warpCnt[wid] = __popc(bal);
    __syncthreads();
    int woff = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < ORBFE_FAST_THREADS / 32; ++w) {
      const int c = s_warpCnt[w];
      if (w < wid) woff += c;
      tot += c;
    }
    if (emit) {
      const int y = t / iw, x = t - y * iw;
      const int pos = base + woff + __popc(bal & ((1u << lane) - 1u));
      if (pos < L.cellCap) list[pos] = orbfe_pack(x + xoff, y + yoff, s);
    }
    base += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) *cnt = min(base, L.cellCap);
}
