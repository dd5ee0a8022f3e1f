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
#define ORBFE_FAST_MAXG 16

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

// ---- fused band-segment kernel ----------------------------------------------------------------------
// One CTA = up to L.fG consecutive grid cells of one cell row of one level of one image.  The cells'
// detection windows tile the level's inner region exactly (origin (19,19), stride wCell x hCell, SURVEY
// A.2), so the CTA stages ONE pixel tile (hCell+6 rows x fG*wCell+6 columns, word-aligned, coalesced
// 4-byte loads) and runs:
//   1. compass pre-test, 4 pixels per instruction (SIMD-in-register u8): a 9-arc of the 16-ring always
//      contains two ADJACENT compass points (ring 0/4/8/12), so pixels without two adjacent compass
//      points both > v+t or both < v-t are rejected; survivors (~10 %) are compacted into a queue;
//   2. exact FAST-9 test + corner score on the queue only, all lanes busy;
//   3. strict 3x3 NMS restricted to the cell (neighbours outside the cell's window count 0) + the
//      per-cell "any survivor >= iniThFAST" flag that drives the minThFAST fallback (:753-757);
//   4. one warp per cell: row-major ordered emission with warp ballots (the order cv::FAST returns).
#ifndef ORBFE_EMU
#define ORBFE_VADDUS4(a, b) __vaddus4(a, b)
#define ORBFE_VSUBUS4(a, b) __vsubus4(a, b)
#define ORBFE_VCMPGTU4(a, b) __vcmpgtu4(a, b)
#define ORBFE_VCMPLTU4(a, b) __vcmpltu4(a, b)
#else
static inline unsigned orbfe_emu_b4(unsigned a, unsigned b, int op) {
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) {
    const int x = (a >> (8 * i)) & 0xff, y = (b >> (8 * i)) & 0xff;
    int v = 0;
    if (op == 0) v = x + y > 255 ? 255 : x + y;
    if (op == 1) v = x - y < 0 ? 0 : x - y;
    if (op == 2) v = x > y ? 0xff : 0;
    if (op == 3) v = x < y ? 0xff : 0;
    r |= (unsigned)v << (8 * i);
  }
  return r;
}
#define ORBFE_VADDUS4(a, b) orbfe_emu_b4(a, b, 0)
#define ORBFE_VSUBUS4(a, b) orbfe_emu_b4(a, b, 1)
#define ORBFE_VCMPGTU4(a, b) orbfe_emu_b4(a, b, 2)
#define ORBFE_VCMPLTU4(a, b) orbfe_emu_b4(a, b, 3)
#endif

__global__ void __launch_bounds__(ORBFE_FAST_THREADS)
k_fast_cells(const __grid_constant__ Geom g, const uint8_t* __restrict__ pyr, int* __restrict__ cellCnt,
             unsigned* __restrict__ cellList, const int pitchW, const int maxRows, const int queueCap) {
  ORBFE_DYN_SMEM(smem);
  unsigned* tileW = reinterpret_cast<unsigned*>(smem);               // [maxRows][pitchW] pixels, later NMS survivors
  unsigned* scoreW = tileW + (size_t)maxRows * pitchW;               // [maxRows][pitchW] corner scores
  unsigned short* queue = reinterpret_cast<unsigned short*>(scoreW + (size_t)maxRows * pitchW);
  __shared__ int s_qn;
  __shared__ int s_any[ORBFE_FAST_MAXG];
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int slot = blockIdx.y;
  int level = 0;
  for (int l = 1; l < g.nlevels; ++l)
    if ((int)blockIdx.x >= g.lv[l].fastBase) level = l;
  const LevelGeom& L = g.lv[level];
  const int bi = blockIdx.x - L.fastBase;
  const int i = bi / L.fSegs, j0 = (bi - i * L.fSegs) * L.fG;
  const int nj = min(L.fG, L.nCols - j0);
  int* cnt = cellCnt + (size_t)slot * g.totalCells + L.cellBase + i * L.nCols + j0;
  unsigned* list = cellList + (size_t)slot * g.cellListStride + L.cellListOff + (size_t)(i * L.nCols + j0) * L.cellCap;
  const int iniY = ORBFE_MINB + i * L.hCell, iniX = ORBFE_MINB + j0 * L.wCell;
  const int maxY = min(iniY + L.hCell + 6, L.maxBY);
  const int maxX = min(iniX + nj * L.wCell + 6, L.maxBX);
  const int rows = maxY - iniY;
  // skipped rows/cells (orb_extractor.cpp:735,744) and sub-images too small for FAST yield nothing
  if (rows < 7 || maxX - iniX < 7) {
    if (tid < nj) cnt[tid] = 0;
    return;
  }
  const int pitchB = pitchW * 4;
  const int gx0 = (iniX + ORBFE_EDGE) & ~3;                          // padded-plane column of tile column 0
  const int tw = (maxX + ORBFE_EDGE - gx0 + 3) >> 2;                  // tile width in words
  const uint8_t* src = pyr + (size_t)slot * g.pyrStride + L.planeOff + (size_t)(iniY + ORBFE_EDGE) * L.pitch + gx0;
  for (int r = wid; r < rows; r += ORBFE_FAST_THREADS / 32) {
    const unsigned* srow = reinterpret_cast<const unsigned*>(src + (size_t)r * L.pitch);
    for (int c = lane; c < pitchW; c += 32) {
      tileW[r * pitchW + c] = c < tw ? __ldg(srow + c) : 0u;
      scoreW[r * pitchW + c] = 0u;
    }
  }
  if (tid == 0) s_qn = 0;
  if (tid < ORBFE_FAST_MAXG) s_any[tid] = 0;
  __syncthreads();
  // ---- 1. compass pre-test on words
  const int ix0 = iniX + 3 + ORBFE_EDGE - gx0, ix1 = maxX - 3 + ORBFE_EDGE - gx0;  // inner columns (tile coords)
  const int w0 = ix0 >> 2, nWi = ((ix1 - 1) >> 2) - w0 + 1, nRi = rows - 6;
  const unsigned th4 = (unsigned)g.minTh * 0x01010101u;
  for (int t = tid; t < nRi * nWi; t += ORBFE_FAST_THREADS) {
    const int ry = t / nWi, wx = w0 + t - ry * nWi, y = 3 + ry;
    const unsigned* row = tileW + y * pitchW + wx;
    const unsigned c = row[0], up = row[-3 * pitchW], dn = row[3 * pitchW];
    const unsigned prev = wx > 0 ? row[-1] : 0u, next = wx + 1 < pitchW ? row[1] : 0u;
    const unsigned lf = __funnelshift_r(prev, c, 8), rt = __funnelshift_r(c, next, 24);
    const unsigned hi = ORBFE_VADDUS4(c, th4), lo = ORBFE_VSUBUS4(c, th4);
    const unsigned B0 = ORBFE_VCMPGTU4(dn, hi), B4 = ORBFE_VCMPGTU4(rt, hi), B8 = ORBFE_VCMPGTU4(up, hi), B12 = ORBFE_VCMPGTU4(lf, hi);
    const unsigned D0 = ORBFE_VCMPLTU4(dn, lo), D4 = ORBFE_VCMPLTU4(rt, lo), D8 = ORBFE_VCMPLTU4(up, lo), D12 = ORBFE_VCMPLTU4(lf, lo);
    unsigned m = ((B0 | B8) & (B4 | B12)) | ((D0 | D8) & (D4 | D12));
    // (B0&B4)|(B4&B8)|(B8&B12)|(B12&B0) == (B0|B8)&(B4|B12)
    if (m == 0u) continue;
    const int xb = 4 * wx;
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if (xb + b < ix0 || xb + b >= ix1) m &= ~(0xffu << (8 * b));
    const int n = __popc(m & 0x01010101u);
    if (n == 0) continue;
    int pos = atomicAdd(&s_qn, n);
#pragma unroll
    for (int b = 0; b < 4; ++b)
      if ((m >> (8 * b)) & 1u) { if (pos < queueCap) queue[pos] = (unsigned short)((y << 9) | (xb + b)); ++pos; }
  }
  __syncthreads();
  // ---- 2. exact test + score on the queue
  const int qn = min(s_qn, queueCap);
  const uint8_t* tileB = reinterpret_cast<const uint8_t*>(tileW);
  uint8_t* scoreB = reinterpret_cast<uint8_t*>(scoreW);
  for (int e = tid; e < qn; e += ORBFE_FAST_THREADS) {
    const int code = queue[e], x = code & 511, y = code >> 9;
    const int s = orbfe_fast_score(tileB + y * pitchB + x, pitchB, g.minTh);
    if (s > 0) scoreB[y * pitchB + x] = (uint8_t)s;
  }
  __syncthreads();
  // the pixel tile is no longer needed: it becomes the plane of NMS survivors
  for (int t = tid; t < rows * pitchW; t += ORBFE_FAST_THREADS) tileW[t] = 0u;
  __syncthreads();
  // ---- 3. NMS inside the cell + fallback flag
  uint8_t* nmsB = reinterpret_cast<uint8_t*>(tileW);
  for (int e = tid; e < qn; e += ORBFE_FAST_THREADS) {
    const int code = queue[e], x = code & 511, y = code >> 9;
    const uint8_t* c = scoreB + y * pitchB + x;
    const int s = c[0];
    if (s == 0) continue;
    const int jl = (x - ix0) / L.wCell;
    const int cx0 = ix0 + jl * L.wCell, cx1 = min(cx0 + L.wCell, ix1);
    const bool hasL = x > cx0, hasR = x + 1 < cx1;  // rows outside the inner band hold score 0 already
    bool keep = s > c[-pitchB] && s > c[pitchB];
    if (hasL) keep = keep && s > c[-1] && s > c[-pitchB - 1] && s > c[pitchB - 1];
    if (hasR) keep = keep && s > c[1] && s > c[-pitchB + 1] && s > c[pitchB + 1];
    if (keep) {
      nmsB[y * pitchB + x] = (uint8_t)s;
      if (s >= g.iniTh) s_any[jl] = 1;
    }
  }
  __syncthreads();
  // ---- 4. ordered emission, one warp per cell
  for (int jl = wid; jl < nj; jl += ORBFE_FAST_THREADS / 32) {
    const int cx0 = ix0 + jl * L.wCell, cx1 = min(cx0 + L.wCell, ix1);
    const int th = s_any[jl] ? g.iniTh : g.minTh;
    unsigned* out = list + (size_t)jl * L.cellCap;
    int base = 0;
    if (cx1 > cx0) {
      for (int y = 3; y < rows - 3; ++y)
        for (int c0 = cx0; c0 < cx1; c0 += 32) {
          const int x = c0 + lane;
          const int s = x < cx1 ? nmsB[y * pitchB + x] : 0;
          const bool emit = s >= th;
          const unsigned bal = __ballot_sync(0xffffffffu, emit);
          if (emit) {
            const int pos = base + __popc(bal & ((1u << lane) - 1u));
            if (pos < L.cellCap) out[pos] = orbfe_pack(x + gx0 - ORBFE_EDGE - ORBFE_MINB, y + iniY - ORBFE_MINB, s);
          }
          base += __popc(bal);
        }
    }
    if (lane == 0) cnt[jl] = min(base, L.cellCap);
  }
}
