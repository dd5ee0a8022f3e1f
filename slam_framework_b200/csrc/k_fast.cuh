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
#ifndef ORBFE_FAST_FGCAP
#define ORBFE_FAST_FGCAP 8   // upper bound of cells per CTA
#endif
#ifndef ORBFE_FAST_X2
#define ORBFE_FAST_X2 0  // 1: two queue entries per lane on packed u16x2 networks (fewer instructions, but 64 regs => 4 CTAs/SM: slower)
#endif
#ifndef ORBFE_FAST_PITCHW
#define ORBFE_FAST_PITCHW 61  // compile-time tile pitch (words) of the common geometry; odd => no column bank conflicts
#endif

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
// 4-byte loads) and runs up to two rounds, first at iniThFAST over the whole tile, then at minThFAST
// over the cells that produced no keypoint (the post-NMS emptiness fallback of :753-757; cv::FAST at
// threshold t == "score >= t" on the threshold-independent score map):
//   1. compass pre-test, 4 pixels per instruction (VABSDIFF4 + carry trick): a 9-arc of the 16-ring
//      always contains two ADJACENT compass points (ring 0/4/8/12), so a pixel without two adjacent
//      compass points differing from it by more than t cannot be a corner; survivors are compacted
//      into a shared-memory queue;
//   2. exact corner score (max_arc min_9 of +-(ring - v), 3-input min/max networks) on the queue only,
//      all lanes busy; corner at t <=> score >= t;
//   3. strict 3x3 NMS restricted to the cell (neighbours outside the cell's window count 0);
//   4. one warp per cell: row-major ordered emission with warp ballots (the order cv::FAST returns),
//      visiting only rows that hold a survivor.
#define ORBFE_FAST_ROWWORDS 3  // survivor row masks: up to 96 tile rows

// exact FAST-9/16 corner score of the pixel at p (shared-memory tile, byte pitch tp):
//   score + 1 = max( max_arc min_9 (ring - v),  max_arc min_9 (v - ring) ).
// Both halves run through ONE min3/min3/max3 network on packed u16x2 lanes (VIMNMX3.U16x2): the low half
// carries (ring - v) + 256, the high half (v - ring) + 256, both in [1, 511]; the packing is one IMAD per
// ring pixel:  (d + 256) | (256 - d) << 16  ==  C + ring * (1 - 65536)  with C folding the centre value.
__device__ __forceinline__ int orbfe_fast_score3(const uint8_t* p, int tp) {
  const unsigned K = 1u - 65536u;                                  // d -> d in the low half, -d in the high half
  const unsigned C = (256u | (256u << 16)) - (unsigned)p[0] * K;   // centre folded in
  unsigned d[16];
#define ORBFE_D(k, off) d[k] = (unsigned)p[off] * K + C
  ORBFE_D(0, 3 * tp);       ORBFE_D(1, 3 * tp + 1);   ORBFE_D(2, 2 * tp + 2);   ORBFE_D(3, tp + 3);
  ORBFE_D(4, 3);            ORBFE_D(5, -tp + 3);      ORBFE_D(6, -2 * tp + 2);  ORBFE_D(7, -3 * tp + 1);
  ORBFE_D(8, -3 * tp);      ORBFE_D(9, -3 * tp - 1);  ORBFE_D(10, -2 * tp - 2); ORBFE_D(11, -tp - 3);
  ORBFE_D(12, -3);          ORBFE_D(13, tp - 3);      ORBFE_D(14, 2 * tp - 2);  ORBFE_D(15, 3 * tp - 1);
#undef ORBFE_D
  // The 16 arcs pair up: with A_k = min d[k+1..k+8] (k even), the arcs starting at k and at k+1 are min(d[k], A_k) and
  // min(A_k, d[k+9]), and the larger of the two is min(A_k, max(d[k], d[k+9])).  A_k is the min of four adjacent pairs
  // p[j] = min(d[2j+1], d[2j+2]):  8 (pairs) + 8 (three pairs) + 8 (max of the two ends) + 8 (min3 of the rest) + 4 (max tree)
  // = 36 two-lane min/max instructions instead of the 41 of the min3 / min3 / max3 network.
  unsigned pr[8], s3[8], u[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) pr[j] = __vminu2(d[2 * j + 1], d[(2 * j + 2) & 15]);
#pragma unroll
  for (int j = 0; j < 8; ++j) s3[j] = __vimin3_u16x2(pr[j], pr[(j + 1) & 7], pr[(j + 2) & 7]);
#pragma unroll
  for (int j = 0; j < 8; ++j) u[j] = __vimin3_u16x2(s3[j], pr[(j + 3) & 7], __vmaxu2(d[2 * j], d[(2 * j + 9) & 15]));
  unsigned best = __vimax3_u16x2(u[0], u[1], u[2]);
  best = __vimax3_u16x2(best, u[3], u[4]);
  best = __vimax3_u16x2(best, u[5], u[6]);
  best = __vmaxu2(best, u[7]);
  return (int)max(best & 0xffffu, best >> 16) - 257;
}

// two KITTI-shaped entries of the queue per lane: the same min/max networks on packed u16x2 lanes
// (VIMNMX3.U16x2); differences are biased by +256 so that they stay positive 16-bit values
__device__ __forceinline__ void orbfe_fast_score3_x2(const uint8_t* p0, const uint8_t* p1, int tp, int& s0, int& s1) {
  const unsigned bias = (256u - p0[0]) | ((256u - p1[0]) << 16);
  unsigned d[16];
#define ORBFE_D(k, off) d[k] = ((unsigned)p0[off] | ((unsigned)p1[off] << 16)) + bias
  ORBFE_D(0, 3 * tp);       ORBFE_D(1, 3 * tp + 1);   ORBFE_D(2, 2 * tp + 2);   ORBFE_D(3, tp + 3);
  ORBFE_D(4, 3);            ORBFE_D(5, -tp + 3);      ORBFE_D(6, -2 * tp + 2);  ORBFE_D(7, -3 * tp + 1);
  ORBFE_D(8, -3 * tp);      ORBFE_D(9, -3 * tp - 1);  ORBFE_D(10, -2 * tp - 2); ORBFE_D(11, -tp - 3);
  ORBFE_D(12, -3);          ORBFE_D(13, tp - 3);      ORBFE_D(14, 2 * tp - 2);  ORBFE_D(15, 3 * tp - 1);
#undef ORBFE_D
  unsigned lo3[16], hi3[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    lo3[i] = __vimin3_u16x2(d[i], d[(i + 1) & 15], d[(i + 2) & 15]);
    hi3[i] = __vimax3_u16x2(d[i], d[(i + 1) & 15], d[(i + 2) & 15]);
  }
  unsigned bright = 0u, dark = 0xffffffffu;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    bright = __vmaxu2(bright, __vimin3_u16x2(lo3[i], lo3[(i + 3) & 15], lo3[(i + 6) & 15]));
    dark = __vminu2(dark, __vimax3_u16x2(hi3[i], hi3[(i + 3) & 15], hi3[(i + 6) & 15]));
  }
  s0 = max((int)(bright & 0xffffu) - 256, 256 - (int)(dark & 0xffffu)) - 1;
  s1 = max((int)(bright >> 16) - 256, 256 - (int)(dark >> 16)) - 1;
}

// PW = compile-time tile pitch in words (ring offsets become immediates); PW = 0: run-time pitch
// (cells wider than the fixed tile)
template <int PW>
__global__ void __launch_bounds__(ORBFE_FAST_THREADS)
k_fast_cells(const __grid_constant__ Geom g, const uint8_t* __restrict__ pyr, int* __restrict__ cellCnt,
             unsigned* __restrict__ cellList, const int pitchWArg, const int maxRows, const int queueCap) {
  const int pitchW = PW ? PW : pitchWArg;
  const int xbits = 9;  // queue code = y << xbits | x
  ORBFE_DYN_SMEM(smem);
  unsigned* tileW = reinterpret_cast<unsigned*>(smem);               // [maxRows][pitchW] pixels
  unsigned* scoreW = tileW + (size_t)maxRows * pitchW;               // [maxRows][pitchW] corner scores
  unsigned* bitsW = scoreW + (size_t)maxRows * pitchW;               // [maxRows][pitchW/8+1] NMS survivor bits
  unsigned short* queue = reinterpret_cast<unsigned short*>(bitsW + (size_t)maxRows * (pitchW / 8 + 1));
  __shared__ int s_qn;
  __shared__ int s_any[ORBFE_FAST_MAXG];
  __shared__ int s_fall[ORBFE_FAST_MAXG];
  __shared__ unsigned s_rowmask[ORBFE_FAST_MAXG][ORBFE_FAST_ROWWORDS];
  __shared__ uint8_t s_colCell[512];
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
  const int pitchB = pitchW * 4, bitsP = pitchW / 8 + 1;
  const int gx0 = (iniX + ORBFE_EDGE) & ~3;                          // padded-plane column of tile column 0
  const int tw = (maxX + ORBFE_EDGE - gx0 + 3) >> 2;                  // tile width in words
  const int ix0 = iniX + 3 + ORBFE_EDGE - gx0, ix1 = maxX - 3 + ORBFE_EDGE - gx0;  // inner columns (tile coords)
  const uint8_t* src = pyr + (size_t)slot * g.pyrStride + L.planeOff + (size_t)(iniY + ORBFE_EDGE) * L.pitch + gx0;
  for (int r = wid; r < rows; r += ORBFE_FAST_THREADS / 32) {
    const unsigned* srow = reinterpret_cast<const unsigned*>(src + (size_t)r * L.pitch);
    for (int c = lane; c < pitchW; c += 32) {
      tileW[r * pitchW + c] = c < tw ? __ldg(srow + c) : 0u;
      scoreW[r * pitchW + c] = 0u;
    }
    for (int c = lane; c < bitsP; c += 32) bitsW[r * bitsP + c] = 0u;
  }
  if (tid < ORBFE_FAST_MAXG) {
    s_any[tid] = 0;
#pragma unroll
    for (int k = 0; k < ORBFE_FAST_ROWWORDS; ++k) s_rowmask[tid][k] = 0u;
  }
  for (int x = tid; x < pitchB && x < 512; x += ORBFE_FAST_THREADS)
    s_colCell[x] = (uint8_t)((x >= ix0 && x < ix1) ? (x - ix0) / L.wCell : 0xff);
  const uint8_t* tileB = reinterpret_cast<const uint8_t*>(tileW);
  uint8_t* scoreB = reinterpret_cast<uint8_t*>(scoreW);
  const int w0 = ix0 >> 2, nWi = ((ix1 - 1) >> 2) - w0 + 1;

  for (int round = 0; round < 2; ++round) {
    const int th = round == 0 ? g.iniTh : g.minTh;
    if (tid == 0) s_qn = 0;
    if (tid < ORBFE_FAST_MAXG) s_fall[tid] = !s_any[tid];  // cells in play this round (snapshot: step 3 updates s_any)
    __syncthreads();
    if (round == 1) {  // block-uniform: does any cell of this CTA need the minThFAST fallback?
      bool need = false;
      for (int jl = 0; jl < nj; ++jl) need = need || !s_any[jl];
      if (!need || g.minTh >= g.iniTh) break;
    }
    // ---- 1. compass pre-test on words: flag bit 7 of byte b <=> |ring - centre| > th
    const unsigned K4 = (unsigned)(th < 128 ? 127 - th : 255 - th) * 0x01010101u;
    for (int y = 3 + wid; y < rows - 3; y += ORBFE_FAST_THREADS / 32)
      for (int wi = lane; wi < nWi; wi += 32) {
        const int wx = w0 + wi, xb = 4 * wx;
        if (round == 1) {  // only the fallback cells are re-examined
          const unsigned ca = s_colCell[min(max(xb, ix0), ix1 - 1)], cb = s_colCell[max(min(xb + 3, ix1 - 1), ix0)];
          if (s_any[ca] && s_any[cb]) continue;
        }
        const unsigned* row = tileW + y * pitchW + wx;
        const unsigned c = row[0], up = row[-3 * pitchW], dn = row[3 * pitchW];
        const unsigned prev = wx > 0 ? row[-1] : 0u, next = wx + 1 < pitchW ? row[1] : 0u;
        const unsigned lf = __funnelshift_r(prev, c, 8), rt = __funnelshift_r(c, next, 24);
        const unsigned a0 = __vabsdiffu4(dn, c), a4 = __vabsdiffu4(rt, c), a8 = __vabsdiffu4(up, c), a12 = __vabsdiffu4(lf, c);
        unsigned f0, f4, f8, f12;
        if (th < 128) {
          f0 = ((a0 & 0x7f7f7f7fu) + K4) | a0;   f4 = ((a4 & 0x7f7f7f7fu) + K4) | a4;
          f8 = ((a8 & 0x7f7f7f7fu) + K4) | a8;   f12 = ((a12 & 0x7f7f7f7fu) + K4) | a12;
        } else {
          f0 = ((a0 & 0x7f7f7f7fu) + K4) & a0;   f4 = ((a4 & 0x7f7f7f7fu) + K4) & a4;
          f8 = ((a8 & 0x7f7f7f7fu) + K4) & a8;   f12 = ((a12 & 0x7f7f7f7fu) + K4) & a12;
        }
        unsigned m = (f0 | f8) & (f4 | f12) & 0x80808080u;  // two adjacent compass points differ by > th
        if (m == 0u) continue;
        if (xb < ix0 || xb + 3 >= ix1 || round == 1) {  // edge words of the band / fallback round: per-byte filter
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const int x = xb + b;
            if (x < ix0 || x >= ix1 || (round == 1 && s_any[s_colCell[x]])) m &= ~(0x80u << (8 * b));
          }
        }
        const int n = __popc(m);
        if (n == 0) continue;
        int pos = atomicAdd(&s_qn, n);
#pragma unroll
        for (int b = 0; b < 4; ++b)
          if ((m >> (8 * b + 7)) & 1u) { if (pos < queueCap) queue[pos] = (unsigned short)((y << xbits) | (xb + b)); ++pos; }
      }
    __syncthreads();
    // ---- 2. exact score on the queue; corner at th <=> score >= th
    // The queue holds queueCap entries, sized so that 6 CTAs fit an SM rather than for the worst case (A/B: FAST -7 %).  A tile
    // whose pre-test passes more pixels than that (dense noise at minThFAST) is block-uniformly switched to the dense form of
    // steps 2 and 3: every pixel of the cells in play is scored and NMS-tested, no queue.  Identical results: the pre-test is
    // only a necessary condition, a pixel it rejects has score < th.
    const bool dense = s_qn > queueCap;
    const int qn = min(s_qn, queueCap);
    const int xmask = (1 << xbits) - 1;
    const int innerW = ix1 - ix0, nWork = dense ? innerW * (rows - 6) : qn;
    // work item e -> pixel (x, y) of the tile; false = not in play (dense form, cell already has keypoints)
    auto item = [&](const int e, int& x, int& y) -> bool {
      if (!dense) {
        const int code = queue[e];
        x = code & xmask;
        y = code >> xbits;
        return true;
      }
      const int yy = e / innerW;
      x = ix0 + (e - yy * innerW);
      y = 3 + yy;
      return round == 0 || s_fall[s_colCell[x]];
    };
    for (int e = tid; e < nWork; e += ORBFE_FAST_THREADS) {
      int x, y;
      if (!item(e, x, y)) continue;
      const int o0 = y * pitchB + x;
      const int s0 = orbfe_fast_score3(tileB + o0, pitchB);
      if (s0 >= th) scoreB[o0] = (uint8_t)s0;
    }
    __syncthreads();
    // ---- 3. NMS inside the cell; survivors -> bit plane + row masks; keypoint found => no fallback
    for (int e = tid; e < nWork; e += ORBFE_FAST_THREADS) {
      int x, y;
      if (!item(e, x, y)) continue;
      const uint8_t* c = scoreB + y * pitchB + x;
      const int s = c[0];
      if (s == 0) continue;
      const int jl = s_colCell[x];
      const int cx0 = ix0 + jl * L.wCell, cx1 = min(cx0 + L.wCell, ix1);
      const bool hasL = x > cx0, hasR = x + 1 < cx1;  // rows outside the inner band hold score 0 already
      bool keep = s > c[-pitchB] && s > c[pitchB];
      if (hasL) keep = keep && s > c[-1] && s > c[-pitchB - 1] && s > c[pitchB - 1];
      if (hasR) keep = keep && s > c[1] && s > c[-pitchB + 1] && s > c[pitchB + 1];
      if (keep) {
        atomicOr(&bitsW[y * bitsP + (x >> 5)], 1u << (x & 31));
        atomicOr(&s_rowmask[jl][y >> 5], 1u << (y & 31));
        s_any[jl] = 1;
      }
    }
    __syncthreads();
  }
  // ---- 4. ordered emission, one warp per cell, only rows that hold a survivor
  for (int jl = wid; jl < nj; jl += ORBFE_FAST_THREADS / 32) {
    const int cx0 = ix0 + jl * L.wCell, cx1 = min(cx0 + L.wCell, ix1);
    unsigned* out = list + (size_t)jl * L.cellCap;
    int base = 0;
    if (cx1 > cx0) {
#pragma unroll
      for (int k = 0; k < ORBFE_FAST_ROWWORDS; ++k) {
        unsigned rm = s_rowmask[jl][k];
        while (rm) {
          const int y = 32 * k + __ffs((int)rm) - 1;
          rm &= rm - 1;
          for (int c0 = cx0; c0 < cx1; c0 += 32) {
            const int x = c0 + lane;
            const bool emit = x < cx1 && ((bitsW[y * bitsP + (x >> 5)] >> (x & 31)) & 1u);
            const unsigned bal = __ballot_sync(0xffffffffu, emit);
            if (emit) {
              const int pos = base + __popc(bal & ((1u << lane) - 1u));
              if (pos < L.cellCap)
                out[pos] = orbfe_pack(x + gx0 - ORBFE_EDGE - ORBFE_MINB, y + iniY - ORBFE_MINB, scoreB[y * pitchB + x]);
            }
            base += __popc(bal);
          }
        }
      }
    }
    if (lane == 0) cnt[jl] = min(base, L.cellCap);
  }
}
