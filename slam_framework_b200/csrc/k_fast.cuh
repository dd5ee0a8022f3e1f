// k_fast.cuh -- E2+E3: grid FAST-9/16 with adaptive threshold, per-cell NMS and ordered select.
// Follows the grid loop of ComputeKeyPointsOctTree (orb_extractor.cpp:706-770) and
// cv::FAST(cell, kps, th, true):
//   * the corner score (max threshold at which the pixel is still a corner) does not depend on the threshold
//     passed in; cv::FAST at threshold t == "score >= t" on that score map;
//   * NMS is strict (>) over the 8 neighbours INSIDE the cell's detection window (outside counts 0);
//   * survivors with score >= iniThFAST are emitted; a cell with none re-runs at minThFAST (the post-NMS
//     emptiness fallback, :753-757);
//   * emission order is row-major inside the cell (the order cv::FAST returns); cells are consumed in row-major
//     cell order by the quad-tree kernel.
//
// One CTA = up to L.fG consecutive grid cells of one cell row of one level of one image.  The cells' detection
// windows tile the level's inner region exactly (origin (16,16), stride wCell x hCell), so the CTA stages ONE
// pixel tile, (hCell+6) rows x 256 bytes, with a single TMA box copy (orbfe_tma.cuh) and runs up to two rounds
// (iniThFAST on every cell, minThFAST on the cells still empty).  Every phase of a round is driven by a
// compacted work list, so that all 32 lanes of a warp are busy in the expensive steps:
//   1. 8-point pre-test, 4 pixels per instruction (VABSDIFF4 + carry trick), ONCE per tile for BOTH thresholds (the
//      absolute differences are shared): a 9-arc of the 16-ring contains at least one point of every antipodal pair,
//      so a pixel is a corner only if, in each of the pairs (0,8), (4,12), (2,10), (6,14), a ring point differs from
//      it by more than t.  Branch-free: the flags of up to 16 words accumulate in two registers per thread and
//      threshold;
//   2. one warp-aggregated reservation per warp, then each thread unpacks its flag bits into the candidate queue Q1
//      (round 1: only the bits of cells that found nothing in round 0).  Measured alternatives: building the queue
//      with per-item warp ballots (word queue + expansion) costs twice the instructions of the per-thread unpack;
//   3. exact corner score (max_arc min_9 of +-(ring - v), packed u16x2 min/max networks) on Q1; corners
//      (score >= t) go to the score plane and, by warp ballot, to the corner queue Q2;
//   4. strict 3x3 NMS restricted to the cell on Q2 -> per-cell survivor lists;
//   5. after both rounds, one warp per cell: a survivor's output slot is the number of survivors of its cell that
//      precede it in row-major order (counted against the cell's short list): ordered emission without a serial walk.
// A tile whose pre-test passes more pixels than Q1 holds (dense noise at minThFAST) switches, block-uniformly,
// to the dense form of steps 3 and 4 (every pixel of the cells in play): identical results, the pre-test is
// only a necessary condition.
#pragma once
#include "orbfe_common.cuh"
#include "orbfe_tma.cuh"

#ifndef ORBFE_FAST_WARPS
#define ORBFE_FAST_WARPS 8      // warps per CTA (A/B on B200, 64 pairs: see DESIGN.md)
#endif
#define ORBFE_FAST_THREADS (32 * ORBFE_FAST_WARPS)
#define ORBFE_FAST_ITS (64 / ORBFE_FAST_WARPS)   // row iterations of the pre-test: rows 3 + wid + WARPS * it cover 64 inner rows
#define ORBFE_FAST_NACC (ORBFE_FAST_ITS / 4)     // flag registers per threshold: 8 word items (4 row iterations x 2 columns) each
#define ORBFE_FAST_TP 256       // tile pitch in bytes == TMA box width
#define ORBFE_FAST_TPW 64       // ... in words
#define ORBFE_FAST_MAXG 8       // cells per CTA (upper bound)
// A TMA box must start on a 16-byte boundary of the plane row (measured on B200: any other start coordinate faults with
// "illegal instruction"), so the tile starts at the 16-byte boundary at or below the column 1 px left of the detection
// window: window column c is tile column c + 1 + off, off in [0, 15], and the first inner column is 4 + off.
#define ORBFE_FAST_MAXW 234     // fG * wCell <= 234  (16 + fG*wCell + 6 <= 256)
#define ORBFE_FAST_MAXROWS 72   // tile rows (hCell + 6 <= 65)
#ifndef ORBFE_FAST_MINB
#define ORBFE_FAST_MINB 5
#endif

// per-CTA work descriptor (host-built): level | cell row << 4 | first cell << 16
__host__ __device__ __forceinline__ unsigned orbfe_fast_task(int level, int i, int j0) {
  return (unsigned)level | ((unsigned)i << 4) | ((unsigned)j0 << 16);
}

// exact FAST-9/16 corner score of the pixel at p (shared-memory tile, byte pitch tp):
//   score + 1 = max( max_arc min_9 (ring - v),  max_arc min_9 (v - ring) ).
// Both halves run through ONE min/max network on packed u16x2 lanes (VIMNMX3.U16x2): the low half carries
// (ring - v) + 256, the high half (v - ring) + 256, both in [1, 511]; the packing is one IMAD per ring pixel:
//   (d + 256) | (256 - d) << 16  ==  C + ring * (1 - 65536)  with C folding the centre value.
__device__ __forceinline__ int orbfe_fast_score3(const uint8_t* p, const int tp) {
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
  // = 36 two-lane min/max instructions.
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

// bytes of tile word wx whose column lies in [c0, c1) -> 0x80 per byte
__device__ __forceinline__ unsigned orbfe_fast_colmask(const int wx, const int c0, const int c1) {
  const int lo = max(c0 - 4 * wx, 0), hi = min(c1 - 4 * wx, 4);
  if (hi <= lo) return 0u;
  return (0x80808080u >> (8 * (4 - hi))) & (0x80808080u << (8 * lo));
}

// warp-aggregated reservation of the warp's queue slots: returns the warp's first slot (warp-uniform)
__device__ __forceinline__ int orbfe_fast_reserve(const int cnt, int* counter, const int lane) {
  const int total = __reduce_add_sync(0xffffffffu, cnt);
  int base = 0;
  if (lane == 0 && total > 0) base = atomicAdd(counter, total);
  return __shfl_sync(0xffffffffu, base, 0);
}

struct FastSmemLayout {
  int scoreOff, q1Off, q2Off, listOff, total;  // bytes from the start of dynamic shared memory (tile at 0)
};
__host__ __device__ __forceinline__ FastSmemLayout orbfe_fast_layout(const int maxRows, const int q1Cap, const int q2Cap, const int listCap) {
  FastSmemLayout s;
  s.scoreOff = maxRows * ORBFE_FAST_TP + 16;            // + 16: the pre-test of the last word reads one word past the tile
  s.q1Off = s.scoreOff + maxRows * ORBFE_FAST_TP;
  s.q2Off = s.q1Off + ((q1Cap * 2 + 15) & ~15);
  s.listOff = s.q2Off + ((q2Cap * 2 + 15) & ~15);
  s.total = s.listOff + ((ORBFE_FAST_MAXG * listCap * 2 + 15) & ~15);
  return s;
}

// step 2: a flag register of a thread (items kBase .. kBase + 7; bit 8b + 7 - (k & 7) = byte b of item k; item k = 2 * it + j
// is the word (row 3 + wid + WARPS * it, column wLo + lane + 32 j)) -> queue codes y << 8 | x from the warp's slot `pos` on.
// ROUND-MAJOR order: in every pass each lane that still has a flag contributes ONE entry, so that 32 consecutive queue
// entries come from (nearly) 32 different lanes = 32 different word columns = 32 different shared-memory banks.  The score
// step gathers 16 ring bytes per entry; with the tile pitch a multiple of 128 bytes the bank of a pixel depends on its column
// only, and the lane-major order this replaces (all flags of lane 0, then lane 1, ...) put up to 8 pixels of one column into
// one warp of the score loop (47 M bank conflicts per 128 frames, shared-memory pipe at 67 %).  Returns the warp's next slot.
__device__ __forceinline__ int orbfe_fast_unpack(unsigned acc, const int kBase, int pos, const int cap, unsigned short* q,
                                                 const unsigned baseCode, const unsigned ltMask) {
  unsigned any;
  while ((any = __ballot_sync(0xffffffffu, acc != 0u)) != 0u) {
    if (acc) {
      const int p = __ffs((int)acc) - 1;
      acc &= acc - 1;
      const int k = kBase + 7 - (p & 7);
      const int slot = pos + __popc(any & ltMask);
      if (slot < cap) q[slot] = (unsigned short)(baseCode + (k >> 1) * (ORBFE_FAST_WARPS << 8) + ((k & 1) << 7) + (p >> 3));
    }
    pos += __popc(any);
  }
  return pos;
}

// LOW: both thresholds are below 128 (the usual case; the other instance handles any thresholds)
template <bool LOW>
__global__ void __launch_bounds__(ORBFE_FAST_THREADS, ORBFE_FAST_MINB)
k_fast_cells(const __grid_constant__ Geom g, const uint8_t* __restrict__ pyr, const CUtensorMap* __restrict__ tmaps,
             const unsigned* __restrict__ tasks, int* __restrict__ cellCnt, unsigned* __restrict__ cellList,
             const int maxRows, const int q1Cap, const int q2Cap, const int listCap) {
  ORBFE_DYN_SMEM(smem);
  const FastSmemLayout lay = orbfe_fast_layout(maxRows, q1Cap, q2Cap, listCap);
  unsigned* tileW = reinterpret_cast<unsigned*>(smem);
  const uint8_t* tileB = smem;
  uint8_t* scoreB = smem + lay.scoreOff;
  unsigned short* q1 = reinterpret_cast<unsigned short*>(smem + lay.q1Off);
  unsigned short* q2 = reinterpret_cast<unsigned short*>(smem + lay.q2Off);
  unsigned short* lists = reinterpret_cast<unsigned short*>(smem + lay.listOff);  // [cell][listCap] survivors y << 8 | x
  __shared__ __align__(8) unsigned long long s_bar;
  __shared__ int s_qn[2][2];                         // [round][Q1 / Q2] entries pushed
  __shared__ int s_cellN[ORBFE_FAST_MAXG];           // survivors per cell
  __shared__ uint8_t s_colCell[ORBFE_FAST_TP];       // tile column -> cell of this CTA (0xff outside the inner band)
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int slot = blockIdx.y;
  const unsigned task = __ldg(tasks + blockIdx.x);
  const int level = task & 15, i = (task >> 4) & 0xfff, j0 = task >> 16;
  const LevelGeom& L = g.lv[level];
  const int wCell = L.wCell;
  const int nj = min(L.fG, L.nCols - j0);
  int* cnt = cellCnt + (size_t)slot * g.totalCells + L.cellBase + i * L.nCols + j0;
  const int iniY = ORBFE_MINB + i * L.hCell, iniX = ORBFE_MINB + j0 * wCell;
  const int rows = min(iniY + L.hCell + 6, L.maxBY) - iniY;
  const int winW = min(iniX + nj * wCell + 6, L.maxBX) - iniX;
  // skipped rows/cells (orb_extractor.cpp:735,744) and sub-images too small for FAST yield nothing
  if (rows < 7 || winW < 7) {
    if (tid < nj) cnt[tid] = 0;
    return;
  }
  const int gx0 = (iniX + ORBFE_EDGE - 1) & ~15;   // padded-plane column of tile column 0
  const int off = iniX + ORBFE_EDGE - 1 - gx0;     // window column c -> tile column c + 1 + off
  const int ix0 = 4 + off, ix1 = winW - 2 + off;   // inner columns (tile coordinates)
  if (tid == 0) {
    orbfe_tile_barrier_init(&s_bar);
    s_qn[0][0] = s_qn[0][1] = s_qn[1][0] = s_qn[1][1] = 0;
  }
  if (tid < ORBFE_FAST_MAXG) s_cellN[tid] = 0;
  __syncthreads();
  if (tid == 0) {
    OrbfeTmaPlane P;
    P.base = pyr + L.planeOff; P.sliceStride = g.pyrStride; P.pitch = L.pitch; P.rows = L.h + 2 * ORBFE_EDGE;
    P.slices = gridDim.y; P.boxW = ORBFE_FAST_TP; P.boxH = maxRows;
    orbfe_tmap_acquire(tmaps + level);
    orbfe_tile_issue(smem, &s_bar, tmaps + level, P, gx0, iniY + ORBFE_EDGE, slot);
  }
  {  // while the tile is in flight: clear the score plane, build the column -> cell table
    uint4* z = reinterpret_cast<uint4*>(scoreB);
    for (int k = tid; k < rows * (ORBFE_FAST_TP / 16); k += ORBFE_FAST_THREADS) z[k] = make_uint4(0u, 0u, 0u, 0u);
    // (x - ix0) / wCell by multiply-shift: exact for x - ix0 < 256 and wCell <= 234
    for (int x = tid; x < ORBFE_FAST_TP; x += ORBFE_FAST_THREADS)
      s_colCell[x] = (uint8_t)((x >= ix0 && x < ix1) ? ((unsigned)(x - ix0) * (unsigned)L.wCellMagic) >> 16 : 0xffu);
  }
  const int yEnd = rows - 3;                       // inner rows [3, yEnd)
  const int wLo = ix0 >> 2, nW = (ix1 - 1) >> 2;   // inner words wLo .. nW (wLo >= 1: the pre-test reads word wx - 1)
  const unsigned mask0 = orbfe_fast_colmask(wLo + lane, ix0, ix1), mask1 = orbfe_fast_colmask(wLo + lane + 32, ix0, ix1);
  const int thA = g.iniTh, thB = g.minTh;
  const bool twoRounds = thB < thA;
  const unsigned KA = (unsigned)(thA < 128 ? 127 - thA : 255 - thA) * 0x01010101u;
  const unsigned KB = (unsigned)(thB < 128 ? 127 - thB : 255 - thB) * 0x01010101u;
  const unsigned baseCode = (unsigned)(((3 + wid) << 8) | (4 * (wLo + lane)));
  orbfe_tile_wait(&s_bar, 0);
  __syncthreads();

  // ---- 1. pre-test at both thresholds -> flag registers
  unsigned fa[ORBFE_FAST_NACC], fb[ORBFE_FAST_NACC];  // a: iniThFAST, b: minThFAST
#pragma unroll
  for (int q = 0; q < ORBFE_FAST_NACC; ++q) fa[q] = fb[q] = 0u;
#pragma unroll
  for (int it = 0; it < ORBFE_FAST_ITS; ++it) {
    const int y = 3 + wid + ORBFE_FAST_WARPS * it;
    if (y < yEnd) {                                // warp-uniform
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int wx = wLo + lane + 32 * j;
        if (wx <= nW) {
          const unsigned* row = tileW + y * ORBFE_FAST_TPW + wx;
          const unsigned c = row[0];
          const unsigned r0 = row[3 * ORBFE_FAST_TPW], r8 = row[-3 * ORBFE_FAST_TPW];                      // ring 0 (0,3), 8 (0,-3)
          const unsigned r12 = __funnelshift_r(row[-1], c, 8), r4 = __funnelshift_r(c, row[1], 24);      // ring 12 (-3,0), 4 (3,0)
          const unsigned* rd = row + 2 * ORBFE_FAST_TPW;
          const unsigned* ru = row - 2 * ORBFE_FAST_TPW;
          const unsigned dc = rd[0], uc = ru[0];
          const unsigned r2 = __funnelshift_r(dc, rd[1], 16), r14 = __funnelshift_r(rd[-1], dc, 16);     // ring 2 (2,2), 14 (-2,2)
          const unsigned r6 = __funnelshift_r(uc, ru[1], 16), r10 = __funnelshift_r(ru[-1], uc, 16);     // ring 6 (2,-2), 10 (-2,-2)
          const unsigned d0 = __vabsdiffu4(r0, c), d4 = __vabsdiffu4(r4, c), d8 = __vabsdiffu4(r8, c), d12 = __vabsdiffu4(r12, c);
          const unsigned d2 = __vabsdiffu4(r2, c), d6 = __vabsdiffu4(r6, c), d10 = __vabsdiffu4(r10, c), d14 = __vabsdiffu4(r14, c);
          const unsigned l0 = d0 & 0x7f7f7f7fu, l4 = d4 & 0x7f7f7f7fu, l8 = d8 & 0x7f7f7f7fu, l12 = d12 & 0x7f7f7f7fu;
          const unsigned l2 = d2 & 0x7f7f7f7fu, l6 = d6 & 0x7f7f7f7fu, l10 = d10 & 0x7f7f7f7fu, l14 = d14 & 0x7f7f7f7fu;
          const unsigned mk = j ? mask1 : mask0;
          // |d| > th:  th < 128: (|d| & 127) + (127 - th) carries into bit 7, or bit 7 of |d| is set;
          //            th >= 128: bit 7 of |d| is set AND the low 7 bits exceed th - 128
#define ORBFE_GT_LO(l, d, K) (((l) + (K)) | (d))
#define ORBFE_GT_HI(l, d, K) (((l) + (K)) & (d))
#define ORBFE_PAIRS(GT, K)                                                                                                      \
  ((GT(l0, d0, K) | GT(l8, d8, K)) & (GT(l4, d4, K) | GT(l12, d12, K)) & (GT(l2, d2, K) | GT(l10, d10, K)) & (GT(l6, d6, K) | GT(l14, d14, K)) & mk)
          unsigned mA, mB = 0u;
          if (LOW || thA < 128) mA = ORBFE_PAIRS(ORBFE_GT_LO, KA);
          else mA = ORBFE_PAIRS(ORBFE_GT_HI, KA);
          if (twoRounds) {
            if (LOW || thB < 128) mB = ORBFE_PAIRS(ORBFE_GT_LO, KB);
            else mB = ORBFE_PAIRS(ORBFE_GT_HI, KB);
          }
#undef ORBFE_PAIRS
#undef ORBFE_GT_LO
#undef ORBFE_GT_HI
          fa[it >> 2] |= mA >> (2 * (it & 3) + j);
          fb[it >> 2] |= mB >> (2 * (it & 3) + j);
        }
      }
    }
  }

  unsigned fallMask = 0u;                          // round 1: cells in play
  for (int round = 0; round < 2; ++round) {
    const int th = round == 0 ? thA : thB;
    int* qn1 = &s_qn[round][0];
    int* qn2 = &s_qn[round][1];
    // ---- 2. flag registers -> Q1
    if (round == 1) {
      if (!twoRounds) break;
      for (int jl = 0; jl < nj; ++jl) fallMask |= (s_cellN[jl] ? 0u : 1u) << jl;
      if (fallMask == 0u) break;                   // block-uniform
      // keep the flags of the cells in play: byte b of the items with j = 0 sits in bits 8b + {7,5,3,1}, j = 1 in 8b + {6,4,2,0}
      unsigned keep = 0u;
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int x = 4 * (wLo + lane + 32 * j) + b;
          const unsigned c = x < ORBFE_FAST_TP ? s_colCell[x] : 0xffu;
          if (c != 0xffu && ((fallMask >> c) & 1u)) keep |= (j ? 0x55u : 0xaau) << (8 * b);
        }
#pragma unroll
      for (int q = 0; q < ORBFE_FAST_NACC; ++q) fa[q] = fb[q] & keep;
    }
    {
      int cntA = 0;
#pragma unroll
      for (int q = 0; q < ORBFE_FAST_NACC; ++q) cntA += __popc(fa[q]);
      int pos = orbfe_fast_reserve(cntA, qn1, lane);
#pragma unroll
      for (int q = 0; q < ORBFE_FAST_NACC; ++q) pos = orbfe_fast_unpack(fa[q], 8 * q, pos, q1Cap, q1, baseCode, (1u << lane) - 1u);
    }
    __syncthreads();
    // ---- 3. exact score; corner at th <=> score >= th.  Corners -> score plane + Q2 (warp ballot)
    const bool dense = *qn1 > q1Cap;
    if (!dense) {
      const int qn = *qn1;
      for (int e0 = wid * 32; e0 < qn; e0 += ORBFE_FAST_THREADS) {
        const int e = e0 + lane;
        bool corner = false;
        unsigned short code = 0;
        if (e < qn) {
          code = q1[e];
          const int o = (code >> 8) * ORBFE_FAST_TP + (code & 255);
          const int s = orbfe_fast_score3(tileB + o, ORBFE_FAST_TP);
          corner = s >= th;
          if (corner) scoreB[o] = (uint8_t)s;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, corner);
        if (bal) {
          int base = 0;
          if (lane == 0) base = atomicAdd(qn2, __popc(bal));
          base = __shfl_sync(0xffffffffu, base, 0);
          const int pos = base + __popc(bal & ((1u << lane) - 1u));
          if (corner && pos < q2Cap) q2[pos] = code;
        }
      }
    } else {
      const int innerW = ix1 - ix0, nWork = innerW * (yEnd - 3);
      for (int e = tid; e < nWork; e += ORBFE_FAST_THREADS) {
        const int yy = e / innerW, x = ix0 + (e - yy * innerW), y = 3 + yy;
        if (round == 1 && !((fallMask >> s_colCell[x]) & 1u)) continue;
        const int o = y * ORBFE_FAST_TP + x;
        const int s = orbfe_fast_score3(tileB + o, ORBFE_FAST_TP);
        if (s >= th) scoreB[o] = (uint8_t)s;
      }
    }
    __syncthreads();
    // ---- 4. NMS inside the cell; survivors -> the cell's list
    const bool denseNms = dense || *qn2 > q2Cap;
    auto nms = [&](const int x, const int y) {
      const uint8_t* c = scoreB + y * ORBFE_FAST_TP + x;
      const int s = c[0];
      const int jl = s_colCell[x];
      const int cx0 = ix0 + jl * wCell, cx1 = min(cx0 + wCell, ix1);
      // rows outside the inner band hold score 0; columns outside the cell's window count 0
      const int l0 = x > cx0 ? -1 : 0, r0 = x + 1 < cx1 ? 1 : 0;
      int m = max((int)c[-ORBFE_FAST_TP], (int)c[ORBFE_FAST_TP]);
      m = __vimax3_s32(m, (int)c[l0 - ORBFE_FAST_TP], (int)c[l0 + ORBFE_FAST_TP]);
      m = __vimax3_s32(m, (int)c[r0 - ORBFE_FAST_TP], (int)c[r0 + ORBFE_FAST_TP]);
      if (l0) m = max(m, (int)c[-1]);
      if (r0) m = max(m, (int)c[1]);
      if (s > m) {
        const int pos = atomicAdd(&s_cellN[jl], 1);
        if (pos < listCap) lists[jl * listCap + pos] = (unsigned short)((y << 8) | x);
      }
    };
    if (!denseNms) {
      const int qn = *qn2;
      for (int e = tid; e < qn; e += ORBFE_FAST_THREADS) {
        const unsigned code = q2[e];
        nms(code & 255, code >> 8);
      }
    } else {
      const int nWw = nW - wLo + 1, nWork = nWw * (yEnd - 3);
      for (int e = tid; e < nWork; e += ORBFE_FAST_THREADS) {
        const int yy = e / nWw, wx = wLo + (e - yy * nWw), y = 3 + yy;
        const unsigned s4 = *reinterpret_cast<const unsigned*>(scoreB + y * ORBFE_FAST_TP + 4 * wx);
        if (s4 == 0u) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const int x = 4 * wx + b;
          if (((s4 >> (8 * b)) & 0xffu) == 0u) continue;
          if (round == 1 && !((fallMask >> s_colCell[x]) & 1u)) continue;  // round-0 corners were handled in round 0
          nms(x, y);
        }
      }
    }
    __syncthreads();
  }
  // ---- 5. ordered emission, one warp per cell: slot = number of the cell's survivors that precede this one in
  // row-major order (key y << 8 | x); the lists are short (a few tens of entries)
  for (int jl = wid; jl < nj; jl += ORBFE_FAST_THREADS / 32) {
    const int n = min(s_cellN[jl], min(listCap, L.cellCap));
    const unsigned short* li = lists + jl * listCap;
    unsigned* out = cellList + (size_t)slot * g.cellListStride + L.cellListOff + (size_t)(i * L.nCols + j0 + jl) * L.cellCap;
    for (int e = lane; e < n; e += 32) {
      const unsigned key = li[e];
      int rank = 0;
      for (int f = 0; f < n; ++f) rank += li[f] < key ? 1 : 0;
      const int x = key & 255, y = key >> 8;
      out[rank] = orbfe_pack(x - 1 - off + iniX - ORBFE_MINB, y + iniY - ORBFE_MINB, scoreB[y * ORBFE_FAST_TP + x]);
    }
    if (lane == 0) cnt[jl] = n;
  }
}
