// k_match.cuh -- M0-M4: the Frame feature grid and the windowed Hamming searches of OrbMatcher.
//
//   k_grid_build      Frame::AssignFeaturesToGrid / PosInGrid (frame.cpp:234-248, 339-346): 64x48 cells,
//                     cell lists stored ix-major so that the cells (ix, iyMin..iyMax) of a query window
//                     are ONE contiguous item range already in the reference's visiting order
//                     (ix outer, iy inner, ascending keypoint index inside a cell).
//   k_match_candidates (phase A, one warp per query, fully parallel): Frame::GetFeaturesInArea
//                     (frame.cpp:348-403) + the static filters of the search routine (octave range,
//                     stereo-right consistency) + DescriptorDistance (orb_matcher.cpp:1630-1646, 8 x
//                     __popc on 32-byte rows).  Writes each query's candidates IN REFERENCE ORDER.
//   k_match_resolve   (phase B, one warp): walks the queries in the reference's serial order and
//                     applies the state that couples them (vMatchedDistance in
//                     SearchForInitialization :306-307, F.SetMapPoint feedback in SearchByProjection
//                     :59-63/:97); lanes scan a query's candidates, best / second-best are reduced with
//                     warp shuffles on the composite key (distance, position), which reproduces the
//                     strict '<' scans of the reference exactly; then the rotation-histogram
//                     consistency check (ComputeThreeMaxima :1584-1625).
#pragma once
#include "orbfe_common.cuh"
#ifndef ORBFE_EMU
#include <cooperative_groups.h>
#endif

#define ORBFE_GRID_COLS 64   // frame.h:104
#define ORBFE_GRID_ROWS 48   // frame.h:105
#define ORBFE_GRID_CELLS (ORBFE_GRID_COLS * ORBFE_GRID_ROWS)
#define ORBFE_HISTO_LENGTH 30  // orb_matcher.cpp:7

struct MatchKp {  // what the matchers read of an undistorted cv::KeyPoint
  float x, y, angle;
  int octave;
};

struct FrameGrid {
  const MatchKp* kp;
  const uint8_t* desc;      // n x 32
  const float* uR;          // n (stereo right coordinate, <= 0: none)
  const int* cellStart;     // ORBFE_GRID_CELLS + 1, ix-major
  const int* cellItems;     // keypoint indices
  const float* lvl;         // 3 x nlevels: scale_factors, level_sigma_sq, inv_level_sigma_sq (orb_extractor.cpp:356-372)
  int n, nlevels;
  float minX, minY, gw, gh;
};

struct MatchQueries {
  const float* x;           // window centre
  const float* y;
  const float* r;           // window half-size
  const float* xr;          // predicted right coordinate (stereo consistency), may be null
  const int* minLevel;
  const int* maxLevel;
  const uint8_t* valid;
  const uint8_t* desc;      // n x 32
  int n;
  int filter;               // per-candidate filter: 0 none, 1 stereo-right consistency (:65-70, :1404-1411), 2 Fuse chi2 (:893-917)
};
enum { ORBFE_FILTER_NONE = 0, ORBFE_FILTER_UR = 1, ORBFE_FILTER_FUSE = 2 };

struct MatchScratch {
  uint2* cand;              // x = keypoint index, y = distance | octave << 16
  int* qOff;
  int* qCnt;
  int* cursor;              // [0] allocation cursor, [1] overflow flag
  int capacity;
};

// ---- M0: grid build (one CTA) -------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
k_grid_build(const MatchKp* __restrict__ kp, const int n, const float minX, const float minY, const float gw,
             const float gh, int* __restrict__ cellStart, int* __restrict__ cellItems) {
  __shared__ int s_cnt[ORBFE_GRID_CELLS];
  __shared__ int s_scan[33];
  const int tid = threadIdx.x, T = blockDim.x;
  for (int c = tid; c < ORBFE_GRID_CELLS; c += T) s_cnt[c] = 0;
  __syncthreads();
  for (int i = tid; i < n; i += T) {
    const int px = (int)roundf(__fdiv_rn(__fsub_rn(kp[i].x, minX), gw));
    const int py = (int)roundf(__fdiv_rn(__fsub_rn(kp[i].y, minY), gh));
    if (px >= 0 && px < ORBFE_GRID_COLS && py >= 0 && py < ORBFE_GRID_ROWS) atomicAdd(&s_cnt[px * ORBFE_GRID_ROWS + py], 1);
  }
  __syncthreads();
  // exclusive scan over the cells: each thread owns 3 consecutive cells (3072 = 1024*3)
  const int per = (ORBFE_GRID_CELLS + T - 1) / T;
  const int c0 = min(tid * per, ORBFE_GRID_CELLS), c1 = min(c0 + per, ORBFE_GRID_CELLS);
  int sum = 0;
  for (int c = c0; c < c1; ++c) sum += s_cnt[c];
  int total;
  int run = orbfe_block_exscan(sum, s_scan, &total);
  for (int c = c0; c < c1; ++c) { const int v = s_cnt[c]; cellStart[c] = run; s_cnt[c] = run; run += v; }
  if (tid == 0) cellStart[ORBFE_GRID_CELLS] = total;
  __syncthreads();
  for (int i = tid; i < n; i += T) {
    const int px = (int)roundf(__fdiv_rn(__fsub_rn(kp[i].x, minX), gw));
    const int py = (int)roundf(__fdiv_rn(__fsub_rn(kp[i].y, minY), gh));
    if (px >= 0 && px < ORBFE_GRID_COLS && py >= 0 && py < ORBFE_GRID_ROWS)
      cellItems[atomicAdd(&s_cnt[px * ORBFE_GRID_ROWS + py], 1)] = i;
  }
  __syncthreads();
  // push_back order = ascending keypoint index: insertion-sort each (short) cell list
  for (int c = tid; c < ORBFE_GRID_CELLS; c += T) {
    const int b = cellStart[c], e = s_cnt[c];
    for (int i = b + 1; i < e; ++i) {
      const int v = cellItems[i];
      int j = i - 1;
      while (j >= b && cellItems[j] > v) { cellItems[j + 1] = cellItems[j]; --j; }
      cellItems[j + 1] = v;
    }
  }
}

// extractor output records (cv::KeyPoint layout, 7 x 4 bytes) -> the matcher's view, on the device
__global__ void __launch_bounds__(256)
k_kp_to_match(const float* __restrict__ kp7, const int n, MatchKp* __restrict__ out, const float* __restrict__ uRin, float* __restrict__ uRout) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float* k = kp7 + (size_t)i * 7;
  MatchKp m;
  m.x = k[0]; m.y = k[1]; m.angle = k[3]; m.octave = __float_as_int(k[5]);
  out[i] = m;
  uRout[i] = uRin ? uRin[i] : -1.0f;
}

// cell range of a query window (frame.cpp:356-369); returns false if the window misses the grid
__device__ __forceinline__ bool orbfe_window_cells(const FrameGrid& F, float x, float y, float r, int& x0, int& x1, int& y0,
                                                   int& y1) {
  x0 = max(0, (int)floorf(__fdiv_rn(__fsub_rn(__fsub_rn(x, F.minX), r), F.gw)));
  x1 = min(ORBFE_GRID_COLS - 1, (int)ceilf(__fdiv_rn(__fadd_rn(__fsub_rn(x, F.minX), r), F.gw)));
  if (x1 < 0 || x0 >= ORBFE_GRID_COLS) return false;
  y0 = max(0, (int)floorf(__fdiv_rn(__fsub_rn(__fsub_rn(y, F.minY), r), F.gh)));
  y1 = min(ORBFE_GRID_ROWS - 1, (int)ceilf(__fdiv_rn(__fadd_rn(__fsub_rn(y, F.minY), r), F.gh)));
  if (y1 < 0 || y0 >= ORBFE_GRID_ROWS) return false;
  return true;
}

// ---- phase A: ordered candidates + distances, one warp per query ----------------------------------
#define ORBFE_MATCH_THREADS 128

// one query, one warp (q < Q.n, warp-uniform)
__device__ __forceinline__ void orbfe_candidates_query(const FrameGrid& F, const MatchQueries& Q, const MatchScratch& S, const int q,
                                                       const int lane) {
  int x0 = 0, x1 = -1, y0 = 0, y1 = -1;
  const float x = Q.x[q], y = Q.y[q], r = Q.r[q];
  const bool live = Q.valid[q] && orbfe_window_cells(F, x, y, r, x0, x1, y0, y1);
  if (!live) {
    if (lane == 0) { S.qOff[q] = 0; S.qCnt[q] = 0; }
    return;
  }
  // upper bound of the list = items of the window's cells; one allocation per query
  int bound = 0;
  for (int ix = x0 + lane; ix <= x1; ix += 32)
    bound += F.cellStart[ix * ORBFE_GRID_ROWS + y1 + 1] - F.cellStart[ix * ORBFE_GRID_ROWS + y0];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bound += __shfl_xor_sync(0xffffffffu, bound, o);
  int base = 0;
  if (lane == 0) base = atomicAdd(&S.cursor[0], bound);
  base = __shfl_sync(0xffffffffu, base, 0);
  if (base + bound > S.capacity) {  // host grows the buffer and re-runs
    if (lane == 0) { S.cursor[1] = 1; S.qOff[q] = 0; S.qCnt[q] = 0; }
    return;
  }
  const int minL = Q.minLevel[q], maxL = Q.maxLevel[q];
  const bool checkLevels = (minL > 0) || (maxL >= 0);  // frame.cpp:372 (quirk kept)
  const float xr = Q.filter ? Q.xr[q] : 0.f;
  const float* invSigma2 = F.lvl + 2 * F.nlevels;
  const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(Q.desc + (size_t)q * 32));
  const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(Q.desc + (size_t)q * 32) + 1);
  int count = 0;
  for (int ix = x0; ix <= x1; ++ix) {
    const int b = F.cellStart[ix * ORBFE_GRID_ROWS + y0], e = F.cellStart[ix * ORBFE_GRID_ROWS + y1 + 1];
    for (int j0 = b; j0 < e; j0 += 32) {
      const int j = j0 + lane;
      bool ok = j < e;
      int idx = 0, oct = 0;
      if (ok) {
        idx = F.cellItems[j];
        const MatchKp k = F.kp[idx];
        oct = k.octave;
        if (checkLevels && (oct < minL || (maxL >= 0 && oct > maxL))) ok = false;
        const float dx = __fsub_rn(k.x, x), dy = __fsub_rn(k.y, y);
        if (!(fabsf(dx) < r && fabsf(dy) < r)) ok = false;
        if (ok && Q.filter == ORBFE_FILTER_UR) {  // orb_matcher.cpp:65-70 / :1404-1411
          const float u = F.uR[idx];
          if (u > 0 && fabsf(__fsub_rn(xr, u)) > r) ok = false;
        } else if (ok && Q.filter == ORBFE_FILTER_FUSE) {  // reprojection error gate of Fuse (orb_matcher.cpp:893-917)
          const float kpr = F.uR[idx];
          const float ex = __fsub_rn(x, k.x), ey = __fsub_rn(y, k.y);
          float e2 = __fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey));
          double lim = 5.99;
          if (kpr >= 0) {
            const float er = __fsub_rn(xr, kpr);
            e2 = __fadd_rn(e2, __fmul_rn(er, er));
            lim = 7.8;
          }
          if ((double)__fmul_rn(e2, invSigma2[oct]) > lim) ok = false;
        }
      }
      const unsigned bal = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(F.desc + (size_t)idx * 32));
        const uint4 b1 = __ldg(reinterpret_cast<const uint4*>(F.desc + (size_t)idx * 32) + 1);
        const int dist = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
                         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
        S.cand[base + count + __popc(bal & ((1u << lane) - 1u))] = make_uint2((unsigned)idx, (unsigned)dist | ((unsigned)oct << 16));
      }
      count += __popc(bal);
    }
  }
  if (lane == 0) { S.qOff[q] = base; S.qCnt[q] = count; }
}

__global__ void __launch_bounds__(ORBFE_MATCH_THREADS)
k_match_candidates(const FrameGrid F, const MatchQueries Q, const MatchScratch S) {
  const int q = blockIdx.x * (ORBFE_MATCH_THREADS / 32) + (threadIdx.x >> 5);
  if (q >= Q.n) return;
  orbfe_candidates_query(F, Q, S, q, threadIdx.x & 31);
}

// ---- phase B: serial-order resolve, one warp ------------------------------------------------------
enum { ORBFE_MODE_INIT = 0, ORBFE_MODE_MAPPOINTS = 1, ORBFE_MODE_LASTFRAME = 2, ORBFE_MODE_BOW = 3, ORBFE_MODE_GENERIC = 4 };
enum { ORBFE_FEEDBACK_NONE = 0, ORBFE_FEEDBACK_HASOBS = 1, ORBFE_FEEDBACK_ALL = 2 };   // who occupies a keypoint once accepted
enum { ORBFE_RATIO_NONE = 0, ORBFE_RATIO_SAMELEVEL = 1, ORBFE_RATIO_BOW = 2 };         // second-best test

struct ResolveArgs {
  int mode;
  int nQ, nKp;               // queries; keypoints of the searched frame
  float nnratio;
  int checkOri;
  const uint8_t* hasObs;      // per query (modes 1, 2)
  const uint8_t* occupiedIn;  // per keypoint (modes 1, 2)
  const float* qAngle;        // per query: angle of the query keypoint (modes 0, 2)
  const MatchKp* kp;          // searched frame keypoints (angles)
  int* out;                   // mode 0: matches12[nQ]; modes 1, 2: assigned[nKp]
  int* evBin;                 // per query: histogram bin of its acceptance, or -1
  int* evIdx;                 // per query: mode 0 -> query index, mode 2 -> keypoint index
  int* result;                // [0] nmatches
  // parallel resolve (k_match_iterate / k_match_finalize), set per routine by the host:
  int thAccept;               // accept bestDist <= thAccept (TH_HIGH 100, TH_LOW 50, ORBdist, ...)
  int ratio;                  // ORBFE_RATIO_*
  int feedback;               // ORBFE_FEEDBACK_*
  int tieLast;                // equal distances: the LAST candidate wins (SearchForTriangulation's 'dist>bestDist' skip, :717)
  int perQuery;               // out[] is indexed by query (value = keypoint) instead of by keypoint (value = query)
};

// 2 smallest of the union of two sorted pairs
__device__ __forceinline__ void orbfe_merge2(unsigned& b, unsigned& s, unsigned ob, unsigned os) {
  const unsigned nb = min(b, ob);
  const unsigned ns = min(max(b, ob), min(s, os));
  b = nb; s = ns;
}

__global__ void __launch_bounds__(32)
k_match_resolve(const ResolveArgs A, const MatchScratch S) {
  ORBFE_DYN_SMEM(smem);
  int* s_state = reinterpret_cast<int*>(smem);  // mode 0: vMatchedDistance[nKp] then vnMatches21[nKp]; else occupied[nKp]
  __shared__ int s_hist[ORBFE_HISTO_LENGTH];
  const int lane = threadIdx.x;
  if (S.cursor[1]) return;  // candidate buffer overflowed: the host re-runs with a larger one
  const int nKp = A.nKp;
  if (A.mode == ORBFE_MODE_INIT) {
    for (int i = lane; i < nKp; i += 32) { s_state[i] = 0x7fffffff; s_state[nKp + i] = -1; }
    for (int i = lane; i < A.nQ; i += 32) A.out[i] = -1;
  } else {
    for (int i = lane; i < nKp; i += 32) { s_state[i] = A.occupiedIn[i]; A.out[i] = -1; }
  }
  for (int i = lane; i < ORBFE_HISTO_LENGTH; i += 32) s_hist[i] = 0;
  for (int i = lane; i < A.nQ; i += 32) A.evBin[i] = -1;
  __syncwarp();
  int nmatches = 0;
  const float factor = 1.0f / ORBFE_HISTO_LENGTH;  // orb_matcher.cpp:275 (the reference's bin-width bug, kept)
  for (int q = 0; q < A.nQ; ++q) {
    const int cnt = S.qCnt[q];
    if (cnt == 0) continue;
    const int off = S.qOff[q];
    // key = dist << 20 | position (position < 2^20); sentinel = "none"
    unsigned best = 0xffffffffu, second = 0xffffffffu;
    for (int c = lane; c < cnt; c += 32) {
      const uint2 cd = S.cand[off + c];
      const int dist = (int)(cd.y & 0xffffu);
      bool ok;
      if (A.mode == ORBFE_MODE_INIT) ok = !(s_state[cd.x] <= dist);  // vMatchedDistance[i2] <= dist => skip (:306)
      else ok = !s_state[cd.x] && dist < 256;                        // occupied; bestDist starts at 256
      if (ok) {
        const unsigned key = ((unsigned)dist << 20) | (unsigned)c;
        if (key < best) { second = best; best = key; } else if (key < second) second = key;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned ob = __shfl_xor_sync(0xffffffffu, best, o);
      const unsigned os = __shfl_xor_sync(0xffffffffu, second, o);
      orbfe_merge2(best, second, ob, os);
    }
    if (best == 0xffffffffu) continue;  // warp-uniform
    if (lane == 0) {
      const int bestDist = (int)(best >> 20);
      const uint2 bc = S.cand[off + (int)(best & 0xfffffu)];
      const int bestIdx = (int)bc.x;
      bool accept = false;
      if (A.mode == ORBFE_MODE_INIT) {
        // bestDist <= TH_LOW && bestDist < (float)bestDist2 * mfNNratio (:320-322); no second => INT_MAX
        const float d2 = second == 0xffffffffu ? (float)0x7fffffff : (float)(int)(second >> 20);
        accept = bestDist <= 50 && (float)bestDist < __fmul_rn(d2, A.nnratio);
        if (accept) {
          const int prev = s_state[nKp + bestIdx];  // vnMatches21
          if (prev >= 0) { A.out[prev] = -1; --nmatches; }
          A.out[q] = bestIdx;
          s_state[nKp + bestIdx] = q;
          s_state[bestIdx] = bestDist;
          ++nmatches;
        }
      } else if (A.mode == ORBFE_MODE_MAPPOINTS) {
        if (bestDist <= 100) {  // TH_HIGH
          accept = true;
          if (second != 0xffffffffu) {
            const uint2 sc = S.cand[off + (int)(second & 0xfffffu)];
            const int bestDist2 = (int)(second >> 20);
            const int bestLevel = (int)(bc.y >> 16), bestLevel2 = (int)(sc.y >> 16);
            if (bestLevel == bestLevel2 && (float)bestDist > __fmul_rn(A.nnratio, (float)bestDist2)) accept = false;
          }
          // no second: bestLevel2 == -1 != bestLevel => accepted (:92-95)
          if (accept) { A.out[bestIdx] = q; s_state[bestIdx] = A.hasObs[q]; ++nmatches; }
        }
      } else {
        if (bestDist <= 100) {
          accept = true;
          A.out[bestIdx] = q; s_state[bestIdx] = A.hasObs[q]; ++nmatches;
        }
      }
      if (accept && A.checkOri && A.mode != ORBFE_MODE_MAPPOINTS) {
        float rot = __fsub_rn(A.qAngle[q], A.kp[bestIdx].angle);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        int bin = (int)roundf(__fmul_rn(rot, factor));
        if (bin == ORBFE_HISTO_LENGTH) bin = 0;
        s_hist[bin]++;
        A.evBin[q] = bin;
        A.evIdx[q] = A.mode == ORBFE_MODE_INIT ? q : bestIdx;
      }
    }
    __syncwarp();
  }
  nmatches = __shfl_sync(0xffffffffu, nmatches, 0);
  if (A.checkOri && A.mode != ORBFE_MODE_MAPPOINTS) {
    // ComputeThreeMaxima (orb_matcher.cpp:1584-1625), lane 0
    int ind1 = -1, ind2 = -1, ind3 = -1;
    if (lane == 0) {
      int max1 = 0, max2 = 0, max3 = 0;
      for (int i = 0; i < ORBFE_HISTO_LENGTH; i++) {
        const int s = s_hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
      }
      if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
      else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
    }
    ind1 = __shfl_sync(0xffffffffu, ind1, 0);
    ind2 = __shfl_sync(0xffffffffu, ind2, 0);
    ind3 = __shfl_sync(0xffffffffu, ind3, 0);
    __syncwarp();
    int removed = 0;
    for (int q0 = 0; q0 < A.nQ; q0 += 32) {
      const int q = q0 + lane;
      if (q < A.nQ) {
        const int bin = A.evBin[q];
        if (bin >= 0 && bin != ind1 && bin != ind2 && bin != ind3) {
          const int idx = A.evIdx[q];
          if (A.mode == ORBFE_MODE_INIT) {
            if (A.out[idx] >= 0) { A.out[idx] = -1; ++removed; }  // :364-370
          } else {
            A.out[idx] = -1; ++removed;  // CurrentFrame.SetMapPoint(idx, NULL); nmatches-- (:1441-1446)
          }
        }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
    nmatches -= removed;
  }
  if (lane == 0) A.result[0] = nmatches;
}

// ---- phase A for the vocabulary-node searches: SearchByBoW(KeyFrame, Frame) (orb_matcher.cpp:133-262),
// SearchByBoW(KeyFrame, KeyFrame) (:499-632) and SearchForTriangulation (:634-802).  The candidates of a query
// feature are the searched frame's features of the same vocabulary node, in the node's own order (DBoW2
// FeatureVector), so the list offsets are known on the host.  Candidates rejected by a static filter keep
// their slot with distance 0xffff (the resolve kernels ignore distances >= 256).  One warp per query.
struct BowFilter {
  const uint8_t* valid2;    // per searched keypoint: usable as a candidate (null = all); :551-557, :701-705
  int tri;                  // 1 = SearchForTriangulation: stereo / epipole / epipolar-line gates
  int onlyStereo;           // bOnlyStereo (:686-688, :709-711)
  const float* qx;          // per query: kp1.pt.x, kp1.pt.y (undistorted)
  const float* qy;
  const uint8_t* qStereo;   // per query: pKF1->right_coords[idx1] >= 0
  float F12[9];             // row-major fundamental matrix
  float ex, ey;             // epipole in the second image (:643-649)
};

struct BowQueries {          // vocabulary-node searches: where a query's descriptor and candidate list come from
  const uint8_t* qDescAll;   // all side-1 descriptors
  const int* qDescIdx;       // per query: row of qDescAll
  const unsigned* featIdx;   // the searched frame's flattened feature-vector indices
  const int* qSrcOff;        // per query: first entry of its node's list in featIdx
  int nQ;
};

// one query, one warp (q < BQ.nQ, warp-uniform)
__device__ __forceinline__ void orbfe_candidates_bow_query(const FrameGrid& F, const BowQueries& BQ, const MatchScratch& S,
                                                           const BowFilter& B, const int q, const int lane) {
  const uint8_t* qDescAll = BQ.qDescAll;
  const int* qDescIdx = BQ.qDescIdx;
  const unsigned* featIdx = BQ.featIdx;
  const int* qSrcOff = BQ.qSrcOff;
  const int cnt = S.qCnt[q], off = S.qOff[q], src = qSrcOff[q];
  const uint8_t* qd = qDescAll + (size_t)qDescIdx[q] * 32;
  const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(qd)), a1 = __ldg(reinterpret_cast<const uint4*>(qd) + 1);
  float la = 0.f, lb = 0.f, lc = 0.f, den = 0.f;
  bool stereo1 = false;
  if (B.tri) {  // epipolar line in the second image l = x1' F12 (CheckDistEpipolarLine, orb_matcher.cpp:114-123)
    const float x1 = B.qx[q], y1 = B.qy[q];
    la = __fadd_rn(__fadd_rn(__fmul_rn(x1, B.F12[0]), __fmul_rn(y1, B.F12[3])), B.F12[6]);
    lb = __fadd_rn(__fadd_rn(__fmul_rn(x1, B.F12[1]), __fmul_rn(y1, B.F12[4])), B.F12[7]);
    lc = __fadd_rn(__fadd_rn(__fmul_rn(x1, B.F12[2]), __fmul_rn(y1, B.F12[5])), B.F12[8]);
    den = __fadd_rn(__fmul_rn(la, la), __fmul_rn(lb, lb));
    stereo1 = B.qStereo[q] != 0;
  }
  for (int c = lane; c < cnt; c += 32) {
    const int idx = (int)featIdx[src + c];
    const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(F.desc + (size_t)idx * 32));
    const uint4 b1 = __ldg(reinterpret_cast<const uint4*>(F.desc + (size_t)idx * 32) + 1);
    int dist = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
               __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
    bool ok = !(B.valid2 && !B.valid2[idx]);
    if (ok && B.tri) {
      const MatchKp k2 = F.kp[idx];
      const bool stereo2 = F.uR[idx] >= 0;
      if (B.onlyStereo && !stereo2) ok = false;
      if (ok && !stereo1 && !stereo2) {  // too close to the epipole (:722-728)
        const float dx = __fsub_rn(B.ex, k2.x), dy = __fsub_rn(B.ey, k2.y);
        if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < __fmul_rn(100.0f, F.lvl[k2.octave])) ok = false;
      }
      if (ok) {  // CheckDistEpipolarLine (:121-130)
        const float num = __fadd_rn(__fadd_rn(__fmul_rn(la, k2.x), __fmul_rn(lb, k2.y)), lc);
        if (den == 0) ok = false;
        else {
          const float dsqr = __fdiv_rn(__fmul_rn(num, num), den);
          ok = (double)dsqr < __dmul_rn(3.84, (double)F.lvl[F.nlevels + k2.octave]);
        }
      }
    }
    S.cand[off + c] = make_uint2((unsigned)idx, ok ? (unsigned)dist : 0xffffu);
  }
}

__global__ void __launch_bounds__(ORBFE_MATCH_THREADS)
k_match_candidates_bow(const FrameGrid F, const BowQueries BQ, const MatchScratch S, const BowFilter B) {
  const int q = blockIdx.x * (ORBFE_MATCH_THREADS / 32) + (threadIdx.x >> 5);
  if (q >= BQ.nQ) return;
  orbfe_candidates_bow_query(F, BQ, S, B, q, threadIdx.x & 31);
}

// SearchBySim3 agreement check (orb_matcher.cpp:1291-1307): match12[i1] = idx2 iff vnMatch1[i1] == idx2 and
// vnMatch2[idx2] == i1
__global__ void __launch_bounds__(256)
k_sim3_agree(const int* __restrict__ vnMatch1, const int n1, const int* __restrict__ vnMatch2, const int n2,
             int* __restrict__ match12, int* __restrict__ nFound) {
  const int i1 = blockIdx.x * 256 + threadIdx.x;
  bool hit = false;
  if (i1 < n1) {
    const int idx2 = vnMatch1[i1];
    hit = idx2 >= 0 && idx2 < n2 && vnMatch2[idx2] == i1;
    match12[i1] = hit ? idx2 : -1;
  }
  const unsigned bal = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && bal) atomicAdd(nFound, __popc(bal));
}

// ---- phase B, parallel form for the SearchByProjection routines (modes 1, 2) --------------------------
// The only coupling between queries is the occupancy feedback: query q skips a keypoint that holds a map
// point with observations, i.e. that was occupied on entry or was taken by an EARLIER accepted query
// q' < q with hasObs(q') (orb_matcher.cpp:59-63 + :97, :1398-1402 + :1424).  Define
//     firstOwner[kp] = min { q' : accepted(q'), hasObs(q'), best(q') = kp };
// then occupied_before_q(kp) = occupiedIn[kp] || firstOwner[kp] < q, and the serial result is the unique
// fixed point of  best = F(firstOwner(best)).  Jacobi iteration from "no owners" reaches it: after
// iteration t the first t queries are final (induction on q), and in practice chains are 2-4 long.
// One kernel launch per iteration, one warp per query; iterations after convergence return at once.
struct JacobiState {
  int* best;      // nQ: keypoint taken by the query in the previous iteration (-1 none, -2 not evaluated yet)
  int* own;       // 3 x nKp rotating firstOwner buffers (0x7f7f7f7f = none)
  int* changed;   // per-iteration "some query changed its answer" flags
};

// iteration t of query q, one warp (q < A.nQ, warp-uniform); raises *changed when the query's answer differs from iteration t-1
__device__ __forceinline__ void orbfe_iterate_query(const ResolveArgs& A, const MatchScratch& S, const JacobiState& J, const int t,
                                                    const int q, const int lane, int* changed) {
  const int* ownPrev = J.own + (size_t)(t % 3) * A.nKp;
  int* ownNext = J.own + (size_t)((t + 1) % 3) * A.nKp;
  const int cnt = S.qCnt[q], off = S.qOff[q];
  unsigned best = 0xffffffffu, second = 0xffffffffu;
  for (int c = lane; c < cnt; c += 32) {
    const uint2 cd = S.cand[off + c];
    const int dist = (int)(cd.y & 0xffffu);
    if (dist < 256 && !(A.occupiedIn && A.occupiedIn[cd.x]) && !(__ldcg(ownPrev + cd.x) < q)) {  // L2: other SMs wrote it
      const unsigned key = ((unsigned)dist << 20) | (unsigned)(A.tieLast ? 0xfffff - c : c);
      if (key < best) { second = best; best = key; } else if (key < second) second = key;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned ob = __shfl_xor_sync(0xffffffffu, best, o);
    const unsigned os = __shfl_xor_sync(0xffffffffu, second, o);
    orbfe_merge2(best, second, ob, os);
  }
  if (lane != 0) return;
  int newBest = -1;
  if (best != 0xffffffffu) {
    const int bestDist = (int)(best >> 20);
    const int bpos = (int)(best & 0xfffffu), spos = (int)(second & 0xfffffu);
    const uint2 bc = S.cand[off + (A.tieLast ? 0xfffff - bpos : bpos)];
    bool accept = bestDist <= A.thAccept;
    if (A.ratio == ORBFE_RATIO_BOW) {  // SearchByBoW: bestDist1 < mfNNratio * bestDist2 (:187-189, :575-577)
      const float d2 = second == 0xffffffffu ? 256.0f : (float)(int)(second >> 20);
      accept = accept && (float)bestDist < __fmul_rn(A.nnratio, d2);
    }
    if (accept && A.ratio == ORBFE_RATIO_SAMELEVEL && second != 0xffffffffu) {
      const uint2 sc = S.cand[off + (A.tieLast ? 0xfffff - spos : spos)];
      if ((bc.y >> 16) == (sc.y >> 16) && (float)bestDist > __fmul_rn(A.nnratio, (float)(int)(second >> 20))) accept = false;
    }
    if (accept) newBest = (int)bc.x;
  }
  if (t == 0 || newBest != J.best[q]) { J.best[q] = newBest; *changed = 1; }
  if (newBest >= 0 && (A.feedback == ORBFE_FEEDBACK_ALL || (A.feedback == ORBFE_FEEDBACK_HASOBS && A.hasObs[q])))
    atomicMin(&ownNext[newBest], q);
}

__global__ void __launch_bounds__(ORBFE_MATCH_THREADS)
k_match_iterate(const ResolveArgs A, const MatchScratch S, const JacobiState J, const int t) {
  if (S.cursor[1]) return;                       // candidate buffer overflowed: the host re-runs
  if (t > 0 && J.changed[t - 1] == 0) return;    // converged
  const int gtid = blockIdx.x * ORBFE_MATCH_THREADS + threadIdx.x, gsz = gridDim.x * ORBFE_MATCH_THREADS;
  int* ownClear = J.own + (size_t)((t + 2) % 3) * A.nKp;   // becomes the write target of iteration t+1
  for (int i = gtid; i < A.nKp; i += gsz) ownClear[i] = 0x7f7f7f7f;
  const int q = blockIdx.x * (ORBFE_MATCH_THREADS / 32) + (threadIdx.x >> 5);
  if (q >= A.nQ) return;
  orbfe_iterate_query(A, S, J, t, q, threadIdx.x & 31, J.changed + t);
}

// after convergence: F.SetMapPoint results (the LAST accepted query on a keypoint wins), the match count and
// the rotation-consistency check (one CTA)
// one CTA (any size)
__device__ __forceinline__ void orbfe_match_finalize_block(const ResolveArgs& A, const JacobiState& J) {
  __shared__ int s_hist[ORBFE_HISTO_LENGTH];
  __shared__ int s_ind[3];
  __shared__ int s_n;
  const int tid = threadIdx.x, T = blockDim.x;
  if (!A.perQuery)
    for (int i = tid; i < A.nKp; i += T) A.out[i] = -1;
  if (tid < ORBFE_HISTO_LENGTH) s_hist[tid] = 0;
  if (tid == 0) s_n = 0;
  __syncthreads();
  const bool ori = A.checkOri != 0;
  const float factor = 1.0f / ORBFE_HISTO_LENGTH;  // orb_matcher.cpp:1322 (the reference's bin-width bug, kept)
  int mine = 0;
  for (int q = tid; q < A.nQ; q += T) {
    const int b = J.best[q];
    int bin = -1;
    if (A.perQuery) A.out[q] = b;
    if (b >= 0) {
      ++mine;
      if (!A.perQuery) atomicMax(&A.out[b], q);
      if (ori) {
        float rot = __fsub_rn(A.qAngle[q], A.kp[b].angle);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        bin = (int)roundf(__fmul_rn(rot, factor));
        if (bin == ORBFE_HISTO_LENGTH) bin = 0;
        atomicAdd(&s_hist[bin], 1);
      }
    }
    A.evBin[q] = bin;
  }
  if (mine) atomicAdd(&s_n, mine);
  __syncthreads();
  if (ori) {
    if (tid == 0) {  // ComputeThreeMaxima (orb_matcher.cpp:1584-1625)
      int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
      for (int i = 0; i < ORBFE_HISTO_LENGTH; i++) {
        const int s = s_hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
      }
      if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
      else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
      s_ind[0] = ind1; s_ind[1] = ind2; s_ind[2] = ind3;
    }
    __syncthreads();
    int removed = 0;
    for (int q = tid; q < A.nQ; q += T) {
      const int bin = A.evBin[q];
      if (bin >= 0 && bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) {
        A.out[A.perQuery ? q : J.best[q]] = -1;  // CurrentFrame.SetMapPoint(idx, NULL); nmatches-- (:1441-1446); vMatches12[idx1] = -1 (:784)
        ++removed;
      }
    }
    // every atomicMax above is complete (barrier) before any thread writes -1
    if (removed) atomicSub(&s_n, removed);
    __syncthreads();
  }
  if (tid == 0) A.result[0] = s_n;
}

__global__ void __launch_bounds__(1024)
k_match_finalize(const ResolveArgs A, const MatchScratch S, const JacobiState J) {
  if (S.cursor[1]) return;
  orbfe_match_finalize_block(A, J);
}

// ---- SearchForInitialization, parallel form (orb_matcher.cpp:264-382) ----------------------------------------------------
// The coupling between queries is vMatchedDistance: query q skips a candidate c when an EARLIER accepted query q' < q took c
// at a distance <= dist(q, c) (:306-307; every accept on c lowers vMatchedDistance[c], so the value q sees is the minimum over
// the earlier acceptors).  As for the projection searches, the serial result is the unique fixed point of
//     acc = F(acceptors(acc) restricted to q' < q)
// and Jacobi iteration from "no acceptors" reaches it (query q is final after q+1 iterations; chains are a few links long in
// practice).  Each keypoint keeps the (query, distance) pairs of its acceptors of the previous iteration in ORBFE_INIT_SLOTS
// slots; a keypoint with more acceptors than that raises `overflow` and the host falls back to the one-warp serial kernel.
#define ORBFE_INIT_SLOTS 8
struct InitJacobi {
  int* acc;        // nQ: accepted target | distance << 22 of the previous iteration (-1 none, -2 not evaluated yet)
  int* cnt;        // 3 x nKp rotating acceptor counts
  uint2* slots;    // 3 x nKp x ORBFE_INIT_SLOTS rotating (query, distance)
  int* changed;    // per-iteration "some query changed its answer" flags
  int* overflow;
};

__device__ __forceinline__ void orbfe_init_iterate_query(const ResolveArgs& A, const MatchScratch& S, const InitJacobi& J, const int t,
                                                         const int q, const int lane, int* changed) {
  const int* cntPrev = J.cnt + (size_t)(t % 3) * A.nKp;
  const uint2* slotPrev = J.slots + (size_t)(t % 3) * A.nKp * ORBFE_INIT_SLOTS;
  int* cntNext = J.cnt + (size_t)((t + 1) % 3) * A.nKp;
  uint2* slotNext = J.slots + (size_t)((t + 1) % 3) * A.nKp * ORBFE_INIT_SLOTS;
  const int cnt = S.qCnt[q], off = S.qOff[q];
  unsigned best = 0xffffffffu, second = 0xffffffffu;
  for (int c = lane; c < cnt; c += 32) {
    const uint2 cd = S.cand[off + c];
    const int dist = (int)(cd.y & 0xffffu);
    int matched = 0x7fffffff;  // vMatchedDistance[i2] as query q sees it
    const int na = min(__ldcg(cntPrev + cd.x), ORBFE_INIT_SLOTS);  // L2: other SMs wrote it
    for (int k = 0; k < na; ++k) {
      const uint2 a = __ldcg(slotPrev + (size_t)cd.x * ORBFE_INIT_SLOTS + k);
      if ((int)a.x < q) matched = min(matched, (int)a.y);
    }
    if (!(matched <= dist)) {
      const unsigned key = ((unsigned)dist << 20) | (unsigned)c;
      if (key < best) { second = best; best = key; } else if (key < second) second = key;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned ob = __shfl_xor_sync(0xffffffffu, best, o);
    const unsigned os = __shfl_xor_sync(0xffffffffu, second, o);
    orbfe_merge2(best, second, ob, os);
  }
  if (lane != 0) return;
  int acc = -1;
  if (best != 0xffffffffu) {
    const int bestDist = (int)(best >> 20);
    const float d2 = second == 0xffffffffu ? (float)0x7fffffff : (float)(int)(second >> 20);
    if (bestDist <= 50 && (float)bestDist < __fmul_rn(d2, A.nnratio)) {  // :320-322
      const int idx = (int)S.cand[off + (int)(best & 0xfffffu)].x;
      acc = idx | (bestDist << 22);
      const int slot = atomicAdd(&cntNext[idx], 1);
      if (slot < ORBFE_INIT_SLOTS) slotNext[(size_t)idx * ORBFE_INIT_SLOTS + slot] = make_uint2((unsigned)q, (unsigned)bestDist);
      else *J.overflow = 1;
    }
  }
  if (t == 0 || acc != J.acc[q]) { J.acc[q] = acc; *changed = 1; }
}

__global__ void __launch_bounds__(ORBFE_MATCH_THREADS)
k_init_iterate(const ResolveArgs A, const MatchScratch S, const InitJacobi J, const int t) {
  if (S.cursor[1]) return;
  if (t > 0 && J.changed[t - 1] == 0) return;
  const int gtid = blockIdx.x * ORBFE_MATCH_THREADS + threadIdx.x, gsz = gridDim.x * ORBFE_MATCH_THREADS;
  int* cntClear = J.cnt + (size_t)((t + 2) % 3) * A.nKp;
  for (int i = gtid; i < A.nKp; i += gsz) cntClear[i] = 0;
  const int q = blockIdx.x * (ORBFE_MATCH_THREADS / 32) + (threadIdx.x >> 5);
  if (q >= A.nQ) return;
  orbfe_init_iterate_query(A, S, J, t, q, threadIdx.x & 31, J.changed + t);
}

// after convergence: vnMatches12 (the LAST accepted query on a keypoint owns it, :325-333), nmatches, rotation check (:339-374;
// the histogram counts every accept, stolen ones included, exactly as rotHist does)
__device__ __forceinline__ void orbfe_init_finalize_block(const ResolveArgs& A, const InitJacobi& J, int* __restrict__ owner) {
  __shared__ int s_hist[ORBFE_HISTO_LENGTH];
  __shared__ int s_ind[3];
  __shared__ int s_n;
  const int tid = threadIdx.x, T = blockDim.x;
  for (int i = tid; i < A.nKp; i += T) owner[i] = -1;
  for (int q = tid; q < A.nQ; q += T) A.out[q] = -1;
  if (tid < ORBFE_HISTO_LENGTH) s_hist[tid] = 0;
  if (tid == 0) s_n = 0;
  __syncthreads();
  for (int q = tid; q < A.nQ; q += T) {
    const int acc = J.acc[q];
    if (acc >= 0) atomicMax(&owner[acc & 0x3fffff], q);
  }
  __syncthreads();
  const float factor = 1.0f / ORBFE_HISTO_LENGTH;  // orb_matcher.cpp:275 (the reference's bin-width bug, kept)
  int mine = 0;
  for (int q = tid; q < A.nQ; q += T) {
    const int acc = J.acc[q];
    int bin = -1;
    if (acc >= 0) {
      const int idx = acc & 0x3fffff;
      if (owner[idx] == q) { A.out[q] = idx; ++mine; }
      if (A.checkOri) {
        float rot = __fsub_rn(A.qAngle[q], A.kp[idx].angle);
        if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
        bin = (int)roundf(__fmul_rn(rot, factor));
        if (bin == ORBFE_HISTO_LENGTH) bin = 0;
        atomicAdd(&s_hist[bin], 1);
      }
    }
    A.evBin[q] = bin;
  }
  if (mine) atomicAdd(&s_n, mine);
  __syncthreads();
  if (A.checkOri) {
    if (tid == 0) {  // ComputeThreeMaxima (orb_matcher.cpp:1584-1625)
      int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
      for (int i = 0; i < ORBFE_HISTO_LENGTH; i++) {
        const int s = s_hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
      }
      if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
      else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
      s_ind[0] = ind1; s_ind[1] = ind2; s_ind[2] = ind3;
    }
    __syncthreads();
    int removed = 0;
    for (int q = tid; q < A.nQ; q += T) {
      const int bin = A.evBin[q];
      if (bin >= 0 && bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2] && A.out[q] >= 0) { A.out[q] = -1; ++removed; }  // :364-370
    }
    if (removed) atomicSub(&s_n, removed);
    __syncthreads();
  }
  if (tid == 0) A.result[0] = s_n;
}

__global__ void __launch_bounds__(1024)
k_init_finalize(const ResolveArgs A, const MatchScratch S, const InitJacobi J, int* __restrict__ owner) {
  if (S.cursor[1] || *J.overflow) return;
  orbfe_init_finalize_block(A, J, owner);
}

#ifndef ORBFE_EMU
// ---- the whole search in ONE cooperative launch ------------------------------------------------------------------------
// candidates -> grid barrier -> Jacobi iterations to the fixed point (one grid barrier each) -> finalize (CTA 0).  The grid
// is persistent (at most what is co-resident); warps stride over the queries.  No host round trip between the phases: the
// host enqueues one H2D copy of the packed queries, this kernel, and the D2H copy of the result.
struct SolveCfg {
  int bow;        // candidates from vocabulary nodes (BowQueries) instead of grid windows
  int init;       // SearchForInitialization coupling (InitJacobi) instead of the occupancy feedback (JacobiState)
  int* flags;     // 3 rotating "some query changed" flags
  int* iowner;    // init: per-keypoint owner scratch
  int* hist;      // ORBFE_HISTO_LENGTH rotation-histogram bins, zeroed by the host with the cursor block
};
// A/B on B200 (tools/probes/solve_probe.py; 20 000 queries against 8000 keypoints / 2000 against 2000): 512 threads, 64 registers
// 0.344 / 0.090 ms; 512 x 3 CTAs (40 registers, spills) 0.482 / 0.091; 256 x 4 0.375 / 0.090; 256 x 6 0.510 / 0.090; 128 x 12 0.549 / 0.090
#ifndef ORBFE_SOLVE_THREADS
#define ORBFE_SOLVE_THREADS 512
#endif
#ifndef ORBFE_SOLVE_MINB
#define ORBFE_SOLVE_MINB 1
#endif

__global__ void __launch_bounds__(ORBFE_SOLVE_THREADS, ORBFE_SOLVE_MINB)
k_match_solve(const FrameGrid F, const MatchQueries Q, const BowQueries BQ, const BowFilter B, const MatchScratch S, const ResolveArgs A,
              const JacobiState J, const InitJacobi IJ, const SolveCfg C) {
  cooperative_groups::grid_group grid = cooperative_groups::this_grid();
  const int lane = threadIdx.x & 31;
  const int gtid = blockIdx.x * ORBFE_SOLVE_THREADS + threadIdx.x, gsz = gridDim.x * ORBFE_SOLVE_THREADS;
  const int gwarp = gtid >> 5, nwarps = gsz >> 5;
  const int nQ = A.nQ;
  // state of iteration 0: "no owners / no acceptors" in the buffers iteration 0 reads (0) and writes (1)
  if (C.init) {
    for (int i = gtid; i < 2 * A.nKp; i += gsz) IJ.cnt[i] = 0;
  } else {
    for (int i = gtid; i < 2 * A.nKp; i += gsz) J.own[i] = 0x7f7f7f7f;
  }
  if (gtid < 3) C.flags[gtid] = 0;
  if (gtid == 0 && C.init) *IJ.overflow = 0;
  // the result arrays start "unmatched" (the finalize phase below only writes accepts)
  if (C.init) {
    for (int i = gtid; i < A.nKp; i += gsz) C.iowner[i] = -1;
    for (int q = gtid; q < nQ; q += gsz) A.out[q] = -1;
  } else if (!A.perQuery) {
    for (int i = gtid; i < A.nKp; i += gsz) A.out[i] = -1;
  }
  if (C.bow) {
    for (int q = gwarp; q < nQ; q += nwarps) orbfe_candidates_bow_query(F, BQ, S, B, q, lane);
  } else {
    for (int q = gwarp; q < nQ; q += nwarps) orbfe_candidates_query(F, Q, S, q, lane);
  }
  grid.sync();
  if (*reinterpret_cast<volatile int*>(S.cursor + 1)) return;   // candidate buffer too small (grid-uniform): the host grows it and re-runs
  const bool coupled = C.init || A.feedback != ORBFE_FEEDBACK_NONE;
  for (int t = 0; t <= nQ; ++t) {
    int* changed = C.flags + t % 3;
    if (gtid == 0) C.flags[(t + 1) % 3] = 0;
    if (C.init) {
      int* cntClear = IJ.cnt + (size_t)((t + 2) % 3) * A.nKp;
      for (int i = gtid; i < A.nKp; i += gsz) cntClear[i] = 0;
      for (int q = gwarp; q < nQ; q += nwarps) orbfe_init_iterate_query(A, S, IJ, t, q, lane, changed);
    } else {
      int* ownClear = J.own + (size_t)((t + 2) % 3) * A.nKp;
      for (int i = gtid; i < A.nKp; i += gsz) ownClear[i] = 0x7f7f7f7f;
      for (int q = gwarp; q < nQ; q += nwarps) orbfe_iterate_query(A, S, J, t, q, lane, changed);
    }
    if (!coupled) break;     // independent queries: one evaluation is the answer
    grid.sync();
    if (*reinterpret_cast<volatile int*>(changed) == 0) break;   // fixed point (grid-uniform)
  }
  if (!coupled) grid.sync();
  // ---- finalize, spread over the grid (the one-CTA forms above are what the emulated build launches) -------------------
  __shared__ int s_hist[ORBFE_HISTO_LENGTH];
  __shared__ int s_ind[3];
  const bool ori = A.checkOri != 0;
  const float factor = 1.0f / ORBFE_HISTO_LENGTH;  // orb_matcher.cpp:275 / :1322 (the reference's bin-width bug, kept)
  if (threadIdx.x < ORBFE_HISTO_LENGTH) s_hist[threadIdx.x] = 0;
  __syncthreads();
  int mine = 0;
  if (C.init) {
    if (*reinterpret_cast<volatile int*>(IJ.overflow)) return;   // grid-uniform: the host falls back to the serial resolve
    // the LAST accepted query on a keypoint owns it (:325-333)
    for (int q = gtid; q < nQ; q += gsz) {
      const int acc = IJ.acc[q];
      if (acc >= 0) atomicMax(&C.iowner[acc & 0x3fffff], q);
    }
    grid.sync();
    for (int q = gtid; q < nQ; q += gsz) {
      const int acc = IJ.acc[q];
      int bin = -1;
      if (acc >= 0) {
        const int idx = acc & 0x3fffff;
        if (__ldcg(C.iowner + idx) == q) { A.out[q] = idx; ++mine; }
        if (ori) {  // the histogram counts every accept, stolen ones included, exactly as rotHist does (:339-374)
          float rot = __fsub_rn(A.qAngle[q], A.kp[idx].angle);
          if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
          bin = (int)roundf(__fmul_rn(rot, factor));
          if (bin == ORBFE_HISTO_LENGTH) bin = 0;
          atomicAdd(&s_hist[bin], 1);
        }
      }
      A.evBin[q] = bin;
    }
  } else {
    for (int q = gtid; q < nQ; q += gsz) {
      const int b = J.best[q];
      int bin = -1;
      if (A.perQuery) A.out[q] = b;
      if (b >= 0) {
        ++mine;
        if (!A.perQuery) atomicMax(&A.out[b], q);   // F.SetMapPoint: the last accepted query on a keypoint wins
        if (ori) {
          float rot = __fsub_rn(A.qAngle[q], A.kp[b].angle);
          if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
          bin = (int)roundf(__fmul_rn(rot, factor));
          if (bin == ORBFE_HISTO_LENGTH) bin = 0;
          atomicAdd(&s_hist[bin], 1);
        }
      }
      A.evBin[q] = bin;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if (lane == 0 && mine) atomicAdd(A.result, mine);
  if (!ori) return;
  __syncthreads();
  if (threadIdx.x < ORBFE_HISTO_LENGTH && s_hist[threadIdx.x]) atomicAdd(C.hist + threadIdx.x, s_hist[threadIdx.x]);
  grid.sync();   // every accept is in out[] / hist[] before anybody removes one
  if (threadIdx.x == 0) {  // ComputeThreeMaxima (orb_matcher.cpp:1584-1625), once per CTA
    int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
    for (int i = 0; i < ORBFE_HISTO_LENGTH; i++) {
      const int h = __ldcg(C.hist + i);
      if (h > max1) { max3 = max2; max2 = max1; max1 = h; ind3 = ind2; ind2 = ind1; ind1 = i; }
      else if (h > max2) { max3 = max2; max2 = h; ind3 = ind2; ind2 = i; }
      else if (h > max3) { max3 = h; ind3 = i; }
    }
    if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
    else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
    s_ind[0] = ind1; s_ind[1] = ind2; s_ind[2] = ind3;
  }
  __syncthreads();
  int removed = 0;
  for (int q = gtid; q < nQ; q += gsz) {
    const int bin = A.evBin[q];
    if (bin >= 0 && bin != s_ind[0] && bin != s_ind[1] && bin != s_ind[2]) {
      if (C.init) {
        if (A.out[q] >= 0) { A.out[q] = -1; ++removed; }  // :364-370
      } else {
        A.out[A.perQuery ? q : J.best[q]] = -1;  // CurrentFrame.SetMapPoint(idx, NULL); nmatches-- (:1441-1446); vMatches12[idx1] = -1 (:784)
        ++removed;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
  if (lane == 0 && removed) atomicSub(A.result, removed);
}
#endif

// ---- OrbMatcher::DescriptorDistance, batched (orb_matcher.cpp:1630-1646) --------------------------
__global__ void __launch_bounds__(256)
k_descriptor_distance(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, const int n, int* __restrict__ d) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(a + (size_t)i * 32)), a1 = __ldg(reinterpret_cast<const uint4*>(a + (size_t)i * 32) + 1);
  const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(b + (size_t)i * 32)), b1 = __ldg(reinterpret_cast<const uint4*>(b + (size_t)i * 32) + 1);
  d[i] = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// ---- Frame::GetFeaturesInArea as a call of its own (one warp) --------------------------------------
__global__ void __launch_bounds__(32)
k_features_in_area(const FrameGrid F, const float x, const float y, const float r, const int minL, const int maxL,
                   int* __restrict__ out, const int capacity, int* __restrict__ nOut) {
  const int lane = threadIdx.x;
  int x0, x1, y0, y1, count = 0;
  if (orbfe_window_cells(F, x, y, r, x0, x1, y0, y1)) {
    const bool checkLevels = (minL > 0) || (maxL >= 0);
    for (int ix = x0; ix <= x1; ++ix) {
      const int b = F.cellStart[ix * ORBFE_GRID_ROWS + y0], e = F.cellStart[ix * ORBFE_GRID_ROWS + y1 + 1];
      for (int j0 = b; j0 < e; j0 += 32) {
        const int j = j0 + lane;
        bool ok = j < e;
        int idx = 0;
        if (ok) {
          idx = F.cellItems[j];
          const MatchKp k = F.kp[idx];
          if (checkLevels && (k.octave < minL || (maxL >= 0 && k.octave > maxL))) ok = false;
          if (!(fabsf(__fsub_rn(k.x, x)) < r && fabsf(__fsub_rn(k.y, y)) < r)) ok = false;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        const int pos = count + __popc(bal & ((1u << lane) - 1u));
        if (ok && pos < capacity) out[pos] = idx;
        count += __popc(bal);
      }
    }
  }
  if (lane == 0) *nOut = count;
}
