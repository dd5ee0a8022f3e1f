// k_stereo.cuh -- S1-S4: Frame::ComputeStereoMatches (src/data/frame.cpp:406-577).
//   k_stereo_rows (CTA per pair): S1, the reference's row table (:415-433).
//   k_stereo_search (warp per left keypoint):
//     S2     lanes stride over vRowIndices[(int)vL], apply
//            the octave (+-1) and uR in [uL-maxD, uL] filters, compute the 256-bit Hamming
//            distance with __popc, and a warp-shuffle argmin on (dist, iR) reproduces the strict
//            '<' scan from bestDist = TH_HIGH (ties -> lowest iR) (:444-488).
//     S3     if bestDist < 75: 11 SADs of 11x11 centre-subtracted patches on the UNBLURRED
//            pyramid level of the left keypoint's octave, lanes own 4 of the 121 pixels, sums
//            reduced by shuffles; parabola fit in non-contracted fp32 (:491-562).
//   k_stereo_median (CTA per pair):
//     S4     median of the accepted SADs = element size/2 of the sorted (SAD, iL) list, found by
//            a two-pass radix select; matches with SAD >= 1.5f*1.4f*median are undone (:565-576).
#pragma once
#include "orbfe_common.cuh"
#include "k_describe.cuh"

#ifndef ORBFE_ST_THREADS
#define ORBFE_ST_THREADS 256
#endif

struct StereoPair {
  const uint8_t* pyrL;          // slot base of the left / right pyramid block
  const uint8_t* pyrR;
  const orbfe_kp_dev* kpL;
  const orbfe_kp_dev* kpR;
  const uint8_t* descL;
  const uint8_t* descR;
  const int* nL;
  const int* nR;
  float* uR;                    // outputs, indexed by left keypoint
  float* depth;
  int* sad;
  int* rowStart;                // S1 row table of the right image: nRows+1 offsets ...
  uint2* rowItems;              // ... into (right-keypoint index | octave << 24, x as float bits): the search filters on
                                // octave and x before it touches the descriptor, so they travel with the index
  int rowCap;
};

// S1 (frame.cpp:415-433): right keypoint iR is pushed into every row yi in [floor(y-r), ceil(y+r)],
// r = 2*scale[octave].  One CTA per pair builds the table (count, scan, fill); the order inside a row
// does not matter because the search reduces on (distance, iR).
__global__ void __launch_bounds__(1024)
k_stereo_rows(const __grid_constant__ Geom g, const StereoPair* __restrict__ pairs, const int maxKp) {
  ORBFE_DYN_SMEM(smem);
  int* s_cnt = reinterpret_cast<int*>(smem);  // nRows + 1
  __shared__ int s_scan[33];
  const StereoPair P = pairs[blockIdx.x];
  const int nRows = g.lv[0].h;
  const int nR = min(*P.nR, maxKp);
  const int tid = threadIdx.x, T = blockDim.x;
  for (int r = tid; r <= nRows; r += T) s_cnt[r] = 0;
  __syncthreads();
  for (int iR = tid; iR < nR; iR += T) {
    const orbfe_kp_dev kr = P.kpR[iR];
    const float r = __fmul_rn(2.0f, g.lv[kr.octave].scale);
    const int maxr = min((int)ceilf(__fadd_rn(kr.y, r)), nRows - 1);
    const int minr = max((int)floorf(__fsub_rn(kr.y, r)), 0);
    for (int yi = minr; yi <= maxr; ++yi) atomicAdd(&s_cnt[yi], 1);
  }
  __syncthreads();
  const int per = (nRows + T - 1) / T;
  const int r0 = min(tid * per, nRows), r1 = min(r0 + per, nRows);
  int sum = 0;
  for (int r = r0; r < r1; ++r) sum += s_cnt[r];
  int total;
  int run = orbfe_block_exscan(sum, s_scan, &total);
  for (int r = r0; r < r1; ++r) { const int c = s_cnt[r]; P.rowStart[r] = run; s_cnt[r] = run; run += c; }
  if (tid == 0) P.rowStart[nRows] = min(total, P.rowCap);
  __syncthreads();
  for (int iR = tid; iR < nR; iR += T) {
    const orbfe_kp_dev kr = P.kpR[iR];
    const float r = __fmul_rn(2.0f, g.lv[kr.octave].scale);
    const int maxr = min((int)ceilf(__fadd_rn(kr.y, r)), nRows - 1);
    const int minr = max((int)floorf(__fsub_rn(kr.y, r)), 0);
    for (int yi = minr; yi <= maxr; ++yi) {
      const int pos = atomicAdd(&s_cnt[yi], 1);
      if (pos < P.rowCap) P.rowItems[pos] = make_uint2((unsigned)iR | ((unsigned)kr.octave << 24), __float_as_uint(kr.x));
    }
  }
}

__device__ __forceinline__ int orbfe_hamming256(const uint4 a0, const uint4 a1, const uint8_t* __restrict__ b) {
  const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(b));
  const uint4 b1 = __ldg(reinterpret_cast<const uint4*>(b) + 1);
  return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) +
         __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

#ifndef ORBFE_ST_MINB
#define ORBFE_ST_MINB 6  // <= 42 registers: 48 warps/SM hide the L2 latency of the gathers (A/B measured)
#endif
__global__ void __launch_bounds__(ORBFE_ST_THREADS, ORBFE_ST_MINB)
k_stereo_search(const __grid_constant__ Geom g, const StereoPair* __restrict__ pairs, const float bf, const float baseline,
                const int maxKp) {
  const StereoPair P = pairs[blockIdx.y];
  const int lane = threadIdx.x & 31;
  const int iL = blockIdx.x * (ORBFE_ST_THREADS / 32) + (threadIdx.x >> 5);
  const int nL = min(*P.nL, maxKp), nR = min(*P.nR, maxKp);
  if (iL >= nL) return;
  const orbfe_kp_dev kpL = P.kpL[iL];
  float uRout = -1.0f, depthOut = -1.0f;
  int sadOut = -1;
  const int levelL = kpL.octave;
  const float vL = kpL.y, uL = kpL.x;
  const int nRows = g.lv[0].h;
  const int row = (int)vL;
  const float minD = 0.f;
  const float maxD = __fdiv_rn(bf, baseline);
  const float minU = __fsub_rn(uL, maxD), maxU = __fsub_rn(uL, minD);
  int bestDist = 100;  // OrbMatcher::TH_HIGH
  int bestIdxR = 0x7fffffff;
  unsigned bestXbits = 0u;  // x of the best right keypoint travels with the argmin: no dependent kpR load after the search
  if (row >= 0 && row < nRows && !(maxU < 0)) {
    const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(P.descL + (size_t)iL * 32));
    const uint4 a1 = __ldg(reinterpret_cast<const uint4*>(P.descL + (size_t)iL * 32) + 1);
    const int cb = P.rowStart[row], ce = min(P.rowStart[row + 1], P.rowCap);
    for (int c = cb + lane; c < ce; c += 32) {
      const uint2 it = P.rowItems[c];  // vRowIndices[row] (frame.cpp:450)
      const int iR = (int)(it.x & 0xffffffu), octR = (int)(it.x >> 24);
      const float xR = __uint_as_float(it.y);
      if (octR < levelL - 1 || octR > levelL + 1) continue;
      if (xR >= minU && xR <= maxU) {
        const int dist = orbfe_hamming256(a0, a1, P.descR + (size_t)iR * 32);
        // the reference scans the row in ascending iR with a strict '<': among equal distances the LOWEST iR wins; the row
        // table here is filled in atomic order, so the tie rule has to be applied inside the lane too, not only in the reduction
        if (dist < bestDist || (dist == bestDist && iR < bestIdxR)) { bestDist = dist; bestIdxR = iR; bestXbits = it.y; }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int od = __shfl_xor_sync(0xffffffffu, bestDist, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bestIdxR, o);
    const unsigned ox = __shfl_xor_sync(0xffffffffu, bestXbits, o);
    if (od < bestDist || (od == bestDist && oi < bestIdxR)) { bestDist = od; bestIdxR = oi; bestXbits = ox; }
  }
  const int thOrbDist = (100 + 50) / 2;
  if (bestDist < thOrbDist) {  // warp-uniform
    const float uR0 = __uint_as_float(bestXbits);  // == P.kpR[bestIdxR].x (k_stereo_rows copied it into the row table)
    const LevelGeom& L = g.lv[levelL];
    const float scaleFactor = __fdiv_rn(1.0f, L.scale);  // mvInvScaleFactor (orb_extractor.cpp:371)
    const float scaleduL = roundf(__fmul_rn(kpL.x, scaleFactor));
    const float scaledvL = roundf(__fmul_rn(kpL.y, scaleFactor));
    const float scaleduR0 = roundf(__fmul_rn(uR0, scaleFactor));
    const int w = 5;
    const float iniu = scaleduR0;             // scaleduR0 + L - w with L == w == 5 (:509)
    const float endu = scaleduR0 + 11.0f;     // scaleduR0 + L + w + 1
    const int yl0 = (int)scaledvL - w, xl0 = (int)scaleduL - w, xr0 = (int)scaleduR0 - w;
    // guards: the reference indexes the level unchecked; outside the padded plane we drop the match
    const bool inside = yl0 >= -ORBFE_EDGE && yl0 + 11 <= L.h + ORBFE_EDGE && xl0 >= -ORBFE_EDGE &&
                        xl0 + 11 <= L.w + ORBFE_EDGE && xr0 - 5 >= -ORBFE_EDGE;
    if (!(iniu < 0 || endu >= (float)L.w) && inside) {
      const uint8_t* pl = P.pyrL + L.planeOff + (size_t)ORBFE_EDGE * L.pitch + ORBFE_EDGE;
      const uint8_t* pr = P.pyrR + L.planeOff + (size_t)ORBFE_EDGE * L.pitch + ORBFE_EDGE;
      const int cL = __ldg(pl + (yl0 + w) * L.pitch + xl0 + w);
      int sums[11];
#pragma unroll
      for (int s = 0; s < 11; ++s) sums[s] = 0;
      int cR[11];
#pragma unroll
      for (int s = 0; s < 11; ++s) cR[s] = __ldg(pr + (yl0 + w) * L.pitch + xr0 + (s - 5) + w);
      for (int p = lane; p < 121; p += 32) {
        const int py = p / 11, px = p - py * 11;
        const int il = (int)__ldg(pl + (yl0 + py) * L.pitch + xl0 + px) - cL;
        const uint8_t* rrow = pr + (yl0 + py) * L.pitch + xr0 + px - 5;
#pragma unroll
        for (int s = 0; s < 11; ++s) {
          const int ir = (int)__ldg(rrow + s) - cR[s];
          const int df = il - ir;
          sums[s] += df < 0 ? -df : df;
        }
      }
#pragma unroll
      for (int s = 0; s < 11; ++s) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sums[s] += __shfl_xor_sync(0xffffffffu, sums[s], o);
      }
      int bestS = 0x7fffffff, bestinc = 0;
#pragma unroll
      for (int s = 0; s < 11; ++s)
        if (sums[s] < bestS) { bestS = sums[s]; bestinc = s - 5; }
      if (bestinc != -5 && bestinc != 5) {
        float d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
        for (int s = 1; s < 10; ++s)
          if (s - 5 == bestinc) { d1 = (float)sums[s - 1]; d2 = (float)sums[s]; d3 = (float)sums[s + 1]; }
        const float num = __fsub_rn(d1, d3);
        const float den = __fmul_rn(2.0f, __fsub_rn(__fadd_rn(d1, d3), __fmul_rn(2.0f, d2)));
        const float deltaR = __fdiv_rn(num, den);
        if (!(deltaR < -1.f || deltaR > 1.f)) {
          float bestuR = __fmul_rn(L.scale, __fadd_rn(__fadd_rn(scaleduR0, (float)bestinc), deltaR));
          float disparity = __fsub_rn(uL, bestuR);
          if (disparity >= minD && disparity < maxD) {
            if (disparity <= 0.f) { disparity = 0.01f; bestuR = __fsub_rn(uL, 0.01f); }
            depthOut = __fdiv_rn(bf, disparity);
            uRout = bestuR;
            sadOut = bestS;
          }
        }
      }
    }
  }
  if (lane == 0) { P.uR[iL] = uRout; P.depth[iL] = depthOut; P.sad[iL] = sadOut; }
}

__global__ void __launch_bounds__(ORBFE_ST_THREADS)
k_stereo_median(const StereoPair* __restrict__ pairs, const int maxKp, int* __restrict__ nMatched) {
  static_assert(ORBFE_ST_THREADS >= 256 && ORBFE_ST_THREADS % 32 == 0, "one histogram bin per thread in the rank search");
  __shared__ int s_hist[256];
  __shared__ int s_sel[4];
  __shared__ int s_scan[33];
  const StereoPair P = pairs[blockIdx.x];
  const int nL = min(*P.nL, maxKp);
  const int tid = threadIdx.x;
  // count accepted matches
  int c = 0;
  for (int i = tid; i < nL; i += ORBFE_ST_THREADS) c += P.sad[i] >= 0;
  // block sum of c via histogram slot 0
  for (int b = tid; b < 256; b += ORBFE_ST_THREADS) s_hist[b] = 0;
  __syncthreads();
  if (c) atomicAdd(&s_hist[0], c);
  __syncthreads();
  const int size = s_hist[0];
  __syncthreads();
  if (size == 0) {  // empty list: the reference's vDistIdx[0] read is UB; the cut is skipped
    if (tid == 0) nMatched[blockIdx.x] = 0;
    return;
  }
  const int rank = size / 2;  // element size/2 of the ascending sort
  // pass 1: high byte of the 16-bit SAD (SAD <= 121*510 = 61710 < 65536)
  for (int b = tid; b < 256; b += ORBFE_ST_THREADS) s_hist[b] = 0;
  __syncthreads();
  for (int i = tid; i < nL; i += ORBFE_ST_THREADS) {
    const int s = P.sad[i];
    if (s >= 0) atomicAdd(&s_hist[(s >> 8) & 0xff], 1);
  }
  __syncthreads();
  {  // the bin that holds element `rank`: exclusive prefix <= rank < inclusive prefix (one bin per thread, block scan)
    const int hcnt = tid < 256 ? s_hist[tid] : 0;
    int tot;
    const int ex = orbfe_block_exscan(hcnt, s_scan, &tot);
    if (tid < 256 && ex <= rank && rank < ex + hcnt) { s_sel[0] = tid; s_sel[1] = rank - ex; }
  }
  __syncthreads();
  const int hiBin = s_sel[0], rank2 = s_sel[1];
  __syncthreads();
  for (int b = tid; b < 256; b += ORBFE_ST_THREADS) s_hist[b] = 0;
  __syncthreads();
  for (int i = tid; i < nL; i += ORBFE_ST_THREADS) {
    const int s = P.sad[i];
    if (s >= 0 && ((s >> 8) & 0xff) == hiBin) atomicAdd(&s_hist[s & 0xff], 1);
  }
  __syncthreads();
  {
    const int hcnt = tid < 256 ? s_hist[tid] : 0;
    int tot;
    const int ex = orbfe_block_exscan(hcnt, s_scan, &tot);
    if (tid < 256 && ex <= rank2 && rank2 < ex + hcnt) s_sel[2] = (hiBin << 8) | tid;
  }
  __syncthreads();
  const float median = (float)s_sel[2];
  const float thDist = __fmul_rn(1.5f * 1.4f, median);
  int kept = 0;
  for (int i = tid; i < nL; i += ORBFE_ST_THREADS) {
    const int s = P.sad[i];
    if (s < 0) continue;
    if (!((float)s < thDist)) { P.uR[i] = -1.0f; P.depth[i] = -1.0f; }
    else ++kept;
  }
  __syncthreads();
  for (int b = tid; b < 256; b += ORBFE_ST_THREADS) s_hist[b] = 0;
  __syncthreads();
  if (kept) atomicAdd(&s_hist[0], kept);
  __syncthreads();
  if (tid == 0) nMatched[blockIdx.x] = s_hist[0];
}
