// k_describe.cuh -- E5 + E7 + E8: orientation, rotated-BRIEF descriptor and keypoint packing.
// One warp per 32 keypoint slots of an image (slots are level-major, the order of
// ORBextractor::Compute's output, orb_extractor.cpp:1016-1048).
//   E5  IC_Angle (orb_extractor.cpp:18-45): lanes own the 31 columns of the radius-15 disc and
//       walk its 31 rows on the UNBLURRED level; m10/m01 are exact int32 sums reduced by
//       warp shuffles; angle = cv::fastAtan2 restated op for op in non-contracted fp32
//       (SURVEY Appendix A.4).
//   E7  computeOrbDescriptor (orb_extractor.cpp:48-88): a = cosf, b = sinf of angle*pi/180 with
//       glibc's sincosf algorithm (double polynomial, verified bit-identical to glibc 2.39 over
//       1.6e8 arguments); lane i produces descriptor byte i from 16 rotated samples
//       of the BLURRED level, sample coordinates cvRound(x*b+y*a), cvRound(x*a-y*b) with
//       separate fmul/fadd (no FMA) and round-half-even.
//   E8  pt += (16,16) is already folded in, pt *= scale[level] for level != 0 (:1039-1045),
//       size = (float)(int)(31*scale), octave = level, class_id = -1.
#pragma once
#include "orbfe_common.cuh"
#include "orbfe_tma.cuh"

#ifndef ORBFE_DESC_THREADS
#define ORBFE_DESC_THREADS 64
#endif
// Both patches of a keypoint arrive by TMA (one cp.async.bulk.tensor each, issued by lane 0, orbfe_tma.cuh): the 31 x 31 disc of
// the UNBLURRED level for the moments (box 48 B x 31 rows) and the 37 x 37 rBRIEF window of the BLURRED level (box 64 B x 37 rows).
// A box starts at the 16-byte boundary at or below the patch's first column, so 31 + 15 <= 48 and 37 + 15 <= 64 bytes per row.
// The tensor maps are written by the host before any launch and never modified, so no tensormap-proxy fence is needed.
// Each warp owns two 128-byte aligned buffers with one mbarrier each: while keypoint k is processed from one, the box of keypoint
// k + 1 lands in the other.  (The previous form staged the window with 14 rounds of per-lane loads and index arithmetic and
// read the disc with 22 global loads per lane: 250 of the ~700 warp instructions per keypoint.)
#define ORBFE_DESC_PYR_BW 48
#define ORBFE_DESC_PYR_BH 31
#define ORBFE_DESC_BLUR_BW 64
#define ORBFE_DESC_BLUR_BH 37
#define ORBFE_DESC_BUF 2432  // bytes per buffer: 37 x 64 rounded up to a multiple of 128

// global (L1-cached) rather than __constant__: each lane reads ITS 32 bytes, and per-lane addresses in
// the constant bank are serialised by the address-divergence unit (58 % ADU busy in the ncu profile)
static __device__ __align__(16) const signed char d_orb_pattern[1024] = {
#include "orb_pattern_31.inc"
};
// circular patch half-widths (orb_extractor.cpp:393-410)
__constant__ signed char c_umax[16] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};

struct orbfe_kp_dev {
  float x, y, size, angle, response;
  int octave, class_id;
};

__device__ __forceinline__ float orbfe_fast_atan2(float y, float x) {
  const float sc = (float)(180 / 3.141592653589793238462643383279502884);
  const float p1 = 0.9997878412794807f * sc, p3 = -0.3258083974640975f * sc;
  const float p5 = 0.1555786518463281f * sc, p7 = -0.04432655554792128f * sc;
  const float ax = fabsf(x), ay = fabsf(y);
  const float eps = (float)2.2204460492503131e-16;
  float a, c, c2;
  if (ax >= ay) {
    c = __fdiv_rn(ay, __fadd_rn(ax, eps));
    c2 = __fmul_rn(c, c);
    a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
  } else {
    c = __fdiv_rn(ax, __fadd_rn(ay, eps));
    c2 = __fmul_rn(c, c);
    a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
  }
  if (x < 0) a = __fsub_rn(180.f, a);
  if (y < 0) a = __fsub_rn(360.f, a);
  return a;
}

// glibc >= 2.28 sinf/cosf (ARM optimized-routines sincosf), argument range |y| < 120 only
// (angles here are in [0, 2*pi]).  All double ops are explicit round-to-nearest, no FMA.
__device__ __forceinline__ float orbfe_sincos_poly(double x, double x2, bool neg_tab, int n) {
  const double S1 = -0x1.555545995a603p-3, S2 = 0x1.1107605230bc4p-7, S3 = -0x1.994eb3774cf24p-13;
  double C0 = 0x1p0, C1 = -0x1.ffffffd0c621cp-2, C2 = 0x1.55553e1068f19p-5, C3 = -0x1.6c087e89a359dp-10,
         C4 = 0x1.99343027bf8c3p-16;
  if (neg_tab) { C0 = -C0; C1 = -C1; C2 = -C2; C3 = -C3; C4 = -C4; }
  if ((n & 1) == 0) {
    const double x3 = __dmul_rn(x, x2);
    const double s1 = __dadd_rn(S2, __dmul_rn(x2, S3));
    const double x7 = __dmul_rn(x3, x2);
    const double s = __dadd_rn(x, __dmul_rn(x3, S1));
    return __double2float_rn(__dadd_rn(s, __dmul_rn(x7, s1)));
  } else {
    const double x4 = __dmul_rn(x2, x2);
    const double c2 = __dadd_rn(C3, __dmul_rn(x2, C4));
    const double c1 = __dadd_rn(C1, __dmul_rn(x2, C2));
    const double x6 = __dmul_rn(x4, x2);
    const double c = __dadd_rn(C0, __dmul_rn(x2, c1));
    return __double2float_rn(__dadd_rn(c, __dmul_rn(x6, c2)));
  }
}
__device__ __forceinline__ unsigned orbfe_abstop12(float x) { return (__float_as_uint(x) >> 20) & 0x7ffu; }
// is_cos = 0: sinf(y), 1: cosf(y)
__device__ __forceinline__ float orbfe_sincosf(float y, int is_cos) {
  double x = (double)y;
  if (orbfe_abstop12(y) < orbfe_abstop12(0x1.921FB6p-1f)) {
    const double s = __dmul_rn(x, x);
    if (orbfe_abstop12(y) < orbfe_abstop12(0x1p-12f)) return is_cos ? 1.0f : y;
    return orbfe_sincos_poly(x, s, false, is_cos);
  }
  const double r = __dmul_rn(x, 0x1.45F306DC9C883p+23);
  const int n = ((int)r + 0x800000) >> 24;
  x = __dsub_rn(x, __dmul_rn((double)n, 0x1.921FB54442D18p0));
  const double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
  return orbfe_sincos_poly(__dmul_rn(x, sgn), __dmul_rn(x, x), (n & 2) != 0, n ^ is_cos);
}

// One warp owns 32 consecutive keypoint slots of an image:
//   phase 1  warp-cooperative moments per keypoint (lane = patch column), lane k keeps (m01,m10)
//            of keypoint k;
//   phase 2  lane-parallel: fastAtan2 + glibc sincosf (the fp64 polynomial runs once per
//            keypoint instead of 32x redundantly) + the cv::KeyPoint record;
//   phase 3  per keypoint, lane i builds descriptor byte i; the lane's 16 pattern point pairs are converted to float
//            once per CTA and kept in shared memory as 8 float4 per lane (per-lane constant-bank reads would be
//            serialised by the address-divergence unit; 32 registers per lane cost occupancy and load slots).
#ifndef ORBFE_DESC_PATSMEM
#define ORBFE_DESC_PATSMEM 1  // pattern points as floats in shared memory instead of 32 registers per lane
#endif
#ifndef ORBFE_DESC_MINB
// A/B on B200, 64 pairs, staged-by-loads form (pattern in registers -> in smem): MINB 12 0.269 -> 0.249 ms; smem + MINB 14
// (72 registers, 28 warps/SM) 0.247; MINB 16 (64 registers, spills) 0.265; MINB 20 0.324; 128-thread CTAs 0.259.
// TMA form: MINB 12 0.197, MINB 14 0.191, MINB 16 0.191, 128-thread CTAs (MINB 7) 0.191
#define ORBFE_DESC_MINB 14
#endif
__global__ void __launch_bounds__(ORBFE_DESC_THREADS, ORBFE_DESC_MINB)
k_orient_describe(const __grid_constant__ Geom g, const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur,
                  const unsigned* __restrict__ lvlKp, const int* __restrict__ lvlCnt, orbfe_kp_dev* __restrict__ kps,
                  uint8_t* __restrict__ desc, int* __restrict__ nKp, const int kpw, const uint2* __restrict__ icw,
                  const CUtensorMap* __restrict__ tmPyr, const CUtensorMap* __restrict__ tmBlur) {
  const int slot = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int wglobal = blockIdx.x * (ORBFE_DESC_THREADS / 32) + (threadIdx.x >> 5);
#if ORBFE_DESC_PATSMEM
  // the pattern as floats in shared memory, one float4 (x0, y0, x1, y1) per comparison and lane: 32 registers less per thread
  __shared__ float4 s_pat[8][32];
  if (threadIdx.x < 32) {
    const uint4 p0 = __ldg(reinterpret_cast<const uint4*>(d_orb_pattern) + 2 * lane);
    const uint4 p1 = __ldg(reinterpret_cast<const uint4*>(d_orb_pattern) + 2 * lane + 1);
    const unsigned w[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
    for (int t = 0; t < 8; ++t)
      s_pat[t][lane] = make_float4((float)(signed char)(w[t] & 0xffu), (float)(signed char)((w[t] >> 8) & 0xffu),
                                   (float)(signed char)((w[t] >> 16) & 0xffu), (float)(signed char)((w[t] >> 24) & 0xffu));
  }
  __syncthreads();
#endif
  __shared__ __align__(128) unsigned char s_buf[ORBFE_DESC_THREADS / 32][2][ORBFE_DESC_BUF];
  __shared__ __align__(8) unsigned long long s_bars[ORBFE_DESC_THREADS / 32][2];
  unsigned char (*wbuf)[ORBFE_DESC_BUF] = s_buf[threadIdx.x >> 5];
  unsigned long long* wbar = s_bars[threadIdx.x >> 5];
  if (lane == 0) { orbfe_tile_barrier_init(&wbar[0]); orbfe_tile_barrier_init(&wbar[1]); }
  __syncwarp();  // the buffers and barriers belong to this warp alone
  unsigned par0 = 0, par1 = 0;      // completed phases of the two barriers (warp-uniform)
  const int* cnt = lvlCnt + (size_t)slot * g.nlevels;
  // kpw keypoints per warp: 32 for large batches (the lane-parallel phase is fully used), fewer when
  // only a frame or two is in flight so that the keypoints spread over more warps (latency)
  const int base = wglobal * kpw;
  // this lane's keypoint: packed position base+lane -> (level, index inside the level)
  int level = -1, idx = base + lane, total = 0;
  for (int l = 0; l < g.nlevels; ++l) {
    const int c = cnt[l];
    if (level < 0 && idx < c) level = l;
    if (level < 0) idx -= c;
    total += c;
  }
  if (lane >= kpw) level = -1;
  if (wglobal == 0 && lane == 0) nKp[slot] = total;
  if (base >= total) return;  // whole warp exits together
  const int nk = min(kpw, total - base);
  int kx = 0, ky = 0, resp = 0;
  if (level >= 0) {
    const unsigned pk = lvlKp[(size_t)slot * g.totalOut + g.lv[level].outOff + idx];
    kx = ORBFE_PX(pk) + ORBFE_MINB; ky = ORBFE_PY(pk) + ORBFE_MINB;  // level coordinates
    resp = ORBFE_PS(pk);
  }
  // ---- E5: intensity centroid on the unblurred level.  A patch row's 31 bytes lie in 9 aligned words; lanes
  // = (row within a group of 3, word), so one load instruction touches 3 rows (few L1 wavefronts) and the
  // patch takes 11 of them.  Per-byte weights (u+15 inside the disc, 0 outside) and the 0/1 disc mask come
  // from a host-built table indexed by (alignment of the row start, |v|, word):
  // m10 = sum(u*I) = sum((u+15)*I) - 15*sum(I),  m01 = sum_v v * sum_u I  (exact integer sums).
  int my10 = 0, my01 = 0;
  const int rg = lane / 9, wi = lane - 9 * rg;  // lanes 27..31 idle in this phase
  // box of keypoint k (warp-uniform arguments) into buffer k & 1
  auto issue_pyr = [&](const int k, const int lv, const int cx, const int cy) {
    if (lane == 0) {
      const LevelGeom& L = g.lv[lv];
      OrbfeTmaPlane P;
      P.base = pyr + L.planeOff; P.sliceStride = g.pyrStride; P.pitch = L.pitch; P.rows = L.h + 2 * ORBFE_EDGE;
      P.slices = gridDim.y; P.boxW = ORBFE_DESC_PYR_BW; P.boxH = ORBFE_DESC_PYR_BH;
      const int col = cx + ORBFE_EDGE - ORBFE_HALF_PATCH;
      orbfe_tile_issue(wbuf[k & 1], &wbar[k & 1], tmPyr + lv, P, col & ~15, cy + ORBFE_EDGE - ORBFE_HALF_PATCH, slot);
    }
  };
  issue_pyr(0, __shfl_sync(0xffffffffu, level, 0), __shfl_sync(0xffffffffu, kx, 0), __shfl_sync(0xffffffffu, ky, 0));
  for (int k = 0; k < nk; ++k) {
    const int cx = __shfl_sync(0xffffffffu, kx, k);
    if (k + 1 < nk)  // buffer (k + 1) & 1 was last read in iteration k - 1, which ended with a __syncwarp
      issue_pyr(k + 1, __shfl_sync(0xffffffffu, level, k + 1), __shfl_sync(0xffffffffu, kx, k + 1), __shfl_sync(0xffffffffu, ky, k + 1));
    const int col = cx + ORBFE_EDGE - ORBFE_HALF_PATCH;  // padded column of a row's first pixel
    if (k & 1) { orbfe_tile_wait_warp(&wbar[1], par1 & 1u); ++par1; } else { orbfe_tile_wait_warp(&wbar[0], par0 & 1u); ++par0; }
    const unsigned* p0 = reinterpret_cast<const unsigned*>(wbuf[k & 1]) + ((col & 15) >> 2) + wi;
    const uint2* wt = icw + (col & 3) * (16 * 9) + wi;
    unsigned su = 0, s1 = 0;
    int m01 = 0;
    if (rg < 3) {
#pragma unroll
      for (int it = 0; it < 11; ++it) {
        const int r = 3 * it + rg;  // patch row 0..30 (v = r - 15)
        if (r < 31) {
          const int v = r - ORBFE_HALF_PATCH;
          const unsigned w = p0[r * (ORBFE_DESC_PYR_BW / 4)];
          const uint2 t = __ldg(wt + (v < 0 ? -v : v) * 9);
          su = __dp4a(w, t.x, su);
          const unsigned rs = __dp4a(w, t.y, 0u);
          s1 += rs;
          m01 += v * (int)rs;
        }
      }
    }
    int m10 = (int)su - ORBFE_HALF_PATCH * (int)s1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m10 += __shfl_xor_sync(0xffffffffu, m10, o);
      m01 += __shfl_xor_sync(0xffffffffu, m01, o);
    }
    if (lane == k) { my10 = m10; my01 = m01; }
    __syncwarp();
  }
  // ---- lane-parallel angle, sin/cos and keypoint record
  float angle = 0.f, a = 1.f, b = 0.f;
  int pos = base + lane;
  if (level >= 0) {
    angle = orbfe_fast_atan2((float)my01, (float)my10);
    const float factorPI = (float)(3.141592653589793238462643383279502884 / 180.f);  // :48
    const float ang = __fmul_rn(angle, factorPI);
    a = orbfe_sincosf(ang, 1);
    b = orbfe_sincosf(ang, 0);
    const LevelGeom& L = g.lv[level];
    orbfe_kp_dev kp;
    const float fx = (float)kx, fy = (float)ky;
    kp.x = level != 0 ? __fmul_rn(fx, L.scale) : fx;
    kp.y = level != 0 ? __fmul_rn(fy, L.scale) : fy;
    kp.size = L.kpSize;
    kp.angle = angle;
    kp.response = (float)resp;
    kp.octave = level;
    kp.class_id = -1;
    kps[(size_t)slot * g.totalOut + pos] = kp;
  }
  // ---- E7: rotated BRIEF on the blurred level; lane i -> descriptor byte i
#if ORBFE_DESC_PATSMEM
#define ORBFE_DESC_PAT(t) const float4 q_ = s_pat[t][lane]; const float qx_[2] = {q_.x, q_.z}, qy_[2] = {q_.y, q_.w};
#define ORBFE_DESC_PX(t, h) qx_[h]
#define ORBFE_DESC_PY(t, h) qy_[h]
#else
  float px[16], py[16];
  {
    const uint4 p0 = __ldg(reinterpret_cast<const uint4*>(d_orb_pattern) + 2 * lane);
    const uint4 p1 = __ldg(reinterpret_cast<const uint4*>(d_orb_pattern) + 2 * lane + 1);
    const unsigned w[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      px[t] = (float)(signed char)((w[t >> 1] >> (16 * (t & 1))) & 0xffu);
      py[t] = (float)(signed char)((w[t >> 1] >> (16 * (t & 1) + 8)) & 0xffu);
    }
  }
#define ORBFE_DESC_PAT(t)
#define ORBFE_DESC_PX(t, h) px[2 * (t) + (h)]
#define ORBFE_DESC_PY(t, h) py[2 * (t) + (h)]
#endif
  const uint8_t* blurSlot = blur + (size_t)slot * g.blurStride;
  uint8_t* dOut = desc + ((size_t)slot * g.totalOut + base) * 32;
  // the pattern reaches 18 px (|offset| <= 18 after rotation): keypoints at least 19 px inside the level need no edge
  // handling and take their window from the TMA box (warp-uniform test; ~97 % of the keypoints)
  auto inside = [&](const int lv, const int cx, const int cy) {
    const LevelGeom& L = g.lv[lv];
    return cx >= 18 && cy >= 18 && cx + 18 < L.w && cy + 18 < L.h;
  };
  auto issue_blur = [&](const int k, const int lv, const int cx, const int cy) {
    if (lane == 0) {
      const LevelGeom& L = g.lv[lv];
      OrbfeTmaPlane P;
      P.base = blur + L.blurOff; P.sliceStride = g.blurStride; P.pitch = L.bpitch; P.rows = L.h;
      // 48-byte rows whenever the window fits (first column at most 11 bytes past the 16-byte boundary: 3 keypoints in 4):
      // with a 64-byte pitch the bank of a sample depends on the row's parity only and the 512 scattered byte reads of a
      // keypoint collide twice as often (ncu: 19 M bank conflicts per 128 frames against 10 M)
      const bool narrow = ((cx - 18) & 15) + ORBFE_DESC_BLUR_BH <= ORBFE_DESC_PYR_BW;
      P.slices = gridDim.y; P.boxW = narrow ? ORBFE_DESC_PYR_BW : ORBFE_DESC_BLUR_BW; P.boxH = ORBFE_DESC_BLUR_BH;
      orbfe_tile_issue(wbuf[k & 1], &wbar[k & 1], tmBlur + (narrow ? g.nlevels + lv : lv), P, (cx - 18) & ~15, cy - 18, slot);
    }
  };
  {
    const int lv0 = __shfl_sync(0xffffffffu, level, 0), cx0 = __shfl_sync(0xffffffffu, kx, 0), cy0 = __shfl_sync(0xffffffffu, ky, 0);
    if (inside(lv0, cx0, cy0)) issue_blur(0, lv0, cx0, cy0);
  }
  for (int k = 0; k < nk; ++k) {
    const int lv = __shfl_sync(0xffffffffu, level, k);
    const int cx = __shfl_sync(0xffffffffu, kx, k), cy = __shfl_sync(0xffffffffu, ky, k);
    const float ca = __shfl_sync(0xffffffffu, a, k), sb = __shfl_sync(0xffffffffu, b, k);
    if (k + 1 < nk) {  // buffer (k + 1) & 1 was last read in iteration k - 1, which ended with a __syncwarp
      const int lvn = __shfl_sync(0xffffffffu, level, k + 1), cxn = __shfl_sync(0xffffffffu, kx, k + 1), cyn = __shfl_sync(0xffffffffu, ky, k + 1);
      if (inside(lvn, cxn, cyn)) issue_blur(k + 1, lvn, cxn, cyn);
    }
    const LevelGeom& L = g.lv[lv];
    const uint8_t* bplane = blurSlot + L.blurOff;
    const int W = L.w, Hh = L.h, bp = L.bpitch;
    unsigned val = 0;
    if (inside(lv, cx, cy)) {
      if (k & 1) { orbfe_tile_wait_warp(&wbar[1], par1 & 1u); ++par1; } else { orbfe_tile_wait_warp(&wbar[0], par0 & 1u); ++par0; }
      const int rowB = ((cx - 18) & 15) + ORBFE_DESC_BLUR_BH <= ORBFE_DESC_PYR_BW ? ORBFE_DESC_PYR_BW : ORBFE_DESC_BLUR_BW;  // as issued
      const uint8_t* centre = wbuf[k & 1] + 18 * rowB + 18 + ((cx - 18) & 15);
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        int tv[2];
        ORBFE_DESC_PAT(t)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float x = ORBFE_DESC_PX(t, h), y = ORBFE_DESC_PY(t, h);
          const int iy = __float2int_rn(__fadd_rn(__fmul_rn(x, sb), __fmul_rn(y, ca)));
          const int ix = __float2int_rn(__fsub_rn(__fmul_rn(x, ca), __fmul_rn(y, sb)));
          tv[h] = (int)centre[iy * rowB + ix];
        }
        val |= (unsigned)(tv[0] < tv[1]) << t;
      }
    } else {
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        int tv[2];
        ORBFE_DESC_PAT(t)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float x = ORBFE_DESC_PX(t, h), y = ORBFE_DESC_PY(t, h);
          const int iy = __float2int_rn(__fadd_rn(__fmul_rn(x, sb), __fmul_rn(y, ca)));
          const int ix = __float2int_rn(__fsub_rn(__fmul_rn(x, ca), __fmul_rn(y, sb)));
          // the reference samples a CONTINUOUS w x h clone (step == w): a column overshoot lands in
          // the adjacent row; a sample outside the buffer is UB there and defined as 0
          int col = cx + ix, row = cy + iy;
          if (col < 0) { col += W; --row; } else if (col >= W) { col -= W; ++row; }
          tv[h] = (row < 0 || row >= Hh) ? 0 : (int)__ldg(bplane + (size_t)row * bp + col);
        }
        val |= (unsigned)(tv[0] < tv[1]) << t;
      }
    }
    dOut[k * 32 + lane] = (uint8_t)val;
    __syncwarp();  // every lane has read buffer k & 1 before the box of keypoint k + 2 may land in it
  }
}
