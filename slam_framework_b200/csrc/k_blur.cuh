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
// One CTA (one warp) = a column strip of 32 output words (128 px) x one vertical segment of a level (about
// ORBFE_BLUR_SEG rows: whole-height strips leave the SMs idle for the last 20 % of the launch).  The input
// arrives as a stream of TMA boxes (144 bytes x ORBFE_BLUR_RB rows; first byte = padded column 16 + 128*tx, a
// multiple of 16 as TMA requires) through a ring of ORBFE_BLUR_NB shared-memory buffers, each with its own
// mbarrier (orbfe_tma.cuh): box b + NB is requested as soon as box b has been consumed, so the copy of the next
// rows overlaps the arithmetic on the current ones and only the 6 halo rows of a segment are fetched twice.  A lane owns
// one output word and marches down the rows; per input row it reads its three words from shared memory and forms
//   * the 4 horizontal sums with 10 IDP.4A against byte-shifted copies of the HALVED taps
//     [9,17,24,28,24,17,9] (every tap is even, so H = 2 H' exactly; the shifted constants replace the
//     byte funnel-shifts of the data);
//   * the column sum from PAIRED half sums: with A_t = H'_t | H'_{t+1} << 16 (one PRMT per row), the
//     tap symmetry folds the 7 products into two IDP.2A,
//       V = 36 (H'_0 + H'_6) + 68 (H'_1 + H'_5) + 96 (H'_2 + H'_4) + 112 H'_3
//         = dp2a(A_0 + swap(A_5), 36|68) + dp2a(A_2 + H'_4, 96|112),
//     because H' <= 32640 lets two rows be added inside a 16-bit lane without a carry.
// 4.5 multiply-pipe instructions per pixel instead of 7: the kernel was bound by that pipe (ncu:
// fmaheavy 61 % of elapsed in the previous form), not by HBM.
#pragma once
#include "orbfe_common.cuh"
#include "orbfe_tma.cuh"

#define ORBFE_BLUR_THREADS 32
#define ORBFE_BLUR_WORDS 32   // output words per strip: every lane produces one
#define ORBFE_BLUR_BOXW 144   // bytes per box row: 32 + 2 words, rounded up to 16 bytes
#ifndef ORBFE_BLUR_RB
#define ORBFE_BLUR_RB 36      // rows per box; a multiple of 6 (the march is unrolled by its 6-row register ring)
#endif
#ifndef ORBFE_BLUR_NB
#define ORBFE_BLUR_NB 2       // boxes in flight.  A/B on B200 (128 frames): whole-height strips 0.129 ms; segments of 112 rows with
                              // NB 3 -> 0.116, NB 2 -> 0.111; RB 18/24/36 and SEG 64..180 within 0.111..0.123 (the previous kernel: 0.167)
#endif
#ifndef ORBFE_BLUR_SEG
#define ORBFE_BLUR_SEG 112    // target output rows per segment
#endif
#define ORBFE_BLUR_TW (4 * ORBFE_BLUR_WORDS)

// per-CTA work descriptor (host-built): level | segment << 4 | strip << 16
__host__ __device__ __forceinline__ unsigned orbfe_blur_task(int level, int seg, int tx) {
  return (unsigned)level | ((unsigned)seg << 4) | ((unsigned)tx << 16);
}

struct BlurConsts { unsigned k[13]; };
static inline BlurConsts orbfe_blur_consts() {
  BlurConsts c = {{0x1c181109u, 0x00091118u,               // pixel 0: taps on w0 bytes 0..3, w1 bytes 0..2
                   0x18110900u, 0x0911181cu,               // pixel 1: w0 bytes 1..3, w1 bytes 0..3
                   0x11090000u, 0x11181c18u, 0x00000009u,  // pixel 2: w0 bytes 2..3, w1, w2 byte 0
                   0x09000000u, 0x181c1811u, 0x00000911u,  // pixel 3: w0 byte 3, w1, w2 bytes 0..1
                   36u | (68u << 8), 96u | (112u << 8),    // column weights on the paired half sums
                   32768u}};                               // rounding
  return c;
}

#ifndef ORBFE_BLUR_MINB
#define ORBFE_BLUR_MINB 16
#endif
__global__ void __launch_bounds__(ORBFE_BLUR_THREADS, ORBFE_BLUR_MINB)
k_blur(const __grid_constant__ Geom g, const uint8_t* __restrict__ pyr, const CUtensorMap* __restrict__ tmaps,
       const unsigned* __restrict__ tasks, uint8_t* __restrict__ blur, const __grid_constant__ BlurConsts kc) {
  // every buffer starts on a 128-byte boundary (TMA destination)
  __shared__ __align__(128) unsigned s_tile[ORBFE_BLUR_NB][(ORBFE_BLUR_RB * (ORBFE_BLUR_BOXW / 4) + 31) / 32 * 32];
  __shared__ __align__(8) unsigned long long s_bar[ORBFE_BLUR_NB];
  const int tid = threadIdx.x, slot = blockIdx.y;
  const unsigned task = __ldg(tasks + blockIdx.x);
  const int level = task & 15, seg = (task >> 4) & 0xfff, tx = task >> 16;
  const LevelGeom& L = g.lv[level];
  const int segH = (L.h + L.tilesY - 1) / L.tilesY;  // L.tilesY = segments of this level
  const int y0 = seg * segH, nrows = min(segH, L.h - y0);
  const int total = nrows + 6;  // input rows: padded rows 16 + y0 .. 16 + y0 + nrows + 5
  const int by = ORBFE_EDGE - 3 + y0;
  const int nBoxes = (total + ORBFE_BLUR_RB - 1) / ORBFE_BLUR_RB;
  OrbfeTmaPlane P;
  P.base = pyr + L.planeOff; P.sliceStride = g.pyrStride; P.pitch = L.pitch; P.rows = L.h + 2 * ORBFE_EDGE;
  P.slices = gridDim.y; P.boxW = ORBFE_BLUR_BOXW; P.boxH = ORBFE_BLUR_RB;
  const int bx = 16 + ORBFE_BLUR_TW * tx;
  if (tid == 0)
    for (int b = 0; b < ORBFE_BLUR_NB; ++b) orbfe_tile_barrier_init(&s_bar[b]);
  __syncthreads();
  if (tid == 0) orbfe_tmap_acquire(tmaps + level);
  if (tid == 0)
    for (int b = 0; b < ORBFE_BLUR_NB && b < nBoxes; ++b)
      orbfe_tile_issue(s_tile[b], &s_bar[b], tmaps + level, P, bx, by + b * ORBFE_BLUR_RB, slot);
  const int k = tx * ORBFE_BLUR_WORDS + tid;  // output word of this lane
  const bool writer = 4 * k < L.w;
  // bpitch is a multiple of 16: the word store is aligned; bytes past w land in row padding
  uint8_t* dst = blur + (size_t)slot * g.blurStride + L.blurOff + (size_t)y0 * L.bpitch + 4 * k;
  const int bpitch = L.bpitch;

  // halved taps 9,17,24,28,24,17,9 against the bytes p .. p+6 of (w0, w1, w2), p = 0..3.  They arrive as a kernel
  // parameter so that IDP reads them straight from the constant bank: as literals ptxas re-materialises them into uniform
  // registers on every row (+13 instructions per row, measured in SASS)
  const unsigned K00 = kc.k[0], K01 = kc.k[1], K10 = kc.k[2], K11 = kc.k[3], K20 = kc.k[4], K21 = kc.k[5], K22 = kc.k[6];
  const unsigned K30 = kc.k[7], K31 = kc.k[8], K32 = kc.k[9], W1 = kc.k[10], W2 = kc.k[11], RND = kc.k[12];
  unsigned A[6][4], H[3][4];  // A[t % 6] = H'_t | H'_{t+1} << 16;  H[t % 3] = H'_t
#pragma unroll
  for (int q = 0; q < 6; ++q)
#pragma unroll
    for (int j = 0; j < 4; ++j) A[q][j] = 0u;
#pragma unroll
  for (int q = 0; q < 3; ++q)
#pragma unroll
    for (int j = 0; j < 4; ++j) H[q][j] = 0u;
  for (int box = 0; box < nBoxes; ++box) {
    const int buf = box % ORBFE_BLUR_NB;
    orbfe_tile_wait(&s_bar[buf], (unsigned)(box / ORBFE_BLUR_NB) & 1u);
    const unsigned* src = s_tile[buf] + tid;
    const int sEnd = min(total, (box + 1) * ORBFE_BLUR_RB);
    for (int s0 = box * ORBFE_BLUR_RB; s0 < sEnd; s0 += 6) {
#pragma unroll
    for (int u = 0; u < 6; ++u) {
      const int s = s0 + u;
      if (s < sEnd) {  // block-uniform
        const unsigned w0 = src[0], w1 = src[1], w2 = src[2];
        src += ORBFE_BLUR_BOXW / 4;
        unsigned Hn[4];
        Hn[0] = __dp4a(w1, K01, __dp4a(w0, K00, 0u));
        Hn[1] = __dp4a(w1, K11, __dp4a(w0, K10, 0u));
        Hn[2] = __dp4a(w2, K22, __dp4a(w1, K21, __dp4a(w0, K20, 0u)));
        Hn[3] = __dp4a(w2, K32, __dp4a(w1, K31, __dp4a(w0, K30, 0u)));
        // u == s % 6 (s0 is a multiple of 6): A_{s-6} = A[u], A_{s-4} = A[(u+2)%6], A_{s-1} -> A[(u+5)%6];
        // H'_{s-1} = H[(u+2)%3], H'_{s-2} = H[(u+1)%3], H'_s -> H[u%3]
        unsigned acc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const unsigned hm1 = H[(u + 2) % 3][j], hm2 = H[(u + 1) % 3][j];
          const unsigned An = __byte_perm(hm1, Hn[j], 0x5410);  // H'_{s-1} | H'_s << 16
          const unsigned Dn = __byte_perm(Hn[j], hm1, 0x5410);  // H'_s | H'_{s-1} << 16
          const unsigned S1 = A[u][j] + Dn;                     // (H'_{s-6} + H'_s) | (H'_{s-5} + H'_{s-1}) << 16
          const unsigned T2 = A[(u + 2) % 6][j] + hm2;          // (H'_{s-4} + H'_{s-2}) | H'_{s-3} << 16
          acc[j] = __dp2a_lo(S1, W1, __dp2a_lo(T2, W2, RND));   // V + 32768 < 2^24: the rounded result is byte 2
          A[(u + 5) % 6][j] = An;
          H[u % 3][j] = Hn[j];
        }
        if (s >= 6 && writer) {  // s counts rows of this segment: the loop below runs s = 0 .. total - 1
          const unsigned out = __byte_perm(__byte_perm(acc[0], acc[1], 0x0062), __byte_perm(acc[2], acc[3], 0x0062), 0x5410);
          *reinterpret_cast<unsigned*>(dst) = out;
          dst += bpitch;
        }
      }
    }
    }
    // the buffer is free once every lane has read its last row: request the box NB ahead into it
    __syncwarp();
    if (tid == 0 && box + ORBFE_BLUR_NB < nBoxes)
      orbfe_tile_issue(s_tile[buf], &s_bar[buf], tmaps + level, P, bx, by + (box + ORBFE_BLUR_NB) * ORBFE_BLUR_RB, slot);
  }
}
