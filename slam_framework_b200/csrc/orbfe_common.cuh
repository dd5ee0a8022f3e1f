// orbfe_common.cuh -- shared definitions of the B200 ORB front-end kernels.
// Device memory layout (per extractor handle, S = max_images slots; all per-slot blocks are
// strided arrays so one launch covers (slot, level, tile)):
//   img      S x imgStride      u8   input frames (pitch = align16(max_w))
//   pyr      S x pyrStride      u8   8 padded planes (w+38)x(h+38), pitch align16, plane align256
//   blur     S x blurStride     u8   8 blurred planes w x h
//   cellCnt  S x totalCells     i32  FAST survivors per 30-px grid cell
//   cellList S x cellListStride u32  per-cell candidate lists (x | y<<12 | score<<24, relative
//                                    to the (16,16) detection origin)
//   cand     S x candStride     u32  per-level candidate list in reference order
//   knode    S x candStride     i32  quad-tree: owning node of each candidate
//   nodes... quad-tree generations, finals list (see k_octree.cuh)
//   lvlKp    S x totalOut       u32  distributed keypoints per level (packed like cand)
//   lvlCnt   S x nlevels        i32
//   kps/desc S x totalOut       cv::KeyPoint (28 B) / 32 B descriptors, level-major order
#pragma once
#include <stdint.h>

#ifndef ORBFE_EMU
#include <cuda_runtime.h>
#endif

#define ORBFE_MAX_LEVELS 12
#define ORBFE_EDGE 19        // EDGE_THRESHOLD, orb_extractor.cpp:15
#define ORBFE_MINB 16        // EDGE_THRESHOLD-3, orb_extractor.cpp:714
#define ORBFE_HALF_PATCH 15  // orb_extractor.cpp:14
#define ORBFE_PATCH 31       // orb_extractor.cpp:13

struct LevelGeom {
  int w, h;              // level size (orb_extractor.cpp:1056)
  int pitch;             // padded plane pitch (bytes)
  int bpitch;            // blurred plane pitch (bytes)
  unsigned planeOff;     // byte offset of the padded plane inside a slot's pyramid block
  unsigned blurOff;      // byte offset of the blurred plane inside a slot's blur block
  int lutXOff, lutYOff;  // offsets into the resize LUT (entries)
  int wlutOff, rlutOff, fastResize;  // per-word LUT of the fast resize kernel; 0 => generic k_pyramid_level
  int area2;             // exact 2x decimation (OpenCV executes INTER_AREA)
  // FAST grid (orb_extractor.cpp:714-728)
  int nCols, nRows, wCell, hCell, maxBX, maxBY;
  int wCellMagic;        // ceil(65536 / wCell): n / wCell == (n * magic) >> 16 for n < 256 (k_fast_cells)
  int cellBase, cellCap;
  int fG, fSegs, fastBase;  // FAST CTAs: fG cells per CTA, fSegs CTAs per cell row
  unsigned cellListOff;  // u32 entries
  // quad-tree (orb_extractor.cpp:480-704)
  int N, nIni, boxW, boxH;
  float hX;
  int candCap;
  unsigned candOff;      // u32 entries
  int nodeCap;
  unsigned nodeOff;      // nodes
  int outCap, outOff;    // distributed keypoints of this level
  // blur tiling
  int tilesX, tilesY, tileBase;
  // pyramid tiling (rows of 4-px words)
  int pyrWords, pyrBlockBase, pyrBlocks;
  // streaming resize (k_pyramid_strip): eligible, vertical segments, TMA box width (bytes), offset into the box-origin table
  int pyrStrip, pyrSegs, pyrBoxW, pyrBoxOff;
  float scale;           // mvScaleFactor[level]
  float kpSize;          // (float)(int)(PATCH_SIZE*scale), orb_extractor.cpp:776
};

struct Geom {
  int nlevels, w0, h0, iniTh, minTh;
  int imgPitch;
  int totalCells, totalTiles, totalOut, totalPyrBlocks, totalFast;
  int sortCap;           // power of two >= max(outCap, N) over levels
  int maxNodeCap;        // max over levels of nodeCap
  unsigned imgStride, pyrStride, blurStride, cellListStride, candStride, nodeStride;
  LevelGeom lv[ORBFE_MAX_LEVELS];
};

// resize LUT entry (cv::resize INTER_LINEAR fixed-point coefficients)
struct ResizeLut {
  int ofs;
  short c0, c1;
};

// packed candidate / keypoint: x (12 bits) | y (12 bits) << 12 | score (8 bits) << 24
__host__ __device__ __forceinline__ unsigned orbfe_pack(int x, int y, int s) {
  return (unsigned)x | ((unsigned)y << 12) | ((unsigned)s << 24);
}
#define ORBFE_PX(p) ((int)((p) & 0xfffu))
#define ORBFE_PY(p) ((int)(((p) >> 12) & 0xfffu))
#define ORBFE_PS(p) ((int)((p) >> 24))

__device__ __forceinline__ int orbfe_reflect101(int p, int n) {
  // cv::borderInterpolate(BORDER_REFLECT_101)
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}

// ---- block-wide exclusive scan of one int per thread (all threads must call) ---------------
// s_warp: shared int[33].  Returns the exclusive prefix; *total = block sum.
__device__ __forceinline__ int orbfe_block_exscan(int v, int* s_warp, int* total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // protect s_warp reuse across consecutive calls
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = lane < nw ? s_warp[lane] : 0;
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < nw) s_warp[lane] = winc - w;
    if (lane == 31) s_warp[32] = winc;
  }
  __syncthreads();
  *total = s_warp[32];
  return s_warp[wid] + inc - v;
}

// ---- block-wide bitonic sort, descending, of n2 (power of two) u64 keys in shared memory ----
// Two forms of the same network.  Throughput form (many CTAs per SM, the kernel is issue-bound): every compare-exchange step
// goes through shared memory with one barrier per step.  Latency form (a frame or two in flight, the kernel is bound by its
// barriers): steps with partner distance j >= 32 as above but one PAIR per thread; all the steps with j < 32 that follow inside
// a merge stage exchange between lanes of one warp (element i lives in lane i & 31, blockDim.x is a multiple of 32), so they
// run in registers with shuffles and no barrier in between: 19 barriers instead of 45 for 512 keys.  A/B on B200, quad-tree:
// single pair 55.8 -> 51.8 us with the latency form, 128 frames 0.119 -> 0.127 ms (64-bit shuffles cost more issue slots than
// the idle lanes of the shared-memory form), hence the switch.
__device__ __forceinline__ void orbfe_block_sort_desc(unsigned long long* s, int n2, const bool latencyForm) {
  if (!latencyForm) {
    for (int k = 2; k <= n2; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        __syncthreads();
        for (int i = threadIdx.x; i < n2; i += blockDim.x) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const unsigned long long a = s[i], b = s[ixj];
            const bool desc = (i & k) == 0;
            if (desc ? (a < b) : (a > b)) { s[i] = b; s[ixj] = a; }
          }
        }
      }
    __syncthreads();
    return;
  }
  for (int k = 2; k <= n2; k <<= 1) {
    int j = k >> 1;
    for (; j >= 32; j >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (n2 >> 1); t += blockDim.x) {
        const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1)), ixj = i | j;   // the t-th pair of this step
        const unsigned long long a = s[i], b = s[ixj];
        const bool desc = (i & k) == 0;
        if (desc ? (a < b) : (a > b)) { s[i] = b; s[ixj] = a; }
      }
    }
    __syncthreads();
    if (j > 0) {
      const int nIter = (n2 + (int)blockDim.x - 1) / (int)blockDim.x;   // every lane of a warp runs the same trip count
      for (int e = 0; e < nIter; ++e) {
        const int i = e * (int)blockDim.x + (int)threadIdx.x;
        unsigned long long v = i < n2 ? s[i] : 0ull;
        const bool desc = (i & k) == 0;
        for (int jj = j; jj > 0; jj >>= 1) {
          const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, jj);
          const bool lower = (i & jj) == 0;
          v = (desc == lower) ? (v > o ? v : o) : (v < o ? v : o);
        }
        if (i < n2) s[i] = v;
      }
    }
  }
  __syncthreads();
}

#ifdef ORBFE_EMU
#define ORBFE_DYN_SMEM(name) unsigned char* name = emu::S().dyn_smem
#else
#define ORBFE_DYN_SMEM(name) extern __shared__ __align__(128) unsigned char name[]  // 128: TMA destinations
#endif
