// orbfe_tma.cuh -- TMA (cp.async.bulk.tensor) + mbarrier plumbing for the tiled kernels (FAST, blur, pyramid).
//
// A kernel stages its pixel tile with ONE instruction: an elected thread arms an mbarrier with the byte count
// of the box and issues a 3-D tensor copy (x = byte column of the padded plane, y = row, z = image slot); the
// other threads do their set-up work and then wait on the barrier.  Out-of-bounds parts of a box are
// zero-filled by the hardware, which is what the callers rely on at the right / bottom plane edges.
//
// Tensor maps are encoded on the host (cuTensorMapEncodeTiled, looked up through cudaGetDriverEntryPoint so
// that the library does not link libcuda) and kept in device global memory, one per pyramid level.
//
// Under ORBFE_EMU (tests/emu, g++) the copy is a plain loop over the plane the map describes and the wait is
// a block barrier -- same call sites, same data.
#pragma once
#include <stdint.h>

#ifndef ORBFE_EMU
#include <cuda.h>
#else
struct alignas(64) CUtensorMap { unsigned long long opaque[16]; };
#endif

// what a tensor map covers (kept beside it for the emulated copy and for bounds asserts)
struct OrbfeTmaPlane {
  const uint8_t* base;       // slot 0, row 0, byte 0
  unsigned long long sliceStride;  // bytes between slots
  int pitch;                 // bytes between rows == extent of dimension 0
  int rows;                  // extent of dimension 1
  int slices;                // extent of dimension 2
  int boxW, boxH;            // box (bytes, rows); box depth is 1
};

#ifndef ORBFE_EMU
__device__ __forceinline__ unsigned orbfe_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void orbfe_mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(orbfe_smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void orbfe_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(orbfe_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void orbfe_mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "ORBFE_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra ORBFE_DONE_%=;\n"
      "bra ORBFE_WAIT_%=;\n"
      "ORBFE_DONE_%=:\n"
      "}\n" ::"r"(orbfe_smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void orbfe_tma_load_3d(void* dst, const CUtensorMap* map, unsigned long long* bar, int x, int y, int z) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          orbfe_smem_u32(dst)),
      "l"(map), "r"(orbfe_smem_u32(bar)), "r"(x), "r"(y), "r"(z)
      : "memory");
}
__device__ __forceinline__ void orbfe_tma_store_3d(const CUtensorMap* map, const void* src, int x, int y, int z) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(orbfe_smem_u32(src)), "r"(x), "r"(y), "r"(z)
               : "memory");
}
__device__ __forceinline__ void orbfe_tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void orbfe_tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void orbfe_tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory must be made visible to the async proxy before a TMA store reads them
__device__ __forceinline__ void orbfe_fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
#endif

// ---- host: tensor map of a stack of `slices` u8 planes (pitch x rows bytes each, `sliceStride` apart) ----------
#if !defined(ORBFE_EMU)
#include <cuda_runtime.h>
typedef CUresult (*OrbfeEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// returns 0 on success
static inline int orbfe_tma_encode(CUtensorMap* out, const OrbfeTmaPlane& P) {
  static OrbfeEncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return -1;
    fn = (OrbfeEncodeTiledFn)p;
  }
  const cuuint64_t dims[3] = {(cuuint64_t)P.pitch, (cuuint64_t)P.rows, (cuuint64_t)P.slices};
  const cuuint64_t strides[2] = {(cuuint64_t)P.pitch, (cuuint64_t)P.sliceStride};
  const cuuint32_t box[3] = {(cuuint32_t)P.boxW, (cuuint32_t)P.boxH, 1u};
  const cuuint32_t es[3] = {1u, 1u, 1u};
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<uint8_t*>(P.base), dims, strides, box, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (int)r;
}
#elif defined(ORBFE_EMU)
static inline int orbfe_tma_encode(CUtensorMap* out, const OrbfeTmaPlane&) { *out = CUtensorMap{}; return 0; }
#endif

// ---- the three calls the kernels use ---------------------------------------------------------------------------
// one thread, once per CTA, followed by a __syncthreads() before anybody issues / waits
__device__ __forceinline__ void orbfe_tile_barrier_init(unsigned long long* bar) {
#ifndef ORBFE_EMU
  orbfe_mbar_init(bar, 1);
#else
  *bar = 0;
#endif
}

// one thread, once per CTA before its first orbfe_tile_issue: the map lives in global memory (written by the host before the
// launch); order its generic-proxy image before the tensormap-proxy reads.  (Not per copy: the fence drops the cached descriptor.)
__device__ __forceinline__ void orbfe_tmap_acquire(const CUtensorMap* map) {
#ifndef ORBFE_EMU
  asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(map) : "memory");
#else
  (void)map;
#endif
}

// one thread: fetch the box whose first byte is (x, y) of slot z into dst (dense [boxH][boxW] bytes, 128-byte aligned)
__device__ __forceinline__ void orbfe_tile_issue(void* dst, unsigned long long* bar, const CUtensorMap* map, const OrbfeTmaPlane& P,
                                                 int x, int y, int z) {
#ifndef ORBFE_EMU
  orbfe_mbar_expect_tx(bar, (unsigned)(P.boxW * P.boxH));
  orbfe_tma_load_3d(dst, map, bar, x, y, z);
#else
  (void)map; (void)bar;
  uint8_t* d = reinterpret_cast<uint8_t*>(dst);
  for (int r = 0; r < P.boxH; ++r)
    for (int c = 0; c < P.boxW; ++c) {
      const int gx = x + c, gy = y + r;
      const bool in = gx >= 0 && gx < P.pitch && gy >= 0 && gy < P.rows && z >= 0 && z < P.slices;
      d[r * P.boxW + c] = in ? P.base[(size_t)z * P.sliceStride + (size_t)gy * P.pitch + gx] : (uint8_t)0;
    }
#endif
}

// the same for a barrier that belongs to ONE warp of the CTA (each warp streams its own boxes): every lane of that warp
__device__ __forceinline__ void orbfe_tile_wait_warp(unsigned long long* bar, unsigned parity) {
#ifndef ORBFE_EMU
  orbfe_mbar_wait(bar, parity);
#else
  (void)bar; (void)parity;
  __syncwarp();
#endif
}

// every thread that reads the tile; `parity` = number of completed phases of this barrier & 1
__device__ __forceinline__ void orbfe_tile_wait(unsigned long long* bar, unsigned parity) {
#ifndef ORBFE_EMU
  orbfe_mbar_wait(bar, parity);
#else
  (void)bar; (void)parity;
  __syncthreads();
#endif
}
