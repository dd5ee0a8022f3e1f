// tma_probe.cu -- standalone check of the TMA plumbing in csrc/orbfe_tma.cuh on a real B200, one test per process
// (a faulting kernel kills the context):  ./tma_probe <test>   (see main for the list)
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_probe tools/probes/tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../slam_framework_b200/csrc/orbfe_tma.cuh"

struct Maps { CUtensorMap m[4]; };

// 0: mbarrier only
__global__ void k_mbar(int* out) {
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) orbfe_mbar_init(&bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(orbfe_smem_u32(&bar)) : "memory");
  orbfe_mbar_wait(&bar, 0);
  if (threadIdx.x == 0) *out = 123;
}
// 1: 1-D bulk copy
__global__ void k_bulk(const uint8_t* src, uint8_t* out, int bytes) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) orbfe_mbar_init(&bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    orbfe_mbar_expect_tx(&bar, bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(orbfe_smem_u32(smem)),
                 "l"(src), "r"(bytes), "r"(orbfe_smem_u32(&bar)) : "memory");
  }
  orbfe_mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}
// 2/3/4: tensor copy, map in a kernel parameter (2: 2-D map, 3: 3-D map) or in global memory (4)
template <int RANK>
__global__ void k_tensor(const __grid_constant__ Maps maps, const CUtensorMap* gmaps, int idx, int x, int y, int z, int boxW, int boxH, uint8_t* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) orbfe_mbar_init(&bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    const CUtensorMap* m = gmaps ? gmaps + idx : &maps.m[idx];
    if (gmaps) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(m) : "memory");
    orbfe_mbar_expect_tx(&bar, (unsigned)(boxW * boxH));
    if (RANK == 3) orbfe_tma_load_3d(smem, m, &bar, x, y, z);
    else
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                       orbfe_smem_u32(smem)), "l"(m), "r"(orbfe_smem_u32(&bar)), "r"(x), "r"(y) : "memory");
  }
  orbfe_mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < boxW * boxH; i += blockDim.x) out[i] = smem[i];
}

static int encode(CUtensorMap* out, void* base, int rank, int pitch, int rows, int slices, size_t stride, int boxW, int boxH) {
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return -1;
  OrbfeEncodeTiledFn fn = (OrbfeEncodeTiledFn)p;
  const cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)rows, (cuuint64_t)slices};
  const cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)stride};
  const cuuint32_t box[3] = {(cuuint32_t)boxW, (cuuint32_t)boxH, 1u};
  const cuuint32_t es[3] = {1u, 1u, 1u};
  return (int)fn(out, CU_TENSOR_MAP_DATA_TYPE_UINT8, rank, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                 CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

int main(int argc, char** argv) {
  const int test = argc > 1 ? atoi(argv[1]) : 0;
  const int boxW = argc > 2 ? atoi(argv[2]) : 256, boxH = argc > 3 ? atoi(argv[3]) : 46;
  const int pitch = 1280, rows = 414, slices = 2;
  const size_t stride = (size_t)pitch * rows + 512;
  std::vector<uint8_t> h(stride * slices);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
  uint8_t *d, *dout;
  cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
  cudaMalloc(&dout, 65536);
  cudaMemset(dout, 0xee, 65536);
  std::vector<uint8_t> o(65536);
  if (test == 0) {
    k_mbar<<<1, 128>>>((int*)dout);
    cudaError_t e = cudaDeviceSynchronize();
    printf("test 0 (mbarrier only): %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
  }
  if (test == 1) {
    cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
    k_bulk<<<1, 128, 16384>>>(d + 256, dout, 8192);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(o.data(), dout, 8192, cudaMemcpyDeviceToHost);
    int mism = 0; for (int i = 0; i < 8192; ++i) mism += o[i] != h[256 + i];
    printf("test 1 (1-D bulk copy): %s, %d mismatches\n", cudaGetErrorString(e), mism);
    return e != cudaSuccess || mism;
  }
  const int rank = test == 2 ? 2 : 3;
  Maps maps;
  for (int i = 0; i < 4; ++i) { int r = encode(&maps.m[i], d, rank, pitch, rows, slices, stride, boxW, boxH); if (r) { printf("encode failed %d\n", r); return 1; } }
  CUtensorMap* dmaps; cudaMalloc(&dmaps, sizeof(maps)); cudaMemcpy(dmaps, &maps, sizeof(maps), cudaMemcpyHostToDevice);
  const int coords[3][3] = {{32, 35, 0}, {1104, 390, 1}, {0, 0, 1}};  // x must be a multiple of 16 bytes
  int bad = 0;
  for (int c = 0; c < 3; ++c) {
    const int x = coords[c][0], y = coords[c][1], z = rank == 2 ? 0 : coords[c][2];
    cudaMemset(dout, 0xee, 65536);
    if (rank == 2) { cudaFuncSetAttribute(k_tensor<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, boxW * boxH);
      k_tensor<2><<<1, 128, boxW * boxH>>>(maps, nullptr, 2, x, y, z, boxW, boxH, dout); }
    else { cudaFuncSetAttribute(k_tensor<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, boxW * boxH);
      k_tensor<3><<<1, 128, boxW * boxH>>>(maps, test == 4 ? dmaps : nullptr, 2, x, y, z, boxW, boxH, dout); }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("test %d coords %d: %s\n", test, c, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(o.data(), dout, boxW * boxH, cudaMemcpyDeviceToHost);
    int mism = 0;
    for (int r = 0; r < boxH; ++r)
      for (int cc = 0; cc < boxW; ++cc) {
        const int gx = x + cc, gy = y + r;
        const uint8_t want = (gx < pitch && gy < rows) ? h[(size_t)z * stride + (size_t)gy * pitch + gx] : 0;
        mism += o[r * boxW + cc] != want;
      }
    printf("test %d box %dx%d coords (%d,%d,%d): %d mismatches\n", test, boxW, boxH, x, y, z, mism);
    bad += mism;
  }
  printf(bad ? "TMA PROBE FAILED\n" : "TMA PROBE OK\n");
  return bad ? 3 : 0;
}
