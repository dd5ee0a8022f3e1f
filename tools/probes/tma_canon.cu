// TMA parameter probe built on the canonical CUDA programming guide example (cuda::barrier + cde):
//   ./tma_canon <elemBytes 1|4> <GW> <GH> <SW> <SH> <x> <y>     one tensor-map configuration per process
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda/barrier>
#include <cstdio>
#include <cstdlib>
#include <vector>
using barrier = cuda::barrier<cuda::thread_scope_block>;
namespace cde = cuda::device::experimental;
__global__ void kernel(const __grid_constant__ CUtensorMap tensor_map, int x, int y, int bytes, unsigned char* out) {
  extern __shared__ __align__(128) unsigned char smem_buffer[];
#pragma nv_diag_suppress static_var_with_dynamic_init
  __shared__ barrier bar;
  if (threadIdx.x == 0) { init(&bar, blockDim.x); cde::fence_proxy_async_shared_cta(); }
  __syncthreads();
  barrier::arrival_token token;
  if (threadIdx.x == 0) {
    cde::cp_async_bulk_tensor_2d_global_to_shared(smem_buffer, &tensor_map, x, y, bar);
    token = cuda::device::barrier_arrive_tx(bar, 1, bytes);
  } else token = bar.arrive();
  bar.wait(std::move(token));
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem_buffer[i];
}
int main(int argc, char** argv) {
  const int eb = atoi(argv[1]), GW = atoi(argv[2]), GH = atoi(argv[3]), SW = atoi(argv[4]), SH = atoi(argv[5]), x = atoi(argv[6]), y = atoi(argv[7]);
  std::vector<unsigned char> h((size_t)GW * GH * eb);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (unsigned char)((i * 2654435761u) >> 13);
  unsigned char *d, *o; cudaMalloc(&d, h.size()); cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice); cudaMalloc(&o, SH * SW * eb);
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  auto fn = (PFN_cuTensorMapEncodeTiled_v12000)p;
  CUtensorMap map{};
  cuuint64_t size[2] = {(cuuint64_t)GW, (cuuint64_t)GH}; cuuint64_t stride[1] = {(cuuint64_t)GW * eb}; cuuint32_t box[2] = {(cuuint32_t)SW, (cuuint32_t)SH}; cuuint32_t es[2] = {1, 1};
  CUresult r = fn(&map, eb == 4 ? CU_TENSOR_MAP_DATA_TYPE_INT32 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, size, stride, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("eb %d G %dx%d box %dx%d at (%d,%d): encode %d  ", eb, GW, GH, SW, SH, x, y, (int)r);
  if (r) { printf("\n"); return 1; }
  const int bytes = SW * SH * eb;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  kernel<<<1, 128, bytes>>>(map, x, y, bytes, o);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s  ", cudaGetErrorString(e));
  if (e == cudaSuccess) { std::vector<unsigned char> ho(bytes); cudaMemcpy(ho.data(), o, bytes, cudaMemcpyDeviceToHost);
    int mism = 0; for (int r2 = 0; r2 < SH; ++r2) for (int c = 0; c < SW * eb; ++c) {
      const int gx = x * eb + c, gy = y + r2; const unsigned char want = (gx < GW * eb && gy < GH) ? h[(size_t)gy * GW * eb + gx] : 0; mism += ho[r2 * SW * eb + c] != want; }
    printf("mismatches %d", mism); }
  printf("\n");
  return 0;
}
