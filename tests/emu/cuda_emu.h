// cuda_emu.h -- TEST-ONLY emulation of the small CUDA subset the orbfe kernels use, so that
// the *same kernel sources* (slam_framework_b200/csrc/*.cuh, orbfe_api.cu) can be compiled with
// g++ -DORBFE_EMU and their logic diffed against the oracle in the GPU-less build container.
//
// This is NOT a CPU fallback of the product: the shipped library (liborbfe.so) is built by nvcc
// only, never contains this header, and the Python package refuses to load the emulated
// library.  The emulated build is produced by tests/emu/build_emu.py into
// tests/emu/liborbfe_emu_TESTONLY.so and is loaded by tests/ only.
//
// Execution model: blocks run one after another; the threads of a block are cooperative fibers
// on one OS thread (hand-rolled x86-64 context switch), scheduled round-robin and switched at
// __syncthreads()/__syncwarp()/warp collectives.  Races are therefore NOT detected -- that is
// compute-sanitizer's job on the GPU box -- but barrier / collective logic is exercised for real.
#pragma once
#ifndef ORBFE_EMU
#error "cuda_emu.h is only for -DORBFE_EMU test builds"
#endif

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <vector>
#include <sys/mman.h>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define __constant__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct int2 { int x, y; };
struct int4 { int x, y, z, w; };
struct short4 { short x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline int2 make_int2(int a, int b) { return int2{a, b}; }
static inline int4 make_int4(int a, int b, int c, int d) { return int4{a, b, c, d}; }
static inline short4 make_short4(short a, short b, short c, short d) { return short4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }

namespace emu {

struct Warp {
  int live = 0;        // lanes that have not exited
  int count = 0;       // arrivals at the current warp barrier
  unsigned gen = 0;
  unsigned live_mask = 0;
  unsigned long long slot[32];
};
struct Fiber {
  void* sp = nullptr;
  char* stack = nullptr;
  uint3 tid{0, 0, 0};
  int lin = 0, lane = 0, warp = 0;
  bool done = false;
};
struct State {
  dim3 grid, block;
  uint3 bid{0, 0, 0};
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  Fiber* cur = nullptr;
  void* sched_sp = nullptr;
  int live = 0, bar_count = 0;
  unsigned bar_gen = 0;
  long long bar_acc = 0;  // accumulator for __syncthreads_count/or
  long long bar_result = 0;
  std::function<void()> body;
  unsigned char* dyn_smem = nullptr;
  size_t dyn_cap = 0;
};
inline State& S() { static State s; return s; }

extern "C" void orbfe_emu_switch(void** from_sp, void* to_sp);
#if defined(__x86_64__)
__asm__(
    ".text\n.weak orbfe_emu_switch\n.type orbfe_emu_switch,@function\norbfe_emu_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n  ret\n"
    ".size orbfe_emu_switch,.-orbfe_emu_switch\n");
#else
#error "cuda_emu.h: only x86-64 is supported"
#endif

inline void yield() { State& s = S(); orbfe_emu_switch(&s.cur->sp, s.sched_sp); }

inline void release_block_if_complete() {
  State& s = S();
  if (s.live > 0 && s.bar_count == s.live) { s.bar_count = 0; s.bar_result = s.bar_acc; s.bar_acc = 0; s.bar_gen++; }
}
inline void release_warp_if_complete(Warp& w) {
  if (w.live > 0 && w.count == w.live) { w.count = 0; w.gen++; }
}
inline void fiber_exit() {
  State& s = S();
  Fiber* f = s.cur;
  f->done = true;
  s.live--;
  Warp& w = s.warps[f->warp];
  w.live--;
  w.live_mask &= ~(1u << f->lane);
  release_block_if_complete();
  release_warp_if_complete(w);
  yield();
  abort();  // never resumed
}
inline void trampoline() {
  S().body();
  fiber_exit();
}
inline long long block_barrier(long long contrib) {
  State& s = S();
  const unsigned gen = s.bar_gen;
  s.bar_acc += contrib;
  s.bar_count++;
  release_block_if_complete();
  while (s.bar_gen == gen) yield();
  return s.bar_result;
}
inline void warp_barrier() {
  State& s = S();
  Warp& w = s.warps[s.cur->warp];
  const unsigned gen = w.gen;
  w.count++;
  release_warp_if_complete(w);
  while (w.gen == gen) yield();
}

constexpr size_t kStack = 128 * 1024;

// launches from different host threads (Frame's stereo constructor extracts on two std::threads) run one after the other
inline std::mutex& launch_mutex() { static std::mutex m; return m; }

template <class F>
void launch(dim3 grid, dim3 block, size_t smem, F&& f) {
  std::lock_guard<std::mutex> serialise(launch_mutex());
  State& s = S();
  s.grid = grid; s.block = block;
  const int nthreads = (int)(block.x * block.y * block.z);
  if ((int)s.fibers.size() < nthreads) {
    const size_t old = s.fibers.size();
    s.fibers.resize(nthreads);
    for (size_t i = old; i < s.fibers.size(); ++i) {
      void* m = mmap(nullptr, kStack, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
      if (m == MAP_FAILED) { perror("mmap"); abort(); }
      s.fibers[i].stack = (char*)m;
    }
  }
  if (smem > s.dyn_cap) { free(s.dyn_smem); s.dyn_smem = (unsigned char*)aligned_alloc(128, (smem + 127) / 128 * 128); s.dyn_cap = smem; }
  s.body = std::function<void()>(f);
  const int nwarps = (nthreads + 31) / 32;
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        s.bid = uint3{bx, by, bz};
        s.warps.assign(nwarps, Warp());
        s.live = nthreads; s.bar_count = 0; s.bar_gen = 0; s.bar_acc = 0;
        for (int t = 0; t < nthreads; ++t) {
          Fiber& fb = s.fibers[t];
          fb.lin = t; fb.lane = t & 31; fb.warp = t >> 5; fb.done = false;
          fb.tid = uint3{(unsigned)(t % block.x), (unsigned)((t / block.x) % block.y), (unsigned)(t / (block.x * block.y))};
          s.warps[fb.warp].live++;
          s.warps[fb.warp].live_mask |= 1u << fb.lane;
          // initial stack: 6 callee-saved slots, entry address, fake return address
          uintptr_t top = ((uintptr_t)fb.stack + kStack) & ~(uintptr_t)15;
          void** sp = (void**)(top - 64);
          for (int i = 0; i < 6; ++i) sp[i] = nullptr;
          sp[6] = (void*)&trampoline;
          sp[7] = nullptr;
          fb.sp = (void*)sp;
        }
        while (s.live > 0)
          for (int t = 0; t < nthreads; ++t) {
            Fiber& fb = s.fibers[t];
            if (fb.done) continue;
            s.cur = &fb;
            orbfe_emu_switch(&s.sched_sp, fb.sp);
          }
      }
  s.cur = nullptr;
}

template <class T>
inline T warp_exchange(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  State& s = S();
  Warp& w = s.warps[s.cur->warp];
  unsigned long long raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  w.slot[s.cur->lane] = raw;
  warp_barrier();
  unsigned long long got = raw;
  if (src_lane >= 0 && src_lane < 32 && (w.live_mask >> src_lane & 1)) got = w.slot[src_lane];
  warp_barrier();
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}

}  // namespace emu

#define threadIdx (emu::S().cur->tid)
#define blockIdx (emu::S().bid)
#define blockDim (emu::S().block)
#define gridDim (emu::S().grid)
#define warpSize 32

static inline void __syncthreads() { emu::block_barrier(0); }
static inline int __syncthreads_count(int p) { return (int)emu::block_barrier(p ? 1 : 0); }
static inline int __syncthreads_or(int p) { return emu::block_barrier(p ? 1 : 0) != 0; }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline unsigned __activemask() { return emu::S().warps[emu::S().cur->warp].live_mask; }
template <class T> static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return emu::warp_exchange(v, src); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) {
  return emu::warp_exchange(v, emu::S().cur->lane ^ m);
}
template <class T> static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
  const int l = emu::S().cur->lane + (int)d;
  return emu::warp_exchange(v, l < 32 ? l : emu::S().cur->lane);
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
  const int l = emu::S().cur->lane - (int)d;
  return emu::warp_exchange(v, l >= 0 ? l : emu::S().cur->lane);
}
static inline unsigned __ballot_sync(unsigned, int p) {
  emu::State& s = emu::S();
  emu::Warp& w = s.warps[s.cur->warp];
  w.slot[s.cur->lane] = p ? 1 : 0;
  emu::warp_barrier();
  unsigned r = 0;
  for (int i = 0; i < 32; ++i)
    if ((w.live_mask >> i & 1) && w.slot[i]) r |= 1u << i;
  emu::warp_barrier();
  return r;
}
static inline int __reduce_add_sync(unsigned, int v) {
  emu::State& s = emu::S();
  emu::Warp& w = s.warps[s.cur->warp];
  w.slot[s.cur->lane] = (unsigned long long)(long long)v;
  emu::warp_barrier();
  long long r = 0;
  for (int i = 0; i < 32; ++i)
    if (w.live_mask >> i & 1) r += (long long)w.slot[i];
  emu::warp_barrier();
  return (int)r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, !p) == 0; }

// ---- atomics (single OS thread => plain read-modify-write) ------------------------------
template <class T> static inline T atomicAdd(T* a, T v) { T o = *a; *a = o + v; return o; }
template <class T> static inline T atomicSub(T* a, T v) { T o = *a; *a = o - v; return o; }
template <class T> static inline T atomicMax(T* a, T v) { T o = *a; if (v > o) *a = v; return o; }
template <class T> static inline T atomicMin(T* a, T v) { T o = *a; if (v < o) *a = v; return o; }
template <class T> static inline T atomicOr(T* a, T v) { T o = *a; *a = o | v; return o; }
template <class T> static inline T atomicAnd(T* a, T v) { T o = *a; *a = o & v; return o; }
template <class T> static inline T atomicExch(T* a, T v) { T o = *a; *a = v; return o; }
template <class T> static inline T atomicCAS(T* a, T c, T v) { T o = *a; if (o == c) *a = v; return o; }

// ---- math / bit intrinsics ------------------------------------------------------------------
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }
static inline int __float2int_rn(float v) { return (int)lrintf(v); }
static inline int __float2int_rd(float v) { return (int)floorf(v); }
static inline int __float2int_ru(float v) { return (int)ceilf(v); }
static inline int __float2int_rz(float v) { return (int)v; }
static inline int __double2int_rn(double v) { return (int)lrint(v); }
static inline float __int2float_rn(int v) { return (float)v; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __ddiv_rn(double a, double b) { volatile double r = a / b; return r; }
static inline double __dsqrt_rn(double a) { volatile double r = std::sqrt(a); return r; }
static inline float __double2float_rn(double a) { return (float)a; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
template <class T> static inline T __ldg(const T* p) { return *p; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned sel) {
  const unsigned long long v = ((unsigned long long)b << 32) | a;
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) {
    const unsigned s = (sel >> (4 * i)) & 0xf;
    unsigned byte = (unsigned)(v >> (8 * (s & 7))) & 0xff;
    if (s & 8) byte = (byte & 0x80) ? 0xff : 0x00;
    r |= byte << (8 * i);
  }
  return r;
}
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
  const unsigned long long v = ((unsigned long long)hi << 32) | lo;
  return (unsigned)(v >> (sh & 31));
}
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
  for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xffu) * ((b >> (8 * i)) & 0xffu);
  return c;
}
static inline unsigned __dp2a_lo(unsigned a, unsigned b, unsigned c) {
  return c + (a & 0xffffu) * (b & 0xffu) + (a >> 16) * ((b >> 8) & 0xffu);
}
static inline unsigned __vabsdiffu4(unsigned a, unsigned b) {
  unsigned r = 0;
  for (int i = 0; i < 4; ++i) { const int x = (a >> (8 * i)) & 0xff, y = (b >> (8 * i)) & 0xff; r |= (unsigned)std::abs(x - y) << (8 * i); }
  return r;
}
static inline unsigned emu_u16x2(unsigned a, unsigned b, unsigned c, bool mx) {
  unsigned r = 0;
  for (int i = 0; i < 2; ++i) {
    const unsigned x = (a >> (16 * i)) & 0xffff, y = (b >> (16 * i)) & 0xffff, z = (c >> (16 * i)) & 0xffff;
    const unsigned v = mx ? std::max(x, std::max(y, z)) : std::min(x, std::min(y, z));
    r |= v << (16 * i);
  }
  return r;
}
static inline unsigned __vimin3_u16x2(unsigned a, unsigned b, unsigned c) { return emu_u16x2(a, b, c, false); }
static inline unsigned __vimax3_u16x2(unsigned a, unsigned b, unsigned c) { return emu_u16x2(a, b, c, true); }
static inline unsigned __vminu2(unsigned a, unsigned b) { return emu_u16x2(a, b, b, false); }
static inline unsigned __vmaxu2(unsigned a, unsigned b) { return emu_u16x2(a, b, b, true); }
static inline unsigned __vsub2(unsigned a, unsigned b) {
  return (((a & 0xffff) - (b & 0xffff)) & 0xffff) | ((((a >> 16) - (b >> 16)) & 0xffff) << 16);
}
static inline unsigned __vadd2(unsigned a, unsigned b) {
  return (((a & 0xffff) + (b & 0xffff)) & 0xffff) | ((((a >> 16) + (b >> 16)) & 0xffff) << 16);
}
static inline int __vimin3_s32(int a, int b, int c) { return std::min(a, std::min(b, c)); }
static inline int __vimax3_s32(int a, int b, int c) { return std::max(a, std::max(b, c)); }
using std::max;
using std::min;

// ---- runtime API subset ------------------------------------------------------------------------
typedef int cudaError_t;
typedef struct emuStream_st* cudaStream_t;
typedef struct emuEvent_st* cudaEvent_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaStreamNonBlocking = 1, cudaEventDefault = 0, cudaEventDisableTiming = 2, cudaHostAllocDefault = 0, cudaHostAllocPortable = 1 };
enum cudaDeviceAttr { cudaDevAttrMultiProcessorCount = 16 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline const char* cudaGetErrorString(cudaError_t e) { return e == 0 ? "no error" : "emulated CUDA error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 4; return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
template <class T> static inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)calloc(n ? n : 1, 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
template <class T> static inline cudaError_t cudaMallocHost(T** p, size_t n) { return cudaMalloc(p, n); }
template <class T> static inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memcpy(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpy2DAsync(void* d, size_t dp, const void* s, size_t sp, size_t w, size_t h, cudaMemcpyKind, cudaStream_t = nullptr) {
  for (size_t y = 0; y < h; ++y) std::memcpy((char*)d + y * dp, (const char*)s + y * sp, w);
  return cudaSuccess;
}
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = nullptr) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
template <class K> static inline cudaError_t cudaFuncSetAttribute(K, cudaFuncAttribute, int) { return cudaSuccess; }
