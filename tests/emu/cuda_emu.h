// Minimal CUDA-on-CPU emulation layer.  TEST INFRASTRUCTURE ONLY.
//
// The SIMT kernels under mlff_b200/csrc are written against a small portability header
// (so3k_portable.h).  With -DSO3K_EMU they compile with g++ against this file, so the *same
// kernel source* can be executed on the CPU in the `-m "not gpu"` test-suite to catch indexing
// and logic errors without a GPU.  The emulated library is built into tests/emu/_build/ and is
// loaded only by tests; the product loader (mlff_b200/_lib.py) never looks for it.
//
// Model: blocks run one after another; the threads of a block are real OS threads that meet at
// std::barrier for __syncthreads(); warp collectives go through a per-warp exchange buffer.
// Restrictions the kernels obey: 1-D blocks whose size is a multiple of 32, no early `return`
// ahead of a barrier / warp collective, full-mask warp collectives only.
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

struct uint3_ { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct double2 { double x, y; };
struct double3 { double x, y, z; };
static inline float2 make_float2(float x, float y) { return {x, y}; }
static inline float3 make_float3(float x, float y, float z) { return {x, y, z}; }
static inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
static inline int2 make_int2(int x, int y) { return {x, y}; }
static inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return 0; }
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void* p) { free(p); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }

namespace emu {
struct BlockCtx {
  unsigned nthreads;
  std::barrier<>* block_bar;
  std::vector<std::unique_ptr<std::barrier<>>> warp_bar;
  std::vector<uint64_t> xchg;  // 32 slots per warp
};
inline thread_local uint3_ t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline thread_local BlockCtx* t_ctx = nullptr;
inline unsigned char* g_dyn_smem = nullptr;

template <class Fn>
void launch(dim3 grid, dim3 block, size_t smem, Fn fn) {
  unsigned nt = block.x;
  if (block.y != 1 || block.z != 1 || nt % 32) { fprintf(stderr, "emu: bad block\n"); abort(); }
  std::vector<unsigned char> dyn(smem + 64);
  g_dyn_smem = dyn.data();
  BlockCtx ctx;
  ctx.nthreads = nt;
  std::barrier<> bb(nt);
  ctx.block_bar = &bb;
  for (unsigned w = 0; w < nt / 32; ++w) ctx.warp_bar.emplace_back(new std::barrier<>(32));
  ctx.xchg.assign(nt, 0);
  auto worker = [&](unsigned tid) {
    t_ctx = &ctx;
    t_threadIdx = {tid, 0, 0};
    t_blockDim = {nt, 1, 1};
    t_gridDim = {grid.x, grid.y, grid.z};
    for (unsigned bz = 0; bz < grid.z; ++bz)
      for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
          t_blockIdx = {bx, by, bz};
          fn();
          ctx.block_bar->arrive_and_wait();
        }
  };
  std::vector<std::thread> th;
  for (unsigned t = 1; t < nt; ++t) th.emplace_back(worker, t);
  worker(0);
  for (auto& t : th) t.join();
  g_dyn_smem = nullptr;
}

inline uint64_t warp_xchg(uint64_t v, int src_lane) {
  BlockCtx* c = t_ctx;
  unsigned tid = t_threadIdx.x, w = tid / 32;
  c->xchg[tid] = v;
  c->warp_bar[w]->arrive_and_wait();
  uint64_t r = c->xchg[w * 32 + (src_lane & 31)];
  c->warp_bar[w]->arrive_and_wait();
  return r;
}
template <class T> inline uint64_t to_bits(T v) { uint64_t b = 0; memcpy(&b, &v, sizeof(T)); return b; }
template <class T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static

static inline void __syncthreads() { emu::t_ctx->block_bar->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::t_ctx->warp_bar[emu::t_threadIdx.x / 32]->arrive_and_wait(); }
template <class T> static inline T __shfl_sync(unsigned, T v, int lane) { return emu::from_bits<T>(emu::warp_xchg(emu::to_bits(v), lane)); }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu::from_bits<T>(emu::warp_xchg(emu::to_bits(v), (emu::t_threadIdx.x & 31) ^ m)); }
template <class T> static inline T __shfl_down_sync(unsigned, T v, int d) {
  int l = emu::t_threadIdx.x & 31; int s = l + d; if (s > 31) s = l;
  return emu::from_bits<T>(emu::warp_xchg(emu::to_bits(v), s));
}
template <class T> static inline T __shfl_up_sync(unsigned, T v, int d) {
  int l = emu::t_threadIdx.x & 31; int s = l - d; if (s < 0) s = l;
  return emu::from_bits<T>(emu::warp_xchg(emu::to_bits(v), s));
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned r = 0;
  // 32 exchanges would be slow; gather through the exchange buffer once.
  emu::BlockCtx* c = emu::t_ctx; unsigned tid = emu::t_threadIdx.x, w = tid / 32;
  c->xchg[tid] = pred ? 1 : 0;
  c->warp_bar[w]->arrive_and_wait();
  for (int l = 0; l < 32; ++l) r |= (unsigned)(c->xchg[w * 32 + l] & 1) << l;
  c->warp_bar[w]->arrive_and_wait();
  return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, p) == 0xffffffffu; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }

static inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed); }
static inline double atomicAdd(double* p, double v) { return std::atomic_ref<double>(*p).fetch_add(v, std::memory_order_relaxed); }
static inline int atomicAdd(int* p, int v) { return std::atomic_ref<int>(*p).fetch_add(v, std::memory_order_relaxed); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_add(v, std::memory_order_relaxed); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return std::atomic_ref<unsigned long long>(*p).fetch_add(v, std::memory_order_relaxed); }
static inline int atomicMax(int* p, int v) { int o = *p; while (o < v && !std::atomic_ref<int>(*p).compare_exchange_weak(o, v)) {} return o; }
static inline int atomicOr(int* p, int v) { return std::atomic_ref<int>(*p).fetch_or(v, std::memory_order_relaxed); }
static inline int atomicExch(int* p, int v) { return std::atomic_ref<int>(*p).exchange(v); }

template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
template <class T> static inline T __ldcs(const T* p) { return *p; }
template <class T> static inline void __stcs(T* p, T v) { *p = v; }
// glibc already declares __expf/__cosf/__sinf (same semantics as the accurate versions): nothing to add.
static inline float __fdividef(float a, float b) { return a / b; }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __frcp_rn(float x) { return 1.0f / x; }
static inline float fmaf_(float a, float b, float c) { return fmaf(a, b, c); }
static inline void sincospif(float x, float* s, float* c) { *s = sinf((float)M_PI * x); *c = cosf((float)M_PI * x); }
static inline void sincospi(double x, double* s, double* c) { *s = sin(M_PI * x); *c = cos(M_PI * x); }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
template <class T> static inline T min(T a, T b) { return a < b ? a : b; }
template <class T> static inline T max(T a, T b) { return a > b ? a : b; }
