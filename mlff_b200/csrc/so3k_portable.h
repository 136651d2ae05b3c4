// Portability shim: the SIMT kernels compile either with nvcc for sm_100a (the product) or,
// with -DSO3K_EMU, with g++ against tests/emu/cuda_emu.h (test infrastructure: the same kernel
// source is executed on the CPU in the `-m "not gpu"` suite).  The tcgen05 kernels are
// CUDA-only and are excluded from the emulated build.
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef SO3K_EMU
#include "cuda_emu.h"
#define SO3K_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
  do {                                                                      \
    so3k_count_launch();                                                    \
    emu::launch(grid, block, smem, [&] { kernel(__VA_ARGS__); });           \
  } while (0)
#define SO3K_DYN_SMEM(T, name) \
  T* name = reinterpret_cast<T*>((reinterpret_cast<uintptr_t>(emu::g_dyn_smem) + 15) & ~uintptr_t(15))
#define SO3K_SET_MAX_SMEM(kernel, bytes) ((void)0)
#define SO3K_HD
#define SO3K_SHARED16 alignas(16) static
#else
#include <cuda_runtime.h>
#define SO3K_LAUNCH(kernel, grid, block, smem, stream, ...)                 \
  do {                                                                      \
    so3k_count_launch();                                                    \
    kernel<<<grid, block, smem, stream>>>(__VA_ARGS__);                     \
  } while (0)
#define SO3K_DYN_SMEM(T, name)                                        \
  extern __shared__ __align__(16) unsigned char so3k_dyn_smem_[];     \
  T* name = reinterpret_cast<T*>(so3k_dyn_smem_)
#define SO3K_SET_MAX_SMEM(kernel, bytes) \
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
#define SO3K_HD __host__ __device__
#define SO3K_SHARED16 __shared__ __align__(16)          // a shared-memory array read / written with 128-bit accesses
#endif

void so3k_count_launch();

static inline int so3k_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }
