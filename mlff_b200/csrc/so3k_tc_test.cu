// Diagnostic entry point: one 128 x N x K split-bf16 GEMM tile on tcgen05 with the operand layouts the edge
// kernels use.  Exercised by tests/test_gpu_parity.py to pin descriptor / layout conventions on the device.
#include "so3k_internal.h"
#ifndef SO3K_EMU
#include "so3k_tc.cuh"

namespace {

// variant 0: B given as (N, K) row-major  -> K-major operand            D = A @ B^T
// variant 2: B given as (K, N) row-major  -> image stored K-major with rows = reduction index, consumed as an
//            MN-major operand                                            D = A @ B
// passes: 1 = hi*hi only, 3 = hi*hi + lo*hi + hi*lo
__global__ void __launch_bounds__(128)
k_tc_selftest(int N, int K, int variant, int passes, const float* __restrict__ A, const float* __restrict__ B,
              float* __restrict__ D) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t sboA = (uint32_t)(K / 8) * 128u;
  const int imgRows = variant == 2 ? K : N;       // rows of the B image
  const int imgCols = variant == 2 ? N : K;
  const uint32_t sboB = (uint32_t)(imgCols / 8) * 128u;
  const uint32_t szA = 16u * sboA, szB = (uint32_t)(imgRows / 8) * sboB;
  unsigned char* Ahi = smem;
  unsigned char* Alo = Ahi + szA;
  unsigned char* Bhi = Alo + szA;
  unsigned char* Blo = Bhi + szB;

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::mbar_fence_init();
  }
  // A image: thread = row
  for (int c = 0; c < K / 8; ++c) {
    float v[8];
    for (int q = 0; q < 8; ++q) v[q] = A[(size_t)tid * K + c * 8 + q];
    tc::store_chunk8(Ahi, Alo, tc::canon_off(tid, c, sboA), v);
  }
  for (int r = tid; r < imgRows; r += 128)
    for (int c = 0; c < imgCols / 8; ++c) {
      float v[8];
      for (int q = 0; q < 8; ++q) v[q] = B[(size_t)r * imgCols + c * 8 + q];
      tc::store_chunk8(Bhi, Blo, tc::canon_off(r, c, sboB), v);
    }
  if (variant == 3 || variant == 4) {              // zero A image for the TMEM clear of the layout probe
    for (uint32_t o = (uint32_t)tid * 16u; o < szA; o += 128u * 16u) *reinterpret_cast<uint4*>(Alo + o) = make_uint4(0, 0, 0, 0);
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0 && (variant == 3 || variant == 4)) {
    // Layout probe for M = 64 (round-2 groundwork: two 64-row half tiles per CTA).  TMEM is first zeroed by an
    // M = 128 MMA on an all-zero A image (the lo image, overwritten with zeros below by every thread before the
    // barrier when variant >= 3), then ONE M = 64 GEMM on rows 0..63 of A is written at lane offset 0 (variant 3) or
    // 16 (variant 4); the host sees on which of the 128 lanes the 64 result rows land.
    const uint32_t id128 = tc::make_idesc_bf16(128, N, 0, 0), id64 = tc::make_idesc_bf16(64, N, 0, 0);
    tc::mma_bf16(tmem, tc::make_desc(tc::smem_u32(Alo), 128u, sboA), tc::make_desc(tc::smem_u32(Bhi), 128u, sboB), id128, 0);
    const uint32_t dst = tmem + (variant == 4 ? (16u << 16) : 0u);
    uint32_t acc = 0;
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t aoff = (uint32_t)ks * 256u;
      tc::mma_bf16(dst, tc::make_desc(tc::smem_u32(Ahi) + aoff, 128u, sboA),
                   tc::make_desc(tc::smem_u32(Bhi) + aoff, 128u, sboB), id64, acc);
      acc = 1;
    }
    tc::commit(&bar);
  } else if (variant == 5) {
    // A operand from tensor memory (the layout the filter kernel's epilogue 1 writes IN PLACE over the GEMM1
    // accumulator): k-step ks (16 reduction indices) occupies the 16 columns [16 ks, 16 ks + 16) of the row's lane --
    // 8 columns of packed bf16 hi followed by 8 columns of packed bf16 lo.  D at column 128 (K, N <= 128).
    for (int ks = 0; ks < K / 16; ++ks) {
      uint32_t pk[16];
      for (int q = 0; q < 8; ++q)
        tc::split2(A[(size_t)tid * K + ks * 16 + 2 * q], A[(size_t)tid * K + ks * 16 + 2 * q + 1], &pk[q], &pk[8 + q]);
      tc::tmem_st16(tmem + ((uint32_t)(32 * warp) << 16) + (uint32_t)(16 * ks), pk);
    }
    tc::tmem_st_wait();
    tc::fence_before_sync();
    __syncthreads();
    if (tid == 0) {
      tc::fence_after_sync();
      const uint32_t idesc = tc::make_idesc_bf16(128, N, 0, 0);
      uint32_t acc = 0;
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint32_t aoff = (uint32_t)ks * 256u;
        const uint64_t b_hi = tc::make_desc(tc::smem_u32(Bhi) + aoff, 128u, sboB);
        const uint64_t b_lo = tc::make_desc(tc::smem_u32(Blo) + aoff, 128u, sboB);
        const uint32_t a_hi = tmem + (uint32_t)(16 * ks), a_lo = a_hi + 8u;
        tc::mma_bf16_ts(tmem + 128u, a_hi, b_hi, idesc, acc);
        acc = 1;
        if (passes == 3) {
          tc::mma_bf16_ts(tmem + 128u, a_lo, b_hi, idesc, 1);
          tc::mma_bf16_ts(tmem + 128u, a_hi, b_lo, idesc, 1);
        }
      }
      tc::commit(&bar);
    }
  } else if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_bf16(128, N, 0, variant == 2 ? 1 : 0);
    uint32_t acc = 0;
    for (int ks = 0; ks < K / 16; ++ks) {
      const uint32_t aoff = (uint32_t)ks * 256u;
      uint64_t a_hi = tc::make_desc(tc::smem_u32(Ahi) + aoff, 128u, sboA);
      uint64_t a_lo = tc::make_desc(tc::smem_u32(Alo) + aoff, 128u, sboA);
      uint64_t b_hi, b_lo;
      if (variant == 2) {      // MN-major: LBO = distance between 8-row k groups (= image SBO), SBO = 128
        const uint32_t boff = (uint32_t)ks * 2u * sboB;
        b_hi = tc::make_desc(tc::smem_u32(Bhi) + boff, sboB, 128u);
        b_lo = tc::make_desc(tc::smem_u32(Blo) + boff, sboB, 128u);
      } else {
        b_hi = tc::make_desc(tc::smem_u32(Bhi) + aoff, 128u, sboB);
        b_lo = tc::make_desc(tc::smem_u32(Blo) + aoff, 128u, sboB);
      }
      tc::mma_bf16(tmem, a_hi, b_hi, idesc, acc);
      acc = 1;
      if (passes == 3) {
        tc::mma_bf16(tmem, a_lo, b_hi, idesc, 1);
        tc::mma_bf16(tmem, a_hi, b_lo, idesc, 1);
      }
    }
    tc::commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::fence_after_sync();
  const uint32_t lane_base = (uint32_t)(32 * (warp & 3)) << 16;
  const uint32_t dcol = variant == 5 ? 128u : 0u;
  for (int c = 0; c < N; c += 8) {
    float v[8];
    tc::tmem_ld8(tmem + dcol + lane_base + (uint32_t)c, v);
    tc::tmem_ld_wait();
    for (int q = 0; q < 8; ++q) D[(size_t)tid * N + c + q] = v[q];
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}


// Micro-benchmark: cycles per tcgen05.mma (M=128, N, K=16, bf16) for the two operand layouts.
// layout 0: canonical no-swizzle (LBO=128, SBO=(K/8)*128) ; layout 2: SWIZZLE_128B K-major (64-wide K blocks, SBO=1024)
__global__ void __launch_bounds__(128)
k_tc_bench(int N, int layout, int reps, long long* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 256);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::mbar_fence_init(); }
  for (int k = tid; k < 48 * 1024 / 4; k += 128) reinterpret_cast<uint32_t*>(smem)[k] = 0x3f803f80u;
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc_bf16(128, N, 0, 0);
    const uint32_t a0 = tc::smem_u32(smem), b0 = a0 + 20 * 1024;      // A: 128 x 64 (16 KB) ; B: N x 64 (<= 32 KB... N<=224)
    long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
      for (int ks = 0; ks < 4; ++ks) {
        uint64_t ad, bd;
        if (layout == 0) {
          ad = tc::make_desc(a0 + ks * 256, 128u, 8 * 128u);
          bd = tc::make_desc(b0 + ks * 256, 128u, 8 * 128u);
        } else {
          ad = tc::make_desc(a0 + ks * 32, 16u, 1024u) | ((uint64_t)2 << 61);
          bd = tc::make_desc(b0 + ks * 32, 16u, 1024u) | ((uint64_t)2 << 61);
        }
        tc::mma_bf16(tmem, ad, bd, idesc, 1);
      }
    }
    tc::commit(&bar);
    tc::mbar_wait(&bar, 0);
    long long t1 = clock64();
    out[0] = t1 - t0;
    out[1] = (long long)reps * 4;
  }
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

}  // namespace
#endif

extern "C" int so3k_tc_selftest(int32_t N, int32_t K, int32_t variant, int32_t passes, const float* A, const float* B,
                                float* D, void* stream) {
#ifdef SO3K_EMU
  (void)N; (void)K; (void)variant; (void)passes; (void)A; (void)B; (void)D; (void)stream;
  return SO3K_ERR_CONFIG;
#else
  if (N % 16 || N < 16 || N > 256 || K % 16 || K < 16 || !A || !B || !D) return SO3K_ERR_ARG;
  if (variant == 5 && (N > 128 || K > 128)) return SO3K_ERR_ARG;
  size_t smem = 2 * (size_t)16 * (K / 8) * 128 + 2 * (size_t)N * K * 2 + 1024;
  if (smem > 227 * 1024) return SO3K_ERR_ARG;
  cudaFuncSetAttribute(k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  so3k_count_launch();
  k_tc_selftest<<<1, 128, smem, (cudaStream_t)stream>>>(N, K, variant, passes, A, B, D);
  return cudaGetLastError() == cudaSuccess ? SO3K_OK : SO3K_ERR_CUDA;
#endif
}

extern "C" int so3k_tc_bench(int32_t N, int32_t layout, int32_t reps, long long* out_dev, void* stream) {
#ifdef SO3K_EMU
  (void)N; (void)layout; (void)reps; (void)out_dev; (void)stream;
  return SO3K_ERR_CONFIG;
#else
  cudaFuncSetAttribute(k_tc_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  so3k_count_launch();
  k_tc_bench<<<1, 128, 64 * 1024, (cudaStream_t)stream>>>(N, layout, reps, out_dev);
  return cudaGetLastError() == cudaSuccess ? SO3K_OK : SO3K_ERR_CUDA;
#endif
}
