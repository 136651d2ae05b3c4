// tcgen05 / TMEM / mbarrier helpers (sm_100a only; inline PTX).
//
// Operand convention used by the edge kernels: bf16 operands in shared memory in the canonical NO-SWIZZLE layout
// made of 8x8-element core matrices (8 rows of 16 bytes):
//     byte offset (row r, column k) = (r / 8) * SBO + (k / 8) * LBO + (r % 8) * 16 + (k % 8) * 2
// "K-major" operand: r = M/N index, k = reduction index, LBO = distance between core matrices adjacent in K,
//                    SBO = distance between 8-row groups.
// The SAME memory image read as an "MN-major" operand (r = reduction index, k = M/N index) just swaps the roles
// of the two offsets, which lets W2 serve the forward GEMM (K-major) and its transpose in the backward (MN-major)
// from one shared-memory copy.
// fp32 precision is recovered with split operands: x = hi + lo (both bf16), product = hi*hi + lo*hi + hi*lo,
// accumulated in fp32 in TMEM (3 MMAs per K step).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- shared-memory matrix descriptor (SM100 "version 1", no swizzle) ----------------------------------
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;           // descriptor version (Blackwell)
  return d;                          // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0)
}

// ---- instruction descriptor, kind::f16, BF16 x BF16 -> F32 -------------------------------------------
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                       // D format  = F32
       | (1u << 7)                       // A format  = BF16
       | (1u << 10)                      // B format  = BF16
       | ((uint32_t)a_mn_major << 15)    // A major   (0 = K-major)
       | ((uint32_t)b_mn_major << 16)    // B major
       | ((uint32_t)(N >> 3) << 17)      // N >> 3
       | ((uint32_t)(M >> 4) << 24);     // M >> 4
}

__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// A operand from tensor memory ("TS"): row m of A lives in TMEM lane m, its K = 16 bf16 elements packed two per 32-bit
// column (element 2c in the low half of column c) in 8 consecutive columns starting at a_tmem; B from shared memory.
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// all previously issued MMAs of this thread arrive on the mbarrier when complete
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- TMEM allocation (one full warp) ---------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 8 consecutive fp32 columns of this thread's TMEM lane (warp w reads lanes 32*(w%4) .. +31)
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
  uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
               : "r"(taddr));
  v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
  v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// 16 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::
          "r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (tensor core reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// Bounded wait: returns false if the phase did not complete within ~2^26 polls (a lost commit would otherwise
// hang the GPU); callers raise a status flag and bail out.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  uint32_t spins = 0;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 26)) return false;
  } while (!done);
  return true;
}

// one non-blocking-ish probe of a phase (the hardware may suspend the thread for a bounded time)
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

// ---- split fp32 -> bf16 hi / lo ----------------------------------------------------------------------------
// hi = round(x) to bf16 ; lo = round(x - hi) to bf16.  x - hi is exact in fp32 and |lo| <= 2^-9 |x| with either sign
// (a truncating hi would make lo*lo a biased, coherently accumulating error).
__device__ __forceinline__ void split2(float x, float y, uint32_t* hi, uint32_t* lo) {
#ifndef SO3K_SPLIT_INT
  // packed conversion instruction (round-to-nearest-even): 6 instructions per pair; measured 2-4 % faster kernels than the
  // integer-pipe variant below (10 instructions per pair), which was written when F2FP was feared to share the MUFU pipe
  uint32_t h, l;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h) : "f"(y), "f"(x));
  const float rx = x - __uint_as_float(h << 16);
  const float ry = y - __uint_as_float(h & 0xFFFF0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(l) : "f"(ry), "f"(rx));
  *hi = h;
  *lo = l;
#else
  const uint32_t xh = (__float_as_uint(x) + 0x8000u) & 0xFFFF0000u;
  const uint32_t yh = (__float_as_uint(y) + 0x8000u) & 0xFFFF0000u;
  *hi = __byte_perm(xh, yh, 0x7632);                       // {y[31:16], x[31:16]}
  const float lx = x - __uint_as_float(xh);
  const float ly = y - __uint_as_float(yh);
  *lo = __byte_perm(__float_as_uint(lx) + 0x8000u, __float_as_uint(ly) + 0x8000u, 0x7632);
#endif
}
// 8 consecutive columns (one 16-byte core-matrix row) of a canonical K-major operand
__device__ __forceinline__ void store_chunk8(unsigned char* hi_base, unsigned char* lo_base, uint32_t off,
                                             const float* v) {
  uint4 h, l;
  split2(v[0], v[1], &h.x, &l.x);
  split2(v[2], v[3], &h.y, &l.y);
  split2(v[4], v[5], &h.z, &l.z);
  split2(v[6], v[7], &h.w, &l.w);
  *reinterpret_cast<uint4*>(hi_base + off) = h;
  *reinterpret_cast<uint4*>(lo_base + off) = l;
}
// byte offset of (row r, 8-column chunk c) in a canonical operand with LBO = 128 and the given SBO
__device__ __forceinline__ uint32_t canon_off(int r, int chunk, uint32_t sbo) {
  return (uint32_t)(r >> 3) * sbo + (uint32_t)chunk * 128u + (uint32_t)(r & 7) * 16u;
}

}  // namespace tc
