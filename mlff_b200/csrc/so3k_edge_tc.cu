// The dense part of the edge path on the 5th-generation tensor cores (tcgen05.mma + TMEM): both filter MLPs and
// their transposes, evaluated once per UNDIRECTED pair (the filter depends on d_ij and on |chi_j - chi_i|^2 only, so
// w_ij = w_ji).  Same math as the fp32 CUDA-core kernels of so3k_edge.cu:
//   chi_ij, l0 contraction     mlff/nn/layer/so3krates_layer.py:89-93 ; mlff/sph_ops/contract.py:181-188
//   RadialSphericalFilter      mlff/nn/layer/so3krates_layer.py:449-465 ; MLP mlff/nn/mlp/mlp.py:21-27
//   backward                   SURVEY.md appendix B (reverse of the above)
// Everything that gathers or reduces over neighbours (attention coefficients mlff/nn/layer/so3krates_layer.py:568-586,
// aggregation, q/k gradients) lives in the streaming kernels of so3k_agg.cu, which exchange pair-indexed arrays with
// the kernels below: filters w[pair][F] (out), filter cotangents wbar[pair][F] (in).
//
// Design (one filter block -- fb or gb -- per launch, persistent CTAs, 1 CTA / SM, 512 threads):
//   * the block's weights live in shared memory for the whole kernel as bf16 hi/lo operand images in the canonical
//     no-swizzle K-major layout (so3k_tc.cuh; measured at the tcgen05 issue floor, profiles/tc_mma_layout_bench.py):
//     W1 (N=Fp, K=n_rbf), Ws1 (N=NS, K=16: the first layer of the spherical branch) and W2c (N=Fp, K=KC) -- the radial
//     and spherical second layers stacked along K, so w = [silu(a1) | silu(a_s)] @ W2c is ONE GEMM.  The backward reads
//     the same W2c image as an MN-major operand (W2c^T) -- no transposed copy;
//   * a tile = 128 consecutive pairs = M of one UMMA (M=128, cta_group::1).  Row r of the tile is TMEM lane r; warp w
//     works on lane quadrant (w & 3) and on column quarter (w >> 2): 4 threads per row;
//   * every product is 3 MMAs (hi*hi + lo*hi + hi*lo) accumulating in fp32 in TMEM;
//   * both first layers run as MMAs into one accumulator whose column k is hidden unit k (the spherical one is written
//     over the radial padding columns), so epilogue 1 is one uniform loop: tcgen05.ld, + bias, silu, hi/lo split,
//     store of the NEXT GEMM's A-operand image;
//   * per-pair inputs come from 32-byte records (k_pair_records), prefetched one tile ahead: no dependent loads.
// Kernels:
//   k_pair_records    per pair: d and the l0 contractions of (chi_j - chi_i)
//   k_filter_tc       records -> operand images -> GEMM1 -> silu -> GEMM2 -> + bias -> staging tile -> one bulk copy
//                     (cp.async.bulk) of the tile's contiguous 128 x F block of w
//   k_edge_bwd2_tc    wbar tile -> operand image -> GEMM3 (hbar = wbar W2c^T); pre-activations and the forward-mode
//                     radial derivative from two GEMM1s (rbf, rbf'); epilogue: radial cotangent dbar and l0-contraction
//                     cotangents gbar, added onto the canonical edge of the pair
// Algorithmic FLOPs per k_filter_tc launch: 2 C (K F + F^2 + nl F/4 + F^2/4), C pairs; the tensor pipe
// issues 3x that on padded shapes (Fp x KC).
#include "so3k_internal.h"
#include "so3k_edge.h"
#ifndef SO3K_EMU
#include "so3k_tc.cuh"
#include <cstdio>
#include <cstring>

namespace {

constexpr int NC = 512;        // compute threads per CTA (16 warps: 4 TMEM lane quadrants x 4 column quarters)
constexpr int NW = NC / 32;    // compute warps
constexpr int NT = NC;         // (a 17th, issue-only warp was tried: the warp-allocation granularity of 4 then caps registers at 96/thread)
constexpr int ISSUER = 0;      // thread that issues the MMAs
constexpr int TM = 128;        // pairs per tile

struct TcDims {
  int Fp, KC, K0, NS;          // padded N, padded K of GEMM2, K of GEMM1, padded width of the spherical hidden layer
  uint32_t sboW1, sboW2, sboA0, sboA1;
  uint32_t szW1, szW2, szA0, szA1;   // bytes of ONE (hi or lo) image
  uint32_t szWs, szA0s;              // spherical first layer: weight image (NS x 16) and operand image (TM x 16)
  uint32_t off_W1, off_W2, off_Ws, off_A0s, off_A1, off_small, total;
};
constexpr uint32_t SBO_S = 256u;     // 8-row group stride of the K = 16 images of the spherical first layer

__host__ __device__ inline TcDims make_tc_dims(const So3kDims& D) {
  TcDims t;
  t.Fp = (D.F + 15) & ~15;
  t.KC = (D.F + D.Fs + 15) & ~15;
  t.K0 = D.K;
  t.NS = (D.Fs + 15) & ~15;
  t.sboW1 = (uint32_t)(t.K0 / 8) * 128u;
  t.sboW2 = (uint32_t)(t.KC / 8) * 128u;
  t.sboA0 = t.sboW1;
  t.sboA1 = t.sboW2;
  t.szW1 = (uint32_t)(t.Fp / 8) * t.sboW1;
  t.szW2 = (uint32_t)(t.Fp / 8) * t.sboW2;
  t.szA0 = (uint32_t)(TM / 8) * t.sboA0;
  t.szA1 = (uint32_t)(TM / 8) * t.sboA1;
  t.szWs = (uint32_t)(t.NS / 8) * SBO_S;
  t.szA0s = (uint32_t)(TM / 8) * SBO_S;
  t.off_W1 = 0;
  t.off_W2 = t.off_W1 + 2 * t.szW1;
  t.off_Ws = t.off_W2 + 2 * t.szW2;
  t.off_A0s = t.off_Ws + 2 * t.szWs;                                     // own buffer: written by S0 of the next tile
  t.off_A1 = t.off_A0s + 2 * t.szA0s;
  uint32_t region = 2 * t.szA1;                                         // A1 hi | lo
  const uint32_t wb_bytes = 2u * (uint32_t)(TM / 8) * (uint32_t)(t.Fp / 8) * 128u;   // backward 2: wbar image
  if (wb_bytes > region) region = wb_bytes;
  if (4u * t.szA0 > region) region = 4u * t.szA0;                        // backward 2: rbf and rbf' images
  t.off_small = t.off_A1 + ((region + 127u) & ~127u);
  // small arrays (floats): b1x (KC), b2p (Fp), Ws1 (nl*Fs)
  uint32_t small = 4u * ((uint32_t)t.KC + t.Fp + (uint32_t)D.nl * D.Fs);
  t.total = t.off_small + ((small + 127u) & ~127u);
  return t;
}

// ---- weight images (built once at so3k_load_params) -------------------------------------------------------
// image layout in global memory == shared-memory layout: [W1 hi | W1 lo | W2c hi | W2c lo | Ws1 hi | Ws1 lo], then fp32
// [b1x (KC) | b2p (Fp)], b1x = [rad_b1 (F) | sph_b1 (Fs) | 0] indexed like the hidden units (K index of GEMM2)
__global__ void k_build_images(So3kDims D, TcDims T, const float* __restrict__ rW1, const float* __restrict__ rb1,
                               const float* __restrict__ rW2, const float* __restrict__ rb2,
                               const float* __restrict__ sW1, const float* __restrict__ sb1,
                               const float* __restrict__ sW2, const float* __restrict__ sb2,
                               unsigned char* __restrict__ img, float* __restrict__ bias) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int F = D.F;
  // W1 image: rows n = f (Fp), cols k (K0): value rad_W1[k][f]
  if (t < T.Fp * T.K0) {
    int n = t / T.K0, k = t % T.K0;
    float v = (n < F) ? rW1[(size_t)k * F + n] : 0.f;
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    uint32_t off = tc::canon_off(n, k >> 3, T.sboW1) + (uint32_t)(k & 7) * 2u;
    *reinterpret_cast<__nv_bfloat16*>(img + T.off_W1 + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(img + T.off_W1 + T.szW1 + off) = l;
  }
  // W2c image: rows n = f (Fp), cols k = c (KC): c < F: rad_W2[c][f] ; F <= c < F+Fs: sph_W2[c-F][f] ; else 0
  if (t < T.Fp * T.KC) {
    int n = t / T.KC, c = t % T.KC;
    float v = 0.f;
    if (n < F) {
      if (c < F) v = rW2[(size_t)c * F + n];
      else if (c - F < D.Fs) v = sW2[(size_t)(c - F) * F + n];
    }
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    uint32_t off = tc::canon_off(n, c >> 3, T.sboW2) + (uint32_t)(c & 7) * 2u;
    *reinterpret_cast<__nv_bfloat16*>(img + T.off_W2 + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(img + T.off_W2 + T.szW2 + off) = l;
  }
  // Ws1 image: rows n = c (NS), cols k = l (16): value sph_W1[l][c]
  if (t < T.NS * 16) {
    int n = t / 16, k = t % 16;
    float v = (n < D.Fs && k < D.nl) ? sW1[(size_t)k * D.Fs + n] : 0.f;
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
    uint32_t off = (uint32_t)(n >> 3) * SBO_S + (uint32_t)(k >> 3) * 128u + (uint32_t)(n & 7) * 16u + (uint32_t)(k & 7) * 2u;
    *reinterpret_cast<__nv_bfloat16*>(img + T.off_Ws + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(img + T.off_Ws + T.szWs + off) = l;
  }
  if (t < T.KC) bias[t] = (t < F) ? rb1[t] : (t - F < D.Fs ? sb1[t - F] : 0.f);
  if (t < T.Fp) bias[T.KC + t] = (t < F) ? rb2[t] + sb2[t] : 0.f;
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float silu_fast(float a) {
  // silu(a) = a / (1 + exp(-a)) with two MUFU ops (rel. error ~2e-7, far below the bf16x3 GEMM error)
  return a * rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * a));
}
__device__ __forceinline__ void silu_fast_d(float a, float* y, float* dy) {
  float s = rcp_approx(1.0f + ex2_approx(-1.4426950408889634f * a));
  *y = a * s;
  *dy = s * (1.0f + a * (1.0f - s));
}
// shared-memory carve-up common to the tensor-core edge kernels
struct TcSmem {
  unsigned char *W1hi, *W1lo, *W2hi, *W2lo, *Wshi, *Wslo, *A1hi, *A1lo, *A0hi, *A0lo, *A0shi, *A0slo;
  float *region;                 // start of the A1 region viewed as floats (staging tiles alias it)
  float *b1p, *b2p, *Ws1;
};
__device__ __forceinline__ TcSmem tc_carve(unsigned char* smem, const TcDims& T, const So3kDims& D) {
  TcSmem S;
  S.W1hi = smem + T.off_W1;
  S.W1lo = S.W1hi + T.szW1;
  S.W2hi = smem + T.off_W2;
  S.W2lo = S.W2hi + T.szW2;
  S.Wshi = smem + T.off_Ws;
  S.Wslo = S.Wshi + T.szWs;
  S.A1hi = smem + T.off_A1;
  S.A1lo = S.A1hi + T.szA1;
  // radial-basis operand images at the start of the region (dead once the GEMM1s have completed)
  S.A0hi = S.A1hi;
  S.A0lo = S.A0hi + T.szA0;
  S.A0shi = smem + T.off_A0s;
  S.A0slo = S.A0shi + T.szA0s;
  S.region = reinterpret_cast<float*>(smem + T.off_A1);
  S.b1p = reinterpret_cast<float*>(smem + T.off_small);
  S.b2p = S.b1p + T.KC;
  S.Ws1 = S.b2p + T.Fp;
  return S;
}

__device__ __forceinline__ void tc_setup(unsigned char* smem, const TcDims& T, const So3kDims& D, const TcSmem& S,
                                         const unsigned char* __restrict__ img, const float* __restrict__ bias_g,
                                         const float* __restrict__ Ws1_g) {
  const uint4* src = reinterpret_cast<const uint4*>(img);
  uint4* dst = reinterpret_cast<uint4*>(smem);
  const int n16 = (int)(T.off_A0s / 16);
  for (int k = threadIdx.x; k < n16; k += NT) dst[k] = src[k];
  for (int k = threadIdx.x; k < T.KC + T.Fp; k += NT) S.b1p[k] = bias_g[k];
  for (int k = threadIdx.x; k < D.nl * D.Fs; k += NT) S.Ws1[k] = Ws1_g[k];
}

// one GEMM: D[128 x N] = A[128 x 16*ksteps] @ B^T, 3 MMAs per K step (hi*hi + lo*hi + hi*lo); single thread
__device__ __forceinline__ void tc_issue_gemm(uint32_t tD, const unsigned char* Ahi, const unsigned char* Alo, uint32_t sboA,
                                              const unsigned char* Bhi, const unsigned char* Blo, uint32_t sboB, int ksteps,
                                              uint32_t idesc) {
  uint32_t acc = 0;
  const uint32_t ah = tc::smem_u32(Ahi), al = tc::smem_u32(Alo), bh = tc::smem_u32(Bhi), bl = tc::smem_u32(Blo);
  for (int ks = 0; ks < ksteps; ++ks) {
    const uint32_t ko = (uint32_t)ks * 256u;
    const uint64_t a_hi = tc::make_desc(ah + ko, 128u, sboA), a_lo = tc::make_desc(al + ko, 128u, sboA);
    const uint64_t b_hi = tc::make_desc(bh + ko, 128u, sboB), b_lo = tc::make_desc(bl + ko, 128u, sboB);
    tc::mma_bf16(tD, a_hi, b_hi, idesc, acc);
    tc::mma_bf16(tD, a_lo, b_hi, idesc, 1);
    tc::mma_bf16(tD, a_hi, b_lo, idesc, 1);
    acc = 1;
  }
}

// optional phase timing (debug): cycles between phase boundaries accumulated by thread 0 of block 0
__device__ unsigned long long g_tc_phase[3][16];
#define TC_PHASE(kern, idx)                                                        \
  if (TC_TIMING && blockIdx.x == 0 && threadIdx.x == 0) {                          \
    const long long now_ = clock64();                                              \
    atomicAdd(&g_tc_phase[kern][idx], (unsigned long long)(now_ - t_prev_));       \
    t_prev_ = now_;                                                                \
  }
#ifndef TC_TIMING
#define TC_TIMING 0
#endif

#define TC_WAIT_OR_BAIL(bar, par)                              \
  {                                                            \
    bool ok_ = tc::mbar_wait(bar, par);                        \
    if (__syncthreads_or(!ok_)) {                              \
      if (threadIdx.x == 0 && status) atomicOr(status, 8);     \
      break;                                                   \
    }                                                          \
    tc::fence_after_sync();                                    \
  }

// Per-pair record (k_pair_records, once per layer): [d, g_0 .. g_{nl-1}, 0 ...] = distance and the l0 contractions of
// (chi_j - chi_i) -- everything the filter MLP needs about a pair, so the tensor-core kernels read one contiguous
// 32-byte record per row (prefetched one tile ahead) instead of chasing row/col -> chi rows.
constexpr int REC = 8;                                   // floats per record (nl <= 7)
__global__ void k_pair_records(So3kDims D, int pair_cap, const int* __restrict__ n_canon, const int* __restrict__ canon,
                               const int* __restrict__ row, const int* __restrict__ col, const float4* __restrict__ geom,
                               const float* __restrict__ chi, float* __restrict__ rec) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int nv = (*n_canon < pair_cap) ? *n_canon : pair_cap;
  if (c >= nv) return;
  const int e = canon[c];
  const int i = row[e], j = col[e];
  float v[REC];
#pragma unroll
  for (int l = 0; l < REC; ++l) v[l] = 0.f;
  v[0] = geom[e].w;
  for (int k = 0; k < D.M; ++k) {
    const float cd = chi[(size_t)j * D.M + k] - chi[(size_t)i * D.M + k];
    const float cc = cd * cd * D.m_cg[k];
    const int sg = D.m_seg[k];
#pragma unroll
    for (int l = 0; l < REC - 1; ++l) v[1 + l] += (l == sg) ? cc : 0.f;
  }
  float4* dst = reinterpret_cast<float4*>(rec + (size_t)c * REC);
  dst[0] = make_float4(v[0], v[1], v[2], v[3]);
  dst[1] = make_float4(v[4], v[5], v[6], v[7]);
}

struct TileMeta {
  int e;                        // canonical sorted edge of this row's pair (loaded only when `canon` is given)
  float d;
  float gl[SO3K_MAXDEG];        // l0 contraction of the chi difference (only meaningful for column quarter 0)
};
__device__ __forceinline__ TileMeta tc_prefetch_meta(int r, int qq, int e0, int nv, bool comp, const float* __restrict__ rec,
                                                     const int* __restrict__ canon = nullptr) {
  TileMeta m;
  m.e = 0; m.d = 0.f;
#pragma unroll
  for (int l = 0; l < SO3K_MAXDEG; ++l) m.gl[l] = 0.f;
  const int c = e0 + r;
  if (comp && c < nv) {
    if (qq == 0) {
      const float4 a = *reinterpret_cast<const float4*>(rec + (size_t)c * REC);
      const float4 b = *reinterpret_cast<const float4*>(rec + (size_t)c * REC + 4);
      m.d = a.x;
      m.gl[0] = a.y; m.gl[1] = a.z; m.gl[2] = a.w; m.gl[3] = b.x; m.gl[4] = b.y; m.gl[5] = b.z; m.gl[6] = b.w;
      if (canon) m.e = canon[c];
    } else {
      m.d = rec[(size_t)c * REC];
    }
  }
  return m;
}

// S0: publish the tile's metadata to shared memory and write the radial-basis operand image(s).
// WITH_D additionally writes the image of d rbf / d d.
template <bool WITH_D>
__device__ __forceinline__ void tc_tile_s0(const So3kDims& D, const TcDims& T, const TcSmem& S, int r, int qq, int ne,
                                           const TileMeta& m, unsigned char* A0dhi, unsigned char* A0dlo) {
  const bool valid = r < ne;
  const float d = m.d;
  if (qq == 0) {
    // operand row of the spherical first layer: [g_0 .. g_{nl-1}, 0 ...] (K = 16)
    float g0[8], g1[8];
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      g0[l] = (valid && l < D.nl) ? m.gl[l] : 0.f;
      g1[l] = 0.f;
    }
    const uint32_t off = (uint32_t)(r >> 3) * SBO_S + (uint32_t)(r & 7) * 16u;
    tc::store_chunk8(S.A0shi, S.A0slo, off, g0);
    tc::store_chunk8(S.A0shi, S.A0slo, off + 128u, g1);
  }
  const float ed = __expf(-d);
  for (int c = qq; c < T.K0 / 8; c += 4) {     // this thread writes K chunks qq, qq+4, ... of its row
    float v[8], dv[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float t = ed - (D.mu0 + (float)(c * 8 + q) * D.dmu);
      const float rb = valid ? __expf(-D.gamma * t * t) : 0.f;
      v[q] = rb;
      if (WITH_D) dv[q] = rb * 2.0f * D.gamma * t * ed;
    }
    tc::store_chunk8(S.A0hi, S.A0lo, tc::canon_off(r, c, T.sboA0), v);
    if (WITH_D) tc::store_chunk8(A0dhi, A0dlo, tc::canon_off(r, c, T.sboA0), dv);
  }
}

// first layer of both MLP branches on the tensor cores: D1[:, 0:Fp] = rbf @ W1 (radial), then D1[:, F:F+NS] = g @ Ws1
// (spherical) written OVER the radial padding columns, so that TMEM column k == hidden unit k == K index of GEMM2
__device__ __forceinline__ void tc_issue_gemm1(const TcDims& T, const So3kDims& D, const TcSmem& S, uint32_t tD1,
                                               const unsigned char* A0hi, const unsigned char* A0lo, bool with_sph) {
  tc_issue_gemm(tD1, A0hi, A0lo, T.sboA0, S.W1hi, S.W1lo, T.sboW1, T.K0 / 16, tc::make_idesc_bf16(TM, T.Fp, 0, 0));
  if (with_sph)
    tc_issue_gemm(tD1 + (uint32_t)D.F, S.A0shi, S.A0slo, SBO_S, S.Wshi, S.Wslo, SBO_S, 1, tc::make_idesc_bf16(TM, T.NS, 0, 0));
}

// epilogue 1: hidden activations of both MLP branches -> A1 operand image (K index: [h1 (F) | h_s (Fs) | 0])
__device__ __forceinline__ void tc_epilogue1(const So3kDims& D, const TcDims& T, const TcSmem& S, uint32_t tD1,
                                             uint32_t lane_sel, int r, int qq) {
  const int nh = D.F + D.Fs;                          // hidden units
  const int nchunk = T.KC / 8;
  const int cbeg = (nchunk * qq) / 4, cend = (nchunk * (qq + 1)) / 4;
  for (int cb = cbeg; cb < cend; cb += 2) {           // 2 TMEM loads in flight
    float v[2][8];
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (cb + u < cend) tc::tmem_ld8(tD1 + lane_sel + (uint32_t)((cb + u) * 8), v[u]);
    tc::tmem_ld_wait();
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (cb + u < cend) {
        const int c0 = (cb + u) * 8;
        if (c0 >= nh) {                                // zero padding of K (warp-uniform)
#pragma unroll
          for (int q = 0; q < 8; ++q) v[u][q] = 0.f;
        } else {
          const float4 ba = *reinterpret_cast<const float4*>(S.b1p + c0);
          const float4 bb = *reinterpret_cast<const float4*>(S.b1p + c0 + 4);
          v[u][0] = silu_fast(v[u][0] + ba.x); v[u][1] = silu_fast(v[u][1] + ba.y);
          v[u][2] = silu_fast(v[u][2] + ba.z); v[u][3] = silu_fast(v[u][3] + ba.w);
          v[u][4] = silu_fast(v[u][4] + bb.x); v[u][5] = silu_fast(v[u][5] + bb.y);
          v[u][6] = silu_fast(v[u][6] + bb.z); v[u][7] = silu_fast(v[u][7] + bb.w);
          if (c0 + 8 > nh) {                           // last partial chunk: units beyond F + Fs are padding
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (c0 + q >= nh) v[u][q] = 0.f;
          }
        }
        tc::store_chunk8(S.A1hi, S.A1lo, tc::canon_off(r, c0 >> 3, T.sboA1), v[u]);
      }
    }
  }
}

// Filter kernel: w[c][f] of one filter block for every canonical pair c (the filter is symmetric in e <-> rev e,
// so it is computed once per undirected pair).  Pure dense work on contiguous tiles of pairs:
//   S0 (radial basis image) -> GEMM1 -> epilogue 1 (both hidden layers -> image) -> GEMM2 -> + bias -> HBM.
// Everything that gathers or reduces over neighbours (attention coefficients, aggregation, the q/k gradients)
// lives in the streaming kernels of so3k_agg.cu, which read w back through the pair index of an edge.
__global__ void __launch_bounds__(NT, 1)
k_filter_tc(So3kDims D, TcDims T, int n_tiles, int pair_cap, const int* __restrict__ n_canon,
            const float* __restrict__ rec, const unsigned char* __restrict__ img,
            const float* __restrict__ bias_g, const float* __restrict__ Ws1_g, float* __restrict__ wst,
            int* __restrict__ status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool comp = warp < NW;
  const int r = 32 * (warp & 3) + lane;      // row of the tile == TMEM lane
  const int qq = comp ? (warp >> 2) : 0;     // column quarter handled by this thread
  const int F = D.F;
  const TcSmem S = tc_carve(smem, T, D);

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init(&bars[1], 1);
    tc::mbar_fence_init();
  }
  tc_setup(smem, T, D, S, img, bias_g, Ws1_g);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tD1 = tmem, tD2 = tmem + 256;
  const uint32_t lane_sel = (uint32_t)(32 * (warp & 3)) << 16;
  const uint32_t idesc = tc::make_idesc_bf16(TM, T.Fp, 0, 0);
  const int nv = (*n_canon < pair_cap) ? *n_canon : pair_cap;
  const int nchunk_d = T.Fp / 8;
  const int cbeg = (nchunk_d * qq) / 4, cend = (nchunk_d * (qq + 1)) / 4;    // this thread's D2 chunks

  uint32_t it = 0;
  // epilogue-2 staging tile [TM][F] fp32 at the END of the operand region (the radial-basis images of the next tile
  // live at its start), drained to HBM by one bulk copy per tile
  const uint32_t stage_off = 2 * T.szA1 - (uint32_t)TM * F * 4u;         // (KC >= F: the tile always fits)
  float* stage = reinterpret_cast<float*>(smem + T.off_A1 + stage_off);
  const bool stage_hits_a0 = stage_off < 2 * T.szA0;                       // small F: S0 itself would overwrite the tile
  TileMeta meta = tc_prefetch_meta(r, qq, blockIdx.x * TM, nv, comp, rec);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int e0 = tile * TM;
    if (e0 >= nv) break;                     // block-uniform
    const int ne = (nv - e0) < TM ? (nv - e0) : TM;
    const uint32_t par = it & 1u;
    long long t_prev_ = TC_TIMING ? clock64() : 0;

    // the previous tile's bulk store must have finished READING the staging tile before it is overwritten (by
    // epilogue 1; by S0 already when the radial-basis images overlap it)
    if (stage_hits_a0) {
      if (tid == ISSUER) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncthreads();
    }
    if (comp) tc_tile_s0<false>(D, T, S, r, qq, ne, meta, nullptr, nullptr);
    if (!stage_hits_a0 && tid == ISSUER) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    TC_PHASE(0, 0);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    TC_PHASE(0, 1);
    if (tid == ISSUER) {
      tc::fence_after_sync();
      tc_issue_gemm1(T, D, S, tD1, S.A0hi, S.A0lo, true);
      tc::commit(&bars[0]);
    }
    // records of this CTA's next tile (consumed by its S0)
    const int e0_next = (tile + (int)gridDim.x) * TM;
    TileMeta meta_nx = tc_prefetch_meta(r, qq, e0_next, nv, comp, rec);
    TC_PHASE(0, 2);
    TC_WAIT_OR_BAIL(&bars[0], par);
    TC_PHASE(0, 3);
    if (comp) tc_epilogue1(D, T, S, tD1, lane_sel, r, qq);
    TC_PHASE(0, 4);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    TC_PHASE(0, 5);
    if (tid == ISSUER) {
      tc::fence_after_sync();
      tc_issue_gemm(tD2, S.A1hi, S.A1lo, T.sboA1, S.W2hi, S.W2lo, T.sboW2, T.KC / 16, idesc);
      tc::commit(&bars[1]);
    }
    meta = meta_nx;
    TC_PHASE(0, 6);
    TC_WAIT_OR_BAIL(&bars[1], par);
    TC_PHASE(0, 7);
    // ---- epilogue 2: w = D2 + b2 -> staging tile (row stride F: F mod 32 == 4 keeps the 16-byte row-per-lane stores
    //      conflict-free) -> one bulk copy to the tile's contiguous block of pair rows ----------------------------------
    if (comp) {
      float* srow = stage + (size_t)r * F;
      for (int cb = cbeg; cb < cend; cb += 2) {
        float v[2][8];
#pragma unroll
        for (int u = 0; u < 2; ++u)
          if (cb + u < cend) tc::tmem_ld8(tD2 + lane_sel + (uint32_t)((cb + u) * 8), v[u]);
        tc::tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const int c0 = (cb + u) * 8;
          if (cb + u < cend && c0 < F) {
            if (c0 + 8 <= F) {
              const float4 ba = *reinterpret_cast<const float4*>(S.b2p + c0);
              const float4 bb = *reinterpret_cast<const float4*>(S.b2p + c0 + 4);
              *reinterpret_cast<float4*>(srow + c0) = make_float4(v[u][0] + ba.x, v[u][1] + ba.y, v[u][2] + ba.z, v[u][3] + ba.w);
              *reinterpret_cast<float4*>(srow + c0 + 4) =
                  make_float4(v[u][4] + bb.x, v[u][5] + bb.y, v[u][6] + bb.z, v[u][7] + bb.w);
            } else {
#pragma unroll
              for (int q = 0; q < 8; ++q)
                if (c0 + q < F) srow[c0 + q] = v[u][q] + S.b2p[c0 + q];
            }
          }
        }
      }
    }
    TC_PHASE(0, 8);
    tc::fence_async_smem();                  // staging tile (generic proxy) -> visible to the bulk-copy engine
    tc::fence_before_sync();
    __syncthreads();
    if (tid == ISSUER) {
      const uint32_t bytes = (uint32_t)ne * (uint32_t)F * 4u;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(wst + (size_t)e0 * F),
                   "r"(tc::smem_u32(stage)), "r"(bytes)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
    TC_PHASE(0, 9);
    if (TC_TIMING && blockIdx.x == 0 && tid == 0) atomicAdd(&g_tc_phase[0][15], 1ull);
  }
  if (tid == ISSUER) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// Backward part 2: radial cotangent and l0-contraction cotangent of one filter block.
// Rows are canonical pairs; the pair's filter cotangent wbar[c][f] comes from k_qk_bwd_stream as a contiguous tile.
//   hbar = wbar @ W2c^T                                                 (GEMM3: the W2c image consumed MN-major)
//   dbar_rad[e] += sum_c hbar[c] silu'(a1[c]) (rbf'(d) @ W1)[c]         (a1 and the forward-mode term from 2 GEMM1s)
//   gbar[e][l]  += sum_c hbar_s[c] silu'(a_s[c]) Ws1[l][c]
__global__ void __launch_bounds__(NT, 1)
k_edge_bwd2_tc(So3kDims D, TcDims T, int Gst, int n_tiles, int pair_cap, const int* __restrict__ n_canon,
               const int* __restrict__ canon, const float* __restrict__ rec, const unsigned char* __restrict__ img,
               const float* __restrict__ bias_g, const float* __restrict__ Ws1_g, const float* __restrict__ wbar,
               float* __restrict__ gbar, float* __restrict__ ebar, int* __restrict__ status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool comp = warp < NW;
  const int r = 32 * (warp & 3) + lane;
  const int qq = comp ? (warp >> 2) : 0;
  const int F = D.F, Fs = D.Fs;
  TcSmem S = tc_carve(smem, T, D);
  // four radial-basis images packed at the start of the region: rbf hi | rbf lo | rbf' hi | rbf' lo
  S.A0hi = reinterpret_cast<unsigned char*>(S.region);
  S.A0lo = S.A0hi + T.szA0;
  unsigned char* A0dhi = S.A0hi + 2 * T.szA0;
  unsigned char* A0dlo = S.A0hi + 3 * T.szA0;
  const uint32_t sboWb = (uint32_t)(T.Fp / 8) * 128u;
  const uint32_t szWb = (uint32_t)(TM / 8) * sboWb;
  unsigned char* Wbhi = reinterpret_cast<unsigned char*>(S.region);   // aliases the rbf images: dead once the GEMM1s are done
  unsigned char* Wblo = Wbhi + szWb;
  float* s_red = S.region;                      // [4 quarters][8][TM]: cross-quarter reduction of (dbar, gbar_l); aliases
                                                // the wbar image, dead once GEMM3 has completed

  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init(&bars[1], 1);
    tc::mbar_fence_init();
  }
  tc_setup(smem, T, D, S, img, bias_g, Ws1_g);
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  // TMEM columns: a (both branches' pre-activations, F + NS <= 192) | forward-mode radial derivative (Fp) | hbar (KC)
  const uint32_t tD1 = tmem, tD1p = tmem + 192u, tD3 = tmem + 192u + (uint32_t)T.Fp;
  const uint32_t lane_sel = (uint32_t)(32 * (warp & 3)) << 16;
  const uint32_t idesc1 = tc::make_idesc_bf16(TM, T.Fp, 0, 0);
  const uint32_t idesc3 = tc::make_idesc_bf16(TM, T.KC, 0, 1);     // B = W2c image read MN-major
  const int nv = (*n_canon < pair_cap) ? *n_canon : pair_cap;   // rows = canonical edges (undirected pairs): the filter
  const int nchunkF = T.Fp / 8;                 // and everything downstream of it are identical for e and rev e, and
  const int nchunkK = T.KC / 8;                 // only dbar_e + dbar_rev / gbar_e + gbar_rev enter the forces
  // this thread's hidden-column chunks: a quarter of the radial range plus every 4th chunk beyond it (the spherical
  // hidden units cost more per column, so they are dealt round-robin)
  const int kcb = (nchunkF * qq) / 4, kce = (nchunkF * (qq + 1)) / 4;
  const int nmine = comp ? (kce - kcb) + ((nchunkK - nchunkF - qq + 3) / 4) : 0;

  uint32_t it = 0;
  TileMeta meta = tc_prefetch_meta(r, qq, blockIdx.x * TM, nv, comp, rec, canon);
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int e0 = tile * TM;
    if (e0 >= nv) break;
    const int ne = (nv - e0) < TM ? (nv - e0) : TM;
    const uint32_t par = it & 1u;
    const int my_e = meta.e;                      // valid for r < ne
    long long t_prev_ = TC_TIMING ? clock64() : 0;

    if (comp) tc_tile_s0<true>(D, T, S, r, qq, ne, meta, A0dhi, A0dlo);
    TC_PHASE(2, 0);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    TC_PHASE(2, 1);
    if (tid == ISSUER) {
      tc::fence_after_sync();
      tc_issue_gemm1(T, D, S, tD1, S.A0hi, S.A0lo, true);
      tc_issue_gemm1(T, D, S, tD1p, A0dhi, A0dlo, false);
      tc::commit(&bars[0]);
    }
    // records of this CTA's next tile (consumed by its S0)
    const int e0_next = (tile + (int)gridDim.x) * TM;
    TileMeta meta_nx = tc_prefetch_meta(r, qq, e0_next, nv, comp, rec, canon);
    // ---- wbar image from the tile's contiguous block of pair rows: work item = (8 rows, 4 chunks of 8 columns),
    //      lane -> (row = lane / 4, chunk = lane % 4); the loads of all of a warp's items are issued before the GEMM1 wait
    const int ncg = (nchunkF + 3) / 4;
    const int nitems = 16 * ncg;
    constexpr int MAXIT = 6;                      // items per warp: 16 * ceil(Fp / 32) / 16 <= 5 for Fp <= 144
    float4 wa[MAXIT], wb2[MAXIT];
#pragma unroll
    for (int v = 0; v < MAXIT; ++v) {
      const int item = warp + v * NW;
      wa[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      wb2[v] = wa[v];
      if (comp && item < nitems) {
        const int rg = item / ncg, cg = item - rg * ncg;
        const int rr = rg * 8 + (lane >> 2), c = cg * 4 + (lane & 3);
        const int f0 = c * 8;
        if (c < nchunkF && f0 < F && rr < ne) {
          const float* src = wbar + (size_t)(e0 + rr) * F + f0;
          wa[v] = __ldcs(reinterpret_cast<const float4*>(src));
          if (f0 + 4 < F) wb2[v] = __ldcs(reinterpret_cast<const float4*>(src + 4));
        }
      }
    }
    TC_PHASE(2, 2);
    TC_WAIT_OR_BAIL(&bars[0], par);
    TC_PHASE(2, 3);
#pragma unroll
    for (int v = 0; v < MAXIT; ++v) {
      const int item = warp + v * NW;
      if (comp && item < nitems) {
        const int rg = item / ncg, cg = item - rg * ncg;
        const int rr = rg * 8 + (lane >> 2), c = cg * 4 + (lane & 3);
        if (c < nchunkF) {
          const float wv[8] = {wa[v].x, wa[v].y, wa[v].z, wa[v].w, wb2[v].x, wb2[v].y, wb2[v].z, wb2[v].w};
          tc::store_chunk8(Wbhi, Wblo, tc::canon_off(rr, c, sboWb), wv);
        }
      }
    }
    TC_PHASE(2, 4);
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    TC_PHASE(2, 5);
    // ---- GEMM3: hbar[128 x KC] = wbar[128 x Fp] @ W2c^T --------------------------------------------------------------
    if (tid == ISSUER) {
      tc::fence_after_sync();
      uint32_t acc = 0;
      const uint32_t ah = tc::smem_u32(Wbhi), al = tc::smem_u32(Wblo), bh = tc::smem_u32(S.W2hi), bl = tc::smem_u32(S.W2lo);
      for (int ks = 0; ks < T.Fp / 16; ++ks) {
        const uint32_t ao = (uint32_t)ks * 256u;
        const uint32_t bo = (uint32_t)ks * 2u * T.sboW2;
        const uint64_t a_hi = tc::make_desc(ah + ao, 128u, sboWb), a_lo = tc::make_desc(al + ao, 128u, sboWb);
        const uint64_t b_hi = tc::make_desc(bh + bo, T.sboW2, 128u), b_lo = tc::make_desc(bl + bo, T.sboW2, 128u);
        tc::mma_bf16(tD3, a_hi, b_hi, idesc3, acc);
        tc::mma_bf16(tD3, a_lo, b_hi, idesc3, 1);
        tc::mma_bf16(tD3, a_hi, b_lo, idesc3, 1);
        acc = 1;
      }
      tc::commit(&bars[1]);
    }
    float gb[SO3K_MAXDEG];
#pragma unroll
    for (int l = 0; l < SO3K_MAXDEG; ++l) gb[l] = 0.f;
    // overlapping GEMM3: pull the next tile's filter-cotangent block (contiguous, TM * F floats) into L2
    if (e0_next < nv) {
      const int nrow = (nv - e0_next) < TM ? (nv - e0_next) : TM;
      const char* pb = reinterpret_cast<const char*>(wbar + (size_t)e0_next * F);
      const int nline = (nrow * F * 4 + 127) / 128;
      for (int k = tid; k < nline; k += NT) asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + (size_t)k * 128));
    }
    meta = meta_nx;
    TC_PHASE(2, 6);
    TC_WAIT_OR_BAIL(&bars[1], par);
    TC_PHASE(2, 7);
    // ---- epilogue: this thread's quarter of the KC hidden columns of its row ------------------------------------------------
    {
      float dbar = 0.f;
      int ci = 0;
      // radial chunks two at a time: six TMEM loads in flight, sixteen independent silu' chains
      for (; comp && ci + 1 < kce - kcb && (kcb + ci + 2) * 8 <= F; ci += 2) {
        const int c0 = (kcb + ci) * 8;
        float hb[2][8], a1[2][8], tp[2][8];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          tc::tmem_ld8(tD3 + lane_sel + (uint32_t)(c0 + 8 * u), hb[u]);
          tc::tmem_ld8(tD1 + lane_sel + (uint32_t)(c0 + 8 * u), a1[u]);
          tc::tmem_ld8(tD1p + lane_sel + (uint32_t)(c0 + 8 * u), tp[u]);
        }
        tc::tmem_ld_wait();
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float4 ba = *reinterpret_cast<const float4*>(S.b1p + c0 + 8 * u);
          const float4 bb = *reinterpret_cast<const float4*>(S.b1p + c0 + 8 * u + 4);
          const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float y, dy;
            silu_fast_d(a1[u][q] + bv[q], &y, &dy);
            dbar = fmaf(hb[u][q] * dy, tp[u][q], dbar);
          }
        }
      }
      for (; ci < nmine; ++ci) {
        const int cc = (ci < kce - kcb) ? kcb + ci : nchunkF + qq + 4 * (ci - (kce - kcb));
        const int c0 = cc * 8;
        if (c0 >= F + Fs) break;                  // warp-uniform: only zero padding left
        float hb[8], a1[8], tp[8];
        tc::tmem_ld8(tD3 + lane_sel + (uint32_t)c0, hb);
        tc::tmem_ld8(tD1 + lane_sel + (uint32_t)c0, a1);                 // pre-activations of both branches
        if (c0 < F) tc::tmem_ld8(tD1p + lane_sel + (uint32_t)c0, tp);    // forward-mode radial derivative
        tc::tmem_ld_wait();
        if (c0 + 8 <= F) {                        // radial chunk: no per-element control flow
          const float4 ba = *reinterpret_cast<const float4*>(S.b1p + c0);
          const float4 bb = *reinterpret_cast<const float4*>(S.b1p + c0 + 4);
          const float bv[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            float y, dy;
            silu_fast_d(a1[q] + bv[q], &y, &dy);
            dbar = fmaf(hb[q] * dy, tp[q], dbar);
          }
        } else {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int k = c0 + q;
            if (k < F + Fs) {
              float y, dy;
              silu_fast_d(a1[q] + S.b1p[k], &y, &dy);
              const float as = hb[q] * dy;
              if (k < F) {
                dbar = fmaf(as, tp[q], dbar);
              } else {
                const int c = k - F;
#pragma unroll
                for (int l = 0; l < 7; ++l)
                  if (l < D.nl) gb[l] = fmaf(as, S.Ws1[l * Fs + c], gb[l]);
              }
            }
          }
        }
      }
      TC_PHASE(2, 8);
      // cross-quarter combination through shared memory (one 32-byte slot per thread)
      if (comp) {
        float* slot = s_red + (size_t)qq * 8 * TM + r;         // lanes = consecutive rows: conflict-free
        slot[0] = dbar;
#pragma unroll
        for (int l = 0; l < 7; ++l)
          if (l < D.nl) slot[(1 + l) * TM] = gb[l];
      }
      tc::fence_before_sync();
      __syncthreads();
      TC_PHASE(2, 9);
      if (comp && qq == 0 && r < ne) {
        // the pair's totals go to the canonical edge; its reverse keeps the zeros of the memset (only the sums
        // ebar[e] + ebar[rev e], gbar[e] + gbar[rev e] enter chi-bar and the forces).  Fire-and-forget reductions: both
        // blocks add into the same slots.
        const size_t e = (size_t)my_e;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          if (k <= D.nl) {
            const float tot = s_red[(0 * 8 + k) * TM + r] + s_red[(1 * 8 + k) * TM + r] + s_red[(2 * 8 + k) * TM + r] +
                              s_red[(3 * 8 + k) * TM + r];
            if (k == 0) atomicAdd(&ebar[2 * e], tot);
            else atomicAdd(&gbar[e * Gst + k - 1], tot);
          }
        }
      }
    }
    __syncthreads();
    TC_PHASE(2, 10);
    if (TC_TIMING && blockIdx.x == 0 && tid == 0) atomicAdd(&g_tc_phase[2][15], 1ull);
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

bool so3k_edge_tc_supported(const So3kDims& D) {
  if (D.F % 4 != 0 || D.K % 16 != 0) return false;
  TcDims T = make_tc_dims(D);
  if (T.Fp > 192 || T.Fp % 16 != 0) return false;               // <= 6 image items per warp in backward part 2
  if (D.H > 8 || D.nl > 7) return false;                        // pair records hold 7 contractions
  if (D.F + T.NS > 192 || 192 + T.Fp + T.KC > 512) return false;   // TMEM columns (backward part 2: a | t' | hbar)
  return T.total + 1024 <= 227u * 1024u;
}

size_t so3k_edge_tc_image_bytes(const So3kDims& D) {
  TcDims T = make_tc_dims(D);
  return ((size_t)T.off_A0s + 255) / 256 * 256 + sizeof(float) * (T.KC + T.Fp) + 256;
}

void so3k_edge_tc_build(const So3kDims& D, const float* p, const BlockParams& bp, unsigned char* dst, TcBlockW* out,
                        const EdgeBlockW& simt, cudaStream_t st) {
  TcDims T = make_tc_dims(D);
  size_t img_bytes = ((size_t)T.off_A0s + 255) / 256 * 256;
  float* bias = reinterpret_cast<float*>(dst + img_bytes);
  int tot = T.Fp * T.KC;
  so3k_count_launch();
  k_build_images<<<so3k_cdiv(tot, 256), 256, 0, st>>>(D, T, p + bp.rad_W1, p + bp.rad_b1, p + bp.rad_W2, p + bp.rad_b2,
                                                       p + bp.sph_W1, p + bp.sph_b1, p + bp.sph_W2, p + bp.sph_b2, dst,
                                                       bias);
  out->img = dst;
  out->bias = bias;
  out->Ws1 = simt.Ws1;
  out->bs1 = simt.bs1;
}

static int tc_grid(int n_tiles, const TcDims& T) {
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    (void)T;
  }
  return n_tiles < n_sm ? n_tiles : n_sm;
}

// filter of one block for every canonical pair -> wst[pair][F]
// per-pair records of the current layer (distance + l0 contractions of the chi differences) -> rec[pair capacity][8]
void so3k_launch_pair_records(const So3kDims& D, const Graph& g, const float* chi, float* rec, cudaStream_t st) {
  if (g.Pt <= 0) return;
  const int pair_cap = so3k_pair_capacity(g.Pt);
  so3k_count_launch();
  k_pair_records<<<so3k_cdiv(pair_cap, 256), 256, 0, st>>>(D, pair_cap, g.n_canon, g.canon, g.row, g.col, g.geom, chi, rec);
}

void so3k_launch_filter_tc(const So3kDims& D, const Graph& g, const TcBlockW& W, const float* rec, float* wst, int* status,
                           cudaStream_t st) {
  if (g.Pt <= 0) return;
  TcDims T = make_tc_dims(D);
  const int pair_cap = so3k_pair_capacity(g.Pt);
  int n_tiles = so3k_cdiv(pair_cap, TM);
  int grid = tc_grid(n_tiles, T);
  cudaFuncSetAttribute(k_filter_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T.total + 1024);
  so3k_count_launch();
  k_filter_tc<<<grid, NT, T.total + 1024, st>>>(D, T, n_tiles, pair_cap, g.n_canon, rec, W.img, W.bias, W.Ws1, wst, status);
}

// radial cotangent (ebar[e][0]) and l0-contraction cotangent (gbar) of both filter blocks, ADDED to zero-initialised
// arrays; wbar = [block][pair capacity][F] filter cotangents from k_qk_bwd_stream
void so3k_launch_edge_bwd2_tc(const So3kDims& D, const Graph& g, const TcBlockW& Wf, const TcBlockW& Wg, const float* rec,
                              const float* wbar, float* gbar, float* ebar, int* status, cudaStream_t st) {
  if (g.Pt <= 0) return;
  TcDims T = make_tc_dims(D);
  int Gst = (D.nl + 3) & ~3;
  const int pair_cap = so3k_pair_capacity(g.Pt);
  int n_tiles = so3k_cdiv(pair_cap, TM);
  int grid = tc_grid(n_tiles, T);
  cudaFuncSetAttribute(k_edge_bwd2_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T.total + 1024);
  for (int blk = 0; blk < 2; ++blk) {
    const TcBlockW& W = blk ? Wg : Wf;
    so3k_count_launch();
    k_edge_bwd2_tc<<<grid, NT, T.total + 1024, st>>>(D, T, Gst, n_tiles, pair_cap, g.n_canon, g.canon, rec, W.img, W.bias,
                                                      W.Ws1, wbar + (size_t)blk * pair_cap * D.F, gbar, ebar, status);
  }
}
extern "C" void so3k_tc_phase_dump(void) {
  unsigned long long h[3][16];
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h, g_tc_phase, sizeof(h));
  // phase names per kernel index (TC_PHASE(kernel, phase)): 0 = k_filter_tc, 2 = k_edge_bwd2_tc
  static const char* const names[3][13] = {
      {"records+images", "sync", "issue GEMM1", "wait", "epilogue1", "sync", "issue GEMM2", "wait", "epilogue2",
       "sync+bulk store", "-", "-", "-"},
      {"-", "-", "-", "-", "-", "-", "-", "-", "-", "-", "-", "-", "-"},
      {"rbf images", "sync", "issue GEMM1s + wbar loads", "wait", "wbar image", "sync", "issue GEMM3 + prefetch", "wait",
       "epilogue", "reduce+sync", "stores+sync", "-", "-"}};
  for (int k = 0; k < 3; ++k) {
    if (!h[k][15]) continue;
    double tot = 0;
    printf("tc phase timing kernel %d (%llu tiles of block 0):", k, h[k][15]);
    for (int p = 0; p < 13; ++p) {
      if (names[k][p][0] == '-' && !h[k][p]) continue;
      printf(" %s=%.0f", names[k][p], (double)h[k][p] / h[k][15]);
      tot += (double)h[k][p] / h[k][15];
    }
    printf(" | total=%.0f cycles/tile\n", tot);
  }
  memset(h, 0, sizeof(h));
  cudaMemcpyToSymbol(g_tc_phase, h, sizeof(h));
}
#else
extern "C" void so3k_tc_phase_dump(void) {}
bool so3k_edge_tc_supported(const So3kDims&) { return false; }
size_t so3k_edge_tc_image_bytes(const So3kDims&) { return 0; }
void so3k_edge_tc_build(const So3kDims&, const float*, const BlockParams&, unsigned char*, TcBlockW*, const EdgeBlockW&,
                        cudaStream_t) {}
void so3k_launch_pair_records(const So3kDims&, const Graph&, const float*, float*, cudaStream_t) {}
void so3k_launch_filter_tc(const So3kDims&, const Graph&, const TcBlockW&, const float*, float*, int*, cudaStream_t) {}
void so3k_launch_edge_bwd2_tc(const So3kDims&, const Graph&, const TcBlockW&, const TcBlockW&, const float*, const float*,
                              float*, float*, int*, cudaStream_t) {}
#endif
