// Per-edge filter MLPs + attention coefficients on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// Same math and same interface as k_edge_fwd (so3k_edge.cu):
//   chi_ij, l0 contraction     mlff/nn/layer/so3krates_layer.py:89-93 ; mlff/sph_ops/contract.py:181-188
//   RadialSphericalFilter      mlff/nn/layer/so3krates_layer.py:449-465 ; MLP mlff/nn/mlp/mlp.py:21-27
//   ConvAttentionCoefficients  mlff/nn/layer/so3krates_layer.py:568-586 ; masking/cutoff :500-501, :551-554
// Design (one filter block -- fb or gb -- per launch, persistent CTAs, 1 CTA / SM, 256 threads):
//   * the block's weights live in shared memory for the whole kernel as bf16 hi/lo operand images in the
//     canonical no-swizzle K-major layout (so3k_tc.cuh): W1 (N=Fp, K=n_rbf) and W2c (N=Fp, K=KC) where the
//     radial and the spherical second layers are stacked along K, so  w = [silu(a1) | silu(a_s)] @ W2c  is ONE GEMM;
//   * a tile = 128 consecutive receiver-sorted edges = the M dimension of one UMMA (M=128, cta_group::1);
//     thread t and thread t+128 own row t (TMEM lane t) and split its columns;
//   * GEMM1  D1[128 x Fp] = rbf @ W1         (3 MMAs per 16-wide K step: hi*hi + lo*hi + hi*lo, fp32 accumulate in TMEM)
//     epilogue 1: tcgen05.ld -> +bias -> silu -> bf16 hi/lo split -> written straight into the A-operand image of
//     GEMM2, together with the tiny spherical hidden layer (K = nl, CUDA cores);
//   * GEMM2  D2[128 x Fp] = A1 @ W2c ;  epilogue 2: tcgen05.ld -> +bias -> * (q_i k_j) -> per-head sums -> alpha.
//     q_i * k_j is gathered with full-line coalesced loads into the (by then dead) A1 region.
//   The F-wide filter never exists outside TMEM/registers.
// Algorithmic FLOPs per launch: 2 * P * (K*F + F*F + nl*F/4 + F*F/4); tensor work issued is 3x that
// on padded shapes (Fp x KC).
#include "so3k_internal.h"
#include "so3k_edge.h"
#ifndef SO3K_EMU
#include "so3k_tc.cuh"

namespace {

constexpr int NT = 256;
constexpr int TM = 128;        // edges per tile

struct TcDims {
  int Fp, KC, K0, QS;          // padded N, padded K of GEMM2, K of GEMM1, row stride (floats) of the qk tile
  uint32_t sboW1, sboW2, sboA0, sboA1;
  uint32_t szW1, szW2, szA0, szA1;   // bytes of ONE (hi or lo) image
  uint32_t off_W1, off_W2, off_A1, off_small, total;
};

__host__ __device__ inline TcDims make_tc_dims(const So3kDims& D) {
  TcDims t;
  t.Fp = (D.F + 15) & ~15;
  t.KC = (D.F + D.Fs + 15) & ~15;
  t.K0 = D.K;
  t.QS = ((D.F - 4 + 31) / 32) * 32 + 4;
  t.sboW1 = (uint32_t)(t.K0 / 8) * 128u;
  t.sboW2 = (uint32_t)(t.KC / 8) * 128u;
  t.sboA0 = t.sboW1;
  t.sboA1 = t.sboW2;
  t.szW1 = (uint32_t)(t.Fp / 8) * t.sboW1;
  t.szW2 = (uint32_t)(t.Fp / 8) * t.sboW2;
  t.szA0 = (uint32_t)(TM / 8) * t.sboA0;
  t.szA1 = (uint32_t)(TM / 8) * t.sboA1;
  t.off_W1 = 0;
  t.off_W2 = t.off_W1 + 2 * t.szW1;
  t.off_A1 = t.off_W2 + 2 * t.szW2;
  uint32_t a1_region = 2 * t.szA1;
  uint32_t qk_bytes = (uint32_t)TM * t.QS * 4u;
  if (qk_bytes > a1_region) a1_region = qk_bytes;
  t.off_small = t.off_A1 + ((a1_region + 127u) & ~127u);
  // small arrays: b1p, b2p (Fp each), Ws1 (nl*Fs), bs1 (Fs), s_phi (TM), s_g (TM*nl), s_part (TM*8*2), s_i, s_j (TM)
  uint32_t small = 4u * (2u * t.Fp + (uint32_t)D.nl * D.Fs + D.Fs + TM + TM * D.nl + TM * 16 + 2 * TM);
  t.total = t.off_small + ((small + 127u) & ~127u);
  return t;
}

// ---- weight images (built once at so3k_load_params) -------------------------------------------------------
// image layout in global memory == shared-memory layout: [W1 hi | W1 lo | W2 hi | W2 lo] then fp32 [b1p | b2p]
__global__ void k_build_images(So3kDims D, TcDims T, const float* __restrict__ rW1, const float* __restrict__ rb1,
                               const float* __restrict__ rW2, const float* __restrict__ rb2,
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
  if (t < T.Fp) {
    bias[t] = (t < F) ? rb1[t] : 0.f;
    bias[T.Fp + t] = (t < F) ? rb2[t] + sb2[t] : 0.f;
  }
}

__device__ __forceinline__ float silu_fast(float a) {
  // silu(a) = a / (1 + exp(-a)); ex2.approx + fast division (rel. error ~1e-7, far below the bf16x3 GEMM error)
  return __fdividef(a, 1.0f + __expf(-a));
}

__global__ void __launch_bounds__(NT, 1)
k_edge_fwd_tc(So3kDims D, TcDims T, int blk, int Ast, int n_tiles, const int* __restrict__ n_valid,
              const int* __restrict__ row, const int* __restrict__ col, const float4* __restrict__ geom,
              const float* __restrict__ chi, const float* __restrict__ proj, const unsigned char* __restrict__ img,
              const float* __restrict__ bias_g, const float* __restrict__ Ws1_g, const float* __restrict__ bs1_g,
              float* __restrict__ alpha, int* __restrict__ status) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t bars[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = 32 * (warp & 3) + lane;      // row of the tile / TMEM lane
  const int hh = warp >> 2;                  // column half handled by this thread
  const int F = D.F, nl = D.nl, Fs = D.Fs;
  const int heads = blk ? D.nl : D.H;
  const int dhd = F / heads;
  const int aoff = blk ? D.H : 0;
  const int qoff = blk ? F : 0;
  const int koff = 2 * F + (blk ? F : 0);
  const size_t F5 = 5 * (size_t)F;
  const float inv = 1.0f / sqrtf((float)dhd);

  unsigned char* W1hi = smem + T.off_W1;
  unsigned char* W1lo = W1hi + T.szW1;
  unsigned char* W2hi = smem + T.off_W2;
  unsigned char* W2lo = W2hi + T.szW2;
  unsigned char* A1hi = smem + T.off_A1;
  unsigned char* A1lo = A1hi + T.szA1;
  unsigned char* A0hi = A1hi;                // aliased: dead once GEMM1 has completed
  unsigned char* A0lo = A1lo;
  float* qk = reinterpret_cast<float*>(smem + T.off_A1);   // aliased: A1 is dead once GEMM2 has completed
  float* b1p = reinterpret_cast<float*>(smem + T.off_small);
  float* b2p = b1p + T.Fp;
  float* Ws1 = b2p + T.Fp;
  float* bs1 = Ws1 + nl * Fs;
  float* s_phi = bs1 + Fs;
  float* s_g = s_phi + TM;
  float* s_part = s_g + TM * nl;             // [TM][8][2]
  int* s_i = reinterpret_cast<int*>(s_part + TM * 16);
  int* s_j = s_i + TM;

  // ---- one-time setup: TMEM, barriers, weights -----------------------------------------------------------
  if (warp == 0) tc::tmem_alloc(&tmem_base_s, 512);
  if (tid == 0) {
    tc::mbar_init(&bars[0], 1);
    tc::mbar_init(&bars[1], 1);
    tc::mbar_fence_init();
  }
  {
    const uint4* src = reinterpret_cast<const uint4*>(img);
    uint4* dst = reinterpret_cast<uint4*>(smem);
    const int n16 = (int)((2 * T.szW1 + 2 * T.szW2) / 16);
    for (int k = tid; k < n16; k += NT) dst[k] = src[k];
    for (int k = tid; k < 2 * T.Fp; k += NT) b1p[k] = bias_g[k];
    for (int k = tid; k < nl * Fs; k += NT) Ws1[k] = Ws1_g[k];
    for (int k = tid; k < Fs; k += NT) bs1[k] = bs1_g[k];
  }
  tc::fence_async_smem();
  tc::fence_before_sync();
  __syncthreads();
  tc::fence_after_sync();
  const uint32_t tmem = tmem_base_s;
  const uint32_t tD1 = tmem, tD2 = tmem + 256;
  const uint32_t lane_sel = (uint32_t)(32 * (warp & 3)) << 16;
  const uint32_t idesc = tc::make_idesc_bf16(TM, T.Fp, 0, 0);
  const int nv = *n_valid;
  const int half_cols = T.Fp / 2;            // D columns per thread (multiple of 8)
  const int nchunk_d = T.Fp / 8;             // K chunks of A1 that come from D1 columns
  const int nchunk_all = T.KC / 8;

  uint32_t it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int e0 = tile * TM;
    if (e0 >= nv) break;                     // block-uniform
    const int ne = (nv - e0) < TM ? (nv - e0) : TM;
    const uint32_t par = it & 1u;

    // ---- S0: per-edge metadata, radial basis -> A0 image ---------------------------------------------------
    float ed = 1.0f;
    {
      int i = 0, j = 0;
      float d = 0.f, ph = 0.f;
      float gl[SO3K_MAXDEG];
      for (int l = 0; l < nl; ++l) gl[l] = 0.f;
      if (r < ne) {
        const int e = e0 + r;
        i = row[e];
        j = col[e];
        d = geom[e].w;
        ph = so3k_cutoff(d, D.r_cut);
        if (hh == 0) {
          for (int m = 0; m < D.M; ++m) {
            float c = chi[(size_t)j * D.M + m] - chi[(size_t)i * D.M + m];
            gl[D.m_seg[m]] += c * c * D.m_cg[m];
          }
        }
      }
      ed = expf(-d);
      if (hh == 0) {
        s_i[r] = i; s_j[r] = j; s_phi[r] = ph;
        for (int l = 0; l < nl; ++l) s_g[r * nl + l] = gl[l];
      }
      // rbf columns: this thread writes K chunks hh, hh+2, ... of its row
      for (int c = hh; c < T.K0 / 8; c += 2) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (r < ne) ? so3k_rbf(ed, c * 8 + q, D) : 0.f;
        tc::store_chunk8(A0hi, A0lo, tc::canon_off(r, c, T.sboA0), v);
      }
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    // ---- GEMM1 ------------------------------------------------------------------------------------------------
    if (tid == 0) {
      tc::fence_after_sync();
      uint32_t acc = 0;
      for (int ks = 0; ks < T.K0 / 16; ++ks) {
        const uint32_t ko = (uint32_t)ks * 256u;
        uint64_t a_hi = tc::make_desc(tc::smem_u32(A0hi) + ko, 128u, T.sboA0);
        uint64_t a_lo = tc::make_desc(tc::smem_u32(A0lo) + ko, 128u, T.sboA0);
        uint64_t b_hi = tc::make_desc(tc::smem_u32(W1hi) + ko, 128u, T.sboW1);
        uint64_t b_lo = tc::make_desc(tc::smem_u32(W1lo) + ko, 128u, T.sboW1);
        tc::mma_bf16(tD1, a_hi, b_hi, idesc, acc);
        tc::mma_bf16(tD1, a_lo, b_hi, idesc, 1);
        tc::mma_bf16(tD1, a_hi, b_lo, idesc, 1);
        acc = 1;
      }
      tc::commit(&bars[0]);
    }
    {
      bool ok = tc::mbar_wait(&bars[0], par);
      if (__syncthreads_or(!ok)) {
        if (tid == 0 && status) atomicOr(status, 8);
        break;
      }
    }
    tc::fence_after_sync();
    // ---- epilogue 1: hidden activations -> A1 image ---------------------------------------------------------------
    {
      float gl[SO3K_MAXDEG];
      for (int l = 0; l < nl; ++l) gl[l] = s_g[r * nl + l];
      auto hidden_s = [&](int c) -> float {          // spherical hidden unit c (< Fs)
        float a = bs1[c];
        for (int l = 0; l < nl; ++l) a = fmaf(gl[l], Ws1[l * Fs + c], a);
        return silu_fast(a);
      };
      for (int cc = 0; cc < half_cols / 8; ++cc) {
        const int c0 = hh * half_cols + cc * 8;       // first column of this chunk (== K index in A1)
        float v[8];
        tc::tmem_ld8(tD1 + lane_sel + (uint32_t)c0, v);
        tc::tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int k = c0 + q;
          float y;
          if (k < F) y = silu_fast(v[q] + b1p[k]);
          else if (k - F < Fs) y = hidden_s(k - F);
          else y = 0.f;
          v[q] = y;
        }
        tc::store_chunk8(A1hi, A1lo, tc::canon_off(r, c0 >> 3, T.sboA1), v);
      }
      // K chunks beyond the D1 columns: remaining spherical hidden units and zero padding
      for (int ch = nchunk_d + hh; ch < nchunk_all; ch += 2) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int k = ch * 8 + q;
          v[q] = (k - F < Fs) ? hidden_s(k - F) : 0.f;
        }
        tc::store_chunk8(A1hi, A1lo, tc::canon_off(r, ch, T.sboA1), v);
      }
    }
    tc::fence_async_smem();
    tc::fence_before_sync();
    __syncthreads();
    // ---- GEMM2 ------------------------------------------------------------------------------------------------
    if (tid == 0) {
      tc::fence_after_sync();
      uint32_t acc = 0;
      for (int ks = 0; ks < T.KC / 16; ++ks) {
        const uint32_t ko = (uint32_t)ks * 256u;
        uint64_t a_hi = tc::make_desc(tc::smem_u32(A1hi) + ko, 128u, T.sboA1);
        uint64_t a_lo = tc::make_desc(tc::smem_u32(A1lo) + ko, 128u, T.sboA1);
        uint64_t b_hi = tc::make_desc(tc::smem_u32(W2hi) + ko, 128u, T.sboW2);
        uint64_t b_lo = tc::make_desc(tc::smem_u32(W2lo) + ko, 128u, T.sboW2);
        tc::mma_bf16(tD2, a_hi, b_hi, idesc, acc);
        tc::mma_bf16(tD2, a_lo, b_hi, idesc, 1);
        tc::mma_bf16(tD2, a_hi, b_lo, idesc, 1);
        acc = 1;
      }
      tc::commit(&bars[1]);
    }
    {
      bool ok = tc::mbar_wait(&bars[1], par);
      if (__syncthreads_or(!ok)) {
        if (tid == 0 && status) atomicOr(status, 8);
        break;
      }
    }
    tc::fence_after_sync();
    // ---- gather q_i * k_j into the (dead) A1 region: one warp per row, full-line loads ---------------------------
    for (int rr = warp; rr < TM; rr += NT / 32) {
      const float* qi = proj + (size_t)s_i[rr] * F5 + qoff;
      const float* kj = proj + (size_t)s_j[rr] * F5 + koff;
      for (int c4 = lane; c4 < F / 4; c4 += 32) {
        float4 q = *reinterpret_cast<const float4*>(qi + 4 * c4);
        float4 k = *reinterpret_cast<const float4*>(kj + 4 * c4);
        *reinterpret_cast<float4*>(qk + (size_t)rr * T.QS + 4 * c4) = make_float4(q.x * k.x, q.y * k.y, q.z * k.z, q.w * k.w);
      }
    }
#pragma unroll
    for (int h = 0; h < 8; ++h) s_part[(r * 8 + h) * 2 + hh] = 0.f;
    __syncthreads();
    // ---- epilogue 2: alpha_h = sum_{f in head h} (w_f + b_f) q_f k_f / sqrt(d_h) * phi -------------------------------
    {
      const int f_begin = hh * half_cols;
      int h = f_begin / dhd;
      int nb = (h + 1) * dhd;
      float acc = 0.f;
      const float* qrow = qk + (size_t)r * T.QS;
      for (int cc = 0; cc < half_cols / 8; ++cc) {
        const int c0 = f_begin + cc * 8;
        float v[8];
        tc::tmem_ld8(tD2 + lane_sel + (uint32_t)c0, v);
        tc::tmem_ld_wait();
        if (c0 < F) {
          float4 qa = *reinterpret_cast<const float4*>(qrow + c0);
          float4 qb = (c0 + 4 < F) ? *reinterpret_cast<const float4*>(qrow + c0 + 4) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float qq[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int f = c0 + q;
            if (f < F) {
              if (f == nb) {
                s_part[(r * 8 + h) * 2 + hh] = acc;
                acc = 0.f;
                ++h;
                nb += dhd;
              }
              acc = fmaf(v[q] + b2p[f], qq[q], acc);
            }
          }
        }
      }
      if (f_begin < F) s_part[(r * 8 + h) * 2 + hh] = acc;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (tid < ne) {
      const float ph = s_phi[tid];
      for (int h = 0; h < heads; ++h) {
        float s = s_part[(tid * 8 + h) * 2 + 0] + s_part[(tid * 8 + h) * 2 + 1];
        alpha[(size_t)(e0 + tid) * Ast + aoff + h] = s * inv * ph;
      }
      if (blk == 1)
        for (int a = D.H + D.nl; a < Ast; ++a) alpha[(size_t)(e0 + tid) * Ast + a] = 0.f;
    }
    __syncthreads();      // qk tile / s_part fully consumed before the next tile overwrites the region
  }
  tc::fence_before_sync();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

bool so3k_edge_tc_supported(const So3kDims& D) {
  if (D.F % 4 != 0 || D.K % 16 != 0) return false;
  TcDims T = make_tc_dims(D);
  if (T.Fp > 256 || T.Fp % 16 != 0) return false;
  if (D.H > 8 || D.nl > 8) return false;
  return T.total + 1024 <= 227u * 1024u;
}

size_t so3k_edge_tc_image_bytes(const So3kDims& D) {
  TcDims T = make_tc_dims(D);
  return ((size_t)(2 * T.szW1 + 2 * T.szW2) + 255) / 256 * 256 + sizeof(float) * 2 * T.Fp + 256;
}

void so3k_edge_tc_build(const So3kDims& D, const float* p, const BlockParams& bp, unsigned char* dst, TcBlockW* out,
                        const EdgeBlockW& simt, cudaStream_t st) {
  TcDims T = make_tc_dims(D);
  size_t img_bytes = ((size_t)(2 * T.szW1 + 2 * T.szW2) + 255) / 256 * 256;
  float* bias = reinterpret_cast<float*>(dst + img_bytes);
  int tot = T.Fp * T.KC;
  so3k_count_launch();
  k_build_images<<<so3k_cdiv(tot, 256), 256, 0, st>>>(D, T, p + bp.rad_W1, p + bp.rad_b1, p + bp.rad_W2, p + bp.rad_b2,
                                                       p + bp.sph_W2, p + bp.sph_b2, dst, bias);
  out->img = dst;
  out->bias = bias;
  out->Ws1 = simt.Ws1;
  out->bs1 = simt.bs1;
}

void so3k_launch_edge_fwd_tc(const So3kDims& D, const Graph& g, const TcBlockW& Wf, const TcBlockW& Wg, const float* chi,
                             const float* proj, float* alpha, int* status, cudaStream_t st) {
  if (g.Pt <= 0) return;
  TcDims T = make_tc_dims(D);
  int Ast = (D.H + D.nl + 3) & ~3;
  int n_tiles = so3k_cdiv(g.Pt, TM);
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaFuncSetAttribute(k_edge_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(T.total + 1024));
  }
  int grid = n_tiles < n_sm ? n_tiles : n_sm;
  for (int blk = 0; blk < 2; ++blk) {
    const TcBlockW& W = blk ? Wg : Wf;
    so3k_count_launch();
    k_edge_fwd_tc<<<grid, NT, T.total + 1024, st>>>(D, T, blk, Ast, n_tiles, g.n_valid, g.row, g.col, g.geom, chi, proj,
                                                     W.img, W.bias, W.Ws1, W.bs1, alpha, status);
  }
}
#else
bool so3k_edge_tc_supported(const So3kDims&) { return false; }
size_t so3k_edge_tc_image_bytes(const So3kDims&) { return 0; }
void so3k_edge_tc_build(const So3kDims&, const float*, const BlockParams&, unsigned char*, TcBlockW*, const EdgeBlockW&,
                        cudaStream_t) {}
void so3k_launch_edge_fwd_tc(const So3kDims&, const Graph&, const TcBlockW&, const TcBlockW&, const float*, const float*,
                             float*, int*, cudaStream_t) {}
#endif
