// Per-edge filter MLPs + attention coefficients (forward) and their analytic data-gradient (backward),
// fp32 FMA path on the CUDA cores.  A CTA owns a tile of TE consecutive receiver-sorted edges; every
// intermediate (radial basis, hidden activations, the F-wide filter w_ij, its cotangent) lives in shared
// memory / registers and never touches HBM -- per edge the kernel reads a 16-byte geometry record, two indices,
// the gathered q_i / k_j rows and chi rows, and writes H+nl attention coefficients.
//   chi_ij, l0 contraction     mlff/nn/layer/so3krates_layer.py:89-93 ; mlff/sph_ops/contract.py:181-188
//   RadialSphericalFilter      mlff/nn/layer/so3krates_layer.py:449-465 ; MLP mlff/nn/mlp/mlp.py:21-27
//   ConvAttentionCoefficients  mlff/nn/layer/so3krates_layer.py:568-586 ; masking/cutoff :500-501, :551-554
// Backward: recompute the tile's forward, then
//   wbar = sbar q_i k_j / sqrt(d_h)              -> W2^T -> silu' -> (forward-mode) d/dd of the radial basis
//   qbar_i += sbar w k_j / sqrt(d_h)             (receiver-indexed)
//   kbar_i += sbar_rev w q_j / sqrt(d_h)         (sender-indexed sum turned receiver-indexed: w_ij == w_ji)
//   gbar   = (hbar_s silu') Ws1^T ; phibar = sum_h alphabar s ; Ybar = alpha' chibar1_i
//   rbar  += dbar u + (ubar - (ubar.u) u) / d
// The two tile GEMMs are register-tiled 8 edges x 4 features per thread with the weights streamed through L1/L2.
#include "so3k_internal.h"
#include "so3k_edge.h"

namespace {

constexpr int NT = 256;

struct Tile {
  int TE, RG;          // edges per tile, row groups (TE = 8 * RG)
  int lda0, lda1, ldp; // leading dimensions of A0 (rbf), A1/D1 (hidden), Pb/Wb (F wide)
  int Fs4, KC;
};

__device__ __forceinline__ void fma4(float* acc, float a, const float4& w) {
  acc[0] = fmaf(a, w.x, acc[0]);
  acc[1] = fmaf(a, w.y, acc[1]);
  acc[2] = fmaf(a, w.z, acc[2]);
  acc[3] = fmaf(a, w.w, acc[3]);
}

// acc[m][0..3] += A[r*8+m][0:Kin] @ W[0:Kin][4c..4c+3]      (Kin % 4 == 0, 16-byte aligned rows)
__device__ __forceinline__ void gemm8x4(const float* As, int lda, int Kin, const float* __restrict__ W, int ldw,
                                        int c, int r, float (*acc)[4]) {
  const float* a = As + (size_t)r * 8 * lda;
  const float* w = W + 4 * c;
  for (int k = 0; k < Kin; k += 4) {
    float4 w0 = *reinterpret_cast<const float4*>(w + (size_t)(k + 0) * ldw);
    float4 w1 = *reinterpret_cast<const float4*>(w + (size_t)(k + 1) * ldw);
    float4 w2 = *reinterpret_cast<const float4*>(w + (size_t)(k + 2) * ldw);
    float4 w3 = *reinterpret_cast<const float4*>(w + (size_t)(k + 3) * ldw);
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      float4 av = *reinterpret_cast<const float4*>(a + m * lda + k);
      fma4(acc[m], av.x, w0);
      fma4(acc[m], av.y, w1);
      fma4(acc[m], av.z, w2);
      fma4(acc[m], av.w, w3);
    }
  }
}

struct Smem {
  float *A0, *A0d, *A1, *D1, *Wb, *Pb;
  float *s_d, *s_ed, *s_phi, *s_g, *s_dbar, *s_phibar, *s_gbar, *s_ab, *s_abr;
  int *s_i, *s_j;
};

__device__ __forceinline__ Smem carve(float* base, const Tile& T, const So3kDims& D, int Ast, bool bwd) {
  Smem s;
  float* p = base;
  s.A0 = p; p += T.TE * T.lda0;
  s.A1 = p; p += T.TE * T.lda1;
  s.Pb = p; p += T.TE * T.ldp;
  s.s_d = p; p += T.TE;
  s.s_ed = p; p += T.TE;
  s.s_phi = p; p += T.TE;
  s.s_g = p; p += T.TE * D.nl;
  s.s_i = reinterpret_cast<int*>(p); p += T.TE;
  s.s_j = reinterpret_cast<int*>(p); p += T.TE;
  s.A0d = s.D1 = s.Wb = s.s_dbar = s.s_phibar = s.s_gbar = s.s_ab = s.s_abr = nullptr;
  if (bwd) {
    p = base + (((p - base) + 3) & ~3);
    s.A0d = p; p += T.TE * T.lda0;
    s.D1 = p; p += T.TE * T.lda1;
    s.Wb = p; p += T.TE * T.ldp;
    s.s_dbar = p; p += T.TE;
    s.s_phibar = p; p += T.TE;
    s.s_gbar = p; p += T.TE * D.nl;
    s.s_ab = p; p += T.TE * Ast;
    s.s_abr = p; p += T.TE * Ast;
  }
  return s;
}

size_t smem_floats(const Tile& T, const So3kDims& D, int Ast, bool bwd) {
  size_t n = (size_t)T.TE * (T.lda0 + T.lda1 + T.ldp + 5 + D.nl);
  if (bwd) {
    n = (n + 3) & ~(size_t)3;
    n += (size_t)T.TE * (T.lda0 + T.lda1 + T.ldp + 2 + D.nl + 2 * Ast);
  }
  return n + 4;
}

// per-edge metadata + l0 contraction of the chi difference (so3krates_layer.py:89-93)
__device__ __forceinline__ void load_meta(const So3kDims& D, const Tile& T, const Smem& S, int e0, int ne,
                                          const int* __restrict__ row, const int* __restrict__ col,
                                          const float4* __restrict__ geom, const float* __restrict__ chi) {
  for (int t = threadIdx.x; t < T.TE; t += NT) {
    int i = 0, j = 0;
    float d = 0.f, ed = 0.f, ph = 0.f;
    float gl[SO3K_MAXDEG];
    for (int l = 0; l < D.nl; ++l) gl[l] = 0.f;
    if (t < ne) {
      int e = e0 + t;
      i = row[e];
      j = col[e];
      d = geom[e].w;
      ed = expf(-d);
      ph = so3k_cutoff(d, D.r_cut);
      for (int m = 0; m < D.M; ++m) {
        float c = chi[(size_t)j * D.M + m] - chi[(size_t)i * D.M + m];
        gl[D.m_seg[m]] += c * c * D.m_cg[m];
      }
    }
    S.s_i[t] = i; S.s_j[t] = j; S.s_d[t] = d; S.s_ed[t] = ed; S.s_phi[t] = ph;
    for (int l = 0; l < D.nl; ++l) S.s_g[t * D.nl + l] = gl[l];
  }
}

// hidden layer of both MLPs for one block: A1 = [silu(rbf W1 + b1) | silu(g Ws1 + bs1) | 0]; optionally D1 = silu'
__device__ __forceinline__ void hidden_layer(const So3kDims& D, const Tile& T, const Smem& S, const EdgeBlockW& W,
                                             bool want_d) {
  const int F = D.F, NC = F / 4;
  if ((int)threadIdx.x < NC * T.RG) {
    const int c = threadIdx.x % NC, r = threadIdx.x / NC;
    float acc[8][4];
#pragma unroll
    for (int m = 0; m < 8; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
    gemm8x4(S.A0, T.lda0, D.K, W.W1, F, c, r, acc);
    float4 b = *reinterpret_cast<const float4*>(W.b1 + 4 * c);
    const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      float y[4], dy[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) so3k_silu_d(acc[m][q] + bb[q], &y[q], &dy[q]);
      int t = r * 8 + m;
      *reinterpret_cast<float4*>(S.A1 + t * T.lda1 + 4 * c) = make_float4(y[0], y[1], y[2], y[3]);
      if (want_d) *reinterpret_cast<float4*>(S.D1 + t * T.lda1 + 4 * c) = make_float4(dy[0], dy[1], dy[2], dy[3]);
    }
  }
  for (int idx = threadIdx.x; idx < T.TE * T.Fs4; idx += NT) {
    int t = idx / T.Fs4, cc = idx % T.Fs4;
    float y = 0.f, dy = 0.f;
    if (cc < D.Fs) {
      float a = W.bs1[cc];
      for (int l = 0; l < D.nl; ++l) a = fmaf(S.s_g[t * D.nl + l], W.Ws1[l * D.Fs + cc], a);
      so3k_silu_d(a, &y, &dy);
    }
    S.A1[t * T.lda1 + F + cc] = y;
    if (want_d) S.D1[t * T.lda1 + F + cc] = dy;
  }
}

__global__ void __launch_bounds__(NT)
k_edge_fwd(So3kDims D, Tile T, int Ast, const int* __restrict__ n_valid, const int* __restrict__ row,
           const int* __restrict__ col, const float4* __restrict__ geom, const float* __restrict__ chi,
           const float* __restrict__ proj, EdgeBlockW Wf, EdgeBlockW Wg, float* __restrict__ alpha) {
  SO3K_DYN_SMEM(float, smem);
  const int nv = *n_valid;
  const int e0 = blockIdx.x * T.TE;
  if (e0 >= nv) return;                          // block-uniform
  const int ne = (nv - e0) < T.TE ? (nv - e0) : T.TE;
  const Smem S = carve(smem, T, D, Ast, false);
  const int F = D.F, NC = F / 4;
  const size_t F5 = 5 * (size_t)F;

  load_meta(D, T, S, e0, ne, row, col, geom, chi);
  __syncthreads();
  for (int idx = threadIdx.x; idx < T.TE * D.K; idx += NT) {
    int t = idx / D.K, k = idx % D.K;
    S.A0[t * T.lda0 + k] = (t < ne) ? so3k_rbf(S.s_ed[t], k, D) : 0.f;
  }
  __syncthreads();

  for (int blk = 0; blk < 2; ++blk) {
    const EdgeBlockW& W = blk ? Wg : Wf;
    const int heads = blk ? D.nl : D.H;
    const int dhd = F / heads;
    const int aoff = blk ? D.H : 0;
    const int qoff = blk ? F : 0;
    const int koff = 2 * F + (blk ? F : 0);
    hidden_layer(D, T, S, W, false);
    __syncthreads();
    if ((int)threadIdx.x < NC * T.RG) {
      const int c = threadIdx.x % NC, r = threadIdx.x / NC;
      float acc[8][4];
#pragma unroll
      for (int m = 0; m < 8; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
      gemm8x4(S.A1, T.lda1, T.KC, W.W2c, F, c, r, acc);
      float4 b = *reinterpret_cast<const float4*>(W.b2c + 4 * c);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        int t = r * 8 + m;
        float4 q = *reinterpret_cast<const float4*>(proj + (size_t)S.s_i[t] * F5 + qoff + 4 * c);
        float4 k = *reinterpret_cast<const float4*>(proj + (size_t)S.s_j[t] * F5 + koff + 4 * c);
        *reinterpret_cast<float4*>(S.Pb + t * T.ldp + 4 * c) =
            make_float4((acc[m][0] + b.x) * q.x * k.x, (acc[m][1] + b.y) * q.y * k.y, (acc[m][2] + b.z) * q.z * k.z,
                        (acc[m][3] + b.w) * q.w * k.w);
      }
    }
    __syncthreads();
    const float inv = 1.0f / sqrtf((float)dhd);
    for (int idx = threadIdx.x; idx < T.TE * heads; idx += NT) {
      int t = idx / heads, h = idx % heads;
      if (t < ne) {
        const float* p = S.Pb + t * T.ldp + h * dhd;
        float s = 0.f;
        for (int f = 0; f < dhd; ++f) s += p[f];
        alpha[(size_t)(e0 + t) * Ast + aoff + h] = s * inv * S.s_phi[t];
      }
    }
    __syncthreads();
  }
  for (int idx = threadIdx.x; idx < ne * (Ast - D.H - D.nl); idx += NT) {
    int npad = Ast - D.H - D.nl;
    int t = idx / npad, a = idx % npad;
    alpha[(size_t)(e0 + t) * Ast + D.H + D.nl + a] = 0.f;
  }
}

__global__ void __launch_bounds__(NT)
k_edge_bwd(So3kDims D, Tile T, int Ast, int Gst, const int* __restrict__ n_valid, const int* __restrict__ row,
           const int* __restrict__ col, const int* __restrict__ rev, const float4* __restrict__ geom,
           const float* __restrict__ chi, const float* __restrict__ proj, EdgeBlockW Wf, EdgeBlockW Wg,
           const float* __restrict__ alphabar, int parts, float* __restrict__ gproj, float* __restrict__ gbar,
           float* __restrict__ ebar) {
  SO3K_DYN_SMEM(float, smem);
  const int nv = *n_valid;
  const int e0 = blockIdx.x * T.TE;
  if (e0 >= nv) return;                          // block-uniform
  const int ne = (nv - e0) < T.TE ? (nv - e0) : T.TE;
  const Smem S = carve(smem, T, D, Ast, true);
  const int F = D.F, NC = F / 4, NC3 = T.KC / 4;
  const size_t F5 = 5 * (size_t)F;

  load_meta(D, T, S, e0, ne, row, col, geom, chi);
  for (int idx = threadIdx.x; idx < T.TE * Ast; idx += NT) {
    int t = idx / Ast, a = idx % Ast;
    float v = 0.f, vr = 0.f;
    if (t < ne) {
      v = alphabar[(size_t)(e0 + t) * Ast + a];
      vr = alphabar[(size_t)rev[e0 + t] * Ast + a];
    }
    S.s_ab[idx] = v;
    S.s_abr[idx] = vr;
  }
  for (int t = threadIdx.x; t < T.TE; t += NT) {
    S.s_dbar[t] = 0.f;
    S.s_phibar[t] = 0.f;
    for (int l = 0; l < D.nl; ++l) S.s_gbar[t * D.nl + l] = 0.f;
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < T.TE * D.K; idx += NT) {
    int t = idx / D.K, k = idx % D.K;
    float v = 0.f, dv = 0.f;
    if (t < ne) {
      v = so3k_rbf(S.s_ed[t], k, D);
      dv = so3k_rbf_d(S.s_ed[t], k, v, D);
    }
    S.A0[t * T.lda0 + k] = v;
    S.A0d[t * T.lda0 + k] = dv;
  }
  __syncthreads();

  for (int blk = 0; blk < 2; ++blk) {
    const EdgeBlockW& W = blk ? Wg : Wf;
    const int heads = blk ? D.nl : D.H;
    const int dhd = F / heads;
    const int aoff = blk ? D.H : 0;
    const int qoff = blk ? F : 0;
    const int koff = 2 * F + (blk ? F : 0);
    const float inv = 1.0f / sqrtf((float)dhd);
    hidden_layer(D, T, S, W, true);
    __syncthreads();

    // ---- filter w (registers) --------------------------------------------------------------------
    const bool act = (int)threadIdx.x < NC * T.RG;
    const int c = threadIdx.x % NC, r = threadIdx.x / NC;
    float wreg[8][4];
#pragma unroll
    for (int m = 0; m < 8; ++m) wreg[m][0] = wreg[m][1] = wreg[m][2] = wreg[m][3] = 0.f;
    if (act) {
      gemm8x4(S.A1, T.lda1, T.KC, W.W2c, F, c, r, wreg);
      float4 b = *reinterpret_cast<const float4*>(W.b2c + 4 * c);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        wreg[m][0] += b.x; wreg[m][1] += b.y; wreg[m][2] += b.z; wreg[m][3] += b.w;
      }
      // Pb = q k w (for the pre-cutoff coefficient s) ; Wb = wbar = sbar q k / sqrt(dh)
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        int t = r * 8 + m;
        float4 q = *reinterpret_cast<const float4*>(proj + (size_t)S.s_i[t] * F5 + qoff + 4 * c);
        float4 k = *reinterpret_cast<const float4*>(proj + (size_t)S.s_j[t] * F5 + koff + 4 * c);
        const float qq[4] = {q.x, q.y, q.z, q.w}, kk[4] = {k.x, k.y, k.z, k.w};
        float pb[4], wb[4];
#pragma unroll
        for (int x = 0; x < 4; ++x) {
          int h = (4 * c + x) / dhd;
          float sb = S.s_ab[t * Ast + aoff + h] * S.s_phi[t] * inv;
          float qk = qq[x] * kk[x];
          pb[x] = qk * wreg[m][x];
          wb[x] = sb * qk;
        }
        *reinterpret_cast<float4*>(S.Pb + t * T.ldp + 4 * c) = make_float4(pb[0], pb[1], pb[2], pb[3]);
        *reinterpret_cast<float4*>(S.Wb + t * T.ldp + 4 * c) = make_float4(wb[0], wb[1], wb[2], wb[3]);
      }
    }
    __syncthreads();
    // phibar += sum_h alphabar_h * s_h
    for (int t = threadIdx.x; t < ne; t += NT) {
      float acc = 0.f;
      for (int h = 0; h < heads; ++h) {
        const float* p = S.Pb + t * T.ldp + h * dhd;
        float s = 0.f;
        for (int f = 0; f < dhd; ++f) s += p[f];
        acc = fmaf(S.s_ab[t * Ast + aoff + h], s * inv, acc);
      }
      S.s_phibar[t] += acc;
    }
    __syncthreads();
    // ---- qbar (receiver rows of this tile) and kbar (through the reverse edges) --------------------
    for (int pass = 0; pass < 2; ++pass) {
      if (act) {
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          int t = r * 8 + m;
          const float* other = pass == 0 ? proj + (size_t)S.s_j[t] * F5 + koff + 4 * c    // k_j
                                         : proj + (size_t)S.s_j[t] * F5 + qoff + 4 * c;   // q_j
          float4 o = *reinterpret_cast<const float4*>(other);
          const float oo[4] = {o.x, o.y, o.z, o.w};
          const float* ab = pass == 0 ? S.s_ab : S.s_abr;
          float pb[4];
#pragma unroll
          for (int x = 0; x < 4; ++x) {
            int h = (4 * c + x) / dhd;
            float sb = ab[t * Ast + aoff + h] * S.s_phi[t] * inv;
            pb[x] = sb * wreg[m][x] * oo[x];
          }
          *reinterpret_cast<float4*>(S.Pb + t * T.ldp + 4 * c) = make_float4(pb[0], pb[1], pb[2], pb[3]);
        }
      }
      __syncthreads();
      const int goff = pass == 0 ? qoff : koff;
      for (int f = threadIdx.x; f < ((parts & 1) ? F : 0); f += NT) {
        int cur = S.s_i[0];
        float acc = 0.f;
        for (int t = 0; t < ne; ++t) {
          int it = S.s_i[t];
          if (it != cur) {
            atomicAdd(&gproj[(size_t)cur * F5 + goff + f], acc);
            acc = 0.f;
            cur = it;
          }
          acc += S.Pb[t * T.ldp + f];
        }
        atomicAdd(&gproj[(size_t)cur * F5 + goff + f], acc);
      }
      __syncthreads();
    }
    // ---- hbar = wbar @ W2c^T ; abar = hbar * silu'  (in place in D1) ---------------------------------
    if ((int)threadIdx.x < NC3 * T.RG) {
      const int c3 = threadIdx.x % NC3, r3 = threadIdx.x / NC3;
      float acc[8][4];
#pragma unroll
      for (int m = 0; m < 8; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
      gemm8x4(S.Wb, T.ldp, F, W.W2cT, T.KC, c3, r3, acc);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        float4* dp = reinterpret_cast<float4*>(S.D1 + (r3 * 8 + m) * T.lda1 + 4 * c3);
        float4 dv = *dp;
        *dp = make_float4(dv.x * acc[m][0], dv.y * acc[m][1], dv.z * acc[m][2], dv.w * acc[m][3]);
      }
    }
    __syncthreads();
    // ---- forward-mode radial derivative: (rbf' W1) . abar1 ---------------------------------------------
    if (act) {
      float acc[8][4];
#pragma unroll
      for (int m = 0; m < 8; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
      gemm8x4(S.A0d, T.lda0, D.K, W.W1, F, c, r, acc);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        int t = r * 8 + m;
        float4 dv = *reinterpret_cast<const float4*>(S.D1 + t * T.lda1 + 4 * c);
        *reinterpret_cast<float4*>(S.Pb + t * T.ldp + 4 * c) =
            make_float4(dv.x * acc[m][0], dv.y * acc[m][1], dv.z * acc[m][2], dv.w * acc[m][3]);
      }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < ne; t += NT) {
      float s = 0.f;
      const float* p = S.Pb + t * T.ldp;
      for (int f = 0; f < F; ++f) s += p[f];
      S.s_dbar[t] += s;
      const float* as = S.D1 + t * T.lda1 + F;
      for (int l = 0; l < D.nl; ++l) {
        float gsum = 0.f;
        for (int cc = 0; cc < D.Fs; ++cc) gsum = fmaf(as[cc], W.Ws1[l * D.Fs + cc], gsum);
        S.s_gbar[t * D.nl + l] += gsum;
      }
    }
    __syncthreads();
  }

  // ---- per-edge outputs: radial cotangent, cutoff cotangent, l0-contraction cotangent ------------------------
  for (int t = threadIdx.x; t < ne; t += NT) {
    const int e = e0 + t;
    if (parts & 1) ebar[2 * (size_t)e + 1] = S.s_phibar[t];
    if (parts & 2) {
      ebar[2 * (size_t)e + 0] = S.s_dbar[t];
      for (int l = 0; l < D.nl; ++l) gbar[(size_t)e * Gst + l] = S.s_gbar[t * D.nl + l];
    }
  }
}

// Chain the per-edge cotangents to the displacement vector (thread per edge):
//   dbar = dbar_rad + phibar * phi'(d) ; Ybar = alpha' chibar1_i -> ubar ; rbar += dbar u + (ubar - (ubar.u) u) / d
__global__ void k_edge_finish(So3kDims D, int Ast, const int* __restrict__ n_valid, const int* __restrict__ row,
                              const float4* __restrict__ geom, const float* __restrict__ alpha,
                              const float* __restrict__ chibar1, const float* __restrict__ ebar,
                              float* __restrict__ rbar) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < *n_valid) {
    const float4 g = geom[e];
    const float d = g.w;
    const int i = row[e];
    float dbar = ebar[2 * (size_t)e + 0] + ebar[2 * (size_t)e + 1] * so3k_cutoff_d(d, D.r_cut);
    float ybar[SO3K_MAXM];
    for (int m = 0; m < D.M; ++m)
      ybar[m] = alpha[(size_t)e * Ast + D.H + D.m_seg[m]] * chibar1[(size_t)i * D.M + m];
    float ux = 0.f, uy = 0.f, uz = 0.f;
    for (int s = 0; s < D.nl; ++s) so3k_sph_vjp(D.deg[s], g.x, g.y, g.z, ybar + D.seg_off[s], &ux, &uy, &uz);
    float rx = 0.f, ry = 0.f, rz = 0.f;
    if (d > 0.f) {
      float dot = ux * g.x + uy * g.y + uz * g.z;
      float id = 1.0f / d;
      rx = dbar * g.x + (ux - dot * g.x) * id;
      ry = dbar * g.y + (uy - dot * g.y) * id;
      rz = dbar * g.z + (uz - dot * g.z) * id;
    }
    rbar[3 * (size_t)e + 0] += rx;
    rbar[3 * (size_t)e + 1] += ry;
    rbar[3 * (size_t)e + 2] += rz;
  }
}

Tile make_tile(const So3kDims& D) {
  Tile T;
  T.Fs4 = (D.Fs + 3) & ~3;
  T.KC = D.F + T.Fs4;
  int ncmax = T.KC / 4;
  int rg = NT / ncmax;
  if (rg > 8) rg = 8;
  if (rg < 1) rg = 1;
  T.RG = rg;
  T.TE = 8 * rg;
  T.lda0 = D.K + 4;
  T.lda1 = T.KC + 4;
  T.ldp = D.F + 4;
  return T;
}

}  // namespace

namespace {
__global__ void k_repack_block(int K, int F, int Fs, int nl, int KC, const float* __restrict__ rW1,
                               const float* __restrict__ rb1, const float* __restrict__ rW2,
                               const float* __restrict__ rb2, const float* __restrict__ sW1,
                               const float* __restrict__ sb1, const float* __restrict__ sW2,
                               const float* __restrict__ sb2, float* W1, float* b1, float* W2c, float* W2cT,
                               float* b2c, float* Ws1, float* bs1) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < K * F) W1[t] = rW1[t];
  if (t < F) {
    b1[t] = rb1[t];
    b2c[t] = rb2[t] + sb2[t];
  }
  if (t < KC * F) {
    int k = t / F, f = t % F;
    float v = 0.f;
    if (k < F) v = rW2[(size_t)k * F + f];
    else if (k - F < Fs) v = sW2[(size_t)(k - F) * F + f];
    W2c[t] = v;
    W2cT[(size_t)f * KC + k] = v;
  }
  if (t < nl * Fs) Ws1[t] = sW1[t];
  if (t < Fs) bs1[t] = sb1[t];
}
inline size_t r4(size_t n) { return (n + 3) & ~(size_t)3; }
}  // namespace

size_t so3k_edge_repack_floats(const So3kDims& D) {
  size_t KC = D.F + ((D.Fs + 3) & ~3);
  return r4((size_t)D.K * D.F) + r4(D.F) + 2 * r4(KC * D.F) + r4(D.F) + r4((size_t)D.nl * D.Fs) + r4(D.Fs);
}

void so3k_edge_repack(const So3kDims& D, const float* p, const BlockParams& bp, float* dst, EdgeBlockW* out,
                      cudaStream_t st) {
  int KC = D.F + ((D.Fs + 3) & ~3);
  float* W1 = dst;
  float* b1 = W1 + r4((size_t)D.K * D.F);
  float* W2c = b1 + r4(D.F);
  float* W2cT = W2c + r4((size_t)KC * D.F);
  float* b2c = W2cT + r4((size_t)KC * D.F);
  float* Ws1 = b2c + r4(D.F);
  float* bs1 = Ws1 + r4((size_t)D.nl * D.Fs);
  int tot = KC * D.F;
  if (D.K * D.F > tot) tot = D.K * D.F;
  SO3K_LAUNCH(k_repack_block, dim3(so3k_cdiv(tot, 256)), dim3(256), 0, st, D.K, D.F, D.Fs, D.nl, KC, p + bp.rad_W1,
              p + bp.rad_b1, p + bp.rad_W2, p + bp.rad_b2, p + bp.sph_W1, p + bp.sph_b1, p + bp.sph_W2, p + bp.sph_b2,
              W1, b1, W2c, W2cT, b2c, Ws1, bs1);
  out->W1 = W1; out->b1 = b1; out->W2c = W2c; out->W2cT = W2cT; out->b2c = b2c; out->Ws1 = Ws1; out->bs1 = bs1;
}

bool so3k_edge_supported(const So3kDims& D) {
  if (D.F % 4 != 0 || D.K % 4 != 0) return false;
  if ((D.F + ((D.Fs + 3) & ~3)) / 4 > NT) return false;
  return true;
}

void so3k_launch_edge_fwd(const So3kDims& D, const Graph& g, const EdgeBlockW& Wf, const EdgeBlockW& Wg,
                          const float* chi, const float* proj, float* alpha, cudaStream_t st) {
  if (g.Pt <= 0) return;
  Tile T = make_tile(D);
  int Ast = (D.H + D.nl + 3) & ~3;
  size_t smem = sizeof(float) * smem_floats(T, D, Ast, false);
  SO3K_SET_MAX_SMEM(k_edge_fwd, smem);
  SO3K_LAUNCH(k_edge_fwd, dim3(so3k_cdiv(g.Pt, T.TE)), dim3(NT), smem, st, D, T, Ast, g.n_valid, g.row, g.col, g.geom,
              chi, proj, Wf, Wg, alpha);
}

void so3k_launch_edge_bwd(const So3kDims& D, const Graph& g, const EdgeBlockW& Wf, const EdgeBlockW& Wg,
                          const float* chi, const float* proj, const float* alphabar, int parts, float* gproj,
                          float* gbar, float* ebar, cudaStream_t st) {
  if (g.Pt <= 0) return;
  Tile T = make_tile(D);
  int Ast = (D.H + D.nl + 3) & ~3;
  int Gst = (D.nl + 3) & ~3;
  size_t smem = sizeof(float) * smem_floats(T, D, Ast, true);
  SO3K_SET_MAX_SMEM(k_edge_bwd, smem);
  SO3K_LAUNCH(k_edge_bwd, dim3(so3k_cdiv(g.Pt, T.TE)), dim3(NT), smem, st, D, T, Ast, Gst, g.n_valid, g.row, g.col, g.rev,
              g.geom, chi, proj, Wf, Wg, alphabar, parts, gproj, gbar, ebar);
}

void so3k_launch_edge_finish(const So3kDims& D, const Graph& g, const float* alpha, const float* chibar1,
                             const float* ebar, float* rbar, cudaStream_t st) {
  if (g.Pt <= 0) return;
  int Ast = (D.H + D.nl + 3) & ~3;
  SO3K_LAUNCH(k_edge_finish, dim3(so3k_cdiv(g.Pt, 256)), dim3(256), 0, st, D, Ast, g.n_valid, g.row, g.geom, alpha, chibar1,
              ebar, rbar);
}
