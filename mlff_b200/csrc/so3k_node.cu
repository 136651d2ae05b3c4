// Node-level (per-atom) kernels: embedding, q/k/v projections, interaction block, energy read-out and
// their data-gradients.  These are < 2 % of the FLOPs of a step;
// they are written as small shared-memory tiled FMA kernels: a block owns TA atoms, the atom rows sit in
// shared memory TRANSPOSED ([column][atom], column stride LDA) so that one broadcast 16-byte load feeds four FMAs,
// each thread owns one output column and keeps TA accumulators so every weight element is read once per block
// (coalesced across the threads of a warp).  NT is sized to the ~F output columns of a pass; several blocks share an SM.
//   AtomTypeEmbed            mlff/nn/embed/embed.py:227-229
//   q / k / v Dense          mlff/nn/layer/so3krates_layer.py:583-584, 607  (per head, no bias)
//   InteractionBlock         mlff/nn/layer/so3krates_layer.py:355-379 (+ skip :163-164)
//   Energy                   mlff/nn/observable/observable.py:77-91
//   LayerNorm (optional)     mlff/nn/layer/so3krates_layer.py:103-106, 153-156, 170-174 (flax nn.LayerNorm defaults)
//   ResidualMLP (optional)   mlff/nn/mlp/mlp.py:30-72, so3krates_layer.py:149-150, 166-167
#include "so3k_internal.h"

namespace {

constexpr int TA = 16;        // atoms per block
constexpr int LDA = TA + 4;   // column stride of a transposed tile: 16-byte aligned, 20 mod 32 spreads the banks
constexpr int NT = 160;       // threads per block

// acc[a] += w * col[a], a < TA, col = one column of a transposed tile (all lanes read the same 16 bytes: broadcast)
__device__ __forceinline__ void axpy_col(const float* col, float w, float* acc) {
#pragma unroll
  for (int k = 0; k < TA / 4; ++k) {
    const float4 v = *reinterpret_cast<const float4*>(col + 4 * k);
    acc[4 * k + 0] = fmaf(v.x, w, acc[4 * k + 0]);
    acc[4 * k + 1] = fmaf(v.y, w, acc[4 * k + 1]);
    acc[4 * k + 2] = fmaf(v.z, w, acc[4 * k + 2]);
    acc[4 * k + 3] = fmaf(v.w, w, acc[4 * k + 3]);
  }
}
// acc[a] += sum_c xs[c0 + c][a] * W[c * ldw + o]
__device__ __forceinline__ void col_dot(const float* xs, int c0, int nC, const float* __restrict__ W, int ldw, int o,
                                        float* acc) {
  for (int c = 0; c < nC; ++c) axpy_col(xs + (size_t)(c0 + c) * LDA, W[(size_t)c * ldw + o], acc);
}
__global__ void k_embed(So3kDims D, int N, const int* __restrict__ z, const float* __restrict__ emb,
                        float* __restrict__ x, float* __restrict__ chi, int* __restrict__ status) {
  long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long tot = (long long)N * D.F;
  if (t < tot) {
    int a = (int)(t / D.F), f = (int)(t % D.F);
    int za = z[a];
    float v = 0.f;
    if (za > 0 && za < D.nemb) v = emb[(size_t)za * D.F + f];
    else if (za != 0 && f == 0 && status) atomicOr(status, SO3K_STATUS_BAD_Z);
    x[t] = v;                                   // z == 0: point_mask = 0 -> exact zero (embed.py:228)
  }
  if (t < (long long)N * D.M) chi[t] = 0.f;     // chi0 = 0 (embed.py:196-197)
}

// proj[a][5F] = [q_fb | q_gb | k_fb | k_gb | v_fb]
__global__ void __launch_bounds__(NT)
k_node_proj(So3kDims D, int N, const float* __restrict__ x, const float* __restrict__ Wq_f,
            const float* __restrict__ Wq_g, const float* __restrict__ Wk_f, const float* __restrict__ Wk_g,
            const float* __restrict__ Wv_f, float* __restrict__ proj) {
  SO3K_DYN_SMEM(float, xs);                     // [F][LDA]
  const int F = D.F;
  const int a0 = blockIdx.x * TA;
  for (int t = threadIdx.x; t < TA * F; t += NT) {
    int a = t / F, f = t % F;
    xs[f * LDA + a] = (a0 + a < N) ? x[(size_t)(a0 + a) * F + f] : 0.f;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < 5 * F; o += NT) {
    int seg = o / F, f = o % F;
    const float* W;
    int dsz;
    switch (seg) {
      case 0: W = Wq_f; dsz = D.dh; break;
      case 1: W = Wq_g; dsz = D.dg; break;
      case 2: W = Wk_f; dsz = D.dh; break;
      case 3: W = Wk_g; dsz = D.dg; break;
      default: W = Wv_f; dsz = D.dh; break;
    }
    int hd = f / dsz, oo = f % dsz;
    float acc[TA];
#pragma unroll
    for (int a = 0; a < TA; ++a) acc[a] = 0.f;
    col_dot(xs, hd * dsz, dsz, W + (size_t)hd * dsz * dsz, dsz, oo, acc);
#pragma unroll
    for (int a = 0; a < TA; ++a)
      if (a0 + a < N) proj[(size_t)(a0 + a) * 5 * F + o] = acc[a];
  }
}

// xbar[a][f] += sum over the five projections of  gproj[a][seg][head(f) block] @ W[seg][head]^T
// (the W* pointers are the TRANSPOSED per-head matrices: thread = column c reads WT[o][c], coalesced)
__global__ void __launch_bounds__(NT)
k_node_proj_bwd(So3kDims D, int N, const float* __restrict__ gproj, const float* __restrict__ Wq_f,
                const float* __restrict__ Wq_g, const float* __restrict__ Wk_f, const float* __restrict__ Wk_g,
                const float* __restrict__ Wv_f, float* __restrict__ xbar) {
  SO3K_DYN_SMEM(float, gs);                     // [5F][LDA]
  const int F = D.F, F5 = 5 * D.F;
  const int a0 = blockIdx.x * TA;
  for (int t = threadIdx.x; t < TA * F5; t += NT) {
    int a = t / F5, o = t % F5;
    gs[o * LDA + a] = (a0 + a < N) ? gproj[(size_t)(a0 + a) * F5 + o] : 0.f;
  }
  __syncthreads();
  for (int f = threadIdx.x; f < F; f += NT) {
    float acc[TA];
#pragma unroll
    for (int a = 0; a < TA; ++a) acc[a] = 0.f;
    for (int seg = 0; seg < 5; ++seg) {
      const float* W;
      int dsz;
      switch (seg) {
        case 0: W = Wq_f; dsz = D.dh; break;
        case 1: W = Wq_g; dsz = D.dg; break;
        case 2: W = Wk_f; dsz = D.dh; break;
        case 3: W = Wk_g; dsz = D.dg; break;
        default: W = Wv_f; dsz = D.dh; break;
      }
      int hd = f / dsz, c = f % dsz;
      col_dot(gs, seg * F + hd * dsz, dsz, W + (size_t)hd * dsz * dsz, dsz, c, acc);
    }
#pragma unroll
    for (int a = 0; a < TA; ++a)
      if (a0 + a < N) xbar[(size_t)(a0 + a) * F + f] += acc[a];
  }
}

// x2 = xbase + a ; chi2 = chi1 * (1 + repeat(b)) ; [a|b] = [x1 | l0(chi1)] W + bias ; xbase = nullptr: x1
// (with layer normalisation x1 is the normalised input and xbase the un-normalised skip connection)
__global__ void __launch_bounds__(NT)
k_interaction(So3kDims D, int N, const float* __restrict__ x1, const float* __restrict__ chi1,
              const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ xbase,
              float* __restrict__ x2, float* __restrict__ chi2, float* __restrict__ bcoef) {
  SO3K_DYN_SMEM(float, sm);
  const int F = D.F, nl = D.nl, M = D.M, FI = D.F + D.nl;
  float* ys = sm;                 // [FI][LDA]
  float* cs = ys + FI * LDA;      // [TA][M]
  float* bs = cs + TA * M;        // [TA][nl]
  const int a0 = blockIdx.x * TA;
  for (int t = threadIdx.x; t < TA * F; t += NT) {
    int a = t / F, f = t % F;
    ys[f * LDA + a] = (a0 + a < N) ? x1[(size_t)(a0 + a) * F + f] : 0.f;
  }
  for (int t = threadIdx.x; t < TA * M; t += NT) {
    int a = t / M, m = t % M;
    cs[t] = (a0 + a < N) ? chi1[(size_t)(a0 + a) * M + m] : 0.f;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < TA * nl; t += NT) {
    int a = t / nl, l = t % nl;
    float s = 0.f;
    for (int m = D.seg_off[l]; m < D.seg_off[l + 1]; ++m) s += cs[a * M + m] * cs[a * M + m] * D.m_cg[m];
    ys[(F + l) * LDA + a] = s;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < FI; o += NT) {
    float acc[TA];
    float bo = bias[o];
#pragma unroll
    for (int a = 0; a < TA; ++a) acc[a] = bo;
    col_dot(ys, 0, FI, W, FI, o, acc);
    if (o < F) {
#pragma unroll
      for (int a = 0; a < TA; ++a)
        if (a0 + a < N)
          x2[(size_t)(a0 + a) * F + o] = (xbase ? xbase[(size_t)(a0 + a) * F + o] : ys[o * LDA + a]) + acc[a];
    } else {
#pragma unroll
      for (int a = 0; a < TA; ++a) {
        bs[a * nl + (o - F)] = acc[a];
        if (a0 + a < N) bcoef[(size_t)(a0 + a) * nl + (o - F)] = acc[a];
      }
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < TA * M; t += NT) {
    int a = t / M, m = t % M;
    if (a0 + a < N) chi2[(size_t)(a0 + a) * M + m] = cs[t] + bs[a * nl + D.m_seg[m]] * cs[t];
  }
}

__global__ void __launch_bounds__(NT)
k_interaction_bwd(So3kDims D, int N, const float* __restrict__ chi1, const float* __restrict__ bcoef,
                  const float* __restrict__ W, const float* __restrict__ xbar_out,
                  const float* __restrict__ chibar_out, int skip, float* __restrict__ xbar1,
                  float* __restrict__ chibar1) {
  SO3K_DYN_SMEM(float, sm);
  const int F = D.F, nl = D.nl, M = D.M, FI = D.F + D.nl;
  float* gs = sm;                 // [FI][LDA] = [xbar_out | bbar] transposed
  float* cs = gs + FI * LDA;      // [TA][M]   chi1
  float* cb = cs + TA * M;        // [TA][M]   chibar_out
  float* Gb = cb + TA * M;        // [TA][nl]  gradient wrt l0(chi1)
  const int a0 = blockIdx.x * TA;
  for (int t = threadIdx.x; t < TA * F; t += NT) {
    int a = t / F, f = t % F;
    gs[f * LDA + a] = (a0 + a < N) ? xbar_out[(size_t)(a0 + a) * F + f] : 0.f;
  }
  for (int t = threadIdx.x; t < TA * M; t += NT) {
    int a = t / M, m = t % M;
    bool ok = a0 + a < N;
    cs[t] = ok ? chi1[(size_t)(a0 + a) * M + m] : 0.f;
    cb[t] = ok ? chibar_out[(size_t)(a0 + a) * M + m] : 0.f;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < TA * nl; t += NT) {
    int a = t / nl, l = t % nl;
    float s = 0.f;
    for (int m = D.seg_off[l]; m < D.seg_off[l + 1]; ++m) s += cb[a * M + m] * cs[a * M + m];
    gs[(F + l) * LDA + a] = s;    // bbar
  }
  __syncthreads();
  for (int c = threadIdx.x; c < FI; c += NT) {
    float acc[TA];
#pragma unroll
    for (int a = 0; a < TA; ++a) acc[a] = 0.f;
    col_dot(gs, 0, FI, W, FI, c, acc);              // W = int_W^T
    if (c < F) {
#pragma unroll
      for (int a = 0; a < TA; ++a)
        if (a0 + a < N) xbar1[(size_t)(a0 + a) * F + c] = skip ? gs[c * LDA + a] + acc[a] : acc[a];
    } else {
#pragma unroll
      for (int a = 0; a < TA; ++a) Gb[a * nl + (c - F)] = acc[a];
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < TA * M; t += NT) {
    int a = t / M, m = t % M;
    if (a0 + a < N) {
      int l = D.m_seg[m];
      float b = bcoef[(size_t)(a0 + a) * nl + l];
      chibar1[(size_t)(a0 + a) * M + m] = cb[t] * (1.0f + b) + 2.0f * D.m_cg[m] * cs[t] * Gb[a * nl + l];
    }
  }
}

// e = (Dense(F->1)(silu(Dense(F->F)(x)))) * scale[z] + shift[z], masked ; xbar = d(sum e)/dx
__global__ void __launch_bounds__(NT)
k_readout(So3kDims D, int N, const float* __restrict__ x, const int* __restrict__ z, const float* __restrict__ W1,
          const float* __restrict__ W1T, const float* __restrict__ b1, const float* __restrict__ W2,
          const float* __restrict__ b2, const float* __restrict__ scale, const float* __restrict__ shift, float out_scale,
          float* __restrict__ e_atom, float* __restrict__ xbar) {
  SO3K_DYN_SMEM(float, sm);
  const int F = D.F;
  float* xs = sm;               // [F][LDA]
  float* ab = xs + F * LDA;     // [F][LDA] abar
  float* es = ab + F * LDA;     // [TA][F]  silu(a) * W2
  float* coef = es + TA * F;    // [TA]
  const int a0 = blockIdx.x * TA;
  for (int t = threadIdx.x; t < TA * F; t += NT) {
    int a = t / F, f = t % F;
    xs[f * LDA + a] = (a0 + a < N) ? x[(size_t)(a0 + a) * F + f] : 0.f;
  }
  if (threadIdx.x < TA) {
    int a = a0 + threadIdx.x;
    float c = 0.f;
    if (a < N) {
      int za = z[a];
      if (za > 0 && za < D.nemb) c = scale[za] * out_scale;
    }
    coef[threadIdx.x] = c;
  }
  __syncthreads();
  for (int o = threadIdx.x; o < F; o += NT) {
    float acc[TA];
    float bo = b1[o];
#pragma unroll
    for (int a = 0; a < TA; ++a) acc[a] = bo;
    col_dot(xs, 0, F, W1, F, o, acc);
    float w2 = W2[o];
#pragma unroll
    for (int a = 0; a < TA; ++a) {
      float s, ds;
      so3k_silu_d(acc[a], &s, &ds);
      es[a * F + o] = s * w2;
      ab[o * LDA + a] = ds * w2 * coef[a];
    }
  }
  __syncthreads();
  if (threadIdx.x < TA) {
    int a = a0 + threadIdx.x;
    if (a < N) {
      float s = 0.f;
      for (int o = 0; o < F; ++o) s += es[threadIdx.x * F + o];
      s += b2[0];
      int za = z[a];
      float e = 0.f;
      if (za > 0 && za < D.nemb) e = (scale[za] * s + shift[za]) * out_scale;
      e_atom[a] = e;
    }
  }
  for (int c = threadIdx.x; c < F; c += NT) {
    float acc[TA];
#pragma unroll
    for (int a = 0; a < TA; ++a) acc[a] = 0.f;
    col_dot(ab, 0, F, W1T, F, c, acc);
#pragma unroll
    for (int a = 0; a < TA; ++a)
      if (a0 + a < N) xbar[(size_t)(a0 + a) * F + c] = acc[a];
  }
}

// E[b] = sum_a e_atom[b][a]   (float64 accumulation, one warp per structure chunk)
__global__ void k_energy_sum(int B, int n, const float* __restrict__ e_atom, float offset, float* __restrict__ E) {
  __shared__ double s_part[NT / 32];
  int b = blockIdx.x;
  double s = 0.0;
  for (int a = threadIdx.x; a < n; a += NT) s += (double)e_atom[(size_t)b * n + a];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < NT / 32; ++w) t += s_part[w];
    E[b] = (float)(t - (double)offset);
  }
}

// ---- optional branches ------------------------------------------------------------------------------------
// flax nn.LayerNorm defaults: statistics over the last axis, var = max(0, E[x^2] - E[x]^2) (use_fast_variance),
// eps = 1e-6, learnable scale and bias.  One warp per atom row.  (Rows of padding atoms are normalised like any
// other: they have no edges and are masked in the read-out, so the reference's safe_mask there is unobservable.)
constexpr float LN_EPS = 1e-6f;
__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__global__ void __launch_bounds__(128)
k_layernorm(int N, int F, const float* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ bias,
            float* __restrict__ y) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= N) return;                 // warp-uniform
  const float* xr = x + (size_t)row * F;
  float s1 = 0.f, s2 = 0.f;
  for (int f = lane; f < F; f += 32) {
    const float v = xr[f];
    s1 += v;
    s2 = fmaf(v, v, s2);
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const float mean = s1 / (float)F;
  const float var = fmaxf(0.f, s2 / (float)F - mean * mean);
  const float r = rsqrtf(var + LN_EPS);
  for (int f = lane; f < F; f += 32) y[(size_t)row * F + f] = (xr[f] - mean) * r * scale[f] + bias[f];
}

// out = add + r * (g - mean(g) - xhat * mean(g * xhat)),  g = gy * scale,  xhat = (x - mean) * r
// gy / add / out are not __restrict__: out may alias either (every element is read before it is written, by the
// lane that writes it)
__global__ void __launch_bounds__(128)
k_layernorm_bwd(int N, int F, const float* __restrict__ x, const float* __restrict__ scale, const float* gy,
                const float* add, float* out) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= N) return;
  const float* xr = x + (size_t)row * F;
  const float* gr = gy + (size_t)row * F;
  float s1 = 0.f, s2 = 0.f;
  for (int f = lane; f < F; f += 32) {
    const float v = xr[f];
    s1 += v;
    s2 = fmaf(v, v, s2);
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const float mean = s1 / (float)F;
  const float var = fmaxf(0.f, s2 / (float)F - mean * mean);
  const float r = rsqrtf(var + LN_EPS);
  float g1 = 0.f, g2 = 0.f;
  for (int f = lane; f < F; f += 32) {
    const float g = gr[f] * scale[f];
    g1 += g;
    g2 = fmaf(g, (xr[f] - mean) * r, g2);
  }
  g1 = warp_sum(g1) / (float)F;
  g2 = warp_sum(g2) / (float)F;
  for (int f = lane; f < F; f += 32) {
    const float g = gr[f] * scale[f];
    const float v = r * (g - g1 - (xr[f] - mean) * r * g2);
    out[(size_t)row * F + f] = add ? add[(size_t)row * F + f] + v : v;
  }
}

// tile[c][a] (transposed, column stride LDA) <- rows a0.. of a row-major (N,F) array
__device__ __forceinline__ void load_tile_T(float* tile, const float* src, int a0, int N, int F) {
  for (int t = threadIdx.x; t < TA * F; t += NT) {
    int a = t / F, f = t % F;
    tile[f * LDA + a] = (a0 + a < N) ? src[(size_t)(a0 + a) * F + f] : 0.f;
  }
}

// ResidualMLP(num_residuals=1, num_blocks_per_residual=2, use_bias=False), mlp.py:30-72:
//   h1 = silu(x) W0 ; r = x + silu(h1) W1 ; y = silu(r) W2
__global__ void __launch_bounds__(NT)
k_resmlp(So3kDims D, int N, const float* __restrict__ x, const float* __restrict__ W0, const float* __restrict__ W1,
         const float* __restrict__ W2, float* __restrict__ y) {
  SO3K_DYN_SMEM(float, sm);
  const int F = D.F;
  float* X = sm;                  // [F][LDA] x
  float* T0 = X + F * LDA;        // silu(x), then silu(r)
  float* T1 = T0 + F * LDA;       // silu(h1)
  const int a0 = blockIdx.x * TA;
  load_tile_T(X, x, a0, N, F);
  __syncthreads();
  for (int t = threadIdx.x; t < F * TA; t += NT) {
    const int k = (t / TA) * LDA + (t % TA);
    T0[k] = so3k_silu(X[k]);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < F; o += NT) {
    float acc[TA];
#pragma unroll
    for (int a = 0; a < TA; ++a) acc[a] = 0.f;
    col_dot(T0, 0, F, W0, F, o, acc);
#pragma unroll
    for (int a = 0; a < TA; ++a) T1[o * LDA + a] = so3k_silu(acc[a]);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < F; o += NT) {
    float acc[TA];
#pragma unroll
    for (int a = 0; a < TA; ++a) acc[a] = X[o * LDA + a];
    col_dot(T1, 0, F, W1, F, o, acc);
#pragma unroll
    for (int a = 0; a < TA; ++a) T0[o * LDA + a] = so3k_silu(acc[a]);     // T0 was last read before the barrier above
  }
  __syncthreads();
  for (int o = threadIdx.x; o < F; o += NT) {
    float acc[TA];
#pragma unroll
    for (int a = 0; a < TA; ++a) acc[a] = 0.f;
    col_dot(T0, 0, F, W2, F, o, acc);
#pragma unroll
    for (int a = 0; a < TA; ++a)
      if (a0 + a < N) y[(size_t)(a0 + a) * F + o] = acc[a];
  }
}

// xbar = rbar + (hbar1 W0^T) silu'(x) ; rbar = (ybar W2^T) silu'(r) ; hbar1 = (rbar W1^T) silu'(h1)
// h1 and r are recomputed from x.  W0 / W1 are the forward matrices, W0T / W1T / W2T their transposes.
// ybar / xbar are not __restrict__: xbar may alias ybar (a block reads its rows of ybar before it writes them).
__global__ void __launch_bounds__(NT)
k_resmlp_bwd(So3kDims D, int N, const float* __restrict__ x, const float* __restrict__ W0,
             const float* __restrict__ W1, const float* __restrict__ W0T, const float* __restrict__ W1T,
             const float* __restrict__ W2T, const float* ybar, float* xbar) {
  SO3K_DYN_SMEM(float, sm);
  const int F = D.F;
  float* X = sm;                  // [F][LDA] x
  float* H1 = X + F * LDA;        // h1, then hbar1
  float* T0 = H1 + F * LDA;       // silu(x), then r, then rbar
  float* T1 = T0 + F * LDA;       // silu(h1), then ybar
  const int a0 = blockIdx.x * TA;
  load_tile_T(X, x, a0, N, F);
  __syncthreads();
  for (int t = threadIdx.x; t < F * TA; t += NT) {
    const int k = (t / TA) * LDA + (t % TA);
    T0[k] = so3k_silu(X[k]);
  }
  __syncthreads();
  for (int o = threadIdx.x; o < F; o += NT) {
    float acc[TA];
#pragma unroll
    for (int a = 0; a < TA; ++a) acc[a] = 0.f;
    col_dot(T0, 0, F, W0, F, o, acc);
#pragma unroll
    for (int a = 0; a < TA; ++a) {
      H1[o * LDA + a] = acc[a];
      T1[o * LDA + a] = so3k_silu(acc[a]);
    }
  }
  __syncthreads();
  for (int o = threadIdx.x; o < F; o += NT) {
    float acc[TA];
#pragma unroll
    for (int a = 0; a < TA; ++a) acc[a] = X[o * LDA + a];
    col_dot(T1, 0, F, W1, F, o, acc);
#pragma unroll
    for (int a = 0; a < TA; ++a) T0[o * LDA + a] = acc[a];                // r
  }
  __syncthreads();
  load_tile_T(T1, ybar, a0, N, F);
  __syncthreads();
  for (int c = threadIdx.x; c < F; c += NT) {
    float acc[TA];
#pragma unroll
    for (int a = 0; a < TA; ++a) acc[a] = 0.f;
    col_dot(T1, 0, F, W2T, F, c, acc);
#pragma unroll
    for (int a = 0; a < TA; ++a) {
      float s, ds;
      so3k_silu_d(T0[c * LDA + a], &s, &ds);
      T0[c * LDA + a] = acc[a] * ds;                                      // rbar (column c is read by this thread only)
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < F; c += NT) {
    float acc[TA];
#pragma unroll
    for (int a = 0; a < TA; ++a) acc[a] = 0.f;
    col_dot(T0, 0, F, W1T, F, c, acc);
#pragma unroll
    for (int a = 0; a < TA; ++a) {
      float s, ds;
      so3k_silu_d(H1[c * LDA + a], &s, &ds);
      H1[c * LDA + a] = acc[a] * ds;                                      // hbar1
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < F; c += NT) {
    float acc[TA];
#pragma unroll
    for (int a = 0; a < TA; ++a) acc[a] = 0.f;
    col_dot(H1, 0, F, W0T, F, c, acc);
#pragma unroll
    for (int a = 0; a < TA; ++a) {
      float s, ds;
      so3k_silu_d(X[c * LDA + a], &s, &ds);
      if (a0 + a < N) xbar[(size_t)(a0 + a) * F + c] = T0[c * LDA + a] + acc[a] * ds;
    }
  }
}

}  // namespace

void so3k_launch_embed(const So3kDims& D, int N, const int* z, const float* emb, float* x, float* chi,
                       int* dev_status, cudaStream_t st) {
  long long tot = (long long)N * (D.F > D.M ? D.F : D.M);
  SO3K_LAUNCH(k_embed, dim3(so3k_cdiv(tot, 256)), dim3(256), 0, st, D, N, z, emb, x, chi, dev_status);
}

void so3k_launch_node_proj(const So3kDims& D, int N, const float* x, const float* p, const LayerParams& lp,
                           float* proj, cudaStream_t st) {
  size_t smem = sizeof(float) * LDA * D.F;
  SO3K_SET_MAX_SMEM(k_node_proj, smem);
  SO3K_LAUNCH(k_node_proj, dim3(so3k_cdiv(N, TA)), dim3(NT), smem, st, D, N, x, p + lp.fb.Wq, p + lp.gb.Wq,
              p + lp.fb.Wk, p + lp.gb.Wk, p + lp.fb.Wv, proj);
}

// dst[b][j][i] = src[b][i][j], n x n matrices
__global__ void k_transpose_sq(int batch, int n, const float* __restrict__ src, float* __restrict__ dst) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < batch * n * n) {
    const int b = t / (n * n), r = t % (n * n), i = r / n, j = r % n;
    dst[(size_t)b * n * n + (size_t)j * n + i] = src[t];
  }
}
static void transpose_sq(int batch, int n, const float* src, float* dst, cudaStream_t st) {
  SO3K_LAUNCH(k_transpose_sq, dim3(so3k_cdiv(batch * n * n, 256)), dim3(256), 0, st, batch, n, src, dst);
}

void so3k_launch_transpose_params(const So3kDims& D, const ModelLayout& ml, const float* p, float* pT, cudaStream_t st) {
  for (size_t t = 0; t < ml.layers.size(); ++t) {
    const LayerParams& lp = ml.layers[t];
    transpose_sq(D.H, D.dh, p + lp.fb.Wq, pT + lp.fb.Wq, st);
    transpose_sq(D.H, D.dh, p + lp.fb.Wk, pT + lp.fb.Wk, st);
    transpose_sq(D.H, D.dh, p + lp.fb.Wv, pT + lp.fb.Wv, st);
    transpose_sq(D.nl, D.dg, p + lp.gb.Wq, pT + lp.gb.Wq, st);
    transpose_sq(D.nl, D.dg, p + lp.gb.Wk, pT + lp.gb.Wk, st);
    transpose_sq(1, D.F + D.nl, p + lp.int_W, pT + lp.int_W, st);
    for (int r = 0; r < 2; ++r)
      if (ml.node_opts & (r ? SO3K_OPT_RES2 : SO3K_OPT_RES1))
        for (int k = 0; k < 3; ++k) transpose_sq(1, D.F, p + lp.res[r][k], pT + lp.res[r][k], st);
  }
  transpose_sq(1, D.F, p + ml.out_W1, pT + ml.out_W1, st);
}

void so3k_launch_node_proj_bwd(const So3kDims& D, int N, const float* gproj, const float* p, const LayerParams& lp,
                               float* xbar, cudaStream_t st) {
  size_t smem = sizeof(float) * LDA * 5 * D.F;
  SO3K_SET_MAX_SMEM(k_node_proj_bwd, smem);
  SO3K_LAUNCH(k_node_proj_bwd, dim3(so3k_cdiv(N, TA)), dim3(NT), smem, st, D, N, gproj, p + lp.fb.Wq, p + lp.gb.Wq,
              p + lp.fb.Wk, p + lp.gb.Wk, p + lp.fb.Wv, xbar);
}

void so3k_launch_interaction(const So3kDims& D, int N, const float* x1, const float* chi1, const float* p,
                             const LayerParams& lp, float* x2, float* chi2, float* bcoef, cudaStream_t st,
                             const float* xbase) {
  size_t smem = sizeof(float) * (LDA * (D.F + D.nl) + TA * (D.M + D.nl));
  SO3K_SET_MAX_SMEM(k_interaction, smem);
  SO3K_LAUNCH(k_interaction, dim3(so3k_cdiv(N, TA)), dim3(NT), smem, st, D, N, x1, chi1, p + lp.int_W, p + lp.int_b,
              xbase, x2, chi2, bcoef);
}

void so3k_launch_interaction_bwd(const So3kDims& D, int N, const float* chi1, const float* bcoef, const float* p,
                                 const LayerParams& lp, const float* xbar_out, const float* chibar_out, float* xbar1,
                                 float* chibar1, cudaStream_t st, bool skip) {
  size_t smem = sizeof(float) * (LDA * (D.F + D.nl) + TA * (2 * D.M + D.nl));
  SO3K_SET_MAX_SMEM(k_interaction_bwd, smem);
  SO3K_LAUNCH(k_interaction_bwd, dim3(so3k_cdiv(N, TA)), dim3(NT), smem, st, D, N, chi1, bcoef, p + lp.int_W,
              xbar_out, chibar_out, skip ? 1 : 0, xbar1, chibar1);
}

void so3k_launch_readout(const So3kDims& D, int N, const float* x, const int* z, const float* p, const float* pT,
                         const ModelLayout& ml, float out_scale, float* e_atom, float* xbar, cudaStream_t st) {
  size_t smem = sizeof(float) * (2 * LDA * D.F + TA * D.F + TA);
  SO3K_SET_MAX_SMEM(k_readout, smem);
  SO3K_LAUNCH(k_readout, dim3(so3k_cdiv(N, TA)), dim3(NT), smem, st, D, N, x, z, p + ml.out_W1, pT + ml.out_W1,
              p + ml.out_b1, p + ml.out_W2, p + ml.out_b2, p + ml.scale, p + ml.shift, out_scale, e_atom, xbar);
}

void so3k_launch_energy_sum(int B, int n, const float* e_atom, float offset, float* E, cudaStream_t st) {
  SO3K_LAUNCH(k_energy_sum, dim3(B), dim3(NT), 0, st, B, n, e_atom, offset, E);
}

void so3k_launch_layernorm(int N, int F, const float* x, const float* scale, const float* bias, float* y, cudaStream_t st) {
  SO3K_LAUNCH(k_layernorm, dim3(so3k_cdiv(N, 4)), dim3(128), 0, st, N, F, x, scale, bias, y);
}

void so3k_launch_layernorm_bwd(int N, int F, const float* x, const float* scale, const float* gy, const float* add,
                               float* out, cudaStream_t st) {
  SO3K_LAUNCH(k_layernorm_bwd, dim3(so3k_cdiv(N, 4)), dim3(128), 0, st, N, F, x, scale, gy, add, out);
}

void so3k_launch_resmlp(const So3kDims& D, int N, const float* x, const float* p, const size_t* W, float* y,
                        cudaStream_t st) {
  size_t smem = sizeof(float) * 3 * LDA * D.F;
  SO3K_SET_MAX_SMEM(k_resmlp, smem);
  SO3K_LAUNCH(k_resmlp, dim3(so3k_cdiv(N, TA)), dim3(NT), smem, st, D, N, x, p + W[0], p + W[1], p + W[2], y);
}

void so3k_launch_resmlp_bwd(const So3kDims& D, int N, const float* x, const float* p, const float* pT, const size_t* W,
                            const float* ybar, float* xbar, cudaStream_t st) {
  size_t smem = sizeof(float) * 4 * LDA * D.F;
  SO3K_SET_MAX_SMEM(k_resmlp_bwd, smem);
  SO3K_LAUNCH(k_resmlp_bwd, dim3(so3k_cdiv(N, TA)), dim3(NT), smem, st, D, N, x, p + W[0], p + W[1], pT + W[0],
              pT + W[1], pT + W[2], ybar, xbar);
}

// ---- halo rows of the domain decomposition (so3k_dd_pack / so3k_dd_unpack) -------------------------------------------------
// pack: out[r][0 .. wa + wb) = [a[send_idx[r]][0 .. wa) | b[send_idx[r]][0 .. wb)] -- the owned rows a peer needs, both buffers
// of an exchange in one message row; unpack: a[n_own + r] = in[r][0 .. wa), b[n_own + r] = in[r][wa ..) -- the ghost rows.
namespace {
__global__ void k_dd_pack(int n_rows, int wa, int wb, const int* __restrict__ send_idx, const float* __restrict__ a,
                          const float* __restrict__ b, float* __restrict__ out) {
  const int w = wa + wb;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_rows * w) return;
  const int r = (int)(idx / w), c = (int)(idx % w);
  const size_t src = (size_t)send_idx[r];
  out[idx] = c < wa ? a[src * wa + c] : b[src * wb + (c - wa)];
}
__global__ void k_dd_unpack(int n_rows, int n_own, int wa, int wb, const float* __restrict__ in, float* __restrict__ a,
                            float* __restrict__ b) {
  const int w = wa + wb;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)n_rows * w) return;
  const int r = (int)(idx / w), c = (int)(idx % w);
  if (c < wa) a[(size_t)(n_own + r) * wa + c] = in[idx];
  else b[(size_t)(n_own + r) * wb + (c - wa)] = in[idx];
}
}  // namespace

void so3k_launch_dd_pack(int n_rows, int wa, int wb, const int* send_idx, const float* a, const float* b, float* out,
                         cudaStream_t st) {
  if (n_rows <= 0) return;
  SO3K_LAUNCH(k_dd_pack, dim3(so3k_cdiv((long long)n_rows * (wa + wb), 256)), dim3(256), 0, st, n_rows, wa, wb, send_idx, a, b, out);
}
void so3k_launch_dd_unpack(int n_rows, int n_own, int wa, int wb, const float* in, float* a, float* b, cudaStream_t st) {
  if (n_rows <= 0) return;
  SO3K_LAUNCH(k_dd_unpack, dim3(so3k_cdiv((long long)n_rows * (wa + wb), 256)), dim3(256), 0, st, n_rows, n_own, wa, wb, in, a, b);
}
