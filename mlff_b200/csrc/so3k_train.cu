// Training seam: cotangents of (E, F) -> parameter gradients (the reference's value_and_grad of a loss that
// contains the force, mlff/training/run.py:34-50, loss.py:42-73, observable_function.py:127-149).
//
// Formulation.  With F = -dE/dR the parameter gradient of a loss l(E, F) is
//     dl/dtheta = sum_b Ebar_b dE_b/dtheta  -  d/deps [ dE_tot/dtheta (R + eps Fbar) ]_(eps = 0) ,
// i.e. the FORWARD-mode derivative, along the position direction Rdot = -Fbar, of the reverse-mode parameter gradient.
// Both are obtained by ONE sweep "forward pass + reverse pass with parameter gradients" carried out in dual numbers
// (value, tangent): positions enter as (R, -Fbar), the read-out seed as (1, Ebar_b); the tangent component of the
// accumulated parameter gradient is dl/dtheta.  Parameters carry no tangent, so every dense contraction with a weight
// matrix is an ordinary fp32 GEMM over the two planes stacked along the row axis, and every weight gradient
// x0^T ybar1 + x1^T ybar0 is one GEMM whose reduction runs over both planes (with the planes of ybar swapped).  The
// kernels are templates over the scalar type: T = float gives the plain first-order dE/dtheta (no Fbar).
//
// Arrays are plane-major: [plane][row][col], plane stride = rows * cols, rows = atoms N or edge slots Pt (sorted CSR
// order of so3k_graph.cu; rows >= n_valid are zero in every cotangent array, so reductions over rows may run over all
// slots).  Reference rows: a5-a19 of SURVEY.md section 8 (embed.py:66-169, so3krates_layer.py:57-178, 449-465, 526-608,
// observable.py:62-93) plus the optional branches: LayerNorm pre / post (so3krates_layer.py:103-106, 153-156, 170-174),
// ResidualMLP 1 / 2 (mlp.py:30-72) and the ZBL repulsion with its ten scalars (observable.py:110-183).
#include "so3k_internal.h"
#include "so3k_edge.h"

namespace {

// ---- forward-mode dual numbers ---------------------------------------------------------------------------------
struct Dual { float p, t; };
SO3K_HD inline Dual mkd(float p, float t) { Dual r; r.p = p; r.t = t; return r; }
SO3K_HD inline Dual operator+(Dual a, Dual b) { return mkd(a.p + b.p, a.t + b.t); }
SO3K_HD inline Dual operator-(Dual a, Dual b) { return mkd(a.p - b.p, a.t - b.t); }
SO3K_HD inline Dual operator*(Dual a, Dual b) { return mkd(a.p * b.p, a.p * b.t + a.t * b.p); }
SO3K_HD inline Dual operator/(Dual a, Dual b) { float q = a.p / b.p; return mkd(q, (a.t - q * b.t) / b.p); }
SO3K_HD inline Dual operator+(Dual a, float b) { return mkd(a.p + b, a.t); }
SO3K_HD inline Dual operator+(float a, Dual b) { return mkd(a + b.p, b.t); }
SO3K_HD inline Dual operator-(Dual a, float b) { return mkd(a.p - b, a.t); }
SO3K_HD inline Dual operator-(float a, Dual b) { return mkd(a - b.p, -b.t); }
SO3K_HD inline Dual operator*(Dual a, float b) { return mkd(a.p * b, a.t * b); }
SO3K_HD inline Dual operator*(float a, Dual b) { return mkd(a * b.p, a * b.t); }
SO3K_HD inline Dual operator/(float a, Dual b) { float q = a / b.p; return mkd(q, -q * b.t / b.p); }
SO3K_HD inline Dual operator-(Dual a) { return mkd(-a.p, -a.t); }
SO3K_HD inline Dual& operator+=(Dual& a, Dual b) { a.p += b.p; a.t += b.t; return a; }

SO3K_HD inline float t_exp(float a) { return expf(a); }
SO3K_HD inline Dual t_exp(Dual a) { float e = expf(a.p); return mkd(e, e * a.t); }
SO3K_HD inline float t_sqrt(float a) { return sqrtf(a); }
SO3K_HD inline Dual t_sqrt(Dual a) { float s = sqrtf(a.p); return mkd(s, s > 0.f ? 0.5f * a.t / s : 0.f); }
SO3K_HD inline float t_cos(float a) { return cosf(a); }
SO3K_HD inline Dual t_cos(Dual a) { return mkd(cosf(a.p), -sinf(a.p) * a.t); }
SO3K_HD inline float t_max0(float a) { return a > 0.f ? a : 0.f; }
SO3K_HD inline Dual t_max0(Dual a) { return a.p > 0.f ? a : mkd(0.f, 0.f); }
SO3K_HD inline float prim(float a) { return a; }
SO3K_HD inline float prim(Dual a) { return a.p; }
// the component that is written into the parameter-gradient blob
SO3K_HD inline float gout(float a) { return a; }
SO3K_HD inline float gout(Dual a) { return a.t; }

template <class T> struct Tr;
template <> struct Tr<float> {
  static constexpr int planes = 1;
  SO3K_HD static inline float zero() { return 0.f; }
  SO3K_HD static inline float ld(const float* p, size_t i, size_t) { return p[i]; }
  SO3K_HD static inline void st(float* p, size_t i, size_t, float v) { p[i] = v; }
  SO3K_HD static inline float make(float p, float) { return p; }
};
template <> struct Tr<Dual> {
  static constexpr int planes = 2;
  SO3K_HD static inline Dual zero() { return mkd(0.f, 0.f); }
  SO3K_HD static inline Dual ld(const float* p, size_t i, size_t ps) { return mkd(p[i], p[i + ps]); }
  SO3K_HD static inline void st(float* p, size_t i, size_t ps, Dual v) { p[i] = v.p; p[i + ps] = v.t; }
  SO3K_HD static inline Dual make(float p, float t) { return mkd(p, t); }
};

template <class T> SO3K_HD inline T t_sigmoid(T a) { return 1.0f / (1.0f + t_exp(-a)); }
template <class T> SO3K_HD inline T t_silu(T a) { return a * t_sigmoid(a); }
template <class T> SO3K_HD inline T t_silu_prime(T a) { T s = t_sigmoid(a); return s * (1.0f + a * (1.0f - s)); }

__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ Dual warp_sum(Dual v) { return mkd(warp_sum(v.p), warp_sum(v.t)); }

constexpr int TB = 256;
constexpr int TN_PARTS = 296;     // row chunks of a weight-gradient reduction (2 per SM)

// ---- geometry: r, d, u, phi, Y_lm, rbf per sorted edge (embed.py:88-127, radial.py:163-166, cutoff radial.py:50-51) ----
template <class T>
__global__ void k_tr_geom(So3kDims D, int N, int Pt, int P, const int* __restrict__ rowptr, const int* __restrict__ row,
                          const int* __restrict__ col, const int* __restrict__ src, const float* __restrict__ R,
                          const float* __restrict__ Rdot, const float* __restrict__ strain_dot /* (B,3,3) or null */,
                          const float* __restrict__ cell, const int* __restrict__ cell_offset, float* __restrict__ phi,
                          float* __restrict__ Y, float* __restrict__ rbf, float* __restrict__ dist) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Pt) return;
  const size_t ps1 = (size_t)Pt, psM = (size_t)Pt * D.M, psK = (size_t)Pt * D.K;
  if (e >= rowptr[N]) {
    Tr<T>::st(phi, e, ps1, Tr<T>::zero());
    Tr<T>::st(dist, e, ps1, Tr<T>::zero());
    for (int m = 0; m < D.M; ++m) Tr<T>::st(Y, (size_t)e * D.M + m, psM, Tr<T>::zero());
    for (int k = 0; k < D.K; ++k) Tr<T>::st(rbf, (size_t)e * D.K + k, psK, Tr<T>::zero());
    return;
  }
  const int i = row[e], j = col[e], s = src[e];
  T r[3];
  for (int c = 0; c < 3; ++c) {
    float rp = R[(size_t)j * 3 + c] - R[(size_t)i * 3 + c];
    if (cell && cell_offset) {
      const float* cb = cell + (size_t)(s / P) * 9;
      const int* o = cell_offset + (size_t)s * 3;
      rp += (float)o[0] * cb[0 + c] + (float)o[1] * cb[3 + c] + (float)o[2] * cb[6 + c];
    }
    r[c] = Tr<T>::make(rp, Rdot ? Rdot[(size_t)j * 3 + c] - Rdot[(size_t)i * 3 + c] : 0.f);
  }
  if (Tr<T>::planes == 2 && strain_dot) {
    // stress cotangent: positions and cell are strained by the symmetrised eps (observable_function.py:172-174), so the edge
    // vector moves as r -> (1 + eps_s) r: tangent rdot += sym(Sbar_b) r
    const float* sb = strain_dot + (size_t)(s / P) * 9;
    const float rp[3] = {prim(r[0]), prim(r[1]), prim(r[2])};
    for (int c = 0; c < 3; ++c) {
      float tdot = 0.f;
      for (int dd = 0; dd < 3; ++dd) tdot += 0.5f * (sb[3 * c + dd] + sb[3 * dd + c]) * rp[dd];
      r[c] = r[c] + Tr<T>::make(0.f, tdot);
    }
  }
  T d = t_sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  T u[3];
  if (prim(d) > 0.f) { for (int c = 0; c < 3; ++c) u[c] = r[c] / d; }
  else { for (int c = 0; c < 3; ++c) u[c] = Tr<T>::zero(); }
  T ph = Tr<T>::zero();
  if (prim(d) < D.r_cut) ph = 0.5f * (t_cos(d * (3.14159265358979323846f / D.r_cut)) + 1.0f);
  Tr<T>::st(phi, e, ps1, ph);
  Tr<T>::st(dist, e, ps1, d);
  T yl[9];
  for (int sgm = 0; sgm < D.nl; ++sgm) {
    const int l = D.deg[sgm];
    so3k_sph_g<T>(l, u[0], u[1], u[2], yl);
    for (int m = 0; m < 2 * l + 1; ++m) Tr<T>::st(Y, (size_t)e * D.M + D.seg_off[sgm] + m, psM, yl[m]);
  }
  T ed = t_exp(-d);
  for (int k = 0; k < D.K; ++k) {
    T t = ed - (D.mu0 + (float)k * D.dmu);
    Tr<T>::st(rbf, (size_t)e * D.K + k, psK, t_exp(-(D.gamma * (t * t))));
  }
}

// x0 = Embed[z] * point_mask (embed.py:227-229); tangent plane and chi0 are zeroed by the caller
__global__ void k_tr_embed(int N, int F, int nemb, const int* __restrict__ z, const float* __restrict__ emb,
                           float* __restrict__ x, int* __restrict__ status) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * F) return;
  const int n = (int)(idx / F), f = (int)(idx % F);
  const int zz = z[n];
  float v = 0.f;
  if (zz > 0 && zz < nemb) v = emb[(size_t)zz * F + f];
  else if (zz != 0 && f == 0 && status) atomicOr(status, SO3K_STATUS_BAD_Z);
  x[idx] = v;
}

// g_el = c_l sum_{m in l} (chi_j - chi_i)_m^2   (so3krates_layer.py:89-93, contract.py:181-188)
template <class T>
__global__ void k_tr_gdiff(So3kDims D, int N, int Pt, const int* __restrict__ rowptr, const int* __restrict__ row,
                           const int* __restrict__ col, const float* __restrict__ chi, float* __restrict__ g) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Pt) return;
  const size_t psc = (size_t)N * D.M, psg = (size_t)Pt * D.nl;
  T acc[SO3K_MAXDEG];
  for (int l = 0; l < D.nl; ++l) acc[l] = Tr<T>::zero();
  if (e < rowptr[N]) {
    const int i = row[e], j = col[e];
    for (int s = 0; s < D.nl; ++s)
      for (int m = D.seg_off[s]; m < D.seg_off[s + 1]; ++m) {
        T c = Tr<T>::ld(chi, (size_t)j * D.M + m, psc) - Tr<T>::ld(chi, (size_t)i * D.M + m, psc);
        acc[s] += D.m_cg[m] * (c * c);
      }
  }
  for (int l = 0; l < D.nl; ++l) Tr<T>::st(g, (size_t)e * D.nl + l, psg, acc[l]);
}

// spherical branch, first layer: As = g Ws1 + bs1 (P, Fs) and Hs = silu(As)  (so3krates_layer.py:459-463)
template <class T>
__global__ void k_tr_sph1(int Pt, int nl, int Fs, const float* __restrict__ g, const float* __restrict__ Ws1,
                          const float* __restrict__ bs1, float* __restrict__ As, float* __restrict__ Hs) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)Pt * Fs) return;
  const int e = (int)(idx / Fs), c = (int)(idx % Fs);
  const size_t psg = (size_t)Pt * nl, psa = (size_t)Pt * Fs;
  T a = Tr<T>::make(bs1[c], 0.f);
  for (int l = 0; l < nl; ++l) a += Tr<T>::ld(g, (size_t)e * nl + l, psg) * Ws1[l * Fs + c];
  if (As) Tr<T>::st(As, idx, psa, a);
  Tr<T>::st(Hs, idx, psa, t_silu(a));
}

template <class T>
__global__ void k_tr_silu(size_t count, const float* __restrict__ A, float* __restrict__ H) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  Tr<T>::st(H, idx, count, t_silu(Tr<T>::ld(A, idx, count)));
}
// Abar = Hbar * silu'(A), in place on Hbar
template <class T>
__global__ void k_tr_silu_bwd(size_t count, const float* __restrict__ A, float* __restrict__ Hbar) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= count) return;
  Tr<T>::st(Hbar, idx, count, Tr<T>::ld(Hbar, idx, count) * t_silu_prime(Tr<T>::ld(A, idx, count)));
}

// per-head Dense without bias: y[n, h d + c] = sum_r x[n, h d + r] W[h, r, c]   (so3krates_layer.py:472-484, 568-608)
template <class T>
__global__ void k_tr_headproj(int N, int F, int d, const float* __restrict__ x, const float* __restrict__ W,
                              float* __restrict__ y) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * F) return;
  const int n = (int)(idx / F), f = (int)(idx % F), h = f / d, c = f % d;
  const size_t ps = (size_t)N * F;
  T acc = Tr<T>::zero();
  const float* Wh = W + (size_t)h * d * d;
  for (int r = 0; r < d; ++r) acc += Tr<T>::ld(x, (size_t)n * F + h * d + r, ps) * Wh[r * d + c];
  Tr<T>::st(y, idx, ps, acc);
}
// xbar[n, h d + r] += sum_c ybar[n, h d + c] W[h, r, c]
template <class T>
__global__ void k_tr_headproj_bwd_x(int N, int F, int d, const float* __restrict__ ybar, const float* __restrict__ W,
                                    float* __restrict__ xbar) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * F) return;
  const int n = (int)(idx / F), f = (int)(idx % F), h = f / d, r = f % d;
  const size_t ps = (size_t)N * F;
  T acc = Tr<T>::zero();
  const float* Wh = W + (size_t)h * d * d + (size_t)r * d;
  for (int c = 0; c < d; ++c) acc += Tr<T>::ld(ybar, (size_t)n * F + h * d + c, ps) * Wh[c];
  Tr<T>::st(xbar, idx, ps, Tr<T>::ld(xbar, idx, ps) + acc);
}
// attention coefficients alpha[e, a] = phi_e / sqrt(d_head) * sum_{f in head a} q_i w_e k_j   (so3krates_layer.py:583-586,
// 500-501, 551-554): heads 0..H-1 of the feature block, H..H+nl-1 of the geometric block.  One warp per edge.
template <class T>
__global__ void k_tr_alpha(So3kDims D, int N, int Pt, const int* __restrict__ rowptr, const int* __restrict__ row,
                           const int* __restrict__ col, const float* __restrict__ q, const float* __restrict__ k,
                           const float* __restrict__ qg, const float* __restrict__ kg, const float* __restrict__ wf,
                           const float* __restrict__ wg, const float* __restrict__ phi, float* __restrict__ alpha) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (e >= Pt) return;
  const int A = D.H + D.nl;
  const size_t psn = (size_t)N * D.F, pse = (size_t)Pt * D.F, psa = (size_t)Pt * A;
  if (e >= rowptr[N]) {
    if (lane < A) Tr<T>::st(alpha, (size_t)e * A + lane, psa, Tr<T>::zero());
    return;
  }
  const int i = row[e], j = col[e];
  const T ph = Tr<T>::ld(phi, e, (size_t)Pt);
  for (int a = 0; a < A; ++a) {
    const bool fb = a < D.H;
    const int dd = fb ? D.dh : D.dg, f0 = (fb ? a : a - D.H) * dd;
    const float* qq = fb ? q : qg; const float* kk = fb ? k : kg; const float* ww = fb ? wf : wg;
    T acc = Tr<T>::zero();
    for (int c = lane; c < dd; c += 32) {
      const int f = f0 + c;
      acc += Tr<T>::ld(qq, (size_t)i * D.F + f, psn) * Tr<T>::ld(ww, (size_t)e * D.F + f, pse) *
             Tr<T>::ld(kk, (size_t)j * D.F + f, psn);
    }
    acc = warp_sum(acc);
    if (lane == 0) Tr<T>::st(alpha, (size_t)e * A + a, psa, acc * ph * (1.0f / sqrtf((float)dd)));
  }
}

// x1 = x + sum_e alpha_fb v_j ; chi1 = chi + sum_e alpha_gb Y_e   (so3krates_layer.py:607-608, 563-564, 146-147): warp per receiver
template <class T>
__global__ void k_tr_aggregate(So3kDims D, int N, int Pt, const int* __restrict__ rowptr, const int* __restrict__ col,
                               const float* __restrict__ x, const float* __restrict__ chi, const float* __restrict__ v,
                               const float* __restrict__ alpha, const float* __restrict__ Y, float* __restrict__ x1,
                               float* __restrict__ chi1) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= N) return;
  const int A = D.H + D.nl;
  const size_t psn = (size_t)N * D.F, psm = (size_t)N * D.M, psa = (size_t)Pt * A, psy = (size_t)Pt * D.M;
  const int e0 = rowptr[i], e1 = rowptr[i + 1];
  for (int f = lane; f < D.F; f += 32) {
    const int h = f / D.dh;
    T acc = Tr<T>::ld(x, (size_t)i * D.F + f, psn);
    for (int e = e0; e < e1; ++e)
      acc += Tr<T>::ld(alpha, (size_t)e * A + h, psa) * Tr<T>::ld(v, (size_t)col[e] * D.F + f, psn);
    Tr<T>::st(x1, (size_t)i * D.F + f, psn, acc);
  }
  for (int m = lane; m < D.M; m += 32) {
    const int a = D.H + D.m_seg[m];
    T acc = Tr<T>::ld(chi, (size_t)i * D.M + m, psm);
    for (int e = e0; e < e1; ++e)
      acc += Tr<T>::ld(alpha, (size_t)e * A + a, psa) * Tr<T>::ld(Y, (size_t)e * D.M + m, psy);
    Tr<T>::st(chi1, (size_t)i * D.M + m, psm, acc);
  }
}

// cat = [x1 | l0(chi1)]   (so3krates_layer.py:355-379)
template <class T>
__global__ void k_tr_intpre(So3kDims D, int N, const float* __restrict__ x1, const float* __restrict__ chi1,
                            float* __restrict__ cat) {
  const int C = D.F + D.nl;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * C) return;
  const int n = (int)(idx / C), c = (int)(idx % C);
  const size_t psn = (size_t)N * D.F, psm = (size_t)N * D.M, psc = (size_t)N * C;
  T v;
  if (c < D.F) v = Tr<T>::ld(x1, (size_t)n * D.F + c, psn);
  else {
    const int s = c - D.F;
    v = Tr<T>::zero();
    for (int m = D.seg_off[s]; m < D.seg_off[s + 1]; ++m) {
      T t = Tr<T>::ld(chi1, (size_t)n * D.M + m, psm);
      v += D.m_cg[m] * (t * t);
    }
  }
  Tr<T>::st(cat, idx, psc, v);
}
// x2 = x1 + a ; chi2 = chi1 (1 + b_l)
template <class T>
__global__ void k_tr_intpost(So3kDims D, int N, const float* __restrict__ x1, const float* __restrict__ chi1,
                             const float* __restrict__ ab, float* __restrict__ x2, float* __restrict__ chi2) {
  const int C = D.F + D.M, Cab = D.F + D.nl;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * C) return;
  const int n = (int)(idx / C), c = (int)(idx % C);
  const size_t psn = (size_t)N * D.F, psm = (size_t)N * D.M, psab = (size_t)N * Cab;
  if (c < D.F) {
    Tr<T>::st(x2, (size_t)n * D.F + c, psn, Tr<T>::ld(x1, (size_t)n * D.F + c, psn) + Tr<T>::ld(ab, (size_t)n * Cab + c, psab));
  } else {
    const int m = c - D.F;
    T b = Tr<T>::ld(ab, (size_t)n * Cab + D.F + D.m_seg[m], psab);
    Tr<T>::st(chi2, (size_t)n * D.M + m, psm, Tr<T>::ld(chi1, (size_t)n * D.M + m, psm) * (1.0f + b));
  }
}
// abbar = [x2bar | sum_{m in l} chi2bar_m chi1_m]
template <class T>
__global__ void k_tr_intpost_bwd(So3kDims D, int N, const float* __restrict__ chi1, const float* __restrict__ x2bar,
                                 const float* __restrict__ chi2bar, float* __restrict__ abbar) {
  const int C = D.F + D.nl;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * C) return;
  const int n = (int)(idx / C), c = (int)(idx % C);
  const size_t psn = (size_t)N * D.F, psm = (size_t)N * D.M, psc = (size_t)N * C;
  T v;
  if (c < D.F) v = Tr<T>::ld(x2bar, (size_t)n * D.F + c, psn);
  else {
    const int s = c - D.F;
    v = Tr<T>::zero();
    for (int m = D.seg_off[s]; m < D.seg_off[s + 1]; ++m)
      v += Tr<T>::ld(chi2bar, (size_t)n * D.M + m, psm) * Tr<T>::ld(chi1, (size_t)n * D.M + m, psm);
  }
  Tr<T>::st(abbar, idx, psc, v);
}
// x1bar = x2bar + catbar[:, :F] ; chi1bar = chi2bar (1 + b_l) + catbar[:, F + l] 2 c_l chi1
template <class T>
__global__ void k_tr_intpre_bwd(So3kDims D, int N, const float* __restrict__ chi1, const float* __restrict__ ab,
                                const float* __restrict__ x2bar, const float* __restrict__ chi2bar,
                                const float* __restrict__ catbar, float* __restrict__ x1bar, float* __restrict__ chi1bar,
                                float* __restrict__ xnbar /* null, or (LayerNorm before the block): catbar[:, :F] goes here */) {
  const int C = D.F + D.M, Cab = D.F + D.nl;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * C) return;
  const int n = (int)(idx / C), c = (int)(idx % C);
  const size_t psn = (size_t)N * D.F, psm = (size_t)N * D.M, psab = (size_t)N * Cab;
  if (c < D.F) {
    if (xnbar) {
      Tr<T>::st(xnbar, (size_t)n * D.F + c, psn, Tr<T>::ld(catbar, (size_t)n * Cab + c, psab));
      Tr<T>::st(x1bar, (size_t)n * D.F + c, psn, Tr<T>::ld(x2bar, (size_t)n * D.F + c, psn));
    } else {
      Tr<T>::st(x1bar, (size_t)n * D.F + c, psn,
                Tr<T>::ld(x2bar, (size_t)n * D.F + c, psn) + Tr<T>::ld(catbar, (size_t)n * Cab + c, psab));
    }
  } else {
    const int m = c - D.F, s = D.m_seg[m];
    T b = Tr<T>::ld(ab, (size_t)n * Cab + D.F + s, psab);
    T gb = Tr<T>::ld(catbar, (size_t)n * Cab + D.F + s, psab);
    T c1 = Tr<T>::ld(chi1, (size_t)n * D.M + m, psm);
    Tr<T>::st(chi1bar, (size_t)n * D.M + m, psm,
              Tr<T>::ld(chi2bar, (size_t)n * D.M + m, psm) * (1.0f + b) + gb * (2.0f * D.m_cg[m]) * c1);
  }
}

// flax LayerNorm over the F features of a row (so3krates_layer.py:103-106, 153-156, 170-174; eps 1e-6, fast variance
// max(0, E[x^2] - E[x]^2), scale and bias).  One warp per atom row.
template <class T>
__global__ void k_tr_layernorm(int N, int F, const float* __restrict__ x, const float* __restrict__ scale,
                               const float* __restrict__ bias, float* __restrict__ y) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= N) return;
  const size_t ps = (size_t)N * F;
  T s1 = Tr<T>::zero(), s2 = Tr<T>::zero();
  for (int f = lane; f < F; f += 32) {
    T v = Tr<T>::ld(x, (size_t)n * F + f, ps);
    s1 += v;
    s2 += v * v;
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const float invF = 1.0f / (float)F;
  T mu = s1 * invF;
  T r = 1.0f / t_sqrt(t_max0(s2 * invF - mu * mu) + 1e-6f);
  for (int f = lane; f < F; f += 32)
    Tr<T>::st(y, (size_t)n * F + f, ps, (Tr<T>::ld(x, (size_t)n * F + f, ps) - mu) * r * scale[f] + bias[f]);
}
// out = (add ? add : 0) + dLayerNorm(x)^T ybar ; gs = ybar * xhat (its column sum is the scale gradient; the bias gradient is
// the column sum of ybar).  out may alias ybar or add.
template <class T>
__global__ void k_tr_layernorm_bwd(int N, int F, const float* __restrict__ x, const float* __restrict__ scale,
                                   const float* __restrict__ ybar, const float* __restrict__ add, float* __restrict__ out,
                                   float* __restrict__ gs) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= N) return;
  const size_t ps = (size_t)N * F;
  T s1 = Tr<T>::zero(), s2 = Tr<T>::zero();
  for (int f = lane; f < F; f += 32) {
    T v = Tr<T>::ld(x, (size_t)n * F + f, ps);
    s1 += v;
    s2 += v * v;
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  const float invF = 1.0f / (float)F;
  T mu = s1 * invF;
  T r = 1.0f / t_sqrt(t_max0(s2 * invF - mu * mu) + 1e-6f);
  T m1 = Tr<T>::zero(), m2 = Tr<T>::zero();
  for (int f = lane; f < F; f += 32) {
    T xh = (Tr<T>::ld(x, (size_t)n * F + f, ps) - mu) * r;
    T gh = Tr<T>::ld(ybar, (size_t)n * F + f, ps) * scale[f];
    m1 += gh;
    m2 += gh * xh;
  }
  m1 = warp_sum(m1) * invF;
  m2 = warp_sum(m2) * invF;
  for (int f = lane; f < F; f += 32) {
    const size_t i = (size_t)n * F + f;
    T xh = (Tr<T>::ld(x, i, ps) - mu) * r;
    T yb = Tr<T>::ld(ybar, i, ps);
    T v = r * (yb * scale[f] - m1 - xh * m2);
    if (add) v += Tr<T>::ld(add, i, ps);
    Tr<T>::st(gs, i, ps, yb * xh);
    Tr<T>::st(out, i, ps, v);
  }
}
// out = a + b (all planes)
template <class T>
__global__ void k_tr_add(size_t count, const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < count * Tr<T>::planes) out[idx] = a[idx] + b[idx];
}

// read-out (observable.py:80-83): e_i = (silu(a_i) . w2 + b2) scale[z] + shift[z], masked; and its reverse with the
// per-structure seed: abar = seed scale[z] w2 silu'(a) ; hs = seed scale[z] silu(a) (its column sum is w2bar) ;
// ebar_i = seed scale[z] (its sum is b2bar).  One warp per atom.
template <class T>
__global__ void k_tr_readout(int N, int F, int n_per, int nemb, const int* __restrict__ z, const float* __restrict__ a,
                             const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ scale,
                             const float* __restrict__ shift, const float* __restrict__ seed_t /* (B) or null */,
                             float seed_p, float* __restrict__ e_atom, float* __restrict__ abar, float* __restrict__ hs,
                             float* __restrict__ ebar) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (n >= N) return;
  const size_t psn = (size_t)N * F;
  const int zz = z[n];
  const bool ok = zz > 0 && zz < nemb;
  const float sc = ok ? scale[zz] : 0.f;
  // seed of structure b: T = Dual -> (1, Ebar_b) ; T = float -> Ebar_b (or 1)
  const int b = n / n_per;
  T seed = Tr<T>::planes == 2 ? Tr<T>::make(seed_p, seed_t ? seed_t[b] : 0.f)
                              : Tr<T>::make(seed_t ? seed_t[b] : seed_p, 0.f);
  T sb = seed * sc;
  T acc = Tr<T>::zero();
  for (int f = lane; f < F; f += 32) {
    T av = Tr<T>::ld(a, (size_t)n * F + f, psn);
    T s = t_silu(av);
    acc += s * w2[f];
    Tr<T>::st(abar, (size_t)n * F + f, psn, sb * w2[f] * t_silu_prime(av));
    Tr<T>::st(hs, (size_t)n * F + f, psn, sb * s);
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    T e = ok ? (acc + b2[0]) * sc + shift[zz] : Tr<T>::zero();
    Tr<T>::st(e_atom, n, (size_t)N, e);
    Tr<T>::st(ebar, n, (size_t)N, sb);
  }
}
__global__ void k_tr_energy_sum(int B, int n_per, const float* __restrict__ e_atom, float offset, float* __restrict__ E) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  double s = 0.0;
  for (int i = 0; i < n_per; ++i) s += (double)e_atom[(size_t)b * n_per + i];
  E[b] = (float)(s - (double)offset);            // per-structure ZBL shift (observable.py:83-86)
}

// ---- reverse sweep over the edges --------------------------------------------------------------------------------
// sbar[e, a] = alphabar[e, a] phi_e / sqrt(d_a) with alphabar_fb = <x1bar_i, v_j>_h, alphabar_gb = <chi1bar_i, Y_e>_l.  Warp per edge.
template <class T>
__global__ void k_tr_alpha_bwd(So3kDims D, int N, int Pt, const int* __restrict__ rowptr, const int* __restrict__ row,
                               const int* __restrict__ col, const float* __restrict__ x1bar,
                               const float* __restrict__ chi1bar, const float* __restrict__ v, const float* __restrict__ Y,
                               const float* __restrict__ phi, float* __restrict__ sbar) {
  const int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (e >= Pt) return;
  const int A = D.H + D.nl;
  const size_t psn = (size_t)N * D.F, psm = (size_t)N * D.M, psa = (size_t)Pt * A, psy = (size_t)Pt * D.M;
  if (e >= rowptr[N]) {
    if (lane < A) Tr<T>::st(sbar, (size_t)e * A + lane, psa, Tr<T>::zero());
    return;
  }
  const int i = row[e], j = col[e];
  const T ph = Tr<T>::ld(phi, e, (size_t)Pt);
  for (int h = 0; h < D.H; ++h) {
    T acc = Tr<T>::zero();
    for (int c = lane; c < D.dh; c += 32) {
      const int f = h * D.dh + c;
      acc += Tr<T>::ld(x1bar, (size_t)i * D.F + f, psn) * Tr<T>::ld(v, (size_t)j * D.F + f, psn);
    }
    acc = warp_sum(acc);
    if (lane == 0) Tr<T>::st(sbar, (size_t)e * A + h, psa, acc * ph * (1.0f / sqrtf((float)D.dh)));
  }
  for (int s = 0; s < D.nl; ++s) {
    T acc = Tr<T>::zero();
    for (int m = D.seg_off[s] + lane; m < D.seg_off[s + 1]; m += 32)
      acc += Tr<T>::ld(chi1bar, (size_t)i * D.M + m, psm) * Tr<T>::ld(Y, (size_t)e * D.M + m, psy);
    acc = warp_sum(acc);
    if (lane == 0) Tr<T>::st(sbar, (size_t)e * A + D.H + s, psa, acc * ph * (1.0f / sqrtf((float)D.dg)));
  }
}
// vbar_i = sum_{e in row i} alpha[rev e, h] x1bar[col e]  (the sender-indexed sum made receiver-indexed through rev)
template <class T>
__global__ void k_tr_vbar(So3kDims D, int N, int Pt, const int* __restrict__ rowptr, const int* __restrict__ col,
                          const int* __restrict__ rev, const float* __restrict__ alpha, const float* __restrict__ x1bar,
                          float* __restrict__ vbar) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= N) return;
  const int A = D.H + D.nl;
  const size_t psn = (size_t)N * D.F, psa = (size_t)Pt * A;
  const int e0 = rowptr[i], e1 = rowptr[i + 1];
  for (int f = lane; f < D.F; f += 32) {
    const int h = f / D.dh;
    T acc = Tr<T>::zero();
    for (int e = e0; e < e1; ++e)
      acc += Tr<T>::ld(alpha, (size_t)rev[e] * A + h, psa) * Tr<T>::ld(x1bar, (size_t)col[e] * D.F + f, psn);
    Tr<T>::st(vbar, (size_t)i * D.F + f, psn, acc);
  }
}
// wbar[e, f] = sbar[e, head(f)] q_i[f] k_j[f]   (blk 0: feature block heads, 1: geometric block heads)
template <class T>
__global__ void k_tr_wbar(So3kDims D, int N, int Pt, int blk, const int* __restrict__ rowptr, const int* __restrict__ row,
                          const int* __restrict__ col, const float* __restrict__ q, const float* __restrict__ k,
                          const float* __restrict__ sbar, float* __restrict__ wbar) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)Pt * D.F) return;
  const int e = (int)(idx / D.F), f = (int)(idx % D.F);
  const int A = D.H + D.nl;
  const size_t psn = (size_t)N * D.F, pse = (size_t)Pt * D.F, psa = (size_t)Pt * A;
  T v = Tr<T>::zero();
  if (e < rowptr[N]) {
    const int a = blk ? D.H + f / D.dg : f / D.dh;
    v = Tr<T>::ld(sbar, (size_t)e * A + a, psa) * Tr<T>::ld(q, (size_t)row[e] * D.F + f, psn) *
        Tr<T>::ld(k, (size_t)col[e] * D.F + f, psn);
  }
  Tr<T>::st(wbar, idx, pse, v);
}
// qbar_i = sum_e sbar_e w_e k_j ; kbar_i = sum_e sbar_(rev e) w_(rev e) q_j.  Warp per receiver.
template <class T>
__global__ void k_tr_qkbar(So3kDims D, int N, int Pt, int blk, const int* __restrict__ rowptr, const int* __restrict__ col,
                           const int* __restrict__ rev, const float* __restrict__ q, const float* __restrict__ k,
                           const float* __restrict__ w, const float* __restrict__ sbar, float* __restrict__ qbar,
                           float* __restrict__ kbar) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= N) return;
  const int A = D.H + D.nl;
  const size_t psn = (size_t)N * D.F, pse = (size_t)Pt * D.F, psa = (size_t)Pt * A;
  const int e0 = rowptr[i], e1 = rowptr[i + 1];
  for (int f = lane; f < D.F; f += 32) {
    const int a = blk ? D.H + f / D.dg : f / D.dh;
    T aq = Tr<T>::zero(), ak = Tr<T>::zero();
    for (int e = e0; e < e1; ++e) {
      const int j = col[e], r = rev[e];
      aq += Tr<T>::ld(sbar, (size_t)e * A + a, psa) * Tr<T>::ld(w, (size_t)e * D.F + f, pse) *
            Tr<T>::ld(k, (size_t)j * D.F + f, psn);
      ak += Tr<T>::ld(sbar, (size_t)r * A + a, psa) * Tr<T>::ld(w, (size_t)r * D.F + f, pse) *
            Tr<T>::ld(q, (size_t)j * D.F + f, psn);
    }
    Tr<T>::st(qbar, (size_t)i * D.F + f, psn, aq);
    Tr<T>::st(kbar, (size_t)i * D.F + f, psn, ak);
  }
}
// spherical branch reverse: Asbar = (wbar Ws2^T) silu'(As) is in `asbar`; gbar[e, l] (+)= sum_c asbar[e, c] Ws1[l, c]
template <class T>
__global__ void k_tr_sph1_bwd(int Pt, int nl, int Fs, int accumulate, const float* __restrict__ asbar,
                              const float* __restrict__ Ws1, float* __restrict__ gbar) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)Pt * nl) return;
  const int e = (int)(idx / nl), l = (int)(idx % nl);
  const size_t psa = (size_t)Pt * Fs, psg = (size_t)Pt * nl;
  T acc = accumulate ? Tr<T>::ld(gbar, idx, psg) : Tr<T>::zero();
  for (int c = 0; c < Fs; ++c) acc += Tr<T>::ld(asbar, (size_t)e * Fs + c, psa) * Ws1[l * Fs + c];
  Tr<T>::st(gbar, idx, psg, acc);
}
// chibar_in[i, m] = chi1bar[i, m] - 2 c_l sum_{e in row i} (chi_j - chi_i)_m (gbar[e, l] + gbar[rev e, l])
template <class T>
__global__ void k_tr_gdiff_bwd(So3kDims D, int N, int Pt, const int* __restrict__ rowptr, const int* __restrict__ col,
                               const int* __restrict__ rev, const float* __restrict__ chi, const float* __restrict__ gbar,
                               const float* __restrict__ chi1bar, float* __restrict__ chibar) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= N) return;
  const size_t psm = (size_t)N * D.M, psg = (size_t)Pt * D.nl;
  const int e0 = rowptr[i], e1 = rowptr[i + 1];
  for (int m = lane; m < D.M; m += 32) {
    const int s = D.m_seg[m];
    T ci = Tr<T>::ld(chi, (size_t)i * D.M + m, psm);
    T acc = Tr<T>::zero();
    for (int e = e0; e < e1; ++e) {
      T c = Tr<T>::ld(chi, (size_t)col[e] * D.M + m, psm) - ci;
      acc += c * (Tr<T>::ld(gbar, (size_t)e * D.nl + s, psg) + Tr<T>::ld(gbar, (size_t)rev[e] * D.nl + s, psg));
    }
    Tr<T>::st(chibar, (size_t)i * D.M + m, psm, Tr<T>::ld(chi1bar, (size_t)i * D.M + m, psm) - (2.0f * D.m_cg[m]) * acc);
  }
}
// embedding gradient: Ebar[z_n, f] += x0bar[n, f]
template <class T>
__global__ void k_tr_embed_bwd(int N, int F, int nemb, const int* __restrict__ z, const float* __restrict__ xbar,
                               float* __restrict__ gemb) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (size_t)N * F) return;
  const int n = (int)(idx / F), f = (int)(idx % F);
  const int zz = z[n];
  if (zz > 0 && zz < nemb) atomicAdd(&gemb[(size_t)zz * F + f], gout(Tr<T>::ld(xbar, idx, (size_t)N * F)));
}
__global__ void k_tr_axpy_neg(size_t count, const float* __restrict__ a, float* __restrict__ out) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx < count) out[idx] = -a[idx];
}

// ---- Ziegler-Biersack-Littmark repulsion (observable.py:110-183): energy and the gradients of its ten scalars ------------
// The ten raw scalars (softplus at use) enter through a second level of forward mode: D2<T> = (value, d / d theta_k) over T
// (T itself float, or Dual in the position direction), one evaluation of the pair energy per scalar k.  The tangent part,
// times the read-out seed, is d(seed E_rep)/d theta_k -- with T = Dual its tangent component is the mixed second derivative
// the loss gradient needs.
template <class T> struct D2 { T p, t; };
template <class T> SO3K_HD inline D2<T> mk2(T p, T t) { D2<T> r; r.p = p; r.t = t; return r; }
template <class T> SO3K_HD inline D2<T> operator+(D2<T> a, D2<T> b) { return mk2<T>(a.p + b.p, a.t + b.t); }
template <class T> SO3K_HD inline D2<T> operator*(D2<T> a, D2<T> b) { return mk2<T>(a.p * b.p, a.p * b.t + a.t * b.p); }
template <class T> SO3K_HD inline D2<T> operator/(D2<T> a, D2<T> b) { T q = a.p / b.p; return mk2<T>(q, (a.t - q * b.t) / b.p); }
template <class T> SO3K_HD inline D2<T> scale2(D2<T> a, T s) { return mk2<T>(a.p * s, a.t * s); }
template <class T> SO3K_HD inline D2<T> d2_exp(D2<T> a) { T e = t_exp(a.p); return mk2<T>(e, e * a.t); }
template <class T> SO3K_HD inline D2<T> neg2(D2<T> a) { return mk2<T>(Tr<T>::zero() - a.p, Tr<T>::zero() - a.t); }
// softplus(raw) of scalar k, seeded when k == which
template <class T> SO3K_HD inline D2<T> zbl_par(const float* raw, int k, int which) {
  const float x = raw[k];
  const float sp = x > 20.f ? x : log1pf(expf(x));
  const float sg = 1.0f / (1.0f + expf(-x));
  return mk2<T>(Tr<T>::make(sp, 0.f), Tr<T>::make(k == which ? sg : 0.f, 0.f));
}
// sigma(u) = exp(-1 / u) for u > 0, else 0 (observable.py:17-18)
template <class T> SO3K_HD inline T zbl_sigma_t(T u) { return prim(u) > 0.f ? t_exp(Tr<T>::zero() - 1.0f / u) : Tr<T>::zero(); }

// e_atom[i] += sum_{e in row i} e_pair ; grad[zbl + k] += seed_b(i) * d e_pair / d theta_k.  One warp per receiver row.
template <class T>
__global__ void k_tr_zbl(So3kDims D, int N, int Pt, int n_per, const int* __restrict__ rowptr, const int* __restrict__ col,
                         const int* __restrict__ z, const float* __restrict__ raw, const float* __restrict__ dist,
                         const float* __restrict__ phi, const float* __restrict__ seed_t, float seed_p,
                         float* __restrict__ e_atom, float* __restrict__ gzbl) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= N) return;
  const float zi = (float)z[i];
  const int e0 = rowptr[i], e1 = rowptr[i + 1];
  const int b = i / n_per;
  const T seed = Tr<T>::planes == 2 ? Tr<T>::make(seed_p, seed_t ? seed_t[b] : 0.f) : Tr<T>::make(seed_t ? seed_t[b] : seed_p, 0.f);
  T esum = Tr<T>::zero();
  float gk[10];
#pragma unroll
  for (int k = 0; k < 10; ++k) gk[k] = 0.f;
  for (int e = e0 + lane; e < e1; e += 32) {
    const float zj = (float)z[col[e]];
    const T d = Tr<T>::ld(dist, e, (size_t)Pt), ph = Tr<T>::ld(phi, e, (size_t)Pt);
    if (!(prim(d) > 0.f && zi > 0.f && zj > 0.f)) continue;
    const T cc = d * (1.0f / 1.5f);
    const T sa = zbl_sigma_t<T>(1.0f - cc), sb = zbl_sigma_t<T>(cc);
    const T pre = (0.5f * 14.399645351950548f * zi * zj) * (sa / (sa + sb)) * ph / d;
    const float lzi = logf(zi), lzj = logf(zj);
    for (int which = -1; which < 10; ++which) {         // which = -1: the energy itself ; 0..9: d / d theta_which
      D2<T> a[4], c[4];
      for (int k = 0; k < 4; ++k) { a[k] = zbl_par<T>(raw, k, which); c[k] = zbl_par<T>(raw, 4 + k, which); }
      const D2<T> csum = c[0] + c[1] + c[2] + c[3];
      const D2<T> pw = zbl_par<T>(raw, 8, which), dd = zbl_par<T>(raw, 9, which);
      // s = (z_i^p + z_j^p) d_par ; y = sum_k (c_k / csum) exp(-a_k s d)
      const D2<T> s = (d2_exp<T>(scale2<T>(pw, Tr<T>::make(lzi, 0.f))) + d2_exp<T>(scale2<T>(pw, Tr<T>::make(lzj, 0.f)))) * dd;
      D2<T> y = mk2<T>(Tr<T>::zero(), Tr<T>::zero());
      for (int k = 0; k < 4; ++k) y = y + (c[k] / csum) * d2_exp<T>(neg2<T>(scale2<T>(a[k] * s, d)));
      if (which < 0) esum += pre * y.p;
      else gk[which] += gout(seed * (pre * y.t));
    }
  }
  esum = warp_sum(esum);
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    const float v = warp_sum(gk[k]);
    if (lane == 0 && v != 0.f) atomicAdd(&gzbl[k], v);
  }
  if (lane == 0) Tr<T>::st(e_atom, i, (size_t)N, Tr<T>::ld(e_atom, i, (size_t)N) + esum);
}

// ---- dense contractions (plain fp32 over the stacked planes) -------------------------------------------------------
// C[r, n] = (beta ? C[r, n] : 0) + sum_k A[r, k] B(k, n) + (r < bias_rows ? bias[n] : 0) ; B(k, n) = Bp[k sbk + n sbn]
// (sbk = ldb, sbn = 1: C = A B ; sbk = 1, sbn = ldb: C = A B^T).  128 x 64 tile, 256 threads, 8 x 4 outputs per thread.
constexpr int GM = 128, GN = 64, GK = 16;
__global__ void __launch_bounds__(256) k_tr_gemm(int Mr, int Nc, int Kd, const float* __restrict__ A, int lda,
                                                 const float* __restrict__ Bp, int sbk, int sbn,
                                                 const float* __restrict__ bias, int bias_rows, int beta,
                                                 float* __restrict__ C, int ldc) {
  SO3K_SHARED16 float As[GK][GM + 4];
  SO3K_SHARED16 float Bs[GK][GN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;        // tx: 4 columns, ty: 8 rows
  const int r0 = blockIdx.y * GM, c0 = blockIdx.x * GN;
  float acc[8][4];
  for (int a = 0; a < 8; ++a) for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  // the next K slab is fetched into registers while the current one is multiplied out of shared memory
  float ra[8], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int t = tid + 256 * q, r = t / GK, kk = t % GK;
      ra[q] = (r0 + r < Mr && k0 + kk < Kd) ? A[(size_t)(r0 + r) * lda + k0 + kk] : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int t = tid + 256 * q;
      const int kk = sbn == 1 ? t / GN : t % GK, c = sbn == 1 ? t % GN : t / GK;
      rb[q] = (k0 + kk < Kd && c0 + c < Nc) ? Bp[(size_t)(k0 + kk) * sbk + (size_t)(c0 + c) * sbn] : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < Kd; k0 += GK) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int t = tid + 256 * q;
      As[t % GK][t / GK] = ra[q];
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int t = tid + 256 * q;
      Bs[sbn == 1 ? t / GN : t % GK][sbn == 1 ? t % GN : t / GK] = rb[q];
    }
    __syncthreads();
    if (k0 + GK < Kd) fetch(k0 + GK);
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
  for (int a = 0; a < 8; ++a) {
    const int r = r0 + ty * 8 + a;
    if (r >= Mr) continue;
    for (int b = 0; b < 4; ++b) {
      const int c = c0 + tx * 4 + b;
      if (c >= Nc) continue;
      float v = acc[a][b];
      if (bias && r < bias_rows) v += bias[c];
      if (beta) v += C[(size_t)r * ldc + c];
      C[(size_t)r * ldc + c] = v;
    }
  }
}
// Weight gradient: G[k, n] += sum_r A'[r, k] B'[r, n] over the stacked rows r < planes R, where plane p of B' is plane
// (planes - 1 - p) of B (x0^T ybar1 + x1^T ybar0; one plane: x^T ybar).  A block owns a 144 x 144 output tile (the whole
// matrix for F = 132) and a chunk of rows: 9 x 9 outputs per thread, 18 shared-memory loads per 81 FMAs; the per-chunk
// partial sums go to a scratch array and k_tr_reduce_parts adds them onto the gradient in a fixed order (no atomics).
constexpr int WT = 144, WR = 8, WPT = 9;
template <int WPK>                // output rows per thread: the tile is 16 WPK rows (Kd) x 144 columns (Nc)
__global__ void __launch_bounds__(256) k_tr_gemm_tn(int R, int planes, int Kd, int Nc, int chunk,
                                                    const float* __restrict__ A, int lda, const float* __restrict__ B,
                                                    int ldb, float* __restrict__ part) {
  __shared__ float As[WR][16 * WPK];
  __shared__ float Bs[WR][WT];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  constexpr int WTK = 16 * WPK;
  const int k0 = blockIdx.y * WTK, c0 = blockIdx.x * WT;
  const int Rt = R * planes;                                   // stacked rows (fits an int: planes * edge slots)
  const int rbeg = blockIdx.z * chunk, rend = rbeg + chunk < Rt ? rbeg + chunk : Rt;
  float acc[WPK][WPT];
#pragma unroll
  for (int a = 0; a < WPK; ++a)
#pragma unroll
    for (int b = 0; b < WPT; ++b) acc[a][b] = 0.f;
  // 8 rows x 144 columns of both operands per step: 4.5 elements per thread, fetched a step ahead into registers
  constexpr int NLD = (WR * WT + 255) / 256;
  float ra[NLD], rb[NLD];
  auto fetch = [&](int rr) {
#pragma unroll
    for (int q = 0; q < NLD; ++q) {
      const int t = tid + 256 * q, r = t / WT, c = t - r * WT;
      const int gr = rr + r;
      float av = 0.f, bv = 0.f;
      if (t < WR * WT && gr < rend) {
        if (c < WTK && k0 + c < Kd) av = A[(size_t)gr * lda + k0 + c];
        if (c0 + c < Nc) {
          const int pb = gr >= R ? 1 : 0, rb_ = gr - pb * R;     // planes <= 2
          bv = B[(size_t)((planes - 1 - pb) * R + rb_) * ldb + c0 + c];
        }
      }
      ra[q] = av;
      rb[q] = bv;
    }
  };
  fetch(rbeg);
  for (int rr = rbeg; rr < rend; rr += WR) {
#pragma unroll
    for (int q = 0; q < NLD; ++q) {
      const int t = tid + 256 * q, r = t / WT, c = t - r * WT;
      if (t < WR * WT) {
        if (c < WTK) As[r][c] = ra[q];
        Bs[r][c] = rb[q];
      }
    }
    __syncthreads();
    if (rr + WR < rend) fetch(rr + WR);
#pragma unroll
    for (int r = 0; r < WR; ++r) {
      float av[WPK], bv[WPT];
#pragma unroll
      for (int a = 0; a < WPK; ++a) av[a] = As[r][ty + 16 * a];
#pragma unroll
      for (int b = 0; b < WPT; ++b) bv[b] = Bs[r][tx + 16 * b];
#pragma unroll
      for (int a = 0; a < WPK; ++a)
#pragma unroll
        for (int b = 0; b < WPT; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
  float* dst = part + (size_t)blockIdx.z * Kd * Nc;
#pragma unroll
  for (int a = 0; a < WPK; ++a) {
    const int k = k0 + ty + 16 * a;
    if (k >= Kd) continue;
#pragma unroll
    for (int b = 0; b < WPT; ++b) {
      const int c = c0 + tx + 16 * b;
      if (c < Nc) dst[(size_t)k * Nc + c] = acc[a][b];
    }
  }
}
// G[i] += sum_z part[z][i]  (i < count = Kd Nc of the call that filled `part`): 32 outputs x 8 partial-sum lanes per block, so
// that the nz (up to 296) loads of an output are eight independent streams; the lanes are combined in a fixed order
__global__ void __launch_bounds__(256) k_tr_reduce_parts(int count, int nz, const float* __restrict__ part, float* __restrict__ G) {
  __shared__ float sh[8][33];
  const int cx = threadIdx.x & 31, zy = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + cx;
  float acc = 0.f;
  if (i < count)
    for (int z = zy; z < nz; z += 8) acc += part[(size_t)z * count + i];
  sh[zy][cx] = acc;
  __syncthreads();
  if (zy == 0 && i < count) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += sh[k][cx];
    G[i] += v;
  }
}
// per-head blocks of a dense (F, F) product: Wbar[h, r, c] += Gd[h d + r, h d + c]
__global__ void k_tr_extract_heads(int F, int d, const float* __restrict__ Gd, float* __restrict__ Wbar) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;
  const int heads = F / d;
  if (o >= heads * d * d) return;
  const int h = o / (d * d), r = (o / d) % d, c = o % d;
  Wbar[o] += Gd[(size_t)(h * d + r) * F + h * d + c];
}
// out[c] += sum_r X[r, c] over rows [0, R) of ONE plane (the caller passes the plane that belongs in the gradient):
// 32 columns x 8 row lanes per block, `chunk` rows per block
__global__ void __launch_bounds__(256) k_tr_colsum(int R, int C, int chunk, const float* __restrict__ X, float* __restrict__ out) {
  __shared__ float sh[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  const int r0 = blockIdx.y * chunk, r1 = min(R, r0 + chunk);
  float acc = 0.f;
  if (c < C)
    for (int r = r0 + ry; r < r1; r += 8) acc += X[(size_t)r * C + c];
  sh[ry][cx] = acc;
  __syncthreads();
  if (ry == 0 && c < C) {
    float v = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) v += sh[k][cx];
    atomicAdd(&out[c], v);
  }
}

// ---- loss (training/loss.py:13-29, 42-73) --------------------------------------------------------------------------
// sums[0..3] += (sum of squared energy residuals, their count, sum of squared force residuals, their count) over the
// entries whose target is not NaN (and, for forces, whose atom is real: z != 0)
__global__ void k_loss_sums(int B, int n, const float* __restrict__ E, const float* __restrict__ F,
                            const float* __restrict__ E_true, const float* __restrict__ F_true,
                            const int* __restrict__ z, double* __restrict__ sums) {
  __shared__ double sh[4][TB / 32];
  double s[4] = {0.0, 0.0, 0.0, 0.0};
  const size_t total = (size_t)B * n * 3;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    if (idx < (size_t)B) {
      const float t = E_true[idx];
      if (t == t) { const double r = (double)t - (double)E[idx]; s[0] += r * r; s[1] += 1.0; }
    }
    const float t = F_true[idx];
    if (t == t && z[idx / 3] != 0) { const double r = (double)t - (double)F[idx]; s[2] += r * r; s[3] += 1.0; }
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int q = 0; q < 4; ++q) {
    double v = s[q];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh[q][w] = v;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int k = 0; k < TB / 32; ++k) v += sh[threadIdx.x][k];
    atomicAdd(&sums[threadIdx.x], v);
  }
}
// Ebar = w_E 2 (E - E_true) / n_E ; Fbar = w_F 2 (F - F_true) / n_F (0 where masked) ; loss[0..2] = (loss, mse_E, mse_F)
__global__ void k_loss_cotangents(int B, int n, const float* __restrict__ E, const float* __restrict__ F,
                                  const float* __restrict__ E_true, const float* __restrict__ F_true,
                                  const int* __restrict__ z, const double* __restrict__ sums, float w_E, float w_F,
                                  float* __restrict__ E_bar, float* __restrict__ F_bar, float* __restrict__ loss) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const double nE = sums[1], nF = sums[3];
  if (idx == 0 && loss) {
    const double lE = nE > 0 ? sums[0] / nE : 0.0, lF = nF > 0 ? sums[2] / nF : 0.0;
    loss[0] = (float)(w_E * lE + w_F * lF);
    loss[1] = (float)lE;
    loss[2] = (float)lF;
  }
  if (idx < (size_t)B) {
    const float t = E_true[idx];
    E_bar[idx] = (t == t && nE > 0) ? (float)(2.0 * w_E * ((double)E[idx] - (double)t) / nE) : 0.f;
  }
  if (idx < (size_t)B * n * 3) {
    const float t = F_true[idx];
    F_bar[idx] = (t == t && z[idx / 3] != 0 && nF > 0) ? (float)(2.0 * w_F * ((double)F[idx] - (double)t) / nF) : 0.f;
  }
}

// ---- optimiser (training/optimizer.py:64-97: clip_by_global_norm -> scale_by_adam -> add_decayed_weights -> scale(-lr)) ----
__global__ void k_sumsq(size_t n, const float* __restrict__ g, double* __restrict__ out, long long* __restrict__ step_dev) {
  __shared__ double sh[TB / 32];
  if (step_dev && blockIdx.x == 0 && threadIdx.x == 0) *step_dev += 1;       // the update count of k_adam (launched next)
  double s = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    s += (double)g[i] * (double)g[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double v = 0.0;
    for (int k = 0; k < TB / 32; ++k) v += sh[k];
    atomicAdd(out, v);
  }
}
__global__ void k_adam(size_t n, float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                       float* __restrict__ v, const unsigned char* __restrict__ frozen,
                       const unsigned char* __restrict__ decay_mask, const double* __restrict__ sumsq, float clip,
                       float lr, float b1, float b2, float eps, float eps_root, float wd, float bc1, float bc2,
                       const long long* __restrict__ step_dev) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (frozen && frozen[i]) return;
  if (step_dev) {                         // bias corrections from the device-resident update count (CUDA-graph replays)
    const float t = (float)*step_dev;
    bc1 = 1.0f - powf(b1, t);
    bc2 = 1.0f - powf(b2, t);
  }
  float gi = g[i];
  if (clip > 0.f) {
    const float nrm = (float)sqrt(*sumsq);
    if (nrm > clip) gi *= clip / nrm;      // optax.clip_by_global_norm
  }
  const float mi = b1 * m[i] + (1.0f - b1) * gi;
  const float vi = b2 * v[i] + (1.0f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  float u = (mi / bc1) / (sqrtf(vi / bc2 + eps_root) + eps);
  if (wd != 0.f && (!decay_mask || decay_mask[i])) u += wd * p[i];
  p[i] -= lr * u;
}

// ---- host side ---------------------------------------------------------------------------------------------------
inline size_t a256(size_t n) { return (n + 255) & ~(size_t)255; }

struct TrainWs {
  Graph g;
  int* scratch;
  float *Rdot;
  float *phi, *Y, *rbf, *dist;                 // geometry (T planes)
  float *x, *chi;                              // (L+1) states
  float *gl;                                   // per layer (Pt, nl)
  float *A1[2], *As[2], *w[2];                 // per layer and block
  float *proj;                                 // per layer: q, k, v, qg, kg
  float *alpha, *x1, *chi1, *ab;               // per layer
  float *aout, *e_atom, *ebar;
  // temporaries
  float *H, *Hs, *cat, *abbar, *catbar;
  float *xbar, *chibar, *x1bar, *chi1bar, *sbar, *wbar, *projbar, *gbar, *hsbar;
  float *part, *dense;                         // weight-gradient partial sums (TN_PARTS x C x C), dense (F, F) product
  unsigned char* bimg;                         // weight operand image of the tensor-core GEMM in flight
  // optional node-level branches (SO3K_OPT_*; null when absent), per layer: LayerNorm-1 output, states between the
  // blocks, ResidualMLP intermediates (a1 = silu(x) W0, r = x + silu(a1) W1) ; temporaries
  float *xn1, *x1b, *x2a, *x2b, *ra1[2], *rr[2];
  float *tA, *tB, *tC, *xnbar;
  size_t bytes;
};

TrainWs carve_train(const So3kDims& D, int N, int Pt, int planes, void* base, int opts = 0) {
  TrainWs w;
  char* p = (char*)base;
  auto take = [&](size_t bytes) { char* r = p; p += a256(bytes); return (void*)r; };
  const size_t F = D.F, M = D.M, nl = D.nl, L = D.L, K = D.K, Fs = D.Fs, A = D.H + D.nl, C = D.F + D.nl;
  const size_t Pe = Pt > 0 ? Pt : 1, pl = planes, fl = sizeof(float);
  w.g.N = N; w.g.Pt = Pt;
  w.g.rowptr = (int*)take(sizeof(int) * (N + 1));
  w.g.col = (int*)take(sizeof(int) * Pe);
  w.g.row = (int*)take(sizeof(int) * Pe);
  w.g.src = (int*)take(sizeof(int) * Pe);
  w.g.rev = (int*)take(sizeof(int) * Pe);
  w.g.canon = (int*)take(sizeof(int) * (Pe + 4));
  w.g.n_canon = w.g.canon + Pe;
  w.g.geom = (float4*)take(sizeof(float4) * Pe);
  w.g.pid = (int*)take(sizeof(int) * Pe);
  w.g.n_valid = w.g.rowptr + N;
  w.scratch = (int*)take(sizeof(int) * so3k_graph_ws_ints(N, Pt));
  w.Rdot = (float*)take(fl * N * 3);
  w.phi = (float*)take(fl * pl * Pe);
  w.Y = (float*)take(fl * pl * Pe * M);
  w.rbf = (float*)take(fl * pl * Pe * K);
  w.dist = (float*)take(fl * pl * Pe);
  w.x = (float*)take(fl * pl * (L + 1) * N * F);
  w.chi = (float*)take(fl * pl * (L + 1) * N * M);
  w.gl = (float*)take(fl * pl * L * Pe * nl);
  for (int b = 0; b < 2; ++b) {
    w.A1[b] = (float*)take(fl * pl * L * Pe * F);
    w.As[b] = (float*)take(fl * pl * L * Pe * Fs);
    w.w[b] = (float*)take(fl * pl * L * Pe * F);
  }
  w.proj = (float*)take(fl * pl * L * 5 * N * F);
  w.alpha = (float*)take(fl * pl * L * Pe * A);
  w.x1 = (float*)take(fl * pl * L * N * F);
  w.chi1 = (float*)take(fl * pl * L * N * M);
  w.ab = (float*)take(fl * pl * L * N * C);
  w.aout = (float*)take(fl * pl * N * F);
  w.e_atom = (float*)take(fl * pl * N);
  w.ebar = (float*)take(fl * pl * N);
  w.H = (float*)take(fl * pl * Pe * F);
  w.Hs = (float*)take(fl * pl * Pe * Fs);
  w.cat = (float*)take(fl * pl * N * C);
  w.abbar = (float*)take(fl * pl * N * C);
  w.catbar = (float*)take(fl * pl * N * C);
  w.xbar = (float*)take(fl * pl * N * F);
  w.chibar = (float*)take(fl * pl * N * M);
  w.x1bar = (float*)take(fl * pl * N * F);
  w.chi1bar = (float*)take(fl * pl * N * M);
  w.sbar = (float*)take(fl * pl * Pe * A);
  w.wbar = (float*)take(fl * pl * Pe * F);
  w.projbar = (float*)take(fl * pl * 5 * N * F);
  w.gbar = (float*)take(fl * pl * Pe * nl);
  w.hsbar = (float*)take(fl * pl * Pe * Fs);
  w.part = (float*)take(fl * (size_t)TN_PARTS * C * C);
  w.dense = (float*)take(fl * F * F);
  w.bimg = (unsigned char*)take(so3k_tc_gemm_scratch_bytes());
  const bool ln = opts & SO3K_OPT_LN, post = ln && (opts & SO3K_OPT_LN_POST), res1 = opts & SO3K_OPT_RES1, res2 = opts & SO3K_OPT_RES2;
  auto layer_nf = [&](bool on) { return on ? (float*)take(fl * pl * L * N * F) : (float*)nullptr; };
  w.xn1 = layer_nf(ln);
  w.x1b = layer_nf(res1);
  w.x2a = layer_nf(res2);
  w.x2b = layer_nf(post);
  w.ra1[0] = layer_nf(res1); w.rr[0] = layer_nf(res1);
  w.ra1[1] = layer_nf(res2); w.rr[1] = layer_nf(res2);
  const bool any = opts != 0;
  w.tA = any ? (float*)take(fl * pl * N * F) : nullptr;
  w.tB = any ? (float*)take(fl * pl * N * F) : nullptr;
  w.tC = any ? (float*)take(fl * pl * N * F) : nullptr;
  w.xnbar = any ? (float*)take(fl * pl * N * F) : nullptr;
  w.bytes = (size_t)(p - (char*)base);
  return w;
}

inline dim3 g1(size_t count) { return dim3((unsigned)((count + TB - 1) / TB)); }
inline dim3 gw(size_t warps) { return dim3((unsigned)((warps * 32 + TB - 1) / TB)); }

// With SO3K_FLAG_TENSOR the large dense layers of the sweep run on the tensor cores (so3k_tc_gemm: split-bf16 operands, fp32
// accumulation); shapes outside that kernel's limits, and the fp32 mode, use the SIMT GEMM below.
struct TcCtx { bool on; unsigned char* scratch; int* status; };
// C (rows, Nc) = A (rows, Kd) @ W (Kd, Nc) + bias on the first bias_rows rows
void gemm_nn(int rows, int Nc, int Kd, const float* A, const float* W, const float* bias, int bias_rows, int beta,
             float* C, cudaStream_t st, const TcCtx* tc = nullptr) {
  if (rows <= 0) return;
  if (tc && tc->on && !beta &&
      so3k_tc_gemm(rows, Nc, Kd, A, Kd, W, Nc, 1, bias, bias_rows, C, Nc, tc->scratch, tc->status, st))
    return;
  SO3K_LAUNCH(k_tr_gemm, dim3(so3k_cdiv(Nc, GN), so3k_cdiv(rows, GM)), dim3(256), 0, st, rows, Nc, Kd, A, Kd, W, Nc, 1,
              bias, bias_rows, beta, C, Nc);
}
// C (rows, Kd) = A (rows, Nc) @ W (Kd, Nc)^T
void gemm_nt(int rows, int Kd, int Nc, const float* A, const float* W, float* C, cudaStream_t st, const TcCtx* tc = nullptr) {
  if (rows <= 0) return;
  if (tc && tc->on && so3k_tc_gemm(rows, Kd, Nc, A, Nc, W, 1, Nc, nullptr, 0, C, Kd, tc->scratch, tc->status, st)) return;
  SO3K_LAUNCH(k_tr_gemm, dim3(so3k_cdiv(Kd, GN), so3k_cdiv(rows, GM)), dim3(256), 0, st, rows, Kd, Nc, A, Nc, W, 1, Nc,
              (const float*)nullptr, 0, 0, C, Kd);
}
// G (Kd, Nc) += X^T Ybar with the dual-number product rule; `part` = scratch of TN_PARTS * Kd * Nc floats
void gemm_tn(int R, int planes, int Kd, int Nc, const float* X, int ldx, const float* Ybar, int ldy, float* G, float* part,
             cudaStream_t st) {
  if (R <= 0) return;
  const long long Rt = (long long)R * planes;
  // output rows per thread: 1 (Kd <= 16), 2 (<= 32), 3 (<= 48) or 9 (tiles of 144 rows) -- a small Kd does not pay for a full tile
  const int wpk = Kd <= 16 ? 1 : Kd <= 32 ? 2 : Kd <= 48 ? 3 : 9;
  const int tiles = so3k_cdiv(Nc, WT) * so3k_cdiv(Kd, 16 * wpk);
  int nz = TN_PARTS / tiles;
  if (nz < 1) nz = 1;
  int chunk = (int)((Rt + nz - 1) / nz);
  chunk = (chunk + WR - 1) / WR * WR;
  if (chunk < 64) chunk = 64;
  nz = (int)((Rt + chunk - 1) / chunk);
  const dim3 grid(so3k_cdiv(Nc, WT), so3k_cdiv(Kd, 16 * wpk), nz);
  switch (wpk) {
    case 1: SO3K_LAUNCH(k_tr_gemm_tn<1>, grid, dim3(256), 0, st, R, planes, Kd, Nc, chunk, X, ldx, Ybar, ldy, part); break;
    case 2: SO3K_LAUNCH(k_tr_gemm_tn<2>, grid, dim3(256), 0, st, R, planes, Kd, Nc, chunk, X, ldx, Ybar, ldy, part); break;
    case 3: SO3K_LAUNCH(k_tr_gemm_tn<3>, grid, dim3(256), 0, st, R, planes, Kd, Nc, chunk, X, ldx, Ybar, ldy, part); break;
    default: SO3K_LAUNCH(k_tr_gemm_tn<9>, grid, dim3(256), 0, st, R, planes, Kd, Nc, chunk, X, ldx, Ybar, ldy, part); break;
  }
  SO3K_LAUNCH(k_tr_reduce_parts, dim3(so3k_cdiv((long long)Kd * Nc, 32)), dim3(256), 0, st, Kd * Nc, nz, part, G);
}
// out (C) += column sums of the gradient plane of X (R, C)
void colsum(int R, int C, int planes, const float* X, float* out, cudaStream_t st) {
  if (R <= 0) return;
  const int chunk = 256;
  SO3K_LAUNCH(k_tr_colsum, dim3(so3k_cdiv(C, 32), so3k_cdiv(R, chunk)), dim3(256), 0, st, R, C, chunk,
              X + (size_t)(planes - 1) * R * C, out);
}

template <class T>
int run_vjp(so3k_handle* h, int B, int n, int P, const float* R, const int* z, const int* idx_i, const int* idx_j,
            const float* cell, const int* cell_offset, const float* E_bar, const float* F_bar, const float* S_bar, float* grad,
            float* E_out, void* workspace, int* dev_status, cudaStream_t st) {
  const So3kDims& D = h->dims;
  const ModelLayout& ml = h->layout;
  const int N = B * n, Pt = B * P, pl = Tr<T>::planes;
  const size_t F = D.F, M = D.M, nl = D.nl, K = D.K, Fs = D.Fs, A = D.H + D.nl, C = D.F + D.nl;
  const size_t Pe = Pt > 0 ? Pt : 1;
  const float* prm = h->params;
  const int opts = ml.node_opts;
  const bool ln = opts & SO3K_OPT_LN, post = ln && (opts & SO3K_OPT_LN_POST), res1 = opts & SO3K_OPT_RES1, res2 = opts & SO3K_OPT_RES2;
  TrainWs w = carve_train(D, N, Pt, pl, workspace, opts);
  cudaMemsetAsync(grad, 0, sizeof(float) * ml.total, st);
  so3k_build_graph(D, B, n, P, R, idx_i, idx_j, cell, cell_offset, nullptr, nullptr, w.g, w.scratch, dev_status, st);
  const Graph& g = w.g;
  const float* Rdot = nullptr;
  if (pl == 2) {
    // position tangent = -Fbar
    cudaMemsetAsync(w.Rdot, 0, sizeof(float) * N * 3, st);
    if (F_bar) SO3K_LAUNCH(k_tr_axpy_neg, g1((size_t)N * 3), dim3(TB), 0, st, (size_t)N * 3, F_bar, w.Rdot);
    Rdot = w.Rdot;
  }
  SO3K_LAUNCH(k_tr_geom<T>, g1(Pe), dim3(TB), 0, st, D, N, Pt, P, g.rowptr, g.row, g.col, g.src, R, Rdot, pl == 2 ? S_bar : nullptr,
              cell, cell_offset, w.phi, w.Y, w.rbf, w.dist);
  // x0, chi0
  const size_t sx = (size_t)pl * N * F, sc = (size_t)pl * N * M;       // per-state strides
  cudaMemsetAsync(w.x, 0, sizeof(float) * sx, st);
  cudaMemsetAsync(w.chi, 0, sizeof(float) * sc, st);
  SO3K_LAUNCH(k_tr_embed, g1((size_t)N * F), dim3(TB), 0, st, N, (int)F, D.nemb, z, prm + ml.embed, w.x, dev_status);

  const int rowsE = pl * (int)Pe, rowsN = pl * N;
  const size_t NF = (size_t)N * F;
  const TcCtx tcx = {(h->cfg.flags & SO3K_FLAG_TENSOR) != 0, w.bimg, dev_status};
  const TcCtx* tc = &tcx;
  auto Lp = [&](float* base, size_t per, int t) { return base + (size_t)t * pl * per; };
  // ResidualMLP (mlp.py:30-72, bias-free): y = silu(x + silu(silu(x) W0) W1) W2 ; keeps a1 = silu(x) W0 and r = x + silu(a1) W1
  auto resmlp_fwd = [&](const size_t* Wk, const float* x, float* a1, float* rr, float* y) {
    SO3K_LAUNCH(k_tr_silu<T>, g1(NF), dim3(TB), 0, st, NF, x, w.tA);
    gemm_nn(rowsN, (int)F, (int)F, w.tA, prm + Wk[0], nullptr, 0, 0, a1, st);
    SO3K_LAUNCH(k_tr_silu<T>, g1(NF), dim3(TB), 0, st, NF, a1, w.tA);
    gemm_nn(rowsN, (int)F, (int)F, w.tA, prm + Wk[1], nullptr, 0, 0, w.tB, st);
    SO3K_LAUNCH(k_tr_add<T>, g1(NF * pl), dim3(TB), 0, st, NF, x, w.tB, rr);
    SO3K_LAUNCH(k_tr_silu<T>, g1(NF), dim3(TB), 0, st, NF, rr, w.tA);
    gemm_nn(rowsN, (int)F, (int)F, w.tA, prm + Wk[2], nullptr, 0, 0, y, st);
  };
  // reverse: xbar <- dResidualMLP(x)^T ybar (in place on `bar`), weight gradients into the blob
  auto resmlp_bwd = [&](const size_t* Wk, const float* x, const float* a1, const float* rr, float* bar) {
    SO3K_LAUNCH(k_tr_silu<T>, g1(NF), dim3(TB), 0, st, NF, rr, w.tA);                       // h2
    gemm_tn(N, pl, (int)F, (int)F, w.tA, (int)F, bar, (int)F, grad + Wk[2], w.part, st);
    gemm_nt(rowsN, (int)F, (int)F, bar, prm + Wk[2], w.tB, st);                              // h2bar
    SO3K_LAUNCH(k_tr_silu_bwd<T>, g1(NF), dim3(TB), 0, st, NF, rr, w.tB);                   // rbar
    SO3K_LAUNCH(k_tr_silu<T>, g1(NF), dim3(TB), 0, st, NF, a1, w.tA);                       // h1
    gemm_tn(N, pl, (int)F, (int)F, w.tA, (int)F, w.tB, (int)F, grad + Wk[1], w.part, st);
    gemm_nt(rowsN, (int)F, (int)F, w.tB, prm + Wk[1], w.tC, st);                             // h1bar
    SO3K_LAUNCH(k_tr_silu_bwd<T>, g1(NF), dim3(TB), 0, st, NF, a1, w.tC);                   // a1bar
    SO3K_LAUNCH(k_tr_silu<T>, g1(NF), dim3(TB), 0, st, NF, x, w.tA);                        // h0
    gemm_tn(N, pl, (int)F, (int)F, w.tA, (int)F, w.tC, (int)F, grad + Wk[0], w.part, st);
    gemm_nt(rowsN, (int)F, (int)F, w.tC, prm + Wk[0], w.tA, st);                             // h0bar
    SO3K_LAUNCH(k_tr_silu_bwd<T>, g1(NF), dim3(TB), 0, st, NF, x, w.tA);                    // h0bar silu'(x)
    SO3K_LAUNCH(k_tr_add<T>, g1(NF * pl), dim3(TB), 0, st, NF, w.tB, w.tA, bar);            // xbar = rbar + ...
  };
  // reverse of LayerNorm k of layer t: out = (add ? add : 0) + dLN(x)^T ybar, scale / bias gradients into the blob
  auto ln_bwd = [&](const LayerParams& lp, int k, const float* x, const float* ybar, const float* add, float* out) {
    colsum(N, (int)F, pl, ybar, grad + lp.ln_b[k], st);
    SO3K_LAUNCH(k_tr_layernorm_bwd<T>, gw(N), dim3(TB), 0, st, N, (int)F, x, prm + lp.ln_s[k], ybar, add, out, w.tC);
    colsum(N, (int)F, pl, w.tC, grad + lp.ln_s[k], st);
  };
  for (int t = 0; t < D.L; ++t) {
    const LayerParams& lp = ml.layers[t];
    float* xt = w.x + t * sx; float* ct = w.chi + t * sc;
    float* gl = Lp(w.gl, Pe * nl, t);
    SO3K_LAUNCH(k_tr_gdiff<T>, g1(Pe), dim3(TB), 0, st, D, N, Pt, g.rowptr, g.row, g.col, ct, gl);
    for (int b = 0; b < 2; ++b) {
      const BlockParams& bp = b ? lp.gb : lp.fb;
      float* A1 = Lp(w.A1[b], Pe * F, t); float* As = Lp(w.As[b], Pe * Fs, t); float* wf = Lp(w.w[b], Pe * F, t);
      gemm_nn(rowsE, (int)F, (int)K, w.rbf, prm + bp.rad_W1, prm + bp.rad_b1, (int)Pe, 0, A1, st, tc);
      SO3K_LAUNCH(k_tr_silu<T>, g1(Pe * F), dim3(TB), 0, st, Pe * F, A1, w.H);
      gemm_nn(rowsE, (int)F, (int)F, w.H, prm + bp.rad_W2, prm + bp.rad_b2, (int)Pe, 0, wf, st, tc);
      SO3K_LAUNCH(k_tr_sph1<T>, g1(Pe * Fs), dim3(TB), 0, st, Pt, (int)nl, (int)Fs, gl, prm + bp.sph_W1, prm + bp.sph_b1,
                  As, w.Hs);
      gemm_nn(rowsE, (int)F, (int)Fs, w.Hs, prm + bp.sph_W2, prm + bp.sph_b2, (int)Pe, 1, wf, st);
    }
    // the projections read LayerNorm 1 of x_t (so3krates_layer.py:103-106); the skip connection keeps x_t (:146)
    const float* xin = xt;
    if (ln) {
      float* xn1 = Lp(w.xn1, NF, t);
      SO3K_LAUNCH(k_tr_layernorm<T>, gw(N), dim3(TB), 0, st, N, (int)F, xt, prm + lp.ln_s[0], prm + lp.ln_b[0], xn1);
      xin = xn1;
    }
    float* pj = Lp(w.proj, 5 * NF, t);
    float* q = pj, *k = pj + pl * NF, *v = pj + 2 * pl * NF, *qg = pj + 3 * pl * NF, *kg = pj + 4 * pl * NF;
    SO3K_LAUNCH(k_tr_headproj<T>, g1(NF), dim3(TB), 0, st, N, (int)F, D.dh, xin, prm + lp.fb.Wq, q);
    SO3K_LAUNCH(k_tr_headproj<T>, g1(NF), dim3(TB), 0, st, N, (int)F, D.dh, xin, prm + lp.fb.Wk, k);
    SO3K_LAUNCH(k_tr_headproj<T>, g1(NF), dim3(TB), 0, st, N, (int)F, D.dh, xin, prm + lp.fb.Wv, v);
    SO3K_LAUNCH(k_tr_headproj<T>, g1(NF), dim3(TB), 0, st, N, (int)F, D.dg, xin, prm + lp.gb.Wq, qg);
    SO3K_LAUNCH(k_tr_headproj<T>, g1(NF), dim3(TB), 0, st, N, (int)F, D.dg, xin, prm + lp.gb.Wk, kg);
    float* al = Lp(w.alpha, Pe * A, t);
    SO3K_LAUNCH(k_tr_alpha<T>, gw(Pe), dim3(TB), 0, st, D, N, Pt, g.rowptr, g.row, g.col, q, k, qg, kg,
                Lp(w.w[0], Pe * F, t), Lp(w.w[1], Pe * F, t), w.phi, al);
    // states between the blocks; each aliases its successor when the branch is absent
    float* x1a = Lp(w.x1, NF, t);                                     // x + x_local
    float* x1b = res1 ? Lp(w.x1b, NF, t) : x1a;                       // first skip state
    float* x2b = post ? Lp(w.x2b, NF, t) : xt + sx;                   // layer output before the post LayerNorm
    float* x2a = res2 ? Lp(w.x2a, NF, t) : x2b;                       // second skip state before ResidualMLP 2
    float* c1 = Lp(w.chi1, (size_t)N * M, t); float* ab = Lp(w.ab, (size_t)N * C, t);
    SO3K_LAUNCH(k_tr_aggregate<T>, gw(N), dim3(TB), 0, st, D, N, Pt, g.rowptr, g.col, xt, ct, v, al, w.Y, x1a, c1);
    if (res1) resmlp_fwd(lp.res[0], x1a, Lp(w.ra1[0], NF, t), Lp(w.rr[0], NF, t), x1b);
    const float* xi2 = x1b;
    if (ln) {
      SO3K_LAUNCH(k_tr_layernorm<T>, gw(N), dim3(TB), 0, st, N, (int)F, x1b, prm + lp.ln_s[1], prm + lp.ln_b[1], w.xnbar);
      xi2 = w.xnbar;
    }
    SO3K_LAUNCH(k_tr_intpre<T>, g1((size_t)N * C), dim3(TB), 0, st, D, N, xi2, c1, w.cat);
    gemm_nn(rowsN, (int)C, (int)C, w.cat, prm + lp.int_W, prm + lp.int_b, N, 0, ab, st);
    SO3K_LAUNCH(k_tr_intpost<T>, g1((size_t)N * (F + M)), dim3(TB), 0, st, D, N, x1b, c1, ab, x2a, ct + sc);
    if (res2) resmlp_fwd(lp.res[1], x2a, Lp(w.ra1[1], NF, t), Lp(w.rr[1], NF, t), x2b);
    if (post) SO3K_LAUNCH(k_tr_layernorm<T>, gw(N), dim3(TB), 0, st, N, (int)F, x2b, prm + lp.ln_s[2], prm + lp.ln_b[2], xt + sx);
  }
  // read-out and its reverse
  float* xL = w.x + (size_t)D.L * sx;
  gemm_nn(rowsN, (int)F, (int)F, xL, prm + ml.out_W1, prm + ml.out_b1, N, 0, w.aout, st);
  SO3K_LAUNCH(k_tr_readout<T>, gw(N), dim3(TB), 0, st, N, (int)F, n, D.nemb, z, w.aout, prm + ml.out_W2, prm + ml.out_b2,
              prm + ml.scale, prm + ml.shift, E_bar, 1.0f, w.e_atom, w.x1bar /* abar */, w.cat /* hs (N,F) fits in (N,C) */,
              w.ebar);
  if (ml.has_zbl)       // Energy(zbl_repulsion=True): pair repulsion added to the atoms' energies, gradients of its ten scalars
    SO3K_LAUNCH(k_tr_zbl<T>, gw(N), dim3(TB), 0, st, D, N, Pt, n, g.rowptr, g.col, z, prm + ml.zbl, w.dist, w.phi, E_bar, 1.0f,
                w.e_atom, grad + ml.zbl);
  if (E_out) SO3K_LAUNCH(k_tr_energy_sum, g1(B), dim3(TB), 0, st, B, n, w.e_atom, ml.has_zbl ? h->cfg.zbl_shift : 0.f, E_out);
  colsum(N, (int)F, pl, w.cat, grad + ml.out_W2, st);
  colsum(N, 1, pl, w.ebar, grad + ml.out_b2, st);
  float* abar = w.x1bar;
  gemm_tn(N, pl, (int)F, (int)F, xL, (int)F, abar, (int)F, grad + ml.out_W1, w.part, st);
  colsum(N, (int)F, pl, abar, grad + ml.out_b1, st);
  gemm_nt(rowsN, (int)F, (int)F, abar, prm + ml.out_W1, w.xbar, st);
  cudaMemsetAsync(w.chibar, 0, sizeof(float) * sc, st);

  for (int t = D.L - 1; t >= 0; --t) {
    const LayerParams& lp = ml.layers[t];
    float* xt = w.x + t * sx; float* ct = w.chi + t * sc;
    float* x1a = Lp(w.x1, NF, t);
    float* x1b = res1 ? Lp(w.x1b, NF, t) : x1a;
    float* x2b = post ? Lp(w.x2b, NF, t) : xt + sx;
    float* x2a = res2 ? Lp(w.x2a, NF, t) : x2b;
    float* c1 = Lp(w.chi1, (size_t)N * M, t); float* ab = Lp(w.ab, (size_t)N * C, t);
    float* al = Lp(w.alpha, Pe * A, t);
    float* gl = Lp(w.gl, Pe * nl, t);
    float* pj = Lp(w.proj, 5 * NF, t);
    const size_t pn = (size_t)pl * NF;
    float* proj[5] = {pj, pj + pn, pj + 2 * pn, pj + 3 * pn, pj + 4 * pn};          // q k v qg kg
    float* pbar[5] = {w.projbar, w.projbar + pn, w.projbar + 2 * pn, w.projbar + 3 * pn, w.projbar + 4 * pn};
    // the optional branches above the interaction block, last to first (in place on xbar)
    if (post) ln_bwd(lp, 2, x2b, w.xbar, nullptr, w.xbar);
    if (res2) resmlp_bwd(lp.res[1], x2a, Lp(w.ra1[1], NF, t), Lp(w.rr[1], NF, t), w.xbar);
    // interaction block (its x input: LayerNorm 2 of the first skip state when LayerNorm is on)
    const float* xi2 = x1b;
    if (ln) {
      SO3K_LAUNCH(k_tr_layernorm<T>, gw(N), dim3(TB), 0, st, N, (int)F, x1b, prm + lp.ln_s[1], prm + lp.ln_b[1], w.tA);
      xi2 = w.tA;
    }
    SO3K_LAUNCH(k_tr_intpost_bwd<T>, g1((size_t)N * C), dim3(TB), 0, st, D, N, c1, w.xbar, w.chibar, w.abbar);
    SO3K_LAUNCH(k_tr_intpre<T>, g1((size_t)N * C), dim3(TB), 0, st, D, N, xi2, c1, w.cat);
    gemm_tn(N, pl, (int)C, (int)C, w.cat, (int)C, w.abbar, (int)C, grad + lp.int_W, w.part, st);
    colsum(N, (int)C, pl, w.abbar, grad + lp.int_b, st);
    gemm_nt(rowsN, (int)C, (int)C, w.abbar, prm + lp.int_W, w.catbar, st);
    SO3K_LAUNCH(k_tr_intpre_bwd<T>, g1((size_t)N * (F + M)), dim3(TB), 0, st, D, N, c1, ab, w.xbar, w.chibar, w.catbar,
                w.x1bar, w.chi1bar, ln ? w.xnbar : (float*)nullptr);
    if (ln) ln_bwd(lp, 1, x1b, w.xnbar, w.x1bar, w.x1bar);           // + the skip connection already in x1bar
    if (res1) resmlp_bwd(lp.res[0], x1a, Lp(w.ra1[0], NF, t), Lp(w.rr[0], NF, t), w.x1bar);
    // attention + aggregation
    SO3K_LAUNCH(k_tr_alpha_bwd<T>, gw(Pe), dim3(TB), 0, st, D, N, Pt, g.rowptr, g.row, g.col, w.x1bar, w.chi1bar, proj[2],
                w.Y, w.phi, w.sbar);
    SO3K_LAUNCH(k_tr_vbar<T>, gw(N), dim3(TB), 0, st, D, N, Pt, g.rowptr, g.col, g.rev, al, w.x1bar, pbar[2]);
    for (int b = 0; b < 2; ++b) {
      const BlockParams& bp = b ? lp.gb : lp.fb;
      float* A1 = Lp(w.A1[b], Pe * F, t); float* As = Lp(w.As[b], Pe * Fs, t); float* wf = Lp(w.w[b], Pe * F, t);
      const float* qq = b ? proj[3] : proj[0]; const float* kk = b ? proj[4] : proj[1];
      SO3K_LAUNCH(k_tr_qkbar<T>, gw(N), dim3(TB), 0, st, D, N, Pt, b, g.rowptr, g.col, g.rev, qq, kk, wf, w.sbar,
                  b ? pbar[3] : pbar[0], b ? pbar[4] : pbar[1]);
      SO3K_LAUNCH(k_tr_wbar<T>, g1(Pe * F), dim3(TB), 0, st, D, N, Pt, b, g.rowptr, g.row, g.col, qq, kk, w.sbar, w.wbar);
      // filter MLP, radial branch: w = silu(rbf W1 + b1) W2 + b2
      colsum((int)Pe, (int)F, pl, w.wbar, grad + bp.rad_b2, st);
      colsum((int)Pe, (int)F, pl, w.wbar, grad + bp.sph_b2, st);
      SO3K_LAUNCH(k_tr_silu<T>, g1(Pe * F), dim3(TB), 0, st, Pe * F, A1, w.H);
      gemm_tn((int)Pe, pl, (int)F, (int)F, w.H, (int)F, w.wbar, (int)F, grad + bp.rad_W2, w.part, st);
      gemm_nt(rowsE, (int)F, (int)F, w.wbar, prm + bp.rad_W2, w.H, st, tc);          // Hbar
      SO3K_LAUNCH(k_tr_silu_bwd<T>, g1(Pe * F), dim3(TB), 0, st, Pe * F, A1, w.H);   // A1bar
      gemm_tn((int)Pe, pl, (int)K, (int)F, w.rbf, (int)K, w.H, (int)F, grad + bp.rad_W1, w.part, st);
      colsum((int)Pe, (int)F, pl, w.H, grad + bp.rad_b1, st);
      // spherical branch: w += silu(g Ws1 + bs1) Ws2 + bs2
      SO3K_LAUNCH(k_tr_silu<T>, g1(Pe * Fs), dim3(TB), 0, st, Pe * Fs, As, w.Hs);
      gemm_tn((int)Pe, pl, (int)Fs, (int)F, w.Hs, (int)Fs, w.wbar, (int)F, grad + bp.sph_W2, w.part, st);
      gemm_nt(rowsE, (int)Fs, (int)F, w.wbar, prm + bp.sph_W2, w.hsbar, st, tc);
      SO3K_LAUNCH(k_tr_silu_bwd<T>, g1(Pe * Fs), dim3(TB), 0, st, Pe * Fs, As, w.hsbar);   // Asbar
      gemm_tn((int)Pe, pl, (int)nl, (int)Fs, gl, (int)nl, w.hsbar, (int)Fs, grad + bp.sph_W1, w.part, st);
      colsum((int)Pe, (int)Fs, pl, w.hsbar, grad + bp.sph_b1, st);
      SO3K_LAUNCH(k_tr_sph1_bwd<T>, g1(Pe * nl), dim3(TB), 0, st, Pt, (int)nl, (int)Fs, b, w.hsbar, prm + bp.sph_W1, w.gbar);
    }
    // cotangent of (x_t, chi_t): skip connection + projections (through LayerNorm 1 when it is on) + chi differences
    const float* xin = ln ? Lp(w.xn1, NF, t) : xt;                   // what the projections read
    const float* Wp[5] = {prm + lp.fb.Wq, prm + lp.fb.Wk, prm + lp.fb.Wv, prm + lp.gb.Wq, prm + lp.gb.Wk};
    float* Gp[5] = {grad + lp.fb.Wq, grad + lp.fb.Wk, grad + lp.fb.Wv, grad + lp.gb.Wq, grad + lp.gb.Wk};
    float* pxbar = w.xbar;                                           // receives the projections' x cotangent
    if (ln) {
      pxbar = w.xnbar;
      cudaMemsetAsync(w.xnbar, 0, sizeof(float) * sx, st);
    } else {
      cudaMemcpyAsync(w.xbar, w.x1bar, sizeof(float) * sx, cudaMemcpyDeviceToDevice, st);
    }
    for (int k5 = 0; k5 < 5; ++k5) {
      const int d = k5 < 3 ? D.dh : D.dg;
      const int heads = (int)F / d;
      cudaMemsetAsync(w.dense, 0, sizeof(float) * F * F, st);
      gemm_tn(N, pl, (int)F, (int)F, xin, (int)F, pbar[k5], (int)F, w.dense, w.part, st);
      SO3K_LAUNCH(k_tr_extract_heads, g1((size_t)heads * d * d), dim3(TB), 0, st, (int)F, d, w.dense, Gp[k5]);
      SO3K_LAUNCH(k_tr_headproj_bwd_x<T>, g1(NF), dim3(TB), 0, st, N, (int)F, d, pbar[k5], Wp[k5], pxbar);
    }
    if (ln) ln_bwd(lp, 0, xt, w.xnbar, w.x1bar, w.xbar);             // xbar = x1bar (skip) + dLN1^T (projections' cotangent)
    if (t > 0)
      SO3K_LAUNCH(k_tr_gdiff_bwd<T>, gw(N), dim3(TB), 0, st, D, N, Pt, g.rowptr, g.col, g.rev, ct, w.gbar, w.chi1bar, w.chibar);
  }
  // x0 = Embed[z]: xbar now holds the cotangent of x0
  SO3K_LAUNCH(k_tr_embed_bwd<T>, g1(NF), dim3(TB), 0, st, N, (int)F, D.nemb, z, w.xbar, grad + ml.embed);
  return SO3K_OK;
}

}  // namespace

extern "C" {

size_t so3k_train_workspace_bytes(const so3k_handle* h, int64_t n_atoms_total, int64_t n_edge_slots_total) {
  if (!h || n_atoms_total < 0 || n_edge_slots_total < 0) return 0;
  return carve_train(h->dims, (int)n_atoms_total, (int)n_edge_slots_total, 2, nullptr, h->layout.node_opts).bytes + 256;
}

int so3k_energy_forces_vjp(so3k_handle* h, int32_t B, int32_t n, int32_t P, const float* R, const int32_t* z,
                           const int32_t* idx_i, const int32_t* idx_j, const float* cell, const int32_t* cell_offset,
                           const float* E_bar, const float* F_bar, const float* S_bar, float* grad, float* E_out,
                           void* workspace, size_t workspace_bytes, int32_t* dev_status, void* stream) {
  if (!h || !R || !z || !idx_i || !idx_j || !grad || !workspace || B <= 0 || n <= 0 || P < 0) return SO3K_ERR_ARG;
  if (!h->have_params) return SO3K_ERR_NOPARAMS;
  if ((cell == nullptr) != (cell_offset == nullptr)) return SO3K_ERR_ARG;
  if (workspace_bytes < so3k_train_workspace_bytes(h, (int64_t)B * n, (int64_t)B * P)) return SO3K_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  void* ws = (void*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  int rc = (F_bar || S_bar)
               ? run_vjp<Dual>(h, B, n, P, R, z, idx_i, idx_j, cell, cell_offset, E_bar, F_bar, S_bar, grad, E_out, ws, dev_status, st)
               : run_vjp<float>(h, B, n, P, R, z, idx_i, idx_j, cell, cell_offset, E_bar, nullptr, nullptr, grad, E_out, ws, dev_status, st);
  if (rc != SO3K_OK) return rc;
  return cudaPeekAtLastError() == cudaSuccess ? SO3K_OK : SO3K_ERR_CUDA;
}

int so3k_loss_sums(int32_t B, int32_t n, const float* E, const float* F, const float* E_true, const float* F_true,
                   const int32_t* z, double* sums, void* stream) {
  if (!E || !F || !E_true || !F_true || !z || !sums || B <= 0 || n <= 0) return SO3K_ERR_ARG;
  const size_t total = (size_t)B * n * 3;
  int blocks = (int)((total + TB - 1) / TB);
  if (blocks > 1024) blocks = 1024;
  SO3K_LAUNCH(k_loss_sums, dim3(blocks), dim3(TB), 0, (cudaStream_t)stream, B, n, E, F, E_true, F_true, z, sums);
  return cudaPeekAtLastError() == cudaSuccess ? SO3K_OK : SO3K_ERR_CUDA;
}

int so3k_loss_cotangents(int32_t B, int32_t n, const float* E, const float* F, const float* E_true, const float* F_true,
                         const int32_t* z, const double* sums, float w_energy, float w_force, float* E_bar, float* F_bar,
                         float* loss, void* stream) {
  if (!E || !F || !E_true || !F_true || !z || !sums || !E_bar || !F_bar || B <= 0 || n <= 0) return SO3K_ERR_ARG;
  SO3K_LAUNCH(k_loss_cotangents, g1((size_t)B * n * 3), dim3(TB), 0, (cudaStream_t)stream, B, n, E, F, E_true, F_true, z,
              sums, w_energy, w_force, E_bar, F_bar, loss);
  return cudaPeekAtLastError() == cudaSuccess ? SO3K_OK : SO3K_ERR_CUDA;
}

int so3k_adam_step(size_t n, float* params, const float* grad, float* m, float* v, const uint8_t* frozen,
                   const uint8_t* decay_mask, int64_t step, float lr, float b1, float b2, float eps, float eps_root,
                   float weight_decay, float clip_norm, double* grad_sumsq, int64_t* step_dev, void* stream) {
  if (!params || !grad || !m || !v || !grad_sumsq || (step < 1 && !step_dev)) return SO3K_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  cudaMemsetAsync(grad_sumsq, 0, sizeof(double), st);
  int blocks = (int)((n + TB - 1) / TB);
  SO3K_LAUNCH(k_sumsq, dim3(blocks > 512 ? 512 : blocks), dim3(TB), 0, st, n, grad, grad_sumsq, (long long*)step_dev);
  const float bc1 = 1.0f - powf(b1, (float)step), bc2 = 1.0f - powf(b2, (float)step);
  SO3K_LAUNCH(k_adam, dim3(blocks), dim3(TB), 0, st, n, params, grad, m, v, frozen, decay_mask, grad_sumsq, clip_norm, lr,
              b1, b2, eps, eps_root, weight_decay, bc1, bc2, (const long long*)step_dev);
  return cudaPeekAtLastError() == cudaSuccess ? SO3K_OK : SO3K_ERR_CUDA;
}

}  // extern "C"
