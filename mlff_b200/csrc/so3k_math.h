// Scalar math of the SO3krates hot path, usable from device code and (for the CPU unit tests of
// the formulas) from host code.  References (relative to the reference checkout):
//   PhysNet basis   mlff/basis_function/radial.py:163-166
//   cosine cutoff   mlff/cutoff_function/radial.py:50-51
//   real Y_lm       mlff/basis_function/spherical.py:6-38  (m = -l..l as listed there)
//   silu            jax.nn.silu via mlff/nn/activation_function
#pragma once
#include "so3k_portable.h"
#include <math.h>

#define SO3K_MAXDEG 8
#define SO3K_MAXM 32

// Hyper-parameters and small lookup tables, passed BY VALUE to kernels.
struct So3kDims {
  int F, K, L, H, nl, M, Fs, dh, dg, nemb;
  float r_cut;
  float mu0, dmu, gamma;          // PhysNet basis: mu_k = mu0 + k*dmu ; exp(-gamma (e^-d - mu_k)^2)
  int deg[SO3K_MAXDEG];           // degree value at each position of `degrees`
  int seg_off[SO3K_MAXDEG + 1];   // first flat m index of each position
  int m_seg[SO3K_MAXM];           // flat m -> position in `degrees`
  float m_cg[SO3K_MAXM];          // flat m -> l0-contraction coefficient (reference cg_rep order)
};

SO3K_HD static inline float so3k_silu(float a) {
  float s = 1.0f / (1.0f + expf(-a));
  return a * s;
}
// silu and its derivative
SO3K_HD static inline void so3k_silu_d(float a, float* y, float* dy) {
  float s = 1.0f / (1.0f + expf(-a));
  *y = a * s;
  *dy = s * (1.0f + a * (1.0f - s));
}

SO3K_HD static inline float so3k_cutoff(float d, float r_cut) {
  // 0.5 * (cos(d*pi/r_cut) + 1) for d < r_cut, else 0
  return d < r_cut ? 0.5f * (cosf(d * 3.14159265358979323846f / r_cut) + 1.0f) : 0.0f;
}
SO3K_HD static inline float so3k_cutoff_d(float d, float r_cut) {
  float w = 3.14159265358979323846f / r_cut;
  return d < r_cut ? -0.5f * w * sinf(d * w) : 0.0f;
}

// rbf_k(d) for k = 0..K-1 ; ed = exp(-d)
SO3K_HD static inline float so3k_rbf(float ed, int k, const So3kDims& D) {
  float t = ed - (D.mu0 + (float)k * D.dmu);
  return expf(-D.gamma * t * t);
}
// d rbf_k / d d  = rbf_k * 2 gamma (e^-d - mu_k) e^-d
SO3K_HD static inline float so3k_rbf_d(float ed, int k, float rbf, const So3kDims& D) {
  float t = ed - (D.mu0 + (float)k * D.dmu);
  return rbf * 2.0f * D.gamma * t * ed;
}

// ---- real spherical harmonics --------------------------------------------------------------
#define SO3K_C1 0.4886025119029199f    /* sqrt(3/(4pi)) */
#define SO3K_C2A 1.0925484305920792f   /* 1/2 sqrt(15/pi) */
#define SO3K_C2B 0.31539156525252005f  /* 1/4 sqrt(5/pi) */
#define SO3K_C2C 0.5462742152960396f   /* 1/4 sqrt(15/pi) */
#define SO3K_C3A 0.5900435899266435f   /* 1/4 sqrt(35/(2pi)) */
#define SO3K_C3B 2.890611442640554f    /* 1/2 sqrt(105/pi) */
#define SO3K_C3C 0.4570457994644658f   /* 1/4 sqrt(21/(2pi)) */
#define SO3K_C3D 0.3731763325901154f   /* 1/4 sqrt(7/pi) */
#define SO3K_C3E 1.445305721320277f    /* 1/4 sqrt(105/pi) */
#define SO3K_C4A 2.5033429417967046f   /* 3/4 sqrt(35/pi) */
#define SO3K_C4B 1.7701307697799304f   /* 3/4 sqrt(35/(2pi)) */
#define SO3K_C4C 0.9461746957575601f   /* 3/4 sqrt(5/pi) */
#define SO3K_C4D 0.6690465435572892f   /* 3/4 sqrt(5/(2pi)) */
#define SO3K_C4E 0.10578554691520431f  /* 3/16 sqrt(1/pi) */
#define SO3K_C4F 0.47308734787878004f  /* 3/8 sqrt(5/pi) */
#define SO3K_C4G 0.6258357354491761f   /* 3/16 sqrt(35/pi) */
#define SO3K_C0 0.28209479177387814f   /* 1/2 sqrt(1/pi) */

// Y_l(x,y,z) -> out[0..2l] ; T = float, or a forward-mode dual number (so3k_train.cu) with float * T, T * T, T +- T defined
template <class T>
SO3K_HD static inline void so3k_sph_g(int l, T x, T y, T z, T* out) {
  switch (l) {
    case 0:
      out[0] = x * 0.0f + SO3K_C0;
      break;
    case 1:
      out[0] = SO3K_C1 * y; out[1] = SO3K_C1 * z; out[2] = SO3K_C1 * x;
      break;
    case 2:
      out[0] = SO3K_C2A * x * y;
      out[1] = SO3K_C2A * y * z;
      out[2] = SO3K_C2B * (3.0f * z * z - 1.0f);
      out[3] = SO3K_C2A * x * z;
      out[4] = SO3K_C2C * (x * x - y * y);
      break;
    case 3:
      out[0] = SO3K_C3A * y * (3.0f * x * x - y * y);
      out[1] = SO3K_C3B * x * y * z;
      out[2] = SO3K_C3C * y * (5.0f * z * z - 1.0f);
      out[3] = SO3K_C3D * (5.0f * z * z * z - 3.0f * z);
      out[4] = SO3K_C3C * x * (5.0f * z * z - 1.0f);
      out[5] = SO3K_C3E * (x * x - y * y) * z;
      out[6] = SO3K_C3A * x * (x * x - 3.0f * y * y);
      break;
    case 4: {
      T x2 = x * x, y2 = y * y, z2 = z * z;
      out[0] = SO3K_C4A * x * y * (x2 - y2);
      out[1] = SO3K_C4B * y * (3.0f * x2 - y2) * z;
      out[2] = SO3K_C4C * x * y * (7.0f * z2 - 1.0f);
      out[3] = SO3K_C4D * y * (7.0f * z2 * z - 3.0f * z);
      out[4] = SO3K_C4E * (35.0f * z2 * z2 - 30.0f * z2 + 3.0f);
      out[5] = SO3K_C4D * x * (7.0f * z2 * z - 3.0f * z);
      out[6] = SO3K_C4F * (x2 - y2) * (7.0f * z2 - 1.0f);
      out[7] = SO3K_C4B * x * (x2 - 3.0f * y2) * z;
      out[8] = SO3K_C4G * (x2 * (x2 - 3.0f * y2) - y2 * (3.0f * x2 - y2));
    } break;
    default:
      break;
  }
}

SO3K_HD static inline void so3k_sph(int l, float x, float y, float z, float* out) { so3k_sph_g<float>(l, x, y, z, out); }

// Accumulate g += sum_m ybar[m] * dY_lm/d(x,y,z)  (x,y,z treated as independent variables of the
// polynomials above, exactly what reverse-mode AD of spherical.py does).
SO3K_HD static inline void so3k_sph_vjp(int l, float x, float y, float z, const float* yb,
                                        float* gx, float* gy, float* gz) {
  float ax = 0.f, ay = 0.f, az = 0.f;
  switch (l) {
    case 1:
      ax = SO3K_C1 * yb[2]; ay = SO3K_C1 * yb[0]; az = SO3K_C1 * yb[1];
      break;
    case 2:
      ax = SO3K_C2A * (y * yb[0] + z * yb[3]) + 2.0f * SO3K_C2C * x * yb[4];
      ay = SO3K_C2A * (x * yb[0] + z * yb[1]) - 2.0f * SO3K_C2C * y * yb[4];
      az = SO3K_C2A * (y * yb[1] + x * yb[3]) + 6.0f * SO3K_C2B * z * yb[2];
      break;
    case 3: {
      float x2 = x * x, y2 = y * y, z2 = z * z;
      ax = SO3K_C3A * (6.0f * x * y * yb[0] + (3.0f * x2 - 3.0f * y2) * yb[6]) + SO3K_C3B * y * z * yb[1] +
           SO3K_C3C * (5.0f * z2 - 1.0f) * yb[4] + SO3K_C3E * 2.0f * x * z * yb[5];
      ay = SO3K_C3A * ((3.0f * x2 - 3.0f * y2) * yb[0] - 6.0f * x * y * yb[6]) + SO3K_C3B * x * z * yb[1] +
           SO3K_C3C * (5.0f * z2 - 1.0f) * yb[2] - SO3K_C3E * 2.0f * y * z * yb[5];
      az = SO3K_C3B * x * y * yb[1] + SO3K_C3C * 10.0f * z * (y * yb[2] + x * yb[4]) +
           SO3K_C3D * (15.0f * z2 - 3.0f) * yb[3] + SO3K_C3E * (x2 - y2) * yb[5];
    } break;
    case 4: {
      float x2 = x * x, y2 = y * y, z2 = z * z;
      float p7 = 7.0f * z2 - 1.0f;            // 7z^2 - 1
      float q7 = 7.0f * z2 * z - 3.0f * z;    // 7z^3 - 3z
      float r7 = 21.0f * z2 - 3.0f;           // d q7 / dz
      ax = SO3K_C4A * (3.0f * x2 * y - y2 * y) * yb[0] + SO3K_C4B * 6.0f * x * y * z * yb[1] +
           SO3K_C4C * y * p7 * yb[2] + SO3K_C4D * q7 * yb[5] + SO3K_C4F * 2.0f * x * p7 * yb[6] +
           SO3K_C4B * (3.0f * x2 - 3.0f * y2) * z * yb[7] + SO3K_C4G * (4.0f * x2 * x - 12.0f * x * y2) * yb[8];
      ay = SO3K_C4A * (x2 * x - 3.0f * x * y2) * yb[0] + SO3K_C4B * (3.0f * x2 - 3.0f * y2) * z * yb[1] +
           SO3K_C4C * x * p7 * yb[2] + SO3K_C4D * q7 * yb[3] - SO3K_C4F * 2.0f * y * p7 * yb[6] -
           SO3K_C4B * 6.0f * x * y * z * yb[7] + SO3K_C4G * (4.0f * y2 * y - 12.0f * x2 * y) * yb[8];
      az = SO3K_C4B * y * (3.0f * x2 - y2) * yb[1] + SO3K_C4C * 14.0f * x * y * z * yb[2] +
           SO3K_C4D * r7 * (y * yb[3] + x * yb[5]) + SO3K_C4E * (140.0f * z2 * z - 60.0f * z) * yb[4] +
           SO3K_C4F * (x2 - y2) * 14.0f * z * yb[6] + SO3K_C4B * x * (x2 - 3.0f * y2) * yb[7];
    } break;
    default:
      break;
  }
  *gx += ax; *gy += ay; *gz += az;
}

// all degrees of the model: out[0..M)
SO3K_HD static inline void so3k_sph_all(const So3kDims& D, float x, float y, float z, float* out) {
  for (int s = 0; s < D.nl; ++s) so3k_sph(D.deg[s], x, y, z, out + D.seg_off[s]);
}
