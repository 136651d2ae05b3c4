// Fused edge featurisation: one coalesced pass over edge vectors producing the radial basis,
// the cutoff, unit vectors and real spherical harmonics.
//   GeometryEmbed.__call__      mlff/nn/embed/embed.py:88-127
//   PhysNetBasis                mlff/basis_function/radial.py:163-166
//   cosine_cutoff_fn            mlff/cutoff_function/radial.py:50-51
//   add_cell_offsets            mlff/cutoff_function/pbc.py:17-18
//   fn_Y1..fn_Y4                mlff/basis_function/spherical.py:6-38
// Layout: a warp owns 32 consecutive edges.  Phase 1 is lane-per-edge (gather R[i], R[j], geometry,
// Y_lm into registers); phase 2 re-distributes so that every global store is a full 128-byte line:
// the (32 x K) rbf block is written one edge-row per iteration with lane == k, and the (32 x M) / (32 x 3)
// blocks are transposed through a per-warp shared-memory tile and streamed out linearly.
// Algorithmic bytes / edge: 8 (idx) + 12 (cell offset) in, 4*(K + M + 5) out (228 B at K=32, M=15)
// plus 12 B per atom of positions.
#include "so3k_internal.h"

namespace {

constexpr int FEAT_WARPS = 8;

__global__ void __launch_bounds__(FEAT_WARPS * 32)
k_featurise(So3kDims D, int n, int P, const float* __restrict__ R, const int* __restrict__ idx_i,
            const int* __restrict__ idx_j, const float* __restrict__ cell, const int* __restrict__ cell_offset,
            const float* __restrict__ disp, float* __restrict__ rbf, float* __restrict__ phi, float* __restrict__ dout,
            float* __restrict__ unit, float* __restrict__ sph) {
  __shared__ float s_tile[FEAT_WARPS][32 * (SO3K_MAXM + 1)];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long e0 = ((long long)blockIdx.x * FEAT_WARPS + w) * 32;
  const long long e = e0 + lane;
  float* tile = s_tile[w];

  // ---- phase 1: lane per edge -------------------------------------------------------------------
  float ux = 0.f, uy = 0.f, uz = 0.f, d = 0.f, ph = 0.f, ed = 0.f;
  bool valid = false;
  if (e < P) {
    int i = idx_i[e], j = idx_j[e];
    valid = (i >= 0) && (i < n) && (j >= 0) && (j < n);
    if (valid) {
      float rx, ry, rz;
      if (disp) {
        rx = disp[3 * e + 0]; ry = disp[3 * e + 1]; rz = disp[3 * e + 2];
      } else {
        rx = R[3 * (size_t)j + 0] - R[3 * (size_t)i + 0];
        ry = R[3 * (size_t)j + 1] - R[3 * (size_t)i + 1];
        rz = R[3 * (size_t)j + 2] - R[3 * (size_t)i + 2];
        if (cell && cell_offset) {
          float s0 = (float)cell_offset[3 * e + 0], s1 = (float)cell_offset[3 * e + 1], s2 = (float)cell_offset[3 * e + 2];
          rx += s0 * cell[0] + s1 * cell[3] + s2 * cell[6];
          ry += s0 * cell[1] + s1 * cell[4] + s2 * cell[7];
          rz += s0 * cell[2] + s1 * cell[5] + s2 * cell[8];
        }
      }
      d = sqrtf(rx * rx + ry * ry + rz * rz);
      if (d > 0.f) { ux = rx / d; uy = ry / d; uz = rz / d; }
      ph = so3k_cutoff(d, D.r_cut);
      ed = expf(-d);
    }
  }
  if (e < P) {
    if (phi) phi[e] = ph;
    if (dout) dout[e] = d;
  }
  const int nrow = (P - e0) < 32 ? (int)(P - e0) : 32;   // edges of this warp that exist (<=0: none)

  // ---- unit vectors: transpose through smem, linear store ------------------------------------------
  if (unit) {
    tile[lane * 3 + 0] = ux; tile[lane * 3 + 1] = uy; tile[lane * 3 + 2] = uz;
    __syncwarp();
    for (int t = lane; t < nrow * 3; t += 32) unit[e0 * 3 + t] = tile[t];
    __syncwarp();
  }
  // ---- spherical harmonics --------------------------------------------------------------------------
  if (sph) {
    float y[SO3K_MAXM];
    for (int m = 0; m < D.M; ++m) y[m] = 0.f;
    if (valid) so3k_sph_all(D, ux, uy, uz, y);
    for (int m = 0; m < D.M; ++m) tile[lane * D.M + m] = y[m];
    __syncwarp();
    for (int t = lane; t < nrow * D.M; t += 32) sph[e0 * D.M + t] = tile[t];
    __syncwarp();
  }
  // ---- radial basis: one edge-row per iteration, lane == k ------------------------------------------
  if (rbf) {
    const int lpr = D.K >> 2;                                // lanes per edge row when every lane writes 16 bytes
    if ((D.K & 3) == 0 && lpr <= 32 && (32 % lpr) == 0) {
      // vectorised: a store instruction writes 32 / lpr complete rows (one row at K = 128, four at K = 32)
      const int rpi = 32 / lpr, sub = lane / lpr, k0 = 4 * (lane % lpr);
      const float edv = valid ? ed : -1.0f;                  // ed = exp(-d) > 0 for a valid edge
      for (int q0 = 0; q0 < 32; q0 += rpi) {
        const int q = q0 + sub;
        const float edq = __shfl_sync(0xffffffffu, edv, q);
        if (q < nrow) {
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (edq >= 0.f) {
            const float t0 = edq - (D.mu0 + (float)k0 * D.dmu);
            const float t1 = t0 - D.dmu, t2 = t1 - D.dmu, t3 = t2 - D.dmu;
            v = make_float4(expf(-D.gamma * t0 * t0), expf(-D.gamma * t1 * t1), expf(-D.gamma * t2 * t2),
                            expf(-D.gamma * t3 * t3));
          }
          *reinterpret_cast<float4*>(rbf + (e0 + q) * D.K + k0) = v;
        }
      }
    } else {
      for (int q = 0; q < 32; ++q) {
        float edq = __shfl_sync(0xffffffffu, ed, q);
        int vq = __shfl_sync(0xffffffffu, valid ? 1 : 0, q);
        if (q < nrow) {
          for (int k = lane; k < D.K; k += 32) rbf[(e0 + q) * D.K + k] = vq ? so3k_rbf(edq, k, D) : 0.f;
        }
      }
    }
  }
}

}  // namespace

void so3k_launch_featurise(const So3kDims& D, int n, int P, const float* R, const int* idx_i, const int* idx_j,
                           const float* cell, const int* cell_offset, const float* disp, float* rbf, float* phi,
                           float* d, float* unit, float* sph, cudaStream_t st) {
  if (P <= 0) return;
  int blocks = so3k_cdiv(P, FEAT_WARPS * 32);
  SO3K_LAUNCH(k_featurise, dim3(blocks), dim3(FEAT_WARPS * 32), 0, st, D, n, P, R, idx_i, idx_j, cell, cell_offset,
              disp, rbf, phi, d, unit, sph);
}
