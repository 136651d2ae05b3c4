// Receiver-indexed aggregations over the receiver-sorted CSR -- no atomics.
// A warp owns one receiver row.  It walks the row in chunks of 32 edges: the chunk's senders and attention
// coefficients are staged in shared memory (coalesced lane-per-edge loads), then the warp loops over the staged
// edges with lane == feature, gathering the F-wide sender rows with full 128-byte requests.  Quantities that are
// narrow per edge (the M spherical-harmonic coordinates, 3-vectors) are accumulated lane-per-edge and combined
// with warp-shuffle reductions at the end of the row.
//   AttentionAggregation   mlff/nn/layer/so3krates_layer.py:607-608   x_local = segment_sum(alpha * v[idx_j], idx_i)
//   SphConvAttention       mlff/nn/layer/so3krates_layer.py:563-564   chi_local = segment_sum(repeat(alpha') * Y, idx_i)
//   skip connections       mlff/nn/layer/so3krates_layer.py:146-147
// and the transposes needed by the analytic force backward.  Sender-indexed reductions are
// turned into receiver-indexed ones through the reverse-edge permutation (the edge list is symmetric).
// Algorithmic bytes (compulsory traffic, fp32): feature aggregation (4H+4)*P + (8F+4)*N, SPHC aggregation
// (4nl+16+4)*P + 8M*N with Y recomputed from the 16-byte (u,d) record instead of read (SURVEY 8(d) counts 4M/edge).
#include "so3k_internal.h"

namespace {

// Warps (receiver rows) per block.  Small blocks + explicit minimum-blocks bounds set the occupancy of the three streaming
// kernels (c4, ms per step, profiles/r03d_ab_stream_policies.txt): 8 warps -> 4: k_alpha_bwd 3.18 -> 2.56 (80 registers,
// 24 resident warps), -> 2.38 at 63 registers / 32 warps; k_alpha_aggregate_scan 6.49 -> 6.05 at 96 registers / 20 warps;
// k_qk_bwd_stream 7.49 -> 7.05 single-buffered at 96 registers / 20 warps (it spills below that).
#ifndef SO3K_AGG_WPB
#define SO3K_AGG_WPB 4
#endif
constexpr int WPB = SO3K_AGG_WPB;   // warps (rows) per block
constexpr int AMAX = 16;        // H + nl <= AMAX (stride in the alpha arrays is a multiple of 4)

// Head partial of feature slot t (features 32 t + lane) when the head width DHC is a compile-time constant >= 32: a
// slot then touches at most two heads, split at a compile-time lane index -- 1 or 2 predicated adds instead of one
// compare-select-add per head.
template <int DHC, int HPAD>
__device__ __forceinline__ void head_accum(float* val, int t, int lane, float p) {
  const int f0 = 32 * t;
  const int hlo = f0 / DHC;
  const int b = (hlo + 1) * DHC - f0;          // lanes < b belong to head hlo, the rest to hlo + 1
  if (b >= 32) {
    if (hlo < HPAD) val[hlo] += p;
  } else {
    if (hlo < HPAD) val[hlo] += (lane < b) ? p : 0.f;
    if (hlo + 1 < HPAD) val[hlo + 1] += (lane >= b) ? p : 0.f;
  }
}

// A load the compiler may not move or merge (asm volatile): used where a whole batch of independent row loads has to be
// in flight before its first use.  Read-only path, L1-allocating like a plain load.
__device__ __forceinline__ float4 ld_pinned4(const float* p) {
#ifdef SO3K_EMU
  return *reinterpret_cast<const float4*>(p);
#else
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
#endif
}

// Filter rows (w[pid e]) are read exactly twice per sweep -- by the two directed edges of the pair, from two different
// SMs -- so they never re-hit in L1 and their second reader is their last use.  Load policy (template parameter MODE):
//   0  plain read-only loads
//   1  filter rows bypass L1 (L1::no_allocate): L1 keeps the sender rows, which neighbouring receivers share
//   2  1 + the second reader (the sender's row comes before the receiver's: j < i) marks the line L2::evict_first
//   3  2 + sender rows carry an L2::evict_last hint
// Measured on c4 (profiles/README.md, r03d): k_alpha_aggregate_scan 6.66 / 6.60 / 6.48 / 6.53 ms per step for modes 0-3,
// k_qk_bwd_stream 7.49 / 7.87 / 9.33 / 10.5 (the policy operand costs it registers) -> 2 and 0.
#ifndef SO3K_AGG_W_POLICY
#define SO3K_AGG_W_POLICY 2
#endif
#ifndef SO3K_QK_W_POLICY
#define SO3K_QK_W_POLICY 0
#endif
struct WPolicy { unsigned long long first, normal, last; };
template <int MODE>
__device__ __forceinline__ WPolicy make_w_policy() {
  WPolicy P;
  P.first = P.normal = P.last = 0ull;
#ifndef SO3K_EMU
  if (MODE >= 2) {
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(P.first));
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(P.normal));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(P.last));
  }
#endif
  return P;
}
template <int MODE>
__device__ __forceinline__ float ld_w(const float* p, unsigned long long pol) {
#ifdef SO3K_EMU
  return *p;
#else
  float v;
  if (MODE == 0) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else if (MODE == 1) asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
#endif
}
template <int MODE>
__device__ __forceinline__ float4 ld_w4(const float* p, unsigned long long pol) {
#ifdef SO3K_EMU
  return *reinterpret_cast<const float4*>(p);
#else
  float4 v;
  if (MODE == 0)
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  else if (MODE == 1)
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  else
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
#endif
}
template <int MODE>
__device__ __forceinline__ float ld_row(const float* p, unsigned long long pol) {      // sender rows
#ifdef SO3K_EMU
  return *p;
#else
  float v;
  if (MODE < 3) asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  else asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
#endif
}
template <int MODE>
__device__ __forceinline__ float4 ld_row4(const float* p, unsigned long long pol) {
#ifdef SO3K_EMU
  return *reinterpret_cast<const float4*>(p);
#else
  float4 v;
  if (MODE < 3)
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  else
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
#endif
}

__device__ __forceinline__ float warp_sum(float v) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// x1 = x + sum_e alpha_fb[e][h(f)] v[col e][f] ;  chi1 = chi + sum_e alpha_gb[e][l(m)] Y_m(u_e)
template <int TN>                               // TN = ceil(F / 32): feature slots per lane
__global__ void __launch_bounds__(WPB * 32)
k_aggregate(So3kDims D, int N, int Ast, const int* __restrict__ rowptr, const int* __restrict__ col,
            const float4* __restrict__ geom, const float* __restrict__ x, const float* __restrict__ chi,
            const float* __restrict__ proj, const float* __restrict__ alpha, float* __restrict__ x1,
            float* __restrict__ chi1) {
  __shared__ float s_a[WPB][32 * AMAX];
  __shared__ int s_j[WPB][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * WPB + w;
  const int F = D.F, M = D.M, H = D.H;
  const bool rowok = i < N;
  const int beg = rowok ? rowptr[i] : 0, end = rowok ? rowptr[i + 1] : 0;
  float accx[TN];
  int hsel[TN];
#pragma unroll
  for (int t = 0; t < TN; ++t) {
    accx[t] = 0.f;
    int f = lane + 32 * t;
    hsel[t] = (f < F) ? f / D.dh : 0;
  }
  float cacc[SO3K_MAXM];
  for (int m = 0; m < M; ++m) cacc[m] = 0.f;
  const size_t F5 = 5 * (size_t)F;
  for (int e0 = beg; e0 < end; e0 += 32) {
    const int e = e0 + lane;
    const int cnt = (end - e0) < 32 ? (end - e0) : 32;
    int j = 0;
    if (e < end) {
      j = col[e];
      for (int a = 0; a < Ast; ++a) s_a[w][lane * Ast + a] = alpha[(size_t)e * Ast + a];
      // SPHC part, lane-per-edge
      float4 g = geom[e];
      float y[SO3K_MAXM];
      so3k_sph_all(D, g.x, g.y, g.z, y);
      for (int m = 0; m < M; ++m) cacc[m] = fmaf(s_a[w][lane * Ast + H + D.m_seg[m]], y[m], cacc[m]);
    }
    s_j[w][lane] = j;
    __syncwarp();
    // four sender rows in flight per iteration (the gathers are latency-bound otherwise)
    for (int q = 0; q < cnt; q += 4) {
      float vv[4][TN];
      float am[4];
      int qk[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        qk[k] = (q + k < cnt) ? q + k : cnt - 1;
        am[k] = (q + k < cnt) ? 1.f : 0.f;
        const float* vj = proj + (size_t)s_j[w][qk[k]] * F5 + 4 * (size_t)F;
#pragma unroll
        for (int t = 0; t < TN; ++t) {
          const int f = lane + 32 * t;
          vv[k][t] = ((t + 1 < TN) || f < F) ? vj[f] : 0.f;
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float* aq = &s_a[w][qk[k] * Ast];
#pragma unroll
        for (int t = 0; t < TN; ++t) accx[t] = fmaf(am[k] * aq[hsel[t]], vv[k][t], accx[t]);
      }
    }
    __syncwarp();
  }
  // full-warp reductions of the lane-per-edge accumulators (every lane of the warp takes part)
  for (int m = 0; m < M; ++m) {
    float s = warp_sum(cacc[m]);
    if (rowok && lane == 0) chi1[(size_t)i * M + m] = chi[(size_t)i * M + m] + s;
  }
  if (rowok) {
#pragma unroll
    for (int t = 0; t < TN; ++t) {
      int f = lane + 32 * t;
      if (f < F) x1[(size_t)i * F + f] = x[(size_t)i * F + f] + accx[t];
    }
  }
}

// Tensor-core path: attention coefficients from the pair-indexed filters, fused with both aggregations.
//   alpha_fb[e][h] = phi_e / sqrt(F/H)  sum_{f in h} w_fb[pid e][f] q_fb[i][f] k_fb[j][f]       (so3krates_layer.py:590-606)
//   alpha_gb[e][l] = phi_e / sqrt(F/nl) sum_{f in l} w_gb[pid e][f] q_gb[i][f] k_gb[j][f]       (so3krates_layer.py:545-562)
//   x1 = x + sum_e alpha_fb[e][h(f)] v[j][f] ;  chi1 = chi + sum_e alpha_gb[e][l(m)] Y_m(u_e)
// EP edges per iteration; their EP * 2 * HS per-lane head partials (HS slots per block) are reduced over the warp by one
// halving butterfly (16 shuffles), as in k_alpha_bwd.
template <int TN, int HS, int DHF, int DHG>     // TN = ceil(F / 32) ; HS = 4 (H, nl <= 4: two edges per pass) or 8 ;
__global__ void __launch_bounds__(WPB * 32)     // DHF / DHG = compile-time head widths F/H, F/nl (0: run-time)
k_alpha_aggregate(So3kDims D, int N, int Ast, const int* __restrict__ rowptr, const int* __restrict__ col,
                  const int* __restrict__ pid, const float4* __restrict__ geom, const float* __restrict__ x,
                  const float* __restrict__ chi, const float* __restrict__ proj, const float* __restrict__ wst_f,
                  const float* __restrict__ wst_g, float* __restrict__ alpha, float* __restrict__ x1,
                  float* __restrict__ chi1) {
  constexpr int EP = 8 / HS;
  __shared__ float s_a[WPB][32 * AMAX];
  __shared__ float s_ph[WPB][32];
  __shared__ int s_j[WPB][32];
  __shared__ int s_p[WPB][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * WPB + w;
  const int F = D.F, M = D.M, H = D.H, nl = D.nl;
  const bool rowok = i < N;
  const int beg = rowok ? rowptr[i] : 0, end = rowok ? rowptr[i + 1] : 0;
  const size_t F5 = 5 * (size_t)F;
  const float inv_f = 1.0f / sqrtf((float)(F / H)), inv_g = 1.0f / sqrtf((float)(F / nl));
  float accx[TN], qf[TN], qg[TN];
  int hf[TN], hg[TN];
#pragma unroll
  for (int t = 0; t < TN; ++t) {
    accx[t] = 0.f;
    const int f = lane + 32 * t;
    const bool ok = rowok && f < F;
    qf[t] = ok ? proj[(size_t)i * F5 + f] : 0.f;           // zero beyond F: those slots add nothing
    qg[t] = ok ? proj[(size_t)i * F5 + F + f] : 0.f;
    hf[t] = (f < F) ? f / (F / H) : 0;
    hg[t] = (f < F) ? f / (F / nl) : 0;
  }
  float cacc[SO3K_MAXM];
  for (int m = 0; m < M; ++m) cacc[m] = 0.f;
  for (int e0 = beg; e0 < end; e0 += 32) {
    const int e = e0 + lane;
    const int cnt = (end - e0) < 32 ? (end - e0) : 32;
    float4 g = make_float4(0.f, 0.f, 1.f, 0.f);
    if (e < end) {
      g = geom[e];
      s_j[w][lane] = col[e];
      s_p[w][lane] = pid[e];
      s_ph[w][lane] = so3k_cutoff(g.w, D.r_cut);
      for (int a = H + nl; a < Ast; ++a) s_a[w][lane * Ast + a] = 0.f;
    }
    __syncwarp();
    for (int q = 0; q < cnt; q += EP) {
      float wf[EP][TN], wg[EP][TN], kf[EP][TN], kg[EP][TN], vv[EP][TN];
      int qk[EP];
#pragma unroll
      for (int k = 0; k < EP; ++k) {
        qk[k] = (q + k < cnt) ? q + k : cnt - 1;
        const float* pj = proj + (size_t)s_j[w][qk[k]] * F5;
        const float* pwf = wst_f + (size_t)s_p[w][qk[k]] * F;
        const float* pwg = wst_g + (size_t)s_p[w][qk[k]] * F;
#pragma unroll
        for (int t = 0; t < TN; ++t) {
          const int f = lane + 32 * t;
          const bool ok = (t + 1 < TN) || f < F;
          wf[k][t] = ok ? pwf[f] : 0.f;
          wg[k][t] = ok ? pwg[f] : 0.f;
          kf[k][t] = ok ? pj[2 * F + f] : 0.f;
          kg[k][t] = ok ? pj[3 * F + f] : 0.f;
          vv[k][t] = ok ? pj[4 * F + f] : 0.f;
        }
      }
      float val[16];                           // [edge k][block][head slot]
#pragma unroll
      for (int u = 0; u < 16; ++u) val[u] = 0.f;
#pragma unroll
      for (int k = 0; k < EP; ++k) {
#pragma unroll
        for (int t = 0; t < TN; ++t) {
          const float pf = wf[k][t] * qf[t] * kf[k][t];
          const float pg = wg[k][t] * qg[t] * kg[k][t];
          if (DHF >= 32 && DHG >= 32) {
            head_accum<(DHF >= 32 ? DHF : 32), HS>(val + k * 2 * HS, t, lane, pf);
            head_accum<(DHG >= 32 ? DHG : 32), HS>(val + k * 2 * HS + HS, t, lane, pg);
          } else {
#pragma unroll
            for (int h = 0; h < HS; ++h) {
              val[k * 2 * HS + h] += (hf[t] == h) ? pf : 0.f;
              val[k * 2 * HS + HS + h] += (hg[t] == h) ? pg : 0.f;
            }
          }
        }
      }
#pragma unroll
      for (int st = 0; st < 4; ++st) {         // halving butterfly: afterwards a lane holds value index (lane >> 1)
        const int off = 16 >> st, half = 8 >> st;
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int u = 0; u < half; ++u) {
          const float send = up ? val[u] : val[u + half];
          const float keep = up ? val[u + half] : val[u];
          val[u] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      val[0] += __shfl_xor_sync(0xffffffffu, val[0], 1);
      {
        const int idx = lane >> 1, k = idx / (2 * HS), a = idx % (2 * HS);
        if ((lane & 1) == 0 && q + k < cnt) {
          const float ph = s_ph[w][q + k];
          if (a < HS) {
            if (a < H) s_a[w][(q + k) * Ast + a] = val[0] * inv_f * ph;
          } else if (a - HS < nl) {
            s_a[w][(q + k) * Ast + H + a - HS] = val[0] * inv_g * ph;
          }
        }
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < EP; ++k) {
        const float m = (q + k < cnt) ? 1.f : 0.f;
        const float* aq = &s_a[w][qk[k] * Ast];
#pragma unroll
        for (int t = 0; t < TN; ++t) accx[t] = fmaf(m * aq[hf[t]], vv[k][t], accx[t]);
      }
    }
    __syncwarp();
    if (e < end) {                             // SPHC part, lane-per-edge
      float y[SO3K_MAXM];
      so3k_sph_all(D, g.x, g.y, g.z, y);
      for (int m = 0; m < M; ++m) cacc[m] = fmaf(s_a[w][lane * Ast + H + D.m_seg[m]], y[m], cacc[m]);
    }
    for (int t = lane; t < cnt * Ast; t += 32) alpha[(size_t)e0 * Ast + t] = s_a[w][t];
    __syncwarp();
  }
  for (int m = 0; m < M; ++m) {
    float s = warp_sum(cacc[m]);
    if (rowok && lane == 0) chi1[(size_t)i * M + m] = chi[(size_t)i * M + m] + s;
  }
  if (rowok) {
#pragma unroll
    for (int t = 0; t < TN; ++t) {
      int f = lane + 32 * t;
      if (f < F) x1[(size_t)i * F + f] = x[(size_t)i * F + f] + accx[t];
    }
  }
}

// =====================================================================================================================
// Fast variant of k_alpha_aggregate for rows of at most 160 features (one 128-bit chunk per lane + optional tail scalar)
// whose head layout lets a warp scan form the head sums (so3k_scan_ok): a lane's chunk touches at most two adjacent heads
// (head width >= 4), and the tail scalars together with the chunks of the last ntail lanes lie in the LAST head.  The
// default model (F = 132, 4 heads of 33 / 3 heads of 44) qualifies.  Measured on c4: 6.74 ms per step against 7.31 ms for
// the general kernel above.  (The same treatment of k_alpha_bwd and k_qk_bwd_stream was SLOWER than the lane-per-feature
// kernels -- 3.71 vs 3.12 ms and 8.95 vs 8.04 ms -- and is not kept; a shared-memory staging of the per-feature products
// for the head sums was slower still: 70 % of the LSU wavefront peak, profiles/README.md r02c.)
// Two things make it cheaper:
//   * row addresses: every array has a per-lane base pointer (array + this lane's float offset), so a row address is one
//     64-bit multiply-add (base + row * stride) instead of a chain of integer instructions per load;
//   * head sums: a lane folds its chunk's products into (lo, hi) = (part in the chunk's first head, part in the next), an
//     inclusive warp scan of lo + hi gives running totals, and head h = X[last lane of h] - X[last lane of h - 1] with
//     X = scan - hi: 8 shuffles and ~25 arithmetic instructions per edge and block, no shared memory.
struct HeadScan {
  float m[4];          // 1 where element c of the lane's chunk belongs to the chunk's first head h0, 0 where to h0 + 1
  int h0, h1;          // heads of the chunk (h1 = min(h0 + 1, heads - 1))
  int src_hi, src_lo;  // lanes whose X holds the running total through head `lane` / through head `lane - 1`
  float use_lo;        // 1 for lanes 1 .. heads - 1 (subtract the previous running total), else 0
};
__device__ __forceinline__ HeadScan make_head_scan(int lane, int F, int heads) {
  HeadScan s;
  const int dw = F / heads, f0 = 4 * lane;
  int h0 = f0 / dw;
  if (h0 > heads - 1) h0 = heads - 1;
  const int nlo = (h0 + 1) * dw - f0;                  // elements of the chunk inside head h0 (>= 4: all of them)
#pragma unroll
  for (int c = 0; c < 4; ++c) s.m[c] = (c < nlo || h0 == heads - 1) ? 1.f : 0.f;
  s.h0 = h0;
  s.h1 = h0 + 1 < heads ? h0 + 1 : heads - 1;
  const int o = lane < heads ? lane : heads - 1;
  int bh = ((o + 1) * dw - 1) / 4, bl = o > 0 ? (o * dw - 1) / 4 : 0;
  s.src_hi = (o == heads - 1 || bh > 31) ? 31 : bh;    // the last head runs to the end of the row (tail included)
  s.src_lo = bl > 31 ? 31 : bl;
  s.use_lo = (lane > 0 && lane < heads) ? 1.f : 0.f;
  return s;
}
// head sums of one edge: lane h < heads returns the sum over head h of the per-feature products (p = this lane's chunk,
// tailp = its tail scalar's product or 0)
template <bool TAIL>
__device__ __forceinline__ float head_scan_sums(const float* p, float tailp, const HeadScan& s, int lane) {
  float lo = s.m[0] * p[0];
  lo = fmaf(s.m[1], p[1], lo);
  lo = fmaf(s.m[2], p[2], lo);
  lo = fmaf(s.m[3], p[3], lo);
  float tot = (p[0] + p[1]) + (p[2] + p[3]);
  const float hi = tot - lo;
  if (TAIL) tot += __shfl_sync(0xffffffffu, tailp, 31 - lane);     // tail scalars sit in lanes 0.. : hand them to the last lanes
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const float v = __shfl_up_sync(0xffffffffu, tot, d);
    tot += (lane >= d) ? v : 0.f;
  }
  const float X = tot - hi;
  const float a = __shfl_sync(0xffffffffu, X, s.src_hi);
  const float b = __shfl_sync(0xffffffffu, X, s.src_lo);
  return fmaf(-s.use_lo, b, a);
}
// per-element coefficients of a lane's chunk from per-head values held by lanes 0 .. heads - 1
__device__ __forceinline__ void head_scan_coef(float per_head, const HeadScan& s, float* coef) {
  const float a0 = __shfl_sync(0xffffffffu, per_head, s.h0), a1 = __shfl_sync(0xffffffffu, per_head, s.h1);
  const float d = a0 - a1;
#pragma unroll
  for (int c = 0; c < 4; ++c) coef[c] = fmaf(s.m[c], d, a1);
}
static bool so3k_scan_ok(int F, int heads) {
  if (F > 160 || F % 4 || heads < 1 || F % heads) return false;
  const int dw = F / heads, ntail = F > 128 ? F - 128 : 0;
  if (dw < 4) return false;
  return ntail == 0 || (heads - 1) * dw <= 128 - 4 * ntail;
}

// k_alpha_aggregate, fast variant (see there for the math)
#ifndef SO3K_AGG_MINB
#define SO3K_AGG_MINB (20 / WPB)      /* 20 resident warps per SM at <= 96 registers */
#endif
#ifndef SO3K_AGG_EPB
#define SO3K_AGG_EPB 2
#endif
template <bool TAIL>
__global__ void __launch_bounds__(WPB * 32, SO3K_AGG_MINB)
k_alpha_aggregate_scan(So3kDims D, int N, int Ast, const int* __restrict__ rowptr, const int* __restrict__ col,
                       const int* __restrict__ pid, const float4* __restrict__ geom, const float* __restrict__ x,
                       const float* __restrict__ chi, const float* __restrict__ proj, const float* __restrict__ wst_f,
                       const float* __restrict__ wst_g, float* __restrict__ alpha, float* __restrict__ x1,
                       float* __restrict__ chi1) {
  constexpr int EPB = SO3K_AGG_EPB, WP = SO3K_AGG_W_POLICY;
  __shared__ float s_a[WPB][32 * AMAX];
  __shared__ float s_ph[WPB][32];
  __shared__ int s_j[WPB][32];
  __shared__ int s_p[WPB][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * WPB + w;
  const int F = D.F, M = D.M, H = D.H, nl = D.nl, A = H + nl;
  const bool rowok = i < N;
  const int beg = rowok ? rowptr[i] : 0, end = rowok ? rowptr[i + 1] : 0;
  const unsigned F5 = 5u * (unsigned)F, Fu = (unsigned)F;
  const float inv_f = 1.0f / sqrtf((float)(F / H)), inv_g = 1.0f / sqrtf((float)(F / nl));
  const bool okv = 4 * lane < F, okt = TAIL && 128 + lane < F;
  const int offv = okv ? 4 * lane : 0, offt = okt ? 128 + lane : 0;          // clamped: masked lanes read a valid address
  // per-lane base pointers: row address = base + row * stride
  const float* const wf_v = wst_f + offv; const float* const wf_t = wst_f + offt;
  const float* const wg_v = wst_g + offv; const float* const wg_t = wst_g + offt;
  const float* const kf_v = proj + 2 * F + offv; const float* const kf_t = proj + 2 * F + offt;
  const float* const kg_v = proj + 3 * F + offv; const float* const kg_t = proj + 3 * F + offt;
  const float* const vv_v = proj + 4 * F + offv; const float* const vv_t = proj + 4 * F + offt;
  const float mv = okv ? 1.f : 0.f, mt = okt ? 1.f : 0.f;
  const WPolicy wpol = make_w_policy<WP>();
  const HeadScan hsf = make_head_scan(lane, F, H), hsg = make_head_scan(lane, F, nl);
  float4 qf = make_float4(0.f, 0.f, 0.f, 0.f), qg = qf;
  float qft = 0.f, qgt = 0.f;
  if (rowok) {
    const float* pi = proj + (size_t)i * F5;
    if (okv) { qf = ld_pinned4(pi + offv); qg = ld_pinned4(pi + F + offv); }
    if (okt) { qft = pi[offt]; qgt = pi[F + offt]; }
  }
  float accx[4] = {0.f, 0.f, 0.f, 0.f}, acct = 0.f;
  float cacc[SO3K_MAXM];
  for (int m = 0; m < M; ++m) cacc[m] = 0.f;
  for (int e0 = beg; e0 < end; e0 += 32) {
    const int e = e0 + lane;
    const int cnt = (end - e0) < 32 ? (end - e0) : 32;
    float4 g = make_float4(0.f, 0.f, 1.f, 0.f);
    if (e < end) {
      g = geom[e];
      s_j[w][lane] = col[e];
      s_p[w][lane] = pid[e];
      s_ph[w][lane] = so3k_cutoff(g.w, D.r_cut);
      for (int a = A; a < Ast; ++a) s_a[w][lane * Ast + a] = 0.f;
    }
    __syncwarp();
    for (int q = 0; q < cnt; q += EPB) {
      float4 wf[EPB], wg[EPB], kf[EPB], kg[EPB], vv[EPB];
      float wft[EPB], wgt[EPB], kft[EPB], kgt[EPB], vvt[EPB];
#pragma unroll
      for (int k = 0; k < EPB; ++k) {
        const int qc = (q + k < cnt) ? q + k : cnt - 1;
        const unsigned jr = (unsigned)s_j[w][qc], pr = (unsigned)s_p[w][qc];
        const unsigned long long pw = (int)jr < i ? wpol.first : wpol.normal;      // second reader of the pair's rows
        wf[k] = ld_w4<WP>(wf_v + (size_t)pr * Fu, pw);
        wg[k] = ld_w4<WP>(wg_v + (size_t)pr * Fu, pw);
        kf[k] = ld_row4<WP>(kf_v + (size_t)jr * F5, wpol.last);
        kg[k] = ld_row4<WP>(kg_v + (size_t)jr * F5, wpol.last);
        vv[k] = ld_row4<WP>(vv_v + (size_t)jr * F5, wpol.last);
        if (TAIL) {
          wft[k] = ld_w<WP>(wf_t + (size_t)pr * Fu, pw);
          wgt[k] = ld_w<WP>(wg_t + (size_t)pr * Fu, pw);
          kft[k] = ld_row<WP>(kf_t + (size_t)jr * F5, wpol.last);
          kgt[k] = ld_row<WP>(kg_t + (size_t)jr * F5, wpol.last);
          vvt[k] = ld_row<WP>(vv_t + (size_t)jr * F5, wpol.last);
        }
      }
#pragma unroll
      for (int k = 0; k < EPB; ++k) {
        const bool live = q + k < cnt;                               // warp-uniform
        // masked lanes (beyond F) loaded a valid but foreign address: their products are zeroed by mv / mt
        const float pf[4] = {mv * wf[k].x * (qf.x * kf[k].x), mv * wf[k].y * (qf.y * kf[k].y),
                             mv * wf[k].z * (qf.z * kf[k].z), mv * wf[k].w * (qf.w * kf[k].w)};
        const float pg[4] = {mv * wg[k].x * (qg.x * kg[k].x), mv * wg[k].y * (qg.y * kg[k].y),
                             mv * wg[k].z * (qg.z * kg[k].z), mv * wg[k].w * (qg.w * kg[k].w)};
        const float pft = TAIL ? mt * wft[k] * (qft * kft[k]) : 0.f, pgt = TAIL ? mt * wgt[k] * (qgt * kgt[k]) : 0.f;
        const float ph = s_ph[w][live ? q + k : cnt - 1];
        const float af = head_scan_sums<TAIL>(pf, pft, hsf, lane) * (inv_f * ph);    // lane h < H: alpha_fb[h]
        const float ag = head_scan_sums<TAIL>(pg, pgt, hsg, lane) * (inv_g * ph);    // lane l < nl: alpha_gb[l]
        if (live) {
          if (lane < H) s_a[w][(q + k) * Ast + lane] = af;
          if (lane < nl) s_a[w][(q + k) * Ast + H + lane] = ag;
          float cf[4];
          head_scan_coef(af, hsf, cf);
          accx[0] = fmaf(cf[0], vv[k].x, accx[0]);
          accx[1] = fmaf(cf[1], vv[k].y, accx[1]);
          accx[2] = fmaf(cf[2], vv[k].z, accx[2]);
          accx[3] = fmaf(cf[3], vv[k].w, accx[3]);
          if (TAIL) acct = fmaf(__shfl_sync(0xffffffffu, af, H - 1), vvt[k], acct);
        }
      }
    }
    __syncwarp();
    if (e < end) {                             // SPHC part, lane-per-edge
      float y[SO3K_MAXM];
      so3k_sph_all(D, g.x, g.y, g.z, y);
      for (int m = 0; m < M; ++m) cacc[m] = fmaf(s_a[w][lane * Ast + H + D.m_seg[m]], y[m], cacc[m]);
    }
    for (int t = lane; t < cnt * Ast; t += 32) alpha[(size_t)e0 * Ast + t] = s_a[w][t];
    __syncwarp();
  }
  for (int m = 0; m < M; ++m) {
    float s = warp_sum(cacc[m]);
    if (rowok && lane == 0) chi1[(size_t)i * M + m] = chi[(size_t)i * M + m] + s;
  }
  if (rowok) {
    const float* xi = x + (size_t)i * Fu;
    float* xo = x1 + (size_t)i * Fu;
    if (okv) {
      const float4 t = *reinterpret_cast<const float4*>(xi + offv);
      *reinterpret_cast<float4*>(xo + offv) = make_float4(t.x + accx[0], t.y + accx[1], t.z + accx[2], t.w + accx[3]);
    }
    if (okt) xo[offt] = xi[offt] + acct;
  }
}

// alphabar_fb[e][h] = sum_{f in h} xbar1[i][f] v[j][f] ; alphabar_gb[e][l] = sum_{m in l} chibar1[i][m] Y_m(u_e)
// gproj[i].v[f]     = sum_{e in row i} alpha_fb[rev e][h(f)] xbar1[col e][f]        (= d/dv_i, sender-indexed sum)
// EP = 16 / HP edges per iteration: their EP * HP per-lane head partials are reduced over the warp with ONE halving
// butterfly (16 shuffles for all of them instead of 5 per value).
#ifndef SO3K_ABWD_MINB
#define SO3K_ABWD_MINB (TN <= 5 ? 32 / WPB : 16 / WPB)     /* 63 registers hold up to 5 feature slots per lane */
#endif
#ifndef SO3K_QK_SINGLE
#define SO3K_QK_SINGLE 1
#endif
#ifndef SO3K_QK_MINB
#define SO3K_QK_MINB (SO3K_QK_SINGLE ? 20 / WPB : 16 / WPB)
#endif
template <int TN, int HP, int DHC>              // TN = ceil(F / 32) ; HP = heads padded to 4 or 8 ; DHC = compile-time
__global__ void __launch_bounds__(WPB * 32, SO3K_ABWD_MINB)     // head width F/H (0: run-time)
k_alpha_bwd(So3kDims D, int N, int Ast, const int* __restrict__ rowptr, const int* __restrict__ col,
            const int* __restrict__ rev, const float4* __restrict__ geom, const float* __restrict__ proj,
            const float* __restrict__ alpha, const float* __restrict__ xbar1, const float* __restrict__ chibar1,
            float* __restrict__ alphabar, float* __restrict__ gproj) {
  constexpr int EP = 16 / HP;
  __shared__ float s_a[WPB][32 * AMAX];     // alpha_fb of the reverse edges
  __shared__ float s_o[WPB][32 * AMAX];     // alphabar staging (coalesced store)
  __shared__ int s_j[WPB][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * WPB + w;
  const int F = D.F, M = D.M, H = D.H, nl = D.nl;
  const bool rowok = i < N;
  const int beg = rowok ? rowptr[i] : 0, end = rowok ? rowptr[i + 1] : 0;
  const size_t F5 = 5 * (size_t)F;
  float xb[TN], accv[TN];
  int hsel[TN];
#pragma unroll
  for (int t = 0; t < TN; ++t) {
    int f = lane + 32 * t;
    xb[t] = (rowok && f < F) ? xbar1[(size_t)i * F + f] : 0.f;
    accv[t] = 0.f;
    hsel[t] = (f < F) ? f / D.dh : 0;       // xb = 0 beyond F: those slots add nothing
  }
  float cb[SO3K_MAXM];
  for (int m = 0; m < M; ++m) cb[m] = rowok ? chibar1[(size_t)i * M + m] : 0.f;
  for (int e0 = beg; e0 < end; e0 += 32) {
    const int e = e0 + lane;
    const int cnt = (end - e0) < 32 ? (end - e0) : 32;
    int j = 0;
    if (e < end) {
      j = col[e];
      int re = rev[e];
      for (int h = 0; h < H; ++h) s_a[w][lane * Ast + h] = alpha[(size_t)re * Ast + h];
      float4 g = geom[e];
      float y[SO3K_MAXM];
      so3k_sph_all(D, g.x, g.y, g.z, y);
      for (int l = 0; l < nl; ++l) {
        float s = 0.f;
        for (int m = D.seg_off[l]; m < D.seg_off[l + 1]; ++m) s = fmaf(cb[m], y[m], s);
        s_o[w][lane * Ast + H + l] = s;
      }
      for (int a = H + nl; a < Ast; ++a) s_o[w][lane * Ast + a] = 0.f;
    }
    s_j[w][lane] = j;
    __syncwarp();
    for (int q = 0; q < cnt; q += EP) {
      float vv[EP][TN], xx[EP][TN];
      int qk[EP];
      float am[EP];
#pragma unroll
      for (int k = 0; k < EP; ++k) {
        qk[k] = (q + k < cnt) ? q + k : cnt - 1;
        am[k] = (q + k < cnt) ? 1.f : 0.f;
        const int jq = s_j[w][qk[k]];
        const float* vj = proj + (size_t)jq * F5 + 4 * (size_t)F;
        const float* xj = xbar1 + (size_t)jq * F;
#pragma unroll
        for (int t = 0; t < TN; ++t) {
          const int f = lane + 32 * t;
          const bool ok = (t + 1 < TN) || f < F;
          vv[k][t] = ok ? vj[f] : 0.f;
          xx[k][t] = ok ? xj[f] : 0.f;
        }
      }
      float val[16];                           // [edge k][head h]: this lane's partial of alphabar_fb
#pragma unroll
      for (int u = 0; u < 16; ++u) val[u] = 0.f;
#pragma unroll
      for (int k = 0; k < EP; ++k) {
        const float* aq = &s_a[w][qk[k] * Ast];
#pragma unroll
        for (int t = 0; t < TN; ++t) {
          const float p = xb[t] * vv[k][t];
          if (DHC >= 32) {
            head_accum<(DHC >= 32 ? DHC : 32), HP>(val + k * HP, t, lane, p);
          } else {
#pragma unroll
            for (int h = 0; h < HP; ++h) val[k * HP + h] += (hsel[t] == h) ? p : 0.f;
          }
          accv[t] = fmaf(am[k] * aq[hsel[t]], xx[k][t], accv[t]);
        }
      }
      // halving butterfly: after the stages with offsets 16, 8, 4, 2 a lane holds value index (lane >> 1); the last
      // stage completes the sum over the remaining lane bit
#pragma unroll
      for (int st = 0; st < 4; ++st) {
        const int off = 16 >> st, half = 8 >> st;
        const bool up = (lane & off) != 0;
#pragma unroll
        for (int u = 0; u < half; ++u) {
          const float send = up ? val[u] : val[u + half];
          const float keep = up ? val[u + half] : val[u];
          val[u] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
      }
      val[0] += __shfl_xor_sync(0xffffffffu, val[0], 1);
      {
        const int idx = lane >> 1, k = idx / HP, h = idx % HP;
        if ((lane & 1) == 0 && h < H && q + k < cnt) s_o[w][(q + k) * Ast + h] = val[0];
      }
    }
    __syncwarp();
    for (int t = lane; t < cnt * Ast; t += 32) alphabar[(size_t)e0 * Ast + t] = s_o[w][t];
    __syncwarp();
  }
  if (rowok) {
#pragma unroll
    for (int t = 0; t < TN; ++t) {
      int f = lane + 32 * t;
      if (f < F) gproj[(size_t)i * F5 + 4 * (size_t)F + f] = accv[t];
    }
  }
}

// q/k projection gradients of one filter block (blockIdx.y) from its pair-indexed filters w[pid e][f] (kept from the
// forward with SO3K_FLAG_KEEP_FILTERS, else recomputed by k_filter_tc just before):
//   qbar_i[f] = sum_{e in row i} ab_e[h(f)]  w_e[f] k_j[f]        ab_e  = alphabar[e]     phi_e / sqrt(d_h)
//   kbar_i[f] = sum_{e in row i} abr_e[h(f)] w_e[f] q_j[f]        abr_e = alphabar[rev e] phi_e / sqrt(d_h)
// (w and phi are symmetric in e <-> rev e), the cutoff cotangent ebar[e][1] = sum_a alphabar[e][a] alpha[e][a] / phi_e,
// and, for the canonical edge of every pair, the filter cotangent of the PAIR
//   wbar[pid e][f] = ab_e[h(f)] q_i[f] k_j[f] + abr_e[h(f)] q_j[f] k_i[f]
// which the tensor-core kernel of backward part 2 then reads as a contiguous tile.
// Pure streaming: F floats of filter per edge and block, 2F floats of sender q/k rows (L2), no atomics.
template <int TN>                               // TN = ceil(F / 32): feature slots per lane
__global__ void __launch_bounds__(WPB * 32, SO3K_QK_MINB)
k_qk_bwd_stream(So3kDims D, int N, int Ast, const int* __restrict__ rowptr, const int* __restrict__ col,
                const int* __restrict__ rev, const int* __restrict__ pid, const float4* __restrict__ geom,
                const float* __restrict__ proj, const float* __restrict__ alpha, const float* __restrict__ alphabar,
                const float* __restrict__ wst, size_t blk_stride, float* __restrict__ gproj, float* __restrict__ ebar,
                float* __restrict__ wbar) {
  __shared__ float s_ab[WPB][32 * 8];
  __shared__ float s_abr[WPB][32 * 8];
  __shared__ int s_j[WPB][32];
  __shared__ int s_p[WPB][32];              // pair index ; bit 31 set: e is the canonical edge of its pair
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * WPB + w;
  const int blk = blockIdx.y;
  const int F = D.F, H = D.H, nl = D.nl;
  const int heads = blk ? nl : H, aoff = blk ? H : 0;
  const int qoff = blk ? F : 0, koff = 2 * F + (blk ? F : 0);
  const bool rowok = i < N;
  const int beg = rowok ? rowptr[i] : 0, end = rowok ? rowptr[i + 1] : 0;
  const size_t F5 = 5 * (size_t)F;
  const float inv = 1.0f / sqrtf((float)(F / heads));
  wst += blk * blk_stride;
  wbar += blk * blk_stride;
  constexpr int WP = SO3K_QK_W_POLICY;
  const WPolicy wpol = make_w_policy<WP>();
  float aq[TN], ak[TN], qi[TN], ki[TN];
  int hf[TN];
#pragma unroll
  for (int t = 0; t < TN; ++t) {
    aq[t] = ak[t] = 0.f;
    const int f = lane + 32 * t;
    const bool ok = rowok && f < F;
    qi[t] = ok ? proj[(size_t)i * F5 + qoff + f] : 0.f;
    ki[t] = ok ? proj[(size_t)i * F5 + koff + f] : 0.f;
    hf[t] = (f < F) ? f / (F / heads) : 0;
  }
  for (int e0 = beg; e0 < end; e0 += 32) {
    const int e = e0 + lane;
    const int cnt = (end - e0) < 32 ? (end - e0) : 32;
    if (e < end) {
      const int re = rev[e];
      s_j[w][lane] = col[e];
      s_p[w][lane] = pid[e] | (e < re ? (int)0x80000000 : 0);
      const float ph = so3k_cutoff(geom[e].w, D.r_cut);
      const float sc = ph * inv;
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        s_ab[w][lane * 8 + h] = (h < heads) ? alphabar[(size_t)e * Ast + aoff + h] * sc : 0.f;
        s_abr[w][lane * 8 + h] = (h < heads) ? alphabar[(size_t)re * Ast + aoff + h] * sc : 0.f;
      }
      if (blk == 0) {
        float pb = 0.f;
        for (int a = 0; a < H + nl; ++a) pb = fmaf(alphabar[(size_t)e * Ast + a], alpha[(size_t)e * Ast + a], pb);
        ebar[2 * (size_t)e + 1] = ph > 1e-30f ? pb / ph : 0.f;
      }
    }
    __syncwarp();
    // Two edges per batch; the 6 * TN row loads of a batch are pinned (asm volatile) ahead of its first use -- left to
    // itself ptxas sinks the loads towards their first use and the kernel becomes latency-bound.  SO3K_QK_SINGLE = 0 keeps
    // two batches in flight per warp (explicit double buffering, 128 registers, 16 resident warps); the default issues one
    // batch at a time at 96 registers and lets 20 resident warps cover the latency (7.24 -> 7.05 ms per c4 step).
    // Indices beyond the chunk are clamped (valid addresses), their terms masked.
    auto load_batch = [&](int q, float (&wv)[2][TN], float (&kj)[2][TN], float (&qj)[2][TN], int (&pp)[2]) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int qc = (q + k < cnt) ? q + k : cnt - 1;
        pp[k] = s_p[w][qc];
        const float* pj = proj + (size_t)s_j[w][qc] * F5;
        const float* pw = wst + (size_t)(pp[k] & 0x7fffffff) * F;
        const unsigned long long pol = pp[k] < 0 ? wpol.normal : wpol.first;      // canonical edge = first reader
#pragma unroll
        for (int t = 0; t < TN; ++t) {
          const int f = lane + 32 * t;
          const bool ok = (t + 1 < TN) || f < F;
          wv[k][t] = ok ? ld_w<WP>(pw + f, pol) : 0.f;
          kj[k][t] = ok ? ld_row<WP>(pj + koff + f, wpol.last) : 0.f;
          qj[k][t] = ok ? ld_row<WP>(pj + qoff + f, wpol.last) : 0.f;
        }
      }
    };
    auto use_batch = [&](int q, const float (&wv)[2][TN], const float (&kj)[2][TN], const float (&qj)[2][TN],
                         const int (&pp)[2]) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const bool live = q + k < cnt;                              // warp-uniform
        const int qc = live ? q + k : cnt - 1;
        const float m = live ? 1.f : 0.f;
        const float* ab = &s_ab[w][qc * 8];
        const float* abr = &s_abr[w][qc * 8];
        const bool canon = (pp[k] < 0) && live;
        float* wb = wbar + (size_t)(pp[k] & 0x7fffffff) * F;
#pragma unroll
        for (int t = 0; t < TN; ++t) {
          const float a0 = ab[hf[t]], a1 = abr[hf[t]];
          aq[t] = fmaf(m * a0 * wv[k][t], kj[k][t], aq[t]);
          ak[t] = fmaf(m * a1 * wv[k][t], qj[k][t], ak[t]);
          const int f = lane + 32 * t;
          if (canon && ((t + 1 < TN) || f < F)) __stcs(wb + f, fmaf(a0 * qi[t], kj[k][t], a1 * qj[k][t] * ki[t]));
        }
      }
    };
#if SO3K_QK_SINGLE
    float wA[2][TN], kA[2][TN], qA[2][TN];
    int pA[2];
    for (int q = 0; q < cnt; q += 2) {
      load_batch(q, wA, kA, qA, pA);
      use_batch(q, wA, kA, qA, pA);
    }
#else
    float wA[2][TN], kA[2][TN], qA[2][TN], wB[2][TN], kB[2][TN], qB[2][TN];
    int pA[2], pB[2];
    load_batch(0, wA, kA, qA, pA);
    for (int q = 0; q < cnt; q += 4) {
      load_batch(q + 2, wB, kB, qB, pB);
      use_batch(q, wA, kA, qA, pA);
      load_batch(q + 4, wA, kA, qA, pA);
      use_batch(q + 2, wB, kB, qB, pB);
    }
#endif
    __syncwarp();
  }
  if (rowok) {
    float* gi = gproj + (size_t)i * F5;
#pragma unroll
    for (int t = 0; t < TN; ++t) {
      const int f = lane + 32 * t;
      if (f < F) {
        gi[qoff + f] = aq[t];
        gi[koff + f] = ak[t];
      }
    }
  }
}

// chibar[i][m] = chibar1[i][m] - 2 cg[m] sum_{e in row i} c_em (gbar[e][l(m)] + gbar[rev e][l(m)]),  c_e = chi[j]-chi[i]
__global__ void __launch_bounds__(WPB * 32)
k_chi_bwd(So3kDims D, int N, int Gst, const int* __restrict__ rowptr, const int* __restrict__ col,
          const int* __restrict__ rev, const float* __restrict__ chi, const float* __restrict__ gbar,
          const float* __restrict__ chibar1, float* __restrict__ chibar) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * WPB + w;
  const int M = D.M;
  const bool rowok = i < N;
  const int beg = rowok ? rowptr[i] : 0, end = rowok ? rowptr[i + 1] : 0;
  float acc[SO3K_MAXM], ci[SO3K_MAXM];
  for (int m = 0; m < M; ++m) {
    acc[m] = 0.f;
    ci[m] = rowok ? chi[(size_t)i * M + m] : 0.f;
  }
  for (int e = beg + lane; e < end; e += 32) {
    int j = col[e], re = rev[e];
    float gs[SO3K_MAXDEG];
    for (int l = 0; l < D.nl; ++l) gs[l] = gbar[(size_t)e * Gst + l] + gbar[(size_t)re * Gst + l];
    for (int m = 0; m < M; ++m) {
      float c = chi[(size_t)j * M + m] - ci[m];
      acc[m] = fmaf(c, gs[D.m_seg[m]], acc[m]);
    }
  }
  for (int m = 0; m < M; ++m) {
    float s = warp_sum(acc[m]);
    if (rowok && lane == 0) chibar[(size_t)i * M + m] = chibar1[(size_t)i * M + m] - 2.0f * D.m_cg[m] * s;
  }
}

// F_i = sum_{e in row i} (rbar[e] - rbar[rev e]) ; dE_dedges[src e] = rbar[e]
__global__ void __launch_bounds__(WPB * 32)
k_forces(int N, const int* __restrict__ rowptr, const int* __restrict__ rev, const int* __restrict__ src,
         const float* __restrict__ rbar, float* __restrict__ forces, float* __restrict__ dE_dedges) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * WPB + w;
  const bool rowok = i < N;
  const int beg = rowok ? rowptr[i] : 0, end = rowok ? rowptr[i + 1] : 0;
  float fx = 0.f, fy = 0.f, fz = 0.f;
  for (int e = beg + lane; e < end; e += 32) {
    int re = rev[e];
    float ax = rbar[3 * (size_t)e + 0], ay = rbar[3 * (size_t)e + 1], az = rbar[3 * (size_t)e + 2];
    fx += ax - rbar[3 * (size_t)re + 0];
    fy += ay - rbar[3 * (size_t)re + 1];
    fz += az - rbar[3 * (size_t)re + 2];
    if (dE_dedges) {
      size_t s = (size_t)src[e];
      dE_dedges[3 * s + 0] = ax; dE_dedges[3 * s + 1] = ay; dE_dedges[3 * s + 2] = az;
    }
  }
  fx = warp_sum(fx); fy = warp_sum(fy); fz = warp_sum(fz);
  if (rowok && lane == 0 && forces) {
    forces[3 * (size_t)i + 0] = fx; forces[3 * (size_t)i + 1] = fy; forces[3 * (size_t)i + 2] = fz;
  }
}

// stress[b] += sym( sum_{e in rows of structure b} rbar_e (x) r_e ),  r_e = d_e u_e : dE/d(strain), the quantity the
// reference obtains by differentiating through strained positions and cell (observable_function.py:152-186)
__global__ void __launch_bounds__(WPB * 32)
k_stress(int N, int n_per, const int* __restrict__ rowptr, const float4* __restrict__ geom, const float* __restrict__ rbar,
         float* __restrict__ stress) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * WPB + w;
  const bool rowok = i < N;
  const int beg = rowok ? rowptr[i] : 0, end = rowok ? rowptr[i + 1] : 0;
  float s[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};               // xx, yy, zz, xy, xz, yz
  for (int e = beg + lane; e < end; e += 32) {
    const float4 g = geom[e];
    const float rx = g.x * g.w, ry = g.y * g.w, rz = g.z * g.w;
    const float ax = rbar[3 * (size_t)e + 0], ay = rbar[3 * (size_t)e + 1], az = rbar[3 * (size_t)e + 2];
    s[0] = fmaf(ax, rx, s[0]);
    s[1] = fmaf(ay, ry, s[1]);
    s[2] = fmaf(az, rz, s[2]);
    s[3] += 0.5f * (ax * ry + ay * rx);
    s[4] += 0.5f * (ax * rz + az * rx);
    s[5] += 0.5f * (ay * rz + az * ry);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) s[k] = warp_sum(s[k]);
  if (rowok && lane == 0 && beg < end) {
    float* o = stress + 9 * (size_t)(i / n_per);
    atomicAdd(o + 0, s[0]); atomicAdd(o + 4, s[1]); atomicAdd(o + 8, s[2]);
    atomicAdd(o + 1, s[3]); atomicAdd(o + 3, s[3]);
    atomicAdd(o + 2, s[4]); atomicAdd(o + 6, s[4]);
    atomicAdd(o + 5, s[5]); atomicAdd(o + 7, s[5]);
  }
}

// ZBLRepulsion (mlff/nn/observable/observable.py:131-183), receiver rows:
//   e_edge = 1/2 w(d) ke phi(d) z_i z_j / d * sum_k c_k exp(-a_k d (z_i^p + z_j^p) dpar),  w = switching_fn(d, 0, 1.5)
// and its derivative with respect to d, added to the edge cotangent rbar (E depends on d only: dE/dr = dE/dd u).
struct ZblP { float a[4], c[4], p, dpar; };
__device__ __forceinline__ float zbl_softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }
__device__ __forceinline__ float zbl_sigma(float x, float* ds) {          // exp(-1/x), x > 0 ; derivative exp(-1/x)/x^2
  if (x > 0.f) {
    const float s = expf(-1.0f / x);
    *ds = s / (x * x);
    return s;
  }
  *ds = 0.f;
  return 0.f;
}
__global__ void __launch_bounds__(WPB * 32)
k_zbl(So3kDims D, int N, const int* __restrict__ rowptr, const int* __restrict__ col, const float4* __restrict__ geom,
      const int* __restrict__ z, const float* __restrict__ raw, float out_scale, float atom_shift,
      float* __restrict__ e_atom, float* __restrict__ rbar) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * WPB + w;
  const bool rowok = i < N;
  const int beg = rowok ? rowptr[i] : 0, end = rowok ? rowptr[i + 1] : 0;
  ZblP P;
  float csum = 0.f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    P.a[k] = zbl_softplus(raw[k]);
    P.c[k] = zbl_softplus(raw[4 + k]);
    csum += P.c[k];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) P.c[k] /= csum;
  P.p = zbl_softplus(raw[8]);
  P.dpar = zbl_softplus(raw[9]);
  const float ke = 14.399645351950548f;
  const float zi = rowok ? (float)z[i] : 0.f;
  const float zip = zi > 0.f ? powf(zi, P.p) : 0.f;
  float esum = 0.f;
  for (int e = beg + lane; e < end; e += 32) {
    const float4 g = geom[e];
    const float d = g.w;
    const float zj = (float)z[col[e]];
    if (d > 0.f && zi > 0.f && zj > 0.f) {
      const float cc = d * (1.0f / 1.5f);
      float dA, dB;
      const float A = zbl_sigma(1.0f - cc, &dA), Bv = zbl_sigma(cc, &dB);
      const float den = A + Bv;
      const float wsw = A / den;
      const float dw = ((-dA) * Bv - A * dB) / (den * den) * (1.0f / 1.5f);
      const float ph = so3k_cutoff(d, D.r_cut), dph = so3k_cutoff_d(d, D.r_cut);
      const float s = (zip + powf(zj, P.p)) * P.dpar;
      float y = 0.f, dy = 0.f;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float t = P.c[k] * expf(-P.a[k] * s * d);
        y += t;
        dy -= t * P.a[k] * s;
      }
      const float pre = 0.5f * ke * zi * zj * out_scale;
      const float f = wsw * ph * y / d;
      const float df = (dw * ph * y + wsw * dph * y + wsw * ph * dy) / d - f / d;
      esum = fmaf(pre, f, esum);
      const float gd = pre * df;
      rbar[3 * (size_t)e + 0] += gd * g.x;
      rbar[3 * (size_t)e + 1] += gd * g.y;
      rbar[3 * (size_t)e + 2] += gd * g.z;
    }
  }
  esum = warp_sum(esum);
  if (rowok && lane == 0) e_atom[i] += esum - atom_shift * out_scale;
}

}  // namespace

void so3k_launch_zbl(const So3kDims& D, const Graph& g, const int* z, const float* zbl_raw, float out_scale,
                     float atom_shift, float* e_atom, float* rbar, cudaStream_t st) {
  SO3K_LAUNCH(k_zbl, dim3(so3k_cdiv(g.N, WPB)), dim3(WPB * 32), 0, st, D, g.N, g.rowptr, g.col, g.geom, z, zbl_raw,
              out_scale, atom_shift, e_atom, rbar);
}

void so3k_launch_stress(const Graph& g, int B, int n_per, const float* rbar, float* stress, cudaStream_t st) {
  cudaMemsetAsync(stress, 0, sizeof(float) * 9 * (size_t)B, st);
  SO3K_LAUNCH(k_stress, dim3(so3k_cdiv(g.N, WPB)), dim3(WPB * 32), 0, st, g.N, n_per, g.rowptr, g.geom, rbar, stress);
}

void so3k_launch_aggregate(const So3kDims& D, const Graph& g, const float* x, const float* chi, const float* proj,
                           const float* alpha, float* x1, float* chi1, cudaStream_t st) {
  int Ast = (D.H + D.nl + 3) & ~3;
#define SO3K_AGG_CASE(TN_)                                                                                             \
  case TN_:                                                                                                            \
    SO3K_LAUNCH(k_aggregate<TN_>, dim3(so3k_cdiv(g.N, WPB)), dim3(WPB * 32), 0, st, D, g.N, Ast, g.rowptr, g.col,      \
                g.geom, x, chi, proj, alpha, x1, chi1);                                                                \
    break;
  switch (so3k_cdiv(D.F, 32)) {
    SO3K_AGG_CASE(1) SO3K_AGG_CASE(2) SO3K_AGG_CASE(3) SO3K_AGG_CASE(4) SO3K_AGG_CASE(5) SO3K_AGG_CASE(6)
    SO3K_AGG_CASE(7) SO3K_AGG_CASE(8)
    default: break;
  }
#undef SO3K_AGG_CASE
}

void so3k_launch_alpha_aggregate(const So3kDims& D, const Graph& g, const float* x, const float* chi, const float* proj,
                                 const float* wst_f, const float* wst_g, float* alpha, float* x1, float* chi1,
                                 cudaStream_t st) {
  int Ast = (D.H + D.nl + 3) & ~3;
  if (so3k_scan_ok(D.F, D.H) && so3k_scan_ok(D.F, D.nl)) {
    if (D.F > 128)
      SO3K_LAUNCH(k_alpha_aggregate_scan<true>, dim3(so3k_cdiv(g.N, WPB)), dim3(WPB * 32), 0, st, D, g.N, Ast, g.rowptr, g.col,
                  g.pid, g.geom, x, chi, proj, wst_f, wst_g, alpha, x1, chi1);
    else
      SO3K_LAUNCH(k_alpha_aggregate_scan<false>, dim3(so3k_cdiv(g.N, WPB)), dim3(WPB * 32), 0, st, D, g.N, Ast, g.rowptr, g.col,
                  g.pid, g.geom, x, chi, proj, wst_f, wst_g, alpha, x1, chi1);
    return;
  }
#define SO3K_AA_CASE(TN_)                                                                                              \
  case TN_:                                                                                                            \
    if (D.H <= 4 && D.nl <= 4)                                                                                         \
      SO3K_LAUNCH((k_alpha_aggregate<TN_, 4, 0, 0>), dim3(so3k_cdiv(g.N, WPB)), dim3(WPB * 32), 0, st, D, g.N, Ast,    \
                  g.rowptr, g.col, g.pid, g.geom, x, chi, proj, wst_f, wst_g, alpha, x1, chi1);                        \
    else                                                                                                               \
      SO3K_LAUNCH((k_alpha_aggregate<TN_, 8, 0, 0>), dim3(so3k_cdiv(g.N, WPB)), dim3(WPB * 32), 0, st, D, g.N, Ast,    \
                  g.rowptr, g.col, g.pid, g.geom, x, chi, proj, wst_f, wst_g, alpha, x1, chi1);                        \
    break;
  // (compile-time head widths, as in so3k_launch_alpha_bwd, bring nothing here: 7.3 ms per step either way -- the kernel
  //  is bound by its 50 row loads per iteration, and with cheaper arithmetic ptxas keeps fewer of them in flight)
  switch (so3k_cdiv(D.F, 32)) {
    SO3K_AA_CASE(1) SO3K_AA_CASE(2) SO3K_AA_CASE(3) SO3K_AA_CASE(4) SO3K_AA_CASE(5) SO3K_AA_CASE(6) SO3K_AA_CASE(7)
    SO3K_AA_CASE(8)
    default: break;
  }
#undef SO3K_AA_CASE
}

void so3k_launch_alpha_bwd(const So3kDims& D, const Graph& g, const float* proj, const float* alpha,
                           const float* xbar1, const float* chibar1, float* alphabar, float* gproj, cudaStream_t st) {
  int Ast = (D.H + D.nl + 3) & ~3;
#define SO3K_AB_CASE(TN_)                                                                                              \
  case TN_:                                                                                                            \
    if (D.H <= 4)                                                                                                      \
      SO3K_LAUNCH((k_alpha_bwd<TN_, 4, 0>), dim3(so3k_cdiv(g.N, WPB)), dim3(WPB * 32), 0, st, D, g.N, Ast, g.rowptr,   \
                  g.col, g.rev, g.geom, proj, alpha, xbar1, chibar1, alphabar, gproj);                                 \
    else                                                                                                               \
      SO3K_LAUNCH((k_alpha_bwd<TN_, 8, 0>), dim3(so3k_cdiv(g.N, WPB)), dim3(WPB * 32), 0, st, D, g.N, Ast, g.rowptr,   \
                  g.col, g.rev, g.geom, proj, alpha, xbar1, chibar1, alphabar, gproj);                                 \
    break;
  if (D.F == 132 && D.H == 4) {                    // the reference's default shape: head boundaries known at compile time
    SO3K_LAUNCH((k_alpha_bwd<5, 4, 33>), dim3(so3k_cdiv(g.N, WPB)), dim3(WPB * 32), 0, st, D, g.N, Ast, g.rowptr, g.col,
                g.rev, g.geom, proj, alpha, xbar1, chibar1, alphabar, gproj);
    return;
  }
  switch (so3k_cdiv(D.F, 32)) {
    SO3K_AB_CASE(1) SO3K_AB_CASE(2) SO3K_AB_CASE(3) SO3K_AB_CASE(4) SO3K_AB_CASE(5) SO3K_AB_CASE(6) SO3K_AB_CASE(7)
    SO3K_AB_CASE(8)
    default: break;
  }
#undef SO3K_AB_CASE
}

// wst / wbar: [block][pair capacity][F] (block stride = so3k_pair_capacity(Pt) * F floats)
void so3k_launch_qk_bwd_stream(const So3kDims& D, const Graph& g, const float* proj, const float* alpha,
                               const float* alphabar, const float* wst, float* gproj, float* ebar, float* wbar,
                               cudaStream_t st) {
  int Ast = (D.H + D.nl + 3) & ~3;
  const size_t blk_stride = (size_t)so3k_pair_capacity(g.Pt) * D.F;
#define SO3K_QK_CASE(TN_)                                                                                              \
  case TN_:                                                                                                            \
    SO3K_LAUNCH(k_qk_bwd_stream<TN_>, dim3(so3k_cdiv(g.N, WPB), 2), dim3(WPB * 32), 0, st, D, g.N, Ast, g.rowptr,      \
                g.col, g.rev, g.pid, g.geom, proj, alpha, alphabar, wst, blk_stride, gproj, ebar, wbar);               \
    break;
  switch (so3k_cdiv(D.F, 32)) {
    SO3K_QK_CASE(1) SO3K_QK_CASE(2) SO3K_QK_CASE(3) SO3K_QK_CASE(4) SO3K_QK_CASE(5) SO3K_QK_CASE(6) SO3K_QK_CASE(7)
    SO3K_QK_CASE(8)
    default: break;
  }
#undef SO3K_QK_CASE
}

void so3k_launch_chi_bwd(const So3kDims& D, const Graph& g, const float* chi, const float* gbar,
                         const float* chibar1, float* chibar, cudaStream_t st) {
  int Gst = (D.nl + 3) & ~3;
  SO3K_LAUNCH(k_chi_bwd, dim3(so3k_cdiv(g.N, WPB)), dim3(WPB * 32), 0, st, D, g.N, Gst, g.rowptr, g.col, g.rev, chi, gbar,
              chibar1, chibar);
}

void so3k_launch_forces(const So3kDims& D, const Graph& g, int P_slots, const float* rbar, float* forces,
                        float* dE_dedges, cudaStream_t st) {
  (void)D;
  if (dE_dedges) cudaMemsetAsync(dE_dedges, 0, sizeof(float) * 3 * (size_t)P_slots, st);
  SO3K_LAUNCH(k_forces, dim3(so3k_cdiv(g.N, WPB)), dim3(WPB * 32), 0, st, g.N, g.rowptr, g.rev, g.src, rbar, forces,
              dE_dedges);
}
