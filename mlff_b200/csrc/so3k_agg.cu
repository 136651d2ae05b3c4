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

constexpr int WPB = 8;          // warps (rows) per block
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
__device__ __forceinline__ float ld_pinned(const float* p) {
#ifdef SO3K_EMU
  return *p;
#else
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
#endif
}

__device__ __forceinline__ float4 ld_pinned4(const float* p) {
#ifdef SO3K_EMU
  return *reinterpret_cast<const float4*>(p);
#else
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
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

// ---- lane <-> feature map of the vectorised streaming kernels ---------------------------------------------------------
// A lane owns NV 16-byte chunks of an F-wide row -- chunk index lane + 32 s, features 4 (lane + 32 s) .. + 3 -- and, with
// TAIL, one scalar of the remainder (feature 128 NV + lane).  F = 132 is <NV = 1, TAIL>: ONE 128-bit load per lane covers
// features 0..127 and lanes 0..3 add features 128..131 with a predicated 32-bit load -- two load instructions per row
// where a lane-per-feature layout needs five.  Element e of a lane: e < 4 NV vector element, e == 4 NV the tail scalar.
template <int NV, bool TAIL>
struct LaneMap {
  static constexpr int E = 4 * NV + (TAIL ? 1 : 0);
  __device__ __forceinline__ static int feature(int lane, int e) {
    return e < 4 * NV ? 4 * (lane + 32 * (e >> 2)) + (e & 3) : 128 * NV + lane;
  }
  // row p (16-byte aligned, F % 4 == 0) -> v[E]; elements beyond F read as 0
  __device__ __forceinline__ static void load(const float* __restrict__ p, int lane, int F, float* v) {
#pragma unroll
    for (int s = 0; s < NV; ++s) {
      const int c = lane + 32 * s;
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (4 * c < F) t = ld_pinned4(p + 4 * c);
      v[4 * s] = t.x; v[4 * s + 1] = t.y; v[4 * s + 2] = t.z; v[4 * s + 3] = t.w;
    }
    if (TAIL) {
      const int f = 128 * NV + lane;
      v[4 * NV] = (f < F) ? ld_pinned(p + f) : 0.f;
    }
  }
  __device__ __forceinline__ static void store(float* __restrict__ p, int lane, int F, const float* v) {
#pragma unroll
    for (int s = 0; s < NV; ++s) {
      const int c = lane + 32 * s;
      if (4 * c < F) *reinterpret_cast<float4*>(p + 4 * c) = make_float4(v[4 * s], v[4 * s + 1], v[4 * s + 2], v[4 * s + 3]);
    }
    if (TAIL) {
      const int f = 128 * NV + lane;
      if (f < F) p[f] = v[4 * NV];
    }
  }
};
// floats of product staging per (edge, block) and warp: F rounded up to 8, + 8 to move successive rows off the bank period
SO3K_HD static inline int so3k_spw(int F) { return ((F + 7) & ~7) + 8; }

// Who reduces what after a batch of EPB edges has staged its per-feature products in shared memory: NOUT = EPB * A head
// sums (A heads per edge).  With NOUT <= 16 two lanes share a sum (halves of the head's feature range, combined by one
// shuffle), otherwise one lane per sum.  All fields are loop-invariant per lane.
struct HeadReducer {
  int k, a, f0, len;        // edge of the batch, head slot, feature range [f0, f0 + len)
  bool split, owner;        // two lanes per sum ; this lane writes the result
  // sum of the staged products of this lane's range: four independent chains, so the shared-memory loads pipeline
  // instead of forming one load -> add -> load chain
  __device__ __forceinline__ float sum(const float* __restrict__ s_w) const {
    const float* sp = s_w + f0;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int t = 0;
    for (; t + 4 <= len; t += 4) {
      s0 += sp[t];
      s1 += sp[t + 1];
      s2 += sp[t + 2];
      s3 += sp[t + 3];
    }
    if (t < len) s0 += sp[t];
    if (t + 1 < len) s1 += sp[t + 1];
    if (t + 2 < len) s2 += sp[t + 2];
    return (s0 + s1) + (s2 + s3);
  }
};

// Tensor-core path: attention coefficients from the pair-indexed filters, fused with both aggregations.
//   alpha_fb[e][h] = phi_e / sqrt(F/H)  sum_{f in h} w_fb[pid e][f] q_fb[i][f] k_fb[j][f]       (so3krates_layer.py:590-606)
//   alpha_gb[e][l] = phi_e / sqrt(F/nl) sum_{f in l} w_gb[pid e][f] q_gb[i][f] k_gb[j][f]       (so3krates_layer.py:545-562)
//   x1 = x + sum_e alpha_fb[e][h(f)] v[j][f] ;  chi1 = chi + sum_e alpha_gb[e][l(m)] Y_m(u_e)
// A warp owns a receiver row; two edges per pass.  Their five rows each (two filter rows, the sender's k_fb, k_gb, v)
// arrive as 128-bit lane chunks (LaneMap); the per-feature products q w k go to a shared-memory staging row in feature
// order, and the 2 (H + nl) head sums are formed by lanes that each add up (half of) one head's contiguous range -- a
// dozen shared-memory reads instead of a select-and-add per (feature, head) and a 16-value butterfly.
template <int NV, bool TAIL>
__global__ void __launch_bounds__(WPB * 32)
k_alpha_aggregate(So3kDims D, int N, int Ast, const int* __restrict__ rowptr, const int* __restrict__ col,
                  const int* __restrict__ pid, const float4* __restrict__ geom, const float* __restrict__ x,
                  const float* __restrict__ chi, const float* __restrict__ proj, const float* __restrict__ wst_f,
                  const float* __restrict__ wst_g, float* __restrict__ alpha, float* __restrict__ x1,
                  float* __restrict__ chi1) {
  using LM = LaneMap<NV, TAIL>;
  constexpr int E = LM::E, EPB = 2;
  __shared__ float s_a[WPB][32 * AMAX];
  __shared__ float s_ph[WPB][32];
  __shared__ int s_j[WPB][32];
  __shared__ int s_p[WPB][32];
  SO3K_DYN_SMEM(float, s_w_all);             // [WPB][EPB * 2][SPW]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * WPB + w;
  const int F = D.F, M = D.M, H = D.H, nl = D.nl, A = H + nl;
  const int SPW = so3k_spw(F);
  float* s_w = s_w_all + (size_t)w * EPB * 2 * SPW;
  const bool rowok = i < N;
  const int beg = rowok ? rowptr[i] : 0, end = rowok ? rowptr[i + 1] : 0;
  const size_t F5 = 5 * (size_t)F;
  const float inv_f = 1.0f / sqrtf((float)(F / H)), inv_g = 1.0f / sqrtf((float)(F / nl));
  float accx[E], qf[E], qg[E];
  int hx[E];
  if (rowok) {
    LM::load(proj + (size_t)i * F5, lane, F, qf);
    LM::load(proj + (size_t)i * F5 + F, lane, F, qg);
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    accx[e] = 0.f;
    if (!rowok) qf[e] = qg[e] = 0.f;
    const int f = LM::feature(lane, e);
    hx[e] = (f < F) ? f / (F / H) : 0;
  }
  // this lane's share of the head sums of a batch
  HeadReducer hr;
  {
    const int nout = EPB * A;
    hr.split = nout <= 16;
    const int o = hr.split ? (lane & 15) : lane, half = hr.split ? (lane >> 4) : 0;
    hr.k = o / A;
    hr.a = o - hr.k * A;
    const bool ok = o < nout;
    const int blk = hr.a >= H, h = blk ? hr.a - H : hr.a, dh = blk ? F / nl : F / H;
    const int part = hr.split ? (dh + 1) / 2 : dh;
    hr.f0 = (hr.k * 2 + blk) * SPW + h * dh + half * part;
    hr.len = ok ? (half ? dh - part : part) : 0;
    hr.owner = ok && half == 0;
  }
  const float hr_scale = hr.a >= H ? inv_g : inv_f;
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
      float vv[EPB][E];
      int qk[EPB];
      {
        float wf[EPB][E], wg[EPB][E], kf[EPB][E], kg[EPB][E];
#pragma unroll
        for (int k = 0; k < EPB; ++k) {
          qk[k] = (q + k < cnt) ? q + k : cnt - 1;
          const float* pj = proj + (size_t)s_j[w][qk[k]] * F5;
          LM::load(wst_f + (size_t)s_p[w][qk[k]] * F, lane, F, wf[k]);
          LM::load(wst_g + (size_t)s_p[w][qk[k]] * F, lane, F, wg[k]);
          LM::load(pj + 2 * F, lane, F, kf[k]);
          LM::load(pj + 3 * F, lane, F, kg[k]);
          LM::load(pj + 4 * F, lane, F, vv[k]);
        }
#pragma unroll
        for (int k = 0; k < EPB; ++k) {
          float pf[E], pg[E];
#pragma unroll
          for (int e2 = 0; e2 < E; ++e2) {
            pf[e2] = wf[k][e2] * (qf[e2] * kf[k][e2]);
            pg[e2] = wg[k][e2] * (qg[e2] * kg[k][e2]);
          }
          LM::store(&s_w[(k * 2 + 0) * SPW], lane, F, pf);
          LM::store(&s_w[(k * 2 + 1) * SPW], lane, F, pg);
        }
      }
      __syncwarp();
      {
        float sum = hr.sum(s_w);
        if (hr.split) sum += __shfl_xor_sync(0xffffffffu, sum, 16);
        if (hr.owner && q + hr.k < cnt) s_a[w][(q + hr.k) * Ast + hr.a] = sum * hr_scale * s_ph[w][q + hr.k];
      }
      __syncwarp();
#pragma unroll
      for (int k = 0; k < EPB; ++k) {
        const float m = (q + k < cnt) ? 1.f : 0.f;
        const float* aq = &s_a[w][qk[k] * Ast];
#pragma unroll
        for (int e2 = 0; e2 < E; ++e2) accx[e2] = fmaf(m * aq[hx[e2]], vv[k][e2], accx[e2]);
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
    float xi[E];
    LM::load(x + (size_t)i * F, lane, F, xi);
#pragma unroll
    for (int e2 = 0; e2 < E; ++e2) xi[e2] += accx[e2];
    LM::store(x1 + (size_t)i * F, lane, F, xi);
  }
}

// alphabar_fb[e][h] = sum_{f in h} xbar1[i][f] v[j][f] ; alphabar_gb[e][l] = sum_{m in l} chibar1[i][m] Y_m(u_e)
// gproj[i].v[f]     = sum_{e in row i} alpha_fb[rev e][h(f)] xbar1[col e][f]        (= d/dv_i, sender-indexed sum)
// Four edges per pass (two 128-bit row loads each); their 4 H head sums are formed from products staged in shared
// memory, like in k_alpha_aggregate.
template <int NV, bool TAIL>
__global__ void __launch_bounds__(WPB * 32)
k_alpha_bwd(So3kDims D, int N, int Ast, const int* __restrict__ rowptr, const int* __restrict__ col,
            const int* __restrict__ rev, const float4* __restrict__ geom, const float* __restrict__ proj,
            const float* __restrict__ alpha, const float* __restrict__ xbar1, const float* __restrict__ chibar1,
            float* __restrict__ alphabar, float* __restrict__ gproj) {
  using LM = LaneMap<NV, TAIL>;
  constexpr int E = LM::E, EPB = 4;
  __shared__ float s_a[WPB][32 * AMAX];     // alpha_fb of the reverse edges
  __shared__ float s_o[WPB][32 * AMAX];     // alphabar staging (coalesced store)
  __shared__ int s_j[WPB][32];
  SO3K_DYN_SMEM(float, s_w_all);             // [WPB][EPB][SPW]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * WPB + w;
  const int F = D.F, M = D.M, H = D.H, nl = D.nl;
  const int SPW = so3k_spw(F);
  float* s_w = s_w_all + (size_t)w * EPB * SPW;
  const bool rowok = i < N;
  const int beg = rowok ? rowptr[i] : 0, end = rowok ? rowptr[i + 1] : 0;
  const size_t F5 = 5 * (size_t)F;
  float xb[E], accv[E];
  int hx[E];
  if (rowok) LM::load(xbar1 + (size_t)i * F, lane, F, xb);
#pragma unroll
  for (int e = 0; e < E; ++e) {
    if (!rowok) xb[e] = 0.f;
    accv[e] = 0.f;
    const int f = LM::feature(lane, e);
    hx[e] = (f < F) ? f / (F / H) : 0;
  }
  HeadReducer hr;
  {
    const int nout = EPB * H;                 // <= 32 (H <= 8)
    hr.split = nout <= 16;
    const int o = hr.split ? (lane & 15) : lane, half = hr.split ? (lane >> 4) : 0;
    hr.k = o / H;
    hr.a = o - hr.k * H;
    const bool ok = o < nout;
    const int dh = F / H;
    const int part = hr.split ? (dh + 1) / 2 : dh;
    hr.f0 = hr.k * SPW + hr.a * dh + half * part;
    hr.len = ok ? (half ? dh - part : part) : 0;
    hr.owner = ok && half == 0;
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
        float sm = 0.f;
        for (int m = D.seg_off[l]; m < D.seg_off[l + 1]; ++m) sm = fmaf(cb[m], y[m], sm);
        s_o[w][lane * Ast + H + l] = sm;
      }
      for (int a = H + nl; a < Ast; ++a) s_o[w][lane * Ast + a] = 0.f;
    }
    s_j[w][lane] = j;
    __syncwarp();
    for (int q = 0; q < cnt; q += EPB) {
      float xx[EPB][E];
      int qk[EPB];
      {
        float vv[EPB][E];
#pragma unroll
        for (int k = 0; k < EPB; ++k) {
          qk[k] = (q + k < cnt) ? q + k : cnt - 1;
          const int jq = s_j[w][qk[k]];
          LM::load(proj + (size_t)jq * F5 + 4 * (size_t)F, lane, F, vv[k]);
          LM::load(xbar1 + (size_t)jq * F, lane, F, xx[k]);
        }
#pragma unroll
        for (int k = 0; k < EPB; ++k) {
          float pr[E];
#pragma unroll
          for (int e2 = 0; e2 < E; ++e2) pr[e2] = xb[e2] * vv[k][e2];
          LM::store(&s_w[k * SPW], lane, F, pr);
        }
      }
      __syncwarp();
      {
        float sum = hr.sum(s_w);
        if (hr.split) sum += __shfl_xor_sync(0xffffffffu, sum, 16);
        if (hr.owner && q + hr.k < cnt) s_o[w][(q + hr.k) * Ast + hr.a] = sum;
      }
#pragma unroll
      for (int k = 0; k < EPB; ++k) {
        const float m = (q + k < cnt) ? 1.f : 0.f;
        const float* aq = &s_a[w][qk[k] * Ast];
#pragma unroll
        for (int e2 = 0; e2 < E; ++e2) accv[e2] = fmaf(m * aq[hx[e2]], xx[k][e2], accv[e2]);
      }
      __syncwarp();
    }
    __syncwarp();
    for (int t = lane; t < cnt * Ast; t += 32) alphabar[(size_t)e0 * Ast + t] = s_o[w][t];
    __syncwarp();
  }
  if (rowok) LM::store(gproj + (size_t)i * F5 + 4 * (size_t)F, lane, F, accv);
}

// q/k projection gradients of one filter block (blockIdx.y) from its pair-indexed filters w[pid e][f] (kept from the
// forward with SO3K_FLAG_KEEP_FILTERS, else recomputed by k_filter_tc just before):
//   qbar_i[f] = sum_{e in row i} ab_e[h(f)]  w_e[f] k_j[f]        ab_e  = alphabar[e]     phi_e / sqrt(d_h)
//   kbar_i[f] = sum_{e in row i} abr_e[h(f)] w_e[f] q_j[f]        abr_e = alphabar[rev e] phi_e / sqrt(d_h)
// (w and phi are symmetric in e <-> rev e), the cutoff cotangent ebar[e][1] = sum_a alphabar[e][a] alpha[e][a] / phi_e,
// and, for the canonical edge of every pair, the filter cotangent of the PAIR
//   wbar[pid e][f] = ab_e[h(f)] q_i[f] k_j[f] + abr_e[h(f)] q_j[f] k_i[f]
// which the tensor-core kernel of backward part 2 then reads as a contiguous tile.
// Pure streaming: F floats of filter per edge and block, 2F floats of sender q/k rows (L2), no atomics; rows move as
// 128-bit lane chunks (LaneMap).
template <int NV, bool TAIL>
__global__ void __launch_bounds__(WPB * 32)
k_qk_bwd_stream(So3kDims D, int N, int Ast, const int* __restrict__ rowptr, const int* __restrict__ col,
                const int* __restrict__ rev, const int* __restrict__ pid, const float4* __restrict__ geom,
                const float* __restrict__ proj, const float* __restrict__ alpha, const float* __restrict__ alphabar,
                const float* __restrict__ wst, size_t blk_stride, float* __restrict__ gproj, float* __restrict__ ebar,
                float* __restrict__ wbar) {
  using LM = LaneMap<NV, TAIL>;
  constexpr int E = LM::E;
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
  float aq[E], ak[E], qi[E], ki[E];
  int hf[E];
  if (rowok) {
    LM::load(proj + (size_t)i * F5 + qoff, lane, F, qi);
    LM::load(proj + (size_t)i * F5 + koff, lane, F, ki);
  }
#pragma unroll
  for (int e2 = 0; e2 < E; ++e2) {
    aq[e2] = ak[e2] = 0.f;
    if (!rowok) qi[e2] = ki[e2] = 0.f;
    const int f = LM::feature(lane, e2);
    hf[e2] = (f < F) ? f / (F / heads) : 0;
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
    // Two edges per batch, two batches in flight: the row loads of the NEXT batch are issued before the current one is
    // consumed (explicit double buffering with pinned loads).  Indices beyond the chunk are clamped (valid addresses),
    // their terms masked.
    auto load_batch = [&](int q, float (&wv)[2][E], float (&kj)[2][E], float (&qj)[2][E], int (&pp)[2]) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int qc = (q + k < cnt) ? q + k : cnt - 1;
        pp[k] = s_p[w][qc];
        const float* pj = proj + (size_t)s_j[w][qc] * F5;
        LM::load(wst + (size_t)(pp[k] & 0x7fffffff) * F, lane, F, wv[k]);
        LM::load(pj + koff, lane, F, kj[k]);
        LM::load(pj + qoff, lane, F, qj[k]);
      }
    };
    auto use_batch = [&](int q, const float (&wv)[2][E], const float (&kj)[2][E], const float (&qj)[2][E],
                         const int (&pp)[2]) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const bool live = q + k < cnt;                              // warp-uniform
        const int qc = live ? q + k : cnt - 1;
        const float m = live ? 1.f : 0.f;
        const float* ab = &s_ab[w][qc * 8];
        const float* abr = &s_abr[w][qc * 8];
        const bool canon = (pp[k] < 0) && live;
        float wb[E];
#pragma unroll
        for (int e2 = 0; e2 < E; ++e2) {
          const float a0 = ab[hf[e2]], a1 = abr[hf[e2]];
          aq[e2] = fmaf(m * a0 * wv[k][e2], kj[k][e2], aq[e2]);
          ak[e2] = fmaf(m * a1 * wv[k][e2], qj[k][e2], ak[e2]);
          wb[e2] = fmaf(a0 * qi[e2], kj[k][e2], a1 * qj[k][e2] * ki[e2]);
        }
        if (canon) LM::store(wbar + (size_t)(pp[k] & 0x7fffffff) * F, lane, F, wb);
      }
    };
    float wA[2][E], kA[2][E], qA[2][E], wB[2][E], kB[2][E], qB[2][E];
    int pA[2], pB[2];
    load_batch(0, wA, kA, qA, pA);
    for (int q = 0; q < cnt; q += 4) {
      load_batch(q + 2, wB, kB, qB, pB);
      use_batch(q, wA, kA, qA, pA);
      load_batch(q + 4, wA, kA, qA, pA);
      use_batch(q + 2, wB, kB, qB, pB);
    }
    __syncwarp();
  }
  if (rowok) {
    LM::store(gproj + (size_t)i * F5 + qoff, lane, F, aq);
    LM::store(gproj + (size_t)i * F5 + koff, lane, F, ak);
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

// NV / TAIL of the lane map for an F-wide row (LaneMap): F <= 128: one chunk per lane; 128 < F <= 160: + tail scalar;
// else two chunks per lane (F <= 256)
#define SO3K_LANEMAP_DISPATCH(F_, CALL)                  \
  do {                                                   \
    if ((F_) <= 128) { CALL(1, false); }                 \
    else if ((F_) <= 160) { CALL(1, true); }             \
    else { CALL(2, false); }                             \
  } while (0)

void so3k_launch_alpha_aggregate(const So3kDims& D, const Graph& g, const float* x, const float* chi, const float* proj,
                                 const float* wst_f, const float* wst_g, float* alpha, float* x1, float* chi1,
                                 cudaStream_t st) {
  int Ast = (D.H + D.nl + 3) & ~3;
  const size_t smem = sizeof(float) * WPB * 4 * (size_t)so3k_spw(D.F);      // product staging: 2 edges x 2 blocks per warp
#define SO3K_AA_CALL(NV_, TAIL_)                                                                                     \
  SO3K_SET_MAX_SMEM((k_alpha_aggregate<NV_, TAIL_>), smem);                                                          \
  SO3K_LAUNCH((k_alpha_aggregate<NV_, TAIL_>), dim3(so3k_cdiv(g.N, WPB)), dim3(WPB * 32), smem, st, D, g.N, Ast,     \
              g.rowptr, g.col, g.pid, g.geom, x, chi, proj, wst_f, wst_g, alpha, x1, chi1)
  SO3K_LANEMAP_DISPATCH(D.F, SO3K_AA_CALL);
#undef SO3K_AA_CALL
}

void so3k_launch_alpha_bwd(const So3kDims& D, const Graph& g, const float* proj, const float* alpha,
                           const float* xbar1, const float* chibar1, float* alphabar, float* gproj, cudaStream_t st) {
  int Ast = (D.H + D.nl + 3) & ~3;
  const size_t smem = sizeof(float) * WPB * 4 * (size_t)so3k_spw(D.F);      // product staging: 4 edges per warp
#define SO3K_AB_CALL(NV_, TAIL_)                                                                                      \
  SO3K_SET_MAX_SMEM((k_alpha_bwd<NV_, TAIL_>), smem);                                                                 \
  SO3K_LAUNCH((k_alpha_bwd<NV_, TAIL_>), dim3(so3k_cdiv(g.N, WPB)), dim3(WPB * 32), smem, st, D, g.N, Ast, g.rowptr,  \
              g.col, g.rev, g.geom, proj, alpha, xbar1, chibar1, alphabar, gproj)
  SO3K_LANEMAP_DISPATCH(D.F, SO3K_AB_CALL);
#undef SO3K_AB_CALL
}

// wst / wbar: [block][pair capacity][F] (block stride = so3k_pair_capacity(Pt) * F floats)
void so3k_launch_qk_bwd_stream(const So3kDims& D, const Graph& g, const float* proj, const float* alpha,
                               const float* alphabar, const float* wst, float* gproj, float* ebar, float* wbar,
                               cudaStream_t st) {
  int Ast = (D.H + D.nl + 3) & ~3;
  const size_t blk_stride = (size_t)so3k_pair_capacity(g.Pt) * D.F;
#define SO3K_QK_CALL(NV_, TAIL_)                                                                                       \
  SO3K_LAUNCH((k_qk_bwd_stream<NV_, TAIL_>), dim3(so3k_cdiv(g.N, WPB), 2), dim3(WPB * 32), 0, st, D, g.N, Ast,         \
              g.rowptr, g.col, g.rev, g.pid, g.geom, proj, alpha, alphabar, wst, blk_stride, gproj, ebar, wbar)
  SO3K_LANEMAP_DISPATCH(D.F, SO3K_QK_CALL);
#undef SO3K_QK_CALL
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
