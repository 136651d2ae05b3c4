// Receiver-sorted CSR construction from the reference's padded index lists.
//
// Input conventions (reference): idx_i / idx_j (B,P) local to each structure, padding = -1
// (mlff/nn/stacknet/stacknet.py:131) or >= n (glp sentinel, mlff/mdx/calculator.py:133); an optional
// boolean mask (Graph.mask, mlff/utils/structures.py:5).  Output: edges flattened over the batch,
// padding removed, stably sorted by receiver i (original slot order inside a row), with
//   rowptr (N+1), row/col (receiver / sender, global atom ids), src (original slot), geom (u, d),
//   rev (index of the reverse edge (j,i,-S); the list must be symmetric, as every list the reference
//   builds is -- otherwise SO3K_STATUS_ASYMMETRIC is raised).
// Geometry follows GeometryEmbed (mlff/nn/embed/embed.py:88-119): r = R[j]-R[i] (+ cell_offset @ cell,
// mlff/cutoff_function/pbc.py:17-18) or the supplied displacement, d = |r|, u = r/d (0 where d == 0).
#include "so3k_internal.h"

namespace {

// r = R[j] - R[i] (+ cell_offset @ cell) or the supplied displacement, in the arithmetic every user of the edge shares
// (k_count / k_fill decide with it, k_rowsort_geom stores it): the reverse edge evaluates the exact negative, so both
// directions of a pair always get the same distance.
__device__ __forceinline__ void edge_vec(int e, int i, int j, int b, const float* __restrict__ R,
                                         const float* __restrict__ cell, const int* __restrict__ cell_offset,
                                         const float* __restrict__ disp, float* rx, float* ry, float* rz) {
  if (disp) {
    *rx = disp[3 * (size_t)e + 0]; *ry = disp[3 * (size_t)e + 1]; *rz = disp[3 * (size_t)e + 2];
  } else {
    *rx = R[3 * (size_t)j + 0] - R[3 * (size_t)i + 0];
    *ry = R[3 * (size_t)j + 1] - R[3 * (size_t)i + 1];
    *rz = R[3 * (size_t)j + 2] - R[3 * (size_t)i + 2];
    if (cell && cell_offset) {
      const float* c = cell + 9 * (size_t)b;
      float s0 = (float)cell_offset[3 * (size_t)e + 0];
      float s1 = (float)cell_offset[3 * (size_t)e + 1];
      float s2 = (float)cell_offset[3 * (size_t)e + 2];
      *rx += s0 * c[0] + s1 * c[3] + s2 * c[6];
      *ry += s0 * c[1] + s1 * c[4] + s2 * c[7];
      *rz += s0 * c[2] + s1 * c[5] + s2 * c[8];
    }
  }
}

// A slot is an edge of the graph if its indices are real (not -1 / >= n padding), its mask is set, and -- r_cut > 0 -- its
// length is below the cutoff.  Pairs at d >= r_cut contribute exactly zero to the energy and to every derivative (the
// cosine cutoff and its slope vanish there), so lists built with a skin (MD, domain decomposition: r_cut + skin) cost the
// model kernels nothing for their skin pairs; the results are those of the reference, which carries them as zeros.
__device__ __forceinline__ bool edge_valid(int il, int jl, int n, const unsigned char* mask, int e, int b, float r_cut,
                                           const float* __restrict__ R, const float* __restrict__ cell,
                                           const int* __restrict__ cell_offset, const float* __restrict__ disp) {
  bool ok = (il >= 0) && (il < n) && (jl >= 0) && (jl < n);
  if (mask) ok = ok && (mask[e] != 0);
  if (ok && r_cut > 0.f) {
    float rx, ry, rz;
    edge_vec(e, b * n + il, b * n + jl, b, R, cell, cell_offset, disp, &rx, &ry, &rz);
    ok = sqrtf(rx * rx + ry * ry + rz * rz) < r_cut;
  }
  return ok;
}

__global__ void k_count(int Pt, int n, int P, float r_cut, const int* __restrict__ idx_i, const int* __restrict__ idx_j,
                        const unsigned char* __restrict__ mask, const float* __restrict__ R, const float* __restrict__ cell,
                        const int* __restrict__ cell_offset, const float* __restrict__ disp, int* __restrict__ deg) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < Pt) {
    int b = e / P;
    int il = idx_i[e], jl = idx_j[e];
    if (edge_valid(il, jl, n, mask, e, b, r_cut, R, cell, cell_offset, disp)) atomicAdd(&deg[b * n + il], 1);
  }
}

// ---- exclusive scan (three small passes; N is at most a few million) ----------------------------
constexpr int SCAN_T = 256;
constexpr int SCAN_PER = 4;                 // items per thread
constexpr int SCAN_BLK = SCAN_T * SCAN_PER;

__global__ void k_scan_local(int n, const int* __restrict__ in, int* __restrict__ out, int* __restrict__ bsum) {
  __shared__ int s_warp[SCAN_T / 32];
  int base = blockIdx.x * SCAN_BLK + threadIdx.x * SCAN_PER;
  int v[SCAN_PER];
  int tsum = 0;
  for (int k = 0; k < SCAN_PER; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    tsum += v[k];
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = tsum;
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[w] = inc;
  __syncthreads();
  if (w == 0) {
    int ws = (lane < SCAN_T / 32) ? s_warp[lane] : 0;
    int winc = ws;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < SCAN_T / 32) s_warp[lane] = winc - ws;   // exclusive warp offsets
    if (lane == SCAN_T / 32 - 1) bsum[blockIdx.x] = winc;
  }
  __syncthreads();
  int excl = inc - tsum + s_warp[w];
  for (int k = 0; k < SCAN_PER; ++k) {
    if (base + k < n) out[base + k] = excl;
    excl += v[k];
  }
}

// single block: exclusive scan of the block sums (nb <= a few thousand), sequential chunks of 256
__global__ void k_scan_bsum(int nb, int* __restrict__ bsum, int* __restrict__ total) {
  __shared__ int s_warp[SCAN_T / 32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += SCAN_T) {
    int i = base + threadIdx.x;
    int v = (i < nb) ? bsum[i] : 0;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[w] = inc;
    __syncthreads();
    int woff = 0;
    for (int k = 0; k < w; ++k) woff += s_warp[k];
    int carry = s_carry;
    if (i < nb) bsum[i] = carry + woff + inc - v;
    __syncthreads();
    if (threadIdx.x == SCAN_T - 1) s_carry = carry + woff + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}

__global__ void k_scan_add(int n, int* __restrict__ out, const int* __restrict__ bsum, const int* __restrict__ total) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += bsum[i / SCAN_BLK];
  if (i == 0) out[n] = *total;   // rowptr[N]
}

__global__ void k_fill(int Pt, int n, int P, float r_cut, const int* __restrict__ idx_i, const int* __restrict__ idx_j,
                       const unsigned char* __restrict__ mask, const float* __restrict__ R, const float* __restrict__ cell,
                       const int* __restrict__ cell_offset, const float* __restrict__ disp, const int* __restrict__ rowptr,
                       int* __restrict__ cursor, int* __restrict__ tmp_src) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < Pt) {
    int b = e / P;
    int il = idx_i[e], jl = idx_j[e];
    if (edge_valid(il, jl, n, mask, e, b, r_cut, R, cell, cell_offset, disp)) {
      int i = b * n + il;
      int slot = rowptr[i] + atomicAdd(&cursor[i], 1);
      tmp_src[slot] = e;
    }
  }
}

// Rank-sort every row by original slot (stable order == the caller's order), emit row/col/src and the
// edge geometry.  One thread per sorted slot.
__global__ void k_rowsort_geom(int N, int n, int P, const int* __restrict__ rowptr, const int* __restrict__ deg,
                               const int* __restrict__ tmp_src, const int* __restrict__ idx_j,
                               const float* __restrict__ R, const float* __restrict__ cell,
                               const int* __restrict__ cell_offset, const float* __restrict__ disp,
                               int* __restrict__ row, int* __restrict__ col, int* __restrict__ src,
                               float4* __restrict__ geom, const int* __restrict__ slot_row /* row of unsorted slot */) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  int nvalid = rowptr[N];
  if (s < nvalid) {
    int i = slot_row[s];
    int beg = rowptr[i], end = beg + deg[i];
    int me = tmp_src[s];
    int rank = 0;
    for (int t = beg; t < end; ++t) rank += (tmp_src[t] < me) ? 1 : 0;
    int dst = beg + rank;
    int b = me / P;
    int j = b * n + idx_j[me];
    row[dst] = i;
    col[dst] = j;
    src[dst] = me;
    float rx, ry, rz;
    edge_vec(me, i, j, b, R, cell, cell_offset, disp, &rx, &ry, &rz);
    float d = sqrtf(rx * rx + ry * ry + rz * rz);
    geom[dst] = make_float4(d > 0.f ? rx / d : 0.f, d > 0.f ? ry / d : 0.f, d > 0.f ? rz / d : 0.f, d);
  }
}

// row id of every (unsorted) slot: slot_row[s] = i for s in [rowptr[i], rowptr[i+1])
__global__ void k_slot_row(int N, const int* __restrict__ rowptr, int* __restrict__ slot_row) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
    int beg = rowptr[i], end = rowptr[i + 1];
    for (int s = beg; s < end; ++s) slot_row[s] = i;
  }
}

// reverse edge: for e = (i -> receiver, j -> sender) find e' in row j with col == i whose
// displacement is the negative of e's (handles several periodic images of the same pair).
__global__ void k_rev(int N, const int* __restrict__ rowptr, const int* __restrict__ row, const int* __restrict__ col,
                      const float4* __restrict__ geom, int* __restrict__ rev, int* __restrict__ status) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  int nvalid = rowptr[N];
  if (e < nvalid) {
    int i = row[e], j = col[e];
    float4 g = geom[e];
    float rx = g.x * g.w, ry = g.y * g.w, rz = g.z * g.w;
    int best = -1;
    float bestd = 3.0e38f;
    int beg = rowptr[j], end = rowptr[j + 1];
    for (int t = beg; t < end; ++t) {
      if (col[t] == i) {
        float4 h = geom[t];
        float dx = h.x * h.w + rx, dy = h.y * h.w + ry, dz = h.z * h.w + rz;
        float dd = dx * dx + dy * dy + dz * dz;
        if (dd < bestd) { bestd = dd; best = t; }
      }
    }
    if (best < 0) {
      best = e;   // keep indices in range; results are flagged invalid
      if (status) atomicOr(status, SO3K_STATUS_ASYMMETRIC);
    }
    rev[e] = best;
  }
}

// canonical edges: one representative per undirected pair {e, rev e} (the one with the smaller sorted index)
__global__ void k_canon_flag(int N, const int* __restrict__ rowptr, const int* __restrict__ rev, int* __restrict__ flag,
                             int* __restrict__ pid, int Pt) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < Pt) {
    flag[e] = (e < rowptr[N] && e < rev[e]) ? 1 : 0;
    pid[e] = 0;         // edges without a partner (asymmetric list, flagged by k_rev) keep an in-range index
  }
}
__global__ void k_canon_fill(int Pt, const int* __restrict__ flag, const int* __restrict__ pos, const int* __restrict__ rev,
                             int* __restrict__ canon, int* __restrict__ n_canon, int* __restrict__ pid) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < Pt && flag[e]) {
    const int c = pos[e];
    canon[c] = e;
    const int cc = c < so3k_pair_capacity(Pt) ? c : 0;
    pid[e] = cc;
    pid[rev[e]] = cc;
  }
  if (e == 0) *n_canon = pos[Pt];
}

}  // namespace

size_t so3k_graph_ws_ints(int N, int Pt) {
  // deg (N+1) + cursor (N) + tmp_src (Pt) + slot_row (Pt) + bsum (cdiv(N+1,SCAN_BLK)+1) + total (1)
  return (size_t)(N + 1) + N + 2 * (size_t)Pt + so3k_cdiv(N + 1, SCAN_BLK) + 8 +
         (size_t)(Pt + 1) + so3k_cdiv(Pt + 1, SCAN_BLK) + 16;          // + canonical-edge scan
}

void so3k_build_graph(const So3kDims& D, int B, int n, int P, const float* R, const int* idx_i, const int* idx_j,
                      const float* cell, const int* cell_offset, const float* disp, const unsigned char* mask,
                      Graph& g, int* scratch, int* dev_status, cudaStream_t st) {
  const int N = B * n, Pt = B * P;
  const float r_cut = D.r_cut;
  g.N = N;
  g.Pt = Pt;
  int* deg = scratch;                    // N+1
  int* cursor = deg + (N + 1);           // N
  int* tmp_src = cursor + N;             // Pt
  int* slot_row = tmp_src + Pt;          // Pt
  int nb = so3k_cdiv(N, SCAN_BLK);
  int* bsum = slot_row + Pt;             // nb
  int* total = bsum + nb + 1;
  cudaMemsetAsync(deg, 0, sizeof(int) * (size_t)(2 * N + 1), st);   // deg + cursor
  const int T = 256;
  if (Pt > 0) SO3K_LAUNCH(k_count, dim3(so3k_cdiv(Pt, T)), dim3(T), 0, st, Pt, n, P, r_cut, idx_i, idx_j, mask, R, cell, cell_offset, disp, deg);
  SO3K_LAUNCH(k_scan_local, dim3(nb), dim3(SCAN_T), 0, st, N, deg, g.rowptr, bsum);
  SO3K_LAUNCH(k_scan_bsum, dim3(1), dim3(SCAN_T), 0, st, nb, bsum, total);
  SO3K_LAUNCH(k_scan_add, dim3(so3k_cdiv(N, T)), dim3(T), 0, st, N, g.rowptr, bsum, total);
  g.n_valid = g.rowptr + N;
  if (Pt > 0) {
    SO3K_LAUNCH(k_fill, dim3(so3k_cdiv(Pt, T)), dim3(T), 0, st, Pt, n, P, r_cut, idx_i, idx_j, mask, R, cell, cell_offset, disp, g.rowptr, cursor,
                tmp_src);
    SO3K_LAUNCH(k_slot_row, dim3(so3k_cdiv(N, T)), dim3(T), 0, st, N, g.rowptr, slot_row);
    SO3K_LAUNCH(k_rowsort_geom, dim3(so3k_cdiv(Pt, T)), dim3(T), 0, st, N, n, P, g.rowptr, deg, tmp_src, idx_j, R, cell,
                cell_offset, disp, g.row, g.col, g.src, g.geom, slot_row);
    SO3K_LAUNCH(k_rev, dim3(so3k_cdiv(Pt, T)), dim3(T), 0, st, N, g.rowptr, g.row, g.col, g.geom, g.rev, dev_status);
    // canonical edge list (tmp_src is free again: reuse it for the flags)
    int* flag = tmp_src;
    int* cpos = total + 8;                         // (Pt + 1)
    int nbp = so3k_cdiv(Pt, SCAN_BLK);
    int* cbsum = cpos + Pt + 1;                    // (nbp + 2)
    SO3K_LAUNCH(k_canon_flag, dim3(so3k_cdiv(Pt, T)), dim3(T), 0, st, N, g.rowptr, g.rev, flag, g.pid, Pt);
    SO3K_LAUNCH(k_scan_local, dim3(nbp), dim3(SCAN_T), 0, st, Pt, flag, cpos, cbsum);
    SO3K_LAUNCH(k_scan_bsum, dim3(1), dim3(SCAN_T), 0, st, nbp, cbsum, cbsum + nbp);
    SO3K_LAUNCH(k_scan_add, dim3(so3k_cdiv(Pt, T)), dim3(T), 0, st, Pt, cpos, cbsum, cbsum + nbp);
    SO3K_LAUNCH(k_canon_fill, dim3(so3k_cdiv(Pt, T)), dim3(T), 0, st, Pt, flag, cpos, g.rev, g.canon, g.n_canon, g.pid);
  }
}
