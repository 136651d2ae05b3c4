// Device neighbour list: cell-list (binning) build with warp-ballot compaction, float64 geometry.
//
// Replaces the host-side builders of the reference:
//   pbc_neighbor_list / get_pbc_neighbors  mlff/indexing/indices.py:194-275, 111-191
//       = ase.neighborlist.primitive_neighbor_list('ijS', pbc, cell, positions, cutoff, self_interaction=False):
//         all images S with |R_j - R_i + S@cell| < cutoff (strict), (i != j or S != 0), S = 0 on non-periodic
//         axes, cell rows = lattice vectors                                                      [SO3K_NL_ASE]
//   get_indices                            mlff/indexing/indices.py:56-60
//       non-periodic, 0 < d <= cutoff (inclusive), z_i * z_j != 0, row-major order              [SO3K_NL_DENSE]
//       minimum-image list of the MD path (glp quadratic_neighbor_list, mlff/mdx/atoms.py:488-529): one image per
//       ordered pair i != j, fractional displacement wrapped to [-1/2, 1/2), d < cutoff           [SO3K_NL_MIC]
// Output rows are sorted by i, then by (j, S) -- the reference's order for the dense list, and a canonical order
// for the ASE list (whose consumers are order-independent: segment_sum).  Padding: idx = -1, shifts = 0
// (indices.py:226-231); when more than `capacity` edges exist SO3K_STATUS_OVERFLOW is raised and *count holds
// the true number so the caller can re-allocate (indices.py:264-272).
//
// Pipeline (all on the caller's stream, no host synchronisation):
//   k_bbox      fractional coordinates, their extent along non-periodic axes
//   k_setup     bin grid: nb[a] = floor(extent_a / cutoff) (>= 1, capped), search range ceil(cutoff / bin height)
//   k_bin       wrap periodic coordinates, bin index, per-bin counts        -> scan -> k_bin_fill (bin-sorted atoms)
//   k_pairs<0>  one warp per atom: walk the neighbouring bins, lanes test candidates in float64,
//               __ballot_sync + __popc count                                 -> scan (row offsets, total)
//   k_pairs<1>  same walk, lanes write their hits at rowptr[i] + prefix-popc (warp-ballot compaction)
//   k_rowsort   rank-sort each row by (j, S) into the caller's arrays, pad the tail
// Algorithmic bytes: 24 N (positions) + 20 P (i, j, S as int32).
#include "so3k_internal.h"

#ifdef SO3K_EMU
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
#endif

namespace {

struct NLGrid {
  double inv[9];      // inverse of the (completed) cell: frac = pos @ inv
  double cell[9];     // completed cell, rows = lattice vectors
  double lo[3];       // lower fractional bound of the bin grid
  double span[3];     // fractional extent of the grid
  int nb[3];          // bins per axis
  int ns[3];          // neighbouring bins to visit in each direction
  int pbc[3];
  int nbins;
};

struct NLSetup {
  double cell[9];
  double inv[9];
  double height[3];
  int pbc[3];
  double cutoff;
  int max_bins;
  int mode;
};

constexpr int T = 256;
constexpr int SCAN_PER = 4;
constexpr int SCAN_BLK = T * SCAN_PER;

__device__ __forceinline__ void frac_of(const NLSetup& S, const double* p, double* f) {
  f[0] = p[0] * S.inv[0] + p[1] * S.inv[3] + p[2] * S.inv[6];
  f[1] = p[0] * S.inv[1] + p[1] * S.inv[4] + p[2] * S.inv[7];
  f[2] = p[0] * S.inv[2] + p[1] * S.inv[5] + p[2] * S.inv[8];
}

// single block: min / max of fractional coordinates
__global__ void k_bbox(NLSetup S, int N, const double* __restrict__ pos, double* __restrict__ bbox /* 6 */) {
  __shared__ double s_lo[3][T / 32], s_hi[3][T / 32];
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int i = threadIdx.x; i < N; i += T) {
    double f[3];
    frac_of(S, pos + 3 * (size_t)i, f);
    for (int a = 0; a < 3; ++a) {
      lo[a] = f[a] < lo[a] ? f[a] : lo[a];
      hi[a] = f[a] > hi[a] ? f[a] : hi[a];
    }
  }
  for (int a = 0; a < 3; ++a) {
    for (int o = 16; o > 0; o >>= 1) {
      double l2 = __shfl_xor_sync(0xffffffffu, lo[a], o), h2 = __shfl_xor_sync(0xffffffffu, hi[a], o);
      lo[a] = l2 < lo[a] ? l2 : lo[a];
      hi[a] = h2 > hi[a] ? h2 : hi[a];
    }
    if ((threadIdx.x & 31) == 0) {
      s_lo[a][threadIdx.x >> 5] = lo[a];
      s_hi[a][threadIdx.x >> 5] = hi[a];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int a = 0; a < 3; ++a) {
      double l = s_lo[a][0], h = s_hi[a][0];
      for (int w = 1; w < T / 32; ++w) {
        l = s_lo[a][w] < l ? s_lo[a][w] : l;
        h = s_hi[a][w] > h ? s_hi[a][w] : h;
      }
      bbox[a] = l;
      bbox[3 + a] = h;
    }
  }
}

__global__ void k_setup(NLSetup S, const double* __restrict__ bbox, NLGrid* __restrict__ G) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    NLGrid g;
    for (int k = 0; k < 9; ++k) { g.inv[k] = S.inv[k]; g.cell[k] = S.cell[k]; }
    double ext[3];
    for (int a = 0; a < 3; ++a) {
      g.pbc[a] = S.pbc[a];
      if (S.pbc[a]) { g.lo[a] = 0.0; g.span[a] = 1.0; }
      else {
        double l = bbox[a], h = bbox[3 + a];
        double pad = 1e-9 * (1.0 + (h - l));
        g.lo[a] = l - pad;
        g.span[a] = (h - l) + 2.0 * pad;
      }
      ext[a] = g.span[a] * S.height[a];           // perpendicular extent in length units
      int nb = (int)floor(ext[a] / S.cutoff);
      g.nb[a] = nb < 1 ? 1 : nb;
    }
    // cap the number of bins (sparse / elongated systems)
    for (int it = 0; it < 64; ++it) {
      double tot = (double)g.nb[0] * (double)g.nb[1] * (double)g.nb[2];
      if (tot <= (double)S.max_bins) break;
      int a = 0;
      if (g.nb[1] > g.nb[a]) a = 1;
      if (g.nb[2] > g.nb[a]) a = 2;
      g.nb[a] = (g.nb[a] + 1) / 2;
    }
    for (int a = 0; a < 3; ++a) {
      double hbin = ext[a] / (double)g.nb[a];
      int ns = (int)ceil(S.cutoff / hbin);
      if (!S.pbc[a] && ns > g.nb[a] - 1) ns = g.nb[a] - 1;     // nothing outside the grid
      g.ns[a] = ns < 0 ? 0 : ns;
    }
    g.nbins = g.nb[0] * g.nb[1] * g.nb[2];
    *G = g;
  }
}

__device__ __forceinline__ void bin_coords(const NLGrid& G, const double* f, int* c, int* w) {
  for (int a = 0; a < 3; ++a) {
    double s = f[a];
    int wa = 0;
    if (G.pbc[a]) {
      double fl = floor(s);
      wa = (int)fl;
      s -= fl;                                   // [0,1)
    }
    int b = (int)floor((s - G.lo[a]) / G.span[a] * (double)G.nb[a]);
    if (b < 0) b = 0;
    if (b >= G.nb[a]) b = G.nb[a] - 1;
    c[a] = b;
    w[a] = wa;
  }
}

__global__ void k_bin(int N, const double* __restrict__ pos, const NLGrid* __restrict__ Gp, int* __restrict__ bin_of,
                      int* __restrict__ wrap, int* __restrict__ bin_cnt) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
    const NLGrid& G = *Gp;
    double f[3];
    const double* p = pos + 3 * (size_t)i;
    f[0] = p[0] * G.inv[0] + p[1] * G.inv[3] + p[2] * G.inv[6];
    f[1] = p[0] * G.inv[1] + p[1] * G.inv[4] + p[2] * G.inv[7];
    f[2] = p[0] * G.inv[2] + p[1] * G.inv[5] + p[2] * G.inv[8];
    int c[3], w[3];
    bin_coords(G, f, c, w);
    int b = (c[0] * G.nb[1] + c[1]) * G.nb[2] + c[2];
    bin_of[i] = b;
    wrap[3 * (size_t)i + 0] = w[0]; wrap[3 * (size_t)i + 1] = w[1]; wrap[3 * (size_t)i + 2] = w[2];
    atomicAdd(&bin_cnt[b], 1);
  }
}

// ---- exclusive scan over a fixed-size int array (same scheme as graph.cu) ---------------------------
__global__ void k_scan_local(int n, const int* __restrict__ in, int* __restrict__ out, int* __restrict__ bsum) {
  __shared__ int s_warp[T / 32];
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
    int ws = (lane < T / 32) ? s_warp[lane] : 0;
    int winc = ws;
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    if (lane < T / 32) s_warp[lane] = winc - ws;
    if (lane == T / 32 - 1) bsum[blockIdx.x] = winc;
  }
  __syncthreads();
  int excl = inc - tsum + s_warp[w];
  for (int k = 0; k < SCAN_PER; ++k) {
    if (base + k < n) out[base + k] = excl;
    excl += v[k];
  }
}
__global__ void k_scan_bsum(int nb, int* __restrict__ bsum, int* __restrict__ total) {
  __shared__ int s_warp[T / 32];
  __shared__ int s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < nb; base += T) {
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
    if (threadIdx.x == T - 1) s_carry = carry + woff + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}
__global__ void k_scan_add(int n, int* __restrict__ out, const int* __restrict__ bsum, const int* __restrict__ total) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] += bsum[i / SCAN_BLK];
  if (i == 0) out[n] = *total;
}
void exclusive_scan(int n, const int* in, int* out /* n+1 */, int* bsum, cudaStream_t st) {
  int nb = so3k_cdiv(n, SCAN_BLK);
  SO3K_LAUNCH(k_scan_local, dim3(nb), dim3(T), 0, st, n, in, out, bsum);
  SO3K_LAUNCH(k_scan_bsum, dim3(1), dim3(T), 0, st, nb, bsum, bsum + nb);
  SO3K_LAUNCH(k_scan_add, dim3(so3k_cdiv(n, T)), dim3(T), 0, st, n, out, bsum, bsum + nb);
}

// place atoms into their bins; inside a bin atoms are ordered by index (rank by counting smaller indices
// that fell into the same bin is O(n_bin^2) but bins hold ~10 atoms) -> deterministic
__global__ void k_bin_fill(int N, const int* __restrict__ bin_of, const int* __restrict__ bin_start,
                           int* __restrict__ cursor, int* __restrict__ tmp_atoms) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < N) {
    int b = bin_of[i];
    int slot = bin_start[b] + atomicAdd(&cursor[b], 1);
    tmp_atoms[slot] = i;
  }
}
__global__ void k_bin_sort(int N, const int* __restrict__ bin_of, const int* __restrict__ bin_start,
                           const int* __restrict__ tmp_atoms, const double* __restrict__ pos,
                           const int* __restrict__ wrap, int* __restrict__ atoms, double* __restrict__ spos,
                           int* __restrict__ swrap) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < N) {
    int me = tmp_atoms[s];
    int b = bin_of[me];
    int beg = bin_start[b], end = bin_start[b + 1];
    int rank = 0;
    for (int t = beg; t < end; ++t) rank += (tmp_atoms[t] < me) ? 1 : 0;
    int dst = beg + rank;
    atoms[dst] = me;
    for (int a = 0; a < 3; ++a) {
      spos[3 * (size_t)dst + a] = pos[3 * (size_t)me + a];
      swrap[3 * (size_t)dst + a] = wrap[3 * (size_t)me + a];
    }
  }
}

constexpr int WPB = 8;

// One warp per atom.  FILL == 0: count hits into deg[i].  FILL == 1: write (j, S) at rowptr[i] + running offset.
template <int FILL>
__global__ void __launch_bounds__(WPB * 32)
k_pairs(int N, int mode, double cutoff, const NLGrid* __restrict__ Gp, const double* __restrict__ pos,
        const int* __restrict__ z, const int* __restrict__ bin_of, const int* __restrict__ wrap,
        const int* __restrict__ bin_start, const int* __restrict__ atoms, const double* __restrict__ spos,
        const int* __restrict__ swrap, int* __restrict__ deg, const int* __restrict__ rowptr,
        int* __restrict__ tj, int* __restrict__ ts, long long capacity) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * WPB + w;
  const bool rowok = i < N;
  const NLGrid& G = *Gp;
  int c[3] = {0, 0, 0}, wi[3] = {0, 0, 0};
  double pi[3] = {0, 0, 0};
  int zi = 1;
  if (rowok) {
    int b = bin_of[i];
    c[2] = b % G.nb[2];
    c[1] = (b / G.nb[2]) % G.nb[1];
    c[0] = b / (G.nb[2] * G.nb[1]);
    for (int a = 0; a < 3; ++a) {
      wi[a] = wrap[3 * (size_t)i + a];
      pi[a] = pos[3 * (size_t)i + a];
    }
    if (z) zi = z[i];
  }
  int found = 0;
  const long long base = (FILL && rowok) ? (long long)rowptr[i] : 0;
  const int n0 = rowok ? G.ns[0] : -1, n1 = G.ns[1], n2 = G.ns[2];
  for (int d0 = -n0; d0 <= n0; ++d0) {
    int b0 = c[0] + d0, t0 = 0;
    if (G.pbc[0]) { t0 = (b0 >= 0) ? b0 / G.nb[0] : -((-b0 + G.nb[0] - 1) / G.nb[0]); b0 -= t0 * G.nb[0]; }
    else if (b0 < 0 || b0 >= G.nb[0]) continue;
    for (int d1 = -n1; d1 <= n1; ++d1) {
      int b1 = c[1] + d1, t1 = 0;
      if (G.pbc[1]) { t1 = (b1 >= 0) ? b1 / G.nb[1] : -((-b1 + G.nb[1] - 1) / G.nb[1]); b1 -= t1 * G.nb[1]; }
      else if (b1 < 0 || b1 >= G.nb[1]) continue;
      for (int d2 = -n2; d2 <= n2; ++d2) {
        int b2 = c[2] + d2, t2 = 0;
        if (G.pbc[2]) { t2 = (b2 >= 0) ? b2 / G.nb[2] : -((-b2 + G.nb[2] - 1) / G.nb[2]); b2 -= t2 * G.nb[2]; }
        else if (b2 < 0 || b2 >= G.nb[2]) continue;
        const int bb = (b0 * G.nb[1] + b1) * G.nb[2] + b2;
        const int beg = bin_start[bb], end = bin_start[bb + 1];
        for (int s0 = beg; s0 < end; s0 += 32) {
          const int s = s0 + lane;
          bool hit = false;
          int j = -1, S0 = 0, S1 = 0, S2 = 0;
          if (s < end) {
            j = atoms[s];
            S0 = t0 - swrap[3 * (size_t)s + 0] + wi[0];
            S1 = t1 - swrap[3 * (size_t)s + 1] + wi[1];
            S2 = t2 - swrap[3 * (size_t)s + 2] + wi[2];
            // distance_vector = positions[j] - positions[i] + S @ cell   (float64, no contraction)
            double ox = __dadd_rn(__dadd_rn(__dmul_rn((double)S0, G.cell[0]), __dmul_rn((double)S1, G.cell[3])), __dmul_rn((double)S2, G.cell[6]));
            double oy = __dadd_rn(__dadd_rn(__dmul_rn((double)S0, G.cell[1]), __dmul_rn((double)S1, G.cell[4])), __dmul_rn((double)S2, G.cell[7]));
            double oz = __dadd_rn(__dadd_rn(__dmul_rn((double)S0, G.cell[2]), __dmul_rn((double)S1, G.cell[5])), __dmul_rn((double)S2, G.cell[8]));
            double dx = __dadd_rn(__dadd_rn(spos[3 * (size_t)s + 0], -pi[0]), ox);
            double dy = __dadd_rn(__dadd_rn(spos[3 * (size_t)s + 1], -pi[1]), oy);
            double dz = __dadd_rn(__dadd_rn(spos[3 * (size_t)s + 2], -pi[2]), oz);
            double d = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
            if (mode == SO3K_NL_ASE) {
              hit = (d < cutoff) && !(j == i && S0 == 0 && S1 == 0 && S2 == 0);
            } else if (mode == SO3K_NL_MIC) {
              // glp quadratic_neighbor_list + make_displacement (mlff/mdx/atoms.py:488-529): ONE image per ordered pair,
              // the one whose fractional displacement lies in [-1/2, 1/2) on the periodic axes (frac -= floor(frac + 1/2))
              const double f0 = dx * G.inv[0] + dy * G.inv[3] + dz * G.inv[6];
              const double f1 = dx * G.inv[1] + dy * G.inv[4] + dz * G.inv[7];
              const double f2 = dx * G.inv[2] + dy * G.inv[5] + dz * G.inv[8];
              const bool m0 = !G.pbc[0] || (f0 >= -0.5 && f0 < 0.5);
              const bool m1 = !G.pbc[1] || (f1 >= -0.5 && f1 < 0.5);
              const bool m2 = !G.pbc[2] || (f2 >= -0.5 && f2 < 0.5);
              hit = (d < cutoff) && (j != i) && m0 && m1 && m2;
            } else {
              int zj = z ? z[j] : 1;
              hit = (d <= cutoff) && (d > 0.0) && (zi != 0) && (zj != 0);
            }
          }
          unsigned bal = __ballot_sync(0xffffffffu, hit ? 1 : 0);
          if (FILL) {
            if (hit) {
              long long slot = base + found + __popc(bal & ((1u << lane) - 1u));
              if (slot < capacity) {
                tj[slot] = j;
                ts[3 * (size_t)slot + 0] = S0; ts[3 * (size_t)slot + 1] = S1; ts[3 * (size_t)slot + 2] = S2;
              }
            }
          }
          found += __popc(bal);
        }
      }
    }
  }
  if (!FILL && rowok && lane == 0) deg[i] = found;
}

__device__ __forceinline__ unsigned long long nl_key(int j, int s0, int s1, int s2) {
  return ((unsigned long long)(unsigned)j << 30) | ((unsigned long long)(s0 + 512) << 20) |
         ((unsigned long long)(s1 + 512) << 10) | (unsigned long long)(s2 + 512);
}

// rank-sort each row by (j, S) into the output arrays ; pad the tail ; publish count / overflow
__global__ void k_rowsort(int N, const int* __restrict__ rowptr, const int* __restrict__ tj, const int* __restrict__ ts,
                          long long capacity, int* __restrict__ idx_i, int* __restrict__ idx_j, int* __restrict__ shifts,
                          long long* __restrict__ count, int* __restrict__ status) {
  long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = rowptr[N];
  if (s == 0) {
    if (count) *count = total;
    if (total > capacity && status) atomicOr(status, SO3K_STATUS_OVERFLOW);
  }
  if (s < capacity) {
    if (s < total) {
      // binary search the row of slot s
      int lo = 0, hi = N;                      // rowptr[lo] <= s < rowptr[hi]
      while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if ((long long)rowptr[mid] <= s) lo = mid; else hi = mid;
      }
      const int i = lo;
      long long beg = rowptr[i], end = rowptr[i + 1];
      if (end > capacity) end = capacity;
      unsigned long long me = nl_key(tj[s], ts[3 * s], ts[3 * s + 1], ts[3 * s + 2]);
      int rank = 0;
      for (long long t = beg; t < end; ++t)
        rank += (nl_key(tj[t], ts[3 * t], ts[3 * t + 1], ts[3 * t + 2]) < me) ? 1 : 0;
      long long dst = beg + rank;
      idx_i[dst] = i;
      idx_j[dst] = tj[s];
      shifts[3 * dst + 0] = ts[3 * s + 0]; shifts[3 * dst + 1] = ts[3 * s + 1]; shifts[3 * dst + 2] = ts[3 * s + 2];
    } else {
      idx_i[s] = -1;
      idx_j[s] = -1;
      shifts[3 * s + 0] = 0; shifts[3 * s + 1] = 0; shifts[3 * s + 2] = 0;
    }
  }
}

inline size_t a256(size_t n) { return (n + 255) & ~(size_t)255; }

struct NLWs {
  NLGrid* grid; double* bbox;
  int *bin_of, *wrap, *bin_cnt, *bin_start, *cursor, *tmp_atoms, *atoms, *swrap, *deg, *rowptr, *bsum, *tj, *ts;
  double* spos;
  size_t bytes;
  int max_bins;
};

NLWs carve_nl(long long N, long long capacity, void* base) {
  NLWs w;
  char* p = (char*)base;
  auto take = [&](size_t bytes) { char* r = p; p += a256(bytes); return (void*)r; };
  w.max_bins = (int)(2 * N + 1024);
  size_t cap = capacity > 0 ? (size_t)capacity : 1;
  w.grid = (NLGrid*)take(sizeof(NLGrid));
  w.bbox = (double*)take(sizeof(double) * 8);
  w.spos = (double*)take(sizeof(double) * 3 * N);
  w.bin_of = (int*)take(sizeof(int) * N);
  w.wrap = (int*)take(sizeof(int) * 3 * N);
  w.swrap = (int*)take(sizeof(int) * 3 * N);
  w.bin_cnt = (int*)take(sizeof(int) * (2 * (size_t)w.max_bins + 2));   // bin_cnt + cursor, zeroed together
  w.cursor = w.bin_cnt + w.max_bins + 1;
  w.bin_start = (int*)take(sizeof(int) * (w.max_bins + 2));
  w.tmp_atoms = (int*)take(sizeof(int) * N);
  w.atoms = (int*)take(sizeof(int) * N);
  w.deg = (int*)take(sizeof(int) * (N + 1));
  w.rowptr = (int*)take(sizeof(int) * (N + 2));
  size_t nscan = (size_t)(w.max_bins > N ? w.max_bins : N) + 1;
  w.bsum = (int*)take(sizeof(int) * (nscan / SCAN_BLK + 8));
  w.tj = (int*)take(sizeof(int) * cap);
  w.ts = (int*)take(sizeof(int) * 3 * cap);
  w.bytes = (size_t)(p - (char*)base);
  return w;
}

bool complete_and_invert(const double* cell_in, const unsigned char* pbc, NLSetup* S) {
  double c[9];
  for (int k = 0; k < 9; ++k) c[k] = cell_in ? cell_in[k] : 0.0;
  auto norm = [&](int a) { return sqrt(c[3 * a] * c[3 * a] + c[3 * a + 1] * c[3 * a + 1] + c[3 * a + 2] * c[3 * a + 2]); };
  // ase.geometry.complete_cell: replace missing (zero) lattice vectors by unit vectors orthogonal to the others
  for (int a = 0; a < 3; ++a) {
    if (norm(a) > 1e-12) continue;
    if (pbc && pbc[a]) return false;           // periodic axis without a lattice vector
    // candidates: standard basis vectors, Gram-Schmidt against the existing non-zero vectors
    double best[3] = {0, 0, 0}, bestn = -1.0;
    for (int e = 0; e < 3; ++e) {
      double v[3] = {e == 0 ? 1.0 : 0.0, e == 1 ? 1.0 : 0.0, e == 2 ? 1.0 : 0.0};
      for (int b = 0; b < 3; ++b) {
        if (b == a) continue;
        double nb2 = c[3 * b] * c[3 * b] + c[3 * b + 1] * c[3 * b + 1] + c[3 * b + 2] * c[3 * b + 2];
        if (nb2 < 1e-24) continue;
        double dot = (v[0] * c[3 * b] + v[1] * c[3 * b + 1] + v[2] * c[3 * b + 2]) / nb2;
        for (int k = 0; k < 3; ++k) v[k] -= dot * c[3 * b + k];
      }
      double n2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
      if (n2 > bestn) { bestn = n2; best[0] = v[0]; best[1] = v[1]; best[2] = v[2]; }
    }
    if (bestn < 1e-12) return false;
    double n = sqrt(bestn);
    for (int k = 0; k < 3; ++k) c[3 * a + k] = best[k] / n;
  }
  double det = c[0] * (c[4] * c[8] - c[5] * c[7]) - c[1] * (c[3] * c[8] - c[5] * c[6]) + c[2] * (c[3] * c[7] - c[4] * c[6]);
  if (fabs(det) < 1e-12) return false;
  double inv[9];
  inv[0] = (c[4] * c[8] - c[5] * c[7]) / det; inv[1] = (c[2] * c[7] - c[1] * c[8]) / det; inv[2] = (c[1] * c[5] - c[2] * c[4]) / det;
  inv[3] = (c[5] * c[6] - c[3] * c[8]) / det; inv[4] = (c[0] * c[8] - c[2] * c[6]) / det; inv[5] = (c[2] * c[3] - c[0] * c[5]) / det;
  inv[6] = (c[3] * c[7] - c[4] * c[6]) / det; inv[7] = (c[1] * c[6] - c[0] * c[7]) / det; inv[8] = (c[0] * c[4] - c[1] * c[3]) / det;
  for (int k = 0; k < 9; ++k) { S->cell[k] = c[k]; S->inv[k] = inv[k]; }
  // face-to-face heights: 1 / |column a of inv|
  for (int a = 0; a < 3; ++a) {
    double n = sqrt(inv[a] * inv[a] + inv[3 + a] * inv[3 + a] + inv[6 + a] * inv[6 + a]);
    S->height[a] = 1.0 / n;
    S->pbc[a] = (pbc && pbc[a]) ? 1 : 0;
  }
  return true;
}

}  // namespace

extern "C" {

size_t so3k_neighbor_list_workspace_bytes(int64_t N, int64_t capacity) {
  if (N <= 0 || capacity < 0) return 0;
  return carve_nl(N, capacity, nullptr).bytes + 256;
}

int so3k_neighbor_list(int32_t mode, int64_t N, const double* positions, const int32_t* z, const double* cell,
                       const uint8_t* pbc, double cutoff, int64_t capacity, int32_t* idx_i, int32_t* idx_j,
                       int32_t* shifts, int64_t* count, void* workspace, size_t workspace_bytes, int32_t* dev_status,
                       void* stream) {
  if (N <= 0 || N > 0x3fffffff || !positions || !(cutoff > 0.0) || capacity < 0 || capacity > 0x7fffffff) return SO3K_ERR_ARG;
  if (!idx_i || !idx_j || !shifts || !workspace) return SO3K_ERR_ARG;
  if (mode != SO3K_NL_ASE && mode != SO3K_NL_DENSE && mode != SO3K_NL_MIC) return SO3K_ERR_ARG;
  if (workspace_bytes < so3k_neighbor_list_workspace_bytes(N, capacity)) return SO3K_ERR_WORKSPACE;
  NLSetup S;
  unsigned char nopbc[3] = {0, 0, 0};
  if (!complete_and_invert(cell, mode == SO3K_NL_DENSE ? nopbc : pbc, &S)) return SO3K_ERR_ARG;
  if (mode == SO3K_NL_DENSE) S.pbc[0] = S.pbc[1] = S.pbc[2] = 0;
  S.cutoff = cutoff;
  S.mode = mode;
  NLWs w = carve_nl(N, capacity, workspace);
  S.max_bins = w.max_bins;
  cudaStream_t st = (cudaStream_t)stream;
  const int n = (int)N;
  SO3K_LAUNCH(k_bbox, dim3(1), dim3(T), 0, st, S, n, positions, w.bbox);
  SO3K_LAUNCH(k_setup, dim3(1), dim3(32), 0, st, S, w.bbox, w.grid);
  cudaMemsetAsync(w.bin_cnt, 0, sizeof(int) * (2 * (size_t)w.max_bins + 2), st);
  SO3K_LAUNCH(k_bin, dim3(so3k_cdiv(n, T)), dim3(T), 0, st, n, positions, w.grid, w.bin_of, w.wrap, w.bin_cnt);
  exclusive_scan(w.max_bins, w.bin_cnt, w.bin_start, w.bsum, st);
  SO3K_LAUNCH(k_bin_fill, dim3(so3k_cdiv(n, T)), dim3(T), 0, st, n, w.bin_of, w.bin_start, w.cursor, w.tmp_atoms);
  SO3K_LAUNCH(k_bin_sort, dim3(so3k_cdiv(n, T)), dim3(T), 0, st, n, w.bin_of, w.bin_start, w.tmp_atoms, positions, w.wrap,
              w.atoms, w.spos, w.swrap);
  SO3K_LAUNCH(k_pairs<0>, dim3(so3k_cdiv(n, WPB)), dim3(WPB * 32), 0, st, n, (int)mode, cutoff, w.grid, positions, z,
              w.bin_of, w.wrap, w.bin_start, w.atoms, w.spos, w.swrap, w.deg, (const int*)nullptr, (int*)nullptr,
              (int*)nullptr, (long long)capacity);
  exclusive_scan(n, w.deg, w.rowptr, w.bsum, st);
  SO3K_LAUNCH(k_pairs<1>, dim3(so3k_cdiv(n, WPB)), dim3(WPB * 32), 0, st, n, (int)mode, cutoff, w.grid, positions, z,
              w.bin_of, w.wrap, w.bin_start, w.atoms, w.spos, w.swrap, w.deg, (const int*)w.rowptr, w.tj, w.ts,
              (long long)capacity);
  SO3K_LAUNCH(k_rowsort, dim3(capacity > 0 ? so3k_cdiv(capacity, T) : 1), dim3(T), 0, st, n, w.rowptr, w.tj, w.ts, (long long)capacity, idx_i, idx_j,
              shifts, (long long*)count, dev_status);
  return cudaGetLastError() == cudaSuccess ? SO3K_OK : SO3K_ERR_CUDA;
}

}  // extern "C"
