// Internal structures shared by the translation units of libso3k.
#pragma once
#include "so3k_math.h"
#include "../../include/so3k.h"
#include <string>
#include <vector>

// ---- parameter blob layout (offsets in floats into the caller's blob order, so3k.h) ----------
struct BlockParams {            // one filter block (fb or gb) of one layer
  size_t rad_W1, rad_b1, rad_W2, rad_b2;   // (K,F) (F) (F,F) (F)
  size_t sph_W1, sph_b1, sph_W2, sph_b2;   // (nl,Fs) (Fs) (Fs,F) (F)
  size_t Wq, Wk, Wv;                        // (heads,d,d) each; Wv only for fb
};
struct LayerParams {
  BlockParams fb, gb;
  size_t int_W, int_b;                      // (F+nl,F+nl) (F+nl)
  // optional branches (SO3K_FLAG_LAYER_NORM / _FINAL_LAYER / _RESIDUAL_MLP_1 / _2); 0 when absent
  size_t ln_s[3] = {0, 0, 0}, ln_b[3] = {0, 0, 0};     // LayerNorm scale / bias (F): pre-1, pre-2, post
  size_t res[2][3] = {{0, 0, 0}, {0, 0, 0}};           // ResidualMLP r: Residual layers_0, layers_1, final Dense (F,F)
};
// optional node-level branches of every layer
enum { SO3K_OPT_LN = 1, SO3K_OPT_LN_POST = 2, SO3K_OPT_RES1 = 4, SO3K_OPT_RES2 = 8 };
struct ModelLayout {
  size_t embed;
  std::vector<LayerParams> layers;
  size_t out_W1, out_b1, out_W2, out_b2;    // (F,F) (F) (F,1) (1)
  size_t scale, shift;                      // (nemb) each
  size_t zbl = 0;                           // 10 raw ZBLRepulsion scalars (a1..a4, c1..c4, p, d) with SO3K_FLAG_ZBL
  bool has_zbl = false;
  int node_opts = 0;                        // SO3K_OPT_*
  size_t total;
  std::vector<std::pair<std::string, std::pair<size_t, size_t>>> names;   // name -> (offset, count)
};

struct so3k_handle {
  so3k_config cfg;
  So3kDims dims;
  ModelLayout layout;
  float* params = nullptr;      // device copy of the blob (owned)
  float* params_T = nullptr;    // same layout, with every matrix the backward reads TRANSPOSED (coalesced weight loads)
  bool have_params = false;
};

// ---- graph (CSR) view living in the workspace --------------------------------------------------
struct Graph {
  int N = 0;          // atoms (flattened over the batch)
  int Pt = 0;         // edge slots (flattened, padding included) ; valid edges = rowptr[N] (device)
  int* rowptr;        // (N+1)
  int* col;           // (Pt)  sender j of sorted edge e
  int* row;           // (Pt)  receiver i of sorted edge e
  int* src;           // (Pt)  original slot of sorted edge e
  int* rev;           // (Pt)  sorted index of the reverse edge
  float4* geom;       // (Pt)  (ux, uy, uz, d)
  int* n_valid;       // device scalar = rowptr[N]
  int* canon;         // (Pt)  sorted indices e with e < rev[e]: one representative per undirected pair
  int* n_canon;       // device scalar
  int* pid;           // (Pt)  pair index of sorted edge e: canon[pid[e]] is e or rev[e]
};
// rows of the pair-indexed arrays (a symmetric list has exactly Pt / 2 pairs; malformed lists are clamped to this)
SO3K_HD static inline int so3k_pair_capacity(int Pt) { return Pt / 2 + 1; }


// ---- kernels' host launchers -------------------------------------------------------------------
// graph.cu
size_t so3k_graph_ws_ints(int N, int Pt);
void so3k_build_graph(const So3kDims& D, int B, int n, int P, const float* R, const int* idx_i, const int* idx_j,
                      const float* cell, const int* cell_offset, const float* disp, const unsigned char* mask,
                      Graph& g, int* scratch /* >= so3k_graph_ws_ints */, int* dev_status, cudaStream_t st);
// feat.cu
void so3k_launch_featurise(const So3kDims& D, int n, int P, const float* R, const int* idx_i, const int* idx_j,
                           const float* cell, const int* cell_offset, const float* disp, float* rbf, float* phi,
                           float* d, float* unit, float* sph, cudaStream_t st);
// node.cu
void so3k_launch_embed(const So3kDims& D, int N, const int* z, const float* emb, float* x, float* chi,
                       int* dev_status, cudaStream_t st);
void so3k_launch_node_proj(const So3kDims& D, int N, const float* x, const float* params, const LayerParams& lp,
                           float* proj, cudaStream_t st);
// params_T = blob of so3k_launch_transpose_params (per-head q/k/v matrices, interaction and read-out matrices transposed)
void so3k_launch_transpose_params(const So3kDims& D, const ModelLayout& ml, const float* params, float* params_T,
                                  cudaStream_t st);
void so3k_launch_node_proj_bwd(const So3kDims& D, int N, const float* gproj, const float* params_T,
                               const LayerParams& lp, float* xbar /* += */, cudaStream_t st);
// x2 = xbase + a(x1, chi1) ; xbase = nullptr means x1 (the default layer: input and skip base coincide)
void so3k_launch_interaction(const So3kDims& D, int N, const float* x1, const float* chi1, const float* params,
                             const LayerParams& lp, float* x2, float* chi2, float* bcoef, cudaStream_t st,
                             const float* xbase = nullptr);
// skip = false leaves the identity (skip-connection) term out of xbar1: only the cotangent of the block's x input
void so3k_launch_interaction_bwd(const So3kDims& D, int N, const float* chi1, const float* bcoef,
                                 const float* params_T, const LayerParams& lp, const float* xbar_out,
                                 const float* chibar_out, float* xbar1, float* chibar1, cudaStream_t st,
                                 bool skip = true);
// flax LayerNorm over the F features of every row: y = (x - mean) * rsqrt(var + 1e-6) * scale + bias
void so3k_launch_layernorm(int N, int F, const float* x, const float* scale, const float* bias, float* y, cudaStream_t st);
// out = (add ? add : 0) + dLayerNorm(x)^T gy ; out may alias gy or add
void so3k_launch_layernorm_bwd(int N, int F, const float* x, const float* scale, const float* gy, const float* add,
                               float* out, cudaStream_t st);
// ResidualMLP: y = silu(x + silu(silu(x) W0) W1) W2 ; W = the three (F,F) matrices of LayerParams::res[r]
void so3k_launch_resmlp(const So3kDims& D, int N, const float* x, const float* params, const size_t* W, float* y,
                        cudaStream_t st);
// xbar = dResidualMLP(x)^T ybar ; xbar may alias ybar ; params_T holds the transposed matrices
void so3k_launch_resmlp_bwd(const So3kDims& D, int N, const float* x, const float* params, const float* params_T,
                            const size_t* W, const float* ybar, float* xbar, cudaStream_t st);
void so3k_launch_readout(const So3kDims& D, int N, const float* x, const int* z, const float* params,
                         const float* params_T, const ModelLayout& ml, float out_scale, float* e_atom, float* xbar,
                         cudaStream_t st);
void so3k_launch_energy_sum(int B, int n, const float* e_atom, float offset, float* E, cudaStream_t st);   // E[b] = sum - offset
// Ziegler-Biersack-Littmark repulsion: e_atom[i] += out_scale * (sum_{e in row i} e_edge - atom_shift), rbar[e] += dE/dr_e
void so3k_launch_zbl(const So3kDims& D, const Graph& g, const int* z, const float* zbl_raw, float out_scale,
                     float atom_shift, float* e_atom, float* rbar, cudaStream_t st);
// halo rows of the domain decomposition: gather the rows a peer needs of two node buffers into one message / scatter the
// received rows into the ghost blocks (rows n_own ..)
void so3k_launch_dd_pack(int n_rows, int wa, int wb, const int* send_idx, const float* a, const float* b, float* out,
                         cudaStream_t st);
void so3k_launch_dd_unpack(int n_rows, int n_own, int wa, int wb, const float* in, float* a, float* b, cudaStream_t st);
// agg.cu
void so3k_launch_aggregate(const So3kDims& D, const Graph& g, const float* x, const float* chi, const float* proj,
                           const float* alpha, float* x1, float* chi1, cudaStream_t st);
void so3k_launch_alpha_bwd(const So3kDims& D, const Graph& g, const float* proj, const float* alpha,
                           const float* xbar1, const float* chibar1, float* alphabar, float* gproj /* v grad */,
                           cudaStream_t st);
void so3k_launch_alpha_aggregate(const So3kDims& D, const Graph& g, const float* x, const float* chi, const float* proj,
                                 const float* wst_f, const float* wst_g, float* alpha, float* x1, float* chi1,
                                 cudaStream_t st);
void so3k_launch_qk_bwd_stream(const So3kDims& D, const Graph& g, const float* proj, const float* alpha,
                               const float* alphabar, const float* wst, float* gproj, float* ebar, float* wbar,
                               cudaStream_t st);
void so3k_launch_stress(const Graph& g, int B, int n_per, const float* rbar, float* stress /* (B,3,3) */, cudaStream_t st);
void so3k_launch_chi_bwd(const So3kDims& D, const Graph& g, const float* chi, const float* gbar,
                         const float* chibar1, float* chibar /* out */, cudaStream_t st);
void so3k_launch_forces(const So3kDims& D, const Graph& g, int P_slots, const float* rbar, float* forces /* (N,3) or null */,
                        float* dE_dedges /* (P_slots,3) or null, original order */, cudaStream_t st);
