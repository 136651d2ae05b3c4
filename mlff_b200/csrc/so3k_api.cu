// C ABI of libso3k (include/so3k.h): handle, parameter blob, workspace carving and the orchestration of one
// energy+force evaluation:  CSR build -> embed -> L x [q/k/v projections, edge filter+attention, aggregation,
// interaction] -> read-out -> reverse sweep (analytic data-gradient) -> forces.
//   StackNet.__call__            mlff/nn/stacknet/stacknet.py:40-87
//   So3kratesLayer.__call__      mlff/nn/layer/so3krates_layer.py:57-178
//   get_obs_and_force_fn         mlff/nn/stacknet/observable_function.py:127-149
//   MLFFPotential.potential_fn   mlff/mdx/potential/mlff_potential.py:95-109
#include "so3k_internal.h"
#include "so3k_edge.h"
#include <atomic>
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <new>

static std::atomic<uint64_t> g_launches{0};
void so3k_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

namespace {

// edge_fwd: filter kernels (tensor mode) / k_edge_fwd ; aggregate: k_alpha_aggregate (tensor mode) / k_aggregate ;
// edge_bwd: k_edge_bwd2_tc (tensor mode) / k_edge_bwd ; pair_records and qk_stream exist in tensor mode only
enum { CAT_GRAPH = 0, CAT_EMBED, CAT_NODE_PROJ, CAT_PAIR_RECORDS, CAT_EDGE_FWD, CAT_AGGREGATE, CAT_INTERACTION, CAT_READOUT,
       CAT_INTERACTION_BWD, CAT_ALPHA_BWD, CAT_FILTER_RECOMPUTE, CAT_QK_STREAM, CAT_EDGE_BWD, CAT_EDGE_FINISH, CAT_CHI_BWD,
       CAT_NODE_PROJ_BWD, CAT_FORCES, CAT_COUNT };
const char* const kCatNames =
    "graph,embed,node_proj,pair_records,edge_fwd,aggregate,interaction,readout,interaction_bwd,alpha_bwd,"
    "filter_recompute,qk_stream,edge_bwd,edge_finish,chi_bwd,node_proj_bwd,forces";

// Optional per-kernel-category timing with CUDA events on the caller's stream (bench.py's roofline numbers).
struct Profiler {
  bool on = false;
#ifndef SO3K_EMU
  std::vector<cudaEvent_t> ev;       // pairs (start, stop)
  std::vector<int> cat;
  size_t used = 0;
#endif
  double ms[CAT_COUNT] = {0};
  long long cnt[CAT_COUNT] = {0};
};

struct HandleImpl : so3k_handle {
  float* repack = nullptr;
  std::vector<EdgeBlockW> wf, wg;    // per layer
  unsigned char* tc_images = nullptr;
  std::vector<TcBlockW> tf, tg;      // per layer (tensor-core mode)
  bool use_tc = false;
  bool tc_b2 = true;                 // SO3K_TC_B2=0 (env): debugging aid, radial / l0 cotangents on the fp32 CUDA-core kernel
  bool keep_w = false;               // SO3K_FLAG_KEEP_FILTERS
  Profiler prof;
};

struct Scope {
  Profiler& p;
  cudaStream_t st;
  size_t slot = 0;
  bool active = false;
  Scope(Profiler& p_, int c, cudaStream_t st_) : p(p_), st(st_) {
#ifndef SO3K_EMU
    if (p.on) {
      if (p.used + 2 > p.ev.size()) {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        p.ev.push_back(a);
        p.ev.push_back(b);
        p.cat.push_back(c);
      }
      slot = p.used;
      p.cat[slot / 2] = c;
      p.used += 2;
      cudaEventRecord(p.ev[slot], st);
      active = true;
    }
#else
    (void)c;
#endif
  }
  ~Scope() {
#ifndef SO3K_EMU
    if (active) cudaEventRecord(p.ev[slot + 1], st);
#endif
  }
};
#define PROF(cat) Scope _scope_##cat(h->prof, cat, st)

bool make_dims(const so3k_config& c, So3kDims* D) {
  if (c.F <= 0 || c.n_rbf <= 0 || c.n_layers <= 0 || c.n_heads <= 0 || c.n_degrees <= 0) return false;
  if (c.n_degrees > SO3K_MAXDEG || c.n_heads > 8) return false;
  if (c.F % c.n_heads != 0 || c.F % c.n_degrees != 0) return false;
  if (c.F > 256 || c.F / 4 < 1) return false;
  if (!(c.r_cut > 0.f)) return false;
  memset(D, 0, sizeof(*D));
  D->F = c.F; D->K = c.n_rbf; D->L = c.n_layers; D->H = c.n_heads; D->nl = c.n_degrees;
  D->Fs = c.F / 4;                       // int(F/4), mlff/nn/representation/so3krates.py:64-66
  D->dh = c.F / c.n_heads;
  D->dg = c.F / c.n_degrees;
  D->nemb = c.num_embeddings > 0 ? c.num_embeddings : 100;
  D->r_cut = c.r_cut;
  int M = 0;
  for (int s = 0; s < c.n_degrees; ++s) {
    int l = c.degrees[s];
    if (l < 0 || l > 4) return false;
    D->deg[s] = l;
    D->seg_off[s] = M;
    for (int m = 0; m < 2 * l + 1; ++m) {
      if (M >= SO3K_MAXM) return false;
      D->m_seg[M++] = s;
    }
  }
  D->seg_off[c.n_degrees] = M;
  D->M = M;
  // l0-contraction coefficients in the reference's cg_rep order (sorted unique degrees, tiled by multiplicity):
  // mlff/sph_ops/contract.py:168-172
  int idx = 0;
  for (int l = 0; l <= 4; ++l) {
    int count = 0;
    for (int s = 0; s < c.n_degrees; ++s) count += (c.degrees[s] == l);
    for (int r = 0; r < count; ++r)
      for (int m = 0; m < 2 * l + 1; ++m) D->m_cg[idx++] = (float)(1.0 / std::sqrt(2.0 * l + 1.0));
  }
  // PhysNet basis constants in float32, mlff/basis_function/radial.py:163-165
  float rc = c.r_cut;
  float e_rc = expf(-rc);
  D->mu0 = e_rc;
  D->dmu = c.n_rbf > 1 ? (1.0f - e_rc) / (float)(c.n_rbf - 1) : 0.f;
  float t = (2.0f / (float)c.n_rbf) * (1.0f - e_rc);
  D->gamma = 1.0f / (t * t);
  if (D->H + D->nl > 16) return false;
  return true;
}

void make_layout(const So3kDims& D, ModelLayout* ml, bool zbl, int node_opts) {
  size_t off = 0;
  auto add = [&](const std::string& name, size_t count) {
    size_t o = off;
    ml->names.push_back({name, {o, count}});
    off += count;
    return o;
  };
  const size_t F = D.F, K = D.K, nl = D.nl, Fs = D.Fs, H = D.H, dh = D.dh, dg = D.dg;
  ml->embed = add("embed/embedding", (size_t)D.nemb * F);
  ml->layers.resize(D.L);
  for (int t = 0; t < D.L; ++t) {
    std::string lt = "layers_" + std::to_string(t) + "/";
    for (int b = 0; b < 2; ++b) {
      BlockParams& bp = b ? ml->layers[t].gb : ml->layers[t].fb;
      std::string pre = lt + (b ? "gb/" : "fb/");
      bp.rad_W1 = add(pre + "rad/layers_0/kernel", K * F);
      bp.rad_b1 = add(pre + "rad/layers_0/bias", F);
      bp.rad_W2 = add(pre + "rad/layers_1/kernel", F * F);
      bp.rad_b2 = add(pre + "rad/layers_1/bias", F);
      bp.sph_W1 = add(pre + "sph/layers_0/kernel", nl * Fs);
      bp.sph_b1 = add(pre + "sph/layers_0/bias", Fs);
      bp.sph_W2 = add(pre + "sph/layers_1/kernel", Fs * F);
      bp.sph_b2 = add(pre + "sph/layers_1/bias", F);
      size_t heads = b ? nl : H, d = b ? dg : dh;
      bp.Wq = add(pre + "q/kernel", heads * d * d);
      bp.Wk = add(pre + "k/kernel", heads * d * d);
      bp.Wv = b ? 0 : add(pre + "v/kernel", heads * d * d);
    }
    ml->layers[t].int_W = add(lt + "int/layers_0/kernel", (F + nl) * (F + nl));
    ml->layers[t].int_b = add(lt + "int/layers_0/bias", F + nl);
    LayerParams& lp = ml->layers[t];
    const int n_ln = (node_opts & SO3K_OPT_LN) ? ((node_opts & SO3K_OPT_LN_POST) ? 3 : 2) : 0;
    for (int k = 0; k < n_ln; ++k) {
      lp.ln_s[k] = add(lt + "ln_" + std::to_string(k + 1) + "/scale", F);
      lp.ln_b[k] = add(lt + "ln_" + std::to_string(k + 1) + "/bias", F);
    }
    for (int r = 0; r < 2; ++r) {
      if (!(node_opts & (r ? SO3K_OPT_RES2 : SO3K_OPT_RES1))) continue;
      std::string pre = lt + "res_" + std::to_string(r + 1) + "/";
      lp.res[r][0] = add(pre + "layers_0/kernel", F * F);
      lp.res[r][1] = add(pre + "layers_1/kernel", F * F);
      lp.res[r][2] = add(pre + "out/kernel", F * F);
    }
  }
  ml->node_opts = node_opts;
  ml->out_W1 = add("energy/layers_0/kernel", F * F);
  ml->out_b1 = add("energy/layers_0/bias", F);
  ml->out_W2 = add("energy/layers_1/kernel", F);
  ml->out_b2 = add("energy/layers_1/bias", 1);
  ml->scale = add("energy/per_atom_scale", D.nemb);
  ml->shift = add("energy/per_atom_shift", D.nemb);
  ml->has_zbl = zbl;
  if (zbl) {
    static const char* const zn[10] = {"a1", "a2", "a3", "a4", "c1", "c2", "c3", "c4", "p", "d"};
    for (int k = 0; k < 10; ++k) {
      size_t o = add(std::string("zbl/") + zn[k], 1);
      if (k == 0) ml->zbl = o;
    }
  }
  ml->total = off;
}

inline size_t a256(size_t n) { return (n + 255) & ~(size_t)255; }

// ---- workspace carving ------------------------------------------------------------------------------
struct Workspace {
  Graph g;
  int* scratch;
  float *x, *chi;          // (L+1) states each
  float *x1, *chi1, *bcoef;  // x1: 1 buffer ; chi1, bcoef: per layer
  float *proj, *gproj;
  float *alpha;            // per layer
  float *alphabar, *gbar, *rbar, *ebar;
  float *wst;              // tensor mode: filters [layer or 1][block][pair][F] (all layers with SO3K_FLAG_KEEP_FILTERS)
  int wst_layers;
  float *rec;              // tensor mode: per-pair records of the current layer [pair][8]
  float *wbar;             // tensor mode: pair-indexed filter cotangents of the layer being differentiated [block][pair][F]
  float *wtmp;             // tensor mode, wide model (2 x 2 sub-blocks): one sub-block's filter output [pair][F / 2]
  float *xbar_a, *xbar_b, *chibar_a, *chibar_b;
  float *e_atom;
  // optional node-level branches (SO3K_OPT_*; all null in the default model)
  float *xn;               // LayerNorm output feeding the projections / the interaction block ; backward scratch
  float *x1a;              // per layer: x + x_local, the input of ResidualMLP 1
  float *x1b;              // per layer: first skip state (input of the second LayerNorm)
  float *x2a;              // per layer: second skip state before ResidualMLP 2
  float *x2b;              // per layer: layer output before the post LayerNorm
  // rows whose node-level outputs are needed: all N, or (domain decomposition) the owned rows, which come first -- a
  // ghost's layer output, read-out and cotangents belong to its owner and are overwritten by the halo exchange
  int n_node;
  size_t bytes;
};

// low bits: 0 (fp32 CUDA-core mode), 1 (tensor mode, filters recomputed in the backward) or L (kept) ; the SO3K_OPT_*
// bits of the optional node-level branches ride above them
constexpr int WS_OPT_SHIFT = 16;
static inline int wst_layers_of(const HandleImpl* h) {
  return (h->use_tc ? (h->keep_w ? h->dims.L : 1) : 0) | (h->layout.node_opts << WS_OPT_SHIFT);
}

Workspace carve_ws(const So3kDims& D, int N, int Pt, void* base, int ws_mode) {
  const int wst_layers = ws_mode & ((1 << WS_OPT_SHIFT) - 1), opts = ws_mode >> WS_OPT_SHIFT;
  Workspace w;
  char* p = (char*)base;
  auto take = [&](size_t bytes) {
    char* r = p;
    p += a256(bytes);
    return (void*)r;
  };
  const size_t F = D.F, M = D.M, nl = D.nl, L = D.L;
  const size_t Ast = (D.H + D.nl + 3) & ~3, Gst = (D.nl + 3) & ~3;
  const size_t Pe = Pt > 0 ? Pt : 1;
  w.g.N = N; w.g.Pt = Pt;
  w.n_node = N;
  w.g.rowptr = (int*)take(sizeof(int) * (N + 1));
  w.g.col = (int*)take(sizeof(int) * Pe);
  w.g.row = (int*)take(sizeof(int) * Pe);
  w.g.src = (int*)take(sizeof(int) * Pe);
  w.g.rev = (int*)take(sizeof(int) * Pe);
  w.g.canon = (int*)take(sizeof(int) * (Pe + 4));
  w.g.n_canon = w.g.canon + Pe;
  w.g.geom = (float4*)take(sizeof(float4) * Pe);
  w.g.n_valid = w.g.rowptr + N;
  w.scratch = (int*)take(sizeof(int) * so3k_graph_ws_ints(N, Pt));
  w.x = (float*)take(sizeof(float) * (L + 1) * N * F);
  w.chi = (float*)take(sizeof(float) * (L + 1) * N * M);
  w.x1 = (float*)take(sizeof(float) * N * F);
  w.chi1 = (float*)take(sizeof(float) * L * N * M);
  w.bcoef = (float*)take(sizeof(float) * L * N * nl);
  w.proj = (float*)take(sizeof(float) * L * N * 5 * F);     // per layer: kept for the backward (2 640 B per atom at F=132)
  w.gproj = (float*)take(sizeof(float) * N * 5 * F);
  w.alpha = (float*)take(sizeof(float) * L * Pe * Ast);
  w.alphabar = (float*)take(sizeof(float) * Pe * Ast);
  w.gbar = (float*)take(sizeof(float) * Pe * Gst);
  w.rbar = (float*)take(sizeof(float) * Pe * 3);
  w.ebar = (float*)take(sizeof(float) * Pe * 2);
  w.g.pid = (int*)take(sizeof(int) * Pe);
  w.wst_layers = wst_layers;
  w.wst = wst_layers ? (float*)take(sizeof(float) * wst_layers * 2 * (size_t)so3k_pair_capacity(Pt) * F) : nullptr;
  w.rec = wst_layers ? (float*)take(sizeof(float) * 8 * (size_t)so3k_pair_capacity(Pt)) : nullptr;
  w.wbar = wst_layers ? (float*)take(sizeof(float) * 2 * (size_t)so3k_pair_capacity(Pt) * F) : nullptr;
  w.wtmp = (wst_layers && so3k_edge_tc_split(D) == 2) ? (float*)take(sizeof(float) * (size_t)so3k_pair_capacity(Pt) * (F / 2)) : nullptr;
  w.xbar_a = (float*)take(sizeof(float) * N * F);
  w.xbar_b = (float*)take(sizeof(float) * N * F);
  w.chibar_a = (float*)take(sizeof(float) * N * M);
  w.chibar_b = (float*)take(sizeof(float) * N * M);
  w.e_atom = (float*)take(sizeof(float) * N);
  const bool ln = opts & SO3K_OPT_LN, post = ln && (opts & SO3K_OPT_LN_POST);
  w.xn = ln ? (float*)take(sizeof(float) * N * F) : nullptr;
  w.x1a = (opts & SO3K_OPT_RES1) ? (float*)take(sizeof(float) * L * N * F) : nullptr;
  w.x1b = ln ? (float*)take(sizeof(float) * L * N * F) : nullptr;
  w.x2a = (opts & SO3K_OPT_RES2) ? (float*)take(sizeof(float) * L * N * F) : nullptr;
  w.x2b = post ? (float*)take(sizeof(float) * L * N * F) : nullptr;
  w.bytes = (size_t)(p - (char*)base);
  return w;
}

// ---- stages of one evaluation (also exposed one by one through the so3k_dd_* entry points, so that a
// domain-decomposed caller can exchange ghost rows between them) ------------------------------------------------
void stage_begin(HandleImpl* h, Workspace& w, int B, int n, int P, const float* R, const int* z, const int* idx_i,
                 const int* idx_j, const float* cell, const int* cell_offset, const float* disp,
                 const unsigned char* mask, int* dev_status, cudaStream_t st) {
  const So3kDims& D = h->dims;
  { PROF(CAT_GRAPH); so3k_build_graph(D, B, n, P, R, idx_i, idx_j, cell, cell_offset, disp, mask, w.g, w.scratch, dev_status, st); }
  { PROF(CAT_EMBED); so3k_launch_embed(D, B * n, z, h->params + h->layout.embed, w.x, w.chi, dev_status, st); }
}

// x_{t+1}, chi_{t+1} from x_t, chi_t
void stage_layer_fwd(HandleImpl* h, Workspace& w, int t, int* dev_status, cudaStream_t st) {
  const So3kDims& D = h->dims;
  const size_t N = w.g.N, F = D.F, M = D.M, nl = D.nl, Ast = (D.H + D.nl + 3) & ~3, Pe = w.g.Pt > 0 ? w.g.Pt : 1;
  const float* prm = h->params;
  const LayerParams& lp = h->layout.layers[t];
  float* xt = w.x + (size_t)t * N * F;
  float* ct = w.chi + (size_t)t * N * M;
  float* alpha_t = w.alpha + (size_t)t * Pe * Ast;
  float* proj_t = w.proj + (size_t)t * N * 5 * F;
  // optional branches: the states between the blocks (each aliases its successor when the branch is absent)
  const int opts = h->layout.node_opts;
  const bool ln = opts & SO3K_OPT_LN, post = ln && (opts & SO3K_OPT_LN_POST);
  const bool res1 = opts & SO3K_OPT_RES1, res2 = opts & SO3K_OPT_RES2;
  float* x1b = ln ? w.x1b + (size_t)t * N * F : w.x1;              // first skip state
  float* x1a = res1 ? w.x1a + (size_t)t * N * F : x1b;            // x + x_local
  float* x2b = post ? w.x2b + (size_t)t * N * F : xt + N * F;     // layer output before the post LayerNorm
  float* x2a = res2 ? w.x2a + (size_t)t * N * F : x2b;            // second skip state before ResidualMLP 2
  const float* xin = xt;
  if (ln) {
    PROF(CAT_NODE_PROJ);
    so3k_launch_layernorm((int)N, (int)F, xt, prm + lp.ln_s[0], prm + lp.ln_b[0], w.xn, st);
    xin = w.xn;
  }
  { PROF(CAT_NODE_PROJ); so3k_launch_node_proj(D, (int)N, xin, prm, lp, proj_t, st); }
  if (h->use_tc) {
    // filters of both blocks per undirected pair on the tensor cores, then one streaming pass: attention coefficients
    // + both aggregations
    const size_t Pc = (size_t)so3k_pair_capacity(w.g.Pt);
    float* wst_t = w.wst + (size_t)(w.wst_layers > 1 ? t : 0) * 2 * Pc * F;
    { PROF(CAT_PAIR_RECORDS); so3k_launch_pair_records(D, w.g, ct, w.rec, st); }
    {
      PROF(CAT_EDGE_FWD);
      so3k_launch_filter_tc(D, w.g, h->tf[t], w.rec, wst_t, w.wtmp, dev_status, st);
      so3k_launch_filter_tc(D, w.g, h->tg[t], w.rec, wst_t + Pc * F, w.wtmp, dev_status, st);
    }
    PROF(CAT_AGGREGATE);
    so3k_launch_alpha_aggregate(D, w.g, xt, ct, proj_t, wst_t, wst_t + Pc * F, alpha_t, x1a, w.chi1 + (size_t)t * N * M, st);
  } else {
    { PROF(CAT_EDGE_FWD); so3k_launch_edge_fwd(D, w.g, h->wf[t], h->wg[t], ct, proj_t, alpha_t, st); }
    PROF(CAT_AGGREGATE);
    so3k_launch_aggregate(D, w.g, xt, ct, proj_t, alpha_t, x1a, w.chi1 + (size_t)t * N * M, st);
  }
  PROF(CAT_INTERACTION);
  const int Nn = w.n_node;
  if (res1) so3k_launch_resmlp(D, Nn, x1a, prm, lp.res[0], x1b, st);
  if (ln) so3k_launch_layernorm(Nn, (int)F, x1b, prm + lp.ln_s[1], prm + lp.ln_b[1], w.xn, st);
  so3k_launch_interaction(D, Nn, ln ? w.xn : x1b, w.chi1 + (size_t)t * N * M, prm, lp, x2a, ct + N * M,
                          w.bcoef + (size_t)t * N * nl, st, ln ? x1b : nullptr);
  if (res2) so3k_launch_resmlp(D, Nn, x2a, prm, lp.res[1], x2b, st);
  if (post) so3k_launch_layernorm(Nn, (int)F, x2b, prm + lp.ln_s[2], prm + lp.ln_b[2], xt + N * F, st);
}

// e_atom and d(sum e)/dx_L -> xbar_a ; chibar_a = 0 ; rbar = 0
// atom_shift: ZBL shift subtracted from every atom ('per_atom' convention of the MD seam), 0 otherwise
void stage_readout(HandleImpl* h, Workspace& w, const int* z, float out_scale, float atom_shift, float* e_atom,
                   cudaStream_t st) {
  const So3kDims& D = h->dims;
  const size_t N = w.g.N, Pe = w.g.Pt > 0 ? w.g.Pt : 1;
  {
    PROF(CAT_READOUT);
    so3k_launch_readout(D, w.n_node, w.x + (size_t)D.L * N * D.F, z, h->params, h->params_T, h->layout, out_scale, e_atom,
                        w.xbar_a, st);
  }
  cudaMemsetAsync(w.chibar_a, 0, sizeof(float) * N * D.M, st);
  cudaMemsetAsync(w.rbar, 0, sizeof(float) * Pe * 3, st);
  if (h->layout.has_zbl) {
    PROF(CAT_READOUT);
    so3k_launch_zbl(D, w.g, z, h->params + h->layout.zbl, out_scale, atom_shift, e_atom, w.rbar, st);
  }
}

// interaction block of layer t: (xbar_a, chibar_a) = cotangent of (x_{t+1}, chi_{t+1})  ->  (xbar_b, chibar_b) = (xbar1, chibar1)
void stage_layer_bwd_a(HandleImpl* h, Workspace& w, int t, cudaStream_t st) {
  const So3kDims& D = h->dims;
  const size_t N = w.g.N, F = D.F;
  const LayerParams& lp = h->layout.layers[t];
  const int opts = h->layout.node_opts;
  const bool ln = opts & SO3K_OPT_LN, post = ln && (opts & SO3K_OPT_LN_POST);
  const bool res1 = opts & SO3K_OPT_RES1, res2 = opts & SO3K_OPT_RES2;
  PROF(CAT_INTERACTION_BWD);
  const int Nn = w.n_node;
  // the optional branches above the interaction block, last to first (all in place on xbar_a)
  if (post) {
    const float* x2b = w.x2b + (size_t)t * N * F;
    so3k_launch_layernorm_bwd(Nn, (int)F, x2b, h->params + lp.ln_s[2], w.xbar_a, nullptr, w.xbar_a, st);
  }
  if (res2)
    so3k_launch_resmlp_bwd(D, Nn, w.x2a + (size_t)t * N * F, h->params, h->params_T, lp.res[1], w.xbar_a, w.xbar_a, st);
  so3k_launch_interaction_bwd(D, Nn, w.chi1 + (size_t)t * N * D.M, w.bcoef + (size_t)t * N * D.nl, h->params_T,
                              lp, w.xbar_a, w.chibar_a, w.xbar_b, w.chibar_b, st, !ln);
  // ... and those between the first skip connection and the interaction block: xbar_b becomes the cotangent of x + x_local
  if (ln)       // xbar_b holds the cotangent of the normalised input: back through LayerNorm 2, plus the skip connection
    so3k_launch_layernorm_bwd(Nn, (int)F, w.x1b + (size_t)t * N * F, h->params + lp.ln_s[1], w.xbar_b, w.xbar_a,
                              w.xbar_b, st);
  if (res1)
    so3k_launch_resmlp_bwd(D, Nn, w.x1a + (size_t)t * N * F, h->params, h->params_T, lp.res[0], w.xbar_b, w.xbar_b, st);
}

// everything below the interaction block of layer t: edge cotangents, rbar += ..., and (t > 0) the cotangent of
// (x_t, chi_t) back into (xbar_a, chibar_a)
void stage_layer_bwd_b(HandleImpl* h, Workspace& w, int t, int* dev_status, cudaStream_t st) {
  const So3kDims& D = h->dims;
  const size_t N = w.g.N, F = D.F, M = D.M, Ast = (D.H + D.nl + 3) & ~3, Pe = w.g.Pt > 0 ? w.g.Pt : 1;
  const LayerParams& lp = h->layout.layers[t];
  const float* ct = w.chi + (size_t)t * N * M;
  const float* alpha_t = w.alpha + (size_t)t * Pe * Ast;
  const float* proj_t = w.proj + (size_t)t * N * 5 * F;     // kept from the forward
  cudaMemsetAsync(w.gproj, 0, sizeof(float) * N * 5 * F, st);
  { PROF(CAT_ALPHA_BWD); so3k_launch_alpha_bwd(D, w.g, proj_t, alpha_t, w.xbar_b, w.chibar_b, w.alphabar, w.gproj, st); }
  if (h->use_tc) {
    const size_t Pc = (size_t)so3k_pair_capacity(w.g.Pt);
    float* wst_t = w.wst + (size_t)(w.wst_layers > 1 ? t : 0) * 2 * Pc * F;
    { PROF(CAT_PAIR_RECORDS); so3k_launch_pair_records(D, w.g, ct, w.rec, st); }
    if (w.wst_layers == 1) {         // filters not kept: recompute this layer's
      PROF(CAT_FILTER_RECOMPUTE);
      so3k_launch_filter_tc(D, w.g, h->tf[t], w.rec, wst_t, w.wtmp, dev_status, st);
      so3k_launch_filter_tc(D, w.g, h->tg[t], w.rec, wst_t + Pc * F, w.wtmp, dev_status, st);
    }
    const size_t Gst = (D.nl + 3) & ~3;
    cudaMemsetAsync(w.ebar, 0, sizeof(float) * Pe * 2, st);
    cudaMemsetAsync(w.gbar, 0, sizeof(float) * Pe * Gst, st);
    {
      // q/k projection gradients, cutoff cotangent and the pairs' filter cotangents: one streaming pass per block
      PROF(CAT_QK_STREAM);
      so3k_launch_qk_bwd_stream(D, w.g, proj_t, alpha_t, w.alphabar, wst_t, w.gproj, w.ebar, w.wbar, st);
    }
    {
      // radial + l0-contraction cotangents (added onto the canonical edge of each pair)
      PROF(CAT_EDGE_BWD);
      if (h->tc_b2) so3k_launch_edge_bwd2_tc(D, w.g, h->tf[t], h->tg[t], w.rec, w.wbar, w.gbar, w.ebar, dev_status, st);
      else so3k_launch_edge_bwd(D, w.g, h->wf[t], h->wg[t], ct, proj_t, w.alphabar, 2, w.gproj, w.gbar, w.ebar, st);
    }
  } else {
    PROF(CAT_EDGE_BWD);
    so3k_launch_edge_bwd(D, w.g, h->wf[t], h->wg[t], ct, proj_t, w.alphabar, 3, w.gproj, w.gbar, w.ebar, st);
  }
  { PROF(CAT_EDGE_FINISH); so3k_launch_edge_finish(D, w.g, alpha_t, w.chibar_b, w.ebar, w.rbar, st); }
  if (t > 0) {
    // chi0 = 0 and x0 = embedding are constants of the positions: their cotangents are not needed
    { PROF(CAT_CHI_BWD); so3k_launch_chi_bwd(D, w.g, ct, w.gbar, w.chibar_b, w.chibar_a, st); }     // chibar_a <- d/d chi_t
    if (h->layout.node_opts & SO3K_OPT_LN) {
      // the projections read LayerNorm 1 (x_t): their cotangent goes back through it, the skip connection adds xbar_b
      PROF(CAT_NODE_PROJ_BWD);
      cudaMemsetAsync(w.xn, 0, sizeof(float) * N * F, st);
      so3k_launch_node_proj_bwd(D, w.n_node, w.gproj, h->params_T, lp, w.xn, st);
      so3k_launch_layernorm_bwd(w.n_node, (int)F, w.x + (size_t)t * N * F, h->params + lp.ln_s[0], w.xn, w.xbar_b, w.xbar_a, st);
    } else {
      { PROF(CAT_NODE_PROJ_BWD); so3k_launch_node_proj_bwd(D, w.n_node, w.gproj, h->params_T, lp, w.xbar_b, st); }   // xbar_b += proj^T
      cudaMemcpyAsync(w.xbar_a, w.xbar_b, sizeof(float) * N * F, cudaMemcpyDeviceToDevice, st);      // xbar_a <- d/d x_t
    }
  }
}

// One evaluation on an already-flattened problem description.
int run_model(HandleImpl* h, int B, int n, int P, const float* R, const int* z, const int* idx_i, const int* idx_j,
              const float* cell, const int* cell_offset, const float* disp, const unsigned char* mask,
              float out_scale, float* E, float* forces, float* e_atom_out, float* dE_dedges, void* workspace,
              size_t workspace_bytes, int* dev_status, cudaStream_t st) {
  const So3kDims& D = h->dims;
  const int N = B * n, Pt = B * P;
  if (workspace_bytes < so3k_workspace_bytes(h, N, Pt)) return SO3K_ERR_WORKSPACE;
  Workspace w = carve_ws(D, N, Pt, workspace, wst_layers_of(h));
  stage_begin(h, w, B, n, P, R, z, idx_i, idx_j, cell, cell_offset, disp, mask, dev_status, st);
  for (int t = 0; t < D.L; ++t) stage_layer_fwd(h, w, t, dev_status, st);
  float* e_atom = e_atom_out ? e_atom_out : w.e_atom;
  const float zshift = h->layout.has_zbl ? h->cfg.zbl_shift : 0.f;
  const bool md_seam = disp != nullptr;      // 'per_atom' convention: the ZBL shift is subtracted from every atom
  stage_readout(h, w, z, out_scale, md_seam ? zshift : 0.f, e_atom, st);
  if (E) { PROF(CAT_READOUT); so3k_launch_energy_sum(B, n, e_atom, md_seam ? 0.f : zshift * out_scale, E, st); }
  if (!forces && !dE_dedges) return SO3K_OK;
  for (int t = D.L - 1; t >= 0; --t) {
    stage_layer_bwd_a(h, w, t, st);
    stage_layer_bwd_b(h, w, t, dev_status, st);
  }
  { PROF(CAT_FORCES); so3k_launch_forces(D, w.g, Pt, w.rbar, forces, dE_dedges, st); }
  return SO3K_OK;
}

}  // namespace

extern "C" {

const char* so3k_version(void) { return "so3k-b200 0.1.0"; }
uint64_t so3k_launch_count(void) { return g_launches.load(); }

int so3k_create(const so3k_config* cfg, so3k_handle** out) {
  if (!cfg || !out) return SO3K_ERR_ARG;
  So3kDims D;
  if (!make_dims(*cfg, &D)) return SO3K_ERR_CONFIG;
  if (!so3k_edge_supported(D)) return SO3K_ERR_CONFIG;
  if ((cfg->flags & SO3K_FLAG_TENSOR) && !so3k_edge_tc_supported(D)) return SO3K_ERR_CONFIG;
  HandleImpl* h = new (std::nothrow) HandleImpl();
  if (!h) return SO3K_ERR_ARG;
  h->cfg = *cfg;
  h->dims = D;
  h->use_tc = (cfg->flags & SO3K_FLAG_TENSOR) != 0;
  h->keep_w = h->use_tc && (cfg->flags & SO3K_FLAG_KEEP_FILTERS) != 0;
  if (const char* e = getenv("SO3K_TC_B2")) h->tc_b2 = atoi(e) != 0;
  int node_opts = 0;
  if (cfg->flags & SO3K_FLAG_LAYER_NORM) node_opts |= SO3K_OPT_LN | ((cfg->flags & SO3K_FLAG_FINAL_LAYER) ? SO3K_OPT_LN_POST : 0);
  if (cfg->flags & SO3K_FLAG_RESIDUAL_MLP_1) node_opts |= SO3K_OPT_RES1;
  if (cfg->flags & SO3K_FLAG_RESIDUAL_MLP_2) node_opts |= SO3K_OPT_RES2;
  make_layout(D, &h->layout, (cfg->flags & SO3K_FLAG_ZBL) != 0, node_opts);
  *out = h;
  return SO3K_OK;
}

int so3k_destroy(so3k_handle* hh) {
  if (!hh) return SO3K_ERR_ARG;
  HandleImpl* h = static_cast<HandleImpl*>(hh);
  if (h->params) cudaFree(h->params);
  if (h->params_T) cudaFree(h->params_T);
  if (h->repack) cudaFree(h->repack);
  if (h->tc_images) cudaFree(h->tc_images);
  delete h;
  return SO3K_OK;
}

size_t so3k_param_count(const so3k_handle* h) { return h ? h->layout.total : 0; }

int so3k_param_lookup(const so3k_handle* h, const char* name, size_t* offset, size_t* count) {
  if (!h || !name) return SO3K_ERR_ARG;
  for (const auto& kv : h->layout.names) {
    if (kv.first == name) {
      if (offset) *offset = kv.second.first;
      if (count) *count = kv.second.second;
      return SO3K_OK;
    }
  }
  return SO3K_ERR_ARG;
}

int so3k_load_params(so3k_handle* hh, const float* blob_dev, size_t n_floats, void* stream) {
  if (!hh || !blob_dev) return SO3K_ERR_ARG;
  HandleImpl* h = static_cast<HandleImpl*>(hh);
  if (n_floats != h->layout.total) return SO3K_ERR_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const So3kDims& D = h->dims;
  if (!h->params) {
    if (cudaMalloc((void**)&h->params, sizeof(float) * ((h->layout.total + 3) & ~(size_t)3)) != cudaSuccess) return SO3K_ERR_CUDA;
  }
  if (!h->params_T) {
    if (cudaMalloc((void**)&h->params_T, sizeof(float) * ((h->layout.total + 3) & ~(size_t)3)) != cudaSuccess) return SO3K_ERR_CUDA;
  }
  size_t per = so3k_edge_repack_floats(D);
  if (!h->repack) {
    if (cudaMalloc((void**)&h->repack, sizeof(float) * per * 2 * D.L) != cudaSuccess) return SO3K_ERR_CUDA;
  }
  cudaMemcpyAsync(h->params, blob_dev, sizeof(float) * n_floats, cudaMemcpyDeviceToDevice, st);
  cudaMemcpyAsync(h->params_T, blob_dev, sizeof(float) * n_floats, cudaMemcpyDeviceToDevice, st);
  so3k_launch_transpose_params(D, h->layout, h->params, h->params_T, st);
  h->wf.resize(D.L);
  h->wg.resize(D.L);
  for (int t = 0; t < D.L; ++t) {
    so3k_edge_repack(D, h->params, h->layout.layers[t].fb, h->repack + per * (2 * t), &h->wf[t], st);
    so3k_edge_repack(D, h->params, h->layout.layers[t].gb, h->repack + per * (2 * t + 1), &h->wg[t], st);
  }
  if (h->use_tc) {
    size_t ib = so3k_edge_tc_image_bytes(D);
    if (!h->tc_images) {
      if (cudaMalloc((void**)&h->tc_images, ib * 2 * D.L) != cudaSuccess) return SO3K_ERR_CUDA;
    }
    h->tf.resize(D.L);
    h->tg.resize(D.L);
    for (int t = 0; t < D.L; ++t) {
      so3k_edge_tc_build(D, h->params, h->layout.layers[t].fb, h->tc_images + ib * (2 * t), &h->tf[t], h->wf[t], st);
      so3k_edge_tc_build(D, h->params, h->layout.layers[t].gb, h->tc_images + ib * (2 * t + 1), &h->tg[t], h->wg[t], st);
    }
  }
  h->have_params = true;
  return cudaGetLastError() == cudaSuccess ? SO3K_OK : SO3K_ERR_CUDA;
}

size_t so3k_workspace_bytes(const so3k_handle* h, int64_t n_atoms_total, int64_t n_edge_slots_total) {
  if (!h || n_atoms_total <= 0 || n_edge_slots_total < 0) return 0;
  const HandleImpl* hi = static_cast<const HandleImpl*>(h);
  Workspace w = carve_ws(h->dims, (int)n_atoms_total, (int)n_edge_slots_total, nullptr, wst_layers_of(hi));
  return w.bytes + 256;
}

int so3k_energy_forces(so3k_handle* hh, int32_t B, int32_t n, int32_t P, const float* R, const int32_t* z,
                       const int32_t* idx_i, const int32_t* idx_j, const float* cell, const int32_t* cell_offset,
                       float* E, float* F, float* e_atom, void* workspace, size_t workspace_bytes, int32_t* dev_status,
                       void* stream) {
  if (!hh || B <= 0 || n <= 0 || P < 0 || !R || !z || !workspace) return SO3K_ERR_ARG;
  if (P > 0 && (!idx_i || !idx_j)) return SO3K_ERR_ARG;
  if ((cell == nullptr) != (cell_offset == nullptr)) return SO3K_ERR_ARG;
  if ((int64_t)B * n > 0x7fffffff || (int64_t)B * P > 0x7fffffff) return SO3K_ERR_ARG;
  HandleImpl* h = static_cast<HandleImpl*>(hh);
  if (!h->have_params) return SO3K_ERR_NOPARAMS;
  int rc = run_model(h, B, n, P, R, z, idx_i, idx_j, cell, cell_offset, nullptr, nullptr, 1.0f, E, F, e_atom, nullptr,
                     workspace, workspace_bytes, dev_status, (cudaStream_t)stream);
  if (rc != SO3K_OK) return rc;
  return cudaGetLastError() == cudaSuccess ? SO3K_OK : SO3K_ERR_CUDA;
}

int so3k_stress(const so3k_handle* hh, int32_t B, int32_t n, int32_t P, const void* workspace, size_t workspace_bytes,
                float* stress, void* stream) {
  if (!hh || B <= 0 || n <= 0 || P < 0 || !workspace || !stress) return SO3K_ERR_ARG;
  const HandleImpl* h = static_cast<const HandleImpl*>(hh);
  Workspace w = carve_ws(h->dims, B * n, B * P, const_cast<void*>(workspace), wst_layers_of(h));
  if (w.bytes > workspace_bytes) return SO3K_ERR_WORKSPACE;
  so3k_launch_stress(w.g, B, n, w.rbar, stress, (cudaStream_t)stream);
  return cudaGetLastError() == cudaSuccess ? SO3K_OK : SO3K_ERR_CUDA;
}

int so3k_potential_displacements(so3k_handle* hh, int32_t N, int32_t P, const float* edges, const int32_t* z,
                                 const int32_t* centers, const int32_t* others, const uint8_t* mask,
                                 float energy_scale, float* e_atom, float* dE_dedges, float* forces, void* workspace,
                                 size_t workspace_bytes, int32_t* dev_status, void* stream) {
  if (!hh || N <= 0 || P < 0 || !z || !workspace || !e_atom) return SO3K_ERR_ARG;
  if (P > 0 && (!edges || !centers || !others)) return SO3K_ERR_ARG;
  HandleImpl* h = static_cast<HandleImpl*>(hh);
  if (!h->have_params) return SO3K_ERR_NOPARAMS;
  int rc = run_model(h, 1, N, P, nullptr, z, centers, others, nullptr, nullptr, edges, mask, energy_scale, nullptr,
                     forces, e_atom, dE_dedges, workspace, workspace_bytes, dev_status, (cudaStream_t)stream);
  if (rc != SO3K_OK) return rc;
  return cudaGetLastError() == cudaSuccess ? SO3K_OK : SO3K_ERR_CUDA;
}

int so3k_featurise(const so3k_handle* h, int32_t n, int32_t P, const float* R, const int32_t* idx_i,
                   const int32_t* idx_j, const float* cell, const int32_t* cell_offset, const float* displacements,
                   float* rbf, float* phi, float* d, float* unit, float* sph, void* stream) {
  if (!h || n <= 0 || P < 0 || !idx_i || !idx_j) return SO3K_ERR_ARG;
  if (!R && !displacements) return SO3K_ERR_ARG;
  so3k_launch_featurise(h->dims, n, P, R, idx_i, idx_j, cell, cell_offset, displacements, rbf, phi, d, unit, sph,
                        (cudaStream_t)stream);
  return cudaGetLastError() == cudaSuccess ? SO3K_OK : SO3K_ERR_CUDA;
}

// ---- stepwise evaluation for domain decomposition ----------------------------------------------------------
// The local problem of a rank = its owned atoms followed by ghost atoms; edges = all edges received by owned atoms
// plus their reverses ("mirror edges": receiver ghost, sender owned), so the local list is symmetric and every
// reduction the backward needs on an OWNED atom is local.  Between the stages the caller refreshes the GHOST rows of
// the named buffers from their owners (forward: x, chi of layer t+1 ; backward: xbar1, chibar1).
int so3k_dd_begin(so3k_handle* hh, int32_t N, int32_t P, const float* R, const int32_t* z, const int32_t* idx_i,
                  const int32_t* idx_j, const float* cell, const int32_t* cell_offset, void* workspace,
                  size_t workspace_bytes, int32_t* dev_status, void* stream) {
  if (!hh || N <= 0 || P < 0 || !R || !z || !workspace) return SO3K_ERR_ARG;
  HandleImpl* h = static_cast<HandleImpl*>(hh);
  if (!h->have_params) return SO3K_ERR_NOPARAMS;
  if (workspace_bytes < so3k_workspace_bytes(h, N, P)) return SO3K_ERR_WORKSPACE;
  Workspace w = carve_ws(h->dims, N, P, workspace, wst_layers_of(h));
  stage_begin(h, w, 1, N, P, R, z, idx_i, idx_j, cell, cell_offset, nullptr, nullptr, dev_status, (cudaStream_t)stream);
  return cudaGetLastError() == cudaSuccess ? SO3K_OK : SO3K_ERR_CUDA;
}

// stage: 0 = layer forward t ; 1 = read-out (e_atom (N) out, needs z) ; 2 = layer backward A (interaction block) ;
//        3 = layer backward B ; 4 = forces (N,3) out
int so3k_dd_stage_owned(so3k_handle* hh, int32_t stage, int32_t t, int32_t N, int32_t n_owned, int32_t P, const int32_t* z,
                        float* out, void* workspace, int32_t* dev_status, void* stream) {
  if (!hh || !workspace || N <= 0 || n_owned < 0 || n_owned > N) return SO3K_ERR_ARG;
  HandleImpl* h = static_cast<HandleImpl*>(hh);
  if (t < 0 || t >= h->dims.L) return SO3K_ERR_ARG;
  Workspace w = carve_ws(h->dims, N, P, workspace, wst_layers_of(h));
  w.n_node = n_owned;
  cudaStream_t st = (cudaStream_t)stream;
  switch (stage) {
    case 0: stage_layer_fwd(h, w, t, dev_status, st); break;
    case 1: if (!z || !out) return SO3K_ERR_ARG; stage_readout(h, w, z, 1.0f, 0.f, out, st); break;
    case 2: stage_layer_bwd_a(h, w, t, st); break;
    case 3: stage_layer_bwd_b(h, w, t, dev_status, st); break;
    case 4: if (!out) return SO3K_ERR_ARG; so3k_launch_forces(h->dims, w.g, P, w.rbar, out, nullptr, st); break;
    default: return SO3K_ERR_ARG;
  }
  return cudaGetLastError() == cudaSuccess ? SO3K_OK : SO3K_ERR_CUDA;
}
int so3k_dd_stage(so3k_handle* hh, int32_t stage, int32_t t, int32_t N, int32_t P, const int32_t* z, float* out,
                  void* workspace, int32_t* dev_status, void* stream) {
  return so3k_dd_stage_owned(hh, stage, t, N, N, P, z, out, workspace, dev_status, stream);
}

// byte offset inside the workspace of a node buffer: which = 0: x_t (N,F) ; 1: chi_t (N,M) ; 2: xbar1 (N,F) ; 3: chibar1 (N,M)
int so3k_dd_buffer(const so3k_handle* hh, int32_t which, int32_t t, int32_t N, int32_t P, size_t* offset, size_t* row_floats) {
  if (!hh || !offset || !row_floats || N <= 0) return SO3K_ERR_ARG;
  const So3kDims& D = hh->dims;
  if (t < 0 || t > D.L) return SO3K_ERR_ARG;
  const HandleImpl* h = static_cast<const HandleImpl*>(hh);
  Workspace w = carve_ws(D, N, P, nullptr, wst_layers_of(h));
  const char* base = nullptr;
  const char* p = nullptr;
  switch (which) {
    case 0: p = (const char*)(w.x + (size_t)t * N * D.F); *row_floats = D.F; break;
    case 1: p = (const char*)(w.chi + (size_t)t * N * D.M); *row_floats = D.M; break;
    case 2: p = (const char*)w.xbar_b; *row_floats = D.F; break;
    case 3: p = (const char*)w.chibar_b; *row_floats = D.M; break;
    default: return SO3K_ERR_ARG;
  }
  *offset = (size_t)(p - base);
  return SO3K_OK;
}

// the two node buffers of an exchange: pair 0 = (x_t, chi_t), pair 1 = (xbar1, chibar1)
static bool dd_pair(const HandleImpl* h, int pair, int t, int N, int P, void* workspace, float** a, float** b, int* wa, int* wb) {
  const So3kDims& D = h->dims;
  if (t < 0 || t > D.L || (pair != 0 && pair != 1)) return false;
  Workspace w = carve_ws(D, N, P, workspace, wst_layers_of(h));
  *wa = D.F;
  *wb = D.M;
  if (pair == 0) { *a = w.x + (size_t)t * N * D.F; *b = w.chi + (size_t)t * N * D.M; }
  else { *a = w.xbar_b; *b = w.chibar_b; }
  return true;
}
int so3k_dd_pack(const so3k_handle* hh, int32_t pair, int32_t t, int32_t N, int32_t P, const int32_t* send_idx, int32_t n_send,
                 float* out, void* workspace, void* stream) {
  if (!hh || !workspace || N <= 0 || n_send < 0 || (n_send > 0 && (!send_idx || !out))) return SO3K_ERR_ARG;
  float *a, *b;
  int wa, wb;
  if (!dd_pair(static_cast<const HandleImpl*>(hh), pair, t, N, P, workspace, &a, &b, &wa, &wb)) return SO3K_ERR_ARG;
  so3k_launch_dd_pack(n_send, wa, wb, send_idx, a, b, out, (cudaStream_t)stream);
  return cudaGetLastError() == cudaSuccess ? SO3K_OK : SO3K_ERR_CUDA;
}
int so3k_dd_unpack(const so3k_handle* hh, int32_t pair, int32_t t, int32_t N, int32_t n_owned, int32_t P, const float* in,
                   void* workspace, void* stream) {
  if (!hh || !workspace || N <= 0 || n_owned < 0 || n_owned > N || (n_owned < N && !in)) return SO3K_ERR_ARG;
  float *a, *b;
  int wa, wb;
  if (!dd_pair(static_cast<const HandleImpl*>(hh), pair, t, N, P, workspace, &a, &b, &wa, &wb)) return SO3K_ERR_ARG;
  so3k_launch_dd_unpack(N - n_owned, n_owned, wa, wb, in, a, b, (cudaStream_t)stream);
  return cudaGetLastError() == cudaSuccess ? SO3K_OK : SO3K_ERR_CUDA;
}

// ---- instrumentation ----------------------------------------------------------------------------------
const char* so3k_profile_categories(void) { return kCatNames; }

int so3k_profile_enable(so3k_handle* hh, int32_t enable) {
  if (!hh) return SO3K_ERR_ARG;
  static_cast<HandleImpl*>(hh)->prof.on = enable != 0;
  return SO3K_OK;
}

// Waits for the recorded events (this call synchronises), adds their durations to the per-category totals and
// copies the totals out: ms[c] (milliseconds) and count[c] (scopes) for the categories of so3k_profile_categories().
int so3k_profile_collect(so3k_handle* hh, double* ms, int64_t* count, int32_t n_categories, int32_t reset) {
  if (!hh || n_categories < CAT_COUNT) return SO3K_ERR_ARG;
  Profiler& p = static_cast<HandleImpl*>(hh)->prof;
#ifndef SO3K_EMU
  for (size_t k = 0; k + 1 < p.used; k += 2) {
    if (cudaEventSynchronize(p.ev[k + 1]) != cudaSuccess) return SO3K_ERR_CUDA;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, p.ev[k], p.ev[k + 1]) != cudaSuccess) return SO3K_ERR_CUDA;
    p.ms[p.cat[k / 2]] += t;
    p.cnt[p.cat[k / 2]] += 1;
  }
  p.used = 0;
#endif
  for (int c = 0; c < CAT_COUNT; ++c) {
    if (ms) ms[c] = p.ms[c];
    if (count) count[c] = p.cnt[c];
    if (reset) { p.ms[c] = 0; p.cnt[c] = 0; }
  }
  return SO3K_OK;
}

}  // extern "C"
