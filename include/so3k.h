/* so3k.h -- C ABI of the B200-native SO3krates energy+force engine (libso3k.so).
 *
 * The reference (thorben-frank/mlff 0.3.0) is pure Python on JAX/flax: it has NO native
 * operator / plugin / FFI interface.  The boundary below is therefore placed at the two Python
 * seams through which the hot path is reached today; every entry point
 * cites the reference interface it replaces (paths relative to the reference checkout):
 *
 *   so3k_energy_forces          <- obs_fn = jax.vmap(get_obs_and_force_fn(net), (None, 0))
 *                                  mlff/nn/stacknet/observable_function.py:127-149 (+ :41-63 jacrev),
 *                                  StackNet.__call__ mlff/nn/stacknet/stacknet.py:40-87
 *   so3k_potential_displacements<- MLFFPotential.potential_fn  mlff/mdx/potential/mlff_potential.py:95-109
 *                                  + the caller's jax.value_and_grad  mlff/mdx/calculator.py:42-49,
 *                                  mlff/md/calculator.py:101-109
 *   so3k_featurise              <- GeometryEmbed.__call__  mlff/nn/embed/embed.py:66-169
 *                                  (PhysNetBasis radial.py:154-166, cosine_cutoff_fn cutoff radial.py:39-51,
 *                                   fn_Y1..fn_Y4 spherical.py:6-38, add_cell_offsets pbc.py:8-18)
 *   so3k_neighbor_list          <- pbc_neighbor_list / get_pbc_neighbors  mlff/indexing/indices.py:194-275, 111-191
 *                                  (ase.neighborlist.primitive_neighbor_list 'ijS'), and get_indices :15-84
 *   so3k_load_params            <- flax variable dict {'params': ...} (valid_params, mlff/io/checkpoint.py:12-16)
 *
 * Conventions
 *   - plain C, no torch / CUDA types in signatures: device buffers are raw pointers, the CUDA
 *     stream is passed as `void*` (a cudaStream_t; NULL = legacy default stream).
 *   - every function only ENQUEUES work on the caller's stream and never synchronises; the
 *     caller owns all buffers including the workspace; the handle owns only its repacked
 *     copy of the weights.  One handle may be used from several host threads / streams as long
 *     as each concurrent call uses its own workspace.
 *   - return value: SO3K_OK or an error code; nothing throws across the ABI.  Conditions that
 *     are only known on the device (neighbour-list capacity overflow, an edge list that is not
 *     symmetric) are reported through an `int32_t* dev_status` word in device memory, exactly
 *     like PrimitiveNeighbors.overflow (indices.py:264-272): bit flags SO3K_STATUS_*.
 *   - index arrays are int32; padding sentinels follow the reference: idx == -1 (training,
 *     stacknet.py:131) and idx >= n (glp, mdx/calculator.py:133) both mean "no edge";
 *     z == 0 means "padding atom" (stacknet.py:130).
 *   - all model arithmetic is float32 (the reference default); energies of a structure are
 *     accumulated in float64 before the final cast.
 */
#ifndef SO3K_H_
#define SO3K_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SO3K_MAX_DEGREES 8
#define SO3K_MAX_M 32

typedef struct so3k_handle so3k_handle;

/* Hyper-parameters: So3krates(F, n_layer, degrees, n_heads) + GeometryEmbed(n_rbf, r_cut)
 * (mlff/nn/representation/so3krates.py:13-67).  Fixed (default) choices of the reference that
 * this engine implements: PhysNet radial basis, cosine cutoff, sphc=True with
 * sphc_normalization=None (chi0 = 0), radial_spherical filters, filter widths [F,F] / [F/4,F]. */
typedef struct so3k_config {
  int32_t F;               /* feature dimension (132)                        */
  int32_t n_rbf;           /* K, radial basis functions (32)                 */
  int32_t n_layers;        /* So3kratesLayer count (CLI default 3)           */
  int32_t n_heads;         /* feature-block heads H (4); F % H == 0          */
  int32_t n_degrees;       /* len(degrees) = nl; F % nl == 0                 */
  int32_t degrees[SO3K_MAX_DEGREES]; /* e.g. {1,2,3}; repeats allowed, l<=4  */
  float r_cut;             /* 5.0                                            */
  int32_t num_embeddings;  /* 100                                            */
  int32_t flags;           /* SO3K_FLAG_*                                    */
  float zbl_shift;         /* Energy(zbl_repulsion_shift), used with SO3K_FLAG_ZBL; 0 */
} so3k_config;

enum {
  SO3K_FLAG_NONE = 0,
  /* filter-GEMM arithmetic: default = fp32 FMA on the CUDA cores (bit-faithful to fp32);
   * SO3K_FLAG_TENSOR = tcgen05 tensor cores with split-operand (3-pass) error compensation. */
  SO3K_FLAG_TENSOR = 1,
  /* with SO3K_FLAG_TENSOR: the forward keeps every edge's filter vectors (2 * F floats per edge and layer in the
   * workspace -- so3k_workspace_bytes accounts for it) and the backward streams them instead of recomputing the
   * filter MLP for the q/k projection gradients.  Memory for time; results are identical. */
  SO3K_FLAG_KEEP_FILTERS = 2,
  /* Energy(zbl_repulsion=True) (mlff/nn/observable/observable.py:81-84, 110-183; CLI --zbl_repulsion): adds the
   * Ziegler-Biersack-Littmark pair repulsion to the energy and its derivative to the forces.  The parameter blob then
   * ends with the ten RAW ZBLRepulsion scalars zbl/{a1,a2,a3,a4,c1,c2,c3,c4,p,d} (softplus is applied at use, as in the
   * reference).  so3k_config.zbl_shift is subtracted once per structure (so3k_energy_forces, 'per_structure'
   * convention) or from every atom (so3k_potential_displacements, 'per_atom' convention), as the reference does. */
  SO3K_FLAG_ZBL = 4,
  /* Optional So3kratesLayer branches (mlff/nn/layer/so3krates_layer.py:30-42, selected in the reference through
   * So3krates(so3krates_layer_kwargs=...), so they hold for EVERY layer):
   *   SO3K_FLAG_LAYER_NORM      layer_normalization=True: flax LayerNorm (eps 1e-6, scale + bias over the F features)
   *                             before the feature/geometric blocks and before the interaction block (:103-106, :153-156);
   *                             the skip connections keep the un-normalised features (:146, :163)
   *   SO3K_FLAG_FINAL_LAYER     final_layer=True: with LAYER_NORM, a third LayerNorm on the layer's output (:170-174)
   *   SO3K_FLAG_RESIDUAL_MLP_1  residual_mlp_1=True: ResidualMLP (mlff/nn/mlp/mlp.py:30-72; three bias-free F x F
   *   SO3K_FLAG_RESIDUAL_MLP_2  Dense with silu) after the first / second skip connection (:149-150, :166-167)
   * The parameter blob grows per layer as documented at so3k_load_params. */
  SO3K_FLAG_LAYER_NORM = 8,
  SO3K_FLAG_RESIDUAL_MLP_1 = 16,
  SO3K_FLAG_RESIDUAL_MLP_2 = 32,
  SO3K_FLAG_FINAL_LAYER = 64
};

enum {
  SO3K_OK = 0,
  SO3K_ERR_CONFIG = 1,     /* unsupported hyper-parameters                     */
  SO3K_ERR_ARG = 2,        /* NULL / negative / inconsistent argument          */
  SO3K_ERR_WORKSPACE = 3,  /* workspace too small                              */
  SO3K_ERR_CUDA = 4,       /* a CUDA runtime call failed                       */
  SO3K_ERR_NOPARAMS = 5    /* so3k_load_params has not been called             */
};

/* bits of the device status word */
enum {
  SO3K_STATUS_OVERFLOW = 1,    /* neighbour list did not fit `capacity`          */
  SO3K_STATUS_ASYMMETRIC = 2,  /* edge (i,j) without its reverse (j,i)           */
  SO3K_STATUS_BAD_Z = 4,       /* atomic number outside [0, num_embeddings)      */
  SO3K_STATUS_TC_TIMEOUT = 8   /* a tensor-core kernel gave up waiting on an mbarrier (bounded wait instead of a
                                  hang): the outputs of that call are undefined                              */
};

int so3k_create(const so3k_config* cfg, so3k_handle** out);
int so3k_destroy(so3k_handle* h);
const char* so3k_version(void);

/* ---- parameters --------------------------------------------------------------------------
 * One flat float32 blob in this fixed order (flax `Dense` convention y = x @ kernel + bias,
 * kernel stored (in, out) row-major; leading vmap axes kept):
 *   embed/embedding                         (num_embeddings, F)
 *   for t in 0..n_layers-1:
 *     for blk in (fb, gb):
 *       layers_t/blk/rad/layers_0/{kernel (K,F), bias (F)}
 *       layers_t/blk/rad/layers_1/{kernel (F,F), bias (F)}
 *       layers_t/blk/sph/layers_0/{kernel (nl,F/4), bias (F/4)}
 *       layers_t/blk/sph/layers_1/{kernel (F/4,F), bias (F)}
 *       layers_t/blk/q/kernel (heads, d, d) ; layers_t/blk/k/kernel (heads, d, d)
 *       [fb only] layers_t/fb/v/kernel (H, F/H, F/H)        (fb: heads=H ; gb: heads=nl)
 *     layers_t/int/layers_0/{kernel (F+nl,F+nl), bias (F+nl)}
 *     [SO3K_FLAG_LAYER_NORM]      layers_t/ln_1/{scale (F), bias (F)} ; layers_t/ln_2/{scale, bias}
 *     [.. + SO3K_FLAG_FINAL_LAYER] layers_t/ln_3/{scale, bias}
 *     [SO3K_FLAG_RESIDUAL_MLP_1]  layers_t/res_1/{layers_0/kernel, layers_1/kernel, out/kernel}   (F,F) each
 *     [SO3K_FLAG_RESIDUAL_MLP_2]  layers_t/res_2/{layers_0/kernel, layers_1/kernel, out/kernel}
 *   energy/layers_0/{kernel (F,F), bias (F)} ; energy/layers_1/{kernel (F,1), bias (1)}
 *   energy/per_atom_scale (num_embeddings) ; energy/per_atom_shift (num_embeddings)
 * so3k_param_lookup maps such a name to (offset, count) in the blob. */
size_t so3k_param_count(const so3k_handle* h);
int so3k_param_lookup(const so3k_handle* h, const char* name, size_t* offset, size_t* count);
int so3k_load_params(so3k_handle* h, const float* blob_dev, size_t n_floats, void* stream);

/* ---- workspace -------------------------------------------------------------------------- */
/* bytes needed by so3k_energy_forces / so3k_potential_displacements for a flattened problem of
 * n_atoms_total atoms (B*n) and n_edge_slots_total edge slots (B*P, padding included). */
size_t so3k_workspace_bytes(const so3k_handle* h, int64_t n_atoms_total, int64_t n_edge_slots_total);

/* ---- energy + forces, training / evaluation seam ------------------------------------------
 * Batched, padded inputs exactly as `vmap(obs_fn)` receives them:
 *   R (B,n,3) f32, z (B,n) i32, idx_i / idx_j (B,P) i32 (indices local to the structure),
 *   cell (B,3,3) f32 row-wise lattice vectors or NULL, cell_offset (B,P,3) i32 or NULL.
 * Outputs: E (B) f32 [per_structure energy, observable.py:91], F (B,n,3) f32 = -dE/dR,
 *   e_atom (B,n) f32 or NULL [per_atom convention, observable.py:88-89].
 * dev_status (device int32, may be NULL) receives SO3K_STATUS_* bits (it is OR-ed, clear it first). */
int so3k_energy_forces(so3k_handle* h, int32_t B, int32_t n, int32_t P,
                       const float* R, const int32_t* z, const int32_t* idx_i, const int32_t* idx_j,
                       const float* cell, const int32_t* cell_offset,
                       float* E, float* F, float* e_atom,
                       void* workspace, size_t workspace_bytes, int32_t* dev_status, void* stream);

/* ---- stress ---------------------------------------------------------------------------------
 * dE/d(strain) of the evaluation that last used `workspace` (so3k_energy_forces with the same B, n, P):
 *   stress[b] = sym( sum_e dE/dr_e (x) r_e ) over the edges of structure b, r_e = R_j - R_i + S@cell,
 * i.e. what the reference obtains by differentiating the energy through strained positions and cell,
 * get_energy_force_stress_fn  mlff/nn/stacknet/observable_function.py:152-186 (no volume factor there either).
 * stress: (B,3,3) f32 device buffer. */
int so3k_stress(const so3k_handle* h, int32_t B, int32_t n, int32_t P, const void* workspace, size_t workspace_bytes,
                float* stress, void* stream);

/* ---- MD seam ------------------------------------------------------------------------------
 * Graph(edges, nodes, centers, others, mask) of mlff/utils/structures.py:5 in the
 * 'displacements' convention: edges (P,3) f32 = R_j - R_i after MIC, nodes z (N) i32,
 * centers / others (P) i32 with sentinel N (or -1), mask (P) u8 or NULL.
 * Outputs: e_atom (N) f32 = per-atom energy * energy_scale (mlff_potential.py:80-109);
 *   dE_dedges (P,3) f32 or NULL = d(sum e_atom)/d edges (what the caller's value_and_grad chains
 *   through R_j - R_i); forces (N,3) f32 or NULL = -dE/dR obtained by that chain
 *   (mdx/calculator.py:46-49).
 * With SO3K_FLAG_TENSOR the radial and l0-contraction cotangents are evaluated once per undirected pair and booked on
 * ONE of its two directed edges: dE_dedges then equals the reference's per-edge gradient up to that antisymmetric pair
 * gauge (dE_dedges[e] - dE_dedges[rev e] is exact), so everything the MD caller derives from it -- forces by scatter
 * over centers / others, stress sum_e dE_dedges[e] (x) edges[e] -- is unchanged; per-edge equality holds in fp32 mode.
 */
int so3k_potential_displacements(so3k_handle* h, int32_t N, int32_t P,
                                 const float* edges, const int32_t* z,
                                 const int32_t* centers, const int32_t* others, const uint8_t* mask,
                                 float energy_scale,
                                 float* e_atom, float* dE_dedges, float* forces,
                                 void* workspace, size_t workspace_bytes, int32_t* dev_status, void* stream);

/* ---- training seam: cotangents of (E, F) -> parameter gradients -----------------------------------------------
 * What `jax.value_and_grad(loss_fn)(params, batch)` needs from the model (mlff/training/run.py:34-50, loss.py:42-73 with
 * obs_fn = vmap(get_obs_and_force_fn(net)) or get_energy_force_stress_fn, observable_function.py:127-186): for cotangents
 * E_bar (B) of the energies, F_bar (B,n,3) of the forces of so3k_energy_forces (same padded inputs) and S_bar (B,3,3) of the
 * stress of so3k_stress (each may be NULL = zero),
 *     grad[theta] = sum_b E_bar[b] dE_b/dtheta + sum F_bar dF/dtheta + sum S_bar dstress/dtheta ,
 * a float32 blob in the order of so3k_load_params (overwritten; energy/per_atom_scale, _shift get 0: they are constants
 * of the Energy module, observable.py:49-58).  Because F = -dE/dR this is second order; it is evaluated as the
 * forward-mode derivative along -F_bar (positions) and sym(S_bar) (strain: edge vectors move as r -> (1 + eps) r) of the
 * reverse-mode parameter gradient, in one dual-number sweep (so3k_train.cu).
 * F_bar == NULL and S_bar == NULL select the first-order sweep (grad = sum_b E_bar[b] dE_b/dtheta; E_bar == NULL means all ones).
 * E_out (B) or NULL receives the energies of the sweep.  Every model variant of so3k_config.flags is covered: the optional
 * LayerNorms and ResidualMLPs and the ten ZBL scalars have their gradients in the blob too.  Arithmetic: fp32 on the CUDA
 * cores; with SO3K_FLAG_TENSOR the large forward / input-gradient dense layers run on tcgen05 with split-bf16 operands
 * (fp32 accumulation), the weight gradients stay fp32 FMA. */
size_t so3k_train_workspace_bytes(const so3k_handle* h, int64_t n_atoms_total, int64_t n_edge_slots_total);
int so3k_energy_forces_vjp(so3k_handle* h, int32_t B, int32_t n, int32_t P,
                           const float* R, const int32_t* z, const int32_t* idx_i, const int32_t* idx_j,
                           const float* cell, const int32_t* cell_offset,
                           const float* E_bar, const float* F_bar, const float* S_bar, float* grad, float* E_out,
                           void* workspace, size_t workspace_bytes, int32_t* dev_status, void* stream);

/* Loss of the reference's training CLI (scaled_safe_masked_mse_loss, loss.py:13-29, scale 1; get_loss_fn :42-73):
 *   loss = w_energy mse(E, E_true) + w_force mse(F, F_true), each a mean over the entries whose target is not NaN
 *   (forces: and whose atom is real, z != 0 = node_mask).
 * so3k_loss_sums ADDS (sum sq E residual, n_E, sum sq F residual, n_F) of this rank's batch onto sums[4] (device f64;
 * data-parallel training all-reduces these four numbers so that every rank divides by the global denominators);
 * so3k_loss_cotangents turns them into E_bar (B), F_bar (B,n,3) and loss[3] = (loss, mse_E, mse_F) (device f32). */
int so3k_loss_sums(int32_t B, int32_t n, const float* E, const float* F, const float* E_true, const float* F_true,
                   const int32_t* z, double* sums, void* stream);
int so3k_loss_cotangents(int32_t B, int32_t n, const float* E, const float* F, const float* E_true, const float* F_true,
                         const int32_t* z, const double* sums, float w_energy, float w_force,
                         float* E_bar, float* F_bar, float* loss, void* stream);

/* One update of the reference's optimiser chain (mlff/training/optimizer.py:64-97): optax.clip_by_global_norm
 * (clip_norm <= 0: off) -> scale_by_adam(b1, b2, eps, eps_root) -> add_decayed_weights(weight_decay, mask) ->
 * scale(-lr), in place on the float32 blobs params / m / v (n entries).  step is the 1-based update count (bias
 * correction); frozen / decay_mask are per-entry u8 masks or NULL; *grad_sumsq (device f64) receives |grad|^2
 * (train_metrics['gradients_norm'] squared, run.py:49).  step_dev (device int64, may be NULL): when given, the update count
 * lives on the device -- the call increments it and takes the bias corrections from it, `step` is ignored -- so that a
 * captured CUDA graph of the training step can be replayed.  Call so3k_load_params afterwards. */
int so3k_adam_step(size_t n, float* params, const float* grad, float* m, float* v, const uint8_t* frozen,
                   const uint8_t* decay_mask, int64_t step, float lr, float b1, float b2, float eps, float eps_root,
                   float weight_decay, float clip_norm, double* grad_sumsq, int64_t* step_dev, void* stream);

/* ---- stepwise evaluation (spatial domain decomposition) ----------------------------------------------------
 * No counterpart in the reference (it has no multi-device path).  A rank's local problem is
 * its OWNED atoms followed by GHOST atoms, with all edges received by owned atoms plus their reverses ("mirror
 * edges": receiver ghost, sender owned), which keeps the local list symmetric so every reduction needed on an owned
 * atom is local.  The caller runs the stages in order and, between them, overwrites the ghost rows of the named
 * workspace buffers with their owners' rows (NCCL / peer copies):
 *   so3k_dd_begin                          graph + embedding           (x_0 needs no exchange)
 *   for t: so3k_dd_stage(0, t)             layer forward               -> exchange x_{t+1}, chi_{t+1}
 *   so3k_dd_stage(1)                       read-out: out = e_atom (N)  (sum the OWNED entries, all-reduce)
 *   for t desc: so3k_dd_stage(2, t)        interaction backward        -> exchange xbar1, chibar1
 *               so3k_dd_stage(3, t)        edge + node backward
 *   so3k_dd_stage(4)                       out = forces (N,3); only OWNED rows are complete
 * so3k_dd_buffer returns the byte offset (inside the workspace) and row width of x_t / chi_t / xbar1 / chibar1. */
int so3k_dd_begin(so3k_handle* h, int32_t N, int32_t P, const float* R, const int32_t* z, const int32_t* idx_i,
                  const int32_t* idx_j, const float* cell, const int32_t* cell_offset, void* workspace,
                  size_t workspace_bytes, int32_t* dev_status, void* stream);
int so3k_dd_stage(so3k_handle* h, int32_t stage, int32_t t, int32_t N, int32_t P, const int32_t* z, float* out,
                  void* workspace, int32_t* dev_status, void* stream);
/* The same stage with the node-level work (interaction block, read-out and their reverses, projection gradients) limited
 * to the first n_owned rows: a ghost's layer output, read-out and cotangents are its owner's and are overwritten by the
 * exchange, so computing them locally is wasted (37 % of the rows of a rank at 8 bricks of the 100k-atom box).  Rows
 * >= n_owned of e_atom / forces are then undefined.  so3k_dd_stage == so3k_dd_stage_owned with n_owned = N. */
int so3k_dd_stage_owned(so3k_handle* h, int32_t stage, int32_t t, int32_t N, int32_t n_owned, int32_t P, const int32_t* z,
                        float* out, void* workspace, int32_t* dev_status, void* stream);
int so3k_dd_buffer(const so3k_handle* h, int32_t which, int32_t t, int32_t N, int32_t P, size_t* offset,
                   size_t* row_floats);
/* Halo messages of an exchange, one kernel each.  pair 0 = (x_t, chi_t), pair 1 = (xbar1, chibar1); a message row is the
 * F + M floats [x row | chi row] of one atom.
 *   so3k_dd_pack:   out (n_send, F + M) <- the owned rows send_idx[r] (device int32, the sender's local row ids, grouped
 *                   by destination rank) of both buffers
 *   so3k_dd_unpack: the ghost rows n_owned .. N-1 of both buffers <- in (N - n_owned, F + M), in the receiver's ghost order
 * The transport between the two (NCCL all-to-all, peer copies) is the caller's. */
int so3k_dd_pack(const so3k_handle* h, int32_t pair, int32_t t, int32_t N, int32_t P, const int32_t* send_idx, int32_t n_send,
                 float* out, void* workspace, void* stream);
int so3k_dd_unpack(const so3k_handle* h, int32_t pair, int32_t t, int32_t N, int32_t n_owned, int32_t P, const float* in,
                   void* workspace, void* stream);

/* ---- edge featurisation (standalone; the fused model path recomputes these on chip) ---------
 * Inputs as GeometryEmbed: either positions R (n,3) [+ cell (3,3) row-wise + cell_offset (P,3)]
 * or displacements (P,3).  Any output pointer may be NULL.  rbf (P,K), phi (P), d (P),
 * unit (P,3), sph (P,M); padded edges (idx_i == -1 or >= n) give exact zeros. */
int so3k_featurise(const so3k_handle* h, int32_t n, int32_t P,
                   const float* R, const int32_t* idx_i, const int32_t* idx_j,
                   const float* cell, const int32_t* cell_offset, const float* displacements,
                   float* rbf, float* phi, float* d, float* unit, float* sph, void* stream);

/* ---- neighbour list ------------------------------------------------------------------------
 * Cell-list build on the device.  positions (N,3) f64 and z (N) i32 (may be NULL) are DEVICE
 * buffers; cell (3,3) f64 row-wise (ASE convention; may be NULL / zero along non-periodic axes)
 * and pbc[3] are small HOST arrays (they size the bin grid, like scalar arguments); cutoff f64.
 * mode SO3K_NL_ASE  : all images, |R_j - R_i + S@cell| <  cutoff, (i != j or S != 0)   [indices.py:216-224]
 * mode SO3K_NL_DENSE: non-periodic, 0 < d <= cutoff, z_i * z_j != 0 (z may be NULL)      [indices.py:56-60]
 * mode SO3K_NL_MIC  : minimum-image list of the MD path: one image per ordered pair i != j -- the one whose fractional
 *                     displacement lies in [-1/2, 1/2) on the periodic axes --, d < cutoff; `shifts` holds that image
 *                     (glp quadratic_neighbor_list + make_displacement, mlff/mdx/atoms.py:488-529, md/calculator.py:204-238)
 * Output (device): idx_i, idx_j (capacity) i32 padded with -1, shifts (capacity,3) i32 padded
 * with 0, sorted by i then (j, S); *count (device int64) = number of edges found (may exceed capacity, in
 * which case SO3K_STATUS_OVERFLOW is set and only the first `capacity` edges are stored). */
enum { SO3K_NL_ASE = 0, SO3K_NL_DENSE = 1, SO3K_NL_MIC = 2 };
size_t so3k_neighbor_list_workspace_bytes(int64_t N, int64_t capacity);
int so3k_neighbor_list(int32_t mode, int64_t N, const double* positions, const int32_t* z,
                       const double* cell, const uint8_t* pbc, double cutoff, int64_t capacity,
                       int32_t* idx_i, int32_t* idx_j, int32_t* shifts, int64_t* count,
                       void* workspace, size_t workspace_bytes, int32_t* dev_status, void* stream);

/* ---- instrumentation ------------------------------------------------------------------------
 * Number of kernels this library has launched since process start (bench.py's gpu_launches). */
uint64_t so3k_launch_count(void);
/* Optional per-kernel-category timing with CUDA events recorded on the caller's stream around the launches of
 * so3k_energy_forces / so3k_potential_displacements.  so3k_profile_categories() returns the comma-separated
 * category names; so3k_profile_collect SYNCHRONISES on the recorded events and returns accumulated milliseconds
 * and scope counts per category (n_categories >= number of names). */
const char* so3k_profile_categories(void);
int so3k_profile_enable(so3k_handle* h, int32_t enable);
int so3k_profile_collect(so3k_handle* h, double* ms, int64_t* count, int32_t n_categories, int32_t reset);

/* Diagnostic: one 128 x N x K split-bf16 GEMM tile on the tcgen05 tensor cores with the shared-memory operand
 * layouts of the edge kernels (tests pin the descriptor conventions with it).  A (128,K) f32;
 * variant 0: B (N,K) f32, D = A @ B^T ; variant 2: B (K,N) f32 consumed as an MN-major operand, D = A @ B.
 * passes: 1 (bf16 x bf16) or 3 (split operands).  D (128,N) f32. */
int so3k_tc_selftest(int32_t N, int32_t K, int32_t variant, int32_t passes, const float* A, const float* B,
                     float* D, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SO3K_H_ */
