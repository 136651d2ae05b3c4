// Repacked filter weights of one block (fb or gb) of one layer, as consumed by the edge kernels.
#pragma once
#include "so3k_internal.h"

struct EdgeBlockW {
  const float* W1;    // (K, F)        rad layers_0 kernel
  const float* b1;    // (F)
  const float* W2c;   // (KC, F)       rows [0,F) = rad layers_1 kernel, rows [F,F+Fs) = sph layers_1 kernel, rest 0
  const float* W2cT;  // (F, KC)       transpose of W2c
  const float* b2c;   // (F)           rad layers_1 bias + sph layers_1 bias
  const float* Ws1;   // (nl, Fs)      sph layers_0 kernel
  const float* bs1;   // (Fs)
};

bool so3k_edge_supported(const So3kDims& D);
size_t so3k_edge_repack_floats(const So3kDims& D);   // per block
void so3k_edge_repack(const So3kDims& D, const float* params, const BlockParams& bp, float* dst, EdgeBlockW* out,
                      cudaStream_t st);
void so3k_launch_edge_fwd(const So3kDims& D, const Graph& g, const EdgeBlockW& Wf, const EdgeBlockW& Wg,
                          const float* chi, const float* proj, float* alpha, cudaStream_t st);
// parts: bit 0 = q/k projection gradients (gproj, atomics) and the cutoff cotangent ebar[e][1];
//        bit 1 = radial cotangent ebar[e][0] and the l0-contraction cotangent gbar.
void so3k_launch_edge_bwd(const So3kDims& D, const Graph& g, const EdgeBlockW& Wf, const EdgeBlockW& Wg,
                          const float* chi, const float* proj, const float* alphabar, int parts, float* gproj,
                          float* gbar, float* ebar, cudaStream_t st);
void so3k_launch_edge_finish(const So3kDims& D, const Graph& g, const float* alpha, const float* chibar1,
                             const float* ebar, float* rbar, cudaStream_t st);

// ---- tcgen05 path (so3k_edge_tc.cu) -----------------------------------------------------------------------------
struct TcBlockW {
  const unsigned char* img;   // [W1 hi | lo | W2c hi | lo | Ws1 hi | lo] bf16 operand images (shared-memory layout)
  const float* bias;          // [b1x = (rad b1 | sph b1 | 0) over the KC hidden units | b2c padded to Fp]
  const float* Ws1;           // (nl, Fs) fp32 (epilogue of backward part 2)
  const float* bs1;           // (Fs)
  // wide models (F > 144, e.g. F = 256): the block runs as split x split sub-blocks of the narrow kernels -- sub-block
  // (kh, nh) = hidden units half kh (radial [kh F/2, (kh+1) F/2), spherical [kh Fs/2, ..)) -> output columns half nh; its
  // images sit at simg / sbias[kh * 2 + nh].  split == 1: simg[0] == img.
  int split = 1;
  const unsigned char* simg[4] = {nullptr, nullptr, nullptr, nullptr};
  const float* sbias[4] = {nullptr, nullptr, nullptr, nullptr};
};
bool so3k_edge_tc_supported(const So3kDims& D);
int so3k_edge_tc_split(const So3kDims& D);            // 1: narrow kernels as they are ; 2: 2 x 2 sub-blocks ; 0: unsupported
size_t so3k_edge_tc_image_bytes(const So3kDims& D);   // per block
void so3k_edge_tc_build(const So3kDims& D, const float* params, const BlockParams& bp, unsigned char* dst, TcBlockW* out,
                        const EdgeBlockW& simt, cudaStream_t st);
// filter w[pair][F] of one block for every canonical (undirected) pair of the list
// rec[pair capacity][8]: per canonical pair (d, l0 contractions of chi_j - chi_i) of the current layer
void so3k_launch_pair_records(const So3kDims& D, const Graph& g, const float* chi, float* rec, cudaStream_t st);
// tmp: scratch of pair capacity x F / 2 floats, needed when so3k_edge_tc_split(D) == 2 (sub-block outputs before they are
// combined into the F-wide rows)
void so3k_launch_filter_tc(const So3kDims& D, const Graph& g, const TcBlockW& W, const float* rec, float* wst, float* tmp,
                           int* status, cudaStream_t st);
// radial cotangent ebar[e][0] and l0-contraction cotangent gbar of both blocks, added onto the canonical edge of each
// pair (the arrays must be zero on entry); wbar: pair-indexed filter cotangents from so3k_launch_qk_bwd_stream
void so3k_launch_edge_bwd2_tc(const So3kDims& D, const Graph& g, const TcBlockW& Wf, const TcBlockW& Wg, const float* rec,
                              const float* wbar, float* gbar, float* ebar, int* status, cudaStream_t st);

// C (rows, N) = A (rows, K; row stride lda) @ B (+ bias on the first bias_rows rows) on the tensor cores with split-bf16
// operands; B(k, n) = Bp[k * sbk + n * sbn]; scratch: so3k_tc_gemm_scratch_bytes() bytes for the weight image.  Returns
// false (nothing launched) when the shape is outside the kernel's limits -- the caller then runs its fp32 SIMT GEMM.
size_t so3k_tc_gemm_scratch_bytes();
bool so3k_tc_gemm(int rows, int N, int K, const float* A, int lda, const float* Bp, int sbk, int sbn, const float* bias,
                  int bias_rows, float* C, int ldc, unsigned char* scratch, int* status, cudaStream_t st);
