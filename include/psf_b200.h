/*
 * psf_b200.h -- C ABI of the B200-native PSF chord-factor chain.
 *
 * Plain C, raw device pointers and sizes, one cudaStream_t per call, int status back.
 * No torch types, no allocation inside the library, no global mutable state beyond the
 * one-time cudaFuncSetAttribute for large dynamic shared memory.  Every function is
 * re-entrant and stream-ordered; pointers must belong to the current CUDA device.
 *
 * Reference interfaces replaced (paths relative to the reference repository):
 *   psf_chord_offsets / psf_chord_indices_*  <- get_chord_indices_assym,
 *                                               SyntheticExperiments/psf.py:7-32
 *   psf_chord_mask_host                      <- chord_mask_mat, non-parametric/chord_mask_mat.m:1-38
 *   psf_factor_fwd                           <- forward_host,  spmul/spmul_cuda.cu:31-59
 *                                               (kernel :5-29), bound at spmul_cuda.cu:164
 *   psf_factor_bwd                           <- backward_host, spmul/spmul_cuda.cu:114-159
 *                                               (kernels :61-112), bound at spmul_cuda.cu:165
 *   psf_chain_fwd / psf_chain_bwd            <- the loop over torch_sparse.spmm in
 *                                               PSFNet.forward, SyntheticExperiments/psf.py:167-188
 *                                               (LRA/psf.py:219-240, Genome_Clf/psf.py:209-230)
 *                                               and its autograd backward.
 *
 * Math (sample b, row i, channel c, factor m = 1..n_W, taps k = 0..L-1 with offsets o_k):
 *   V_m[b,i,c] = sum_k W_m[b,i,k] * V_{m-1}[b,(i+o_k) mod n,c]   (+ V_0[b,i,c] if residual)
 */
#ifndef PSF_B200_H
#define PSF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* cudaStream_t without pulling in the CUDA headers */
typedef struct CUstream_st* psf_stream_t;

/* status codes: 0 ok; negative = argument/capability errors; positive = cudaError_t */
#define PSF_OK 0
#define PSF_ERR_ARG (-1)         /* bad shape / null pointer / misaligned buffer            */
#define PSF_ERR_UNSUPPORTED (-2) /* combination not implemented (e.g. too many taps)        */
#define PSF_ERR_WORKSPACE (-3)   /* workspace pointer null or smaller than the query result */

#define PSF_MAX_TAPS 32

typedef enum { PSF_F32 = 0, PSF_F64 = 1 } psf_dtype;

/* Which chain implementation ran / should run. */
typedef enum {
  PSF_PATH_AUTO = 0,     /* fused resident kernel when the shape allows, else per-factor  */
  PSF_PATH_GENERIC = 1,  /* one launch per factor, V through L2/HBM (any shape, f32/f64)  */
  PSF_PATH_FUSED = 2,    /* cluster kernel, V block resident in SMEM across all factors   */
  PSF_PATH_FUSED_GEN1 = 3 /* request only: the first-generation cluster kernel (contiguous row
                            blocks); reported back as PSF_PATH_FUSED                      */
} psf_path;

/* Element strides of a stack of factors W[factor][sample][row][tap].
 * Reference layout  ("row-major", W_m = fs[m](data) of shape [B,n,L], one tensor per m
 * stacked to [n_W,B,n,L]):      factor=B*n*L sample=n*L row=L tap=1
 * Streaming layout  ("tap-major", what the fused kernel wants, [n_W,L,B,n]):
 *                               factor=L*B*n sample=n   row=1 tap=B*n
 *
 * Row order.  psf_chain_desc.interleave = Q > 1 says that the ROW axis of W and dW (not of V) is
 * stored "quad-interleaved": logical row i, quad u = i>>2, sits at position
 *     p(i) = (u mod Q) * (n/Q) + 4 * (u div Q) + (i & 3)
 * i.e. the row quads are dealt round-robin to Q blocks of n/Q rows.  The second-generation fused
 * kernel runs a Q-CTA cluster per sample in which CTA q owns block q; ask
 * psf_chain_preferred_interleave() for the Q a problem wants (1 = natural order) and produce W in
 * that order (row-wise generators only need their input rows permuted once).            */
typedef struct {
  int64_t factor, sample, row, tap;
} psf_wstrides;

int psf_version(void);
const char* psf_error_string(int status);

/* ---- integer part: chord pattern (bit-exact with psf.py:7-32) ------------------------ */
/* offsets_out[k], k < n_link: 0, then 2^(k-1) mod n_vec.  Host memory. */
int psf_chord_offsets(int64_t n_vec, int n_link, int64_t* offsets_out);
/* rows/cols: host arrays of n_vec*n_link int64; entry e=i*n_link+k -> (i, (i+o_k) mod n_vec). */
int psf_chord_indices_host(int64_t n_vec, int n_link, int64_t* rows, int64_t* cols);
/* index_out: device array [2, n_vec*n_link] int64 (rows then cols), what
 * torch.tensor(get_chord_indices_assym(...)).cuda() holds (psf.py:143-145). */
int psf_chord_indices_device(int64_t n_vec, int n_link, int64_t* index_out, psf_stream_t stream);
/* Dense N x N mask (row-major doubles, host), chord_mask_mat.m with random_order=false. */
int psf_chord_mask_host(int64_t N, int b, int symmetric, int self_loop, double* mask_out);

/* ---- one factor (SparseMultiply) ------------------------------------------------------- */
/* Z[b,p,d] = sum_k F[b,p,k] * V[b,(p+offsets[k]) mod n,d].  F [B,n,L], V/Z [B,n,D] contiguous.
 * offsets: HOST array of L non-negative int64 (any values; reduced mod n).  Z is overwritten. */
int psf_factor_fwd(psf_dtype dtype, const void* F, const void* V, void* Z, const int64_t* offsets,
                   int64_t B, int64_t n, int L, int64_t D, psf_stream_t stream);
/* dF[b,p,k] = sum_d dZ[b,p,d]*V[b,(p+o_k) mod n,d];  dV[b,p,d] = sum_k F[b,j,k]*dZ[b,j,d], j=(p-o_k) mod n.
 * dF / dV are overwritten; either may be NULL to skip it. */
int psf_factor_bwd(psf_dtype dtype, const void* dZ, const void* F, const void* V, void* dF, void* dV,
                   const int64_t* offsets, int64_t B, int64_t n, int L, int64_t D, psf_stream_t stream);

/* ---- the fused chain ------------------------------------------------------------------- */
typedef struct {
  int64_t B;      /* samples                                         */
  int64_t n;      /* sequence length (rows)                          */
  int64_t C;      /* channels of V                                   */
  int n_W;        /* number of factors                               */
  int L;          /* taps per row (n_W+1 in PSFNet)                  */
  int residual;   /* 1: V_m = W_m V_{m-1} + V_0  (psf.py:168-169,187-188) */
  psf_dtype dtype;
  psf_wstrides ws;            /* strides of W (and of dW)            */
  const int64_t* offsets;     /* HOST array [L]; NULL = chord offsets 0,1,2,4,... mod n */
  psf_path path;              /* PSF_PATH_AUTO normally              */
  int interleave;             /* row order of W/dW: 0 or 1 natural, Q > 1 quad-interleaved (above);
                                 Q > 1 is accepted only when it equals psf_chain_preferred_interleave() */
} psf_chain_desc;

/* Bytes of device workspace psf_chain_fwd / psf_chain_bwd need for this problem.
 * save_states=1: forward keeps V_0..V_{n_W-1} in the workspace for the backward pass
 * (the same workspace pointer must then be passed to psf_chain_bwd, untouched in between). */
size_t psf_chain_workspace_bytes(const psf_chain_desc* d, int save_states);

/* out[B,n,C] = chain(W, V0).  V0/out contiguous [B,n,C].  Returns status; *path_used (may be
 * NULL) reports which implementation ran. */
int psf_chain_fwd(const psf_chain_desc* d, const void* W, const void* V0, void* out, void* workspace,
                  size_t workspace_bytes, int save_states, int* path_used, psf_stream_t stream);

/* Given dOut[B,n,C] and the workspace written by psf_chain_fwd(save_states=1):
 * dW (same strides as W, every element written) and dV0[B,n,C]. */
int psf_chain_bwd(const psf_chain_desc* d, const void* W, const void* dOut, void* dW, void* dV0,
                  void* workspace, size_t workspace_bytes, int* path_used, psf_stream_t stream);

/* Layout helper: dst (strides ds) = src (strides ss), elementwise copy of an
 * [n_W,B,n,L] factor stack (f32/f64).  Used to move between the reference layout and the
 * streaming layout. */
int psf_w_relayout(psf_dtype dtype, const void* src, psf_wstrides ss, void* dst, psf_wstrides ds,
                   int n_W, int64_t B, int64_t n, int L, psf_stream_t stream);

/* Same with a row order per side (1 = natural, Q > 1 = quad-interleaved; needs n % (4Q) == 0). */
int psf_w_relayout_ex(psf_dtype dtype, const void* src, psf_wstrides ss, int src_interleave, void* dst,
                      psf_wstrides ds, int dst_interleave, int n_W, int64_t B, int64_t n, int L,
                      psf_stream_t stream);

/* Row interleave Q the fastest fused kernel wants for this problem (desc.interleave and desc.ws are
 * ignored): Q = n/2048 for f32, C = 8, n = 2^p in [4096, 16384], L = p+1 chord offsets; n/4096 for
 * n = 32768, 65536; otherwise 1. */
int psf_chain_preferred_interleave(const psf_chain_desc* d);

/* Which fused kernel generation psf_chain_fwd would launch for this descriptor: 2 (quad-interleaved
 * clusters), 1 (contiguous row blocks) or 0 (per-factor path). */
int psf_chain_fused_generation(const psf_chain_desc* d);

/* ---- the generators (fs[m] and g: Linear(E,H) -> GELU -> Linear(H,out), psf.py:110-126; called at
 * psf.py:165 (V = g(data)) and :175 (W = fs[m](data))) -------------------------------------------------
 * All n_W W-generators, and optionally the V generator g, in ONE kernel each way, fp32 FMA.
 * Supported: E = H = 32, L <= 16, C <= 16 (else PSF_ERR_UNSUPPORTED and the caller keeps its own GEMM path).
 *   x [rows,32] row-major, rows = B*n, row order given by `interleave` (1 natural; Q > 1 quad-interleaved
 *   per sample of n rows, as in psf_chain_desc).  Parameters in nn.Linear layout: W1 [.,32,32] (out,in),
 *   b1 [.,32], W2 [.,out,32], b2 [.,out]; `fs` tensors are stacked over the n_W generators, `g` is single.
 *   W_out / dW_in: element (m, l, row) at m*stride_factor + l*stride_tap + row (the chain's tap-major
 *   layout, row order = that of x).  V_out / dV_in: [rows, C] row-major in NATURAL row order (what the chain
 *   takes as V0).  C = 0: no V generator (g pointers ignored).  The backward pass needs
 *   psf_generators_workspace_bytes(n_W) bytes of device workspace (overwritten). */
typedef struct {
  int64_t rows;                       /* B*n                                                   */
  int64_t n;                          /* rows per sample (used when interleave > 1 or C > 0)   */
  int n_W, E, H, L;
  int64_t stride_factor, stride_tap;  /* layout of W_out / dW_in                               */
  int interleave;                     /* row order of x and of W                               */
  int C;                              /* outputs of the V generator g, 0 = none                */
} psf_gen_desc;
typedef struct { const float *W1, *b1, *W2, *b2; } psf_mlp_params;
typedef struct { float *W1, *b1, *W2, *b2; } psf_mlp_grads;

size_t psf_generators_workspace_bytes(int n_W);
/* pre_save (may be NULL): [n_W + (C > 0)][32][rows] floats receiving the pre-activations for the backward pass. */
int psf_generators_fwd(const psf_gen_desc* d, const float* x, psf_mlp_params fs, psf_mlp_params g, float* W_out,
                       float* V_out, float* pre_save, psf_stream_t stream);
/* pre_saved (may be NULL): what the forward pass wrote; NULL = recompute the pre-activations from x. */
int psf_generators_bwd(const psf_gen_desc* d, const float* x, psf_mlp_params fs, psf_mlp_params g,
                       const float* dW_in, const float* dV_in, const float* pre_saved, float* dx, psf_mlp_grads dfs,
                       psf_mlp_grads dg, void* workspace, size_t workspace_bytes, psf_stream_t stream);

/* Launch geometry the fused path would use for this problem (for benches / docs):
 * cluster size, rows per CTA, threads per CTA, dynamic SMEM bytes, resident clusters.
 * Returns PSF_ERR_UNSUPPORTED when the fused path does not cover the shape. */
int psf_chain_fused_plan(const psf_chain_desc* d, int* cluster, int* rows_per_cta, int* threads,
                         int* smem_bytes, int* resident_clusters);

/* ---- Head of the FLATTEN pooling (reference: self.final = nn.Linear(n_vec * n_channels_V, n_class) applied to
 * V.view(B, -1), SyntheticExperiments/psf.py:60,80; LRA/psf.py:13-22,109; Genome_Clf/psf.py) --------------------------
 * logits[b][c] = bias[c] + sum_k V[b][k] W[c][k];  V [B,K] row-major, W [n_class,K] (nn.Linear layout), fp32.
 * Streaming kernels (both directions are HBM-bound passes over V), fixed summation order, no atomics.
 * Supported: n_class <= PSF_HEAD_MAX_CLASSES, K % 4 == 0, B <= 65535, 16-byte aligned pointers; else
 * PSF_ERR_UNSUPPORTED and the caller keeps its own GEMM.  bias may be NULL.  The backward pass writes
 * dV [B,K] and dW [n_class,K] from g = dlogits [B,n_class] (dbias = column sums of g, left to the caller). */
#define PSF_HEAD_MAX_CLASSES 16
size_t psf_head_workspace_bytes(int64_t B, int64_t K, int n_class, int backward);
int psf_head_fwd(const float* V, const float* W, const float* bias, float* out, int64_t B, int64_t K, int n_class,
                 void* workspace, size_t workspace_bytes, psf_stream_t stream);
int psf_head_bwd(const float* V, const float* W, const float* g, float* dV, float* dW, int64_t B, int64_t K, int n_class,
                 void* workspace, size_t workspace_bytes, psf_stream_t stream);

/* ---- init_linear of the Adding model (reference: self.init_linear = nn.Linear(2, embedding_size) applied to the raw
 * [B, n, 2] input, SyntheticExperiments/psf.py:136-141,155-156) ----------------------------------------------------
 * y[row][e] = b[e] + sum_i x[row][i] W[e][i];  x [rows, in] row-major, W [out, in] (nn.Linear layout), fp32.
 * Supported: in <= 4, out % 4 == 0, out <= 128 with 256 % (out / 4) == 0; else PSF_ERR_UNSUPPORTED.
 * Backward: dW [out, in] and db [out] (may be NULL) from dy [rows, out]; the raw input takes no gradient.
 * Fixed summation order, no atomics. */
size_t psf_init_linear_workspace_bytes(int64_t rows, int in_features, int out_features);
int psf_init_linear_fwd(const float* x, const float* W, const float* b, float* y, int64_t rows, int in_features,
                        int out_features, psf_stream_t stream);
int psf_init_linear_bwd(const float* x, const float* dy, float* dW, float* db, int64_t rows, int in_features,
                        int out_features, void* workspace, size_t workspace_bytes, psf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PSF_B200_H */
