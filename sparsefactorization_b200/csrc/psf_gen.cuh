// Declarations shared by the generator translation units (psf_gen_sm100.cu: FP32-FMA / mma.sync kernels and the
// C-ABI entry points; psf_gen_tc_sm100.cu: the tcgen05 / tensor-memory kernels).
#pragma once
#include "psf_internal.cuh"

constexpr int GE = 32;    // embedding width
constexpr int GH = 32;    // hidden width
constexpr int GLP = 16;   // taps padded
constexpr int kAccBlk = GH * GE + GH + GLP * GH + GLP;  // parameter-gradient block of one generator: dW1, db1, dW2 (l-major, padded), db2

struct GenParams {
  const float* x;
  const float* W1;  // [n_W][32][32]
  const float* b1;  // [n_W][32]
  const float* W2;  // [n_W][L][32]
  const float* b2;  // [n_W][L]
  const float* gW1; // optional V generator g (psf.py:126): [32][32], [32], [C][32], [C]; generator index n_W
  const float* gb1;
  const float* gW2;
  const float* gb2;
  float* Wt;        // fwd: out [n_W][L][rows];  bwd: dW (same layout, read)
  float* V;         // fwd: V0 [rows][C] in NATURAL row order;  bwd: dV0 (read)
  float* dx;        // bwd: [rows][32]
  float* ws;        // bwd: [grid][n_gen][kAccBlk]
  float* pre;       // optional [n_gen][32][rows]: pre-activations saved by the forward pass for the backward pass
  long long rows, sWf, sWt;
  int n_W, L;
  int n_gen;        // n_W + (g present)
  int C;            // outputs of g
  int n, Q;         // rows per sample and row interleave of x (to find the natural row of a V0 row)
};

// generator m's parameter tensors and output count
static __device__ __forceinline__ void gen_ptrs(const GenParams& p, int m, const float*& W1, const float*& b1, const float*& W2,
                                         const float*& b2, int& Lm) {
  if (m < p.n_W) {
    W1 = p.W1 + (long long)m * GH * GE;
    b1 = p.b1 + m * GH;
    W2 = p.W2 + (long long)m * p.L * GH;
    b2 = p.b2 + m * p.L;
    Lm = p.L;
  } else {
    W1 = p.gW1;
    b1 = p.gb1;
    W2 = p.gW2;
    b2 = p.gb2;
    Lm = p.C;
  }
}
// natural row index of row `row` of a quad-interleaved x (include/psf_b200.h)
static __device__ __forceinline__ long long natural_row(const GenParams& p, long long row) {
  if (p.Q <= 1) return row;
  const long long b = row / p.n;
  const int pos = (int)(row - b * p.n);
  const int J = p.n / p.Q;
  const int q = pos / J, r = pos - q * J;
  return b * p.n + 4 * ((r >> 2) * p.Q + q) + (r & 3);
}

// tcgen05 kernels (psf_gen_tc_sm100.cu).  Return PSF_ERR_UNSUPPORTED when the shape is not covered (the caller then
// launches the FP32-FMA kernel), else the launch status.
int psf_gen_tc_fwd(const GenParams& p, cudaStream_t st);
// backward: `images` = psf_gen_tc_bwd_image_bytes(n_gen) bytes of 128-byte aligned device scratch (overwritten); p.ws =
// zeroed [*grid_out][n_gen][kAccBlk] accumulator slices (the caller reduces them)
size_t psf_gen_tc_bwd_image_bytes(int n_gen);
int psf_gen_tc_prep(const GenParams& p, void* images, cudaStream_t st);   // weights -> MMA-ready images (both kernels below)
int psf_gen_tc_bwd(const GenParams& p, void* images, int* grid_out, cudaStream_t st);
// second version (psf_gen_tc_bwd2_sm100.cu): parameter gradients on tcgen05 too; whole 128-row tiles only; writes dW1
// TRANSPOSED ([e][h]) into the accumulator slices
int psf_gen_tc_bwd2(const GenParams& p, const void* images, int* grid_out, cudaStream_t st);
