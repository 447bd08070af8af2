// Internal declarations shared by the translation units of libpsf_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include "psf_b200.h"

#define PSF_CUDA_TRY(expr)                       \
  do {                                           \
    cudaError_t _e = (expr);                     \
    if (_e != cudaSuccess) return (int)_e;       \
  } while (0)

// Tap offsets, already reduced mod n, passed by value in kernel parameters.
struct PsfOffsets {
  int o[PSF_MAX_TAPS];
};

// Normalised description used by both chain paths.
struct PsfChainArgs {
  int64_t B, n, C;
  int n_W, L, residual;
  psf_wstrides ws;
  PsfOffsets offs;
  int interleave;  // row order of W/dW: 1 natural, Q > 1 quad-interleaved (see psf_b200.h)
};

static inline size_t psf_round256(size_t x) { return (x + 255) & ~size_t(255); }

// One-time per-device state (cudaFuncSetAttribute opt-ins, occupancy results) is keyed by the device ordinal and guarded
// by a mutex: a process may drive several devices from several threads.
constexpr int PSF_MAX_DEVICES = 64;
static inline int psf_current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= PSF_MAX_DEVICES) return -1;
  return d;
}

// ---- generic (one launch per factor) ----------------------------------------------------
// W_m points at factor m (sample/row/tap strides apply).  res may be null.
template <typename T>
int psf_generic_factor_fwd(const T* Wm, int64_t s_sample, int64_t s_row, int64_t s_tap, const T* V, const T* res,
                           T* Z, const PsfOffsets& offs, int64_t B, int64_t n, int L, int64_t D, cudaStream_t st);
template <typename T>
int psf_generic_factor_dV(const T* Wm, int64_t s_sample, int64_t s_row, int64_t s_tap, const T* dZ, T* dV,
                          const PsfOffsets& offs, int64_t B, int64_t n, int L, int64_t D, cudaStream_t st);
template <typename T>
int psf_generic_factor_dW(const T* dZ, const T* V, T* dWm, int64_t s_sample, int64_t s_row, int64_t s_tap,
                          const PsfOffsets& offs, int64_t B, int64_t n, int L, int64_t D, cudaStream_t st);
template <typename T>
int psf_generic_axpy(T* acc, const T* x, const T* y, int64_t count, cudaStream_t st);  // acc = x + y (y may be null)

int psf_generic_relayout(psf_dtype dtype, const void* src, psf_wstrides ss, int src_interleave, void* dst,
                         psf_wstrides ds, int dst_interleave, int n_W, int64_t B, int64_t n, int L, cudaStream_t st);

template <typename T>
int psf_generic_chain_fwd(const PsfChainArgs& a, const T* W, const T* V0, T* out, T* ws, int save_states,
                          cudaStream_t st);
template <typename T>
int psf_generic_chain_bwd(const PsfChainArgs& a, const T* W, const T* dOut, T* dW, T* dV0, T* ws, cudaStream_t st);

// ---- fused resident cluster kernels (f32, tap-major W, see psf_chain_sm100.cu) -----------
struct PsfFusedPlan {
  int cluster;        // CTAs per sample
  int rows_per_cta;   // RB, multiple of 4
  int threads;        // per CTA
  int smem_bytes;     // dynamic shared memory, forward kernel
  int resident;       // clusters that fit on the device at once, forward kernel
  int smem_bytes_bwd; // same for the backward kernel (different W ring geometry)
  int resident_bwd;
  int stages;         // W ring depth, forward kernel
  int stages_bwd;     // W ring depth, backward kernel
};
int psf_fused_plan(const PsfChainArgs& a, PsfFusedPlan* plan);   // PSF_OK or PSF_ERR_UNSUPPORTED
bool psf_fused_buffers_ok(const PsfChainArgs& a, const void* W, const void* p1, const void* p2, const void* ws);
int psf_fused_chain_fwd(const PsfChainArgs& a, const float* W, const float* V0, float* out, float* ws,
                        int save_states, cudaStream_t st);
int psf_fused_chain_bwd(const PsfChainArgs& a, const float* W, const float* dOut, float* dW, float* dV0, float* ws,
                        cudaStream_t st);

// ---- second-generation fused kernels (quad-interleaved clusters, see psf_chain2_sm100.cu) ---
int psf_chain2_interleave_for(const PsfChainArgs& a);           // Q the kernels want for this shape, 0 = not covered
int psf_chain2_plan(const PsfChainArgs& a, PsfFusedPlan* plan);  // PSF_OK iff shape, layout and a.interleave fit
int psf_chain2_fwd(const PsfChainArgs& a, const float* W, const float* V0, float* out, float* ws, int save_states,
                   cudaStream_t st);
int psf_chain2_bwd(const PsfChainArgs& a, const float* W, const float* dOut, float* dW, float* dV0, float* ws,
                   cudaStream_t st);
