// extern "C" entry points of libpsf_b200.so (declared in include/psf_b200.h).
#include <cstring>

#include "psf_internal.cuh"

namespace {

int normalise_offsets(const int64_t* offsets, int64_t n, int L, PsfOffsets* out) {
  if (L < 1 || L > PSF_MAX_TAPS || n < 1 || n > (int64_t(1) << 30)) return PSF_ERR_ARG;
  for (int k = 0; k < L; ++k) {
    int64_t o;
    if (offsets != nullptr) {
      o = offsets[k];
      if (o < 0) return PSF_ERR_ARG;
    } else {
      o = (k == 0) ? 0 : (int64_t(1) << (k - 1));
    }
    out->o[k] = (int)(o % n);
  }
  for (int k = L; k < PSF_MAX_TAPS; ++k) out->o[k] = 0;
  return PSF_OK;
}

int make_args(const psf_chain_desc* d, PsfChainArgs* a) {
  if (d == nullptr || d->B < 0 || d->n < 1 || d->C < 1 || d->n_W < 0) return PSF_ERR_ARG;
  if (d->dtype != PSF_F32 && d->dtype != PSF_F64) return PSF_ERR_ARG;
  a->B = d->B;
  a->n = d->n;
  a->C = d->C;
  a->n_W = d->n_W;
  a->L = d->L;
  a->residual = d->residual ? 1 : 0;
  a->ws = d->ws;
  a->interleave = d->interleave < 1 ? 1 : d->interleave;
  return normalise_offsets(d->offsets, d->n, d->L, &a->offs);
}

// Which fused generation serves this call: 2, 1, or 0 = per-factor path (then *status says whether that is
// acceptable).  An interleaved row order exists only for generation 2.
int use_fused(const psf_chain_desc* d, const PsfChainArgs& a, const void* W, const void* p1, const void* p2,
              const void* ws, int* status) {
  *status = PSF_OK;
  PsfFusedPlan plan;
  const bool f32 = d->dtype == PSF_F32;
  if (a.interleave > 1) {
    if (d->path != PSF_PATH_GENERIC && d->path != PSF_PATH_FUSED_GEN1 && f32 && psf_chain2_plan(a, &plan) == PSF_OK &&
        psf_fused_buffers_ok(a, W, p1, p2, ws))
      return 2;
    *status = PSF_ERR_UNSUPPORTED;
    return 0;
  }
  if (d->path == PSF_PATH_GENERIC) return 0;
  if (!f32 || !psf_fused_buffers_ok(a, W, p1, p2, ws)) {
    if (d->path != PSF_PATH_AUTO) *status = PSF_ERR_UNSUPPORTED;
    return 0;
  }
  if (d->path != PSF_PATH_FUSED_GEN1 && psf_chain2_plan(a, &plan) == PSF_OK) return 2;
  if (psf_fused_plan(a, &plan) == PSF_OK) return 1;
  if (d->path != PSF_PATH_AUTO) *status = PSF_ERR_UNSUPPORTED;
  return 0;
}

}  // namespace

extern "C" {

int psf_version(void) { return 100; }

const char* psf_error_string(int status) {
  switch (status) {
    case PSF_OK: return "ok";
    case PSF_ERR_ARG: return "psf: invalid argument";
    case PSF_ERR_UNSUPPORTED: return "psf: unsupported shape / layout for the requested path";
    case PSF_ERR_WORKSPACE: return "psf: workspace missing or too small";
    default: return status > 0 ? cudaGetErrorString((cudaError_t)status) : "psf: unknown error";
  }
}

int psf_chord_offsets(int64_t n_vec, int n_link, int64_t* offsets_out) {
  if (n_vec < 1 || n_link < 1 || n_link > 62 || offsets_out == nullptr) return PSF_ERR_ARG;
  offsets_out[0] = 0;
  for (int k = 1; k < n_link; ++k) offsets_out[k] = (int64_t(1) << (k - 1)) % n_vec;
  return PSF_OK;
}

int psf_chord_indices_host(int64_t n_vec, int n_link, int64_t* rows, int64_t* cols) {
  if (n_vec < 1 || n_link < 1 || n_link > 62 || rows == nullptr || cols == nullptr) return PSF_ERR_ARG;
  for (int64_t i = 0; i < n_vec; ++i) {
    int64_t* r = rows + i * n_link;
    int64_t* c = cols + i * n_link;
    r[0] = i;
    c[0] = i;
    for (int k = 1; k < n_link; ++k) {
      r[k] = i;
      c[k] = (i + ((int64_t(1) << (k - 1)) % n_vec)) % n_vec;
    }
  }
  return PSF_OK;
}

int psf_chord_mask_host(int64_t N, int b, int symmetric, int self_loop, double* mask_out) {
  if (N < 1 || b < 2 || mask_out == nullptr) return PSF_ERR_ARG;
  std::memset(mask_out, 0, sizeof(double) * (size_t)N * (size_t)N);
  // L = floor(log_b N): largest L with b^L <= N
  int L = 0;
  for (int64_t pw = b; pw <= N; pw *= b) ++L;
  for (int64_t i = 0; i < N; ++i) {
    int64_t step = 1;
    for (int k = 0; k < L; ++k, step *= b) {
      mask_out[i * N + (i + step) % N] = 1.0;
      if (symmetric) mask_out[i * N + ((i - step) % N + N) % N] = 1.0;
    }
    if (self_loop) mask_out[i * N + i] = 1.0;
  }
  return PSF_OK;
}

}  // extern "C"

namespace {
__global__ void chord_indices_kernel(int64_t* __restrict__ out, int64_t n_vec, int n_link) {
  const int64_t nnz = n_vec * n_link;
  for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < nnz; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / n_link;
    const int k = (int)(e - i * n_link);
    const int64_t off = (k == 0) ? 0 : ((int64_t(1) << (k - 1)) % n_vec);
    out[e] = i;
    out[nnz + e] = (i + off) % n_vec;
  }
}
}  // namespace

extern "C" {

int psf_chord_indices_device(int64_t n_vec, int n_link, int64_t* index_out, psf_stream_t stream) {
  if (n_vec < 1 || n_link < 1 || n_link > 62 || index_out == nullptr) return PSF_ERR_ARG;
  const int64_t nnz = n_vec * n_link;
  int64_t grid = (nnz + 255) / 256;
  if (grid > 148 * 8) grid = 148 * 8;
  chord_indices_kernel<<<(int)grid, 256, 0, (cudaStream_t)stream>>>(index_out, n_vec, n_link);
  return (int)cudaGetLastError();
}

int psf_factor_fwd(psf_dtype dtype, const void* F, const void* V, void* Z, const int64_t* offsets, int64_t B,
                   int64_t n, int L, int64_t D, psf_stream_t stream) {
  if (F == nullptr || V == nullptr || Z == nullptr || B < 0 || D < 0) return PSF_ERR_ARG;
  PsfOffsets offs;
  int rc = normalise_offsets(offsets, n, L, &offs);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PSF_F32)
    return psf_generic_factor_fwd<float>((const float*)F, n * L, L, 1, (const float*)V, nullptr, (float*)Z, offs, B, n, L,
                                         D, st);
  if (dtype == PSF_F64)
    return psf_generic_factor_fwd<double>((const double*)F, n * L, L, 1, (const double*)V, nullptr, (double*)Z, offs, B,
                                          n, L, D, st);
  return PSF_ERR_ARG;
}

int psf_factor_bwd(psf_dtype dtype, const void* dZ, const void* F, const void* V, void* dF, void* dV,
                   const int64_t* offsets, int64_t B, int64_t n, int L, int64_t D, psf_stream_t stream) {
  if (dZ == nullptr || F == nullptr || V == nullptr || B < 0 || D < 0) return PSF_ERR_ARG;
  PsfOffsets offs;
  int rc = normalise_offsets(offsets, n, L, &offs);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == PSF_F32) {
    if (dF) rc = psf_generic_factor_dW<float>((const float*)dZ, (const float*)V, (float*)dF, n * L, L, 1, offs, B, n, L, D, st);
    if (rc) return rc;
    if (dV) rc = psf_generic_factor_dV<float>((const float*)F, n * L, L, 1, (const float*)dZ, (float*)dV, offs, B, n, L, D, st);
    return rc;
  }
  if (dtype == PSF_F64) {
    if (dF) rc = psf_generic_factor_dW<double>((const double*)dZ, (const double*)V, (double*)dF, n * L, L, 1, offs, B, n, L, D, st);
    if (rc) return rc;
    if (dV) rc = psf_generic_factor_dV<double>((const double*)F, n * L, L, 1, (const double*)dZ, (double*)dV, offs, B, n, L, D, st);
    return rc;
  }
  return PSF_ERR_ARG;
}

size_t psf_chain_workspace_bytes(const psf_chain_desc* d, int save_states) {
  if (d == nullptr || d->B < 0 || d->n < 1 || d->C < 1 || d->n_W < 0) return 0;
  const size_t es = d->dtype == PSF_F64 ? 8 : 4;
  const size_t S = (size_t)d->B * (size_t)d->n * (size_t)d->C;
  const size_t slots = save_states ? (size_t)(d->n_W > 0 ? d->n_W : 1) + 3 : 3;
  return psf_round256(slots * S * es);
}

int psf_chain_fwd(const psf_chain_desc* d, const void* W, const void* V0, void* out, void* workspace,
                  size_t workspace_bytes, int save_states, int* path_used, psf_stream_t stream) {
  PsfChainArgs a;
  int rc = make_args(d, &a);
  if (rc) return rc;
  if (a.B == 0) return PSF_OK;
  if (V0 == nullptr || out == nullptr || (W == nullptr && a.n_W > 0)) return PSF_ERR_ARG;
  if (workspace == nullptr || workspace_bytes < psf_chain_workspace_bytes(d, save_states)) return PSF_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  int status;
  const int gen = use_fused(d, a, W, V0, out, workspace, &status);
  if (gen) {
    if (path_used) *path_used = PSF_PATH_FUSED;
    if (gen == 2)
      return psf_chain2_fwd(a, (const float*)W, (const float*)V0, (float*)out, (float*)workspace, save_states, st);
    return psf_fused_chain_fwd(a, (const float*)W, (const float*)V0, (float*)out, (float*)workspace, save_states, st);
  }
  if (status) return status;
  if (path_used) *path_used = PSF_PATH_GENERIC;
  if (d->dtype == PSF_F32)
    return psf_generic_chain_fwd<float>(a, (const float*)W, (const float*)V0, (float*)out, (float*)workspace,
                                        save_states, st);
  return psf_generic_chain_fwd<double>(a, (const double*)W, (const double*)V0, (double*)out, (double*)workspace,
                                       save_states, st);
}

int psf_chain_bwd(const psf_chain_desc* d, const void* W, const void* dOut, void* dW, void* dV0, void* workspace,
                  size_t workspace_bytes, int* path_used, psf_stream_t stream) {
  PsfChainArgs a;
  int rc = make_args(d, &a);
  if (rc) return rc;
  if (a.B == 0) return PSF_OK;
  if (dOut == nullptr || dV0 == nullptr || ((W == nullptr || dW == nullptr) && a.n_W > 0)) return PSF_ERR_ARG;
  if (workspace == nullptr || workspace_bytes < psf_chain_workspace_bytes(d, 1)) return PSF_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  int status;
  // the saved states are laid out per path, so the backward pass must take the path the forward took
  int gen = use_fused(d, a, W, dOut, dV0, workspace, &status);
  if (gen && (reinterpret_cast<uintptr_t>(dW) & 15) != 0) {
    gen = 0;
    if (d->path != PSF_PATH_AUTO || a.interleave > 1) status = PSF_ERR_UNSUPPORTED;
  }
  if (gen) {
    if (path_used) *path_used = PSF_PATH_FUSED;
    if (gen == 2)
      return psf_chain2_bwd(a, (const float*)W, (const float*)dOut, (float*)dW, (float*)dV0, (float*)workspace, st);
    return psf_fused_chain_bwd(a, (const float*)W, (const float*)dOut, (float*)dW, (float*)dV0, (float*)workspace, st);
  }
  if (status) return status;
  if (path_used) *path_used = PSF_PATH_GENERIC;
  if (d->dtype == PSF_F32)
    return psf_generic_chain_bwd<float>(a, (const float*)W, (const float*)dOut, (float*)dW, (float*)dV0,
                                        (float*)workspace, st);
  return psf_generic_chain_bwd<double>(a, (const double*)W, (const double*)dOut, (double*)dW, (double*)dV0,
                                       (double*)workspace, st);
}

int psf_w_relayout(psf_dtype dtype, const void* src, psf_wstrides ss, void* dst, psf_wstrides ds, int n_W, int64_t B,
                   int64_t n, int L, psf_stream_t stream) {
  if (src == nullptr || dst == nullptr || n_W < 0 || B < 0 || n < 0 || L < 1 || L > PSF_MAX_TAPS) return PSF_ERR_ARG;
  return psf_generic_relayout(dtype, src, ss, 1, dst, ds, 1, n_W, B, n, L, (cudaStream_t)stream);
}

int psf_w_relayout_ex(psf_dtype dtype, const void* src, psf_wstrides ss, int src_interleave, void* dst, psf_wstrides ds,
                      int dst_interleave, int n_W, int64_t B, int64_t n, int L, psf_stream_t stream) {
  if (src == nullptr || dst == nullptr || n_W < 0 || B < 0 || n < 0 || L < 1 || L > PSF_MAX_TAPS) return PSF_ERR_ARG;
  const int qs = src_interleave < 1 ? 1 : src_interleave, qd = dst_interleave < 1 ? 1 : dst_interleave;
  if ((qs > 1 && n % (4 * (int64_t)qs)) || (qd > 1 && n % (4 * (int64_t)qd))) return PSF_ERR_ARG;
  return psf_generic_relayout(dtype, src, ss, qs, dst, ds, qd, n_W, B, n, L, (cudaStream_t)stream);
}

int psf_chain_preferred_interleave(const psf_chain_desc* d) {
  PsfChainArgs a;
  if (make_args(d, &a) != PSF_OK || d->dtype != PSF_F32) return 1;
  const int Q = psf_chain2_interleave_for(a);
  return Q > 1 ? Q : 1;
}

int psf_chain_fused_generation(const psf_chain_desc* d) {
  PsfChainArgs a;
  if (make_args(d, &a) != PSF_OK || d->dtype != PSF_F32 || d->path == PSF_PATH_GENERIC) return 0;
  PsfFusedPlan plan;
  if (d->path != PSF_PATH_FUSED_GEN1 && psf_chain2_plan(a, &plan) == PSF_OK) return 2;
  if (a.interleave > 1) return 0;
  return psf_fused_plan(a, &plan) == PSF_OK ? 1 : 0;
}

int psf_chain_fused_plan(const psf_chain_desc* d, int* cluster, int* rows_per_cta, int* threads, int* smem_bytes,
                         int* resident_clusters) {
  PsfChainArgs a;
  int rc = make_args(d, &a);
  if (rc) return rc;
  if (d->dtype != PSF_F32) return PSF_ERR_UNSUPPORTED;
  PsfFusedPlan plan;
  if (d->path == PSF_PATH_FUSED_GEN1 || psf_chain2_plan(a, &plan) != PSF_OK) {
    if (a.interleave > 1) return PSF_ERR_UNSUPPORTED;
    rc = psf_fused_plan(a, &plan);
    if (rc) return rc;
  }
  if (cluster) *cluster = plan.cluster;
  if (rows_per_cta) *rows_per_cta = plan.rows_per_cta;
  if (threads) *threads = plan.threads;
  if (smem_bytes) *smem_bytes = plan.smem_bytes;
  if (resident_clusters) *resident_clusters = plan.resident;
  return PSF_OK;
}

}  // extern "C"
