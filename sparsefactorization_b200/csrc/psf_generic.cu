// Generic per-factor kernels: any n, C, L <= PSF_MAX_TAPS, any offsets, any W strides, f32/f64.
// One launch per factor; V travels through L2/HBM between factors.  This is the path behind
// SparseMultiply (single factor) and the fallback of the chain for shapes the fused cluster
// kernel does not cover.  Gather form everywhere: every output element is owned by one thread,
// no atomics, deterministic (like spmul_cuda.cu:24,79-80,105-108 in the reference).
#include "psf_internal.cuh"

namespace {

constexpr int kThreads = 256;

inline int grid_for(int64_t work, int threads) {
  int64_t g = (work + threads - 1) / threads;
  const int64_t cap = 148 * 32;  // grid-stride beyond ~32 CTAs per SM
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// Z[b,i,d] = sum_k W[b,i,k] * V[b,(i+o_k) mod n,d] (+ res[b,i,d]); taps accumulated k = 0..L-1.
template <typename T>
__global__ void __launch_bounds__(kThreads) factor_fwd_kernel(const T* __restrict__ W, int64_t s_sample, int64_t s_row,
                                                              int64_t s_tap, const T* __restrict__ V,
                                                              const T* __restrict__ res, T* __restrict__ Z,
                                                              PsfOffsets offs, int64_t B, int64_t n, int L, int64_t D) {
  const int64_t total = B * n * D;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d = idx % D;
    const int64_t bi = idx / D;
    const int64_t i = bi % n;
    const int64_t b = bi / n;
    const T* w = W + b * s_sample + i * s_row;
    const T* v = V + b * n * D + d;
    T acc = T(0);
    for (int k = 0; k < L; ++k) {
      int64_t j = i + offs.o[k];
      if (j >= n) j -= n;
      acc += w[k * s_tap] * v[j * D];
    }
    if (res != nullptr) acc += res[idx];
    Z[idx] = acc;
  }
}

// dV[b,j,d] = sum_k W[b,(j-o_k) mod n,k] * dZ[b,(j-o_k) mod n,d]
template <typename T>
__global__ void __launch_bounds__(kThreads) factor_dV_kernel(const T* __restrict__ W, int64_t s_sample, int64_t s_row,
                                                             int64_t s_tap, const T* __restrict__ dZ, T* __restrict__ dV,
                                                             PsfOffsets offs, int64_t B, int64_t n, int L, int64_t D) {
  const int64_t total = B * n * D;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t d = idx % D;
    const int64_t bj = idx / D;
    const int64_t j = bj % n;
    const int64_t b = bj / n;
    const T* w = W + b * s_sample;
    const T* g = dZ + b * n * D + d;
    T acc = T(0);
    for (int k = 0; k < L; ++k) {
      int64_t i = j - offs.o[k];
      if (i < 0) i += n;
      acc += w[i * s_row + k * s_tap] * g[i * D];
    }
    dV[idx] = acc;
  }
}

// dW[b,i,k] = sum_d dZ[b,i,d] * V[b,(i+o_k) mod n,d]
template <typename T>
__global__ void __launch_bounds__(kThreads) factor_dW_kernel(const T* __restrict__ dZ, const T* __restrict__ V,
                                                             T* __restrict__ dW, int64_t s_sample, int64_t s_row,
                                                             int64_t s_tap, PsfOffsets offs, int64_t B, int64_t n, int L,
                                                             int64_t D) {
  const int64_t total = B * n * L;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % L);
    const int64_t bi = idx / L;
    const int64_t i = bi % n;
    const int64_t b = bi / n;
    int64_t j = i + offs.o[k];
    if (j >= n) j -= n;
    const T* g = dZ + (b * n + i) * D;
    const T* v = V + (b * n + j) * D;
    T acc = T(0);
    for (int64_t d = 0; d < D; ++d) acc += g[d] * v[d];
    dW[b * s_sample + i * s_row + k * s_tap] = acc;
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) add_kernel(T* __restrict__ acc, const T* __restrict__ x,
                                                       const T* __restrict__ y, int64_t count) {
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < count;
       idx += (int64_t)gridDim.x * blockDim.x)
    acc[idx] = y ? x[idx] + y[idx] : x[idx];
}

// physical position of logical row i on a quad-interleaved row axis (psf_b200.h): quads u = i>>2 are dealt
// round-robin to Q blocks of n/Q rows
__device__ __forceinline__ int64_t row_pos(int64_t i, int64_t n, int Q) {
  if (Q <= 1) return i;
  const int64_t u = i >> 2;
  return (u % Q) * (n / Q) + 4 * (u / Q) + (i & 3);
}

template <typename T>
__global__ void __launch_bounds__(kThreads) relayout_kernel(const T* __restrict__ src, psf_wstrides ss, int qs,
                                                            T* __restrict__ dst, psf_wstrides ds, int qd, int n_W,
                                                            int64_t B, int64_t n, int L) {
  // 32 rows x L taps tile through shared memory so that both sides see their contiguous axis.
  extern __shared__ unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);  // [32][L+1]
  const int64_t row_tiles = (n + 31) / 32;
  const int64_t total_tiles = (int64_t)n_W * B * row_tiles;
  for (int64_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    const int64_t rt = t % row_tiles;
    const int64_t fb = t / row_tiles;
    const int64_t b = fb % B;
    const int64_t f = fb / B;
    const int64_t r0 = rt * 32;
    const T* s = src + f * ss.factor + b * ss.sample;
    T* d = dst + f * ds.factor + b * ds.sample;
    const int elems = 32 * L;
    // read with the source's fastest axis on consecutive threads
    for (int e = threadIdx.x; e < elems; e += blockDim.x) {
      int r, k;
      if (ss.tap <= ss.row) { k = e % L; r = e / L; } else { r = e % 32; k = e / 32; }
      if (r0 + r < n) tile[r * (L + 1) + k] = s[row_pos(r0 + r, n, qs) * ss.row + k * ss.tap];
    }
    __syncthreads();
    for (int e = threadIdx.x; e < elems; e += blockDim.x) {
      int r, k;
      if (ds.tap <= ds.row) { k = e % L; r = e / L; } else { r = e % 32; k = e / 32; }
      if (r0 + r < n) d[row_pos(r0 + r, n, qd) * ds.row + k * ds.tap] = tile[r * (L + 1) + k];
    }
    __syncthreads();
  }
}

}  // namespace

template <typename T>
int psf_generic_factor_fwd(const T* Wm, int64_t s_sample, int64_t s_row, int64_t s_tap, const T* V, const T* res, T* Z,
                           const PsfOffsets& offs, int64_t B, int64_t n, int L, int64_t D, cudaStream_t st) {
  if (B * n * D == 0) return PSF_OK;
  factor_fwd_kernel<T><<<grid_for(B * n * D, kThreads), kThreads, 0, st>>>(Wm, s_sample, s_row, s_tap, V, res, Z, offs,
                                                                            B, n, L, D);
  return (int)cudaGetLastError();
}

template <typename T>
int psf_generic_factor_dV(const T* Wm, int64_t s_sample, int64_t s_row, int64_t s_tap, const T* dZ, T* dV,
                          const PsfOffsets& offs, int64_t B, int64_t n, int L, int64_t D, cudaStream_t st) {
  if (B * n * D == 0) return PSF_OK;
  factor_dV_kernel<T><<<grid_for(B * n * D, kThreads), kThreads, 0, st>>>(Wm, s_sample, s_row, s_tap, dZ, dV, offs, B,
                                                                           n, L, D);
  return (int)cudaGetLastError();
}

template <typename T>
int psf_generic_factor_dW(const T* dZ, const T* V, T* dWm, int64_t s_sample, int64_t s_row, int64_t s_tap,
                          const PsfOffsets& offs, int64_t B, int64_t n, int L, int64_t D, cudaStream_t st) {
  if (B * n * L == 0) return PSF_OK;
  factor_dW_kernel<T><<<grid_for(B * n * L, kThreads), kThreads, 0, st>>>(dZ, V, dWm, s_sample, s_row, s_tap, offs, B,
                                                                           n, L, D);
  return (int)cudaGetLastError();
}

template <typename T>
int psf_generic_axpy(T* acc, const T* x, const T* y, int64_t count, cudaStream_t st) {
  if (count == 0) return PSF_OK;
  add_kernel<T><<<grid_for(count, kThreads), kThreads, 0, st>>>(acc, x, y, count);
  return (int)cudaGetLastError();
}

// Workspace (elements of T, S = B*n*C):
//   save_states=1: [0, n_W*S) = V_0 .. V_{n_W-1};  [n_W*S, (n_W+3)*S) = scratch for the backward pass
//   save_states=0: [0, 2*S) ping-pong (the allocation is 3*S, the fused path needs a third slot)
template <typename T>
int psf_generic_chain_fwd(const PsfChainArgs& a, const T* W, const T* V0, T* out, T* ws, int save_states,
                          cudaStream_t st) {
  const int64_t S = a.B * a.n * a.C;
  if (S == 0) return PSF_OK;
  const T* res = a.residual ? V0 : nullptr;
  if (a.n_W == 0) return (int)cudaMemcpyAsync(out, V0, S * sizeof(T), cudaMemcpyDeviceToDevice, st);
  if (save_states) PSF_CUDA_TRY(cudaMemcpyAsync(ws, V0, S * sizeof(T), cudaMemcpyDeviceToDevice, st));
  const T* cur = V0;
  for (int m = 1; m <= a.n_W; ++m) {
    T* dst = (m == a.n_W) ? out : (save_states ? ws + (int64_t)m * S : ws + (int64_t)(m & 1) * S);
    int rc = psf_generic_factor_fwd<T>(W + (int64_t)(m - 1) * a.ws.factor, a.ws.sample, a.ws.row, a.ws.tap, cur, res,
                                       dst, a.offs, a.B, a.n, a.L, a.C, st);
    if (rc) return rc;
    cur = dst;
  }
  return PSF_OK;
}

template <typename T>
int psf_generic_chain_bwd(const PsfChainArgs& a, const T* W, const T* dOut, T* dW, T* dV0, T* ws, cudaStream_t st) {
  const int64_t S = a.B * a.n * a.C;
  if (S == 0) return PSF_OK;
  if (a.n_W == 0) return (int)cudaMemcpyAsync(dV0, dOut, S * sizeof(T), cudaMemcpyDeviceToDevice, st);
  T* ga = ws + (int64_t)a.n_W * S;
  T* gb = ga + S;
  T* racc = gb + S;  // sum of G_m, m = n_W..1, when residual
  const T* gcur = dOut;
  for (int m = a.n_W; m >= 1; --m) {
    const T* Wm = W + (int64_t)(m - 1) * a.ws.factor;
    T* dWm = dW + (int64_t)(m - 1) * a.ws.factor;
    const T* Vprev = ws + (int64_t)(m - 1) * S;
    int rc = psf_generic_factor_dW<T>(gcur, Vprev, dWm, a.ws.sample, a.ws.row, a.ws.tap, a.offs, a.B, a.n, a.L, a.C, st);
    if (rc) return rc;
    if (a.residual) {
      rc = psf_generic_axpy<T>(racc, gcur, m == a.n_W ? nullptr : racc, S, st);
      if (rc) return rc;
    }
    T* gnext = (m == 1 && !a.residual) ? dV0 : ((gcur == ga) ? gb : ga);
    rc = psf_generic_factor_dV<T>(Wm, a.ws.sample, a.ws.row, a.ws.tap, gcur, gnext, a.offs, a.B, a.n, a.L, a.C, st);
    if (rc) return rc;
    gcur = gnext;
  }
  if (a.residual) return psf_generic_axpy<T>(dV0, gcur, racc, S, st);
  return PSF_OK;
}

int psf_generic_relayout(psf_dtype dtype, const void* src, psf_wstrides ss, int qs, void* dst, psf_wstrides ds, int qd,
                         int n_W, int64_t B, int64_t n, int L, cudaStream_t st) {
  if ((int64_t)n_W * B * n * L == 0) return PSF_OK;
  const int64_t tiles = (int64_t)n_W * B * ((n + 31) / 32);
  const int grid = (int)(tiles > 148 * 16 ? 148 * 16 : tiles);
  if (dtype == PSF_F32) {
    relayout_kernel<float><<<grid, 256, 32 * (L + 1) * sizeof(float), st>>>((const float*)src, ss, qs, (float*)dst, ds, qd,
                                                                            n_W, B, n, L);
  } else {
    relayout_kernel<double><<<grid, 256, 32 * (L + 1) * sizeof(double), st>>>((const double*)src, ss, qs, (double*)dst,
                                                                              ds, qd, n_W, B, n, L);
  }
  return (int)cudaGetLastError();
}

#define PSF_INSTANTIATE(T)                                                                                              \
  template int psf_generic_factor_fwd<T>(const T*, int64_t, int64_t, int64_t, const T*, const T*, T*, const PsfOffsets&, \
                                         int64_t, int64_t, int, int64_t, cudaStream_t);                                 \
  template int psf_generic_factor_dV<T>(const T*, int64_t, int64_t, int64_t, const T*, T*, const PsfOffsets&, int64_t,  \
                                        int64_t, int, int64_t, cudaStream_t);                                           \
  template int psf_generic_factor_dW<T>(const T*, const T*, T*, int64_t, int64_t, int64_t, const PsfOffsets&, int64_t,  \
                                        int64_t, int, int64_t, cudaStream_t);                                           \
  template int psf_generic_axpy<T>(T*, const T*, const T*, int64_t, cudaStream_t);                                      \
  template int psf_generic_chain_fwd<T>(const PsfChainArgs&, const T*, const T*, T*, T*, int, cudaStream_t);            \
  template int psf_generic_chain_bwd<T>(const PsfChainArgs&, const T*, const T*, T*, T*, T*, cudaStream_t);
PSF_INSTANTIATE(float)
PSF_INSTANTIATE(double)
