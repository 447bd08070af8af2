// Head of the FLATTEN pooling:  logits[b][c] = bias[c] + sum_k V[b][k] W[c][k]   (self.final = nn.Linear(n_vec *
// n_channels_V, n_class), SyntheticExperiments/psf.py:60,80; LRA/psf.py:13-22,109).  K = n C is 131072 at the headline
// shape and n_class is 1..16, so both directions are streaming passes over V bounded by HBM (69 MB at B = 132).  Measured
// inside the training step: 13 us forward, 20 us backward (the library's gemv route: 31 + 85 us in isolation).
// Sums are taken in a fixed order (per-thread, warp tree, chunk order), no atomics.
#include "psf_internal.cuh"

namespace {

constexpr int kHeadChunk = 4096;      // K elements per forward CTA
constexpr int kHeadThreads = 128;
constexpr int kHeadSlices = 4;        // backward: the batch is cut into this many slices with partial dW each

__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// grid (chunks, B): partial[b][chunk][c] = sum over the chunk's k of V[b][k] W[c][k]
template <int NC>
__global__ void __launch_bounds__(kHeadThreads) psf_head_fwd_kernel(const float* __restrict__ V, const float* __restrict__ W,
                                                                     float* __restrict__ part, long long K, int nchunks) {
  const int chunk = blockIdx.x;
  const long long b = blockIdx.y, k0 = (long long)chunk * kHeadChunk;
  float acc[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) acc[c] = 0.f;
  const float* vrow = V + b * K;
#pragma unroll 2
  for (int i = threadIdx.x * 4; i < kHeadChunk; i += kHeadThreads * 4) {
    const long long k = k0 + i;
    if (k < K) {
      const float4 v = ld_stream4(vrow + k);
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(W + (long long)c * K + k));
        acc[c] = fmaf(v.x, w.x, fmaf(v.y, w.y, fmaf(v.z, w.z, fmaf(v.w, w.w, acc[c]))));
      }
    }
  }
  __shared__ float red[kHeadThreads / 32][NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[c] += __shfl_down_sync(0xffffffffu, acc[c], o);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int c = 0; c < NC; ++c) red[threadIdx.x >> 5][c] = acc[c];
  }
  __syncthreads();
  if (threadIdx.x < NC) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kHeadThreads / 32; ++w) s += red[w][threadIdx.x];
    part[(b * nchunks + chunk) * NC + threadIdx.x] = s;
  }
}

// one thread per (b, c): chunks in order, then the bias
__global__ void psf_head_fwd_reduce_kernel(const float* __restrict__ part, const float* __restrict__ bias, float* __restrict__ out,
                                           long long total, int nchunks, int NC) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // b * NC + c
  if (i >= total) return;
  const long long b = i / NC;
  const int c = (int)(i - b * NC);
  float s = 0.f;
  for (int ch = 0; ch < nchunks; ++ch) s += part[(b * nchunks + ch) * NC + c];
  out[i] = s + (bias ? bias[c] : 0.f);
}

// grid (K / 4 / threads, slices): thread owns four consecutive k, walks the rows of its batch slice:
//   dV[b][k] = sum_c g[b][c] W[c][k]      dWpart[slice][c][k] = sum_{b in slice} g[b][c] V[b][k]
template <int NC>
__global__ void __launch_bounds__(kHeadThreads) psf_head_bwd_kernel(const float* __restrict__ V, const float* __restrict__ W,
                                                                     const float* __restrict__ g, float* __restrict__ dV,
                                                                     float* __restrict__ dWpart, long long K, long long B,
                                                                     long long rows_per_slice) {
  const long long k = ((long long)blockIdx.x * kHeadThreads + threadIdx.x) * 4;
  const long long b0 = (long long)blockIdx.y * rows_per_slice;
  const long long b1 = b0 + rows_per_slice < B ? b0 + rows_per_slice : B;
  if (k >= K) return;
  float4 w[NC], dw[NC];
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    w[c] = __ldg(reinterpret_cast<const float4*>(W + (long long)c * K + k));
    dw[c] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll 4
  for (long long b = b0; b < b1; ++b) {
    const float4 v = ld_stream4(V + b * K + k);
    float4 dv = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      const float gc = __ldg(g + b * NC + c);
      dv.x = fmaf(gc, w[c].x, dv.x), dv.y = fmaf(gc, w[c].y, dv.y), dv.z = fmaf(gc, w[c].z, dv.z), dv.w = fmaf(gc, w[c].w, dv.w);
      dw[c].x = fmaf(gc, v.x, dw[c].x), dw[c].y = fmaf(gc, v.y, dw[c].y), dw[c].z = fmaf(gc, v.z, dw[c].z),
      dw[c].w = fmaf(gc, v.w, dw[c].w);
    }
    st_stream4(dV + b * K + k, dv);
  }
#pragma unroll
  for (int c = 0; c < NC; ++c)
    *reinterpret_cast<float4*>(dWpart + ((long long)blockIdx.y * NC + c) * K + k) = dw[c];
}

// dW[c][k] = sum over the slices, in order
__global__ void psf_head_bwd_reduce_kernel(const float* __restrict__ dWpart, float* __restrict__ dW, long long count, int slices) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i >= count) return;
  float4 s = *reinterpret_cast<const float4*>(dWpart + i);
  for (int sl = 1; sl < slices; ++sl) {
    const float4 t = *reinterpret_cast<const float4*>(dWpart + (long long)sl * count + i);
    s.x += t.x, s.y += t.y, s.z += t.z, s.w += t.w;
  }
  *reinterpret_cast<float4*>(dW + i) = s;
}

template <int NC>
int head_fwd(const float* V, const float* W, float* part, long long B, long long K, int nchunks, cudaStream_t st) {
  psf_head_fwd_kernel<NC><<<dim3((unsigned)nchunks, (unsigned)B), kHeadThreads, 0, st>>>(V, W, part, K, nchunks);
  return (int)cudaGetLastError();
}
template <int NC>
int head_bwd(const float* V, const float* W, const float* g, float* dV, float* dWpart, long long B, long long K, int slices,
             cudaStream_t st) {
  const long long rps = (B + slices - 1) / slices;
  const long long threads = K / 4;
  psf_head_bwd_kernel<NC><<<dim3((unsigned)((threads + kHeadThreads - 1) / kHeadThreads), (unsigned)slices), kHeadThreads, 0, st>>>(
      V, W, g, dV, dWpart, K, B, rps);
  return (int)cudaGetLastError();
}

#define PSF_HEAD_DISPATCH(NC, CALL)                                                                   \
  switch (NC) {                                                                                       \
    case 1: return CALL(1); case 2: return CALL(2); case 3: return CALL(3); case 4: return CALL(4);   \
    case 5: return CALL(5); case 6: return CALL(6); case 7: return CALL(7); case 8: return CALL(8);   \
    case 9: return CALL(9); case 10: return CALL(10); case 11: return CALL(11); case 12: return CALL(12); \
    case 13: return CALL(13); case 14: return CALL(14); case 15: return CALL(15); case 16: return CALL(16); \
    default: return PSF_ERR_UNSUPPORTED;                                                              \
  }

int dispatch_fwd(int NC, const float* V, const float* W, float* part, long long B, long long K, int nchunks, cudaStream_t st) {
#define CALL_F(N) head_fwd<N>(V, W, part, B, K, nchunks, st)
  PSF_HEAD_DISPATCH(NC, CALL_F)
#undef CALL_F
}
int dispatch_bwd(int NC, const float* V, const float* W, const float* g, float* dV, float* dWpart, long long B, long long K,
                 int slices, cudaStream_t st) {
#define CALL_B(N) head_bwd<N>(V, W, g, dV, dWpart, B, K, slices, st)
  PSF_HEAD_DISPATCH(NC, CALL_B)
#undef CALL_B
}

bool head_shape_ok(const void* a, const void* b, long long B, long long K, int n_class) {
  return B > 0 && B <= 65535 && K > 0 && (K & 3) == 0 && n_class >= 1 && n_class <= PSF_HEAD_MAX_CLASSES &&
         !(reinterpret_cast<uintptr_t>(a) & 15) && !(reinterpret_cast<uintptr_t>(b) & 15);
}
int head_slices(long long B) { return B >= 2 * kHeadSlices ? kHeadSlices : 1; }

}  // namespace

extern "C" size_t psf_head_workspace_bytes(int64_t B, int64_t K, int n_class, int backward) {
  if (B <= 0 || K <= 0 || n_class < 1) return 0;
  if (!backward) return psf_round256((size_t)B * (size_t)((K + kHeadChunk - 1) / kHeadChunk) * n_class * sizeof(float));
  return psf_round256((size_t)head_slices(B) * (size_t)n_class * (size_t)K * sizeof(float));
}

extern "C" int psf_head_fwd(const float* V, const float* W, const float* bias, float* out, int64_t B, int64_t K, int n_class,
                            void* workspace, size_t workspace_bytes, psf_stream_t stream) {
  if (!V || !W || !out) return PSF_ERR_ARG;
  if (!head_shape_ok(V, W, B, K, n_class)) return PSF_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < psf_head_workspace_bytes(B, K, n_class, 0)) return PSF_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int nchunks = (int)((K + kHeadChunk - 1) / kHeadChunk);
  float* part = static_cast<float*>(workspace);
  const int rc = dispatch_fwd(n_class, V, W, part, B, K, nchunks, st);
  if (rc != 0) return rc;
  const long long total = (long long)B * n_class;
  psf_head_fwd_reduce_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(part, bias, out, total, nchunks, n_class);
  return (int)cudaGetLastError();
}

extern "C" int psf_head_bwd(const float* V, const float* W, const float* g, float* dV, float* dW, int64_t B, int64_t K,
                            int n_class, void* workspace, size_t workspace_bytes, psf_stream_t stream) {
  if (!V || !W || !g || !dV || !dW) return PSF_ERR_ARG;
  if (!head_shape_ok(V, W, B, K, n_class) || (reinterpret_cast<uintptr_t>(dV) & 15) || (reinterpret_cast<uintptr_t>(dW) & 15))
    return PSF_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < psf_head_workspace_bytes(B, K, n_class, 1)) return PSF_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int slices = head_slices(B);
  float* part = slices == 1 ? dW : static_cast<float*>(workspace);
  const int rc = dispatch_bwd(n_class, V, W, g, dV, part, B, K, slices, st);
  if (rc != 0 || slices == 1) return rc;
  const long long count = (long long)n_class * K;
  psf_head_bwd_reduce_kernel<<<(unsigned)((count / 4 + 127) / 128), 128, 0, st>>>(part, dW, count, slices);
  return (int)cudaGetLastError();
}
