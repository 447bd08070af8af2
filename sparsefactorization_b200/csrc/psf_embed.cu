// init_linear of the Adding model: y[row][e] = b[e] + sum_i x[row][i] W[e][i] with a 2-wide input over B n rows
// (add_init_linear_layer, SyntheticExperiments/psf.py:136-141, applied at :155-156).  At the headline shape that is 2.16 M
// rows x 32 outputs: a 277 MB streaming write forward, and on the way back one pass over dy for dW (E x IN) and db (E) -
// the library GEMMs picked for these shapes (SIMT sgemm forward, split-K batched GEMM + two column reductions backward) took
// 0.46 ms of the 8.4 ms step.  Fixed summation order (thread, warp tree, CTA, then CTAs in order), no atomics.
#include "psf_internal.cuh"

namespace {

constexpr int kEmbThreads = 256;
constexpr int kEmbMaxIn = 4;

struct EmbW {                       // W [E][IN] and b [E] by value would not fit kernel parameters for E = 128: read via __ldg
  const float* W;
  const float* b;
};

// thread = (row, four consecutive outputs); G = E / 4 threads per row
template <int IN>
__global__ void __launch_bounds__(kEmbThreads) psf_embed_fwd_kernel(const float* __restrict__ x, EmbW p, float* __restrict__ y,
                                                                     long long rows, int E) {
  const int G = E >> 2;
  const int c4 = threadIdx.x % G, rsub = threadIdx.x / G, rpb = kEmbThreads / G;
  float w[4][IN], bias[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    bias[j] = p.b ? __ldg(p.b + 4 * c4 + j) : 0.f;
#pragma unroll
    for (int i = 0; i < IN; ++i) w[j][i] = __ldg(p.W + (4 * c4 + j) * IN + i);
  }
  if (rsub >= rpb) return;
  for (long long r = (long long)blockIdx.x * rpb + rsub; r < rows; r += (long long)gridDim.x * rpb) {
    float xi[IN];
#pragma unroll
    for (int i = 0; i < IN; ++i) xi[i] = __ldg(x + r * IN + i);
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float s = bias[j];
#pragma unroll
      for (int i = 0; i < IN; ++i) s = fmaf(xi[i], w[j][i], s);
      o[j] = s;
    }
    *reinterpret_cast<float4*>(y + r * E + 4 * c4) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

// per CTA: part[cta][e][IN + 1] = sums over the CTA's rows of dy[row][e] x[row][i] (i < IN) and of dy[row][e]
template <int IN>
__global__ void __launch_bounds__(kEmbThreads) psf_embed_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                     float* __restrict__ part, long long rows, int E) {
  extern __shared__ float red[];                  // [rpb][E][IN + 1]
  const int G = E >> 2;
  const int c4 = threadIdx.x % G, rsub = threadIdx.x / G, rpb = kEmbThreads / G;
  float acc[4][IN + 1];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i <= IN; ++i) acc[j][i] = 0.f;
  if (rsub < rpb) {
    for (long long r = (long long)blockIdx.x * rpb + rsub; r < rows; r += (long long)gridDim.x * rpb) {
      float4 g;
      asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(g.x), "=f"(g.y), "=f"(g.z), "=f"(g.w) : "l"(dy + r * E + 4 * c4));
      const float gv[4] = {g.x, g.y, g.z, g.w};
      float xi[IN];
#pragma unroll
      for (int i = 0; i < IN; ++i) xi[i] = __ldg(x + r * IN + i);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int i = 0; i < IN; ++i) acc[j][i] = fmaf(gv[j], xi[i], acc[j][i]);
        acc[j][IN] += gv[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i <= IN; ++i) red[(rsub * E + 4 * c4 + j) * (IN + 1) + i] = acc[j][i];
  }
  __syncthreads();
  const int per = E * (IN + 1);
  for (int t = threadIdx.x; t < per; t += kEmbThreads) {
    double s = 0.0;                                  // the partial sums are combined in double: db and dW are sums of
    for (int q = 0; q < rpb; ++q) s += (double)red[q * per + t];     // millions of terms with cancellation, and a chain of
    part[(long long)blockIdx.x * per + t] = (float)s;                 // ~1200 fp32 additions would cost 5e-4 absolute
  }
}

// dW[e][i], db[e] = sums of the CTA partials: one block per output element, strided partial sums per thread and a fixed
// tree over the block, in double (a single thread walking the ~1200 partials took 117 us: a chain of L2 latencies)
constexpr int kEmbRedThreads = 128;
__global__ void __launch_bounds__(kEmbRedThreads) psf_embed_bwd_reduce_kernel(const float* __restrict__ part, float* __restrict__ dW,
                                                                               float* __restrict__ db, int ctas, int E, int IN) {
  __shared__ double sh[kEmbRedThreads];
  const int t = blockIdx.x, per = E * (IN + 1);
  double s = 0.0;
  for (int c = threadIdx.x; c < ctas; c += kEmbRedThreads) s += (double)part[(long long)c * per + t];
  sh[threadIdx.x] = s;
  __syncthreads();
#pragma unroll
  for (int o = kEmbRedThreads / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int e = t / (IN + 1), i = t - e * (IN + 1);
    if (i < IN) dW[e * IN + i] = (float)sh[0];
    else if (db) db[e] = (float)sh[0];
  }
}

bool embed_ok(long long rows, int IN, int E) {
  return rows > 0 && IN >= 1 && IN <= kEmbMaxIn && E >= 4 && E <= 128 && (E & 3) == 0 && kEmbThreads % (E >> 2) == 0;
}
int embed_grid(long long rows, int E) {
  int dev = psf_current_device(), sms = 148;
  if (dev >= 0) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long rpb = kEmbThreads / (E >> 2);
  const long long need = (rows + rpb - 1) / rpb;
  const long long cap = (long long)sms * 8;
  return (int)(need < cap ? need : cap);
}

}  // namespace

extern "C" size_t psf_init_linear_workspace_bytes(int64_t rows, int in_features, int out_features) {
  if (!embed_ok(rows, in_features, out_features)) return 0;
  return psf_round256((size_t)embed_grid(rows, out_features) * out_features * (in_features + 1) * sizeof(float));
}

extern "C" int psf_init_linear_fwd(const float* x, const float* W, const float* b, float* y, int64_t rows, int in_features,
                                   int out_features, psf_stream_t stream) {
  if (!x || !W || !y) return PSF_ERR_ARG;
  if (!embed_ok(rows, in_features, out_features) || (reinterpret_cast<uintptr_t>(y) & 15)) return PSF_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = embed_grid(rows, out_features);
  const EmbW p = {W, b};
  switch (in_features) {
    case 1: psf_embed_fwd_kernel<1><<<grid, kEmbThreads, 0, st>>>(x, p, y, rows, out_features); break;
    case 2: psf_embed_fwd_kernel<2><<<grid, kEmbThreads, 0, st>>>(x, p, y, rows, out_features); break;
    case 3: psf_embed_fwd_kernel<3><<<grid, kEmbThreads, 0, st>>>(x, p, y, rows, out_features); break;
    default: psf_embed_fwd_kernel<4><<<grid, kEmbThreads, 0, st>>>(x, p, y, rows, out_features); break;
  }
  return (int)cudaGetLastError();
}

extern "C" int psf_init_linear_bwd(const float* x, const float* dy, float* dW, float* db, int64_t rows, int in_features,
                                   int out_features, void* workspace, size_t workspace_bytes, psf_stream_t stream) {
  if (!x || !dy || !dW) return PSF_ERR_ARG;
  if (!embed_ok(rows, in_features, out_features) || (reinterpret_cast<uintptr_t>(dy) & 15)) return PSF_ERR_UNSUPPORTED;
  if (!workspace || workspace_bytes < psf_init_linear_workspace_bytes(rows, in_features, out_features)) return PSF_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = embed_grid(rows, out_features);
  const int rpb = kEmbThreads / (out_features >> 2);
  const size_t smem = (size_t)rpb * out_features * (in_features + 1) * sizeof(float);
  float* part = static_cast<float*>(workspace);
  switch (in_features) {
    case 1: psf_embed_bwd_kernel<1><<<grid, kEmbThreads, smem, st>>>(x, dy, part, rows, out_features); break;
    case 2: psf_embed_bwd_kernel<2><<<grid, kEmbThreads, smem, st>>>(x, dy, part, rows, out_features); break;
    case 3: psf_embed_bwd_kernel<3><<<grid, kEmbThreads, smem, st>>>(x, dy, part, rows, out_features); break;
    default: psf_embed_bwd_kernel<4><<<grid, kEmbThreads, smem, st>>>(x, dy, part, rows, out_features); break;
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;
  const int per = out_features * (in_features + 1);
  psf_embed_bwd_reduce_kernel<<<per, kEmbRedThreads, 0, st>>>(part, dW, db, grid, out_features, in_features);
  return (int)cudaGetLastError();
}
