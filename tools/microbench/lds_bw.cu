// Microbenchmark: sustained shared-memory operand bandwidth for the chain's access pattern.
// Each thread reads float4 operands (conflict-free: lane -> consecutive 16 B) from 8 channel
// planes at a runtime row offset and feeds packed FMAs, like the inner loop of the fused chain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_bw lds_bw.cu && ./lds_bw
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void fma4(float4& acc, const float4& a, const float4& b) {
  unsigned long long* pc = reinterpret_cast<unsigned long long*>(&acc);
  const unsigned long long* pa = reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long* pb = reinterpret_cast<const unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(pc[0]) : "l"(pa[0]), "l"(pb[0]));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(pc[1]) : "l"(pa[1]), "l"(pb[1]));
}

template <int MODE>  // 0: LDS.128 + FFMA2 (8 ch), 1: LDS.128 only (xor-accumulate), 2: FFMA2 only
__global__ void __launch_bounds__(256, 2) k(float* out, const int* offs, int taps, int iters) {
  extern __shared__ __align__(16) float sm[];
  constexpr int PL = 2048;
  for (int i = threadIdx.x; i < 8 * PL; i += blockDim.x) sm[i] = 1.0f + i * 1e-7f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 acc[8];
  for (int c = 0; c < 8; ++c) acc[c] = make_float4(0, 0, 0, 0);
  float4 w = make_float4(1.0f, 0.5f, 0.25f, 0.125f);
  for (int it = 0; it < iters; ++it) {
    const float* base = sm + ((it & 1) * 1024 + warp * 128 + lane * 4);
    for (int t = 0; t < taps; ++t) {
      const int o = offs[t];
      const float* vp = base + o;
      if (MODE == 0) {
#pragma unroll
        for (int c = 0; c < 8; ++c) fma4(acc[c], w, *reinterpret_cast<const float4*>(vp + c * PL));
      } else if (MODE == 1) {
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4 v = *reinterpret_cast<const float4*>(vp + c * PL);
          acc[c].x += v.x;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) fma4(acc[c], w, acc[(c + 1) & 7]);
      }
    }
  }
  float s = 0;
  for (int c = 0; c < 8; ++c) s += acc[c].x + acc[c].y + acc[c].z + acc[c].w;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int offs_h[16] = {0, 4, 8, 16, 32, 64, 128, 256, 512, 12, 20, 36, 68, 132, 260, 516};
  int* offs;
  float* out;
  cudaMalloc(&offs, sizeof(offs_h));
  cudaMemcpy(offs, offs_h, sizeof(offs_h), cudaMemcpyHostToDevice);
  cudaMalloc(&out, 296 * 256 * 4);
  const int smem = 8 * 2048 * 4 + 48 * 1024;  // same footprint as the fused kernel (2 CTAs / SM)
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int mode = 0; mode < 3; ++mode) {
    auto fn = mode == 0 ? k<0> : (mode == 1 ? k<1> : k<2>);
    cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000, taps = 16;
    fn<<<296, 256, smem>>>(out, offs, taps, 10);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    fn<<<296, 256, smem>>>(out, offs, taps, iters);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    // per SM: 2 CTAs x 8 warps x iters x taps x 8 LDS.128 (512 B each)
    const double bytes_sm = 2.0 * 8 * iters * taps * 8 * 512;
    const double fma_sm = 2.0 * 8 * iters * taps * 8 * 4 * 32;
    printf("mode %d: %.3f ms  smem %.1f B/ns/SM (%.1f B/clk @1.9GHz)  fma %.1f /ns/SM  err=%s\n", mode, ms,
           bytes_sm / (ms * 1e6), bytes_sm / (ms * 1e6) / 1.9, fma_sm / (ms * 1e6), cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
