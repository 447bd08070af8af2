// Variants of the chain inner loop: which code shape sustains the shared-memory operand bandwidth?
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void fma4(float4& acc, const float4& a, const float4& b) {
  unsigned long long* pc = reinterpret_cast<unsigned long long*>(&acc);
  const unsigned long long* pa = reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long* pb = reinterpret_cast<const unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(pc[0]) : "l"(pa[0]), "l"(pb[0]));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(pc[1]) : "l"(pa[1]), "l"(pb[1]));
}
__device__ __forceinline__ void fma4s(float4& acc, const float4& a, const float4& b) {
  acc.x = fmaf(a.x, b.x, acc.x); acc.y = fmaf(a.y, b.y, acc.y);
  acc.z = fmaf(a.z, b.z, acc.z); acc.w = fmaf(a.w, b.w, acc.w);
}
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

constexpr int PL = 2048;

// MODE 0: baseline (8 loads then 8 fma4, FFMA2)        MODE 1: scalar FFMA
// MODE 2: explicit double buffering across taps (FFMA2) MODE 3: double buffering, scalar FFMA
// MODE 4: 4 rows x 4 channels per thread (half the registers), FFMA2
template <int MODE, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k(float* out, const int* offs, int taps, int iters) {
  extern __shared__ __align__(16) float sm[];
  for (int i = threadIdx.x; i < 8 * PL; i += blockDim.x) sm[i] = 1.0f + i * 1e-7f;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float4 acc[8];
  for (int c = 0; c < 8; ++c) acc[c] = make_float4(0, 0, 0, 0);
  float4 w = make_float4(1.0f, 0.5f, 0.25f, 0.125f);
  const int nw = THREADS / 32;
  for (int it = 0; it < iters; ++it) {
    const float* base = sm + (((it & 1) * 1024 + warp * 128) & (PL / 2 - 1)) + lane * 4;
    if (MODE == 0 || MODE == 1) {
      for (int t = 0; t < taps; ++t) {
        const float* vp = base + offs[t];
        float4 v[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) v[c] = lds4(vp + c * PL);
#pragma unroll
        for (int c = 0; c < 8; ++c) { if (MODE == 0) fma4(acc[c], w, v[c]); else fma4s(acc[c], w, v[c]); }
      }
    } else if (MODE == 2 || MODE == 3) {
      float4 va[8], vb[8];
      {
        const float* vp = base + offs[0];
#pragma unroll
        for (int c = 0; c < 8; ++c) va[c] = lds4(vp + c * PL);
      }
      for (int t = 0; t < taps; t += 2) {
        const float* vp1 = base + offs[t + 1];
#pragma unroll
        for (int c = 0; c < 8; ++c) vb[c] = lds4(vp1 + c * PL);
#pragma unroll
        for (int c = 0; c < 8; ++c) { if (MODE == 2) fma4(acc[c], w, va[c]); else fma4s(acc[c], w, va[c]); }
        const float* vp2 = base + offs[(t + 2) & 15];
#pragma unroll
        for (int c = 0; c < 8; ++c) va[c] = lds4(vp2 + c * PL);
#pragma unroll
        for (int c = 0; c < 8; ++c) { if (MODE == 2) fma4(acc[c], w, vb[c]); else fma4s(acc[c], w, vb[c]); }
      }
    } else if (MODE == 4) {
      // half the channels per thread: thread = (row group, channel half); two halves of a warp pair
      const int half = warp & 1;
      const float* b4 = sm + (((it & 1) * 1024 + (warp >> 1) * 128) & (PL / 2 - 1)) + lane * 4 + half * 4 * PL;
      for (int t = 0; t < taps; ++t) {
        const float* vp = b4 + offs[t];
        float4 v[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) v[c] = lds4(vp + c * PL);
#pragma unroll
        for (int c = 0; c < 4; ++c) fma4(acc[c], w, v[c]);
      }
    }
  }
  float s = 0;
  for (int c = 0; c < 8; ++c) s += acc[c].x + acc[c].y + acc[c].z + acc[c].w;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + nw;
}

template <int MODE, int THREADS, int MINB>
void run(const char* name, float* out, const int* offs, int extra_smem_kb) {
  auto fn = k<MODE, THREADS, MINB>;
  const int smem = 8 * PL * 4 + extra_smem_kb * 1024;
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, THREADS, smem);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, fn);
  const int iters = 2000, taps = 16, grid = 148 * nb;
  fn<<<grid, THREADS, smem>>>(out, offs, taps, 10);
  cudaDeviceSynchronize();
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  fn<<<grid, THREADS, smem>>>(out, offs, taps, iters);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double loads = (MODE == 4) ? 4 : 8;
  const double bytes_sm = (double)nb * (THREADS / 32) * iters * taps * loads * 512;
  printf("%-34s blocks/SM %d regs %3d: %.3f ms  smem %.1f B/clk/SM @1.965GHz  (%s)\n", name, nb, fa.numRegs, ms,
         bytes_sm / (ms * 1e6) / 1.965, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  int offs_h[16] = {0, 4, 8, 16, 32, 64, 128, 256, 512, 12, 20, 36, 68, 132, 260, 516};
  int* offs; float* out;
  cudaMalloc(&offs, sizeof(offs_h));
  cudaMemcpy(offs, offs_h, sizeof(offs_h), cudaMemcpyHostToDevice);
  cudaMalloc(&out, 148 * 8 * 1024 * 4);
  run<0, 256, 2>("base FFMA2 256x2", out, offs, 48);
  run<1, 256, 2>("base FFMA  256x2", out, offs, 48);
  run<2, 256, 2>("dbuf FFMA2 256x2", out, offs, 48);
  run<3, 256, 2>("dbuf FFMA  256x2", out, offs, 48);
  run<4, 256, 2>("4ch  FFMA2 256x2", out, offs, 48);
  run<0, 512, 1>("base FFMA2 512x1", out, offs, 128);
  run<4, 512, 2>("4ch  FFMA2 512x2 (32 warps)", out, offs, 48);
  run<0, 384, 2>("base FFMA2 384x2 (24 warps)", out, offs, 48);
  run<1, 384, 2>("base FFMA  384x2 (24 warps)", out, offs, 48);
  run<0, 256, 3>("base FFMA2 256x3 (24 warps)", out, offs, 8);
  return 0;
}
