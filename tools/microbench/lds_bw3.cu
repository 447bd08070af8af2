// Inner-loop shape study with realistic control flow: per-tap offset from kernel params, alignment and
// locality branches, W operand from shared memory, a (completed) mbarrier wait every 4 taps.
//   R = rows per thread (4: LDS.128, 16 warps/SM;  2: LDS.64, 32 warps/SM)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
struct Offs { int o[32]; };
constexpr int PL = 2048;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fma_f2(float2& acc, const float2& a, const float2& b) {
  unsigned long long* pc = reinterpret_cast<unsigned long long*>(&acc);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(*pc) : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
}
__device__ __forceinline__ void fma_f4(float4& acc, const float4& a, const float4& b) {
  unsigned long long* pc = reinterpret_cast<unsigned long long*>(&acc);
  const unsigned long long* pa = reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long* pb = reinterpret_cast<const unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(pc[0]) : "l"(pa[0]), "l"(pb[0]));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(pc[1]) : "l"(pa[1]), "l"(pb[1]));
}

template <int R, int THREADS, int MINB, bool WAITS>
__global__ void __launch_bounds__(THREADS, MINB) k(float* out, Offs offs, int taps, int iters, int blk_end) {
  extern __shared__ __align__(16) float sm[];
  float* ring = sm + 8 * PL;   // 4 taps x 1024 rows
  __shared__ __align__(8) unsigned long long bar;
  __shared__ __align__(16) int soffs[32];
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 32; ++i) soffs[i] = offs.o[i];
  }
  for (int i = threadIdx.x; i < 8 * PL + 4096; i += blockDim.x) sm[i] = 1.0f + i * 1e-7f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)));  // phase 0 complete
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = 32 * R;
  const int j = (warp * tile) % 1024 + lane * R;
  float acc[8][R];
  for (int c = 0; c < 8; ++c) for (int r = 0; r < R; ++r) acc[c][r] = 0.f;
  const float* vown = sm + j;
  const float* wst = ring + (warp * tile) % 1024 + lane * R;
  for (int it = 0; it < iters; ++it) {
    for (int tg = 0; tg < taps / 4; ++tg) {
      if (WAITS) mbar_wait(smem_u32(&bar), 0);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int o = soffs[tg * 4 + t];
        const int src = j + o;
        if ((o & (R - 1)) == 0 && src + R <= blk_end) {
          if (R == 4) {
            const float4 w = *reinterpret_cast<const float4*>(wst + t * 1024);
#pragma unroll
            for (int c = 0; c < 8; ++c) fma_f4(*reinterpret_cast<float4*>(acc[c]), w, *reinterpret_cast<const float4*>(vown + o + c * PL));
          } else {
            const float2 w = *reinterpret_cast<const float2*>(wst + t * 1024);
#pragma unroll
            for (int c = 0; c < 8; ++c) fma_f2(*reinterpret_cast<float2*>(acc[c]), w, *reinterpret_cast<const float2*>(vown + o + c * PL));
          }
        } else {
          for (int c = 0; c < 8; ++c) acc[c][0] += out[c * 1000 + src];  // "remote" path, never taken here
        }
      }
      __syncwarp();
    }
  }
  float s = 0;
  for (int c = 0; c < 8; ++c) for (int r = 0; r < R; ++r) s += acc[c][r];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int R, int THREADS, int MINB, bool WAITS>
void run(const char* name, float* out, const Offs& offs) {
  auto fn = k<R, THREADS, MINB, WAITS>;
  const int smem = 8 * PL * 4 + 16 * 1024 + 24 * 1024;
  cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int nb = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, THREADS, smem);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, fn);
  const int iters = 2000, taps = 16, grid = 148 * nb;
  fn<<<grid, THREADS, smem>>>(out, offs, taps, 10, 1 << 20);
  cudaDeviceSynchronize();
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a);
  fn<<<grid, THREADS, smem>>>(out, offs, taps, iters, 1 << 20);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  const double bytes_sm = (double)nb * (THREADS / 32) * iters * taps * 9 * (32 * R * 4);   // 8 V + 1 W loads
  printf("%-40s blocks/SM %d regs %3d: %.3f ms  smem %.1f B/clk/SM @1.965GHz (%s)\n", name, nb, fa.numRegs, ms,
         bytes_sm / (ms * 1e6) / 1.965, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  Offs offs;
  int o[16] = {0, 4, 8, 16, 32, 64, 128, 256, 512, 12, 20, 36, 68, 132, 260, 516};
  for (int i = 0; i < 32; ++i) offs.o[i] = o[i & 15];
  float* out;
  cudaMalloc(&out, 8 << 20);
  run<4, 256, 2, false>("R=4 16 warps, branches, no waits", out, offs);
  run<4, 256, 2, true>("R=4 16 warps, branches + mbar waits", out, offs);
  run<2, 512, 2, false>("R=2 32 warps, branches, no waits", out, offs);
  run<2, 512, 2, true>("R=2 32 warps, branches + mbar waits", out, offs);
  run<2, 256, 2, true>("R=2 16 warps, branches + mbar waits", out, offs);
  run<4, 512, 1, true>("R=4 16 warps (1 CTA), waits", out, offs);
  return 0;
}
