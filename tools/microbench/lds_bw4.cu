// Clean comparison of consumer layouts with realistic per-tap control flow (offset table in smem,
// alignment + locality branch, W operand from smem, completed-mbarrier wait per 4 taps):
//   A: 4 rows x 8 channels per thread, 16 warps/SM (current kernels)
//   B: 4 rows x 4 channels per thread, lanes 0-15 / 16-31 = channel halves, 32 warps/SM, 64 regs
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
struct Offs { int o[32]; };
constexpr int PL = 2048;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile("{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fma4(float4& acc, const float4& a, const float4& b) {
  unsigned long long* pc = reinterpret_cast<unsigned long long*>(&acc);
  const unsigned long long* pa = reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long* pb = reinterpret_cast<const unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(pc[0]) : "l"(pa[0]), "l"(pb[0]));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(pc[1]) : "l"(pa[1]), "l"(pb[1]));
}
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

template <int CH, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k(float* out, Offs offs, int taps, int iters, int blk_end) {
  extern __shared__ __align__(16) float sm[];
  float* ring = sm + 8 * PL;
  __shared__ __align__(8) unsigned long long bar;
  __shared__ __align__(16) int soffs[32];
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 32; ++i) soffs[i] = offs.o[i];
  }
  for (int i = threadIdx.x; i < 8 * PL + 4096; i += blockDim.x) sm[i] = 1.0f + i * 1e-7f;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)));
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rg = (CH == 8) ? lane : (lane & 15);          // row group inside the warp tile
  const int chalf = (CH == 8) ? 0 : (lane >> 4);          // which 4 channels
  const int tile = (CH == 8) ? 128 : 64;
  const int j = (warp * tile) % 1024 + rg * 4;
  float4 acc[CH];
  for (int c = 0; c < CH; ++c) acc[c] = make_float4(0, 0, 0, 0);
  const float* vown = sm + j + chalf * 4 * PL;
  const float* wst = ring + (warp * tile) % 1024 + rg * 4;
  for (int it = 0; it < iters; ++it) {
    for (int tg = 0; tg < taps / 4; ++tg) {
      mbar_wait(smem_u32(&bar), 0);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int o = soffs[tg * 4 + t];
        const int src = j + o;
        if ((o & 3) == 0 && src + 4 <= blk_end) {
          const float4 w = lds4(wst + t * 1024);
          float4 v[CH];
#pragma unroll
          for (int c = 0; c < CH; ++c) v[c] = lds4(vown + o + c * PL);
#pragma unroll
          for (int c = 0; c < CH; ++c) fma4(acc[c], w, v[c]);
        } else {
          for (int c = 0; c < CH; ++c) acc[c].x += out[c * 1000 + src];
        }
      }
      __syncwarp();
    }
  }
  float s = 0;
  for (int c = 0; c < CH; ++c) s += acc[c].x + acc[c].y + acc[c].z + acc[c].w;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int CH, int THREADS, int MINB>
void run(const char* name, float* out, const Offs& offs) {
  auto fn = k<CH, THREADS, MINB>;
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
  // useful operand bytes: per warp-tap CH V quads x 32 lanes x 16 B (the W quad is counted once per 128 rows)
  const double rows_per_sm = (double)nb * (THREADS / 32) * ((CH == 8) ? 128 : 64);
  const double v_bytes = rows_per_sm / 4 * 8 * 16.0 * (CH == 8 ? 1 : 0.5) * iters * taps;   // V operand bytes
  const double fma = rows_per_sm * (CH == 8 ? 8 : 4) * iters * taps;
  printf("%-44s blocks/SM %d regs %3d: %.3f ms  %.1f FMA/clk/SM (%s)\n", name, nb, fa.numRegs, ms,
         fma / (ms * 1e6) / 1.965, cudaGetErrorString(cudaGetLastError()));
  (void)v_bytes;
}

int main() {
  Offs offs;
  int o[16] = {0, 4, 8, 16, 32, 64, 128, 256, 512, 12, 20, 36, 68, 132, 260, 516};
  for (int i = 0; i < 32; ++i) offs.o[i] = o[i & 15];
  float* out;
  cudaMalloc(&out, 8 << 20);
  printf("peak for this operand pattern: 32 FMA/clk/SM (one 4 B shared-memory operand per FMA at 128 B/clk)\n");
  run<8, 256, 2>("A: 4 rows x 8 ch, 16 warps", out, offs);
  run<4, 512, 2>("B: 4 rows x 4 ch (lane halves), 32 warps", out, offs);
  run<4, 256, 2>("B: 4 rows x 4 ch (lane halves), 16 warps", out, offs);
  run<8, 512, 1>("A: 4 rows x 8 ch, 16 warps in one CTA", out, offs);
  return 0;
}
