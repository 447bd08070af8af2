// tcgen05 / TMEM probe for the generator kernels (sm_100a).  Stand-alone:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/umma_probe tools/microbench/umma_probe.cu
//   ./tools/microbench/umma_probe
// Answers, on the real part, the questions the generator design rests on:
//   1. the no-swizzle ("interleave") shared-memory canonical layout, K-major and MN-major, for kind::tf32
//   2. whether tf32 operands are truncated or must be pre-rounded (decides the 3xTF32 split)
//   3. A operand from tensor memory
//   4. cycles per tcgen05.mma for the small N the generators need
// Every wait is bounded (trap after ~1 s) so a wrong descriptor cannot hang the box.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e_ = (x);                                                          \
    if (e_ != cudaSuccess) {                                                       \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
      exit(2);                                                                     \
    }                                                                              \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity) {
  long long t0 = clock64();
  while (!mbar_try(bar, parity)) {
    if (clock64() - t0 > 2000000000LL) {
      printf("mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base offset 0, lbo mode 0, layout type 0 = no swizzle
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "f"(v[0]),
               "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}

struct Params {
  const float* A;   // logical [M=128][K] row-major
  const float* B;   // logical [N][K] row-major
  const float* A2;  // optional second operand pair for the 3xTF32 test (lo parts); null = single product
  const float* B2;
  float* D;         // [128][N]
  long long* cycles;
  int N, K;
  int a_mn, b_mn;   // operand major: 0 = K-major, 1 = MN-major
  int a_tmem;       // 1: A from tensor memory
  int reps;         // timing repetitions of the whole MMA sequence
};

// Tile X[r][c] (R rows, Cc cols) in 8 x 16-byte core matrices: addr = (r/8)*RG + (c/4)*128 + (r%8)*16 + (c%4)*4,
// RG = (Cc/4)*128.  K-major use (M/N = r, K = c): SBO = RG, LBO = 128.  MN-major use (M/N = c, K = r): SBO = 128, LBO = RG.
__device__ __forceinline__ uint32_t core_off(int r, int c, int Cc) {
  return (uint32_t)((r >> 3) * (Cc / 4) * 128 + (c >> 2) * 128 + (r & 7) * 16 + (c & 3) * 4);
}

__global__ void __launch_bounds__(128, 1) umma_probe(const Params p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int M = 128, N = p.N, K = p.K;
  uint8_t* sA = smem;
  uint8_t* sB = sA + M * K * 4;
  uint8_t* sA2 = sB + N * K * 4;
  uint8_t* sB2 = sA2 + M * K * 4;
  // stage operands.  K-major: stored tensor is [mn rows][k cols]; MN-major: stored tensor is [k rows][mn cols]
  for (int i = tid; i < M * K; i += 128) {
    const int m = i / K, k = i % K;
    const uint32_t off = p.a_mn ? core_off(k, m, M) : core_off(m, k, K);
    *reinterpret_cast<float*>(sA + off) = p.A[i];
    if (p.A2) *reinterpret_cast<float*>(sA2 + off) = p.A2[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    const int n = i / K, k = i % K;
    const uint32_t off = p.b_mn ? core_off(k, n, N) : core_off(n, k, K);
    *reinterpret_cast<float*>(sB + off) = p.B[i];
    if (p.B2) *reinterpret_cast<float*>(sB2 + off) = p.B2[i];
  }
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // generic-proxy smem writes -> async proxy (tensor core) reads
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tslot;
  const uint32_t tD = tbase;            // columns [0, N)
  const uint32_t tA = tbase + 256;      // columns [256, 256 + K) : A operand in tensor memory
  const uint32_t lane_base = ((uint32_t)(warp & 3) * 32u) << 16;

  if (p.a_tmem) {  // thread = row m writes its K values into columns [256, 256+K) of its lane
    for (int k0 = 0; k0 < K; k0 += 8) {
      float v[8];
      for (int j = 0; j < 8; ++j) v[j] = p.A[tid * K + k0 + j];
      tmem_st8(tA + lane_base + k0, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  const uint32_t idesc = make_idesc(M, N, p.a_mn, p.b_mn);
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    // K-major: SBO = row-group stride, LBO = 128, k-step advance 256.  MN-major: SBO = 128, LBO = k-group stride = advance.
    const uint32_t a_rg = p.a_mn ? (M / 4) * 128 : (K / 4) * 128;
    const uint32_t b_rg = p.b_mn ? (N / 4) * 128 : (K / 4) * 128;
    const uint32_t a_lbo = p.a_mn ? a_rg : 128, a_sbo = p.a_mn ? 128 : a_rg, a_adv = p.a_mn ? a_rg : 256;
    const uint32_t b_lbo = p.b_mn ? b_rg : 128, b_sbo = p.b_mn ? 128 : b_rg, b_adv = p.b_mn ? b_rg : 256;
    t0 = clock64();
    for (int rep = 0; rep < p.reps; ++rep) {
      for (int pass = 0; pass < (p.A2 ? 3 : 1); ++pass) {
        const uint8_t* a_src = (pass == 1) ? sA2 : sA;   // pass 0: hi*hi, 1: lo*hi, 2: hi*lo
        const uint8_t* b_src = (pass == 2) ? sB2 : sB;
        for (int ks = 0; ks < K / 8; ++ks) {
          const uint64_t bd = make_desc(smem_u32(b_src) + ks * b_adv, b_lbo, b_sbo);
          const uint32_t acc = (rep | pass | ks) ? 1u : 0u;
          if (p.a_tmem) {
            mma_ts(tD, tA + ks * 8, bd, idesc, acc);
          } else {
            const uint64_t ad = make_desc(smem_u32(a_src) + ks * a_adv, a_lbo, a_sbo);
            mma_ss(tD, ad, bd, idesc, acc);
          }
        }
      }
    }
    mma_commit(smem_u32(&bar));
  }
  mbar_wait_bounded(smem_u32(&bar), 0);
  if (tid == 0) {
    t1 = clock64();
    p.cycles[0] = t1 - t0;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tmem_ld16(tD + lane_base + c0, v);
    for (int j = 0; j < 16; ++j) p.D[tid * N + c0 + j] = v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512u) : "memory");
}

static float trunc_tf32(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u &= 0xffffe000u;
  memcpy(&v, &u, 4);
  return v;
}
static float round_tf32(float v) {  // round to nearest (ties away), what cvt.rna.tf32 does
  uint32_t u;
  memcpy(&u, &v, 4);
  u += 0x1000u;
  u &= 0xffffe000u;
  memcpy(&v, &u, 4);
  return v;
}

struct Result {
  double err_trunc, err_round, err_full;
  long long cycles;
  double bias_ulp, rms_ulp;   // (D - exact product of the truncated inputs) * sign(exact) / ulp(exact): mean and rms
};

static Result run(int N, int K, int a_mn, int b_mn, int a_tmem, int reps, bool split3) {
  const int M = 128;
  std::vector<float> A(M * K), B(N * K), A2, B2;
  for (auto& v : A) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  for (auto& v : B) v = (float)rand() / RAND_MAX * 2.f - 1.f;
  if (split3) {
    A2.resize(M * K);
    B2.resize(N * K);
    for (int i = 0; i < M * K; ++i) A2[i] = A[i] - trunc_tf32(A[i]);
    for (int i = 0; i < N * K; ++i) B2[i] = B[i] - trunc_tf32(B[i]);
  }
  float *dA, *dB, *dA2 = nullptr, *dB2 = nullptr, *dD;
  long long* dC;
  CK(cudaMalloc(&dA, A.size() * 4));
  CK(cudaMalloc(&dB, B.size() * 4));
  CK(cudaMalloc(&dD, M * N * 4));
  CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  if (split3) {
    CK(cudaMalloc(&dA2, A.size() * 4));
    CK(cudaMalloc(&dB2, B.size() * 4));
    CK(cudaMemcpy(dA2, A2.data(), A.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB2, B2.data(), B.size() * 4, cudaMemcpyHostToDevice));
  }
  Params p = {dA, dB, dA2, dB2, dD, dC, N, K, a_mn, b_mn, a_tmem, reps};
  const size_t smem = (size_t)(2 * M * K + 2 * N * K) * 4;
  CK(cudaFuncSetAttribute(umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  umma_probe<<<1, 128, smem>>>(p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> D(M * N);
  long long cyc;
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  Result r = {0, 0, 0, cyc, 0, 0};
  double sb = 0, sq = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double st = 0, sr = 0, sf = 0;
      for (int k = 0; k < K; ++k) {
        st += (double)trunc_tf32(A[m * K + k]) * trunc_tf32(B[n * K + k]);
        sr += (double)round_tf32(A[m * K + k]) * round_tf32(B[n * K + k]);
        sf += (double)A[m * K + k] * B[n * K + k];
      }
      const double d = D[m * N + n] / (double)reps;
      r.err_trunc = fmax(r.err_trunc, fabs(d - st));
      r.err_round = fmax(r.err_round, fabs(d - sr));
      r.err_full = fmax(r.err_full, fabs(d - sf));
      if (reps == 1 && st != 0) {
        int ex;
        frexp(st, &ex);
        const double ulp = ldexp(1.0, ex - 24);
        const double e = (d - st) / ulp * (st > 0 ? 1.0 : -1.0);
        sb += e;
        sq += e * e;
      }
    }
  r.bias_ulp = sb / (M * N);
  r.rms_ulp = sqrt(sq / (M * N));
  cudaFree(dA);
  cudaFree(dB);
  cudaFree(dD);
  cudaFree(dC);
  if (dA2) cudaFree(dA2);
  if (dB2) cudaFree(dB2);
  return r;
}

int main2();
int main() {
  srand(1);
  printf("test                               N    K  a_mn b_mn a_tm | err_vs_trunc err_vs_round err_vs_full | cycles  cyc/mma\n");
  struct Case {
    const char* name;
    int N, K, a_mn, b_mn, a_tm, reps;
    bool split;
  } cases[] = {
      {"K-major A, K-major B K=8", 32, 8, 0, 0, 0, 1, false},
      {"K-major A, K-major B K=16", 32, 16, 0, 0, 0, 1, false},
      {"K-major A, K-major B", 32, 32, 0, 0, 0, 1, false},
      {"K-major A, K-major B K=128", 32, 128, 0, 0, 0, 1, false},
      {"K-major A, K-major B N=240", 240, 32, 0, 0, 0, 1, false},
      {"K-major A, K-major B N=16", 16, 32, 0, 0, 0, 1, false},
      {"K-major A, MN-major B", 32, 32, 0, 1, 0, 1, false},
      {"MN-major A, MN-major B", 32, 128, 1, 1, 0, 1, false},
      {"MN-major A, MN-major B N=64", 64, 128, 1, 1, 0, 1, false},
      {"MN-major A, K-major B", 32, 32, 1, 0, 0, 1, false},
      {"TMEM A, K-major B", 32, 32, 0, 0, 1, 1, false},
      {"TMEM A, MN-major B N=16", 16, 32, 0, 1, 1, 1, false},
      {"3xTF32 K-major", 32, 32, 0, 0, 0, 1, true},
      {"3xTF32 MN/MN K=128", 32, 128, 1, 1, 0, 1, true},
      {"timing N=16 K=32 x64", 16, 32, 0, 0, 0, 64, false},
      {"timing N=32 K=32 x64", 32, 32, 0, 0, 0, 64, false},
      {"timing N=64 K=32 x64", 64, 32, 0, 0, 0, 64, false},
      {"timing N=128 K=32 x64", 128, 32, 0, 0, 0, 64, false},
      {"timing N=240 K=32 x64", 240, 32, 0, 0, 0, 64, false},
      {"timing N=32 K=128 MN/MN x64", 32, 128, 1, 1, 0, 64, false},
      {"timing N=64 K=128 MN/MN x64", 64, 128, 1, 1, 0, 64, false},
      {"timing TMEM A N=16 K=32 x64", 16, 32, 0, 0, 1, 64, false},
      {"timing TMEM A N=32 K=32 x64", 32, 32, 0, 0, 1, 64, false},
  };
  for (const Case& c : cases) {
    Result r = run(c.N, c.K, c.a_mn, c.b_mn, c.a_tm, c.reps, c.split);
    const int nmma = c.reps * (c.K / 8) * (c.split ? 3 : 1);
    printf("%-32s %4d %4d   %d    %d    %d   | %11.3e %11.3e %11.3e | %7lld %7.1f | bias %+.3f ulp rms %.3f\n", c.name, c.N, c.K, c.a_mn,
           c.b_mn, c.a_tm, r.err_trunc, r.err_round, r.err_full, r.cycles, (double)r.cycles / nmma, r.bias_ulp, r.rms_ulp);
    fflush(stdout);
  }
  return main2();
}

// ------------------------------------------------------------------------------------------------------------
// Layout discovery: one shared-memory word of the operand under test is 1.0, everything else 0; the other operand
// is a (known-good, K-major) partial identity, so D shows which (mn, k) element that word is.  M = 128, N = 16, K = 8.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}

struct DiscParams {
  int layout;       // descriptor layout type of the operand under test: 0 none, 1 = 128B swizzle with 32B base (tf32 MN-major)
  int which;        // 0: operand under test is A (MN-major), 1: B (MN-major), 2: A K-major, 3: B K-major
  uint32_t lbo, sbo;
  int* out;         // per offset bit: [mn, k, count]
  int nbits;
};

__global__ void __launch_bounds__(128, 1) umma_discover(const DiscParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  __shared__ int found[3];
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sT = smem;                 // operand under test: 64 KB
  uint8_t* sI = smem + 65536;         // identity operand, K-major core matrices, [128 or 16][8]: RG = 256
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(32u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tD = tslot;
  const uint32_t lane_base = ((uint32_t)(warp & 3) * 32u) << 16;
  const bool testA = (p.which == 0 || p.which == 2);
  const int t_mn = (p.which < 2) ? 1 : 0;
  uint32_t phase = 0;
  for (int bit = 0; bit <= p.nbits; ++bit) {
    const uint32_t off = bit == 0 ? 0u : (1u << (bit + 1));   // 0, 4, 8, ..., byte offsets
    for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<float*>(sT)[i] = 0.f;
    for (int i = tid; i < 128 * 8; i += 128) reinterpret_cast<float*>(sI)[i] = 0.f;
    if (tid < 3) found[tid] = tid == 2 ? 0 : -1;
    __syncthreads();
    if (tid == 0) *reinterpret_cast<float*>(sT + off) = 1.f;
    if (tid < 8) *reinterpret_cast<float*>(sI + core_off(tid, tid, 8)) = 1.f;   // I[r][r] = 1, r < 8
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
        const uint64_t td = make_desc(smem_u32(sT), p.lbo, p.sbo) | ((uint64_t)p.layout << 61);
        const uint64_t id = make_desc(smem_u32(sI), 128, 256);
        const uint32_t idesc = make_idesc(128, 16, testA ? t_mn : 0, testA ? 0 : t_mn);
        if (testA) mma_ss(tD, td, id, idesc, 0);
        else mma_ss(tD, id, td, idesc, 0);
        mma_commit(smem_u32(&bar));
      }
      __syncwarp();
    }
    mbar_wait_bounded(smem_u32(&bar), phase);
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float v[16];
    tmem_ld16(tD + lane_base, v);
    for (int j = 0; j < 16; ++j)
      if (v[j] != 0.f) {
        // testA: D[m][n] = A[m][k=n] (n < 8)  -> mn = tid, k = j.   testB: D[m][n] = B[n][k=m] (m < 8) -> mn = j, k = tid
        found[0] = testA ? tid : j;
        found[1] = testA ? j : tid;
        atomicAdd(&found[2], 1);
      }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 3) p.out[bit * 3 + tid] = found[tid];
    __syncthreads();
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tD), "r"(32u) : "memory");
}

// ------------------------------------------------------------------------------------------------------------
// Issue-rate probe: one elected lane issues `n` MMAs (M = 128, K = 8 each) round-robin over `nacc` accumulators.
// ------------------------------------------------------------------------------------------------------------
struct RateParams {
  int n;            // MMAs issued (multiple of 8)
  long long* cycles;
};
// N, NACC (accumulators used round-robin), ATM (A from tensor memory) are compile-time so that the issue loop is
// nothing but UTCHMMA + a handful of uniform adds
template <int N, int NACC, int ATM>
__global__ void __launch_bounds__(128, 1) umma_rate(const RateParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 63);
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tD = tslot;
  long long t0 = 0;
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, N, 0, 0);
      const uint64_t ad0 = make_desc(smem_u32(smem), 128, 1024);
      const uint64_t bd0 = make_desc(smem_u32(smem) + 32768, 128, 1024);
      t0 = clock64();
      for (int i = 0; i < p.n; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t acc = tD + (uint32_t)((j % NACC) * N);
          if (ATM) mma_ts(acc, tD + 480 + (j & 3) * 8, bd0 + (uint64_t)((j & 3) * 16), idesc, 1);
          else mma_ss(acc, ad0 + (uint64_t)((j & 3) * 16), bd0 + (uint64_t)((j & 3) * 16), idesc, 1);
        }
      }
      mma_commit(smem_u32(&bar));
      p.cycles[1] = clock64() - t0;    // issue time alone
    }
    __syncwarp();
  }
  mbar_wait_bounded(smem_u32(&bar), 0);
  if (warp == 0 && t0 != 0) p.cycles[0] = clock64() - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tD), "r"(512u) : "memory");
}

template <int N, int NACC, int ATM>
static void rate_case(long long* dC) {
  CK(cudaFuncSetAttribute(umma_rate<N, NACC, ATM>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  RateParams p = {2048, dC};
  umma_rate<N, NACC, ATM><<<1, 128, 160 * 1024>>>(p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  long long c[2];
  CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost));
  printf("      %4d  %d     %d    | %8.1f        %8.1f   (floor %d)\n", N, NACC, ATM, c[0] / 2048.0, c[1] / 2048.0, N / 2);
}

// ------------------------------------------------------------------------------------------------------------
// Sequence probe: the MMA stream of the backward generator kernel (A from tensor memory at changing columns, B either
// K-major no-swizzle or MN-major SWIZZLE_128B_BASE32B, one accumulator), optionally with NOISE warps that keep
// tcgen05.ld / tcgen05.st traffic (and shared-memory stores) going on the same SM.
// ------------------------------------------------------------------------------------------------------------
template <int N, int BMN, int NOISE>
__global__ void __launch_bounds__(128 + 512, 1) umma_seq(const RateParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  __shared__ volatile int stop;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 63);
  if (tid == 0) {
    stop = 0;
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tD = tslot;
  long long t0 = 0;
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, N, 0, BMN);
      const uint32_t b0 = smem_u32(smem);
      t0 = clock64();
      for (int i = 0; i < p.n; i += 16) {
#pragma unroll
        for (int ks = 0; ks < 16; ++ks) {
          const uint64_t bd = BMN ? (make_desc(b0 + ks * 1024, 16384, 512) | ((uint64_t)1 << 61))
                                  : make_desc(b0 + (ks & 3) * 256 + (ks >> 2) * 4096, 128, 1024);
          mma_ts(tD, tD + 256 + ks * 8, bd, idesc, 1);
        }
      }
      mma_commit(smem_u32(&bar));
      p.cycles[1] = clock64() - t0;
    }
    __syncwarp();
    mbar_wait_bounded(smem_u32(&bar), 0);
    if (t0 != 0) p.cycles[0] = clock64() - t0;
    stop = 1;
  } else if (warp >= 4 && NOISE == 3) {
    // ALU-bound neighbours on every scheduler (the epilogue warps of the generator kernels): does the issuing thread starve?
    float a = (float)tid, b = 1.0001f, c = 0.5f, d = 0.25f;
    while (!stop) {
#pragma unroll
      for (int j = 0; j < 64; ++j) {
        a = fmaf(a, b, c);
        c = fmaf(c, b, d);
        d = fmaf(d, b, a);
      }
    }
    if (a + c + d == 12345.f) p.cycles[1] = 0;
  } else if (warp >= 4 && NOISE) {
    // lanes of this warp's quarter, columns 384.. : load, touch, store back (as the epilogue and helper warps do)
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    float v[16];
    float* sm = reinterpret_cast<float*>(smem + 96 * 1024) + tid * 4;
    while (!stop) {
      tmem_ld16(tD + lane_base + 384 + (warp >> 2) * 16, v);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 16; ++j) v[j] = v[j] * 1.0001f + 1.f;
      if (NOISE > 1) *reinterpret_cast<float4*>(sm) = make_float4(v[0], v[1], v[2], v[3]);
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::
                   "r"(tD + lane_base + 384 + (warp >> 2) * 16), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]),
                   "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]) : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tD), "r"(512u) : "memory");
}

template <int N, int BMN, int NOISE>
static void seq_case(long long* dC) {
  CK(cudaFuncSetAttribute(umma_seq<N, BMN, NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  RateParams p = {2048, dC};
  umma_seq<N, BMN, NOISE><<<1, 128 + 512, 160 * 1024>>>(p);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  long long c[2];
  CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost));
  printf("seq   %4d  b_mn %d noise %d | %8.1f        %8.1f   (floor %d)\n", N, BMN, NOISE, c[0] / 2048.0, c[1] / 2048.0, N / 2);
}

// ------------------------------------------------------------------------------------------------------------
// Burst probe: latency of a short dependent chain of MMAs (issue -> commit -> mbarrier completion), which is what the
// pipelined generator kernels pay per stage (as opposed to the steady-state rate above).
// ------------------------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128, 1) umma_burst(int count, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.001f * (i & 63);
  if (tid == 0) {
    mbar_init(smem_u32(&bar), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tD = tslot;
  if (warp == 0) {
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, N, 0, 0);
      const uint32_t b0 = smem_u32(smem);
      long long issue = 0, done = 0;
      uint32_t phase = 0;
      for (int r = 0; r < reps; ++r) {
        const long long t0 = clock64();
        for (int i = 0; i < count; ++i)
          mma_ts(tD, tD + 256 + (i & 3) * 8, make_desc(b0 + (i & 3) * 256, 128, 1024), idesc, i != 0);
        mma_commit(smem_u32(&bar));
        const long long t1 = clock64();
        mbar_wait_bounded(smem_u32(&bar), phase);
        phase ^= 1;
        const long long t2 = clock64();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (r > 0) issue += t1 - t0, done += t2 - t0;
      }
      out[0] = issue / (reps - 1);
      out[1] = done / (reps - 1);
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tD), "r"(512u) : "memory");
}
template <int N>
static void burst_case(long long* dC, int count) {
  CK(cudaFuncSetAttribute(umma_burst<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  umma_burst<N><<<1, 128, 64 * 1024>>>(count, 9, dC);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  long long c[2];
  CK(cudaMemcpy(c, dC, 16, cudaMemcpyDeviceToHost));
  printf("burst N %3d x %2d MMAs | issued after %5lld, complete after %5lld cycles (floor %d)\n", N, count, c[0], c[1], count * N / 2);
}

int main2() {
  // layout discovery
  int* dOut;
  CK(cudaMalloc(&dOut, 64 * 3 * 4));
  CK(cudaFuncSetAttribute(umma_discover, cudaFuncAttributeMaxDynamicSharedMemorySize, 70 * 1024));
  const char* names[4] = {"A MN-major", "B MN-major", "A K-major", "B K-major"};
  for (int which = 0; which < 4; ++which) {
    DiscParams p = {0, which, 8192u, 2048u, dOut, 14};
    umma_discover<<<1, 128, 69 * 1024>>>(p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    int h[64 * 3];
    CK(cudaMemcpy(h, dOut, sizeof(h), cudaMemcpyDeviceToHost));
    printf("discover %s (LBO=8192 SBO=2048): byte offset -> (mn, k) [nonzero count]\n", names[which]);
    for (int bit = 0; bit <= 14; ++bit)
      printf("   off %6u -> mn %4d k %3d  [%d]\n", bit == 0 ? 0u : (1u << (bit + 1)), h[bit * 3], h[bit * 3 + 1], h[bit * 3 + 2]);
  }
  for (int which = 0; which < 2; ++which) {
    DiscParams p = {1, which, 16384u, 2048u, dOut, 14};
    umma_discover<<<1, 128, 69 * 1024>>>(p);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    int h[64 * 3];
    CK(cudaMemcpy(h, dOut, sizeof(h), cudaMemcpyDeviceToHost));
    printf("discover %s with layout SWIZZLE_128B_BASE32B (LBO=16384 SBO=2048): byte offset -> (mn, k) [nonzero count]\n", names[which]);
    for (int bit = 0; bit <= 14; ++bit)
      printf("   off %6u -> mn %4d k %3d  [%d]\n", bit == 0 ? 0u : (1u << (bit + 1)), h[bit * 3], h[bit * 3 + 1], h[bit * 3 + 2]);
  }
  // issue rate
  long long* dC;
  CK(cudaMalloc(&dC, 16));
  printf("rate:   N nacc a_tmem | total cyc/mma   issue cyc/mma\n");
  rate_case<16, 1, 0>(dC);  rate_case<16, 4, 0>(dC);
  rate_case<32, 1, 0>(dC);  rate_case<32, 4, 0>(dC);
  rate_case<64, 1, 0>(dC);  rate_case<64, 4, 0>(dC);
  rate_case<128, 1, 0>(dC); rate_case<128, 2, 0>(dC);
  rate_case<240, 1, 0>(dC); rate_case<240, 2, 0>(dC);
  rate_case<256, 1, 0>(dC);
  rate_case<16, 1, 1>(dC);  rate_case<16, 4, 1>(dC);
  rate_case<32, 1, 1>(dC);  rate_case<32, 4, 1>(dC);
  rate_case<64, 1, 1>(dC);  rate_case<64, 4, 1>(dC);
  rate_case<128, 1, 1>(dC); rate_case<240, 1, 1>(dC);
  seq_case<32, 0, 0>(dC); seq_case<32, 0, 1>(dC); seq_case<32, 0, 2>(dC);
  seq_case<64, 1, 0>(dC); seq_case<64, 1, 1>(dC); seq_case<64, 1, 2>(dC);
  seq_case<64, 0, 0>(dC); seq_case<16, 0, 0>(dC); seq_case<16, 0, 1>(dC);
  seq_case<32, 0, 3>(dC); seq_case<64, 1, 3>(dC);
  for (int c : {1, 2, 4, 12, 32}) burst_case<32>(dC, c);
  for (int c : {1, 4, 16, 32}) burst_case<64>(dC, c);
  return 0;
}
