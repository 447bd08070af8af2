// tcgen05 / tensor-memory helpers shared by the generator kernels (sm_100a).
#pragma once
#include <stdio.h>
#include "psf_ptx.cuh"

namespace {

__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
               : "=r"(ok)
               : "r"(bar), "r"(parity)
               : "memory");
  return ok != 0;
}
// Bounded waits: a protocol error traps (the launch fails with an error) instead of hanging the device.  A failed
// try_wait comes back within a few cycles, so every retry backs off with nanosleep: a spinning warp would otherwise take
// a third of its scheduler's issue slots from the warps doing the work (measured, profiles/r02_gen_notes.md).
#ifdef PSF_TC_DEBUG   // name the barrier a protocol error is stuck on (byte offset in shared memory) before trapping
#define PSF_TC_TIMEOUT(bar, parity) do { if ((threadIdx.x & 31) == 0 || (threadIdx.x >> 5) >= 20) printf("mbarrier timeout: block %d warp %d smem offset %u parity %u\n", blockIdx.x, threadIdx.x >> 5, bar, parity); } while (0)
#define PSF_TC_SPINS (1u << 23)
#define PSF_TC_GIVE_UP return      /* keep going so that every stuck warp reports */
#else
#define PSF_TC_TIMEOUT(bar, parity)
#define PSF_TC_SPINS (1u << 26)
#define PSF_TC_GIVE_UP __trap()
#endif
__device__ __forceinline__ void mbar_wait_tc(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    __nanosleep(20);
    if (++spins > PSF_TC_SPINS) {
      PSF_TC_TIMEOUT(bar, parity);
      PSF_TC_GIVE_UP;
    }
  }
}
// same for the lone MMA-issuing threads: one lane polling costs next to nothing in issue slots, and every cycle it
// sleeps past an arrival is added to a hop of the pipeline
__device__ __forceinline__ void mbar_wait_hot(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    if (++spins > PSF_TC_SPINS) {
      PSF_TC_TIMEOUT(bar, parity);
      PSF_TC_GIVE_UP;
    }
  }
}
// same for warps with slack (loaders)
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    __nanosleep(100);
    if (++spins > PSF_TC_SPINS / 16) {
      PSF_TC_TIMEOUT(bar, parity);
      PSF_TC_GIVE_UP;
    }
  }
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n.reg .pred p;\nelect.sync _|p, 0xffffffff;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(pred));
  return pred != 0;
}
// K-major, no-swizzle canonical layout: 8-row x 16-byte core matrices; LBO = distance of the two 16-byte K chunks of
// one MMA, SBO = distance of 8-row groups (layout verified by tools/microbench/umma_probe.cu)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {  // tf32 x tf32 -> f32, A and B K-major
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
               "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(acc)
               : "memory");
}
// the same with the shared-memory descriptor as two words (descriptors of one operand differ in the low word only)
__device__ __forceinline__ void mma_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t bdesc_lo, uint32_t bdesc_hi, uint32_t idesc,
                                        uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\n.reg .b64 bd;\nsetp.ne.b32 p, %5, 0;\nmov.b64 bd, {%2, %3};\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n}\n" ::"r"(d_tmem),
      "r"(a_tmem), "r"(bdesc_lo), "r"(bdesc_hi), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld_tc() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st_tc() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32f(uint32_t taddr, float (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]),
        "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]),
        "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16f(uint32_t taddr, float (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]),
        "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32f(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]),
      "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]), "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]),
      "f"(v[21]), "f"(v[22]), "f"(v[23]), "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]),
      "f"(v[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16f(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "f"(v[8]), "f"(v[9]), "f"(v[10]),
      "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st8f(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "f"(v[0]),
               "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_ld8f(uint32_t taddr, float (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "r"(taddr)
               : "memory");
}
// 3xTF32 operand split: hi = v rounded to nearest tf32 (stored explicitly - the tensor core would truncate), lo = v - hi
// (exact; at most 12 significant bits, of which the tensor core keeps 11: the split loses < 2^-24 |v|)
__device__ __forceinline__ float tf32_hi(float v) { return __uint_as_float((__float_as_uint(v) + 0x1000u) & 0xffffe000u); }
__device__ __forceinline__ float tf32_lo(float v) { return v - tf32_hi(v); }

// Packed FP32 pairs (FMUL2 / FFMA2 on sm_100: one issue slot for two values).  The epilogue warps of the generator
// kernels are bound by issue slots, and their per-element arithmetic is the same polynomial with the same constants.
__device__ __forceinline__ uint64_t pack2f(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void unpack2f(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t mul2f(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma2f(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t dup2f(float a) { return pack2f(a, a); }
// Veltkamp's form of the tf32 split on a pair: t = (2^13 + 1) v, hi = t - (t - v) is v rounded to nearest with 11
// significant bits - exactly a tf32 - and lo = v - hi is exact; four packed instructions for (hi, lo) of two values
__device__ __forceinline__ void tf32_split2(uint64_t v, uint64_t& hi, uint64_t& lo) {
  const uint64_t m1 = dup2f(-1.f);
  const uint64_t t = mul2f(v, dup2f(8193.f));
  hi = fma2f(fma2f(v, m1, t), m1, t);                   // t - (t - v)
  lo = fma2f(hi, m1, v);
}

// The tensor core adds into its fp32 accumulator by truncation (measured: tools/microbench/umma_probe.cu,
// tools/gen_bias_check.py), so a K = 32 product - four full-magnitude accumulations after the small 3xTF32 terms - comes
// out biased towards zero, where round-to-nearest FMA chains are unbiased: every W shrinks by 1.5e-7 on average (1.7 ulp),
// and the 14 chained factors add that up coherently (logits of the n = 8192 Temporal Order model off by 2e-6 relative,
// twice the error of an FP32-FMA evaluation).  Stepping the result one unit in the last place away from zero removes the
// mean error (measured residual 0.006 ulp on default-initialised generators, rms as the FP32-FMA kernel).
__device__ __forceinline__ float tc_unbias(float v) { return __uint_as_float(__float_as_uint(v) + 1u); }

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
}  // namespace
