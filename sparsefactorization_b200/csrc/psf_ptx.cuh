// PTX / register-shuffle helpers shared by the fused chain kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

constexpr int CT = 8;  // channels held per thread == C of the fused paths

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "PSF_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra PSF_DONE_%=;\n"
      "bra PSF_WAIT_%=;\n"
      "PSF_DONE_%=:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// 1-D bulk async copy global -> shared (TMA), completion counted on an mbarrier.
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// same, with an L2 evict-first policy: W is streamed once, it must not push the exchanged states out of L2
__device__ __forceinline__ void tma_load_1d_stream(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar,
                                                   unsigned long long policy) {
  if (policy == 0) {
    tma_load_1d(dst, src, bytes, bar);
    return;
  }
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          dst),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ unsigned long long make_evict_first_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// generic-proxy writes (st.global) -> later async-proxy reads (TMA) of the same bytes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
// Hardware cluster barrier; release/acquire at cluster scope orders the global-memory exchange.
__device__ __forceinline__ void cluster_arrive() {
  asm volatile("barrier.cluster.arrive.release;" ::: "memory");
}
__device__ __forceinline__ void cluster_wait() {
  asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void consumer_bar_sync(int nthreads) {
  asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

__device__ __forceinline__ float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st_cg4(float* p, float4 v) { __stcg(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ void st_cs4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
// same, but the compiler may not narrow it to the lanes that are used (narrow loads at a 16 B lane
// stride are bank conflicted)
__device__ __forceinline__ float4 lds4_full(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(smem_u32(p)));
  return v;
}

// acc += a * b on 4 lanes as two packed fma.rn.f32x2 (FFMA2 on sm_100): half the FMA issue slots
__device__ __forceinline__ void fma4(float4& acc, const float4& a, const float4& b) {
  unsigned long long* pc = reinterpret_cast<unsigned long long*>(&acc);
  const unsigned long long* pa = reinterpret_cast<const unsigned long long*>(&a);
  const unsigned long long* pb = reinterpret_cast<const unsigned long long*>(&b);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(pc[0]) : "l"(pa[0]), "l"(pb[0]));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(pc[1]) : "l"(pa[1]), "l"(pb[1]));
}
__device__ __forceinline__ void fma2(float& c0, float& c1, float a0, float a1, float b0, float b1) {
  float2 c = make_float2(c0, c1), a = make_float2(a0, a1), b = make_float2(b0, b1);
  unsigned long long* pc = reinterpret_cast<unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(*pc)
      : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)));
  c0 = c.x;
  c1 = c.y;
}
__device__ __forceinline__ void add4(float4& acc, const float4& a) {
  acc.x += a.x;
  acc.y += a.y;
  acc.z += a.z;
  acc.w += a.w;
}
// rows [s, s+4) out of the two aligned quads X = rows [0,4), Y = rows [4,8); s in 1..3 (warp-uniform)
__device__ __forceinline__ float4 select_quad(const float4& X, const float4& Y, int s) {
  if (s == 1) return make_float4(X.y, X.z, X.w, Y.x);
  if (s == 2) return make_float4(X.z, X.w, Y.x, Y.y);
  return make_float4(X.w, Y.x, Y.y, Y.z);
}

// Natural [row][8] <-> per-thread channel planes (4 rows x 8 channels), pure register shuffling.
// same with a row stride of `rs` floats (C > 8 channels: the 8-channel slab [base, base + 8) of every row)
__device__ __forceinline__ void load_rows_strided(const float* __restrict__ base, int rs, float4 (&pl)[CT]) {
  float4 lo[4], hi[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    lo[r] = __ldcs(reinterpret_cast<const float4*>(base + (long long)r * rs));
    hi[r] = __ldcs(reinterpret_cast<const float4*>(base + (long long)r * rs + 4));
  }
  pl[0] = make_float4(lo[0].x, lo[1].x, lo[2].x, lo[3].x);
  pl[1] = make_float4(lo[0].y, lo[1].y, lo[2].y, lo[3].y);
  pl[2] = make_float4(lo[0].z, lo[1].z, lo[2].z, lo[3].z);
  pl[3] = make_float4(lo[0].w, lo[1].w, lo[2].w, lo[3].w);
  pl[4] = make_float4(hi[0].x, hi[1].x, hi[2].x, hi[3].x);
  pl[5] = make_float4(hi[0].y, hi[1].y, hi[2].y, hi[3].y);
  pl[6] = make_float4(hi[0].z, hi[1].z, hi[2].z, hi[3].z);
  pl[7] = make_float4(hi[0].w, hi[1].w, hi[2].w, hi[3].w);
}
__device__ __forceinline__ void store_rows_strided(float* __restrict__ base, int rs, const float4 (&pl)[CT]) {
  __stcs(reinterpret_cast<float4*>(base), make_float4(pl[0].x, pl[1].x, pl[2].x, pl[3].x));
  __stcs(reinterpret_cast<float4*>(base + 4), make_float4(pl[4].x, pl[5].x, pl[6].x, pl[7].x));
  __stcs(reinterpret_cast<float4*>(base + (long long)rs), make_float4(pl[0].y, pl[1].y, pl[2].y, pl[3].y));
  __stcs(reinterpret_cast<float4*>(base + (long long)rs + 4), make_float4(pl[4].y, pl[5].y, pl[6].y, pl[7].y));
  __stcs(reinterpret_cast<float4*>(base + 2LL * rs), make_float4(pl[0].z, pl[1].z, pl[2].z, pl[3].z));
  __stcs(reinterpret_cast<float4*>(base + 2LL * rs + 4), make_float4(pl[4].z, pl[5].z, pl[6].z, pl[7].z));
  __stcs(reinterpret_cast<float4*>(base + 3LL * rs), make_float4(pl[0].w, pl[1].w, pl[2].w, pl[3].w));
  __stcs(reinterpret_cast<float4*>(base + 3LL * rs + 4), make_float4(pl[4].w, pl[5].w, pl[6].w, pl[7].w));
}
__device__ __forceinline__ void load_rows_natural(const float* __restrict__ base, float4 (&pl)[CT]) {
  float4 lo[4], hi[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    lo[r] = __ldcs(reinterpret_cast<const float4*>(base + r * CT));
    hi[r] = __ldcs(reinterpret_cast<const float4*>(base + r * CT + 4));
  }
  pl[0] = make_float4(lo[0].x, lo[1].x, lo[2].x, lo[3].x);
  pl[1] = make_float4(lo[0].y, lo[1].y, lo[2].y, lo[3].y);
  pl[2] = make_float4(lo[0].z, lo[1].z, lo[2].z, lo[3].z);
  pl[3] = make_float4(lo[0].w, lo[1].w, lo[2].w, lo[3].w);
  pl[4] = make_float4(hi[0].x, hi[1].x, hi[2].x, hi[3].x);
  pl[5] = make_float4(hi[0].y, hi[1].y, hi[2].y, hi[3].y);
  pl[6] = make_float4(hi[0].z, hi[1].z, hi[2].z, hi[3].z);
  pl[7] = make_float4(hi[0].w, hi[1].w, hi[2].w, hi[3].w);
}
__device__ __forceinline__ void store_rows_natural(float* __restrict__ base, const float4 (&pl)[CT]) {
  st_cs4(base + 0 * CT, make_float4(pl[0].x, pl[1].x, pl[2].x, pl[3].x));
  st_cs4(base + 0 * CT + 4, make_float4(pl[4].x, pl[5].x, pl[6].x, pl[7].x));
  st_cs4(base + 1 * CT, make_float4(pl[0].y, pl[1].y, pl[2].y, pl[3].y));
  st_cs4(base + 1 * CT + 4, make_float4(pl[4].y, pl[5].y, pl[6].y, pl[7].y));
  st_cs4(base + 2 * CT, make_float4(pl[0].z, pl[1].z, pl[2].z, pl[3].z));
  st_cs4(base + 2 * CT + 4, make_float4(pl[4].z, pl[5].z, pl[6].z, pl[7].z));
  st_cs4(base + 3 * CT, make_float4(pl[0].w, pl[1].w, pl[2].w, pl[3].w));
  st_cs4(base + 3 * CT + 4, make_float4(pl[4].w, pl[5].w, pl[6].w, pl[7].w));
}

}  // namespace
