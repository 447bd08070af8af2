// Fused chord-factor chain, second generation ("quad-interleaved clusters"), sm_100a, fp32, C = 8 G channels
// (each 8-channel slab of a sample is processed like a sample of its own), n = 2^p rows, p = 7 .. 16, with the
// standard chord offsets 0,1,2,4,...,n/2 (L = p+1 taps).  The description below is for the tuned shape
// J = 2048 (n = 2048 .. 16384); smaller n use one CTA per sample, n = 32768 / 65536 use J = 4096.
//
// One cluster of Q = n/J CTAs owns one sample (J = 2048 rows per CTA).  Unlike the first generation
// (psf_chain_sm100.cu: contiguous row blocks), CTA q owns the row QUADS u = i>>2 with u mod Q == q:
//
//     row i = 4*(lq*Q + q) + lo,   lq = local quad index, lo = 0..3
//
// With the CTA bits at the BOTTOM of the quad index, a chord offset either
//   * stays inside the thread's quad pair     (offsets 1, 2: register shuffles of two quads),
//   * moves to a sibling CTA, same local quad (offsets 4..2Q: "remote" taps, log2 Q of them), or
//   * moves inside the CTA by a power of two  (offsets >= 4Q: local quad lq + 2^t, cyclic in J),
// so there is no halo, the cyclic wrap is an AND on a shared-memory address, every operand access is
// an aligned 128-bit quad, and the whole tap schedule is known at compile time: the inner loop is
// straight-line code (no per-tap locality / alignment branches).  Remote operands are exactly
// log2(Q)/(log2(n)+1) of the reads (3/15 at n = 16384) and are fetched with coalesced ld.global.cg
// from the sibling's published state (L2 resident), staged in registers behind local taps.
//
// W / dW stream: physical layout [n_W][L][B][n'] where n' is the row axis in "quad-interleaved"
// order p(i) = q*J + 4*lq + lo (desc.interleave = Q): each CTA's rows of one tap are one contiguous
// 8 KB run.  W goes global -> registers directly (coalesced 512 B per warp and tap, read exactly
// once, prefetched into registers a few taps ahead - an L2 prefetch one pass ahead measured no gain); it does not pass
// through shared memory, which carries only the V/G operand traffic (the co-critical resource).
//
// State handling is the first generation's: the CTA's block of the current state sits in shared
// memory (channel-major, read-only during a factor), every new state goes to global memory
// (forward: these are the saved states of the backward pass) and is pulled back with TMA bulk
// copies at the factor boundary; one hardware cluster barrier per factor orders the exchange.
//
// Reference semantics: SyntheticExperiments/psf.py:167-188 (loop + residual), spmul_cuda.cu:24,
// :79-80, :105-108 (per-factor formulas).
#include <cstdlib>

#include "psf_internal.cuh"
#include "psf_ptx.cuh"

namespace {

// Launch shape (compile-time knobs): warps per CTA, CTAs per SM, W registers in flight (taps).  The forward pass
// fits 128 registers (8 warps x 2 CTAs per SM); the backward pass keeps two 32-register operand sets (accumulators
// and the stationary V_{m-1} rows) and runs 4 warps x 3 CTAs per SM at 168 registers (16 tiles = 4 even passes).
#ifndef PSF2_NW_FWD
#define PSF2_NW_FWD 8
#endif
#ifndef PSF2_MINB_FWD
#define PSF2_MINB_FWD 2
#endif
#ifndef PSF2_PFW_FWD
#define PSF2_PFW_FWD 4
#endif
#ifndef PSF2_NW_BWD
#define PSF2_NW_BWD 4
#endif
#ifndef PSF2_MINB_BWD
#define PSF2_MINB_BWD 3
#endif
#ifndef PSF2_PFW_BWD
#define PSF2_PFW_BWD 3
#endif
// (The tensor-memory variant of the backward residual accumulator measured slower - profiles/r01_ablation.md - and was removed.)
// timing experiments only (results become wrong): 1 no W loads, 2 no dW stores, 4 no residual traffic,
// 8 no V_{m-1} loads, 16 no remote operand loads, 32 no state stores, 64 no cluster barrier
#ifndef PSF2_ABLATE
#define PSF2_ABLATE 0
#endif
constexpr int kAbl = PSF2_ABLATE;
// per block size J = 2^LOGJ rows (J / 128 warp tiles): warps per CTA and CTAs per SM.  J = 2048 uses the tuned
// shapes above; smaller blocks (n < 2048: one CTA per sample) run one pass with one warp per tile and as many
// CTAs per SM as keep 16 (forward) / 12 (backward) warps resident.
// J = 4096 (n = 32768 / 65536: 128 KB of state, one CTA per SM) runs 16 / 8 warps.
__host__ __device__ constexpr int nw_fwd(int logj) { return logj >= 12 ? 16 : logj == 11 ? PSF2_NW_FWD : ((1 << logj) / 128 < 8 ? (1 << logj) / 128 : 8); }
__host__ __device__ constexpr int minb_fwd(int logj) { return logj >= 12 ? 1 : logj == 11 ? PSF2_MINB_FWD : 16 / nw_fwd(logj); }
__host__ __device__ constexpr int nw_bwd(int logj) { return logj >= 12 ? 8 : logj == 11 ? PSF2_NW_BWD : ((1 << logj) / 128 < 4 ? (1 << logj) / 128 : 4); }
__host__ __device__ constexpr int minb_bwd(int logj) { return logj >= 12 ? 1 : logj == 11 ? PSF2_MINB_BWD : 12 / nw_bwd(logj); }
constexpr int kTileRows = 128;       // rows per warp tile (32 lanes x one quad)

struct Chain2Params {
  const float* W;
  float* dW;
  const float* Vin;   // fwd: V0 [B,n,8];   bwd: dOut [B,n,8]  (natural layout)
  float* Vout;        // fwd: out [B,n,8];  bwd: dV0 [B,n,8]
  float* states;      // [slot][B][Q][8][J]
  float* gx;          // bwd exchange, per cluster [2][ncl][Q][8][J]
  float* racc;        // bwd residual accumulator, per cluster [ncl][Q][8][J]
  long long sWf, sWb, sWt;  // factor / sample / tap strides of W and dW (elements); row stride 1
  int B, n_W, residual, save_states;
  int G;              // 8-channel slabs per sample (C = 8 G): slab s of sample b is pseudo-sample b*G + s
  int dbg;
};

template <int V>
struct IC {
  static constexpr int value = V;
  constexpr operator int() const { return V; }
};
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
  if constexpr (I < N) {
    f(IC<I>{});
    static_for<I + 1, N>(f);
  }
}

// Compile-time geometry and tap schedule.
//   taps: k = 0 (offset 0), 1, 2 (offsets 1, 2), 3+r (remote r, quad offset 2^r), 3+NREM+t (local t,
//   element offset 4<<t inside the CTA's J rows).
//   R-quads: the NR = max(NREM,1) quads that are staged in registers: remote taps, or - without a cluster -
//   the local quad of tap 3, which the offset-1/2 taps need next to the thread's own quad.
//   Schedule of one pass: plain local taps are split into NR+2 groups; R(0) is issued after group 0,
//   consumed after group 1 together with taps 0,1,2 ("head"), R(r) is issued when R(r-1) has been consumed,
//   and when the last R-quad is consumed the same registers fetch the residual operand ("tail"), which is
//   used after the last group.
//   W registers are loaded in consumption order: idx -> tap k = k_at(idx).  The backward pass loads the
//   W values of taps 1,2 separately (they straddle two quads), so its sequence skips them.
template <int LOGJ, int LOGQ, bool BWD>
struct Geo {
  static constexpr int NW = BWD ? nw_bwd(LOGJ) : nw_fwd(LOGJ);
  static constexpr int NT = (1 << LOGJ) / kTileRows;  // warp tiles per CTA block
  static constexpr int J = 1 << LOGJ;
  static constexpr int Q = 1 << LOGQ;
  static constexpr int NQ = J / 4;  // local quads
  static constexpr int NLOC = LOGJ - 2;
  static constexpr int NREM = LOGQ;
  static constexpr int NPASS = (NT + NW - 1) / NW;
  static constexpr int NR = NREM > 0 ? NREM : 1;
  static constexpr int SKIP0 = NREM > 0 ? 0 : 1;  // local tap 0 is R(0) when there is no cluster
  static constexpr int P = NLOC - SKIP0;          // plain local taps
  static constexpr int NGRP = NR + 2;  // the last group covers the latency of the tail loads (residual operand)
  static constexpr int HEADN = BWD ? 2 : 4;
  static constexpr int SEQ = P + HEADN + (NR - 1);
  __host__ __device__ static constexpr int plo(int g) { return g * P / NGRP; }
  __host__ __device__ static constexpr int kloc(int p) { return 3 + NREM + p + SKIP0; }
  __host__ __device__ static constexpr int headk(int h) { return BWD ? (h == 0 ? 3 : 0) : h; }
  __host__ __device__ static constexpr int k_at(int idx) {
    int s = 0;
    for (int g = 0; g < NGRP; ++g) {
      for (int pp = plo(g); pp < plo(g + 1); ++pp) {
        if (s == idx) return kloc(pp);
        ++s;
      }
      if (g == 1) {
        for (int h = 0; h < HEADN; ++h) {
          if (s == idx) return headk(h);
          ++s;
        }
      } else if (g >= 2 && g <= NR) {
        if (s == idx) return 3 + (g - 1);
        ++s;
      }
    }
    return -1;
  }
  __host__ __device__ static constexpr int idx_of(int k) {
    for (int i = 0; i < SEQ; ++i)
      if (k_at(i) == k) return i;
    return -1;
  }
};

__device__ __forceinline__ float4 ldg_w4(const float* p) {  // W: read once (streaming: evict-first in L1 and L2)
  float4 v;
  if constexpr (kAbl & 1) return make_float4(0.5f, 0.25f, 0.125f, 1.f);
  asm volatile("ld.global.cs.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ldg_w2(const float* p) {
  float2 v;
  asm volatile("ld.global.cs.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float ldg_w1(const float* p) {
  float v;
  asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// exchanged state written by sibling CTAs during this kernel: L2-coherent load, pinned in program order
__device__ __forceinline__ float4 ldcg4_v(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
// exchange buffers of the backward pass (reused by the cluster sample after sample): keep them in L2
__device__ __forceinline__ unsigned long long make_evict_last_policy() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ float4 ldcg4_keep(const float* p, unsigned long long pol) {
  float4 v;
  asm volatile("ld.global.cg.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p), "l"(pol)
               : "memory");
  return v;
}
__device__ __forceinline__ void stcg4_keep(float* p, const float4& v, unsigned long long pol) {
  asm volatile("st.global.cg.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w), "l"(pol)
               : "memory");
}


__device__ __forceinline__ void red_add4g(float* p, const float4& v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void red_add1g(float* p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void st_cs2(float* p, float a, float b) { __stcs(reinterpret_cast<float2*>(p), make_float2(a, b)); }

// this CTA's block of a published state -> shared memory (8 channel planes of J floats, contiguous)
template <int J>
__device__ __forceinline__ void refresh_block(uint32_t state_u32, uint32_t vbar, const float* blk) {
  mbar_expect_tx(vbar, (uint32_t)(CT * J * 4));
#pragma unroll
  for (int ch = 0; ch < CT; ++ch) tma_load_1d(state_u32 + (uint32_t)(ch * J) * 4u, blk + ch * J, (uint32_t)J * 4u, vbar);
}

// ---------------------------------------------------------------------------------------------
// Forward
// ---------------------------------------------------------------------------------------------
template <int LOGJ, int LOGQ>
__global__ void __launch_bounds__(nw_fwd(LOGJ) * 32, minb_fwd(LOGJ)) psf_chain2_fwd_kernel(const Chain2Params p) {
  using G = Geo<LOGJ, LOGQ, false>;
  constexpr int J = G::J, Q = G::Q, NQ = G::NQ, kNW = G::NW;
  constexpr int PFW = PSF2_PFW_FWD;  // W registers in flight (taps)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* state = reinterpret_cast<float*>(smem_raw);
  const uint32_t state_u32 = smem_u32(state);
  const uint32_t vbar = state_u32 + (uint32_t)(CT * J) * 4u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x & (Q - 1);
  const int cid = blockIdx.x >> LOGQ;
  const int ncl = gridDim.x >> LOGQ;
  constexpr int n = J * Q;
  constexpr long long plane = (long long)CT * n;  // floats per sample per state
  constexpr int blk = CT * J;                     // floats per CTA block

  if (threadIdx.x == 0) {
    mbar_init(vbar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  uint32_t vparity = 0;
  bool cw_pending = false;

  const int NP = p.B * p.G;           // pseudo-samples (sample, 8-channel slab); states are indexed by them
  const int rs = CT * p.G;            // floats per row of V0 / out
  for (int bp = cid; bp < NP; bp += ncl) {
    const int b = bp / p.G;           // W belongs to the sample
    const int c0 = (bp - b * p.G) * CT;
    float* st0 = p.states + (long long)bp * plane;  // slot 0 = V_0 (residual operand, saved state 0)
    // ---- prologue: V0 rows of this CTA (natural layout) -> state slot 0 ------------------------
#pragma unroll 1
    for (int pp = 0; pp < G::NPASS; ++pp) {
      if (pp * kNW + warp >= G::NT) continue;
      const int lq = (pp * kNW + warp) * 32 + lane;
      const int j = 4 * lq;
      const int i0 = 4 * (lq * Q + q);
      float4 pl[CT];
      load_rows_strided(p.Vin + ((long long)b * n + i0) * rs + c0, rs, pl);
#pragma unroll
      for (int ch = 0; ch < CT; ++ch) st_cg4(st0 + q * blk + ch * J + j, pl[ch]);
    }
    fence_proxy_async();
    __syncthreads();  // the block is written; nobody reads the shared copy any more
    if (threadIdx.x == 0) refresh_block<J>(state_u32, vbar, st0 + q * blk);
    if (Q > 1 && !(kAbl & 64)) {
      cluster_arrive();
      cw_pending = true;
    }

    for (int m = 1; m <= p.n_W; ++m) {
      const int slot_prev = (p.save_states || m == 1) ? (m - 1) : 1 + ((m - 1) & 1);
      const int slot_next = p.save_states ? m : 1 + (m & 1);
      const float* prev = p.states + ((long long)slot_prev * NP + bp) * plane;
      float* next = p.states + ((long long)slot_next * NP + bp) * plane;
      const bool last = (m == p.n_W);
      const float* Wmb = p.W + (long long)(m - 1) * p.sWf + (long long)b * p.sWb + q * J;

      mbar_wait(vbar, vparity);  // V_{m-1} block has landed in shared memory
      vparity ^= 1u;

#pragma unroll 1
      for (int pp = 0; pp < G::NPASS; ++pp) {
        const int tile = pp * kNW + warp;
        if (tile >= G::NT) continue;  // warp-uniform; pass 0 is full, so every warp executes the cluster wait
        const int lq = tile * 32 + lane;
        const int j = 4 * lq;
        const float* wrow = Wmb + j;

        float4 acc[CT];
        float4 R[CT];
        float4 wr[PFW];
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) acc[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
        static_for<0, PFW>([&](auto I) {
          constexpr int i = decltype(I)::value;
          wr[i] = ldg_w4(wrow + (long long)G::k_at(i) * p.sWt);
        });
        // take the W registers of sequence position idx and refill the slot PFW taps ahead
        auto take_w = [&](auto I) -> float4 {
          constexpr int idx = decltype(I)::value;
          const float4 w = wr[idx % PFW];
          if constexpr (idx + PFW < G::SEQ) wr[idx % PFW] = ldg_w4(wrow + (long long)G::k_at(idx + PFW) * p.sWt);
          return w;
        };
        auto issue_R = [&](auto Rr) {
          if constexpr (G::NREM > 0) {
            constexpr int x = 1 << decltype(Rr)::value;
            const int qp = (q + x) & (Q - 1);
            const int lqp = (lq + ((q + x) >> LOGQ)) & (NQ - 1);
            const float* src = prev + qp * blk + 4 * lqp;
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) R[ch] = (kAbl & 16) ? make_float4(1.f, 2.f, 3.f, 4.f) : ldcg4_v(src + ch * J);
          }
        };

        static_for<0, G::NGRP>([&](auto Gi) {
          constexpr int g = decltype(Gi)::value;
          // ---- plain local taps of this group -----------------------------------------------
          static_for<G::plo(g), G::plo(g + 1)>([&](auto Pi) {
            constexpr int pidx = decltype(Pi)::value;
            constexpr int k = G::kloc(pidx);
            constexpr int d = 4 << (k - 3 - G::NREM);
            const float4 w = take_w(IC<G::idx_of(k)>{});
            const float* vp = state + ((j + d) & (J - 1));
            float4 v[CT];
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) v[ch] = lds4(vp + ch * J);
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) fma4(acc[ch], w, v[ch]);
          });
          auto issue_tail = [&]() {  // own rows of V_0 (residual): written by this very thread in the prologue
            if (p.residual && !(kAbl & 4)) {
#pragma unroll
              for (int ch = 0; ch < CT; ++ch) R[ch] = ldcg4_v(st0 + q * blk + ch * J + j);
            }
          };
          if constexpr (g == 0) {
            if (Q > 1 && cw_pending && !(kAbl & 64)) {  // siblings have published V_{m-1}
              cluster_wait();
              cw_pending = false;
            }
            issue_R(IC<0>{});
          } else if constexpr (g == G::NR + 1) {
            if (p.residual) {
#pragma unroll
              for (int ch = 0; ch < CT; ++ch) add4(acc[ch], R[ch]);
            }
          } else if constexpr (g == 1) {
            // ---- head: offsets 0, 1, 2 and tap 3 share the quads A = own rows, Rq = next quad ----
            const float4 w0 = take_w(IC<G::idx_of(0)>{});
            const float4 w1 = take_w(IC<G::idx_of(1)>{});
            const float4 w2 = take_w(IC<G::idx_of(2)>{});
            const float4 w3 = take_w(IC<G::idx_of(3)>{});
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) {
              const float4 A = lds4(state + ch * J + j);
              float4 Rq;
              if constexpr (G::NREM > 0) Rq = R[ch]; else Rq = lds4(state + ch * J + ((j + 4) & (J - 1)));
              // row r: w0[r]*V[r] + w1[r]*V[r+1] + w2[r]*V[r+2] + w3[r]*V[r+4]
              fma4(acc[ch], w0, A);
              fma4(acc[ch], w3, Rq);
              fma2(acc[ch].x, acc[ch].y, w2.x, w2.y, A.z, A.w);
              fma2(acc[ch].z, acc[ch].w, w2.z, w2.w, Rq.x, Rq.y);
              acc[ch].x = fmaf(w1.x, A.y, acc[ch].x);
              acc[ch].y = fmaf(w1.y, A.z, acc[ch].y);
              acc[ch].z = fmaf(w1.z, A.w, acc[ch].z);
              acc[ch].w = fmaf(w1.w, Rq.x, acc[ch].w);
            }
            if constexpr (1 < G::NR) issue_R(IC<1>{}); else issue_tail();
          } else {
            constexpr int r = g - 1;
            const float4 w = take_w(IC<G::idx_of(3 + r)>{});
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) fma4(acc[ch], w, R[ch]);
            if constexpr (g < G::NR) issue_R(IC<g>{}); else issue_tail();
          }
        });

        if (!last) {
#pragma unroll
          for (int ch = 0; ch < CT; ++ch)
            if (!(kAbl & 32) || acc[ch].x == 123.456f) st_cg4(next + q * blk + ch * J + j, acc[ch]);
        } else {
          store_rows_strided(p.Vout + ((long long)b * n + 4 * (lq * Q + q)) * rs + c0, rs, acc);
        }
      }
      if (!last) {
        fence_proxy_async();
        __syncthreads();  // all reads of the shared block and all writes of V_m are done
        if (threadIdx.x == 0) refresh_block<J>(state_u32, vbar, next + q * blk);
        if (Q > 1 && !(kAbl & 64)) {
          cluster_arrive();  // V_m visible cluster-wide once the (deferred) wait returns
          cw_pending = true;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Backward (output-stationary at row j):
//   dV_{m-1}[j,:] += W_m[j-o,k] * G_m[j-o,:]      dW_m[j-o,k] = <G_m[j-o,:], V_{m-1}[j,:]>
// one shifted read of G feeds both; V_{m-1}[j,:] (saved by the forward pass) sits in registers.
// ---------------------------------------------------------------------------------------------
template <int LOGJ, int LOGQ, bool kMulti>
__global__ void __launch_bounds__(nw_bwd(LOGJ) * 32, minb_bwd(LOGJ)) psf_chain2_bwd_kernel(const Chain2Params p) {
  using G = Geo<LOGJ, LOGQ, true>;
  constexpr int J = G::J, Q = G::Q, NQ = G::NQ, kNW = G::NW;
  constexpr int PFW = PSF2_PFW_BWD;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* state = reinterpret_cast<float*>(smem_raw);
  const uint32_t state_u32 = smem_u32(state);
  const uint32_t vbar = state_u32 + (uint32_t)(CT * J) * 4u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x & (Q - 1);
  const int cid = blockIdx.x >> LOGQ;
  const int ncl = gridDim.x >> LOGQ;
  constexpr int n = J * Q;
  constexpr long long plane = (long long)CT * n;
  constexpr int blk = CT * J;

  if (threadIdx.x == 0) {
    mbar_init(vbar, 1);
    fence_barrier_init();
  }
  __syncthreads();
  uint32_t vparity = 0;
  bool cw_pending = false;

  // exchange / residual buffers are per CLUSTER (reused sample after sample) so that they stay in L2
  const unsigned long long keep = make_evict_last_policy();
  float* gx0 = p.gx + (long long)cid * plane;
  float* gx1 = p.gx + ((long long)ncl + cid) * plane;
  float* racc = p.racc + (long long)cid * plane + q * blk;

  const int NP = p.B * p.G;           // pseudo-samples (sample, 8-channel slab)
  const int rs = CT * p.G;            // floats per row of dOut / dV0
  constexpr bool multi = kMulti;      // G > 1: several slabs add into the same dW: red.add on a pre-zeroed buffer
  for (int bp = cid; bp < NP; bp += ncl) {
    const int b = bp / p.G;
    const int c0 = (bp - b * p.G) * CT;
    // ---- prologue: dOut rows of this CTA -> exchange slot 0 ------------------------------------
#pragma unroll 1
    for (int pp = 0; pp < G::NPASS; ++pp) {
      if (pp * kNW + warp >= G::NT) continue;
      const int lq = (pp * kNW + warp) * 32 + lane;
      const int j = 4 * lq;
      const int i0 = 4 * (lq * Q + q);
      float4 pl[CT];
      load_rows_strided(p.Vin + ((long long)b * n + i0) * rs + c0, rs, pl);
#pragma unroll
      for (int ch = 0; ch < CT; ++ch) stcg4_keep(gx0 + q * blk + ch * J + j, pl[ch], keep);
    }
    fence_proxy_async();
    __syncthreads();
    if (threadIdx.x == 0) refresh_block<J>(state_u32, vbar, gx0 + q * blk);
    if (Q > 1 && !(kAbl & 64)) {
      cluster_arrive();
      cw_pending = true;
    }

    for (int m = p.n_W; m >= 1; --m) {
      const int step = p.n_W - m;
      const float* gcur = (step & 1) ? gx1 : gx0;
      float* gnext = (step & 1) ? gx0 : gx1;
      const float* vprev = p.states + ((long long)(m - 1) * NP + bp) * plane + q * blk;  // V_{m-1}, own block
      const float* Wmb = p.W + (long long)(m - 1) * p.sWf + (long long)b * p.sWb;         // sample base
      float* dWmb = p.dW + (long long)(m - 1) * p.sWf + (long long)b * p.sWb;
      const bool last = (m == 1);
      const bool first = (m == p.n_W);
      const bool even_step = !(step & 1);

      mbar_wait(vbar, vparity);  // G_m block has landed in shared memory
      vparity ^= 1u;

#pragma unroll 1
      for (int pp = 0; pp < G::NPASS; ++pp) {
        const int tile = pp * kNW + warp;
        if (tile >= G::NT) continue;  // warp-uniform; pass 0 is full, so every warp executes the cluster wait
        const int lq = tile * 32 + lane;
        const int j = 4 * lq;

        // position (inside one tap plane of the sample) of the W / dW quad that tap k pairs with row quad lq
        auto wpos = [&](auto K) -> int {
          constexpr int k = decltype(K)::value;
          if constexpr (k < 3) {
            return q * J + j;
          } else if constexpr (k < 3 + G::NREM) {
            constexpr int x = 1 << (k - 3);
            const int qp = (q - x) & (Q - 1);
            const int lqp = (lq - (q < x ? 1 : 0)) & (NQ - 1);
            return qp * J + 4 * lqp;
          } else {
            constexpr int d = 4 << (k - 3 - G::NREM);
            return q * J + ((j - d) & (J - 1));
          }
        };

        float4 acc[CT];
        float4 vown[CT];
        float4 R[CT];
        float4 wr[PFW];
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) acc[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int ch = 0; ch < CT; ++ch)
          vown[ch] = (kAbl & 8) ? make_float4(1.f, 2.f, 3.f, 4.f) : __ldcs(reinterpret_cast<const float4*>(vprev + ch * J + j));
        static_for<0, PFW>([&](auto I) {
          constexpr int i = decltype(I)::value;
          constexpr int k = G::k_at(i);
          wr[i] = ldg_w4(Wmb + (long long)k * p.sWt + wpos(IC<k>{}));
        });
        auto take_w = [&](auto I) -> float4 {
          constexpr int idx = decltype(I)::value;
          const float4 w = wr[idx % PFW];
          if constexpr (idx + PFW < G::SEQ) {
            constexpr int kn = G::k_at(idx + PFW);
            wr[idx % PFW] = ldg_w4(Wmb + (long long)kn * p.sWt + wpos(IC<kn>{}));
          }
          return w;
        };
        // W of taps 1, 2 straddles the own quad (rows lo >= offset) and the previous quad
        float4 w1a;
        float2 w2a, w2r;
        float w1r;
        const int pos_own = q * J + j;
        const int pos_r0 = wpos(IC<3>{});  // quad u-1: remote CTA q-1 (or local quad lq-1 without a cluster)
        auto issue_R = [&](auto Rr) {
          constexpr int r = decltype(Rr)::value;
          if constexpr (G::NREM > 0) {
            constexpr int x = 1 << r;
            const int qp = (q - x) & (Q - 1);
            const int lqp = (lq - (q < x ? 1 : 0)) & (NQ - 1);
            const float* src = gcur + qp * blk + 4 * lqp;
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) R[ch] = (kAbl & 16) ? make_float4(1.f, 2.f, 3.f, 4.f) : ldcg4_keep(src + ch * J, keep);
          }
          if constexpr (r == 0) {
            w1a = ldg_w4(Wmb + p.sWt + pos_own);
            w1r = ldg_w1(Wmb + p.sWt + pos_r0 + 3);
            w2a = ldg_w2(Wmb + 2 * p.sWt + pos_own);
            w2r = ldg_w2(Wmb + 2 * p.sWt + pos_r0 + 2);
          }
        };
        // one aligned tap: operand quads g[ch], W quad w, dW quad stored at `pos` of tap k
        auto tap = [&](const float4 (&g)[CT], const float4& w, int k, int pos) {
          float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int ch = 0; ch < CT; ++ch) {
            fma4(acc[ch], w, g[ch]);
            fma4(dw, g[ch], vown[ch]);
          }
          if constexpr (!(kAbl & 2)) {
            if constexpr (multi) red_add4g(dWmb + (long long)k * p.sWt + pos, dw); else st_cs4(dWmb + (long long)k * p.sWt + pos, dw);
          } else if (dw.x == 123.456f) st_cs4(dWmb, dw);
        };

        static_for<0, G::NGRP>([&](auto Gi) {
          constexpr int g = decltype(Gi)::value;
          static_for<G::plo(g), G::plo(g + 1)>([&](auto Pi) {
            constexpr int pidx = decltype(Pi)::value;
            constexpr int k = G::kloc(pidx);
            constexpr int d = 4 << (k - 3 - G::NREM);
            const float4 w = take_w(IC<G::idx_of(k)>{});
            const int js = (j - d) & (J - 1);
            float4 v[CT];
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) v[ch] = lds4(state + ch * J + js);
            tap(v, w, k, q * J + js);
          });
          auto issue_tail = [&]() {  // running sum of G_m over the factors (residual): this thread's own rows
            // the buffer is touched every OTHER factor: an even step adds both G_m (the state) and G_{m-1} (the
            // accumulators it has just finished), so the odd step in between has nothing left to add
            if (p.residual && ((even_step && !first) || (!even_step && last)) && !(kAbl & 4)) {
#pragma unroll
              for (int ch = 0; ch < CT; ++ch) R[ch] = ldcg4_keep(racc + ch * J + j, keep);
            }
          };
          if constexpr (g == 0) {
            if (Q > 1 && cw_pending && !(kAbl & 64)) {
              cluster_wait();
              cw_pending = false;
            }
            issue_R(IC<0>{});
          } else if constexpr (g == G::NR + 1) {
            // residual: dV0 also receives G_m of every factor (V_m = W_m V_{m-1} + V_0)
            if (p.residual && even_step) {
#pragma unroll
              for (int ch = 0; ch < CT; ++ch) {
                float4 own = lds4(state + ch * J + j);      // G_m
                if (!first) add4(own, R[ch]);               // + the pairs before
                if (!last) {
                  add4(own, acc[ch]);                       // + G_{m-1}
                  if (!(kAbl & 4)) stcg4_keep(racc + ch * J + j, own, keep);
                } else {
                  add4(acc[ch], own);                       // dV0 = G_0 + G_1 + pairs
                }
              }
            } else if (p.residual && last) {                // odd last step: G_1 was added one step ago
#pragma unroll
              for (int ch = 0; ch < CT; ++ch) add4(acc[ch], R[ch]);
            }
          } else if constexpr (g == 1) {
            // ---- head: tap 3 on the previous quad Rq, then taps 0, 1, 2 on (Rq, own quad A) ----
            const float4 w3 = take_w(IC<G::idx_of(3)>{});
            const float4 w0 = take_w(IC<G::idx_of(0)>{});
            if constexpr (G::NREM == 0) {
#pragma unroll
              for (int ch = 0; ch < CT; ++ch) R[ch] = lds4(state + ch * J + ((j - 4) & (J - 1)));
            }
            tap(R, w3, 3, pos_r0);
            const float4 w1 = make_float4(w1r, w1a.x, w1a.y, w1a.z);
            const float4 w2 = make_float4(w2r.x, w2r.y, w2a.x, w2a.y);
            float4 dw0 = make_float4(0.f, 0.f, 0.f, 0.f), dw1 = dw0, dw2 = dw0;
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) {
              const float4 A = lds4(state + ch * J + j);
              const float4 g1 = make_float4(R[ch].w, A.x, A.y, A.z);   // rows j-1 .. j+2
              const float4 g2 = make_float4(R[ch].z, R[ch].w, A.x, A.y);  // rows j-2 .. j+1
              fma4(acc[ch], w0, A);
              fma4(dw0, A, vown[ch]);
              fma4(acc[ch], w1, g1);
              fma4(dw1, g1, vown[ch]);
              fma4(acc[ch], w2, g2);
              fma4(dw2, g2, vown[ch]);
            }
            float* d0 = dWmb + pos_own;
            if constexpr (kAbl & 2) {
              if (dw0.x + dw1.x + dw2.x == 123.456f) st_cs4(d0, dw0);
            } else {
            float* d1 = dWmb + p.sWt;
            float* d2 = dWmb + 2 * p.sWt;
            if constexpr (multi) {
              red_add4g(d0, dw0);
              red_add1g(d1 + pos_r0 + 3, dw1.x);
              red_add1g(d1 + pos_own, dw1.y);
              red_add1g(d1 + pos_own + 1, dw1.z);
              red_add1g(d1 + pos_own + 2, dw1.w);
              red_add1g(d2 + pos_r0 + 2, dw2.x);
              red_add1g(d2 + pos_r0 + 3, dw2.y);
              red_add1g(d2 + pos_own, dw2.z);
              red_add1g(d2 + pos_own + 1, dw2.w);
            } else {
              st_cs4(d0, dw0);
              __stcs(d1 + pos_r0 + 3, dw1.x);
              st_cs2(d1 + pos_own, dw1.y, dw1.z);
              __stcs(d1 + pos_own + 2, dw1.w);
              st_cs2(d2 + pos_r0 + 2, dw2.x, dw2.y);
              st_cs2(d2 + pos_own, dw2.z, dw2.w);
            }
            }
            if constexpr (1 < G::NR) issue_R(IC<1>{}); else issue_tail();
          } else {
            constexpr int r = g - 1;
            const float4 w = take_w(IC<G::idx_of(3 + r)>{});
            tap(R, w, 3 + r, wpos(IC<3 + r>{}));
            if constexpr (g < G::NR) issue_R(IC<g>{}); else issue_tail();
          }
        });

        if (!last) {
#pragma unroll
          for (int ch = 0; ch < CT; ++ch)
            if (!(kAbl & 32) || acc[ch].x == 123.456f) stcg4_keep(gnext + q * blk + ch * J + j, acc[ch], keep);
        } else {
          store_rows_strided(p.Vout + ((long long)b * n + 4 * (lq * Q + q)) * rs + c0, rs, acc);
        }
      }
      if (!last) {
        fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) refresh_block<J>(state_u32, vbar, gnext + q * blk);
        if (Q > 1 && !(kAbl & 64)) {
          cluster_arrive();
          cw_pending = true;
        }
      }
    }
    // the exchange buffers belong to the cluster: the next sample may overwrite them only after every
    // CTA has finished reading this sample's last state
    if (Q > 1 && !(kAbl & 64)) {
      cluster_arrive();
      cluster_wait();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
constexpr int kLogJ = 11;      // block size of the clustered shapes n = 2048 .. 16384
constexpr int kLogJMin = 7;    // smallest single-CTA shape (n = 128)
constexpr int kLogJBig = 12;   // n = 32768 (8 CTAs) and n = 65536 (16 CTAs: non-portable cluster size)

int env_int2(const char* name, int dflt) {
  const char* v = std::getenv(name);
  if (v == nullptr || *v == 0) return dflt;
  return std::atoi(v);
}

// kernel table: (logj, logq) with logq > 0 only for logj == kLogJ
#define PSF2_KERNEL_CASES(K)                                            \
  if (logj == 7) return (const void*)K<7, 0>;                           \
  if (logj == 8) return (const void*)K<8, 0>;                           \
  if (logj == 9) return (const void*)K<9, 0>;                           \
  if (logj == 10) return (const void*)K<10, 0>;                         \
  if (logj == 12) return logq == 3 ? (const void*)K<12, 3> : (const void*)K<12, 4>; \
  switch (logq) {                                                       \
    case 0: return (const void*)K<kLogJ, 0>;                            \
    case 1: return (const void*)K<kLogJ, 1>;                            \
    case 2: return (const void*)K<kLogJ, 2>;                            \
    default: return (const void*)K<kLogJ, 3>;                           \
  }
template <int LJ, int LQ>
struct Bwd1 { static const void* get() { return (const void*)psf_chain2_bwd_kernel<LJ, LQ, false>; } };
template <int LJ, int LQ>
struct BwdM { static const void* get() { return (const void*)psf_chain2_bwd_kernel<LJ, LQ, true>; } };
#define PSF2_STRUCT_CASES(S)                                            \
  if (logj == 7) return S<7, 0>::get();                                 \
  if (logj == 8) return S<8, 0>::get();                                 \
  if (logj == 9) return S<9, 0>::get();                                 \
  if (logj == 10) return S<10, 0>::get();                               \
  if (logj == 12) return logq == 3 ? S<12, 3>::get() : S<12, 4>::get(); \
  switch (logq) {                                                       \
    case 0: return S<kLogJ, 0>::get();                                  \
    case 1: return S<kLogJ, 1>::get();                                  \
    case 2: return S<kLogJ, 2>::get();                                  \
    default: return S<kLogJ, 3>::get();                                 \
  }
const void* fwd2_kernel_for(int logj, int logq) { PSF2_KERNEL_CASES(psf_chain2_fwd_kernel) }
const void* bwd2_kernel_for(int logj, int logq, bool multi = false) {
  if (multi) { PSF2_STRUCT_CASES(BwdM) }
  PSF2_STRUCT_CASES(Bwd1)
}

int smem2(int logj) { return CT * (1 << logj) * 4 + 64; }
int threads2(bool bwd, int logj) { return (bwd ? nw_bwd(logj) : nw_fwd(logj)) * 32; }

std::mutex g_mu2;
bool g_attr2_done[PSF_MAX_DEVICES] = {};
int g_res2[PSF_MAX_DEVICES][2][10] = {};  // [device][bwd][slot]: slots 0..3 = logj 7..10, 4..7 = logj 11 with logq 0..3, 8..9 = logj 12

int resident2(bool bwd, int logj, int logq) {   // caller holds g_mu2
  const int dev = psf_current_device();
  if (dev < 0) return 0;
  int& slot = g_res2[dev][bwd ? 1 : 0][logj < kLogJ ? logj - kLogJMin : (logj == kLogJ ? 4 + logq : 8 + (logq - 3))];
  if (slot > 0) return slot;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const void* kernel = bwd ? bwd2_kernel_for(logj, logq) : fwd2_kernel_for(logj, logq);
  const int Q = 1 << logq;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(Q * sms * 32), 1, 1);
  cfg.blockDim = dim3((unsigned)threads2(bwd, logj), 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem2(logj);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)Q;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int ncl = 0;
  if (Q > 8 && (cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
                (bwd && cudaFuncSetAttribute(bwd2_kernel_for(logj, logq, true), cudaFuncAttributeNonPortableClusterSizeAllowed,
                                             1) != cudaSuccess))) {
    cudaGetLastError();
    return 0;
  }
  if (cudaOccupancyMaxActiveClusters(&ncl, kernel, &cfg) != cudaSuccess || ncl < 1) {
    cudaGetLastError();
    return 0;
  }
  slot = ncl;
  return ncl;
}

int launch2(const void* kernel, const Chain2Params& cp, int logj, int logq, int resident, int threads, cudaStream_t st) {
  const int Q = 1 << logq;
  cudaLaunchConfig_t cfg = {};
  const int clusters = resident < cp.B * cp.G ? resident : cp.B * cp.G;
  cfg.gridDim = dim3((unsigned)(clusters * Q), 1, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem2(logj);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)Q;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  void* args[] = {(void*)&cp};
  return (int)cudaLaunchKernelExC(&cfg, kernel, args);
}

void fill2(Chain2Params& cp, const PsfChainArgs& a) {
  cp.sWf = a.ws.factor;
  cp.sWb = a.ws.sample;
  cp.sWt = a.ws.tap;
  cp.B = (int)a.B;
  cp.G = (int)(a.C / CT);
  cp.n_W = a.n_W;
  cp.residual = a.residual;
  cp.dbg = env_int2("PSF_CHAIN2_DEBUG", 0);
}

}  // namespace

// Row interleave (cluster size Q) the second-generation kernels want for this problem, or 0 if they do
// not cover it.  Shape rules only; layout/pointer checks are psf_chain2_plan's.
int psf_chain2_interleave_for(const PsfChainArgs& a) {
  if (env_int2("PSF_CHAIN_V2", 1) == 0) return 0;
  // up to 65536 channels: the dense attention map (chain applied to the n x n identity) is C = n
  if (a.C < CT || a.C % CT != 0 || a.C > 65536 || a.n_W < 1 || a.B < 1 || a.B * (a.C / CT) > (1 << 20)) return 0;
  int logn = 0;
  while ((int64_t(1) << logn) < a.n) ++logn;
  if ((int64_t(1) << logn) != a.n || logn < kLogJMin || logn > kLogJBig + 4) return 0;
  if (a.L != logn + 1) return 0;
  for (int k = 0; k < a.L; ++k)
    if (a.offs.o[k] != (k == 0 ? 0 : (1 << (k - 1)))) return 0;
  return logn <= kLogJ ? 1 : (logn <= kLogJ + 3 ? 1 << (logn - kLogJ) : 1 << (logn - kLogJBig));
}

static void geom2(const PsfChainArgs& a, int Q, int* logj, int* logq) {
  int logn = 0;
  while ((int64_t(1) << logn) < a.n) ++logn;
  *logq = 0;
  while ((1 << *logq) < Q) ++*logq;
  *logj = logn - *logq;
}

int psf_chain2_plan(const PsfChainArgs& a, PsfFusedPlan* plan) {
  const int Q = psf_chain2_interleave_for(a);
  if (Q == 0) return PSF_ERR_UNSUPPORTED;
  const int il = a.interleave < 1 ? 1 : a.interleave;
  if (il != Q) return PSF_ERR_UNSUPPORTED;
  if (a.ws.row != 1 || (a.ws.factor & 3) || (a.ws.sample & 3) || (a.ws.tap & 3)) return PSF_ERR_UNSUPPORTED;
  // C > 8: the slabs of a sample add into one dW, which is zeroed first as ONE contiguous [n_W][L][B][n] block
  if (a.C > CT && (a.ws.sample != a.n || a.ws.tap != a.B * a.n || a.ws.factor != (int64_t)a.L * a.B * a.n))
    return PSF_ERR_UNSUPPORTED;
  int logj, logq;
  geom2(a, Q, &logj, &logq);
  const int dev = psf_current_device();
  if (dev < 0) return PSF_ERR_UNSUPPORTED;
  std::lock_guard<std::mutex> lock(g_mu2);
  if (!g_attr2_done[dev]) {
    for (int lj = kLogJMin; lj <= kLogJBig; ++lj)
      for (int lq = (lj == kLogJBig ? 3 : 0); lq < (lj == kLogJBig ? 5 : (lj == kLogJ ? 4 : 1)); ++lq)
        if (cudaFuncSetAttribute(fwd2_kernel_for(lj, lq), cudaFuncAttributeMaxDynamicSharedMemorySize, smem2(lj)) !=
                cudaSuccess ||
            cudaFuncSetAttribute(bwd2_kernel_for(lj, lq, false), cudaFuncAttributeMaxDynamicSharedMemorySize, smem2(lj)) !=
                cudaSuccess ||
            cudaFuncSetAttribute(bwd2_kernel_for(lj, lq, true), cudaFuncAttributeMaxDynamicSharedMemorySize, smem2(lj)) !=
                cudaSuccess) {
          cudaGetLastError();
          return PSF_ERR_UNSUPPORTED;
        }
    g_attr2_done[dev] = true;
  }
  const int rf = resident2(false, logj, logq), rb = resident2(true, logj, logq);
  if (rf < 1 || rb < 1) return PSF_ERR_UNSUPPORTED;
  plan->cluster = Q;
  plan->rows_per_cta = 1 << logj;
  plan->threads = threads2(false, logj);
  plan->smem_bytes = smem2(logj);
  plan->smem_bytes_bwd = smem2(logj);
  plan->resident = rf;
  plan->resident_bwd = rb;
  plan->stages = 0;
  plan->stages_bwd = 0;
  return PSF_OK;
}

// Workspace (floats, S = B*n*8):  states [0, n_W*S) (save) or [0, 3*S);  then gx 2*S, racc S (backward only).
int psf_chain2_fwd(const PsfChainArgs& a, const float* W, const float* V0, float* out, float* ws, int save_states,
                   cudaStream_t st) {
  PsfFusedPlan plan;
  int rc = psf_chain2_plan(a, &plan);
  if (rc) return rc;
  int logj, logq;
  geom2(a, plan.cluster, &logj, &logq);
  Chain2Params cp = {};
  cp.W = W;
  cp.Vin = V0;
  cp.Vout = out;
  cp.states = ws;
  cp.save_states = save_states;
  fill2(cp, a);
  return launch2(fwd2_kernel_for(logj, logq), cp, logj, logq, plan.resident, threads2(false, logj), st);
}

int psf_chain2_bwd(const PsfChainArgs& a, const float* W, const float* dOut, float* dW, float* dV0, float* ws,
                   cudaStream_t st) {
  PsfFusedPlan plan;
  int rc = psf_chain2_plan(a, &plan);
  if (rc) return rc;
  int logj, logq;
  geom2(a, plan.cluster, &logj, &logq);
  const long long S = a.B * a.n * a.C;
  Chain2Params cp = {};
  cp.W = W;
  cp.dW = dW;
  cp.Vin = dOut;
  cp.Vout = dV0;
  cp.states = ws;
  cp.gx = ws + (long long)a.n_W * S;
  cp.racc = cp.gx + 2 * S;
  cp.save_states = 1;
  fill2(cp, a);
  if (cp.G > 1) PSF_CUDA_TRY(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)a.n_W * a.L * a.B * a.n, st));
  return launch2(bwd2_kernel_for(logj, logq, cp.G > 1), cp, logj, logq, plan.resident_bwd, threads2(true, logj), st);
}
