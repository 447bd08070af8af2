// Fused chord-factor chain for sm_100a (B200): forward and backward, fp32, C = 8 channels.
//
// One thread-block CLUSTER owns one sample.  CTA q of a P-CTA cluster owns the contiguous row
// block [q*RB, (q+1)*RB) for ALL channels.  Structure of one CTA:
//
//   * STATE CACHE.  The current state of the block (V_{m-1} forward, G_m backward) sits in shared
//     memory, channel-major [8][plane] fp32 (64 KB for 2048 rows), and is READ-ONLY during a factor.
//     Every new state is written to global memory (it is L2 resident; forward states double as the
//     saved V_0..V_{n_W-1} of the backward pass), and at the factor boundary the CTA pulls its own
//     block back into shared memory with eight 8 KB TMA bulk copies.  No in-place hazard, so no
//     barrier inside a factor; two CTAs (different samples) share an SM and cover each other's
//     factor-boundary latency.
//   * EXCHANGE.  Operands outside the CTA's block (big taps, halo of small taps, cyclic wrap) are
//     read straight from the previous state in global memory (written by the sibling CTAs one
//     factor earlier); one hardware cluster barrier (release/acquire) per factor orders it.
//   * W STREAM.  W is read exactly once, tap-major ([..][tap][row], row stride 1): the rows a pass
//     covers (warps x 128) are one contiguous run per tap, fetched with cp.async.bulk (TMA 1-D, 4 KB
//     each, L2 evict-first) into a CTA-shared ring of stages.  Consumers wait on a per-slot mbarrier;
//     the warp that releases a slot last (release counter) issues the slot's next stage at once.
//     Chord offsets are arithmetic, there are no COO index gathers anywhere.
//   * MATH.  Every consumer thread owns 4 consecutive rows x 8 channels: all shared/global
//     operand accesses are 128-bit and bank-conflict free, FMAs are packed fma.rn.f32x2.  Taps
//     whose offset is not a multiple of 4 read the two aligned quads that cover them; in the
//     forward pass taps {0,1,2,4} share the same two quads (13 instead of 15 operand loads per
//     row group and channel).  Global operand loads ride a one-deep "pending" slot so their
//     latency hides behind the following local taps; taps are visited big-offset-first,
//     interleaved 1:3 with small ones.
//   * BACKWARD (output-stationary at row j):  one shifted read of G_m[j-o_k,:] feeds both
//     dV_{m-1}[j,:] += W_m[j-o_k,k] * G_m[j-o_k,:]  and  dW_m[j-o_k,k] = <G_m[j-o_k,:], V_{m-1}[j,:]>;
//     all 8 channels live in one thread, so dW needs no cross-CTA reduction and is written once.
//
// Reference semantics: SyntheticExperiments/psf.py:167-188 (loop + residual), spmul_cuda.cu:24,
// :79-80, :105-108 (per-factor formulas).
#include <cstdlib>

#include "psf_internal.cuh"
#include "psf_ptx.cuh"

namespace {

constexpr int kTile = 128;     // rows per warp tile (32 lanes x 4 rows)
constexpr int kMaxCW = 8;      // consumer warps per CTA
constexpr int kMaxStages = 8;  // W ring depth is a launch parameter
constexpr int kMaxThreads = kMaxCW * 32;

// ring geometry: forward stages 4 taps (the {0,1,2,4} head group is one stage); backward stages
// single taps whose windows carry 8 spare floats (16 B aligned, shifted source windows).
constexpr int kTpsF = 4;
#ifndef PSF_TPS_B
#define PSF_TPS_B 4
#endif
constexpr int kTpsB = PSF_TPS_B;  // 4 taps x 2 stages measured best; 1-tap stages with deep rings and 5-tap stages are slower
constexpr int kPadB = 8;

struct FusedParams {
  const float* W;
  float* dW;
  const float* Vin;   // fwd: V0 [B,n,8];   bwd: dOut [B,n,8]
  float* Vout;        // fwd: out [B,n,8];  bwd: dV0 [B,n,8]
  float* states;      // channel-major [slots][B][8][n]
  float* gx;          // bwd exchange  [2][B][8][n]
  float* racc;        // bwd residual accumulator [B][8][n]
  long long sWf, sWb;
  int B, n, n_W, L, residual, save_states;
  int P, RB;
  int head;           // fwd: visits 0..3 are the taps with offsets 0,1,2,4
  int stages;         // W ring depth (of the kernel being launched)
  int ncw;            // warps per CTA (warp 0 / lane 0 also streams W)
  int dbg;            // timing experiments only (PSF_FUSED_DEBUG): 1 no state stores, 2 no cluster barrier,
                      // 4 all taps local, 8 no W stream, 16 no state refresh, 32 no residual
  PsfOffsets ooffs;   // ooffs[i] = offset (mod n) of the tap visited i-th
  long long tapoff[PSF_MAX_TAPS];  // element offset of that tap's plane inside a factor (tap * sWt)
};

// ---------------------------------------------------------------------------------------------
// CTA geometry shared by producer and consumers
// ---------------------------------------------------------------------------------------------
struct Geom {
  int warp, lane, ncw;
  int blk_start, blk_end;  // rows of this CTA
  int ntiles;              // 128-row tiles in the block
  int npass;               // ceil(ntiles / ncw): a pass = one tile per consumer warp
  int ngroups;             // ceil(L / taps per stage)
  int cid, ncl;            // cluster id, number of clusters
  int prows;               // rows covered by a full pass (ncw * 128)
};

__device__ __forceinline__ void make_geom(Geom& g, const FusedParams& p, int tps) {
  g.warp = threadIdx.x >> 5;
  g.lane = threadIdx.x & 31;
  g.ncw = p.ncw;
  const int rank = blockIdx.x % p.P;
  g.cid = blockIdx.x / p.P;
  g.ncl = gridDim.x / p.P;
  g.blk_start = min(p.n, rank * p.RB);
  g.blk_end = min(p.n, g.blk_start + p.RB);
  g.ntiles = (g.blk_end - g.blk_start + kTile - 1) / kTile;
  g.npass = (g.ntiles + g.ncw - 1) / g.ncw;
  g.ngroups = (p.L + tps - 1) / tps;
  g.prows = g.ncw * kTile;
}

// Shared-memory carve-up (bytes from the start of dynamic smem):
//   [0, 8*plane*4)                      state block, channel-major
//   then ring: stages * tps * slot floats   (slot = ncw*128 (+8 backward))
//   then mbarriers: full[stages], empty[stages], vbar
struct Smem {
  float* state;
  uint32_t ring_u32, full_u32, vbar_u32, state_u32;
  int* relcnt;     // [stages] warps that have released each ring slot
  const float* ring;
  int slot;        // floats per tap window
  int stage_f;     // floats per stage
  const int* offs;         // [32] offset of the tap visited i-th (copied from the kernel parameters:
  const long long* tapoff; // [32]  indexing parameter arrays dynamically compiles to select chains)
};

template <int kPlane>
__device__ __forceinline__ void carve(Smem& sm, unsigned char* raw, const FusedParams& p, int tps, int pad) {
  sm.state = reinterpret_cast<float*>(raw);
  float* ring = sm.state + CT * kPlane;
  sm.ring = ring;
  sm.slot = p.ncw * kTile + pad;
  sm.stage_f = tps * sm.slot;
  sm.state_u32 = smem_u32(sm.state);
  sm.ring_u32 = smem_u32(ring);
  sm.full_u32 = sm.ring_u32 + (uint32_t)(p.stages * sm.stage_f) * 4u;
  sm.vbar_u32 = sm.full_u32 + (uint32_t)p.stages * 8u;
  sm.relcnt = reinterpret_cast<int*>(ring + p.stages * sm.stage_f) + 2 * (p.stages + 1);
  unsigned char* tab = reinterpret_cast<unsigned char*>(ring + p.stages * sm.stage_f) + (2 * p.stages + 2) * 8;
  sm.tapoff = reinterpret_cast<const long long*>(tab);
  sm.offs = reinterpret_cast<const int*>(tab + PSF_MAX_TAPS * 8);
}

// ---------------------------------------------------------------------------------------------
// W stream.  Every stage of this CTA has an index gidx = ((sample_iteration * n_W + factor_step) * npass
// + pass) * ngroups + tap_group, consumed in increasing order by all warps.  There is no producer warp
// and no cursor: the warp that releases a ring slot LAST (per-slot release counter) immediately issues
// the stage that goes into that slot next, gidx + stages, decoding the index arithmetically.
// ---------------------------------------------------------------------------------------------
template <bool kBackward, int kTps>
__device__ __forceinline__ void issue_stage_index(int gidx, int slot, const FusedParams& p, const Geom& g,
                                                  const Smem& sm) {
  const int spp = g.ngroups;        // stages per pass
  const int sps = spp * g.npass;    // stages per factor step
  const int spsample = sps * p.n_W;
  if (spsample == 0) return;
  const int it = gidx / spsample;   // which of this cluster's samples
  const int b = g.cid + it * g.ncl;
  if (b >= p.B) return;
  int r = gidx - it * spsample;
  const int mm = r / sps;
  r -= mm * sps;
  const int pp = r / spp;
  const int tg = r - pp * spp;
  const unsigned long long policy = (p.dbg & 64) ? 0ull : make_evict_first_policy();
  const int m = kBackward ? p.n_W - mm : mm + 1;
  const float* wfac = p.W + (long long)(m - 1) * p.sWf + (long long)b * p.sWb;
  const int row0 = g.blk_start + pp * g.prows;
  const int prow = min(g.prows, g.blk_end - row0);
  const uint32_t bar = sm.full_u32 + slot * 8;
  const uint32_t dst = sm.ring_u32 + (uint32_t)(slot * sm.stage_f) * 4u;
  const int i0 = tg * kTps;
  if (prow <= 0) {  // nothing to stream for this pass (cannot happen for non-empty blocks): complete the phase
    mbar_arrive(bar);
    return;
  }
  if (!kBackward) {
    const int kn = min(kTps, p.L - i0);
    const uint32_t bytes = (uint32_t)prow * 4u;
    mbar_expect_tx(bar, bytes * (uint32_t)kn);
#pragma unroll
    for (int t = 0; t < kTps; ++t)
      if (t < kn)
        tma_load_1d_stream(dst + (uint32_t)(t * sm.slot) * 4u, wfac + sm.tapoff[i0 + t] + row0, bytes, bar, policy);
  } else {
    uint32_t total = 0;
#pragma unroll
    for (int t = 0; t < kTps; ++t)
      if (i0 + t < p.L) total += (uint32_t)(prow + ((sm.offs[i0 + t] & 3) ? 4 : 0)) * 4u;
    mbar_expect_tx(bar, total);
#pragma unroll
    for (int t = 0; t < kTps; ++t) {
      if (i0 + t < p.L) {
        const int o = sm.offs[i0 + t];
        int s0 = row0 - o;
        if (s0 < 0) s0 += p.n;
        const int a0 = s0 & ~3;                    // 16 B aligned start of the source window
        const int len = prow + ((o & 3) ? 4 : 0);  // floats
        const int first = min(len, p.n - a0);      // part before the cyclic wrap
        const float* wk = wfac + sm.tapoff[i0 + t];
        const uint32_t d = dst + (uint32_t)(t * sm.slot) * 4u;
        tma_load_1d_stream(d, wk + a0, (uint32_t)first * 4u, bar, policy);
        if (first < len) tma_load_1d_stream(d + (uint32_t)first * 4u, wk, (uint32_t)(len - first) * 4u, bar, policy);
      }
    }
  }
}

// Called by lane 0 of every warp when the warp is done with stage `gs` in `slot`.
template <bool kBackward, int kTps>
__device__ __forceinline__ void release_stage(int gs, int slot, const FusedParams& p, const Geom& g, const Smem& sm) {
  int* cnt = sm.relcnt + slot;
  if (atomicAdd(cnt, 1) == g.ncw - 1) {   // last warp out refills the slot
    *cnt = 0;                              // ordered before the consumers' next use by the release in expect_tx
    issue_stage_index<kBackward, kTps>(gs + p.stages, slot, p, g, sm);
  }
}

// Consumer warp 0 / lane 0: pull this CTA's block of a channel-major global state into shared memory.
template <int kPlane>
__device__ __forceinline__ void refresh_state(const Smem& sm, const Geom& g, const float* state_b, int n) {
  const uint32_t bytes = (uint32_t)(g.blk_end - g.blk_start) * 4u;
  if (bytes == 0) {  // empty block: complete the phase without a copy
    mbar_arrive(sm.vbar_u32);
    return;
  }
  mbar_expect_tx(sm.vbar_u32, bytes * CT);
#pragma unroll
  for (int ch = 0; ch < CT; ++ch)
    tma_load_1d(sm.state_u32 + (uint32_t)(ch * kPlane) * 4u, state_b + (long long)ch * n + g.blk_start, bytes,
                sm.vbar_u32);
}

__device__ __forceinline__ void init_barriers(const FusedParams& p, const Smem& sm) {
  if (threadIdx.x == 0) {
    int* so = const_cast<int*>(sm.offs);
    long long* st = const_cast<long long*>(sm.tapoff);
#pragma unroll
    for (int i = 0; i < PSF_MAX_TAPS; ++i) {
      so[i] = p.ooffs.o[i];
      st[i] = p.tapoff[i];
    }
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(sm.full_u32 + s * 8, 1);
      sm.relcnt[s] = 0;
    }
    mbar_init(sm.vbar_u32, 1);
    fence_barrier_init();
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Forward
// ---------------------------------------------------------------------------------------------
template <int kPlane>
__global__ void __launch_bounds__(kMaxThreads, (kPlane <= 2048 ? 2 : 1)) psf_chain_fwd_kernel(const FusedParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kTps = kTpsF;
  Geom g;
  make_geom(g, p, kTps);
  Smem sm;
  carve<kPlane>(sm, smem_raw, p, kTps, 0);
  init_barriers(p, sm);

  if (threadIdx.x == 0 && !(p.dbg & 8)) {  // prime the ring
    for (int s = 0; s < p.stages; ++s) issue_stage_index<false, kTps>(s, s, p, g, sm);
  }
  int gs = 0;  // stages consumed so far (identical in every warp)

  const int S = p.stages;
  const int nthr_c = g.ncw * 32;
  const float* Vs = sm.state;
  const int n = p.n;
  const long long plane = (long long)CT * n;  // floats per sample per state
  int slot = 0;
  uint32_t parity = 0, vparity = 0;
  bool cw_pending = false;  // arrived on the cluster barrier, wait deferred to the first use of exchanged data
  const bool refresher = (threadIdx.x == 0);

  for (int b = g.cid; b < p.B; b += g.ncl) {
    // ---- prologue: V0 block -> state slot 0 (channel-major) ------------------------------
    float* st0 = p.states + (long long)b * plane;
    for (int pp = 0; pp < g.npass; ++pp) {
      const int j = g.blk_start + (pp * g.ncw + g.warp) * kTile + 4 * g.lane;
      if (j < g.blk_end) {
        float4 pl[CT];
        load_rows_natural(p.Vin + ((long long)b * n + j) * CT, pl);
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) st_cg4(st0 + ch * n + j, pl[ch]);
      }
    }
    fence_proxy_async();
    consumer_bar_sync(nthr_c);  // the block is written; nobody reads the shared copy any more
    if (refresher) refresh_state<kPlane>(sm, g, st0, n);
    cluster_arrive();
    cw_pending = true;
    if (p.dbg & 128) {
      cluster_wait();
      cw_pending = false;
    }

    for (int m = 1; m <= p.n_W; ++m) {
      // slot 0 always keeps V_0 (residual operand); without saved states slots 1/2 alternate
      const int slot_prev = (p.save_states || m == 1) ? (m - 1) : 1 + ((m - 1) & 1);
      const int slot_next = p.save_states ? m : 1 + (m & 1);
      const float* prev = p.states + ((long long)slot_prev * p.B + b) * plane;
      float* next = p.states + ((long long)slot_next * p.B + b) * plane;
      const bool last = (m == p.n_W);

      mbar_wait(sm.vbar_u32, vparity);  // V_{m-1} block has landed in shared memory
      vparity ^= 1u;

      for (int pp = 0; pp < g.npass; ++pp) {
        const int row0 = g.blk_start + (pp * g.ncw + g.warp) * kTile;
        const int j = row0 + 4 * g.lane;
        const bool mine = j < g.blk_end;
        float4 acc[CT];
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) acc[ch] = make_float4(0.f, 0.f, 0.f, 0.f);

        const float* vown = Vs + (j - g.blk_start);
        for (int tg = 0; tg < g.ngroups; ++tg) {
          if (!(p.dbg & 8)) mbar_wait(sm.full_u32 + slot * 8, parity);
          const float* wst = sm.ring + slot * sm.stage_f + g.warp * kTile + 4 * g.lane;
          // The deferred cluster wait must be executed by whole warps: decide it from warp-uniform values
          // (tile origin, tap offset) - does any row of this tile reach past the block with this tap?
          if (tg == 0 && p.head) {
            if (cw_pending && row0 + kTile + 8 > g.blk_end) {
              cluster_wait();
              cw_pending = false;
            }
            if (mine) {
              // taps with offsets 0,1,2,4 share the quads A = rows j..j+3 and Bq = rows j+4..j+7
              const float4 w0 = lds4(wst), w1 = lds4(wst + sm.slot), w2 = lds4(wst + 2 * sm.slot),
                           w3 = lds4(wst + 3 * sm.slot);
              const bool bq_local = (j + 8 <= g.blk_end);
              const int jb = (j + 4 >= n) ? j + 4 - n : j + 4;
#pragma unroll
              for (int ch = 0; ch < CT; ++ch) {
                const float4 A = lds4(vown + ch * kPlane);
                const float4 Bq = bq_local ? lds4(vown + ch * kPlane + 4) : ld_cg4(prev + ch * n + jb);
                // row r: w0[r]*V[j+r] + w1[r]*V[j+r+1] + w2[r]*V[j+r+2] + w3[r]*V[j+r+4]
                fma4(acc[ch], w0, A);                                // offset 0
                fma4(acc[ch], w3, Bq);                               // offset 4
                fma2(acc[ch].x, acc[ch].y, w2.x, w2.y, A.z, A.w);    // offset 2, rows 0-1
                fma2(acc[ch].z, acc[ch].w, w2.z, w2.w, Bq.x, Bq.y);  // offset 2, rows 2-3
                acc[ch].x = fmaf(w1.x, A.y, acc[ch].x);              // offset 1
                acc[ch].y = fmaf(w1.y, A.z, acc[ch].y);
                acc[ch].z = fmaf(w1.z, A.w, acc[ch].z);
                acc[ch].w = fmaf(w1.w, Bq.x, acc[ch].w);
              }
            }
          } else {
#pragma unroll
            for (int t = 0; t < kTps; ++t) {
              const int i = tg * kTps + t;
              if (i < p.L) {
                const int o = sm.offs[i];
                if (cw_pending && row0 + kTile + 8 + o > g.blk_end) {
                  cluster_wait();
                  cw_pending = false;
                }
                if (mine) {
                  const float4 w4 = lds4(wst + t * sm.slot);
                  const int src = j + o;
                  if ((o & 3) == 0) {
                    if (src + 4 <= g.blk_end || (p.dbg & 4)) {
                      const float* vp = (p.dbg & 4) ? Vs + ((src - g.blk_start) & (kPlane - 4)) : vown + o;
                      float4 v[CT];
#pragma unroll
                      for (int ch = 0; ch < CT; ++ch) v[ch] = lds4(vp + ch * kPlane);
#pragma unroll
                      for (int ch = 0; ch < CT; ++ch) fma4(acc[ch], w4, v[ch]);
                    } else {
                      const float* gp = prev + (src >= n ? src - n : src);
                      float4 v[CT];
#pragma unroll
                      for (int ch = 0; ch < CT; ++ch) v[ch] = ld_cg4(gp + ch * n);
#pragma unroll
                      for (int ch = 0; ch < CT; ++ch) fma4(acc[ch], w4, v[ch]);
                    }
                  } else {
                    // unaligned tap: the two aligned quads covering rows src..src+3, select in registers
                    const int sh = o & 3;
                    const int a = src - sh;
                    const bool xl = (a + 4 <= g.blk_end), yl = (a + 8 <= g.blk_end);
                    const int ax = a >= n ? a - n : a;
                    const int ay = (a + 4 >= n) ? a + 4 - n : a + 4;
#pragma unroll
                    for (int ch = 0; ch < CT; ++ch) {
                      const float4 X = xl ? lds4_full(vown + (o - sh) + ch * kPlane) : ld_cg4(prev + ch * n + ax);
                      const float4 Y = yl ? lds4_full(vown + (o - sh) + 4 + ch * kPlane) : ld_cg4(prev + ch * n + ay);
                      fma4(acc[ch], w4, select_quad(X, Y, sh));
                    }
                  }
                }
              }
            }
          }
          __syncwarp();
          if (!(p.dbg & 8)) {
            if (g.lane == 0) release_stage<false, kTps>(gs, slot, p, g, sm);  // this warp is done with the stage
            ++gs;
            if (++slot == S) {
              slot = 0;
              parity ^= 1u;
            }
          }
        }
        if (cw_pending) {  // at the latest before this CTA publishes V_m (state slots may be reused)
          cluster_wait();
          cw_pending = false;
        }
        if (mine) {
          if (p.residual) {
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) add4(acc[ch], ld_cg4(st0 + ch * n + j));
          }
          if (!last) {
            if (!(p.dbg & 1) || acc[0].x == 123.456f) {
#pragma unroll
              for (int ch = 0; ch < CT; ++ch) st_cg4(next + ch * n + j, acc[ch]);
            }
          } else {
            store_rows_natural(p.Vout + ((long long)b * n + j) * CT, acc);
          }
        }
      }
      if (!last) {
        fence_proxy_async();
        consumer_bar_sync(nthr_c);  // all reads of the shared block and all writes of V_m are done
        if (refresher) {
          if (p.dbg & 16) mbar_arrive(sm.vbar_u32); else refresh_state<kPlane>(sm, g, next, n);
        }
        if (!(p.dbg & 2)) {
          cluster_arrive();           // V_m visible cluster-wide once the (deferred) wait returns
          cw_pending = true;
          if (p.dbg & 128) {
            cluster_wait();
            cw_pending = false;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Backward
// ---------------------------------------------------------------------------------------------
template <int kPlane>
__global__ void __launch_bounds__(kMaxThreads, (kPlane <= 2048 ? 2 : 1)) psf_chain_bwd_kernel(const FusedParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kTps = kTpsB;
  Geom g;
  make_geom(g, p, kTps);
  Smem sm;
  carve<kPlane>(sm, smem_raw, p, kTps, kPadB);
  init_barriers(p, sm);

  if (threadIdx.x == 0) {  // prime the ring
    for (int s = 0; s < p.stages; ++s) issue_stage_index<true, kTps>(s, s, p, g, sm);
  }
  int gs = 0;  // stages consumed so far (identical in every warp)

  const int S = p.stages;
  const int nthr_c = g.ncw * 32;
  const float* Gs = sm.state;
  const int n = p.n;
  const long long plane = (long long)CT * n;
  int slot = 0;
  uint32_t parity = 0, vparity = 0;
  bool cw_pending = false;  // arrived on the cluster barrier, wait deferred to the first use of exchanged data
  const bool refresher = (threadIdx.x == 0);

  for (int b = g.cid; b < p.B; b += g.ncl) {
    // exchange / residual buffers are per CLUSTER (reused sample after sample): 3 x 0.5 MB x 33 clusters
    // stays L2 resident, whereas per-sample buffers were written back to HBM once per factor
    float* gx0 = p.gx + (long long)g.cid * plane;            // exchange slot 0
    float* gx1 = p.gx + ((long long)g.ncl + g.cid) * plane;  // exchange slot 1
    float* racc = p.racc + (long long)g.cid * plane;
    // ---- prologue: dOut block -> exchange slot 0 (channel-major) ---------------------------
    for (int pp = 0; pp < g.npass; ++pp) {
      const int j = g.blk_start + (pp * g.ncw + g.warp) * kTile + 4 * g.lane;
      if (j < g.blk_end) {
        float4 pl[CT];
        load_rows_natural(p.Vin + ((long long)b * n + j) * CT, pl);
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) st_cg4(gx0 + ch * n + j, pl[ch]);
      }
    }
    fence_proxy_async();
    consumer_bar_sync(nthr_c);
    if (refresher) refresh_state<kPlane>(sm, g, gx0, n);
    cluster_arrive();
    cw_pending = true;
    if (p.dbg & 128) {
      cluster_wait();
      cw_pending = false;
    }

    for (int m = p.n_W; m >= 1; --m) {
      const int step = p.n_W - m;
      const float* gcur = (step & 1) ? gx1 : gx0;
      float* gnext = (step & 1) ? gx0 : gx1;
      const float* vprev = p.states + ((long long)(m - 1) * p.B + b) * plane;  // V_{m-1}, channel-major
      float* dwbase = p.dW + (long long)(m - 1) * p.sWf + (long long)b * p.sWb;
      const bool last = (m == 1);
      const bool first = (m == p.n_W);

      mbar_wait(sm.vbar_u32, vparity);  // G_m block has landed in shared memory
      vparity ^= 1u;

      for (int pp = 0; pp < g.npass; ++pp) {
        const int row0 = g.blk_start + (pp * g.ncw + g.warp) * kTile;
        const int j = row0 + 4 * g.lane;
        const bool mine = j < g.blk_end;
        float4 acc[CT];
        float4 vown[CT];
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) acc[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (mine) {
#pragma unroll
          for (int ch = 0; ch < CT; ++ch) vown[ch] = ld_cg4(vprev + ch * n + j);
        }
        const float* gown = Gs + (j - g.blk_start);
        for (int tg = 0; tg < g.ngroups; ++tg) {
          mbar_wait(sm.full_u32 + slot * 8, parity);
          const float* wst = sm.ring + slot * sm.stage_f + g.warp * kTile + 4 * g.lane;
          {
#pragma unroll
            for (int t = 0; t < kTps; ++t) {
              const int i = tg * kTps + t;
              if (i < p.L) {
                const int o = sm.offs[i];
                // deferred cluster wait, decided from warp-uniform values (whole warps must execute it)
                if (cw_pending && row0 - o - 8 < g.blk_start) {
                  cluster_wait();
                  cw_pending = false;
                }
                if (!mine) continue;
                const int src = j - o;  // first source row (may be negative: cyclic)
                float* dwk = dwbase + sm.tapoff[i];
                if ((o & 3) == 0) {
                  const float4 w4 = lds4(wst + t * sm.slot);
                  float4 gv[CT];
                  int sw;
                  if (src >= g.blk_start) {
                    sw = src;
#pragma unroll
                    for (int ch = 0; ch < CT; ++ch) gv[ch] = lds4(gown - o + ch * kPlane);
                  } else {
                    sw = src < 0 ? src + n : src;
                    const float* gp = gcur + sw;
#pragma unroll
                    for (int ch = 0; ch < CT; ++ch) gv[ch] = ld_cg4(gp + ch * n);
                  }
                  float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                  for (int ch = 0; ch < CT; ++ch) {
                    fma4(acc[ch], w4, gv[ch]);
                    fma4(dw, gv[ch], vown[ch]);
                  }
                  st_cs4(dwk + sw, dw);
                } else {
                  // unaligned tap: aligned quads X = rows a..a+3, Y = rows a+4..a+7 cover src..src+3.
                  // The staged W window starts at the aligned-down row of (prow0 - o), i.e. `sh` floats early.
                  const int sh = (4 - (o & 3)) & 3;  // src mod 4 (j and prow0 are multiples of 4)
                  const int a = src - sh;
                  const bool xl = (a >= g.blk_start), yl = (a + 4 >= g.blk_start);
                  const int ax = a < 0 ? a + n : a;
                  const int ay = (a + 4 < 0) ? a + 4 + n : a + 4;
                  const float4 w4 = select_quad(lds4_full(wst + t * sm.slot), lds4_full(wst + t * sm.slot + 4), sh);
                  float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                  for (int ch = 0; ch < CT; ++ch) {
                    const float4 X = xl ? lds4_full(gown - o - sh + ch * kPlane) : ld_cg4(gcur + ch * n + ax);
                    const float4 Y = yl ? lds4_full(gown - o - sh + 4 + ch * kPlane) : ld_cg4(gcur + ch * n + ay);
                    const float4 gq = select_quad(X, Y, sh);
                    fma4(acc[ch], w4, gq);
                    fma4(dw, gq, vown[ch]);
                  }
                  // rows src..src+3 (cyclic) of dW_k: not 16 B aligned -> four scalar stores
                  int r0 = src < 0 ? src + n : src;
                  __stcs(dwk + r0, dw.x);
                  r0 = (r0 + 1 == n) ? 0 : r0 + 1;
                  __stcs(dwk + r0, dw.y);
                  r0 = (r0 + 1 == n) ? 0 : r0 + 1;
                  __stcs(dwk + r0, dw.z);
                  r0 = (r0 + 1 == n) ? 0 : r0 + 1;
                  __stcs(dwk + r0, dw.w);
                }
              }
            }
          }
          __syncwarp();
          if (g.lane == 0) release_stage<true, kTps>(gs, slot, p, g, sm);
          ++gs;
          if (++slot == S) {
            slot = 0;
            parity ^= 1u;
          }
        }
        if (cw_pending) {
          cluster_wait();
          cw_pending = false;
        }
        if (mine) {
          // residual: dV0 also receives G_m of every factor (V_m = W_m V_{m-1} + V_0)
          if (p.residual) {
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) {
              float4 own = lds4(gown + ch * kPlane);
              if (!first) add4(own, ld_cg4(racc + ch * n + j));
              if (!last) {
                st_cg4(racc + ch * n + j, own);
              } else {
                add4(acc[ch], own);
              }
            }
          }
          if (!last) {
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) st_cg4(gnext + ch * n + j, acc[ch]);
          } else {
            store_rows_natural(p.Vout + ((long long)b * n + j) * CT, acc);
          }
        }
      }
      if (!last) {
        fence_proxy_async();
        consumer_bar_sync(nthr_c);
        if (refresher) refresh_state<kPlane>(sm, g, gnext, n);
        cluster_arrive();
        cw_pending = true;
        if (p.dbg & 128) {
          cluster_wait();
          cw_pending = false;
        }
      }
    }
    // the exchange buffers belong to the cluster: the next sample may overwrite them only after every
    // CTA has finished reading this sample's last state
    cluster_arrive();
    cluster_wait();
  }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
int plane_for(int RB) { return RB <= 512 ? 512 : (RB <= 2048 ? 2048 : 4096); }

int smem_bytes_for(int RB, int ncw, int stages, bool backward) {  // stages of THAT kernel
  const int slot = ncw * kTile + (backward ? kPadB : 0);
  const int ring = stages * (backward ? kTpsB : kTpsF) * slot;
  return CT * plane_for(RB) * 4 + ring * 4 + (2 * stages + 2) * 8 + PSF_MAX_TAPS * 12;
}

const void* fwd_kernel_for(int RB) {
  switch (plane_for(RB)) {
    case 512: return (const void*)psf_chain_fwd_kernel<512>;
    case 2048: return (const void*)psf_chain_fwd_kernel<2048>;
    default: return (const void*)psf_chain_fwd_kernel<4096>;
  }
}
const void* bwd_kernel_for(int RB) {
  switch (plane_for(RB)) {
    case 512: return (const void*)psf_chain_bwd_kernel<512>;
    case 2048: return (const void*)psf_chain_bwd_kernel<2048>;
    default: return (const void*)psf_chain_bwd_kernel<4096>;
  }
}

int env_int(const char* name, int dflt) {
  const char* v = std::getenv(name);
  if (v == nullptr || *v == 0) return dflt;
  const int x = std::atoi(v);
  return x > 0 ? x : dflt;
}

std::mutex g_mu1;
bool g_attr_done[PSF_MAX_DEVICES] = {};

struct OccCache {
  int dev, P, threads, smem, clusters;
  const void* kernel;
};
OccCache g_occ[32];
int g_occ_n = 0;

int resident_clusters(const void* kernel, int P, int threads, int smem, int sms) {   // caller holds g_mu1
  const int dev = psf_current_device();
  for (int i = 0; i < g_occ_n; ++i)
    if (g_occ[i].dev == dev && g_occ[i].P == P && g_occ[i].threads == threads && g_occ[i].smem == smem && g_occ[i].kernel == kernel)
      return g_occ[i].clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(P * sms * 2), 1, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)P;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int nclusters = 0;
  if (cudaOccupancyMaxActiveClusters(&nclusters, kernel, &cfg) != cudaSuccess || nclusters < 1) {
    cudaGetLastError();
    return 0;
  }
  if (g_occ_n < 32) g_occ[g_occ_n++] = {dev, P, threads, smem, nclusters, kernel};
  return nclusters;
}

// Tap visiting order.  Forward: the {0,1,2,4} head first (if present), then big offsets interleaved
// 1:3 with small ones so every (possibly remote) big tap is followed by local work.  Backward: natural.
void make_order(const PsfChainArgs& a, bool backward, int* order, int* head) {
  const int L = a.L;
  *head = 0;
  for (int i = 0; i < PSF_MAX_TAPS; ++i) order[i] = i < L ? i : 0;
  if (backward) return;
  bool used[PSF_MAX_TAPS] = {false};
  int pos = 0;
  const int want[4] = {0, 1, 2, 4};
  int hk[4] = {-1, -1, -1, -1};
  for (int w = 0; w < 4; ++w)
    for (int k = 0; k < L; ++k)
      if (hk[w] < 0 && !used[k] && a.offs.o[k] == want[w]) {
        hk[w] = k;
        used[k] = true;
      }
  if (hk[0] >= 0 && hk[1] >= 0 && hk[2] >= 0 && hk[3] >= 0 && a.n >= 8) {
    *head = 1;
    for (int w = 0; w < 4; ++w) order[pos++] = hk[w];
  } else {
    for (int k = 0; k < L; ++k) used[k] = false;
  }
  int rest[PSF_MAX_TAPS], nr = 0;
  for (int k = 0; k < L; ++k)
    if (!used[k]) rest[nr++] = k;
  for (int i = 1; i < nr; ++i) {  // insertion sort by offset (stable)
    const int k = rest[i];
    int q = i - 1;
    while (q >= 0 && a.offs.o[rest[q]] > a.offs.o[k]) {
      rest[q + 1] = rest[q];
      --q;
    }
    rest[q + 1] = k;
  }
  const int nb = (nr + 3) / 4;
  int lo = 0, hi = nr - 1, bigs = 0;
  while (lo <= hi) {
    for (int q = 0; q < 3 && lo <= hi; ++q) order[pos++] = rest[lo++];
    if (bigs < nb && lo <= hi) {
      order[pos++] = rest[hi--];
      ++bigs;
    }
  }
}

int launch(const void* kernel, const FusedParams& fp, int P, int threads, int smem, int resident, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  const int clusters = resident < fp.B ? resident : fp.B;
  cfg.gridDim = dim3((unsigned)(clusters * P), 1, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)P;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  void* args[] = {(void*)&fp};
  return (int)cudaLaunchKernelExC(&cfg, kernel, args);
}

void fill_common(FusedParams& fp, const PsfChainArgs& a, const PsfFusedPlan& plan, bool backward) {
  fp.sWf = a.ws.factor;
  fp.sWb = a.ws.sample;
  fp.B = (int)a.B;
  fp.n = (int)a.n;
  fp.n_W = a.n_W;
  fp.L = a.L;
  fp.residual = a.residual;
  fp.P = plan.cluster;
  fp.RB = plan.rows_per_cta;
  fp.stages = backward ? plan.stages_bwd : plan.stages;
  fp.ncw = plan.threads / 32;
  fp.dbg = env_int("PSF_FUSED_DEBUG", 0);
  int order[PSF_MAX_TAPS];
  make_order(a, backward, order, &fp.head);
  for (int i = 0; i < PSF_MAX_TAPS; ++i) {
    fp.ooffs.o[i] = a.offs.o[order[i]];
    fp.tapoff[i] = (long long)order[i] * a.ws.tap;
  }
}

}  // namespace

int psf_fused_plan(const PsfChainArgs& a, PsfFusedPlan* plan) {
  if (a.C != CT || a.n < 4 || (a.n & 3) || a.L < 1 || a.L > PSF_MAX_TAPS || a.n_W < 1 || a.B < 1 ||
      a.B > (1 << 20))
    return PSF_ERR_UNSUPPORTED;
  if (a.ws.row != 1 || (a.ws.factor & 3) || (a.ws.sample & 3) || (a.ws.tap & 3)) return PSF_ERR_UNSUPPORTED;
  int sms = 148;
  const int dev = psf_current_device();
  if (dev < 0) return PSF_ERR_UNSUPPORTED;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // block rows: 2048 (64 KB state, two CTAs per SM) unless that needs more than 8 CTAs per sample
  int max_rows = env_int("PSF_FUSED_MAX_ROWS", 2048);
  if (max_rows > 4096) max_rows = 4096;
  int P = 1;
  while ((a.n + P - 1) / P > max_rows && P < 8) P *= 2;
  if ((a.n + P - 1) / P > max_rows) {
    max_rows = 4096;
    if ((a.n + P - 1) / P > max_rows) return PSF_ERR_UNSUPPORTED;
  }
  // spread small batches over more SMs while a CTA keeps at least 4 warp tiles
  while (P < 8 && a.B * P <= sms && (a.n + 2 * P - 1) / (2 * P) >= 4 * kTile) P *= 2;
  const int per = (int)((a.n + P - 1) / P);
  const int RB = (per + 3) & ~3;
  int ncw = (RB + kTile - 1) / kTile;
  const int max_cw = env_int("PSF_FUSED_CONSUMER_WARPS", kMaxCW);
  if (ncw > max_cw) ncw = max_cw;
  if (ncw > kMaxCW) ncw = kMaxCW;
  const int threads = ncw * 32;
  // two stages: the smaller ring leaves ~36 KB of L1 per SM (spill traffic otherwise goes to L2); deeper rings
  // measured slower (tools/knob_sweep.sh)
  int stages = env_int("PSF_FUSED_STAGES", 2);
  int stages_b = env_int("PSF_FUSED_STAGES_BWD", 2);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) stages = 2;
  if (stages_b > kMaxStages) stages_b = kMaxStages;
  if (stages_b < 2) stages_b = 2;
  while (stages > 2 && smem_bytes_for(RB, ncw, stages, false) > 227 * 1024) --stages;
  while (stages_b > 2 && smem_bytes_for(RB, ncw, stages_b, true) > 227 * 1024) --stages_b;
  if (smem_bytes_for(RB, ncw, stages, false) > 227 * 1024 || smem_bytes_for(RB, ncw, stages_b, true) > 227 * 1024)
    return PSF_ERR_UNSUPPORTED;
  plan->cluster = P;
  plan->rows_per_cta = RB;
  plan->threads = threads;
  plan->stages = stages;
  plan->stages_bwd = stages_b;
  plan->smem_bytes = smem_bytes_for(RB, ncw, stages, false);
  plan->smem_bytes_bwd = smem_bytes_for(RB, ncw, stages_b, true);

  std::lock_guard<std::mutex> lock(g_mu1);
  if (!g_attr_done[dev]) {
    const int planes[3] = {512, 2048, 4096};
    for (int q = 0; q < 3; ++q)
      if (cudaFuncSetAttribute(fwd_kernel_for(planes[q]), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
              cudaSuccess ||
          cudaFuncSetAttribute(bwd_kernel_for(planes[q]), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
              cudaSuccess)
        return PSF_ERR_UNSUPPORTED;
    g_attr_done[dev] = true;
  }
  plan->resident = resident_clusters(fwd_kernel_for(RB), P, threads, plan->smem_bytes, sms);
  plan->resident_bwd = resident_clusters(bwd_kernel_for(RB), P, threads, plan->smem_bytes_bwd, sms);
  if (plan->resident < 1 || plan->resident_bwd < 1) return PSF_ERR_UNSUPPORTED;
  return PSF_OK;
}

bool psf_fused_buffers_ok(const PsfChainArgs& a, const void* W, const void* p1, const void* p2, const void* ws) {
  (void)a;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  return al(W) && al(p1) && al(p2) && al(ws);
}

// Workspace (floats, S = B*n*8):  states [0, n_W*S) (save) or [0, 3*S);  then gx 2*S, racc S (backward only).
int psf_fused_chain_fwd(const PsfChainArgs& a, const float* W, const float* V0, float* out, float* ws, int save_states,
                        cudaStream_t st) {
  PsfFusedPlan plan;
  int rc = psf_fused_plan(a, &plan);
  if (rc) return rc;
  FusedParams fp = {};
  fp.W = W;
  fp.Vin = V0;
  fp.Vout = out;
  fp.states = ws;
  fp.save_states = save_states;
  fill_common(fp, a, plan, false);
  return launch(fwd_kernel_for(plan.rows_per_cta), fp, plan.cluster, plan.threads, plan.smem_bytes, plan.resident, st);
}

int psf_fused_chain_bwd(const PsfChainArgs& a, const float* W, const float* dOut, float* dW, float* dV0, float* ws,
                        cudaStream_t st) {
  PsfFusedPlan plan;
  int rc = psf_fused_plan(a, &plan);
  if (rc) return rc;
  const long long S = a.B * a.n * a.C;
  FusedParams fp = {};
  fp.W = W;
  fp.dW = dW;
  fp.Vin = dOut;
  fp.Vout = dV0;
  fp.states = ws;
  fp.gx = ws + (long long)a.n_W * S;
  fp.racc = fp.gx + 2 * S;
  fp.save_states = 1;
  fill_common(fp, a, plan, true);
  return launch(bwd_kernel_for(plan.rows_per_cta), fp, plan.cluster, plan.threads, plan.smem_bytes_bwd,
                plan.resident_bwd, st);
}
