// Fused chord-factor chain for sm_100a (B200): forward and backward, fp32, C = 8 channels.
//
// One thread-block CLUSTER owns one sample.  CTA q of a P-CTA cluster owns the contiguous row
// block [q*RB, (q+1)*RB) for ALL channels and keeps it in shared memory, channel-major
// ([C][RB] fp32), across all n_W factors.  Facts the design is built on:
//   * chord links only point forward (row i reads rows i+o_k, o_k >= 0), so a block can be
//     updated IN PLACE by an ascending sweep over 128-row warp tiles (descending for the
//     backward pass, whose links point backward): once a tile has been read by itself and by
//     every lower tile, nobody inside the CTA needs its old value again.  No ping-pong buffer,
//     so a 2048-row block costs 64 KB and two CTAs (of different samples) share an SM: while one
//     waits at a barrier the other computes.
//   * operands that fall outside the CTA's block (big taps, the halo of small taps, the cyclic
//     wrap) are read from the previous state in global memory.  That state was written by the
//     sibling CTAs a few microseconds earlier and is L2 resident; a hardware cluster barrier
//     (release/acquire) per factor orders the exchange.  The same buffers are the saved
//     states V_0..V_{n_W-1} the backward pass needs, so training pays nothing extra for it.
//   * W is streamed exactly once: tap-major ([..][tap][row], row stride 1) so that the 128 rows
//     of a warp tile are one contiguous 512 B run per tap.  Lane 0 of every warp prefetches its
//     own runs with cp.async.bulk (TMA, 1-D) into a private 3-stage ring guarded by mbarriers;
//     no producer warp, no cross-warp coupling, chord offsets are arithmetic (no COO gathers).
//   * every thread owns 4 consecutive rows x 8 channels: all shared/global operand accesses
//     are 128-bit and bank-conflict free (consecutive lanes -> consecutive 16 B).  Taps whose
//     offset is not a multiple of 4 read the two aligned quads that cover them and select in
//     registers; in the forward pass taps {0,1,2,4} share the same two quads (13 instead of 15
//     operand loads per row group and channel).
//   * global (remote / residual) operand loads are issued one tap group ahead and consumed after
//     the following local taps ("pending" slot); taps are visited big-offset-first, interleaved
//     with small ones, so that every remote tap has local work to hide behind.
//   * backward (output-stationary at row j):  one shifted read of G_m[j-o_k,:] feeds both
//     dV_{m-1}[j,:] += W_m[j-o_k,k] * G_m[j-o_k,:]  and  dW_m[j-o_k,k] = <G_m[j-o_k,:], V_{m-1}[j,:]>,
//     all C channels live in one thread so dW needs no cross-CTA reduction and is written once.
//
// Reference semantics: SyntheticExperiments/psf.py:167-188 (loop + residual), spmul_cuda.cu:24,
// :79-80, :105-108 (per-factor formulas).
#include <cstdlib>

#include "psf_internal.cuh"

namespace {

constexpr int CT = 8;        // channels held per thread == C of the fast path
constexpr int kTile = 128;   // rows per warp tile (32 lanes x 4 rows)
constexpr int kMaxStages = 8; // W ring depth per warp is a launch parameter (<= 8)
constexpr int kMaxThreads = 512;

// ring geometry: forward stages 4 taps (the {0,1,2,4} head group is one stage) of exactly one tile;
// backward stages 3 taps with 8 spare floats (shifted, 16 B aligned source windows).
constexpr int kTpsF = 4, kSlotF = kTile;
constexpr int kTpsB = 3, kSlotB = kTile + 8;

struct FusedParams {
  const float* W;
  float* dW;
  const float* Vin;   // fwd: V0 [B,n,8];   bwd: dOut [B,n,8]
  float* Vout;        // fwd: out [B,n,8];  bwd: dV0 [B,n,8]
  float* states;      // channel-major [slots][B][8][n]
  float* gx;          // bwd exchange  [2][B][8][n]
  float* racc;        // bwd residual accumulator [B][8][n]
  int64_t sWf, sWb, sWt;
  int B, n, n_W, L, residual, save_states;
  int P, RB;
  int head;           // fwd: order[0..3] are the taps with offsets 0,1,2,4
  int stages;         // W ring depth per warp
  PsfOffsets offs;    // offset of tap k (mod n)
  PsfOffsets order;   // order[i] = tap visited i-th
  PsfOffsets ooffs;   // ooffs[i] = offs[order[i]]
  long long tapoff[PSF_MAX_TAPS];  // order[i] * sWt (elements)
};

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
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async;" ::: "memory");
}
// Hardware cluster barrier; release/acquire at cluster scope orders the global-memory exchange.
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n"
               "barrier.cluster.wait.acquire.aligned;\n" ::
                   : "memory");
}

__device__ __forceinline__ float4 ld_cg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st_cg4(float* p, float4 v) { __stcg(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ void st_cs4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

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

// ---------------------------------------------------------------------------------------------
// Per-warp W prefetch cursor.  The issue order must equal the consume order:
//   for sample b (this cluster's samples) / factor m / chunk c (tiles of this warp) / tap group tg.
// kBackward: m and c run downwards, and the staged rows are shifted by -o_k.
// ---------------------------------------------------------------------------------------------
struct WCursor {
  int b, m, c, tg;
  bool valid;
  int row0, rw;        // first row / rows of the tile being staged
  const float* wptr;   // W + (m-1)*sWf + b*sWb (+ row0 in the forward pass)
};

struct WarpGeom {
  int warp, lane, nwarps;
  int blk_start, blk_end;  // rows of this CTA
  int ntiles;              // 128-row tiles in the block
  int nchunks;             // ceil(ntiles / nwarps)
  int ngroups;             // ceil(L / taps per stage)
  int cid, ncl;            // cluster id, number of clusters
};

__device__ __forceinline__ void make_geom(WarpGeom& g, const FusedParams& p, int tps) {
  g.warp = threadIdx.x >> 5;
  g.lane = threadIdx.x & 31;
  g.nwarps = blockDim.x >> 5;
  const int rank = blockIdx.x % p.P;
  g.cid = blockIdx.x / p.P;
  g.ncl = gridDim.x / p.P;
  g.blk_start = min(p.n, rank * p.RB);
  g.blk_end = min(p.n, g.blk_start + p.RB);
  g.ntiles = (g.blk_end - g.blk_start + kTile - 1) / kTile;
  g.nchunks = (g.ntiles + g.nwarps - 1) / g.nwarps;
  g.ngroups = (p.L + tps - 1) / tps;
}

__device__ __forceinline__ int last_chunk_of_warp(const WarpGeom& g) {
  return (g.ntiles - 1 - g.warp) / g.nwarps;  // largest c with c*nwarps + warp < ntiles
}

template <bool kBackward>
__device__ __forceinline__ void cursor_tile(WCursor& cur, const FusedParams& p, const WarpGeom& g) {
  const int tile = cur.c * g.nwarps + g.warp;
  cur.row0 = g.blk_start + tile * kTile;
  cur.rw = min(kTile, g.blk_end - cur.row0);
  cur.wptr = p.W + (long long)(cur.m - 1) * p.sWf + (long long)cur.b * p.sWb + (kBackward ? 0 : cur.row0);
}

template <bool kBackward>
__device__ __forceinline__ void cursor_init(WCursor& cur, const FusedParams& p, const WarpGeom& g) {
  cur.b = g.cid;
  cur.tg = 0;
  cur.valid = (g.warp < g.ntiles) && (cur.b < p.B) && (p.n_W > 0);
  if (kBackward) {
    cur.m = p.n_W;
    cur.c = cur.valid ? last_chunk_of_warp(g) : 0;
  } else {
    cur.m = 1;
    cur.c = 0;
  }
  if (cur.valid) cursor_tile<kBackward>(cur, p, g);
}

template <bool kBackward>
__device__ __forceinline__ void cursor_next(WCursor& cur, const FusedParams& p, const WarpGeom& g) {
  if (++cur.tg < g.ngroups) return;
  cur.tg = 0;
  bool wrapped = false;
  if (kBackward) {
    if (--cur.c < 0) {
      cur.c = last_chunk_of_warp(g);
      if (--cur.m < 1) {
        cur.m = p.n_W;
        wrapped = true;
      }
    }
  } else {
    ++cur.c;
    if (cur.c * g.nwarps + g.warp >= g.ntiles) {
      cur.c = 0;
      if (++cur.m > p.n_W) {
        cur.m = 1;
        wrapped = true;
      }
    }
  }
  if (wrapped) {
    cur.b += g.ncl;
    cur.valid = cur.b < p.B;
  }
  if (cur.valid) cursor_tile<kBackward>(cur, p, g);
}

// Lane 0: arm the stage barrier and launch the bulk copies of one stage (<= kTps taps).
template <bool kBackward, int kTps, int kSlot>
__device__ __forceinline__ void issue_stage(const WCursor& cur, const FusedParams& p, uint32_t dst, uint32_t bar) {
  const int i0 = cur.tg * kTps;
  if (!kBackward) {
    const uint32_t bytes = (uint32_t)cur.rw * 4u;
    const int kn = min(kTps, p.L - i0);
    mbar_expect_tx(bar, bytes * (uint32_t)kn);
#pragma unroll
    for (int t = 0; t < kTps; ++t)
      if (t < kn) tma_load_1d(dst + t * kSlot * 4, cur.wptr + p.tapoff[i0 + t], bytes, bar);
  } else {
    uint32_t total = 0;
#pragma unroll
    for (int t = 0; t < kTps; ++t)
      if (i0 + t < p.L) total += (uint32_t)(cur.rw + ((p.ooffs.o[i0 + t] & 3) ? 4 : 0)) * 4u;
    mbar_expect_tx(bar, total);
#pragma unroll
    for (int t = 0; t < kTps; ++t) {
      if (i0 + t < p.L) {
        const int o = p.ooffs.o[i0 + t];
        int s0 = cur.row0 - o;
        if (s0 < 0) s0 += p.n;
        const int a0 = s0 & ~3;                      // 16 B aligned start of the source window
        const int len = cur.rw + ((o & 3) ? 4 : 0);  // floats
        const int first = min(len, p.n - a0);        // part before the cyclic wrap
        const float* wk = cur.wptr + p.tapoff[i0 + t];
        tma_load_1d(dst + t * kSlot * 4, wk + a0, (uint32_t)first * 4u, bar);
        if (first < len) tma_load_1d(dst + (t * kSlot + first) * 4, wk, (uint32_t)(len - first) * 4u, bar);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Forward
// ---------------------------------------------------------------------------------------------
template <int kPlane>
__global__ void __launch_bounds__(kMaxThreads, 1) psf_chain_fwd_kernel(const FusedParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kTps = kTpsF, kSlot = kSlotF;
  WarpGeom g;
  make_geom(g, p, kTps);

  float* Vs = reinterpret_cast<float*>(smem_raw);  // [CT][kPlane], rows [0, p.RB) of every plane in use
  constexpr int RB = kPlane;                       // compile-time plane stride -> immediate LDS offsets
  const int S = p.stages;
  float* ring = Vs + CT * RB + (size_t)g.warp * S * kTps * kSlot;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Vs + CT * RB + (size_t)g.nwarps * S * kTps * kSlot) + g.warp * S;

  const uint32_t ring_u32 = smem_u32(ring), bars_u32 = smem_u32(bars);
  if (g.lane == 0) {
    for (int s = 0; s < S; ++s) mbar_init(bars_u32 + s * 8, 1);
    fence_barrier_init();
  }
  __syncwarp();

  WCursor cur;
  cursor_init<false>(cur, p, g);
  for (int s = 0; s < S; ++s) {  // prime the ring
    if (cur.valid) {
      if (g.lane == 0) issue_stage<false, kTps, kSlot>(cur, p, ring_u32 + s * (kTps * kSlot * 4), bars_u32 + s * 8);
      cursor_next<false>(cur, p, g);
    }
  }
  int slot = 0;          // ring slot consumed next (same on all lanes)
  uint32_t parity = 0;   // its mbarrier phase

  const int n = p.n;
  const int64_t plane = (int64_t)CT * n;  // floats per sample per state

  for (int b = g.cid; b < p.B; b += g.ncl) {
    // ---- prologue: V0 block -> shared (channel-major) + state slot 0 --------------------
    float* st0 = p.states + (int64_t)b * plane;  // slot 0 of this sample
    for (int c = 0; c < g.nchunks; ++c) {
      const int tile = c * g.nwarps + g.warp;
      const int j = g.blk_start + tile * kTile + 4 * g.lane;
      if (tile < g.ntiles && j < g.blk_end) {
        float4 pl[CT];
        load_rows_natural(p.Vin + ((int64_t)b * n + j) * CT, pl);
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) {
          *reinterpret_cast<float4*>(Vs + ch * RB + (j - g.blk_start)) = pl[ch];
          st_cg4(st0 + ch * n + j, pl[ch]);
        }
      }
    }
    cluster_sync_all();

    for (int m = 1; m <= p.n_W; ++m) {
      // slot 0 always keeps V_0 (residual operand); without saved states slots 1/2 alternate
      const int slot_prev = (p.save_states || m == 1) ? (m - 1) : 1 + ((m - 1) & 1);
      const int slot_next = p.save_states ? m : 1 + (m & 1);
      const float* prev = p.states + ((int64_t)slot_prev * p.B + b) * plane;
      float* next = p.states + ((int64_t)slot_next * p.B + b) * plane;
      const bool last = (m == p.n_W);

      for (int c = 0; c < g.nchunks; ++c) {
        const int tile = c * g.nwarps + g.warp;
        const bool active = tile < g.ntiles;
        const int row0 = g.blk_start + tile * kTile;
        const int j = row0 + 4 * g.lane;
        const bool mine = active && (j < g.blk_end);
        float4 acc[CT];
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) acc[ch] = make_float4(0.f, 0.f, 0.f, 0.f);

        if (active) {
          // one global-operand tap in flight: issued here, consumed after the next local taps
          bool pend = false;
          float4 pw = make_float4(1.f, 1.f, 1.f, 1.f);
          float4 pv[CT];
          if (mine && p.residual) {  // the residual V_0 rides the pending slot with weight 1
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) pv[ch] = ld_cg4(st0 + ch * n + j);
            pend = true;
          }
          const float* vown = Vs + (j - g.blk_start);
          for (int tg = 0; tg < g.ngroups; ++tg) {
            const int s = slot;
            mbar_wait(bars_u32 + s * 8, parity);
            const float* wst = ring + s * kTps * kSlot + 4 * g.lane;
            if (mine) {
              if (tg == 0 && p.head) {
                // taps with offsets 0,1,2,4 share the quads A = rows j..j+3 and Bq = rows j+4..j+7
                const float4 w0 = lds4(wst), w1 = lds4(wst + kSlot), w2 = lds4(wst + 2 * kSlot),
                             w3 = lds4(wst + 3 * kSlot);
                const bool bq_local = (j + 8 <= g.blk_end);
                const int jb = (j + 4 >= n) ? j + 4 - n : j + 4;
#pragma unroll
                for (int ch = 0; ch < CT; ++ch) {
                  const float4 A = lds4(vown + ch * RB);
                  const float4 Bq = bq_local ? lds4(vown + ch * RB + 4) : ld_cg4(prev + ch * n + jb);
                  // row r: w0[r]*V[j+r] + w1[r]*V[j+r+1] + w2[r]*V[j+r+2] + w3[r]*V[j+r+4]
                  fma4(acc[ch], w0, A);                                   // tap 0
                  fma4(acc[ch], w3, Bq);                                  // tap 4
                  fma2(acc[ch].x, acc[ch].y, w2.x, w2.y, A.z, A.w);       // tap 2, rows 0-1
                  fma2(acc[ch].z, acc[ch].w, w2.z, w2.w, Bq.x, Bq.y);     // tap 2, rows 2-3
                  acc[ch].x = fmaf(w1.x, A.y, acc[ch].x);                 // tap 1
                  acc[ch].y = fmaf(w1.y, A.z, acc[ch].y);
                  acc[ch].z = fmaf(w1.z, A.w, acc[ch].z);
                  acc[ch].w = fmaf(w1.w, Bq.x, acc[ch].w);
                }
              } else {
#pragma unroll
                for (int t = 0; t < kTps; ++t) {
                  const int i = tg * kTps + t;
                  if (i < p.L) {
                    const float4 w4 = lds4(wst + t * kSlot);
                    const int o = p.ooffs.o[i];
                    const int src = j + o;
                    if ((o & 3) == 0) {
                      if (src + 4 <= g.blk_end) {
                        const float* vp = vown + o;
#pragma unroll
                        for (int ch = 0; ch < CT; ++ch) fma4(acc[ch], w4, lds4(vp + ch * RB));
                      } else {
                        if (pend) {
#pragma unroll
                          for (int ch = 0; ch < CT; ++ch) fma4(acc[ch], pw, pv[ch]);
                        }
                        const float* gp = prev + (src >= n ? src - n : src);
                        pw = w4;
#pragma unroll
                        for (int ch = 0; ch < CT; ++ch) pv[ch] = ld_cg4(gp + ch * n);
                        pend = true;
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
                        const float4 X = xl ? lds4(vown + (o - sh) + ch * RB) : ld_cg4(prev + ch * n + ax);
                        const float4 Y = yl ? lds4(vown + (o - sh) + 4 + ch * RB) : ld_cg4(prev + ch * n + ay);
                        fma4(acc[ch], w4, select_quad(X, Y, sh));
                      }
                    }
                  }
                }
              }
            }
            __syncwarp();
            if (++slot == S) {
              slot = 0;
              parity ^= 1u;
            }
            if (cur.valid) {  // refill the slot just drained
              if (g.lane == 0) issue_stage<false, kTps, kSlot>(cur, p, ring_u32 + s * (kTps * kSlot * 4), bars_u32 + s * 8);
              cursor_next<false>(cur, p, g);
            }
          }
          if (pend) {
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) fma4(acc[ch], pw, pv[ch]);
          }
        }
        __syncthreads();  // every read of this chunk's rows (by this chunk and all lower ones) is done
        if (mine) {
          if (!last) {
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) {
              *reinterpret_cast<float4*>(Vs + ch * RB + (j - g.blk_start)) = acc[ch];
              st_cg4(next + ch * n + j, acc[ch]);
            }
          } else {
            store_rows_natural(p.Vout + ((int64_t)b * n + j) * CT, acc);
          }
        }
      }
      cluster_sync_all();  // V_m visible cluster-wide (global) and CTA-wide (shared)
    }
    if (p.n_W == 0) {  // degenerate: out = V0
      for (int c = 0; c < g.nchunks; ++c) {
        const int tile = c * g.nwarps + g.warp;
        const int j = g.blk_start + tile * kTile + 4 * g.lane;
        if (tile < g.ntiles && j < g.blk_end) {
          float4 pl[CT];
#pragma unroll
          for (int ch = 0; ch < CT; ++ch) pl[ch] = lds4(Vs + ch * RB + (j - g.blk_start));
          store_rows_natural(p.Vout + ((int64_t)b * n + j) * CT, pl);
        }
      }
      cluster_sync_all();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Backward
// ---------------------------------------------------------------------------------------------
template <int kPlane>
__global__ void __launch_bounds__(kMaxThreads, 1) psf_chain_bwd_kernel(const FusedParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kTps = kTpsB, kSlot = kSlotB;
  WarpGeom g;
  make_geom(g, p, kTps);

  float* Gs = reinterpret_cast<float*>(smem_raw);  // [CT][kPlane]
  constexpr int RB = kPlane;
  const int S = p.stages;
  float* ring = Gs + CT * RB + (size_t)g.warp * S * kTps * kSlot;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Gs + CT * RB + (size_t)g.nwarps * S * kTps * kSlot) + g.warp * S;

  const uint32_t ring_u32 = smem_u32(ring), bars_u32 = smem_u32(bars);
  if (g.lane == 0) {
    for (int s = 0; s < S; ++s) mbar_init(bars_u32 + s * 8, 1);
    fence_barrier_init();
  }
  __syncwarp();

  WCursor cur;
  cursor_init<true>(cur, p, g);
  for (int s = 0; s < S; ++s) {
    if (cur.valid) {
      if (g.lane == 0) issue_stage<true, kTps, kSlot>(cur, p, ring_u32 + s * (kTps * kSlot * 4), bars_u32 + s * 8);
      cursor_next<true>(cur, p, g);
    }
  }
  int slot = 0;
  uint32_t parity = 0;

  const int n = p.n;
  const int64_t plane = (int64_t)CT * n;

  for (int b = g.cid; b < p.B; b += g.ncl) {
    float* gx0 = p.gx + (int64_t)b * plane;          // exchange slot 0
    float* gx1 = p.gx + ((int64_t)p.B + b) * plane;  // exchange slot 1
    float* racc = p.racc + (int64_t)b * plane;
    // ---- prologue: dOut block -> shared + exchange slot 0 ---------------------------------
    for (int c = 0; c < g.nchunks; ++c) {
      const int tile = c * g.nwarps + g.warp;
      const int j = g.blk_start + tile * kTile + 4 * g.lane;
      if (tile < g.ntiles && j < g.blk_end) {
        float4 pl[CT];
        load_rows_natural(p.Vin + ((int64_t)b * n + j) * CT, pl);
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) {
          *reinterpret_cast<float4*>(Gs + ch * RB + (j - g.blk_start)) = pl[ch];
          st_cg4(gx0 + ch * n + j, pl[ch]);
        }
      }
    }
    cluster_sync_all();

    for (int m = p.n_W; m >= 1; --m) {
      const int step = p.n_W - m;
      const float* gcur = (step & 1) ? gx1 : gx0;
      float* gnext = (step & 1) ? gx0 : gx1;
      const float* vprev = p.states + ((int64_t)(m - 1) * p.B + b) * plane;  // V_{m-1}, channel-major
      float* dwbase = p.dW + (int64_t)(m - 1) * p.sWf + (int64_t)b * p.sWb;
      const bool last = (m == 1);

      for (int c = g.nchunks - 1; c >= 0; --c) {
        const int tile = c * g.nwarps + g.warp;
        const bool active = tile < g.ntiles;
        const int row0 = g.blk_start + tile * kTile;
        const int j = row0 + 4 * g.lane;
        const bool mine = active && (j < g.blk_end);
        float4 acc[CT];
#pragma unroll
        for (int ch = 0; ch < CT; ++ch) acc[ch] = make_float4(0.f, 0.f, 0.f, 0.f);
        float4 rsum[CT];  // residual: racc so far, loaded before the barrier, used after it

        if (active) {
          float4 vown[CT];
          if (mine) {
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) vown[ch] = ld_cg4(vprev + ch * n + j);
          }
          const float* gown = Gs + (j - g.blk_start);
          for (int tg = 0; tg < g.ngroups; ++tg) {
            const int s = slot;
            mbar_wait(bars_u32 + s * 8, parity);
            const float* wst = ring + s * kTps * kSlot + 4 * g.lane;
            if (mine) {
#pragma unroll
              for (int t = 0; t < kTps; ++t) {
                const int i = tg * kTps + t;
                if (i < p.L) {
                  const int o = p.ooffs.o[i];
                  const int src = j - o;  // first source row (may be negative: cyclic)
                  float* dwk = dwbase + p.tapoff[i];
                  if ((o & 3) == 0) {
                    const float4 w4 = lds4(wst + t * kSlot);
                    float4 gv[CT];
                    int sw;
                    if (src >= g.blk_start) {
                      sw = src;
#pragma unroll
                      for (int ch = 0; ch < CT; ++ch) gv[ch] = lds4(gown - o + ch * RB);
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
                    // unaligned tap: aligned quads X = rows a..a+3, Y = rows a+4..a+7 cover src..src+3
                    const int sh = (4 - (o & 3)) & 3;  // src mod 4 (j is a multiple of 4)
                    const int a = src - sh;
                    const bool xl = (a >= g.blk_start), yl = (a + 4 >= g.blk_start);
                    const int ax = a < 0 ? a + n : a;
                    const int ay = (a + 4 < 0) ? a + 4 + n : a + 4;
                    const float4 w4 = select_quad(lds4(wst + t * kSlot), lds4(wst + t * kSlot + 4), sh);
                    float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int ch = 0; ch < CT; ++ch) {
                      const float4 X = xl ? lds4(gown - o - sh + ch * RB) : ld_cg4(gcur + ch * n + ax);
                      const float4 Y = yl ? lds4(gown - o - sh + 4 + ch * RB) : ld_cg4(gcur + ch * n + ay);
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
            if (++slot == S) {
              slot = 0;
              parity ^= 1u;
            }
            if (cur.valid) {
              if (g.lane == 0) issue_stage<true, kTps, kSlot>(cur, p, ring_u32 + s * (kTps * kSlot * 4), bars_u32 + s * 8);
              cursor_next<true>(cur, p, g);
            }
          }
          // residual: fetch the running sum now, its latency hides behind the barrier
          if (mine && p.residual && m != p.n_W) {
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) rsum[ch] = ld_cg4(racc + ch * n + j);
          }
        }
        __syncthreads();
        if (mine) {
          // residual: dV0 also receives G_m of every factor (V_m = W_m V_{m-1} + V_0)
          if (p.residual) {
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) {
              float4 own = lds4(Gs + ch * RB + (j - g.blk_start));
              if (m != p.n_W) add4(own, rsum[ch]);
              if (!last) {
                st_cg4(racc + ch * n + j, own);
              } else {
                add4(acc[ch], own);
              }
            }
          }
          if (!last) {
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) {
              *reinterpret_cast<float4*>(Gs + ch * RB + (j - g.blk_start)) = acc[ch];
              st_cg4(gnext + ch * n + j, acc[ch]);
            }
          } else {
            store_rows_natural(p.Vout + ((int64_t)b * n + j) * CT, acc);
          }
        }
      }
      cluster_sync_all();
    }
    if (p.n_W == 0) {
      for (int c = 0; c < g.nchunks; ++c) {
        const int tile = c * g.nwarps + g.warp;
        const int j = g.blk_start + tile * kTile + 4 * g.lane;
        if (tile < g.ntiles && j < g.blk_end) {
          float4 pl[CT];
#pragma unroll
          for (int ch = 0; ch < CT; ++ch) pl[ch] = lds4(Gs + ch * RB + (j - g.blk_start));
          store_rows_natural(p.Vout + ((int64_t)b * n + j) * CT, pl);
        }
      }
      cluster_sync_all();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
int plane_for(int RB) { return RB <= 512 ? 512 : (RB <= 2048 ? 2048 : 4096); }

int smem_bytes_for(int RB, int threads, int stages, bool backward) {
  const int nwarps = threads / 32;
  const int ring = backward ? stages * kTpsB * kSlotB : stages * kTpsF * kSlotF;
  return CT * plane_for(RB) * 4 + nwarps * ring * 4 + nwarps * stages * 8;
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

bool g_attr_done = false;

struct OccCache {
  int P, threads, smem, backward, clusters;
  const void* kernel;
};
OccCache g_occ[16];
int g_occ_n = 0;

int resident_clusters(const void* kernel, int P, int threads, int smem, int backward, int sms) {
  for (int i = 0; i < g_occ_n; ++i)
    if (g_occ[i].P == P && g_occ[i].threads == threads && g_occ[i].smem == smem && g_occ[i].backward == backward &&
        g_occ[i].kernel == kernel)
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
  if (g_occ_n < 16) g_occ[g_occ_n++] = {P, threads, smem, backward, nclusters, kernel};
  return nclusters;
}

// Tap visiting order.  Forward: the {0,1,2,4} head first (if present), then big offsets interleaved
// 1:3 with small ones so every (possibly remote) big tap is followed by local work.  Backward: natural.
void make_order(const PsfChainArgs& a, bool backward, PsfOffsets* order, int* head) {
  const int L = a.L;
  *head = 0;
  for (int i = 0; i < PSF_MAX_TAPS; ++i) order->o[i] = i < L ? i : 0;
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
    for (int w = 0; w < 4; ++w) order->o[pos++] = hk[w];
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
    if (bigs < nb) {
      order->o[pos++] = rest[hi--];
      ++bigs;
    }
    for (int q = 0; q < 3 && lo <= hi; ++q) order->o[pos++] = rest[lo++];
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
  fp.sWt = a.ws.tap;
  fp.B = (int)a.B;
  fp.n = (int)a.n;
  fp.n_W = a.n_W;
  fp.L = a.L;
  fp.residual = a.residual;
  fp.P = plan.cluster;
  fp.RB = plan.rows_per_cta;
  fp.stages = plan.stages;
  fp.offs = a.offs;
  make_order(a, backward, &fp.order, &fp.head);
  for (int i = 0; i < PSF_MAX_TAPS; ++i) {
    fp.ooffs.o[i] = a.offs.o[fp.order.o[i]];
    fp.tapoff[i] = (long long)fp.order.o[i] * a.ws.tap;
  }
}

}  // namespace

int psf_fused_plan(const PsfChainArgs& a, PsfFusedPlan* plan) {
  if (a.C != CT || a.n < 4 || (a.n & 3) || a.L < 1 || a.L > PSF_MAX_TAPS || a.n_W < 0 || a.B < 1 ||
      a.B > (1 << 20))
    return PSF_ERR_UNSUPPORTED;
  if (a.ws.row != 1 || (a.ws.factor & 3) || (a.ws.sample & 3) || (a.ws.tap & 3)) return PSF_ERR_UNSUPPORTED;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return PSF_ERR_UNSUPPORTED;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // block rows: 2048 (64 KB, two CTAs per SM) unless that needs more than 8 CTAs per sample
  int max_rows = env_int("PSF_FUSED_MAX_ROWS", 2048);
  int max_threads = env_int("PSF_FUSED_MAX_THREADS", 256);
  if (max_threads > kMaxThreads) max_threads = kMaxThreads;
  max_threads &= ~31;
  if (max_threads < 32) max_threads = 32;
  int P = 1;
  while ((a.n + P - 1) / P > max_rows && P < 8) P *= 2;
  if ((a.n + P - 1) / P > max_rows) {
    max_rows = 4096;
    max_threads = kMaxThreads;
    if ((a.n + P - 1) / P > max_rows) return PSF_ERR_UNSUPPORTED;
  }
  // spread small batches over more SMs while a CTA keeps at least 4 warp tiles
  while (P < 8 && a.B * P * 2 <= 2 * sms && (a.n + 2 * P - 1) / (2 * P) >= 4 * kTile) P *= 2;
  const int per = (int)((a.n + P - 1) / P);
  const int RB = (per + 3) & ~3;
  int threads = ((RB + kTile - 1) / kTile) * 32;
  if (threads > max_threads) threads = max_threads;
  plan->cluster = P;
  plan->rows_per_cta = RB;
  plan->threads = threads;
  int stages = env_int("PSF_FUSED_STAGES", 3);
  if (stages > kMaxStages) stages = kMaxStages;
  if (stages < 2) stages = 2;
  while (stages > 2 && smem_bytes_for(RB, threads, stages, true) > 227 * 1024) --stages;
  if (smem_bytes_for(RB, threads, stages, true) > 227 * 1024) return PSF_ERR_UNSUPPORTED;
  plan->stages = stages;
  plan->smem_bytes = smem_bytes_for(RB, threads, stages, false);
  plan->smem_bytes_bwd = smem_bytes_for(RB, threads, stages, true);

  if (!g_attr_done) {
    const int planes[3] = {512, 2048, 4096};
    for (int q = 0; q < 3; ++q)
      if (cudaFuncSetAttribute(fwd_kernel_for(planes[q]), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
              cudaSuccess ||
          cudaFuncSetAttribute(bwd_kernel_for(planes[q]), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) !=
              cudaSuccess)
        return PSF_ERR_UNSUPPORTED;
    g_attr_done = true;
  }
  plan->resident = resident_clusters(fwd_kernel_for(RB), P, threads, plan->smem_bytes, 0, sms);
  plan->resident_bwd = resident_clusters(bwd_kernel_for(RB), P, threads, plan->smem_bytes_bwd, 1, sms);
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
  const int64_t S = a.B * a.n * a.C;
  FusedParams fp = {};
  fp.W = W;
  fp.dW = dW;
  fp.Vin = dOut;
  fp.Vout = dV0;
  fp.states = ws;
  fp.gx = ws + (int64_t)a.n_W * S;
  fp.racc = fp.gx + 2 * S;
  fp.save_states = 1;
  fill_common(fp, a, plan, true);
  return launch(bwd_kernel_for(plan.rows_per_cta), fp, plan.cluster, plan.threads, plan.smem_bytes_bwd,
                plan.resident_bwd, st);
}
