// Fused chord-factor chain for sm_100a (B200): forward and backward, fp32, C = 8 channels.
//
// One thread-block CLUSTER owns one sample.  CTA q of a P-CTA cluster owns the contiguous row
// block [q*RB, (q+1)*RB) for ALL channels and keeps it in shared memory, channel-major
// ([C][RB] fp32, <= 128 KB), across all n_W factors.  Facts the layout is built on:
//   * chord links only point forward (row i reads rows i+o_k, o_k >= 0), so a block can be
//     updated IN PLACE by an ascending sweep over 128-row warp tiles (descending for the
//     backward pass, whose links point backward): once a tile has been read by itself and by
//     every lower tile, nobody inside the CTA needs its old value again.
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
//     are 128-bit and conflict free (consecutive lanes -> consecutive 16 B), taps whose offset
//     is not a multiple of 4 (1 and 2) fall back to 32-bit accesses at the same byte cost.
//   * backward (output-stationary at row j):  one shifted read of G_m[j-o_k,:] feeds both
//     dV_{m-1}[j,:] += W_m[j-o_k,k] * G_m[j-o_k,:]  and  dW_m[j-o_k,k] = <G_m[j-o_k,:], V_{m-1}[j,:]>,
//     all C channels live in one thread so dW needs no cross-CTA reduction and is written once.
//
// Reference semantics: SyntheticExperiments/psf.py:167-188 (loop + residual), spmul_cuda.cu:24,
// :79-80, :105-108 (per-factor formulas).
#include "psf_internal.cuh"

namespace {

constexpr int CT = 8;            // channels held per thread == C of the fast path
constexpr int kTile = 128;       // rows per warp tile (32 lanes x 4 rows)
constexpr int kStages = 3;       // W ring depth per warp
constexpr int kTps = 3;          // taps per ring stage
constexpr int kSlot = kTile + 8; // floats per tap slot (8 spare: shifted, 16 B aligned copies in bwd)
constexpr int kMaxThreads = 512;
constexpr int kMaxBlockRows = 4096;  // 8 ch * 4096 rows * 4 B = 128 KB

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
  PsfOffsets offs;
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
__device__ __forceinline__ float ld_cg1(const float* p) { return __ldcg(p); }
__device__ __forceinline__ void st_cg4(float* p, float4 v) { __stcg(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ void st_cs4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

__device__ __forceinline__ void fma4(float4& acc, const float4& a, const float4& b) {
  acc.x = fmaf(a.x, b.x, acc.x);
  acc.y = fmaf(a.y, b.y, acc.y);
  acc.z = fmaf(a.z, b.z, acc.z);
  acc.w = fmaf(a.w, b.w, acc.w);
}
__device__ __forceinline__ float& f4at(float4& v, int r) { return reinterpret_cast<float*>(&v)[r]; }

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
template <bool kBackward>
struct WCursor {
  int b, m, c, tg;
  bool valid;
};

struct WarpGeom {
  int warp, lane, nwarps;
  int blk_start, blk_end;  // rows of this CTA
  int ntiles;              // 128-row tiles in the block
  int nchunks;             // ceil(ntiles / nwarps)
  int ngroups;             // ceil(L / kTps)
  int cid, ncl;            // cluster id, number of clusters
};

template <bool kBackward>
__device__ __forceinline__ int last_chunk_of_warp(const WarpGeom& g) {
  // largest c with c*nwarps + warp < ntiles (>= 0 because the warp has work)
  return (g.ntiles - 1 - g.warp) / g.nwarps;
}

template <bool kBackward>
__device__ __forceinline__ void cursor_init(WCursor<kBackward>& cur, const FusedParams& p, const WarpGeom& g) {
  cur.b = g.cid;
  cur.tg = 0;
  cur.valid = (g.warp < g.ntiles) && (cur.b < p.B) && (p.n_W > 0);
  if (kBackward) {
    cur.m = p.n_W;
    cur.c = cur.valid ? last_chunk_of_warp<kBackward>(g) : 0;
  } else {
    cur.m = 1;
    cur.c = 0;
  }
}

template <bool kBackward>
__device__ __forceinline__ void cursor_next(WCursor<kBackward>& cur, const FusedParams& p, const WarpGeom& g) {
  if (++cur.tg < g.ngroups) return;
  cur.tg = 0;
  if (kBackward) {
    if (--cur.c >= 0) return;
    cur.c = last_chunk_of_warp<kBackward>(g);
    if (--cur.m >= 1) return;
    cur.m = p.n_W;
  } else {
    ++cur.c;
    if (cur.c * g.nwarps + g.warp < g.ntiles) return;
    cur.c = 0;
    if (++cur.m <= p.n_W) return;
    cur.m = 1;
  }
  cur.b += g.ncl;
  cur.valid = cur.b < p.B;
}

// Lane 0: arm the stage barrier and launch the bulk copies of one stage (<= kTps taps).
template <bool kBackward>
__device__ __forceinline__ void issue_stage(const WCursor<kBackward>& cur, const FusedParams& p, const WarpGeom& g,
                                            float* slot_base, uint32_t bar) {
  const int tile = cur.c * g.nwarps + g.warp;
  const int row0 = g.blk_start + tile * kTile;
  const int rw = min(kTile, g.blk_end - row0);
  const int k0 = cur.tg * kTps;
  const int kn = min(kTps, p.L - k0);
  const float* wbase = p.W + (int64_t)(cur.m - 1) * p.sWf + (int64_t)cur.b * p.sWb;
  // bytes first (expect_tx must precede the copies)
  uint32_t total = 0;
  for (int t = 0; t < kn; ++t) {
    const int o = p.offs.o[k0 + t];
    total += (uint32_t)((kBackward && (o & 3)) ? rw + 4 : rw) * 4u;
  }
  mbar_expect_tx(bar, total);
  for (int t = 0; t < kn; ++t) {
    const int k = k0 + t;
    const float* wk = wbase + (int64_t)k * p.sWt;
    const uint32_t dst = smem_u32(slot_base + t * kSlot);
    if (!kBackward) {
      tma_load_1d(dst, wk + row0, (uint32_t)rw * 4u, bar);
    } else {
      const int o = p.offs.o[k];
      int s0 = row0 - o;
      if (s0 < 0) s0 += p.n;
      const int a0 = s0 & ~3;                      // 16 B aligned start
      const int len = (o & 3) ? rw + 4 : rw;       // floats
      const int first = min(len, p.n - a0);        // before the cyclic wrap
      tma_load_1d(dst, wk + a0, (uint32_t)first * 4u, bar);
      if (first < len) tma_load_1d(dst + (uint32_t)first * 4u, wk, (uint32_t)(len - first) * 4u, bar);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Forward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMaxThreads, 1) psf_chain_fwd_kernel(const FusedParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpGeom g;
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
  g.ngroups = (p.L + kTps - 1) / kTps;

  float* Vs = reinterpret_cast<float*>(smem_raw);  // [CT][RB]
  const int RB = p.RB;
  float* ring = Vs + CT * RB + (size_t)g.warp * kStages * kTps * kSlot;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Vs + CT * RB + (size_t)g.nwarps * kStages * kTps * kSlot) + g.warp * kStages;

  if (g.lane == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
  }
  __syncwarp();

  // prime the ring
  WCursor<false> cur;
  cursor_init(cur, p, g);
  int issued = 0;
  for (int s = 0; s < kStages; ++s) {
    if (cur.valid) {
      if (g.lane == 0) issue_stage(cur, p, g, ring + s * kTps * kSlot, smem_u32(&bars[s]));
      cursor_next(cur, p, g);
      ++issued;
    }
  }
  int consumed = 0;  // stages consumed by this warp (same on all lanes)

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
          st_cg4(st0 + (int64_t)ch * n + j, pl[ch]);
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
          for (int tg = 0; tg < g.ngroups; ++tg) {
            const int s = consumed % kStages;
            mbar_wait(smem_u32(&bars[s]), (uint32_t)((consumed / kStages) & 1));
            const float* wst = ring + s * kTps * kSlot;
            if (mine) {
#pragma unroll
              for (int t = 0; t < kTps; ++t) {
                const int k = tg * kTps + t;
                if (k < p.L) {
                  const float4 w4 = *reinterpret_cast<const float4*>(wst + t * kSlot + 4 * g.lane);
                  const int o = p.offs.o[k];
                  const int src = j + o;
                  if ((o & 3) == 0) {
                    if (src + 4 <= g.blk_end) {
                      const float* vp = Vs + (src - g.blk_start);
#pragma unroll
                      for (int ch = 0; ch < CT; ++ch)
                        fma4(acc[ch], w4, *reinterpret_cast<const float4*>(vp + ch * RB));
                    } else {
                      const int sw = src >= n ? src - n : src;
                      const float* gp = prev + sw;
                      float4 v[CT];
#pragma unroll
                      for (int ch = 0; ch < CT; ++ch) v[ch] = ld_cg4(gp + (int64_t)ch * n);
#pragma unroll
                      for (int ch = 0; ch < CT; ++ch) fma4(acc[ch], w4, v[ch]);
                    }
                  } else {
                    if (src + 4 <= g.blk_end) {
                      const float* vp = Vs + (src - g.blk_start);
#pragma unroll
                      for (int ch = 0; ch < CT; ++ch) {
                        acc[ch].x = fmaf(w4.x, vp[ch * RB + 0], acc[ch].x);
                        acc[ch].y = fmaf(w4.y, vp[ch * RB + 1], acc[ch].y);
                        acc[ch].z = fmaf(w4.z, vp[ch * RB + 2], acc[ch].z);
                        acc[ch].w = fmaf(w4.w, vp[ch * RB + 3], acc[ch].w);
                      }
                    } else {
#pragma unroll
                      for (int r = 0; r < 4; ++r) {
                        const int sr = src + r;
                        const float wr = (r == 0) ? w4.x : (r == 1) ? w4.y : (r == 2) ? w4.z : w4.w;
                        if (sr < g.blk_end) {
                          const float* vp = Vs + (sr - g.blk_start);
#pragma unroll
                          for (int ch = 0; ch < CT; ++ch) f4at(acc[ch], r) = fmaf(wr, vp[ch * RB], f4at(acc[ch], r));
                        } else {
                          const float* gp = prev + (sr >= n ? sr - n : sr);
#pragma unroll
                          for (int ch = 0; ch < CT; ++ch)
                            f4at(acc[ch], r) = fmaf(wr, ld_cg1(gp + (int64_t)ch * n), f4at(acc[ch], r));
                        }
                      }
                    }
                  }
                }
              }
            }
            __syncwarp();
            ++consumed;
            if (cur.valid) {  // refill the slot just drained
              if (g.lane == 0) issue_stage(cur, p, g, ring + s * kTps * kSlot, smem_u32(&bars[s]));
              cursor_next(cur, p, g);
              ++issued;
            }
          }
          if (mine && p.residual) {
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) {
              const float4 r4 = ld_cg4(st0 + (int64_t)ch * n + j);
              acc[ch].x += r4.x;
              acc[ch].y += r4.y;
              acc[ch].z += r4.z;
              acc[ch].w += r4.w;
            }
          }
        }
        __syncthreads();  // every read of this chunk's rows (by this chunk and all lower ones) is done
        if (mine) {
          if (!last) {
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) {
              *reinterpret_cast<float4*>(Vs + ch * RB + (j - g.blk_start)) = acc[ch];
              st_cg4(next + (int64_t)ch * n + j, acc[ch]);
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
          for (int ch = 0; ch < CT; ++ch) pl[ch] = *reinterpret_cast<float4*>(Vs + ch * RB + (j - g.blk_start));
          store_rows_natural(p.Vout + ((int64_t)b * n + j) * CT, pl);
        }
      }
      cluster_sync_all();
    }
  }
  (void)issued;
}

// ---------------------------------------------------------------------------------------------
// Backward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kMaxThreads, 1) psf_chain_bwd_kernel(const FusedParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  WarpGeom g;
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
  g.ngroups = (p.L + kTps - 1) / kTps;

  float* Gs = reinterpret_cast<float*>(smem_raw);  // [CT][RB]
  const int RB = p.RB;
  float* ring = Gs + CT * RB + (size_t)g.warp * kStages * kTps * kSlot;
  uint64_t* bars = reinterpret_cast<uint64_t*>(Gs + CT * RB + (size_t)g.nwarps * kStages * kTps * kSlot) + g.warp * kStages;

  if (g.lane == 0) {
    for (int s = 0; s < kStages; ++s) mbar_init(smem_u32(&bars[s]), 1);
    fence_barrier_init();
  }
  __syncwarp();

  WCursor<true> cur;
  cursor_init(cur, p, g);
  for (int s = 0; s < kStages; ++s) {
    if (cur.valid) {
      if (g.lane == 0) issue_stage(cur, p, g, ring + s * kTps * kSlot, smem_u32(&bars[s]));
      cursor_next(cur, p, g);
    }
  }
  int consumed = 0;

  const int n = p.n;
  const int64_t plane = (int64_t)CT * n;

  for (int b = g.cid; b < p.B; b += g.ncl) {
    float* gx0 = p.gx + (int64_t)b * plane;                    // exchange slot 0
    float* gx1 = p.gx + ((int64_t)p.B + b) * plane;            // exchange slot 1
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
          st_cg4(gx0 + (int64_t)ch * n + j, pl[ch]);
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

        if (active) {
          float4 vown[CT];
          if (mine) {
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) vown[ch] = ld_cg4(vprev + (int64_t)ch * n + j);
          }
          for (int tg = 0; tg < g.ngroups; ++tg) {
            const int s = consumed % kStages;
            mbar_wait(smem_u32(&bars[s]), (uint32_t)((consumed / kStages) & 1));
            const float* wst = ring + s * kTps * kSlot;
            if (mine) {
#pragma unroll
              for (int t = 0; t < kTps; ++t) {
                const int k = tg * kTps + t;
                if (k < p.L) {
                  const int o = p.offs.o[k];
                  const int src = j - o;  // first source row (may be negative: cyclic)
                  float* dwk = dwbase + (int64_t)k * p.sWt;
                  if ((o & 3) == 0) {
                    const float4 w4 = *reinterpret_cast<const float4*>(wst + t * kSlot + 4 * g.lane);
                    float4 gv[CT];
                    int sw;
                    if (src >= g.blk_start) {
                      sw = src;
                      const float* gp = Gs + (src - g.blk_start);
#pragma unroll
                      for (int ch = 0; ch < CT; ++ch) gv[ch] = *reinterpret_cast<const float4*>(gp + ch * RB);
                    } else {
                      sw = src < 0 ? src + n : src;
                      const float* gp = gcur + sw;
#pragma unroll
                      for (int ch = 0; ch < CT; ++ch) gv[ch] = ld_cg4(gp + (int64_t)ch * n);
                    }
                    float4 dw = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int ch = 0; ch < CT; ++ch) {
                      fma4(acc[ch], w4, gv[ch]);
                      fma4(dw, gv[ch], vown[ch]);
                    }
                    st_cs4(dwk + sw, dw);
                  } else {
                    // unaligned tap: 32-bit accesses, per-row local/remote decision
                    int s0 = row0 - o;
                    if (s0 < 0) s0 += n;
                    const int lead = s0 & 3;
                    const float* wrow = wst + t * kSlot + lead + 4 * g.lane;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                      const int sr = src + r;
                      const float wr = wrow[r];
                      float dwr = 0.f;
                      int swr;
                      if (sr >= g.blk_start) {
                        swr = sr;
                        const float* gp = Gs + (sr - g.blk_start);
#pragma unroll
                        for (int ch = 0; ch < CT; ++ch) {
                          const float gvv = gp[ch * RB];
                          f4at(acc[ch], r) = fmaf(wr, gvv, f4at(acc[ch], r));
                          dwr = fmaf(gvv, f4at(vown[ch], r), dwr);
                        }
                      } else {
                        swr = sr < 0 ? sr + n : sr;
                        const float* gp = gcur + swr;
#pragma unroll
                        for (int ch = 0; ch < CT; ++ch) {
                          const float gvv = ld_cg1(gp + (int64_t)ch * n);
                          f4at(acc[ch], r) = fmaf(wr, gvv, f4at(acc[ch], r));
                          dwr = fmaf(gvv, f4at(vown[ch], r), dwr);
                        }
                      }
                      __stcs(dwk + swr, dwr);
                    }
                  }
                }
              }
            }
            __syncwarp();
            ++consumed;
            if (cur.valid) {
              if (g.lane == 0) issue_stage(cur, p, g, ring + s * kTps * kSlot, smem_u32(&bars[s]));
              cursor_next(cur, p, g);
            }
          }
          // residual: dV0 also receives G_m of every factor (V_m = W_m V_{m-1} + V_0)
          if (mine && p.residual) {
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) {
              float4 own = *reinterpret_cast<const float4*>(Gs + ch * RB + (j - g.blk_start));
              if (m != p.n_W) {
                const float4 r4 = ld_cg4(racc + (int64_t)ch * n + j);
                own.x += r4.x;
                own.y += r4.y;
                own.z += r4.z;
                own.w += r4.w;
              }
              if (!last) {
                st_cg4(racc + (int64_t)ch * n + j, own);
              } else {
                acc[ch].x += own.x;
                acc[ch].y += own.y;
                acc[ch].z += own.z;
                acc[ch].w += own.w;
              }
            }
          }
        }
        __syncthreads();
        if (mine) {
          if (!last) {
#pragma unroll
            for (int ch = 0; ch < CT; ++ch) {
              *reinterpret_cast<float4*>(Gs + ch * RB + (j - g.blk_start)) = acc[ch];
              st_cg4(gnext + (int64_t)ch * n + j, acc[ch]);
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
          for (int ch = 0; ch < CT; ++ch) pl[ch] = *reinterpret_cast<float4*>(Gs + ch * RB + (j - g.blk_start));
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
int smem_bytes_for(int RB, int threads) {
  const int nwarps = threads / 32;
  return CT * RB * 4 + nwarps * kStages * kTps * kSlot * 4 + nwarps * kStages * 8;
}

int g_attr_done_fwd = 0, g_attr_done_bwd = 0;

int launch(const void* kernel, const FusedParams& fp, const PsfFusedPlan& plan, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  const int clusters = plan.resident < fp.B ? plan.resident : fp.B;
  cfg.gridDim = dim3((unsigned)(clusters * plan.cluster), 1, 1);
  cfg.blockDim = dim3((unsigned)plan.threads, 1, 1);
  cfg.dynamicSmemBytes = (size_t)plan.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)plan.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  void* args[] = {(void*)&fp};
  return (int)cudaLaunchKernelExC(&cfg, kernel, args);
}

}  // namespace

int psf_fused_plan(const PsfChainArgs& a, PsfFusedPlan* plan) {
  if (a.C != CT || a.n < 4 || (a.n & 3) || a.n > 8 * (int64_t)kMaxBlockRows || a.L < 1 || a.L > PSF_MAX_TAPS ||
      a.n_W < 0 || a.B < 1 || a.B > (1 << 20))
    return PSF_ERR_UNSUPPORTED;
  if (a.ws.row != 1 || (a.ws.factor & 3) || (a.ws.sample & 3) || (a.ws.tap & 3)) return PSF_ERR_UNSUPPORTED;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return PSF_ERR_UNSUPPORTED;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int P = 1;
  while ((a.n + P - 1) / P > kMaxBlockRows) P *= 2;
  // spread small batches over more SMs while a CTA still has at least 4 warp tiles
  while (P < 8 && a.B * P * 2 <= sms && (a.n + 2 * P - 1) / (2 * P) >= 4 * kTile) P *= 2;
  const int per = (int)((a.n + P - 1) / P);
  const int RB = (per + 3) & ~3;
  int threads = ((RB + kTile - 1) / kTile) * 32;
  if (threads > kMaxThreads) threads = kMaxThreads;
  plan->cluster = P;
  plan->rows_per_cta = RB;
  plan->threads = threads;
  plan->smem_bytes = smem_bytes_for(RB, threads);

  if (!g_attr_done_fwd) {
    if (cudaFuncSetAttribute((const void*)psf_chain_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024) != cudaSuccess)
      return PSF_ERR_UNSUPPORTED;
    g_attr_done_fwd = 1;
  }
  if (!g_attr_done_bwd) {
    if (cudaFuncSetAttribute((const void*)psf_chain_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             227 * 1024) != cudaSuccess)
      return PSF_ERR_UNSUPPORTED;
    g_attr_done_bwd = 1;
  }
  // how many clusters of this shape are resident at once (the kernels are persistent over samples)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(P * sms), 1, 1);
  cfg.blockDim = dim3((unsigned)threads, 1, 1);
  cfg.dynamicSmemBytes = (size_t)plan->smem_bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)P;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int nclusters = 0;
  if (cudaOccupancyMaxActiveClusters(&nclusters, (const void*)psf_chain_bwd_kernel, &cfg) != cudaSuccess ||
      nclusters < 1) {
    cudaGetLastError();
    return PSF_ERR_UNSUPPORTED;
  }
  plan->resident = nclusters;
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
  fp.sWf = a.ws.factor;
  fp.sWb = a.ws.sample;
  fp.sWt = a.ws.tap;
  fp.B = (int)a.B;
  fp.n = (int)a.n;
  fp.n_W = a.n_W;
  fp.L = a.L;
  fp.residual = a.residual;
  fp.save_states = save_states;
  fp.P = plan.cluster;
  fp.RB = plan.rows_per_cta;
  fp.offs = a.offs;
  return launch((const void*)psf_chain_fwd_kernel, fp, plan, st);
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
  fp.sWf = a.ws.factor;
  fp.sWb = a.ws.sample;
  fp.sWt = a.ws.tap;
  fp.B = (int)a.B;
  fp.n = (int)a.n;
  fp.n_W = a.n_W;
  fp.L = a.L;
  fp.residual = a.residual;
  fp.save_states = 1;
  fp.P = plan.cluster;
  fp.RB = plan.rows_per_cta;
  fp.offs = a.offs;
  return launch((const void*)psf_chain_bwd_kernel, fp, plan, st);
}
