// Backward pass of the generators, every GEMM-shaped piece on tcgen05 (second version; the first one,
// psf_gen_tc_bwd_sm100.cu, keeps the parameter-gradient products on the SIMT / mma.sync pipes and serves ragged shapes).
//
// Per 128-row tile and generator m = Linear(32,32) -> GELU -> Linear(32,L) (SyntheticExperiments/psf.py:110-126):
//
//   row-local products, A operand in TENSOR MEMORY with lane = row (B = weight images in shared memory, streamed by TMA):
//     pre  [128 x 32] = x    [128 x 32] . W1^T
//     dhid [128 x 32] = dout [128 x 16] . W2
//     dx   [128 x 32] += dpre[128 x 32] . W1                     (one accumulator over the tile's generators)
//   parameter gradients, contraction over the tile's ROWS - so the rows are the K dimension, which lives in the COLUMNS of
//   a tensor-memory A operand and in the rows of an MN-major shared-memory B operand:
//     D[lane][n] += sum_r A_T[lane][r] . B[r][n],   A_T lanes = [ x^T (32 lanes: e) ; dout_m^T (16 lanes: l) ; ones ],
//                                                   B cols   = [ dpre_m (32: h) | hid_m (32: h) ]
//     -> D[e][h] = dW1^T, D[32 + l][32 + h] = dW2, D[48][h] = db1   (db2 = column sums of dout, done by the A_T builder)
//   49 of the 128 lanes carry results; batching four generators into the lanes instead (dpre^T as A) needs the
//   producers' rows transposed and four generators resident at once, which fits neither shared nor tensor memory.
//   The legacy warp-level mma.sync path this replaces costs 16 cycles per m16n8k8 TF32 here (three per product).
//
// x^T stays in the A_T lanes 0..31 for the whole tile; dout^T is rewritten per generator from a shared-memory staging
// tile that TMA fills from the tap-major dW (512 contiguous bytes per tap), which also feeds the epilogue's dout reads.
// B is written by the epilogue warps in the canonical MN-major tf32 layout (128-byte rows of 32 columns, 32-byte chunks
// XOR-swizzled with the row index, layout type "128B swizzle, 32B base"; probed in tools/microbench/umma_probe.cu).
// Every product is error-compensated 3xTF32, small terms first (psf_gen_tc_sm100.cu).
//
// Warp roles (736 threads, one CTA per SM, persistent over tiles):
//   warps 0-15   epilogue: thread = (row, 8 hidden units); tcgen05.ld pre / dhid, Phi / phi, dpre -> tensor memory,
//                [dpre | hid] (hi, lo) -> shared memory B operand, dout -> tensor memory A operand of dhid
//   warps 16-19  helpers, one per 32-lane quarter: x -> tensor memory / shared memory, dx writer;
//                warp 16 also builds x^T and drains D lanes 0..31; warps 17 / 18 alternate building dout^T (+ db2) in the
//                lanes 32..48 / 64..80 and draining them; warp 19 stages the V generator's dV0 (kDsStages steps ahead)
//   warp 20      tensor-memory allocation + the MMA-issuing thread
//   warp 21      TMA: weight images (3 stages);  warp 22  TMA: dout staging tiles (4 stages)
#include "psf_gen.cuh"
#include "psf_tc.cuh"

namespace {

// PSF_TC_DEBUG builds record a timeline of block 0 (event, step, SM clock) that tools/gen_timeline.py prints
#ifdef PSF_TC_DEBUG
constexpr int kTlPerWarp = 2048;
__device__ long long g_timeline[24 * kTlPerWarp * 2];
#define TL_DECL unsigned tl_n = 0
#define TL(ev, step)                                                                               \
  do {                                                                                             \
    if (blockIdx.x == 0 && tl_n < kTlPerWarp) {                                                    \
      long long* q_ = g_timeline + ((size_t)(threadIdx.x >> 5) * kTlPerWarp + tl_n) * 2;           \
      q_[0] = (long long)(ev) * 100000 + (long long)(step) + 1;                                    \
      q_[1] = clock64();                                                                           \
      ++tl_n;                                                                                      \
    }                                                                                              \
  } while (0)
#else
#define TL_DECL
#define TL(ev, step)
#endif

constexpr int kEpi = 16, kHelp0 = 16, kMmaW = 20, kTmaW = 21, kTmaD = 22;
constexpr int kThreads2 = (kTmaD + 1) * 32;
// weight image of one generator (same as psf_gen_tc_bwd_sm100.cu)
constexpr uint32_t kW1 = 0, kW1T = 8192, kW2T = 16384, kB1 = 20480, kImg = 20736;
constexpr int kWStages = 3;
constexpr int kDsStages = 4;
// shared memory map (bytes)
constexpr uint32_t kOffB = 0;                                // B operand: 2 buffers x (hi: dpre 16K, hid 16K | lo: 16K, 16K)
constexpr uint32_t kBBuf = 65536;
constexpr uint32_t kOffW = kOffB + 2 * kBBuf;                // weight images, kWStages stages
constexpr uint32_t kOffDs = kOffW + kWStages * kImg;         // dout staging: kDsStages x [16 taps][128 rows] floats
constexpr uint32_t kDsBuf = 16 * 128 * 4;
constexpr uint32_t kOffBar2 = kOffDs + kDsStages * kDsBuf;
constexpr uint32_t kSmem2 = kOffBar2 + 384;
static_assert(kSmem2 <= 227 * 1024, "shared memory");
static_assert(kOffB % 1024 == 0, "B operand alignment (swizzle atoms)");
// tensor-memory columns
constexpr uint32_t cX = 0, cXlo = 32, cDx = 64, cPre = 96, cDhid = 128, cDout = 160, cDoutLo = 176, cAT = 192, cATlo = 320,
                   cD = 448;

enum {
  Q_XFULL = 0, Q_XEMPTY, Q_ATXFULL, Q_ATXFREE, Q_DXFULL, Q_DXEMPTY,
  Q_DOUTT, Q_PDFULL, Q_DPREFULL, Q_ATDFULL0, Q_ATDFULL1, Q_GRADDONE0, Q_GRADDONE1, Q_DFREE,
  Q_DSFULL0, Q_DSFREE0 = Q_DSFULL0 + kDsStages, Q_BFULL0 = Q_DSFREE0 + kDsStages, Q_BFREE0 = Q_BFULL0 + 2,
  Q_WFULL0 = Q_BFREE0 + 2, Q_WEMPTY0 = Q_WFULL0 + kWStages, Q_COUNT = Q_WEMPTY0 + kWStages
};
static_assert(Q_COUNT * 8 + 8 <= 384, "barrier block");

__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red1(float* p, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void tmem_st4f(uint32_t taddr, const float (&v)[4]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3])
               : "memory");
}
// B operand (MN-major) of a tf32 tcgen05.mma: 128-byte rows, 4-row swizzle atoms, N groups of 32 columns LBO apart
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return make_desc(saddr, lbo, sbo) | ((uint64_t)1 << 61);
}
__host__ __device__ constexpr uint32_t make_idesc_bmn(int M, int N) { return make_idesc(M, N) | (1u << 16); }

__device__ __forceinline__ void normal_cdf_pdf2(float p, float& cdf, float& pdf) {   // as psf_gen_tc_bwd_sm100.cu
  const float ax = fabsf(p);
  const float t = rcp_approx(fmaf(ax, 0.27599917f, 1.f));
  float q = fmaf(-0.22690525f, t, 0.88164259f);
  q = fmaf(q, t, -0.62774200f);
  q = fmaf(q, t, 0.64431879f);
  q = fmaf(q, t, 0.093436931f);
  q = fmaf(q, t, 0.23524893f);
  const float e = ex2_approx(p * p * -0.72134752f);
  const float tail = 0.5f * (q * t) * e;
  cdf = p >= 0.f ? 1.f - tail : tail;
  pdf = 0.39894228f * e;
}

__global__ void __launch_bounds__(kThreads2, 1) psf_gen_bwd_tc2_kernel(const GenParams p, const uint8_t* __restrict__ img) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  TL_DECL;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + kOffBar2;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + kOffBar2 + Q_COUNT * 8);
  const int n_gen = p.n_gen;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };

  // the staging tiles start out zero: tap 15 of a 15-tap generator is never written
  for (int i = tid; i < (int)(kDsStages * kDsBuf / 4); i += kThreads2) reinterpret_cast<float*>(smem + kOffDs)[i] = 0.f;
  if (tid == 0) {
    mbar_init(bar(Q_XFULL), 4);
    mbar_init(bar(Q_XEMPTY), 1);
    mbar_init(bar(Q_ATXFULL), 1);
    mbar_init(bar(Q_ATXFREE), 1);
    mbar_init(bar(Q_DXFULL), 1);
    mbar_init(bar(Q_DXEMPTY), 4);
    mbar_init(bar(Q_DOUTT), 4);
    mbar_init(bar(Q_PDFULL), 1);
    mbar_init(bar(Q_DPREFULL), kEpi);
    mbar_init(bar(Q_ATDFULL0), 1);              // one per dout^T builder warp (they run out of step with each other)
    mbar_init(bar(Q_ATDFULL1), 1);
    mbar_init(bar(Q_GRADDONE0), 1);             // per step parity: the builders only wait on every other step, and a
    mbar_init(bar(Q_GRADDONE1), 1);             // parity wait cannot tell phase k from phase k + 2
    mbar_init(bar(Q_DFREE), 2);
    for (int i = 0; i < kDsStages; ++i) {
      mbar_init(bar(Q_DSFULL0 + i), 1);           // TMA thread (+ bytes); for the V generator the staging warp instead
      mbar_init(bar(Q_DSFREE0 + i), 4 + 1);       // the four helpers (dout -> tensor memory) and the step's dout^T builder
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar(Q_BFULL0 + i), kEpi);
      mbar_init(bar(Q_BFREE0 + i), 1);
    }
    for (int i = 0; i < kWStages; ++i) {
      mbar_init(bar(Q_WFULL0 + i), 1);
      mbar_init(bar(Q_WEMPTY0 + i), 1);
    }
    fence_barrier_init();
  }
  if (warp == kMmaW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  const long long ntiles = p.rows >> 7;                  // rows % 128 == 0 (checked by the launcher)
  const int my_tiles = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const uint32_t total = (uint32_t)my_tiles * (uint32_t)n_gen;
  float* wsc = p.ws + (long long)blockIdx.x * n_gen * kAccBlk;
  auto tile_row0 = [&](uint32_t it) { return (blockIdx.x + (long long)it * gridDim.x) * 128; };

  if (warp < kEpi) {
    // ================================================ epilogue ================================================
    const int quarter = warp & 3, cq = warp >> 2;          // rows 32 quarter .., hidden units 8 cq .., taps 4 cq ..
    const int rl = quarter * 32 + lane;
    const uint32_t lane_base = ((uint32_t)quarter * 32u) << 16;
    const uint32_t tb = tmem + lane_base;
    // B operand address of this thread's 32-byte chunk: row rl, logical chunk cq, XOR-swizzled with the row
    const uint32_t boff = (uint32_t)(rl >> 2) * 512 + (uint32_t)(rl & 3) * 128 + (uint32_t)((cq ^ (rl & 3)) * 32);
    const bool swap = (rl >> 2) & 1;                        // which 16-byte half goes first (halves the bank conflicts)

    for (uint32_t s = 0; s < total; ++s) {
      const uint32_t wst = s % kWStages, bbuf = s & 1;
      float pre[8], dh[8];
      mbar_wait_tc(bar(Q_WFULL0 + wst), (s / kWStages) & 1);      // b1 (TMA write) visible to this thread
      TL(1, s);
      mbar_wait_tc(bar(Q_PDFULL), s & 1);
      TL(2, s);
      tc_fence_after();
      tmem_ld8f(tb + cPre + 8 * cq, pre);
      tmem_ld8f(tb + cDhid + 8 * cq, dh);
      tmem_wait_ld_tc();
      const float* b1 = reinterpret_cast<const float*>(smem + kOffW + wst * kImg + kB1) + 8 * cq;
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        const float pv = tc_unbias(pre[h]) + b1[h];
        float cdf, pdf;
        normal_cdf_pdf2(pv, cdf, pdf);
        pre[h] = dh[h] * fmaf(pv, pdf, cdf);               // dpre
        dh[h] = pv * cdf;                                  // hid
      }
      float ph[8], hh[8];
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        ph[h] = tf32_hi(pre[h]);
        pre[h] -= ph[h];                                   // dpre lo
        hh[h] = tf32_hi(dh[h]);
        dh[h] -= hh[h];                                    // hid lo
      }
      // [dpre | hid] (hi, lo) -> the B operand of the gradient product
      TL(3, s);
      mbar_wait_tc(bar(Q_BFREE0 + bbuf), ((s >> 1) & 1) ^ 1);
      TL(4, s);
      {
        // two 128-bit stores per 32-byte chunk; odd row quads issue the upper half first, so that the eight rows of a warp
        // that share a swizzle phase spread over both 16-byte slots (4 wavefronts per store instead of 8)
        const uint32_t bb = sbase + kOffB + bbuf * kBBuf + boff;
        const uint32_t o1 = swap ? 16u : 0u, o2 = o1 ^ 16u;
        auto put = [&](uint32_t off, const float (&v)[8]) {
          const float a0 = swap ? v[4] : v[0], a1 = swap ? v[5] : v[1], a2 = swap ? v[6] : v[2], a3 = swap ? v[7] : v[3];
          const float c0 = swap ? v[0] : v[4], c1 = swap ? v[1] : v[5], c2 = swap ? v[2] : v[6], c3 = swap ? v[3] : v[7];
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(bb + off + o1), "f"(a0), "f"(a1), "f"(a2), "f"(a3) : "memory");
          asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(bb + off + o2), "f"(c0), "f"(c1), "f"(c2), "f"(c3) : "memory");
        };
        put(0, ph);             // dpre hi : N group 0
        put(16384, hh);         // hid  hi : N group 1
        put(32768, pre);        // dpre lo
        put(49152, dh);         // hid  lo
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      // dpre -> tensor memory (A operand of dx += dpre W1); it overlays pre / dhid, which this thread has read
      tmem_st8f(tb + cPre + 8 * cq, ph);
      tmem_st8f(tb + cDhid + 8 * cq, pre);
      tmem_wait_st_tc();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(bar(Q_DPREFULL));
        mbar_arrive(bar(Q_BFULL0 + bbuf));
      }
      TL(5, s);
    }
  } else if (warp < kMmaW) {
    // ================================================ helpers ================================================
    const int quarter = warp & 3;
    const int rl = quarter * 32 + lane;
    const uint32_t lane_base = ((uint32_t)quarter * 32u) << 16;
    const uint32_t tb = tmem + lane_base;

    auto prepare_x = [&](uint32_t it) {                    // x of tile it: tensor memory (lane = row) and shared memory
      const long long row = tile_row0(it) + rl;
      float4 xin[GE / 4];
      const float4* src = reinterpret_cast<const float4*>(p.x + row * GE);
#pragma unroll
      for (int i = 0; i < GE / 4; ++i) xin[i] = __ldg(src + i);
      mbar_wait_relaxed(bar(Q_XEMPTY), (it & 1) ^ 1);
      tc_fence_after();
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float vh[16], vl[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 q = xin[4 * hf + i];
          const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            vh[4 * i + j] = tf32_hi(e[j]);
            vl[4 * i + j] = e[j] - vh[4 * i + j];
          }
        }
        tmem_st16f(tb + cX + 16 * hf, vh);
        tmem_st16f(tb + cXlo + 16 * hf, vl);
      }
      tmem_wait_st_tc();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(Q_XFULL));
    };
    auto build_xT = [&](uint32_t it) {                     // warp 16: lane e of A_T <- column e of the x tile (hi, lo)
      const float* xt = p.x + tile_row0(it) * GE + lane;   // 128-byte coalesced row reads (the tile is L2 / L1 resident)
      mbar_wait_relaxed(bar(Q_ATXFREE), (it & 1) ^ 1);     // the previous tile's gradient products have read x^T
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float vh[32], vl[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) vl[r] = __ldg(xt + (32 * c + r) * GE);
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          vh[r] = tf32_hi(vl[r]);
          vl[r] -= vh[r];
        }
        tmem_st32f(tb + cAT + 32 * c, vh);
        tmem_st32f(tb + cATlo + 32 * c, vl);
      }
      tmem_wait_st_tc();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(Q_ATXFULL));
    };
    auto write_dx = [&](uint32_t it) {
      const long long row = tile_row0(it) + rl;
      float v[GE];
      mbar_wait_relaxed(bar(Q_DXFULL), it & 1);
      tc_fence_after();
      tmem_ld32f(tb + cDx, v);
      tmem_wait_ld_tc();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(Q_DXEMPTY));
      float4* dst = reinterpret_cast<float4*>(p.dx + row * GE);
#pragma unroll
      for (int i = 0; i < GE / 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    };

    // dout of step s: staging tile -> tensor memory (A operand of dhid = dout W2; lane = row, 16 taps, hi then lo).  Done by
    // the helpers, off the epilogue's critical path; the region is free once pre / dhid of step s - 1 have been produced.
    auto publish_dout = [&](uint32_t s) {
      const uint32_t b = s % kDsStages;
      if (s > 0) mbar_wait_tc(bar(Q_PDFULL), (s - 1) & 1);
      mbar_wait_tc(bar(Q_DSFULL0 + b), (s / kDsStages) & 1);
      tc_fence_after();
      const float* ds = reinterpret_cast<const float*>(smem + kOffDs + b * kDsBuf);
      float hi[GLP], lo[GLP];
#pragma unroll
      for (int l = 0; l < GLP; ++l) {
        const float v = ds[l * 128 + rl];
        hi[l] = tf32_hi(v);
        lo[l] = v - hi[l];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(Q_DSFREE0 + b));
      tmem_st16f(tb + cDout, hi);
      tmem_st16f(tb + cDoutLo, lo);
      tmem_wait_st_tc();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(Q_DOUTT));
    };
    // The V generator's dV0 is [rows, C] in natural row order: warp 19 stages it (transposed) into the dout staging ring.
    auto stage_g = [&](uint32_t s1) {
      const uint32_t b1 = s1 % kDsStages, it1 = s1 / (uint32_t)n_gen;
      float* d1 = reinterpret_cast<float*>(smem + kOffDs + b1 * kDsBuf);
      mbar_wait_relaxed(bar(Q_DSFREE0 + b1), ((s1 / kDsStages) & 1) ^ 1);
#pragma unroll 1
      for (int rr = 0; rr < 4; ++rr) {
        const int r = 32 * rr + lane;
        const float* src = p.V + natural_row(p, tile_row0(it1) + r) * p.C;
#pragma unroll
        for (int c = 0; c < GLP; ++c) d1[c * 128 + r] = c < p.C ? __ldcs(src + c) : 0.f;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(Q_DSFULL0 + b1));
    };
    const bool has_g = p.n_gen > p.n_W;
    if (my_tiles > 0) {
      prepare_x(0);
      if (quarter == 0) build_xT(0);
      if (quarter == 3 && has_g)
        for (uint32_t s1 = 0; s1 < (uint32_t)kDsStages && s1 < total; ++s1)
          if (s1 % (uint32_t)n_gen >= (uint32_t)p.n_W) stage_g(s1);       // buffers that have not been used yet
      publish_dout(0);
    }
    for (uint32_t s = 0; s < total; ++s) {
      const uint32_t m = s % (uint32_t)n_gen, it = s / (uint32_t)n_gen;
      float* acc_m = wsc + (long long)m * kAccBlk;
      // Warps 17 and 18 alternate as dout^T builders: even steps use the A_T lanes 32..48, odd steps the lanes 64..80 of the
      // same columns, so that building step s + 1 overlaps the gradient product of step s (the product of a step also
      // multiplies the other buffer's lanes; those D lanes are never read).
      const bool builder = (quarter == 1 && (s & 1) == 0) || (quarter == 2 && (s & 1) == 1);
      if (builder) {
        // ---- dout^T of step s into this warp's A_T lanes (+ the ones lane), db2 ----
        const uint32_t b = s % kDsStages;
        const float* ds = reinterpret_cast<const float*>(smem + kOffDs + b * kDsBuf);
        TL(10, s);
        mbar_wait_tc(bar(Q_DSFULL0 + b), (s / kDsStages) & 1);
        TL(11, s);
        const int l = lane & 15;
        float sum = 0.f;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          float vh[32], vl[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 q = *reinterpret_cast<const float4*>(ds + l * 128 + 32 * c + 4 * i);
            const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float v = lane < 16 ? e[j] : (lane == 16 ? 1.f : 0.f);
              sum += v;
              vh[4 * i + j] = tf32_hi(v);
              vl[4 * i + j] = v - vh[4 * i + j];
            }
          }
          tmem_st32f(tb + cAT + 32 * c, vh);
          tmem_st32f(tb + cATlo + 32 * c, vl);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(Q_DSFREE0 + b));
        if (lane < GLP) red1(acc_m + GH * GE + GH + GLP * GH + lane, sum);      // db2
        tmem_wait_st_tc();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(Q_ATDFULL0 + (s & 1)));
        TL(12, s);
      }
      if (s + 1 < total) publish_dout(s + 1);              // needs pre / dhid of step s produced (PDFULL)
      if (m == (uint32_t)n_gen - 1 && it + 1 < (uint32_t)my_tiles) prepare_x(it + 1);
      if (quarter == 0 || builder) {
        // ---- drain the gradient accumulator of step s: lanes 0..31 dW1^T (warp 16); the builder's lanes: dW2 and db1 ----
        TL(13, s);
        mbar_wait_tc(bar(Q_GRADDONE0 + (s & 1)), (s >> 1) & 1);
        TL(14, s);
        tc_fence_after();
        float v[32];
        tmem_ld32f(tb + cD + (quarter == 0 ? 0 : 32), v);
        tmem_wait_ld_tc();
        if (quarter == 0) {                                 // acc[e][h]: dW1 transposed (the reduction kernel undoes it)
          float* dst = acc_m + lane * GH;
#pragma unroll
          for (int i = 0; i < 8; ++i) red4(dst + 4 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        } else {
          if (lane < GLP) {                                 // dW2[l][h]
            float* dst = acc_m + GH * GE + GH + lane * GH;
#pragma unroll
            for (int i = 0; i < 8; ++i) red4(dst + 4 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          tmem_ld32f(tb + cD, v);                           // the ones lane holds db1 in the dpre columns
          tmem_wait_ld_tc();
          if (lane == 16) {
            float* dst = acc_m + GH * GE;
#pragma unroll
            for (int i = 0; i < 8; ++i) red4(dst + 4 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(Q_DFREE));
        TL(15, s);
      }
      if (quarter == 3 && has_g) {
        // in step with the pipeline (a parity wait cannot be more than one phase early): once the gradient product of step
        // s is done, the staging buffer of step s has been consumed and is free for step s + kDsStages.  (After
        // prepare_x: the first products of the next tile are issued before this step's gradient product.)
        mbar_wait_relaxed(bar(Q_GRADDONE0 + (s & 1)), (s >> 1) & 1);
        const uint32_t s1 = s + kDsStages;
        if (s1 < total && s1 % (uint32_t)n_gen >= (uint32_t)p.n_W) stage_g(s1);
      }
      if (m == (uint32_t)n_gen - 1) {
        write_dx(it);
        if (quarter == 0 && it + 1 < (uint32_t)my_tiles) build_xT(it + 1);
      }
    }
  } else if (warp == kMmaW) {
    // ================================================ MMA issuer ================================================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, 32), idesc_g = make_idesc_bmn(128, 64);
      auto issue_pd = [&](uint32_t s) {
        const uint32_t m = s % (uint32_t)n_gen, it = s / (uint32_t)n_gen, st = s % kWStages;
        if (m == 0) mbar_wait_tc(bar(Q_XFULL), it & 1);
        mbar_wait_tc(bar(Q_WFULL0 + st), (s / kWStages) & 1);
        mbar_wait_tc(bar(Q_DOUTT), s & 1);
        tc_fence_after();
        const uint32_t w = sbase + kOffW + st * kImg;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_ts(tmem + cPre, tmem + (pass == 0 ? cXlo : cX) + ks * 8,
                   make_desc(w + kW1 + (pass == 1 ? 4096 : 0) + ks * 256, 128, 1024), idesc, (pass | ks) != 0);
#pragma unroll
        for (int pass = 0; pass < 3; ++pass)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            mma_ts(tmem + cDhid, tmem + (pass == 0 ? cDoutLo : cDout) + ks * 8,
                   make_desc(w + kW2T + (pass == 1 ? 2048 : 0) + ks * 256, 128, 512), idesc, (pass | ks) != 0);
        mma_commit(bar(Q_PDFULL));
        if (m == (uint32_t)n_gen - 1) mma_commit(bar(Q_XEMPTY));
      };
      if (total > 0) issue_pd(0);
      for (uint32_t s = 0; s < total; ++s) {
        const uint32_t m = s % (uint32_t)n_gen, it = s / (uint32_t)n_gen, st = s % kWStages, b = s & 1;
        const bool last = m == (uint32_t)n_gen - 1;
        // ---- dx += dpre W1 (A = dpre, which the epilogue wrote over pre / dhid) ----
        TL(20, s);
        mbar_wait_tc(bar(Q_DPREFULL), s & 1);
        TL(21, s);
        if (m == 0) mbar_wait_tc(bar(Q_DXEMPTY), (it & 1) ^ 1);
        tc_fence_after();
        {
          const uint32_t w = sbase + kOffW + st * kImg;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_ts(tmem + cDx, tmem + (pass == 0 ? cDhid : cPre) + ks * 8,
                     make_desc(w + kW1T + (pass == 1 ? 4096 : 0) + ks * 256, 128, 1024), idesc, (m | pass | ks) != 0);
        }
        mma_commit(bar(Q_WEMPTY0 + st));
        if (last) mma_commit(bar(Q_DXFULL));
        TL(22, s);
        if (s + 1 < total) issue_pd(s + 1);
        TL(23, s);
        // ---- D = [x^T ; dout^T ; 1] . [dpre | hid]  (K = the tile's 128 rows) ----
        mbar_wait_tc(bar(Q_BFULL0 + b), (s >> 1) & 1);
        mbar_wait_tc(bar(Q_ATDFULL0 + (s & 1)), (s >> 1) & 1);
        if (m == 0) mbar_wait_tc(bar(Q_ATXFULL), it & 1);
        mbar_wait_tc(bar(Q_DFREE), (s & 1) ^ 1);
        TL(24, s);
        tc_fence_after();
        {
          const uint32_t bb = sbase + kOffB + b * kBBuf;
#pragma unroll 1
          for (int pass = 0; pass < 3; ++pass) {     // A_lo B_hi, A_hi B_lo, A_hi B_hi
            const uint32_t a0 = tmem + (pass == 0 ? cATlo : cAT);
            const uint32_t b0 = bb + (pass == 1 ? 32768 : 0);
#pragma unroll 4
            for (int ks = 0; ks < 16; ++ks)
              mma_ts(tmem + cD, a0 + ks * 8, make_desc_mn(b0 + ks * 1024, 16384, 512), idesc_g, (pass | ks) != 0);
          }
        }
        mma_commit(bar(Q_GRADDONE0 + (s & 1)));
        TL(25, s);
        mma_commit(bar(Q_BFREE0 + b));
        if (last) mma_commit(bar(Q_ATXFREE));
      }
    }
    __syncwarp();
  } else if (warp == kTmaW) {
    // ================================================ TMA: weight images ================================================
    if (lane == 0) {
      for (uint32_t s = 0; s < total; ++s) {
        const uint32_t st = s % kWStages, m = s % (uint32_t)n_gen;
        mbar_wait_relaxed(bar(Q_WEMPTY0 + st), ((s / kWStages) & 1) ^ 1);
        mbar_expect_tx(bar(Q_WFULL0 + st), kImg);
        tma_load_1d(sbase + kOffW + st * kImg, img + (size_t)m * kImg, kImg, bar(Q_WFULL0 + st));
      }
    }
    __syncwarp();
  } else {
    // ================================================ TMA: dout staging ================================================
    // (its own issuer: dW streams from HBM, so it has to run as far ahead as the staging ring allows, not in step with the
    // weight ring)
    if (lane == 0) {
      for (uint32_t s = 0; s < total; ++s) {
        const uint32_t m = s % (uint32_t)n_gen, it = s / (uint32_t)n_gen, b = s % kDsStages;
        if (m >= (uint32_t)p.n_W) continue;                  // the V generator's tile is staged and signalled by warp 19
        mbar_wait_relaxed(bar(Q_DSFREE0 + b), ((s / kDsStages) & 1) ^ 1);
        mbar_expect_tx(bar(Q_DSFULL0 + b), (uint32_t)p.L * 512u);   // dW of generator m, this tile: one 512-byte run per tap
        const float* src = p.Wt + (long long)m * p.sWf + tile_row0(it);
        for (int l = 0; l < p.L; ++l)
          tma_load_1d(sbase + kOffDs + b * kDsBuf + (uint32_t)l * 512u, src + (long long)l * p.sWt, 512u, bar(Q_DSFULL0 + b));
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kMmaW) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

struct DevAttr2 {
  bool done = false;
  int status = 0;
};
DevAttr2 g_tc2_attr[PSF_MAX_DEVICES];

}  // namespace

#ifdef PSF_TC_DEBUG
extern "C" int psf_debug_timeline(long long* out) {     // out: 24 warps x kTlPerWarp x (code, clock); code 0 = unused
  cudaMemcpyFromSymbol(out, g_timeline, sizeof(long long) * 24 * kTlPerWarp * 2);
  void* sym = nullptr;
  cudaGetSymbolAddress(&sym, g_timeline);
  cudaMemset(sym, 0, sizeof(long long) * 24 * kTlPerWarp * 2);
  return kTlPerWarp;
}
#endif

// Shapes the kernel covers: whole 128-row tiles, tap-major dW with 16-byte aligned 512-byte runs per tap.
int psf_gen_tc_bwd2(const GenParams& p, const void* images, int* grid_out, cudaStream_t st) {
  if (p.n_gen < 1 || p.L > GLP || p.C > GLP || p.pre != nullptr) return PSF_ERR_UNSUPPORTED;
  if ((p.rows & 127) || (p.sWt & 3) || (p.sWf & 3) || (reinterpret_cast<uintptr_t>(p.Wt) & 15)) return PSF_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(images) & 127) return PSF_ERR_UNSUPPORTED;
  const int dev = psf_current_device();
  if (dev < 0) return PSF_ERR_UNSUPPORTED;
  {
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (!g_tc2_attr[dev].done) {
      cudaError_t e = cudaFuncSetAttribute(psf_gen_bwd_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem2);
      g_tc2_attr[dev].status = (int)e;
      g_tc2_attr[dev].done = true;
      if (e != cudaSuccess) cudaGetLastError();
    }
  }
  if (g_tc2_attr[dev].status != 0) return PSF_ERR_UNSUPPORTED;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long ntiles = p.rows >> 7;
  const long long grid = ntiles < sms ? ntiles : sms;
  psf_gen_bwd_tc2_kernel<<<(unsigned)grid, kThreads2, kSmem2, st>>>(p, (const uint8_t*)images);
  *grid_out = (int)grid;
  return (int)cudaGetLastError();
}
