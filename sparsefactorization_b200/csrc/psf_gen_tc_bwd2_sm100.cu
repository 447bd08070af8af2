// Backward pass of the generators, every GEMM-shaped piece on tcgen05 (second version; the first one,
// psf_gen_tc_bwd_sm100.cu, keeps the parameter-gradient products on the SIMT / mma.sync pipes and serves ragged shapes and
// 16-output generators).
//
// Per 128-row tile and generator m = Linear(32,32) -> GELU -> Linear(32,L) (SyntheticExperiments/psf.py:110-126):
//
//   row-local products, A operand in TENSOR MEMORY with lane = row (B = weight images in shared memory, streamed by TMA),
//   error-compensated 3xTF32, small terms first (psf_gen_tc_sm100.cu):
//     pre  [128 x 32] = x    [128 x 32] . W1^T
//     dhid [128 x 32] = dout [128 x 16] . W2
//     dx   [128 x 32] += dpre[128 x 32] . W1                     (one accumulator over the tile's generators)
//   parameter gradients, contraction over the tile's ROWS - so the rows are the K dimension, which lives in the COLUMNS of
//   a tensor-memory A operand and in the rows of an MN-major shared-memory B operand:
//     D[lane][n] = sum_r A_T[lane][r] . (B_hi + B_lo)[r][n],      B cols = [ dpre_m (32: h) | hid_m (32: h) ]
//   The hi and lo halves of the A-side split sit in DIFFERENT LANES of the same columns, so that one MMA multiplies both
//   (two passes - B_lo, B_hi - instead of three, and half the tensor-memory columns); the drain adds the lane pairs:
//     lanes   0..15 / 16..31   x^T features  0..15, hi / lo        -> D cols  0..31 = dW1^T rows e
//     lanes  64..79 / 80..95   x^T features 16..31, hi / lo
//     lanes  32..47 / 48..63   dout_m^T taps of EVEN steps, hi / lo -> D cols 32..63 = dW2 rows l; tap lane 15 is a row
//     lanes  96..111/112..127  dout_m^T taps of ODD steps,  hi / lo    of ones: D cols 0..31 = db1  (so L, C <= 15)
//   All 128 lanes are used; still only a quarter of the products are wanted (x^T with hid and dout^T with dpre are not),
//   which is what a 32-wide problem costs on a 128-lane MMA.  Batching four generators into the lanes instead (dpre^T as A)
//   needs the producers' rows transposed and four generators resident at once, which fits neither memory.
//   The legacy warp-level mma.sync path this replaces costs 16 cycles per m16n8k8 TF32 here (three per product).
//
// Tensor pipe per step: 30 + 32 MMAs, about 1.5 k cycles.  pre / dhid and D are double buffered, the products of step s + 2
// are issued before the epilogue of step s + 1 starts, so that the sixteen epilogue warps never wait for the tensor pipe.
//
// dout^T is rewritten per generator from a shared-memory staging tile that TMA fills from the tap-major dW (512 contiguous
// bytes per tap; rows padded to 528 bytes against bank conflicts; dV0 of the V generator as it lies, row-major), which
// also feeds the epilogue's dout reads.
// B is written by the epilogue warps in the canonical MN-major tf32 layout (128-byte rows of 32 columns, 32-byte chunks
// XOR-swizzled with the row index, layout type "128B swizzle, 32B base"; probed in tools/microbench/umma_probe.cu).
//
// Warp roles (768 threads, one CTA per SM, persistent over tiles):
//   warps 0-15   epilogue: thread = (row, 8 hidden units); tcgen05.ld pre / dhid, Phi / phi, dpre -> tensor memory,
//                [dpre | hid] (hi, lo) -> shared memory B operand, dout of step s + 2 -> tensor memory A operand of
//                dhid
//   warps 16-19  helpers, one per 32-lane quarter: x -> tensor memory, dx writer, D drain (red.global into the CTA's slice);
//                warps 16 / 18 build x^T and sum dout over the rows (db2), warps 17 / 19 the dout^T strips of even / odd steps
//   warp 20      tensor-memory allocation + the thread issuing the row-local MMAs;  warp 23  the thread issuing the
//                gradient MMAs
//   warp 21      TMA: weight images (3 stages);  warp 22  TMA: dout staging tiles (4 stages)
#include "psf_gen.cuh"
#include "psf_tc.cuh"

namespace {

// PSF_TC_DEBUG builds record a timeline of block 0 (event, step, SM clock) that tools/gen_timeline.py prints
#ifdef PSF_TC_DEBUG
constexpr int kTlPerWarp = 4096;
__device__ uint2 g_timeline[24 * kTlPerWarp];
#define TL_DECL                                                                                                 \
  uint2* tl_p = g_timeline + (size_t)(threadIdx.x >> 5) * kTlPerWarp;                                           \
  uint2* const tl_end = blockIdx.x == 0 ? tl_p + kTlPerWarp : tl_p
#define TL(ev, step)                                                                                            \
  do {                                                                                                          \
    if (tl_p < tl_end) *tl_p++ = make_uint2((uint32_t)(ev) * 100000u + (uint32_t)(step) + 1u, (uint32_t)clock()); \
  } while (0)
#else
#define TL_DECL
#define TL(ev, step)
#endif

constexpr int kEpi = 16, kHelp0 = 16, kMmaW = 20, kTmaW = 21, kTmaD = 22, kMmaG = 23;
constexpr int kThreads2 = (kMmaG + 1) * 32;
// weight image of one generator (same as psf_gen_tc_bwd_sm100.cu)
constexpr uint32_t kW1 = 0, kW1T = 8192, kW2T = 16384, kB1 = 20480, kImg = 20736;
constexpr int kWStages = 3;
constexpr int kDsStages = 4;
constexpr int kDsRow = 132;                                  // floats per tap row of a staging tile (128 + 4 of padding)
// shared memory map (bytes)
constexpr uint32_t kOffB = 0;                                // B operand: 2 buffers x (hi: dpre 16K, hid 16K | lo: 16K, 16K)
constexpr uint32_t kBBuf = 65536;
constexpr uint32_t kOffW = kOffB + 2 * kBBuf;                // weight images, kWStages stages
constexpr uint32_t kOffDs = kOffW + kWStages * kImg;         // dout staging: kDsStages x [16 taps][kDsRow] floats
constexpr uint32_t kDsBuf = 16 * kDsRow * 4;
constexpr uint32_t kOffBar2 = kOffDs + kDsStages * kDsBuf;
constexpr uint32_t kSmem2 = kOffBar2 + 384;
static_assert(kSmem2 <= 227 * 1024, "shared memory");
static_assert(kOffB % 1024 == 0, "B operand alignment (swizzle atoms)");
static_assert(kOffDs % 16 == 0 && (kDsRow * 4) % 16 == 0, "TMA destination alignment");
// tensor-memory columns: x (hi, lo) | dx | two sets of (pre, dhid) | dout (hi, lo) | A_T | two gradient accumulators
constexpr uint32_t cX = 0, cXlo = 32, cDx = 64, cPD = 96, cDout = 224, cDoutLo = 240, cAT = 256, cD = 384;

enum {
  Q_XFULL = 0, Q_XEMPTY, Q_ATXFULL, Q_DXFULL, Q_DXEMPTY, Q_DOUTT, Q_DPREFULL0, Q_DPREFULL1,
  Q_PDFULL0, Q_PDFULL1, Q_ATDFULL0, Q_ATDFULL1, Q_GRADDONE0, Q_GRADDONE1, Q_DFREE0, Q_DFREE1,
  Q_DSFULL0, Q_DSFREE0 = Q_DSFULL0 + kDsStages, Q_WFULL0 = Q_DSFREE0 + kDsStages, Q_WEMPTY0 = Q_WFULL0 + kWStages,
  Q_COUNT = Q_WEMPTY0 + kWStages
};
static_assert(Q_COUNT * 8 + 8 <= 384, "barrier block");

__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
#if defined(PSF_TC_ABLATE) && PSF_TC_ABLATE == 1
  return;
#endif
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

// the same on two values with packed arithmetic (same operations in the same order per value)
__device__ __forceinline__ void normal_cdf_pdf2x2(uint64_t p2, uint64_t& cdf2, uint64_t& pdf2) {
  float pa, pb;
  unpack2f(p2, pa, pb);
  const uint64_t ax2 = pack2f(fabsf(pa), fabsf(pb));
  float ua, ub;
  unpack2f(fma2f(ax2, dup2f(0.27599917f), dup2f(1.f)), ua, ub);
  const uint64_t t2 = pack2f(rcp_approx(ua), rcp_approx(ub));
  uint64_t q2 = fma2f(dup2f(-0.22690525f), t2, dup2f(0.88164259f));
  q2 = fma2f(q2, t2, dup2f(-0.62774200f));
  q2 = fma2f(q2, t2, dup2f(0.64431879f));
  q2 = fma2f(q2, t2, dup2f(0.093436931f));
  q2 = fma2f(q2, t2, dup2f(0.23524893f));
  float ea, eb;
  unpack2f(mul2f(mul2f(p2, p2), dup2f(-0.72134752f)), ea, eb);
  const uint64_t e2 = pack2f(ex2_approx(ea), ex2_approx(eb));
  const uint64_t tail2 = mul2f(mul2f(dup2f(0.5f), mul2f(q2, t2)), e2);
  float ta, tb, oa, ob;
  unpack2f(tail2, ta, tb);
  unpack2f(fma2f(tail2, dup2f(-1.f), dup2f(1.f)), oa, ob);
  cdf2 = pack2f(pa >= 0.f ? oa : ta, pb >= 0.f ? ob : tb);
  pdf2 = mul2f(dup2f(0.39894228f), e2);
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

  // taps >= L of a staging tile are never written by the copies (L, C <= 15)
  for (int i = tid; i < (int)(kDsStages * kDsBuf / 4); i += kThreads2)
    reinterpret_cast<float*>(smem + kOffDs)[i] = (i % (16 * kDsRow)) >= 15 * kDsRow ? 1.f : 0.f;     // tap row 15: ones (db1)
  if (tid == 0) {
    mbar_init(bar(Q_XFULL), 4);
    mbar_init(bar(Q_XEMPTY), 1);
    mbar_init(bar(Q_ATXFULL), 2);
    mbar_init(bar(Q_DXFULL), 1);
    mbar_init(bar(Q_DXEMPTY), 4);
    mbar_init(bar(Q_DOUTT), kEpi);
    for (int i = 0; i < 2; ++i) {
      // one barrier per step parity: some waiters only look at every other step, and a parity wait cannot tell phase k
      // from phase k + 2
      mbar_init(bar(Q_PDFULL0 + i), 1);
      mbar_init(bar(Q_DPREFULL0 + i), kEpi);
      mbar_init(bar(Q_ATDFULL0 + i), 1);
      mbar_init(bar(Q_GRADDONE0 + i), 1);
      mbar_init(bar(Q_DFREE0 + i), 3);            // the two x^T warps and the step's strip warp
    }
    for (int i = 0; i < kDsStages; ++i) {
      mbar_init(bar(Q_DSFULL0 + i), 1);           // TMA thread (+ bytes)
      mbar_init(bar(Q_DSFREE0 + i), kEpi + 2);    // the epilogue warps (dout -> tensor memory), the step's strip builder and
                                                  // the x^T warp that sums its columns (db2)
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

    // dout of step t, this thread's row and four taps: staging tile -> registers -> tensor memory (A operand of
    // dhid = dout W2; lane = row, 16 taps, hi then lo).  The caller has made sure that the products of the previous step
    // have read the region.  W generators are staged tap-major [tap][row], the V generator row-major [row][C].
    auto publish_dout = [&](uint32_t t, bool is_v) {
      const uint32_t b = t % kDsStages;
#if !defined(PSF_TC_ABLATE) || PSF_TC_ABLATE != 2
      mbar_wait_tc(bar(Q_DSFULL0 + b), (t / kDsStages) & 1);
#endif
      const float* ds = reinterpret_cast<const float*>(smem + kOffDs + b * kDsBuf);
      float hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int l = 4 * cq + j;
        const float v = is_v ? (l < p.C ? ds[rl * p.C + l] : 0.f) : (l < p.L ? ds[l * kDsRow + rl] : 0.f);
        hi[j] = tf32_hi(v);
        lo[j] = v - hi[j];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(Q_DSFREE0 + b));
      tmem_st4f(tb + cDout + 4 * cq, hi);
      tmem_st4f(tb + cDoutLo + 4 * cq, lo);
    };
    auto signal_dout = [&]() {
      tmem_wait_st_tc();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(Q_DOUTT));
    };
    uint32_t m2 = 0;                                        // generator of step s + 2
    auto advance2 = [&]() {
      if (++m2 == (uint32_t)n_gen) m2 = 0;
    };
    if (total > 0) {
      publish_dout(0, m2 >= (uint32_t)p.n_W);
      advance2();
      signal_dout();
    }
    if (total > 1) {
      mbar_wait_tc(bar(Q_PDFULL0), 0);                     // the products of step 0 have read dout(0)
      tc_fence_after();
      publish_dout(1, m2 >= (uint32_t)p.n_W);
      advance2();
      signal_dout();
    }

    for (uint32_t s = 0; s < total; ++s) {
      const uint32_t wst = s % kWStages, par = s & 1;
      const uint32_t cPre = cPD + 64 * par, cDhid = cPre + 32;
      float pre[8], dh[8];
      const bool more = s + 2 < total;
      TL(1, s);
      mbar_wait_tc(bar(Q_PDFULL0 + par), (s >> 1) & 1);
      mbar_wait_tc(bar(Q_WFULL0 + wst), (s / kWStages) & 1);      // b1 (TMA write) visible to this thread; complete by now
      TL(2, s);
      tc_fence_after();
      tmem_ld8f(tb + cPre + 8 * cq, pre);
      tmem_ld8f(tb + cDhid + 8 * cq, dh);
      tmem_wait_ld_tc();
      const float* b1 = reinterpret_cast<const float*>(smem + kOffW + wst * kImg + kB1) + 8 * cq;
      float ph[8], hh[8];
#pragma unroll
      for (int h = 0; h < 8; h += 2) {                     // two hidden units at a time, packed arithmetic
        const uint64_t pv2 = fma2f(pack2f(tc_unbias(pre[h]), tc_unbias(pre[h + 1])), dup2f(1.f), pack2f(b1[h], b1[h + 1]));
        uint64_t cdf2, pdf2, hi, lo;
        normal_cdf_pdf2x2(pv2, cdf2, pdf2);
        const uint64_t dpre2 = mul2f(pack2f(dh[h], dh[h + 1]), fma2f(pv2, pdf2, cdf2));
        const uint64_t hid2 = mul2f(pv2, cdf2);
        tf32_split2(dpre2, hi, lo);
        unpack2f(hi, ph[h], ph[h + 1]);
        unpack2f(lo, pre[h], pre[h + 1]);                  // dpre lo
        tf32_split2(hid2, hi, lo);
        unpack2f(hi, hh[h], hh[h + 1]);
        unpack2f(lo, dh[h], dh[h + 1]);                    // hid lo
      }
      // [dpre | hid] (hi, lo) -> the B operand of the gradient product; the buffer is free once the product of step s - 2
      // is done
      TL(3, s);
      mbar_wait_tc(bar(Q_GRADDONE0 + par), ((s >> 1) & 1) ^ 1);
      TL(4, s);
      {
        // two 128-bit stores per 32-byte chunk; odd row quads issue the upper half first, so that the eight rows of a warp
        // that share a swizzle phase spread over both 16-byte slots (4 wavefronts per store instead of 8)
        const uint32_t bb = sbase + kOffB + par * kBBuf + boff;
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
        TL(6, s);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        TL(7, s);
      }
      // dpre -> tensor memory (A operand of dx += dpre W1); it overlays pre / dhid, which this thread has read
      tmem_st8f(tb + cPre + 8 * cq, ph);
      tmem_st8f(tb + cDhid + 8 * cq, pre);
      // dout of step s + 2, once the products of step s + 1 (issued a step ago) have read the region
      if (more) {
        mbar_wait_tc(bar(Q_PDFULL0 + (par ^ 1)), ((s + 1) >> 1) & 1);
        TL(40, s);
        tc_fence_after();
        publish_dout(s + 2, m2 >= (uint32_t)p.n_W);
        advance2();
      }
      TL(8, s);
      tmem_wait_st_tc();
      TL(9, s);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(Q_DPREFULL0 + par));    // covers dpre, B and dout(s + 2)
      TL(5, s);
    }
  } else if (warp < kMmaW) {
    // ================================================ helpers ================================================
    const int quarter = warp & 3;
    const int rl = quarter * 32 + lane;
    const uint32_t lane_base = ((uint32_t)quarter * 32u) << 16;
    const uint32_t tb = tmem + lane_base;
    const bool xside = (quarter & 1) == 0;                  // warps 16, 18: x^T features 0..15 / 16..31
    const uint32_t par_mine = (uint32_t)(quarter >> 1);     // warps 17, 19: strips of even / odd steps
    const int fl = lane & 15;                               // feature / tap of this lane; lanes >= 16 carry the lo halves
    const bool is_lo = lane >= 16;

    auto prepare_x = [&](uint32_t it) {                    // x of tile it: tensor memory (lane = row)
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
    // warps 16 / 18: A_T lanes <- columns of the x tile (16 features each, hi in the lower half warp, lo in the upper).
    // The caller has seen the previous tile's last gradient product finish.
    auto build_xT = [&](uint32_t it) {
      const float* xt = p.x + tile_row0(it) * GE + (quarter >> 1) * 16 + fl;   // the tile is L2 / L1 resident
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float o[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) o[r] = __ldg(xt + (32 * c + r) * GE);
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          const float h = tf32_hi(o[r]);
          o[r] = is_lo ? o[r] - h : h;
        }
        tmem_st32f(tb + cAT + 32 * c, o);
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
    // warps 17 / 19: dout^T of step t (same parity as the warp) into this warp's A_T lanes - tap fl, hi / lo by half warp,
    // tap lane 15 a row of ones - and db2.  The caller has seen the gradient product of step t - 2 finish.
    auto build_strip = [&](uint32_t t, uint32_t m) {       // m = generator of step t
      const uint32_t b = t % kDsStages;
      const float* ds = reinterpret_cast<const float*>(smem + kOffDs + b * kDsBuf);
      TL(10, t);
      mbar_wait_tc(bar(Q_DSFULL0 + b), (t / kDsStages) & 1);
      TL(11, t);
      if (m < (uint32_t)p.n_W) {
        // staged tap-major [tap][row]; taps >= L hold stale finite values (their D rows are never read) and tap row 15 of
        // every staging tile holds ones (written at kernel start, out of reach of the copies)
        const float* dr = ds + fl * kDsRow;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          float o[32];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 q = *reinterpret_cast<const float4*>(dr + 32 * c + 4 * i);
            const float e[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float h = tf32_hi(e[j]);
              o[4 * i + j] = is_lo ? e[j] - h : h;
            }
          }
          tmem_st32f(tb + cAT + 32 * c, o);
        }
      } else {
        // the V generator: staged row-major [row][C] (the layout of dV0)
        // taps >= C read a zero word (the padding of tap row 14, which no copy reaches), tap 15 the ones row
        const float* dv = fl < p.C ? ds + fl : ds + (fl == 15 ? 15 * kDsRow : 14 * kDsRow + 128);
        const int stride = fl < p.C ? p.C : 0;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          float o[32];
#pragma unroll
          for (int k = 0; k < 32; ++k) {
            const float v = dv[(32 * c + k) * stride];
            const float h = tf32_hi(v);
            o[k] = is_lo ? v - h : h;
          }
          tmem_st32f(tb + cAT + 32 * c, o);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(Q_DSFREE0 + b));
      tmem_wait_st_tc();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(Q_ATDFULL0 + par_mine));
      TL(12, t);
    };
    // warps 16 / 18 (which have slack; the strip warps are the critical ones): db2 = column sums of dout of the even / odd
    // steps, straight from the staging tile (lane = tap; the upper half warp idles)
    auto sum_dout = [&](uint32_t t, uint32_t m) {
      const uint32_t b = t % kDsStages;
      const int Lm = m < (uint32_t)p.n_W ? p.L : p.C;
      const float* ds = reinterpret_cast<const float*>(smem + kOffDs + b * kDsBuf);
      mbar_wait_tc(bar(Q_DSFULL0 + b), (t / kDsStages) & 1);
      float sum = 0.f;
      if (lane < Lm) {
        if (m < (uint32_t)p.n_W) {
          const float4* dr = reinterpret_cast<const float4*>(ds + lane * kDsRow);
#pragma unroll 8
          for (int i = 0; i < 32; ++i) {
            const float4 q = dr[i];
            sum += (q.x + q.y) + (q.z + q.w);
          }
        } else {
#pragma unroll 8
          for (int r = 0; r < 128; ++r) sum += ds[r * p.C + lane];
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(Q_DSFREE0 + b));
      if (lane < Lm) red1(wsc + (long long)m * kAccBlk + GH * GE + GH + GLP * GH + lane, sum);
    };
    const uint32_t par_x = (uint32_t)(quarter >> 1);        // warp 16: even steps, warp 18: odd steps
    if (my_tiles > 0) {
      prepare_x(0);
      if (xside) build_xT(0);
      else if (par_mine < total) build_strip(par_mine, par_mine % (uint32_t)n_gen);
      for (uint32_t t = 1; t <= 2; ++t)                   // tiles whose first products are issued ahead of the loop
        if (t < total && t % (uint32_t)n_gen == 0) prepare_x(t / (uint32_t)n_gen);
      if (xside && par_x < total) sum_dout(par_x, par_x % (uint32_t)n_gen);
    }
    uint32_t m = 0, it = 0;                                // generator and tile of step s (no divisions in the loop:
    uint32_t m2h = 2 % (uint32_t)n_gen;                    // every instruction of a helper costs ~10 cycles), and the
    for (uint32_t s = 0; s < total; ++s) {                 // generator of step s + 2
      const uint32_t par = s & 1;
      const bool last = m == (uint32_t)n_gen - 1;
      float* acc_m = wsc + (long long)m * kAccBlk;
      if (last) write_dx(it);
      const bool strip = !xside && par_mine == par;
      if (xside || strip) {
        // ---- drain the gradient accumulator of step s: hi + lo lanes; x^T lanes: dW1^T, strip lanes: dW2 and db1 ----
        TL(13, s);
        mbar_wait_tc(bar(Q_GRADDONE0 + par), (s >> 1) & 1);
        TL(14, s);
        tc_fence_after();
        const uint32_t d = tb + cD + 64 * par;
        float v[32];
        // (red.global into the CTA's own slice: a plain read-modify-write of the lane's own slots was measured slower - its
        // loads and stores share the LSU with the epilogue's shared-memory traffic)
        if (xside) {                                        // acc[e][h]: dW1 transposed (the reduction kernel undoes it)
          tmem_ld32f(d, v);
          tmem_wait_ld_tc();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(Q_DFREE0 + par));
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += __shfl_down_sync(0xffffffffu, v[i], 16);
          if (lane < 16) {
            float* dst = acc_m + ((quarter >> 1) * 16 + lane) * GH;
#pragma unroll
            for (int i = 0; i < 8; ++i) red4(dst + 4 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        } else {
          const int Lm = m < (uint32_t)p.n_W ? p.L : p.C;
          tmem_ld32f(d + 32, v);                            // dW2[l][h] in the hid columns
          tmem_wait_ld_tc();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += __shfl_down_sync(0xffffffffu, v[i], 16);
          if (lane < Lm) {
            float* dst = acc_m + GH * GE + GH + lane * GH;
#pragma unroll
            for (int i = 0; i < 8; ++i) red4(dst + 4 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          tmem_ld32f(d, v);                                 // the ones lane (15; its lo partner is a zero row): db1 in the dpre columns
          tmem_wait_ld_tc();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar(Q_DFREE0 + par));
          if (lane == 15) {
            float* dst = acc_m + GH * GE;
#pragma unroll
            for (int i = 0; i < 8; ++i) red4(dst + 4 * i, v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        }
        TL(15, s);
        if (strip && s + 2 < total) build_strip(s + 2, m2h);
        if (xside && last && it + 1 < (uint32_t)my_tiles) build_xT(it + 1);
        if (xside && par_x == par && s + 2 < total) sum_dout(s + 2, m2h);
      }
      // x of the next tile, three steps before its first products are issued (n_gen >= 3; shorter chains: see the prologue)
      if (s + 3 < total && (n_gen >= 3 ? m + 3 == (uint32_t)n_gen : (s + 3) % (uint32_t)n_gen == 0))
        prepare_x(n_gen >= 3 ? it + 1 : (s + 3) / (uint32_t)n_gen);
      if (++m == (uint32_t)n_gen) m = 0, ++it;
      if (++m2h == (uint32_t)n_gen) m2h = 0;
    }
  } else if (warp == kMmaW) {
    // ================================================ MMA issuer ================================================
    // The issuing thread shares its scheduler with four busy epilogue warps, so every instruction it spends between two
    // MMAs is paid in issue slots: step counters instead of divisions, descriptor words that differ by constants.
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, 32);
      constexpr uint32_t hiW1 = (1024u >> 4) | (1u << 14), hiW2 = (512u >> 4) | (1u << 14);      // SBO | version
      constexpr uint32_t lboW = (128u >> 4) << 16;
      const uint32_t w0 = ((sbase + kOffW) >> 4) | lboW;
      struct Ctr {
        uint32_t m, it, st, wph;
      };
      auto advance = [&](Ctr& c) {
        if (++c.m == (uint32_t)n_gen) c.m = 0, ++c.it;
        if (++c.st == (uint32_t)kWStages) c.st = 0, c.wph ^= 1;
      };
      // pre = x W1^T and dhid = dout W2 of step t into the (pre, dhid) set t & 1
      auto issue_pd = [&](uint32_t t, const Ctr& c) {
        const uint32_t cPre = tmem + cPD + 64 * (t & 1), cDhid = cPre + 32;
        if (c.m == 0) mbar_wait_hot(bar(Q_XFULL), c.it & 1);
        TL(30, t);
        mbar_wait_hot(bar(Q_WFULL0 + c.st), c.wph);
        TL(31, t);
        tc_fence_after();
        const uint32_t w = w0 + c.st * (kImg >> 4);
#pragma unroll
        for (int pass = 0; pass < 3; ++pass)
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_ts2(cPre, tmem + (pass == 0 ? cXlo : cX) + ks * 8, w + ((kW1 + (pass == 1 ? 4096 : 0) + ks * 256) >> 4), hiW1,
                    idesc, (pass | ks) != 0);
#pragma unroll
        for (int pass = 0; pass < 3; ++pass)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            mma_ts2(cDhid, tmem + (pass == 0 ? cDoutLo : cDout) + ks * 8, w + ((kW2T + (pass == 1 ? 2048 : 0) + ks * 256) >> 4),
                    hiW2, idesc, (pass | ks) != 0);
        mma_commit(bar(Q_PDFULL0 + (t & 1)));
        if (c.m == (uint32_t)n_gen - 1) mma_commit(bar(Q_XEMPTY));
      };
      Ctr cs = {0, 0, 0, 0}, ct = {0, 0, 0, 0};              // step s (dx) and step t = s + 2 (row-local products)
      if (total > 0) mbar_wait_hot(bar(Q_DOUTT), 0), issue_pd(0, ct), advance(ct);
      if (total > 1) mbar_wait_hot(bar(Q_DOUTT), 1), issue_pd(1, ct), advance(ct);
      for (uint32_t s = 0; s < total; ++s) {
        const uint32_t par = s & 1;
        const uint32_t cPre = tmem + cPD + 64 * par, cDhid = cPre + 32;
        const bool last = cs.m == (uint32_t)n_gen - 1;
        // ---- dx += dpre W1 (A = dpre, which the epilogue wrote over pre / dhid; the arrival also covers dout(s + 2)) ----
        TL(20, s);
        mbar_wait_hot(bar(Q_DPREFULL0 + par), (s >> 1) & 1);
        TL(21, s);
        if (cs.m == 0) mbar_wait_hot(bar(Q_DXEMPTY), (cs.it & 1) ^ 1);
        tc_fence_after();
        {
          const uint32_t w = w0 + cs.st * (kImg >> 4);
#pragma unroll
          for (int pass = 0; pass < 3; ++pass)
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              mma_ts2(tmem + cDx, (pass == 0 ? cDhid : cPre) + ks * 8, w + ((kW1T + (pass == 1 ? 4096 : 0) + ks * 256) >> 4), hiW1,
                      idesc, (cs.m | pass | ks) != 0);
        }
        mma_commit(bar(Q_WEMPTY0 + cs.st));
        if (last) mma_commit(bar(Q_DXFULL));
        TL(22, s);
        // ---- the row-local products two steps ahead (the epilogue of step s + 1 stores dout behind them) ----
        if (s + 2 < total) issue_pd(s + 2, ct), advance(ct);
        TL(23, s);
        advance(cs);
      }
    }
    __syncwarp();
  } else if (warp == kMmaG) {
    // ================================== second MMA issuer: the gradient products ==================================
    // (A satisfied mbarrier wait costs this lone thread a few hundred cycles and an MMA block blocks it while the tensor
    // pipe works through the queue, so one issuer for everything is the bottleneck of the step; the two issuers' MMAs touch
    // disjoint accumulators.)
    if (elect_one()) {
      constexpr uint32_t idesc_g = make_idesc_bmn(128, 64);
      constexpr uint32_t hiG = (512u >> 4) | (1u << 14) | (1u << 29);                            // SBO | version | 128B-swizzle, 32B base
      constexpr uint32_t lboG = (16384u >> 4) << 16;
      const uint32_t b0 = ((sbase + kOffB) >> 4) | lboG;
      uint32_t m = 0, it = 0;
      for (uint32_t s = 0; s < total; ++s) {
        const uint32_t par = s & 1, ph = (s >> 1) & 1;
        // ---- D = A_T . [dpre | hid]  (K = the tile's 128 rows; B_lo first) ----
        mbar_wait_hot(bar(Q_DPREFULL0 + par), ph);           // B of step s
        mbar_wait_hot(bar(Q_ATDFULL0 + par), ph);
        TL(26, s);
        if (m == 0) mbar_wait_hot(bar(Q_ATXFULL), it & 1);
        mbar_wait_hot(bar(Q_DFREE0 + par), ph ^ 1);
        TL(24, s);
        tc_fence_after();
        {
          const uint32_t bb = b0 + par * (kBBuf >> 4), dcol = tmem + cD + 64 * par;
#pragma unroll
          for (int pass = 0; pass < 2; ++pass)
#pragma unroll
            for (int ks = 0; ks < 16; ++ks)
              mma_ts2(dcol, tmem + cAT + ks * 8, bb + (((pass == 0 ? 32768 : 0) + ks * 1024) >> 4), hiG, idesc_g, (pass | ks) != 0);
        }
        mma_commit(bar(Q_GRADDONE0 + par));
        TL(25, s);
        if (++m == (uint32_t)n_gen) m = 0, ++it;
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
  } else if (warp == kTmaD) {
    // ================================================ TMA: dout staging ================================================
    // (its own issuer: dW streams from HBM, so it has to run as far ahead as the staging ring allows, not in step with the
    // weight ring)
    // Lane 0 waits for the buffer and announces the bytes; the copies of a tile (one per tap, or one per four rows of an
    // interleaved dV0) are issued by the lanes side by side, so that the ring never waits for address arithmetic.
    uint32_t m = 0, it = 0;
    for (uint32_t s = 0; s < total; ++s) {
      const uint32_t b = s % kDsStages;
      const bool is_v = m >= (uint32_t)p.n_W;
      const uint32_t dst = sbase + kOffDs + b * kDsBuf, full = bar(Q_DSFULL0 + b);
      if (lane == 0) {
        mbar_wait_relaxed(bar(Q_DSFREE0 + b), ((s / kDsStages) & 1) ^ 1);
        mbar_expect_tx(full, (uint32_t)(is_v ? p.C : p.L) * 512u);
      }
      __syncwarp();
      if (!is_v) {
        // dW of generator m, this tile: one 512-byte run per tap
        if (lane < p.L)
          tma_load_1d(dst + (uint32_t)lane * (kDsRow * 4u), p.Wt + (long long)m * p.sWf + (long long)lane * p.sWt + tile_row0(it), 512u, full);
      } else if (p.Q <= 1) {
        // the V generator's dV0 is [rows, C] in natural row order: the tile as it lies
        if (lane == 0) tma_load_1d(dst, p.V + tile_row0(it) * p.C, (uint32_t)p.C * 512u, full);
      } else {
        // x rows quad-interleaved: four natural rows are contiguous
        tma_load_1d(dst + (uint32_t)lane * (uint32_t)p.C * 16u, p.V + natural_row(p, tile_row0(it) + 4 * lane) * p.C, (uint32_t)p.C * 16u, full);
      }
      if (++m == (uint32_t)n_gen) m = 0, ++it;
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
extern "C" int psf_debug_timeline(unsigned* out) {     // out: 24 warps x kTlPerWarp x (code, clock); code 0 = unused
  cudaMemcpyFromSymbol(out, g_timeline, sizeof(uint2) * 24 * kTlPerWarp);
  void* sym = nullptr;
  cudaGetSymbolAddress(&sym, g_timeline);
  cudaMemset(sym, 0, sizeof(uint2) * 24 * kTlPerWarp);
  return kTlPerWarp;
}
#endif

// Shapes the kernel covers: whole 128-row tiles, tap-major dW with 16-byte aligned 512-byte runs per tap.
int psf_gen_tc_bwd2(const GenParams& p, const void* images, int* grid_out, cudaStream_t st) {
  // tap lane 15 of the dout^T strips is the row of ones that yields db1
  // (chains of fewer than three generators take the first-version kernel: the look-ahead of this one - products two steps
  // ahead, x three steps ahead - is only exercised by the tests from three generators per tile on)
  if (p.n_gen < 3 || p.L > GLP - 1 || p.C > GLP - 1 || p.pre != nullptr) return PSF_ERR_UNSUPPORTED;
  if ((p.rows & 127) || (p.sWt & 3) || (p.sWf & 3) || (reinterpret_cast<uintptr_t>(p.Wt) & 15)) return PSF_ERR_UNSUPPORTED;
  if (p.n_gen > p.n_W && (reinterpret_cast<uintptr_t>(p.V) & 15)) return PSF_ERR_UNSUPPORTED;      // dV0 tiles go through TMA
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
