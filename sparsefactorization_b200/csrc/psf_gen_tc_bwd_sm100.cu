// Backward pass of the generators on tcgen05 + tensor memory (sm_100a); forward: psf_gen_tc_sm100.cu.
//
// For a 128-row tile and generator m = Linear(32,32) -> GELU -> Linear(32,L) (SyntheticExperiments/psf.py:110-126):
//
//   tcgen05 (A operand in TENSOR MEMORY, lane = row; B = weights in shared memory, K-major, staged per generator by TMA)
//     pre  [128 x 32] = x    [128 x 32] . W1^T        (recomputed: saving it would be 4 GB)
//     dhid [128 x 32] = dout [128 x 16] . W2
//     dx   [128 x 32] += dpre[128 x 32] . W1          (one accumulator over all generators of the tile)
//   epilogue warps (thread = row): tcgen05.ld of pre / dhid, Phi / phi, hid = pre Phi, dpre = dhid (Phi + pre phi);
//     dpre goes back to tensor memory as the A operand of the dx product
//   parameter gradients (K = rows: tcgen05 would need the operands transposed, and batching four generators into the
//     128 lanes does not fit shared / tensor memory - DESIGN.md): dW1 += dpre^T x as FP32 FMA outer products, dW2 += dout^T
//     hid as warp-level mma.sync 3xTF32 (two different pipes), over the warp's own 32 rows; the four row quarters'
//     partial results are summed through shared memory and added to the CTA's slice of a global workspace
//
// Every product is error-compensated 3xTF32 (psf_gen_tc_sm100.cu); small terms first.
//
// Warp roles (704 threads, one CTA per SM, persistent):
//   warps 0-15  two epilogue groups of eight warps (two warps per 32 rows, 16 hidden units each); consecutive
//               generators alternate between the groups
//   warps 16-19 x loader (tensor memory + shared memory) and dx writer
//   warp 20     tensor-memory allocation + the MMA-issuing thread
//   warp 21     one thread streaming the per-generator weight images (TMA bulk copies, 4-stage ring)
#include "psf_gen.cuh"
#include "psf_tc.cuh"

namespace {

// timing experiments only (results become wrong): 1 no red.global, 2 no mma.sync loop, 4 no dout loads, 8 no Phi/phi math
#ifndef PSF_TC_ABLATE
#define PSF_TC_ABLATE 0
#endif
constexpr int kAbl = PSF_TC_ABLATE;

constexpr int kWarpLoad0 = 16, kWarpMma = 20, kWarpTma = 21;
constexpr int kThreadsBwd = (kWarpTma + 1) * 32;
constexpr int XS = GE + 4;    // padded row of the staged fp32 tiles (conflict-free fragment loads)
constexpr int DS = GLP + 4;
// weight image of one generator (bytes): W1 [n=h][k=e] hi, lo | W1^T [n=e][k=h] hi, lo | W2^T [n=h][k=l] hi, lo | b1 (fp32)
constexpr uint32_t kImgW1 = 0, kImgW1T = 8192, kImgW2T = 16384, kImgB1 = 20480, kImgBytes = 20736;
constexpr int kStages = 4;
// shared memory
constexpr uint32_t kOffXs = kStages * kImgBytes;                     // x tile, [128][XS] floats
constexpr uint32_t kOffStage = kOffXs + 128 * XS * 4;                // per warpgroup: dps, hs, dos
constexpr uint32_t kStageBytes = (2 * 128 * XS + 128 * DS) * 4;
constexpr uint32_t kOffBarB = kOffStage + 2 * kStageBytes;
constexpr uint32_t kSmemBwdTC = kOffBarB + 256;
// tensor-memory columns
constexpr uint32_t kTX = 0, kTXlo = 32, kTDx = 64, kTWg = 96, kTWgStride = 160;
constexpr uint32_t kTPre = 0, kTDhid = 32, kTDout = 64, kTDoutLo = 80, kTDpre = 96, kTDpreLo = 128;

enum {
  B_XFULL = 0, B_XEMPTY, B_XSFULL, B_XSFREE, B_DXFULL, B_DXEMPTY,
  B_WFULL0, B_WEMPTY0 = B_WFULL0 + kStages,
  B_DOUTFULL0 = B_WEMPTY0 + kStages, B_PDFULL0 = B_DOUTFULL0 + 2, B_DPREFULL0 = B_PDFULL0 + 2, B_DPREFREE0 = B_DPREFULL0 + 2,
  B_COUNT = B_DPREFREE0 + 2
};
static_assert(B_COUNT * 8 + 8 <= 256, "barrier block");

__device__ __forceinline__ void red_add4g(float* p, float a, float b, float c, float d) {
  if (kAbl & 1) return;
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add1g(float* p, float a) {
  if (kAbl & 1) return;
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}
// truncating split for the mma.sync operands (the tensor core ignores the low 13 bits of hi and of lo)
__device__ __forceinline__ void split_trunc(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32_sync(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void ffma2_tc(float2& c, const float2& a, const float2& b) {
  unsigned long long* pc = reinterpret_cast<unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(*pc)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
}
__device__ __forceinline__ void wg_bar(int wg) { asm volatile("bar.sync %0, 256;" ::"r"(1 + wg) : "memory"); }

// Phi(p) and phi(p) from one exp(-p^2/2): erfc by the 6-term A&S-type fit of psf_gen_tc_sm100.cu (|error| < 1e-7 in fp32)
__device__ __forceinline__ void normal_cdf_pdf(float p, float& cdf, float& pdf) {
  const float ax = fabsf(p);
  const float t = rcp_approx(fmaf(ax, 0.27599917f, 1.f));
  float q = fmaf(-0.22690525f, t, 0.88164259f);
  q = fmaf(q, t, -0.62774200f);
  q = fmaf(q, t, 0.64431879f);
  q = fmaf(q, t, 0.093436931f);
  q = fmaf(q, t, 0.23524893f);
  const float e = ex2_approx(p * p * -0.72134752f);
  const float tail = 0.5f * (q * t) * e;          // Phi(-|p|)
  cdf = p >= 0.f ? 1.f - tail : tail;
  pdf = 0.39894228f * e;
}

// weights -> MMA-ready images (3xTF32 hi / lo, K-major core-matrix layout), once per backward call
__global__ void psf_gen_tc_prep_kernel(const GenParams p, uint8_t* __restrict__ img) {
  const int m = blockIdx.x;
  const float* W1 = m < p.n_W ? p.W1 + (long long)m * GH * GE : p.gW1;
  const float* W2 = m < p.n_W ? p.W2 + (long long)m * p.L * GH : p.gW2;
  const int Lm = m < p.n_W ? p.L : p.C;
  uint8_t* dst = img + (size_t)m * kImgBytes;
  for (int idx = threadIdx.x; idx < GH * GE; idx += blockDim.x) {
    const int h = idx >> 5, e = idx & 31;
    const float v = W1[idx];
    const uint32_t o1 = kImgW1 + (h >> 3) * 1024 + (e >> 2) * 128 + (h & 7) * 16 + (e & 3) * 4;
    const uint32_t o2 = kImgW1T + (e >> 3) * 1024 + (h >> 2) * 128 + (e & 7) * 16 + (h & 3) * 4;
    *reinterpret_cast<float*>(dst + o1) = tf32_hi(v);
    *reinterpret_cast<float*>(dst + o1 + 4096) = tf32_lo(v);
    *reinterpret_cast<float*>(dst + o2) = tf32_hi(v);
    *reinterpret_cast<float*>(dst + o2 + 4096) = tf32_lo(v);
  }
  if (threadIdx.x < GH) reinterpret_cast<float*>(dst + kImgB1)[threadIdx.x] = (m < p.n_W ? p.b1 + m * GH : p.gb1)[threadIdx.x];
  for (int idx = threadIdx.x; idx < GLP * GH; idx += blockDim.x) {
    const int l = idx >> 5, h = idx & 31;
    const float v = l < Lm ? W2[l * GH + h] : 0.f;
    const uint32_t o = kImgW2T + (h >> 3) * 512 + (l >> 2) * 128 + (h & 7) * 16 + (l & 3) * 4;
    *reinterpret_cast<float*>(dst + o) = tf32_hi(v);
    *reinterpret_cast<float*>(dst + o + 2048) = tf32_lo(v);
  }
}

__global__ void __launch_bounds__(kThreadsBwd, 1) psf_gen_bwd_tc_kernel(const GenParams p, const uint8_t* __restrict__ img) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + kOffBarB;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + kOffBarB + B_COUNT * 8);
  const int n_gen = p.n_gen;
  float* xs = reinterpret_cast<float*>(smem + kOffXs);

  if (tid == 0) {
    mbar_init(bars + 8 * B_XFULL, 4);
    mbar_init(bars + 8 * B_XEMPTY, 1);
    mbar_init(bars + 8 * B_XSFULL, 4);
    mbar_init(bars + 8 * B_XSFREE, 16);
    mbar_init(bars + 8 * B_DXFULL, 1);
    mbar_init(bars + 8 * B_DXEMPTY, 4);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(bars + 8 * (B_WFULL0 + i), 1);
      mbar_init(bars + 8 * (B_WEMPTY0 + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(bars + 8 * (B_DOUTFULL0 + i), 8);
      mbar_init(bars + 8 * (B_PDFULL0 + i), 1);
      mbar_init(bars + 8 * (B_DPREFULL0 + i), 8);
      mbar_init(bars + 8 * (B_DPREFREE0 + i), 1);
    }
    fence_barrier_init();
  }
  if (warp == kWarpMma) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  const long long ntiles = (p.rows + 127) >> 7;
  const int my_tiles = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const uint32_t total = (uint32_t)my_tiles * (uint32_t)n_gen;     // (tile, generator) steps of this CTA, in order
  float* wsc = p.ws + (long long)blockIdx.x * n_gen * kAccBlk;

  if (warp < kWarpLoad0) {
    // ======================================= epilogue warpgroups =======================================
    // warp = (warpgroup, half, quarter): rows 32 quarter .. of the tile (its tensor-memory lanes), hidden units / x columns
    // 16 half .., taps 8 half ..  Two warps per row halve the registers per thread, which doubles the warps per scheduler
    // that hide each other's MUFU / HMMA / tensor-memory latencies (the per-thread work is one long dependent chain).
    const int wg = warp >> 3, half = (warp >> 2) & 1, quarter = warp & 3;
    const int g = lane >> 2, t = lane & 3;
    const int r0 = quarter * 32, rl = r0 + lane;
    const int c0 = half * 16, t0 = half * 8;
    const uint32_t lane_base = ((uint32_t)quarter * 32u) << 16;
    const uint32_t twg = tmem + lane_base + kTWg + kTWgStride * wg;
    float* dps = reinterpret_cast<float*>(smem + kOffStage + wg * kStageBytes);
    float* hs = dps + 128 * XS;
    float* dos = hs + 128 * XS;
    const uint32_t bar_dout = bars + 8 * (B_DOUTFULL0 + wg), bar_pd = bars + 8 * (B_PDFULL0 + wg);
    const uint32_t bar_dpre = bars + 8 * (B_DPREFULL0 + wg), bar_free = bars + 8 * (B_DPREFREE0 + wg);
    const int pair_id = 3 + wg * 4 + quarter;      // named barrier of the two warps that share the rows

    auto row_of = [&](uint32_t s) { return (blockIdx.x + (long long)(s / (uint32_t)n_gen) * gridDim.x) * 128 + rl; };
    auto load_dout = [&](uint32_t s, float (&d)[8]) {      // this warp's 8 taps of its rows
      const int m = (int)(s % (uint32_t)n_gen);
      const long long row = row_of(s);
      const bool live = row < p.rows && !(kAbl & 4);
      if (m < p.n_W) {
        const float* src = p.Wt + (long long)m * p.sWf + row;
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = (live && t0 + i < p.L) ? __ldcs(src + (long long)(t0 + i) * p.sWt) : 0.f;
      } else {
        const float* src = p.V + (live ? natural_row(p, row) : 0) * p.C;
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = (live && t0 + i < p.C) ? __ldcs(src + t0 + i) : 0.f;
      }
    };
    // dout -> tensor memory (A operand of dhid = dout W2), hi then lo
    auto publish_dout = [&](const float (&d)[8]) {
      float hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        hi[i] = tf32_hi(d[i]);
        lo[i] = d[i] - hi[i];
      }
      tmem_st8f(twg + kTDout + t0, hi);
      tmem_st8f(twg + kTDoutLo + t0, lo);
      tmem_wait_st_tc();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_dout);
    };
    auto stage_dout = [&](const float (&d)[8]) {
      *reinterpret_cast<float4*>(dos + rl * DS + t0) = make_float4(d[0], d[1], d[2], d[3]);
      *reinterpret_cast<float4*>(dos + rl * DS + t0 + 4) = make_float4(d[4], d[5], d[6], d[7]);
    };

    float dn[8], dnn[8];      // dout of this warpgroup's next generator and of the one after (streamed from HBM)
    uint32_t k = 0;
    if ((uint32_t)wg < total) {
      load_dout((uint32_t)wg, dn);
      publish_dout(dn);
      stage_dout(dn);
      if ((uint32_t)wg + 2 < total) load_dout((uint32_t)wg + 2, dn);
    }
    for (uint32_t s = (uint32_t)wg; s < total; s += 2, ++k) {
      const int m = (int)(s % (uint32_t)n_gen);
      const uint32_t it = s / (uint32_t)n_gen;
      const bool has_next = s + 2 < total;
      if (s + 4 < total) load_dout(s + 4, dnn);           // two generators ahead: a full iteration to land
      const float* b1 = reinterpret_cast<const float*>(smem + (s % kStages) * kImgBytes + kImgB1) + c0;   // staged with the weights
      float pre[16], dh[16];
      mbar_wait_tc(bars + 8 * (B_WFULL0 + s % kStages), (s / kStages) & 1);   // b1 (TMA write) visible to this thread
      mbar_wait_tc(bar_pd, k & 1);
      tc_fence_after();
      tmem_ld16f(twg + kTPre + c0, pre);
      tmem_ld16f(twg + kTDhid + c0, dh);
      tmem_wait_ld_tc();
      if (has_next) publish_dout(dn);                     // pre / dhid of the next generator are computed meanwhile
#pragma unroll
      for (int h4 = 0; h4 < 4; ++h4) {
        const float4 b = *(reinterpret_cast<const float4*>(b1) + h4);
        const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int h = 4 * h4 + j;
          const float pv = tc_unbias(pre[h]) + bb[j];
          float cdf, pdf;
          if (kAbl & 8) { cdf = 0.5f; pdf = 0.25f; } else normal_cdf_pdf(pv, cdf, pdf);
          pre[h] = dh[h] * fmaf(pv, pdf, cdf);            // dpre
          dh[h] = pv * cdf;                               // hid
        }
      }
      // stage for the parameter-gradient products (this warp's rows and columns)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        *reinterpret_cast<float4*>(dps + rl * XS + c0 + 4 * i) = make_float4(pre[4 * i], pre[4 * i + 1], pre[4 * i + 2], pre[4 * i + 3]);
        *reinterpret_cast<float4*>(hs + rl * XS + c0 + 4 * i) = make_float4(dh[4 * i], dh[4 * i + 1], dh[4 * i + 2], dh[4 * i + 3]);
      }
      // dpre -> tensor memory (A operand of dx += dpre W1)
      mbar_wait_tc(bar_free, (k & 1) ^ 1);
      tc_fence_after();
      {
#pragma unroll
        for (int h = 0; h < 16; ++h) dh[h] = tf32_hi(pre[h]);
        tmem_st16f(twg + kTDpre + c0, dh);
#pragma unroll
        for (int h = 0; h < 16; ++h) pre[h] -= dh[h];
        tmem_st16f(twg + kTDpreLo + c0, pre);
        tmem_wait_st_tc();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_dpre);
      }
      mbar_wait_tc(bars + 8 * B_XSFULL, it & 1);          // x of this tile is in shared memory
      asm volatile("bar.sync %0, 64;" ::"r"(pair_id) : "memory");   // the partner's columns of these rows are staged

      // ---- dW1[h][e] += dpre^T x (this warp: h in its 16 columns), dW2[l][h] += dout^T hid (h in its 16 columns) over
      //      the warp's 32 rows: mma.sync m16n8k8, 3xTF32 ----
      float* acc_m = wsc + (long long)m * kAccBlk;
      // dW1: lane owns the 4 x 4 block h = c0 + 4 (lane >> 3) .., e = 4 (lane & 7) .. and walks the warp's 32 rows
      // (FP32 FMA: the legacy mma.sync path costs 16 cycles per m16n8k8 on this part, three of them per product)
      float2 w1a[4][2];
      float d2[2][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) w1a[i][0] = w1a[i][1] = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) d2[j][q] = 0.f;
      if (!(kAbl & 2)) {
        const float* dp = dps + r0 * XS + c0 + 4 * (lane >> 3);
        const float* xp = xs + r0 * XS + 4 * (lane & 7);
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
          const float4 a = *reinterpret_cast<const float4*>(dp + r * XS);
          const float4 b = *reinterpret_cast<const float4*>(xp + r * XS);
          const float2 b01 = make_float2(b.x, b.y), b23 = make_float2(b.z, b.w);
          const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 aa = make_float2(av[i], av[i]);
            ffma2_tc(w1a[i][0], aa, b01);
            ffma2_tc(w1a[i][1], aa, b23);
          }
        }
      }
#pragma unroll 1
      for (int ks = 0; ks < ((kAbl & 2) ? 0 : 4); ++ks) {        // dW2[l][h] += dout^T hid: mma.sync m16n8k8, 3xTF32
        const int ra = r0 + ks * 8 + t, rb = ra + 4;               // rows k = t and k = t + 4 of this step
        uint32_t oh[4], ol[4];                                     // A = dout[k][l]^T
        split_trunc(dos[ra * DS + g], oh[0], ol[0]);
        split_trunc(dos[ra * DS + g + 8], oh[1], ol[1]);
        split_trunc(dos[rb * DS + g], oh[2], ol[2]);
        split_trunc(dos[rb * DS + g + 8], oh[3], ol[3]);
        uint32_t hh[2][2], hl[2][2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {                              // B = hid[k][h]
          split_trunc(hs[ra * XS + c0 + j * 8 + g], hh[j][0], hl[j][0]);
          split_trunc(hs[rb * XS + c0 + j * 8 + g], hh[j][1], hl[j][1]);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) mma_tf32_sync(d2[j], ol, hh[j]);
#pragma unroll
        for (int j = 0; j < 2; ++j) mma_tf32_sync(d2[j], oh, hl[j]);
#pragma unroll
        for (int j = 0; j < 2; ++j) mma_tf32_sync(d2[j], oh, hh[j]);
      }
      // bias gradients: column sums over the warp's rows (lane = column): db1 by the first half, db2 by the second
      if (half == 0) {
        float sb1 = 0.f;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) sb1 += dps[(r0 + r) * XS + lane];
        red_add1g(acc_m + GH * GE + lane, sb1);
      } else {
        float sb2 = 0.f;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) sb2 += dos[(r0 + r) * DS + (lane & (GLP - 1))];
        if (lane < GLP) red_add1g(acc_m + GH * GE + GH + GLP * GH + lane, sb2);
      }
      // cross-warp sum through shared memory (the dpre / hid tiles are dead once all eight warps are here)
      constexpr int PS = 40;
      static_assert(2 * 128 * XS >= 4 * 48 * PS, "partial sums do not fit the dead tiles");
      float* part = dps;
      wg_bar(wg);
      float* mine = part + quarter * 48 * PS;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(mine + (c0 + 4 * (lane >> 3) + i) * PS + 4 * (lane & 7)) =
            make_float4(w1a[i][0].x, w1a[i][0].y, w1a[i][1].x, w1a[i][1].y);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        *reinterpret_cast<float2*>(mine + (32 + g) * PS + c0 + j * 8 + 2 * t) = make_float2(d2[j][0], d2[j][1]);
        *reinterpret_cast<float2*>(mine + (32 + g + 8) * PS + c0 + j * 8 + 2 * t) = make_float2(d2[j][2], d2[j][3]);
      }
      wg_bar(wg);
      for (int u = (warp & 7) * 32 + lane; u < 48 * 8; u += 256) {
        const int prow = u >> 3, col = (u & 7) * 4;
        float4 sum = *reinterpret_cast<const float4*>(part + prow * PS + col);
#pragma unroll
        for (int w = 1; w < 4; ++w) {
          const float4 v = *reinterpret_cast<const float4*>(part + (w * 48 + prow) * PS + col);
          sum.x += v.x;
          sum.y += v.y;
          sum.z += v.z;
          sum.w += v.w;
        }
        float* dst = prow < 32 ? acc_m + prow * GE + col : acc_m + GH * GE + GH + (prow - 32) * GH + col;
        red_add4g(dst, sum.x, sum.y, sum.z, sum.w);
      }
      wg_bar(wg);
      if (!has_next || (s + 2) / (uint32_t)n_gen != it) {   // this warpgroup is done with the tile's x
        if (lane == 0) mbar_arrive(bars + 8 * B_XSFREE);
      }
      if (has_next) stage_dout(dn);
#pragma unroll
      for (int i = 0; i < 8; ++i) dn[i] = dnn[i];
    }
  } else if (warp < kWarpMma) {
    // ======================================= x loader / dx writer =======================================
    const int quarter = warp & 3;
    const int rl = quarter * 32 + lane;
    const uint32_t lane_base = ((uint32_t)quarter * 32u) << 16;
    auto write_dx = [&](int it) {
      const long long row = (blockIdx.x + (long long)it * gridDim.x) * 128 + rl;
      float v[GE];
      mbar_wait_relaxed(bars + 8 * B_DXFULL, it & 1);
      tc_fence_after();
      tmem_ld32f(tmem + lane_base + kTDx, v);
      tmem_wait_ld_tc();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * B_DXEMPTY);
      if (row < p.rows) {
        float4* dst = reinterpret_cast<float4*>(p.dx + row * GE);
#pragma unroll
        for (int i = 0; i < GE / 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    };
    for (int it = 0; it < my_tiles; ++it) {
      const long long row = (blockIdx.x + (long long)it * gridDim.x) * 128 + rl;
      float4 xin[GE / 4];
      if (row < p.rows) {
        const float4* src = reinterpret_cast<const float4*>(p.x + row * GE);
#pragma unroll
        for (int i = 0; i < GE / 4; ++i) xin[i] = __ldg(src + i);
      } else {
#pragma unroll
        for (int i = 0; i < GE / 4; ++i) xin[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      mbar_wait_relaxed(bars + 8 * B_XEMPTY, (it & 1) ^ 1);     // the pre products of the previous tile are done
      tc_fence_after();
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {                           // 16 columns at a time (register pressure)
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
        tmem_st16f(tmem + lane_base + kTX + 16 * hf, vh);
        tmem_st16f(tmem + lane_base + kTXlo + 16 * hf, vl);
      }
      tmem_wait_st_tc();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * B_XFULL);
      if (it > 0) write_dx(it - 1);
      mbar_wait_relaxed(bars + 8 * B_XSFREE, (it & 1) ^ 1);     // the mma.sync phases of the previous tile are done
#pragma unroll
      for (int i = 0; i < GE / 4; ++i) *reinterpret_cast<float4*>(xs + rl * XS + 4 * i) = xin[i];
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * B_XSFULL);
    }
    if (my_tiles > 0) write_dx(my_tiles - 1);
  } else if (warp == kWarpMma) {
    // ======================================= MMA issuer =======================================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc(128, 32);
      auto issue_pd = [&](uint32_t s) {       // pre = x W1^T and dhid = dout W2 of step s
        const uint32_t m = s % (uint32_t)n_gen, it = s / (uint32_t)n_gen, wg = s & 1, stage = s % kStages;
        if (m == 0) mbar_wait_tc(bars + 8 * B_XFULL, it & 1);
        mbar_wait_tc(bars + 8 * (B_WFULL0 + stage), (s / kStages) & 1);
        mbar_wait_tc(bars + 8 * (B_DOUTFULL0 + wg), (s >> 1) & 1);
        tc_fence_after();
        const uint32_t w = sbase + stage * kImgBytes;
        const uint32_t twg = tmem + kTWg + kTWgStride * wg;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {   // x_lo W1_hi, x_hi W1_lo, x_hi W1_hi
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_ts(twg + kTPre, tmem + (pass == 0 ? kTXlo : kTX) + ks * 8,
                   make_desc(w + kImgW1 + (pass == 1 ? 4096 : 0) + ks * 256, 128, 1024), idesc, (pass | ks) != 0);
        }
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {   // dout_lo W2_hi, dout_hi W2_lo, dout_hi W2_hi  (K = 16 taps)
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
            mma_ts(twg + kTDhid, twg + (pass == 0 ? kTDoutLo : kTDout) + ks * 8,
                   make_desc(w + kImgW2T + (pass == 1 ? 2048 : 0) + ks * 256, 128, 512), idesc, (pass | ks) != 0);
        }
        mma_commit(bars + 8 * (B_PDFULL0 + wg));
        if (m == (uint32_t)n_gen - 1) mma_commit(bars + 8 * B_XEMPTY);
      };
      auto issue_dx = [&](uint32_t s) {       // dx += dpre W1 of step s
        const uint32_t m = s % (uint32_t)n_gen, it = s / (uint32_t)n_gen, wg = s & 1, stage = s % kStages;
        mbar_wait_tc(bars + 8 * (B_DPREFULL0 + wg), (s >> 1) & 1);
        if (m == 0) mbar_wait_tc(bars + 8 * B_DXEMPTY, (it & 1) ^ 1);
        tc_fence_after();
        const uint32_t w = sbase + stage * kImgBytes;
        const uint32_t twg = tmem + kTWg + kTWgStride * wg;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {   // dpre_lo W1_hi, dpre_hi W1_lo, dpre_hi W1_hi
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            mma_ts(tmem + kTDx, twg + (pass == 0 ? kTDpreLo : kTDpre) + ks * 8,
                   make_desc(w + kImgW1T + (pass == 1 ? 4096 : 0) + ks * 256, 128, 1024), idesc, (m | pass | ks) != 0);
        }
        mma_commit(bars + 8 * (B_DPREFREE0 + wg));
        mma_commit(bars + 8 * (B_WEMPTY0 + stage));
        if (m == (uint32_t)n_gen - 1) mma_commit(bars + 8 * B_DXFULL);
      };
      if (total > 0) issue_pd(0);
      if (total > 1) issue_pd(1);
      for (uint32_t s = 0; s < total; ++s) {
        issue_dx(s);
        if (s + 2 < total) issue_pd(s + 2);
      }
    }
    __syncwarp();
  } else {
    // ======================================= weight streamer =======================================
    if (lane == 0) {
      for (uint32_t s = 0; s < total; ++s) {
        const uint32_t stage = s % kStages, m = s % (uint32_t)n_gen;
        mbar_wait_relaxed(bars + 8 * (B_WEMPTY0 + stage), ((s / kStages) & 1) ^ 1);
        mbar_expect_tx(bars + 8 * (B_WFULL0 + stage), kImgBytes);
        tma_load_1d(sbase + stage * kImgBytes, img + (size_t)m * kImgBytes, kImgBytes, bars + 8 * (B_WFULL0 + stage));
      }
    }
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kWarpMma) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

struct DevAttrB {
  bool done = false;
  int status = 0;
};
DevAttrB g_tcb_attr[PSF_MAX_DEVICES];

}  // namespace

size_t psf_gen_tc_bwd_image_bytes(int n_gen) { return (size_t)n_gen * kImgBytes; }

int psf_gen_tc_prep(const GenParams& p, void* images, cudaStream_t st) {
  psf_gen_tc_prep_kernel<<<p.n_gen, 256, 0, st>>>(p, (uint8_t*)images);
  return (int)cudaGetLastError();
}

int psf_gen_tc_bwd(const GenParams& p, void* images, int* grid_out, cudaStream_t st) {
  if (p.n_gen < 2 || p.L > GLP || p.C > GLP || p.pre != nullptr) return PSF_ERR_UNSUPPORTED;
  if (reinterpret_cast<uintptr_t>(images) & 127) return PSF_ERR_UNSUPPORTED;
  const int dev = psf_current_device();
  if (dev < 0) return PSF_ERR_UNSUPPORTED;
  {
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (!g_tcb_attr[dev].done) {
    cudaError_t e = cudaFuncSetAttribute(psf_gen_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBwdTC);
    g_tcb_attr[dev].status = (int)e;
    g_tcb_attr[dev].done = true;
    if (e != cudaSuccess) cudaGetLastError();
    }
  }
  if (g_tcb_attr[dev].status != 0) return PSF_ERR_UNSUPPORTED;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long ntiles = (p.rows + 127) >> 7;
  const long long grid = ntiles < sms ? ntiles : sms;
  psf_gen_bwd_tc_kernel<<<(unsigned)grid, kThreadsBwd, kSmemBwdTC, st>>>(p, (const uint8_t*)images);
  *grid_out = (int)grid;
  return (int)cudaGetLastError();
}
