// Generators on the 5th-generation tensor cores (tcgen05 + tensor memory), sm_100a.
//
// Every fs[m] = Linear(32,32) -> GELU -> Linear(32,L) of PSFNet and the V generator g
// (SyntheticExperiments/psf.py:110-126, evaluated at psf.py:165 and :175) for a 128-row tile at a time:
//
//   layer 1   pre[128 x 128] = x[128 x 32] . W1^T   for four generators at once (N = 4 x 32 hidden units)
//             A = x from TENSOR MEMORY (lane = row, column = e), B = W1 from shared memory (K-major, no swizzle),
//             accumulator in tensor memory, two stages
//   epilogue  one thread per row: tcgen05.ld of its 32 pre-activations, + b1, exact (erf) GELU, hidden units
//             written back to tensor memory as the A operand of layer 2
//   layer 2   out[128 x 16] = hid[128 x 32] . W2^T, A from tensor memory, B = W2 from shared memory; tcgen05.ld,
//             + b2, coalesced stores in the chain's tap-major streaming layout (V0 in its natural [rows, C] layout)
//
// Arithmetic: kind::tf32 truncates its operands to 10 mantissa bits (measured: tools/microbench/umma_probe.cu), so
// every product runs as the error-compensated 3xTF32 split  a b ~ a_lo b_hi + a_hi b_lo + a_hi b_hi  with
// a_hi = a rounded to tf32, a_lo = a - a_hi (exact in fp32), fp32 accumulation; the small
// terms are issued first so that the accumulator only sees four roundings at full magnitude.  Result: within a few
// fp32 ulps of the FP32-FMA kernel (psf_gen_sm100.cu), far inside the chain's rtol 1e-5 / atol 1e-6.
//
// Why A lives in tensor memory: a tcgen05.mma with N <= 64 whose A operand comes from shared memory costs ~40 cycles
// (the 4 KB A read), from tensor memory it runs at its N/2-cycle floor (9 cycles at N = 16; same probe).
//
// Warp roles (416 threads, one CTA per SM, persistent over tiles):
//   warps 0-7   two epilogue warpgroups (warp w owns TMEM lanes 32 (w % 4) ..; the groups alternate generators)
//   warps 8-11  x loader: global -> registers -> (hi, lo) -> tensor memory
//   warp 12     TMEM allocation and the single MMA-issuing thread
#include "psf_gen.cuh"
#include "psf_tc.cuh"

namespace {

constexpr int kEpiWarps = 8, kLoadWarps = 4;
constexpr int kMmaWarp = kEpiWarps + kLoadWarps;
constexpr int kThreadsTC = (kMmaWarp + 1) * 32;
constexpr int kMaxGenTC = 16;             // four chunks of four generators
// tensor-memory columns
constexpr uint32_t kColXhi = 0, kColXlo = 32, kColPre = 64, kColHid = 320, kColOut = 448;
// shared memory (bytes)
constexpr uint32_t kW1Chunk = 32768;      // 128 rows x 128 B, hi then lo
constexpr uint32_t kW2Gen = 4096;         // 16 rows x 128 B, hi then lo
constexpr uint32_t kOffW2 = 4 * kW1Chunk;
constexpr uint32_t kOffB1 = kOffW2 + kMaxGenTC * kW2Gen;
constexpr uint32_t kOffB2 = kOffB1 + kMaxGenTC * GH * 4;
constexpr uint32_t kOffBar = kOffB2 + kMaxGenTC * GLP * 4;
constexpr uint32_t kSmemTC = kOffBar + 128;

enum { BAR_XFULL = 0, BAR_XEMPTY, BAR_PREFULL0, BAR_PREFULL1, BAR_PREEMPTY0, BAR_PREEMPTY1, BAR_HIDFULL0, BAR_HIDFULL1,
       BAR_L2DONE0, BAR_L2DONE1, BAR_COUNT };

// Exact (erf) GELU, x Phi(x), as  max(x, 0) - |x|/2 * erfcx(|x|/sqrt 2) * exp(-x^2/2)  with
// erfcx(z) ~ t P6(t), t = 1/(1 + p z): a refit of Abramowitz-Stegun 7.1.26 with one more term (|erf error| < 8e-9 in
// exact arithmetic); in fp32 the result is within 2.5e-7 absolute of the true value over the whole range - the same
// as evaluating 0.5 x (1 + erff(x / sqrt 2)) in fp32 (4.5e-7).  Two MUFU + 13 FP32 instructions.
__device__ __forceinline__ float gelu_erf(float v) {
  const float ax = fabsf(v);
  const float t = rcp_approx(fmaf(ax, 0.27599917f, 1.f));          // p / sqrt 2, p = 0.39032176
  float q = fmaf(-0.22690525f, t, 0.88164259f);
  q = fmaf(q, t, -0.62774200f);
  q = fmaf(q, t, 0.64431879f);
  q = fmaf(q, t, 0.093436931f);
  q = fmaf(q, t, 0.23524893f);
  const float e = ex2_approx(v * v * -0.72134752f);                  // exp(-v^2 / 2)
  const float s = (0.5f * ax) * (q * t) * e;
  return fmaxf(v, 0.f) - s;
}

// The same on two values with packed arithmetic (same operations in the same order per value: bit-identical results):
// 2 x (abs, max, 2 MUFU) + 15 packed instructions per pair instead of 2 x 15.
__device__ __forceinline__ uint64_t gelu_erf2(uint64_t v2) {
  float va, vb;
  unpack2f(v2, va, vb);
  const uint64_t ax2 = pack2f(fabsf(va), fabsf(vb));
  float ua, ub;
  unpack2f(fma2f(ax2, dup2f(0.27599917f), dup2f(1.f)), ua, ub);
  const uint64_t t2 = pack2f(rcp_approx(ua), rcp_approx(ub));
  uint64_t q2 = fma2f(dup2f(-0.22690525f), t2, dup2f(0.88164259f));
  q2 = fma2f(q2, t2, dup2f(-0.62774200f));
  q2 = fma2f(q2, t2, dup2f(0.64431879f));
  q2 = fma2f(q2, t2, dup2f(0.093436931f));
  q2 = fma2f(q2, t2, dup2f(0.23524893f));
  float ea, eb;
  unpack2f(mul2f(mul2f(v2, v2), dup2f(-0.72134752f)), ea, eb);
  const uint64_t e2 = pack2f(ex2_approx(ea), ex2_approx(eb));
  const uint64_t s2 = mul2f(mul2f(mul2f(dup2f(0.5f), ax2), mul2f(q2, t2)), e2);
  return fma2f(s2, dup2f(-1.f), pack2f(fmaxf(va, 0.f), fmaxf(vb, 0.f)));
}

__global__ void __launch_bounds__(kThreadsTC, 1) psf_gen_fwd_tc_kernel(const GenParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + kOffBar;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(smem + kOffBar + BAR_COUNT * 8);
  const int n_gen = p.n_gen;
  const int nch = (n_gen + 3) >> 2;

  // ---- one-time setup: weights (hi, lo) into the canonical K-major layout, biases, barriers, tensor memory ----
  for (int idx = tid; idx < nch * 4 * GH * GE; idx += kThreadsTC) {
    const int m = idx >> 10, h = (idx >> 5) & 31, e = idx & 31;
    float v = 0.f;
    if (m < n_gen) {
      const float* W1 = m < p.n_W ? p.W1 + (long long)m * GH * GE : p.gW1;
      v = W1[h * GE + e];
    }
    const int r = (m & 3) * 32 + h;
    const uint32_t off = (uint32_t)(m >> 2) * kW1Chunk + (r >> 3) * 1024 + (e >> 2) * 128 + (r & 7) * 16 + (e & 3) * 4;
    *reinterpret_cast<float*>(smem + off) = tf32_hi(v);
    *reinterpret_cast<float*>(smem + off + kW1Chunk / 2) = tf32_lo(v);
  }
  for (int idx = tid; idx < n_gen * GLP * GH; idx += kThreadsTC) {
    const int m = idx >> 9, l = (idx >> 5) & 15, h = idx & 31;
    const float* W2 = m < p.n_W ? p.W2 + (long long)m * p.L * GH : p.gW2;
    const int Lm = m < p.n_W ? p.L : p.C;
    const float v = l < Lm ? W2[l * GH + h] : 0.f;
    const uint32_t off = kOffW2 + (uint32_t)m * kW2Gen + (l >> 3) * 1024 + (h >> 2) * 128 + (l & 7) * 16 + (h & 3) * 4;
    *reinterpret_cast<float*>(smem + off) = tf32_hi(v);
    *reinterpret_cast<float*>(smem + off + kW2Gen / 2) = tf32_lo(v);
  }
  for (int idx = tid; idx < n_gen * GH; idx += kThreadsTC) {
    const int m = idx >> 5;
    reinterpret_cast<float*>(smem + kOffB1)[idx] = (m < p.n_W ? p.b1 + m * GH : p.gb1)[idx & 31];
  }
  for (int idx = tid; idx < n_gen * GLP; idx += kThreadsTC) {
    const int m = idx >> 4, l = idx & 15;
    const int Lm = m < p.n_W ? p.L : p.C;
    reinterpret_cast<float*>(smem + kOffB2)[idx] = l < Lm ? (m < p.n_W ? p.b2 + m * p.L : p.gb2)[l] : 0.f;
  }
  if (tid == 0) {
    mbar_init(bars + 8 * BAR_XFULL, kLoadWarps);
    mbar_init(bars + 8 * BAR_XEMPTY, 1);
    mbar_init(bars + 8 * BAR_PREFULL0, 1);
    mbar_init(bars + 8 * BAR_PREFULL1, 1);
    mbar_init(bars + 8 * BAR_PREEMPTY0, kEpiWarps);
    mbar_init(bars + 8 * BAR_PREEMPTY1, kEpiWarps);
    mbar_init(bars + 8 * BAR_HIDFULL0, 4);
    mbar_init(bars + 8 * BAR_HIDFULL1, 4);
    mbar_init(bars + 8 * BAR_L2DONE0, 1);
    mbar_init(bars + 8 * BAR_L2DONE1, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy weight stores -> tensor-core reads
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tslot;

  const long long ntiles = (p.rows + 127) >> 7;
  const int my_tiles = (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

  if (warp < kEpiWarps) {
    // ================================ epilogue warpgroups ================================
    const int wg = warp >> 2;
    const uint32_t lane_base = ((uint32_t)(warp & 3) * 32u) << 16;
    const uint32_t t_hid = tmem + lane_base + kColHid + 64 * wg;
    const uint32_t t_out = tmem + lane_base + kColOut + 16 * wg;
    const uint32_t bar_hid = bars + 8 * (BAR_HIDFULL0 + wg), bar_l2 = bars + 8 * (BAR_L2DONE0 + wg);
    const float* b1s = reinterpret_cast<const float*>(smem + kOffB1);
    const float* b2s = reinterpret_cast<const float*>(smem + kOffB2);
    uint32_t kk = 0;            // generators this warpgroup has handed to layer 2
    bool pending = false;       // a generator whose layer-2 result is still to be stored
    int pend_m = 0;
    long long pend_row = 0;
    // layer-2 result of the pending generator: TMEM -> + b2 -> global (call once l2_done has been waited for)
    auto store_pending = [&](const float (&o)[16]) {
      if (pend_row >= p.rows) return;
      const float* b2 = b2s + pend_m * GLP;
      if (pend_m < p.n_W) {
        float* dst = p.Wt + (long long)pend_m * p.sWf + pend_row;
#pragma unroll
        for (int l = 0; l < GLP; ++l)
          if (l < p.L) dst[(long long)l * p.sWt] = tc_unbias(o[l]) + b2[l];
      } else {
        float* dst = p.V + natural_row(p, pend_row) * p.C;
#pragma unroll
        for (int c = 0; c < GLP; ++c)
          if (c < p.C) dst[c] = tc_unbias(o[c]) + b2[c];
      }
    };
    for (int it = 0; it < my_tiles; ++it) {
      const long long tile = blockIdx.x + (long long)it * gridDim.x;
      const long long row = tile * 128 + (warp & 3) * 32 + lane;
      for (int c = 0; c < nch; ++c) {
        const uint32_t q = (uint32_t)(it * nch + c), stage = q & 1;
        mbar_wait_tc(bars + 8 * (BAR_PREFULL0 + stage), (q >> 1) & 1);
        tc_fence_after();
        bool released = false;
#pragma unroll 1
        for (int jj = 0; jj < 2; ++jj) {
          const int j = 2 * jj + wg, m = c * 4 + j;
          if (m >= n_gen) break;
          float v[32];
          tmem_ld32f(tmem + lane_base + kColPre + stage * 128 + j * 32, v);
          tmem_wait_ld_tc();
          if (jj == 1 || m + 2 >= n_gen) {      // this warp's last read of the stage: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bars + 8 * (BAR_PREEMPTY0 + stage));
            released = true;
          }
          const float* b1 = b1s + m * GH;
#pragma unroll
          float vh[32];                                    // hidden units: (hi, lo) for the layer-2 A operand
          for (int h4 = 0; h4 < GH / 4; ++h4) {
            const float4 b = *reinterpret_cast<const float4*>(b1 + 4 * h4);
            const uint64_t one = dup2f(1.f);
            const uint64_t g0 = gelu_erf2(fma2f(pack2f(tc_unbias(v[4 * h4 + 0]), tc_unbias(v[4 * h4 + 1])), one, pack2f(b.x, b.y)));
            const uint64_t g1 = gelu_erf2(fma2f(pack2f(tc_unbias(v[4 * h4 + 2]), tc_unbias(v[4 * h4 + 3])), one, pack2f(b.z, b.w)));
            uint64_t hi, lo;
            tf32_split2(g0, hi, lo);
            unpack2f(hi, vh[4 * h4 + 0], vh[4 * h4 + 1]);
            unpack2f(lo, v[4 * h4 + 0], v[4 * h4 + 1]);
            tf32_split2(g1, hi, lo);
            unpack2f(hi, vh[4 * h4 + 2], vh[4 * h4 + 3]);
            unpack2f(lo, v[4 * h4 + 2], v[4 * h4 + 3]);
          }
          float o[16];
          if (pending) {     // the previous generator's layer 2 has finished: its result out, the hid columns free
            mbar_wait_tc(bar_l2, (kk - 1) & 1);
            tc_fence_after();
            tmem_ld16f(t_out, o);
            tmem_wait_ld_tc();
          }
          tmem_st32f(t_hid, vh);
          tmem_st32f(t_hid + 32, v);
          tmem_wait_st_tc();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_hid);
          if (pending) store_pending(o);
          pending = true;
          pend_m = m;
          pend_row = row;
          ++kk;
        }
        if (!released) {   // no generator of this chunk for this warpgroup
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bars + 8 * (BAR_PREEMPTY0 + stage));
        }
      }
    }
    if (pending) {
      float o[16];
      mbar_wait_tc(bar_l2, (kk - 1) & 1);
      tc_fence_after();
      tmem_ld16f(t_out, o);
      tmem_wait_ld_tc();
      store_pending(o);
    }
  } else if (warp < kMmaWarp) {
    // ================================ x loader ================================
    const uint32_t lane_base = ((uint32_t)(warp & 3) * 32u) << 16;
    for (int it = 0; it < my_tiles; ++it) {
      const long long tile = blockIdx.x + (long long)it * gridDim.x;
      const long long row = tile * 128 + (warp & 3) * 32 + lane;
      float v[32];
      if (row < p.rows) {
        const float4* src = reinterpret_cast<const float4*>(p.x + row * GE);
#pragma unroll
        for (int i = 0; i < GE / 4; ++i) {
          const float4 t = __ldg(src + i);
          v[4 * i] = t.x;
          v[4 * i + 1] = t.y;
          v[4 * i + 2] = t.z;
          v[4 * i + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < GE; ++i) v[i] = 0.f;
      }
      mbar_wait_relaxed(bars + 8 * BAR_XEMPTY, (it & 1) ^ 1);   // layer 1 of the previous tile has read x
      tc_fence_after();
      {
        float vh[32];
#pragma unroll
        for (int i = 0; i < GE; ++i) vh[i] = tf32_hi(v[i]);
        tmem_st32f(tmem + lane_base + kColXhi, vh);
#pragma unroll
        for (int i = 0; i < GE; ++i) v[i] -= vh[i];
        tmem_st32f(tmem + lane_base + kColXlo, v);
      }
      tmem_wait_st_tc();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 8 * BAR_XFULL);
    }
  } else {
    // ================================ MMA issuer ================================
    if (elect_one()) {
      constexpr uint32_t idesc1 = make_idesc(128, 128), idesc2 = make_idesc(128, 16);
      const uint32_t Q = (uint32_t)(my_tiles * nch);
      auto issue_pre = [&](uint32_t q) {
        const uint32_t c = q % (uint32_t)nch, it = q / (uint32_t)nch, stage = q & 1;
        if (c == 0) mbar_wait_hot(bars + 8 * BAR_XFULL, it & 1);
        mbar_wait_hot(bars + 8 * (BAR_PREEMPTY0 + stage), ((q >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d = tmem + kColPre + stage * 128;
        const uint32_t w1 = sbase + c * kW1Chunk;
#pragma unroll
        for (int pass = 0; pass < 3; ++pass) {     // x_lo W1_hi, x_hi W1_lo, x_hi W1_hi
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint32_t a = tmem + (pass == 0 ? kColXlo : kColXhi) + ks * 8;
            const uint64_t b = make_desc(w1 + (pass == 1 ? kW1Chunk / 2 : 0) + ks * 256, 128, 1024);
            mma_ts(d, a, b, idesc1, (pass | ks) != 0);
          }
        }
        mma_commit(bars + 8 * (BAR_PREFULL0 + stage));
        if (c == (uint32_t)nch - 1) mma_commit(bars + 8 * BAR_XEMPTY);
      };
      uint32_t k0 = 0, k1 = 0;
      if (Q > 0) issue_pre(0);
      if (Q > 1) issue_pre(1);
      for (uint32_t q = 0; q < Q; ++q) {
        const uint32_t c = q % (uint32_t)nch;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          const int m = (int)c * 4 + j;
          if (m >= n_gen) break;
          const int wg = j & 1;
          mbar_wait_hot(bars + 8 * (BAR_HIDFULL0 + wg), (wg ? k1 : k0) & 1);
          if (wg) ++k1; else ++k0;
          tc_fence_after();
          const uint32_t d = tmem + kColOut + 16 * wg;
          const uint32_t hid = tmem + kColHid + 64 * wg;
          const uint32_t w2 = sbase + kOffW2 + (uint32_t)m * kW2Gen;
#pragma unroll
          for (int pass = 0; pass < 3; ++pass) {   // hid_lo W2_hi, hid_hi W2_lo, hid_hi W2_hi
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint32_t a = hid + (pass == 0 ? 32 : 0) + ks * 8;
              const uint64_t b = make_desc(w2 + (pass == 1 ? kW2Gen / 2 : 0) + ks * 256, 128, 1024);
              mma_ts(d, a, b, idesc2, (pass | ks) != 0);
            }
          }
          mma_commit(bars + 8 * (BAR_L2DONE0 + wg));
        }
        if (q + 2 < Q) issue_pre(q + 2);
      }
    }
    __syncwarp();
  }

  // ---- teardown: every tcgen05 operation has been waited for by its consumer ----
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

struct DevAttr {
  bool done = false;
  int status = 0;
};
DevAttr g_tc_attr[PSF_MAX_DEVICES];   // per device ordinal

}  // namespace

int psf_gen_tc_fwd(const GenParams& p, cudaStream_t st) {
  if (p.n_gen < 1 || p.n_gen > kMaxGenTC || p.L > GLP || p.C > GLP || p.pre != nullptr) return PSF_ERR_UNSUPPORTED;
  const int dev = psf_current_device();
  if (dev < 0) return PSF_ERR_UNSUPPORTED;
  {
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (!g_tc_attr[dev].done) {
    cudaError_t e = cudaFuncSetAttribute(psf_gen_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemTC);
    g_tc_attr[dev].status = (int)e;
    g_tc_attr[dev].done = true;
    if (e != cudaSuccess) cudaGetLastError();
    }
  }
  if (g_tc_attr[dev].status != 0) return PSF_ERR_UNSUPPORTED;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long ntiles = (p.rows + 127) >> 7;
  const long long grid = ntiles < sms ? ntiles : sms;
  psf_gen_fwd_tc_kernel<<<(unsigned)grid, kThreadsTC, kSmemTC, st>>>(p);
  return (int)cudaGetLastError();
}
