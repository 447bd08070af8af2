// Fused W generators for sm_100a: every fs[m] = Linear(E,H) -> GELU -> Linear(H,L) of PSFNet
// (SyntheticExperiments/psf.py:110-126, evaluated at psf.py:175) in ONE kernel each way, fp32 FMA only
// (no TF32: the chain's parity tolerance is rtol 1e-5), E = H = 32, L <= 16.
//
//   forward : x [rows, 32] -> W [n_W][L][rows]  (tap-major: the chain's streaming layout; `rows` = B*n in
//             whatever row order x is given in, so a quad-interleaved x yields a quad-interleaved W)
//             thread = one row: x stays in registers for all n_W generators, weights are broadcast
//             reads from shared memory, hidden activations never leave registers.  HBM traffic is the
//             compulsory 128 B in + 4*L*n_W B out per row (the cuBLAS route wrote and re-read three
//             [rows, n_W*32] buffers).
//   backward: dW [n_W][L][rows] -> dx [rows,32], dW1, db1, dW2, db2.  The pre-activations come from the
//             forward pass when it was asked to save them ([n_gen][32][rows], 4.2 GB at the headline shape,
//             written and read with coalesced streaming accesses behind compute-bound kernels) or are
//             RECOMPUTED from x (1 KFMA + 256 broadcast quads per row and generator).  Per 128-row tile and generator (128 threads, three CTAs per SM so that
//             the phases of different tiles overlap):
//               phase 1 (thread = row): pre, gelu, gelu', dhid = dout W2, dpre, dx += dpre W1;
//                        dpre / hid / dout staged in shared memory,
//               phase 2 (thread = 8x8 output tile x one row group): dW1 += dpre^T x, dW2 += dout^T hid
//                        as register-tiled outer products over the tile's rows, reduced across row
//                        groups with a shuffle reduce-scatter, accumulated into a per-CTA slice of a
//                        global workspace (fire-and-forget red.add by the one owning thread: deterministic).
//             A second tiny kernel sums the per-CTA slices in a fixed order.
#include <stdlib.h>
#include "psf_gen.cuh"

namespace {

constexpr int kBlkFwd = GH * GE + GH + GH * GLP + GLP;  // w1[h][e], b1[h], w2t[h][l], b2[l]  = 1584 floats
constexpr int kBlkBwd = GH * GE + GH + GLP * GH;        // w1[h][e], b1[h], w2[l][h]          = 1568 floats
// phase 2 of the backward kernel (parameter-gradient GEMMs): 1 = tensor cores (mma.sync 3xTF32), 0 = FP32 FMA tiles
#ifndef PSF_GEN_MMA
#define PSF_GEN_MMA 1
#endif
// 1: the whole backward pass on tensor cores (psf_gen_bwd_mma_kernel); 0: phase 1 in FP32 FMA (psf_gen_bwd_kernel)
#ifndef PSF_GEN_BWD_ALL_MMA
#define PSF_GEN_BWD_ALL_MMA 1
#endif
constexpr bool kBwdAllMma = PSF_GEN_BWD_ALL_MMA != 0;
#ifndef PSF_GEN_P2_UNROLL
#define PSF_GEN_P2_UNROLL 2
#endif
constexpr int kP2Unroll = PSF_GEN_P2_UNROLL;  // unrolling of the phase-2 row loops of the backward kernel
constexpr int XS = GE + 4;                               // padded row stride of the staged tiles (bank-conflict free)
constexpr int DS = GLP + 4;

__device__ __forceinline__ void ffma2(float2& c, const float2& a, const float2& b) {
  unsigned long long* pc = reinterpret_cast<unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(*pc)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
}
__device__ __forceinline__ float4 lds4g(const float* p) { return *reinterpret_cast<const float4*>(p); }

// exact GELU as ATen evaluates it in fp32: x * 0.5 * (1 + erf(x / sqrt 2))
__device__ __forceinline__ float gelu_f(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752440f)); }

// Backward-only evaluation of Phi(x) and phi(x) (the standard normal cdf / pdf) for gelu(x) = x Phi(x) and
// gelu'(x) = Phi(x) + x phi(x): erf by Abramowitz-Stegun 7.1.26 shares its exp(-x^2/2) with phi; absolute error
// 3e-7 in fp32, which only reaches gradients (the forward kernel uses erff).  ~16 instructions instead of ~44.
__device__ __forceinline__ void normal_cdf_pdf_fast(float x, float& cdf, float& pdf) {
  const float e = __expf(-0.5f * x * x);
  const float t = __fdividef(1.f, fmaf(0.23164189f, fabsf(x), 1.f));  // 0.3275911 / sqrt 2
  float q = fmaf(1.061405429f, t, -1.453152027f);
  q = fmaf(q, t, 1.421413741f);
  q = fmaf(q, t, -0.284496736f);
  q = fmaf(q, t, 0.254829592f);
  const float half_erf = fmaf(-0.5f * q * t, e, 0.5f);  // 0.5 erf(|x| / sqrt 2)
  cdf = 0.5f + copysignf(half_erf, x);
  pdf = 0.39894228040143267794f * e;
}

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
               : "memory");
}
// fire-and-forget vector add to global memory (no read latency on the issuing thread).  Every accumulator element
// is only ever touched by one thread, in program order, so the sums stay deterministic.
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add1(float* p, float a) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}
__device__ __forceinline__ void red_add2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
// v = hi + lo with hi, lo representable in TF32 (10-bit mantissa): the operand split of the 3xTF32 scheme
// (truncating split: hi = upper 19 bits, lo = v - hi exactly; the tensor core ignores the low 13 mantissa bits of lo)
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(v) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}
// D(16x8) += A(16x8, row) * B(8x8, col), TF32 inputs, fp32 accumulate (warp-level tensor-core instruction)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------
// forward: thread = TWO rows (rows r and r + kFT of a 2*kFT-row tile), so that every broadcast weight quad
// read from shared memory feeds 8 FMAs instead of 4 - with one row per thread the kernel sat at 90 % of the
// shared-memory pipe (a broadcast LDS.128 costs two wavefronts).
// ---------------------------------------------------------------------------------------------
constexpr int kFT = 192;  // threads per CTA; 2 CTAs per SM at <= 168 registers

__global__ void __launch_bounds__(kFT, 2) psf_gen_fwd_kernel(const GenParams p) {
  extern __shared__ __align__(16) float gsm[];
  // all generators' weights -> shared memory, once per (persistent) CTA
  for (int idx = threadIdx.x; idx < p.n_gen * kBlkFwd; idx += kFT) {
    const int m = idx / kBlkFwd;
    int o = idx - m * kBlkFwd;
    const float *W1, *b1, *W2, *b2;
    int Lm;
    gen_ptrs(p, m, W1, b1, W2, b2, Lm);
    float v;
    if (o < GH * GE) {
      v = W1[o];
    } else if ((o -= GH * GE) < GH) {
      v = b1[o];
    } else if ((o -= GH) < GH * GLP) {
      const int h = o / GLP, l = o - h * GLP;
      v = l < Lm ? W2[l * GH + h] : 0.f;
    } else {
      o -= GH * GLP;
      v = o < Lm ? b2[o] : 0.f;
    }
    gsm[idx] = v;
  }
  __syncthreads();

  const long long ntiles = (p.rows + 2 * kFT - 1) / (2 * kFT);
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    long long row[2];
    bool live[2];
    float2 x2[2][GE / 2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      row[r] = tile * (2 * kFT) + r * kFT + threadIdx.x;
      live[r] = row[r] < p.rows;
      const float4* xr = reinterpret_cast<const float4*>(p.x + (live[r] ? row[r] : 0) * GE);
#pragma unroll
      for (int i = 0; i < GE / 4; ++i) {
        const float4 v = __ldg(xr + i);
        x2[r][2 * i] = make_float2(v.x, v.y);
        x2[r][2 * i + 1] = make_float2(v.z, v.w);
      }
    }
    if (!live[0]) continue;
#pragma unroll 1
    for (int m = 0; m < p.n_gen; ++m) {
      const float* w1 = gsm + m * kBlkFwd;
      const float* b1 = w1 + GH * GE;
      const float* w2t = b1 + GH;
      const float* b2 = w2t + GH * GLP;
      float2 out2[2][GLP / 2];
#pragma unroll
      for (int i = 0; i < GLP / 4; ++i) {
        const float4 b = lds4g(b2 + 4 * i);
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          out2[r][2 * i] = make_float2(b.x, b.y);
          out2[r][2 * i + 1] = make_float2(b.z, b.w);
        }
      }
#pragma unroll 1
      for (int hc = 0; hc < GH / 8; ++hc) {
        float2 pre2[2][8];
#pragma unroll
        for (int h = 0; h < 8; ++h) pre2[0][h] = pre2[1][h] = make_float2(0.f, 0.f);
#pragma unroll
        for (int e4 = 0; e4 < GE / 4; ++e4) {
#pragma unroll
          for (int h = 0; h < 8; ++h) {
            const float4 w = lds4g(w1 + (hc * 8 + h) * GE + 4 * e4);
            const float2 wa = make_float2(w.x, w.y), wc = make_float2(w.z, w.w);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
              ffma2(pre2[r][h], wa, x2[r][2 * e4]);
              ffma2(pre2[r][h], wc, x2[r][2 * e4 + 1]);
            }
          }
        }
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const float bias = b1[hc * 8 + h];
          const float pre0 = pre2[0][h].x + pre2[0][h].y + bias, pre1 = pre2[1][h].x + pre2[1][h].y + bias;
          if (p.pre != nullptr) {  // coalesced: consecutive lanes are consecutive rows
            float* ps = p.pre + (long long)(m * GH + hc * 8 + h) * p.rows;
            __stcs(ps + row[0], pre0);
            if (live[1]) __stcs(ps + row[1], pre1);
          }
          const float hid0 = gelu_f(pre0);
          const float hid1 = gelu_f(pre1);
          const float2 hh0 = make_float2(hid0, hid0), hh1 = make_float2(hid1, hid1);
#pragma unroll
          for (int l4 = 0; l4 < GLP / 4; ++l4) {
            const float4 w = lds4g(w2t + (hc * 8 + h) * GLP + 4 * l4);
            const float2 wa = make_float2(w.x, w.y), wc = make_float2(w.z, w.w);
            ffma2(out2[0][2 * l4], wa, hh0);
            ffma2(out2[0][2 * l4 + 1], wc, hh0);
            ffma2(out2[1][2 * l4], wa, hh1);
            ffma2(out2[1][2 * l4 + 1], wc, hh1);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (!live[r]) continue;
        if (m < p.n_W) {
          float* dst = p.Wt + (long long)m * p.sWf + row[r];
#pragma unroll
          for (int l = 0; l < GLP; ++l)
            if (l < p.L) dst[(long long)l * p.sWt] = (l & 1) ? out2[r][l >> 1].y : out2[r][l >> 1].x;
        } else {  // V0 row, natural order, row-major
          float* dst = p.V + natural_row(p, row[r]) * p.C;
#pragma unroll
          for (int c = 0; c < GLP; ++c)
            if (c < p.C) dst[c] = (c & 1) ? out2[r][c >> 1].y : out2[r][c >> 1].x;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
// reduce-scatter step over lane bit M of N+N values: afterwards v[0..N) holds the pairwise sums of the half
// this lane keeps (lanes with the bit clear keep the first half)
template <int N>
__device__ __forceinline__ void rs_step2(float2* v, int M, bool hi) {  // same on N+N float2 values
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float2 send = hi ? v[i] : v[i + N];
    const float2 keep = hi ? v[i + N] : v[i];
    v[i].x = keep.x + __shfl_xor_sync(0xffffffffu, send.x, M);
    v[i].y = keep.y + __shfl_xor_sync(0xffffffffu, send.y, M);
  }
}
template <int N>
__device__ __forceinline__ void rs_step(float* v, int M, bool hi) {
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const float send = hi ? v[i] : v[i + N];
    const float keep = hi ? v[i + N] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, M);
  }
}

template <int kTB, bool kSaved>
__global__ void __launch_bounds__(kTB, 384 / kTB) psf_gen_bwd_kernel(const GenParams p) {
  static_assert(kTB == 64 || kTB == 128, "tile sizes with a phase-2 mapping");
  constexpr int KG1 = kTB / 16;  // row groups of phase 2a (16 output tiles x KG1 = kTB threads)
  constexpr int KG2 = kTB / 8;   // row groups of phase 2b (8 output tiles x KG2 = kTB threads)
  extern __shared__ __align__(16) float gsm[];
  float* xs = gsm;                    // [kTB][XS]   x rows of the tile
  float* dps = xs + kTB * XS;       // [kTB][XS]   dpre
  float* hs = dps + kTB * XS;       // [kTB][XS]   hidden activations
  float* dos = hs + kTB * XS;       // [kTB][DS]   dout (taps padded with zeros)
  float* wb0 = dos + kTB * DS;      // weights of the current generator (filled by cp.async during the previous phase 2)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* wsc = p.ws + (long long)blockIdx.x * p.n_gen * kAccBlk;
  const long long ntiles = (p.rows + kTB - 1) / kTB;
  // request generator m's weights: W1 1024 + b1 32 + W2 Lm*32 floats in 16 B pieces; the
  // padding rows of W2 (taps >= Lm) are zeroed with plain stores
  auto request_weights = [&](int m) {
    const float *W1, *b1g, *W2, *b2g;
    int Lm;
    gen_ptrs(p, m, W1, b1g, W2, b2g, Lm);
    float* dst = wb0;
    const int n16 = (GH * GE + GH + Lm * GH) / 4;
    for (int c = tid; c < n16; c += kTB) {
      const int o = 4 * c;
      const float* src = o < GH * GE ? W1 + o : (o < GH * GE + GH ? b1g + (o - GH * GE) : W2 + (o - GH * GE - GH));
      cp_async16(dst + o, src);
    }
    for (int o = GH * GE + GH + Lm * GH + tid; o < kBlkBwd; o += kTB) dst[o] = 0.f;
  };

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row = tile * kTB + tid;
    const bool live = row < p.rows;
    float2 dx2[GE / 2];
    float4 xin[GE / 4];
#pragma unroll
    for (int i = 0; i < GE / 4; ++i) {
      xin[i] = live ? __ldg(reinterpret_cast<const float4*>(p.x + row * GE) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      dx2[2 * i] = make_float2(0.f, 0.f);
      dx2[2 * i + 1] = make_float2(0.f, 0.f);
    }
    __syncthreads();  // phase 2 of the previous tile has finished reading xs and the weight buffers
#pragma unroll
    for (int i = 0; i < GE / 4; ++i) *reinterpret_cast<float4*>(xs + tid * XS + 4 * i) = xin[i];
    request_weights(0);
    cp_async_commit();

    const long long nat = (live && p.n_gen > p.n_W) ? natural_row(p, row) : 0;
    // incoming gradient of generator mm for this thread's row (streamed from HBM: fetched one generator ahead,
    // behind phase 2, so that its latency is not on the tile's critical path)
    auto load_dout = [&](int mm, float (&d)[GLP]) {
      if (mm < p.n_W) {
        const float* src = p.Wt + (long long)mm * p.sWf + row;
#pragma unroll
        for (int l = 0; l < GLP; ++l) d[l] = (live && l < p.L) ? __ldcs(src + (long long)l * p.sWt) : 0.f;
      } else {
        const float* src = p.V + nat * p.C;
#pragma unroll
        for (int c = 0; c < GLP; ++c) d[c] = (live && c < p.C) ? __ldcs(src + c) : 0.f;
      }
    };
    float dnext[GLP];
    load_dout(0, dnext);
#pragma unroll 1
    for (int m = 0; m < p.n_gen; ++m) {
      __syncthreads();  // phase 2 of the previous generator is done with dps / hs / dos
      const float* w1 = wb0;
      const float* b1 = w1 + GH * GE;
      const float* w2 = b1 + GH;                       // [GLP][GH], rows >= Lm are zero
      float dout[GLP];
      {
#pragma unroll
        for (int l = 0; l < GLP; ++l) dout[l] = dnext[l];
#pragma unroll
        for (int i = 0; i < GLP / 4; ++i)
          *reinterpret_cast<float4*>(dos + tid * DS + 4 * i) = make_float4(dout[4 * i], dout[4 * i + 1], dout[4 * i + 2], dout[4 * i + 3]);
      }
      cp_async_wait<0>();  // this generator's weights (requested one phase 2 ago) have landed
      __syncthreads();

      // ---- phase 1: this thread's row (x comes back from the staged tile: it is not kept in registers
      //      across phase 2, whose 8 x 8 accumulator tile needs them) ------------------------------
      float2 x2[GE / 2];
      if constexpr (!kSaved) {
#pragma unroll
        for (int i = 0; i < GE / 4; ++i) {
          const float4 v = lds4g(xs + tid * XS + 4 * i);
          x2[2 * i] = make_float2(v.x, v.y);
          x2[2 * i + 1] = make_float2(v.z, v.w);
        }
      }
#pragma unroll 1
      for (int hc = 0; hc < GH / 8; ++hc) {
        float2 pre2[8];
        float prs[8];
        if constexpr (kSaved) {  // pre-activations saved by the forward pass: coalesced loads, used after dhid
          const float* ps = p.pre + (long long)(m * GH + hc * 8) * p.rows + (live ? row : 0);
#pragma unroll
          for (int h = 0; h < 8; ++h) prs[h] = __ldcs(ps + (long long)h * p.rows);
        } else {
#pragma unroll
          for (int h = 0; h < 8; ++h) pre2[h] = make_float2(0.f, 0.f);
#pragma unroll
          for (int e4 = 0; e4 < GE / 4; ++e4) {
#pragma unroll
            for (int h = 0; h < 8; ++h) {
              const float4 w = lds4g(w1 + (hc * 8 + h) * GE + 4 * e4);
              ffma2(pre2[h], make_float2(w.x, w.y), x2[2 * e4]);
              ffma2(pre2[h], make_float2(w.z, w.w), x2[2 * e4 + 1]);
            }
          }
        }
        float2 dh2[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) dh2[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int l = 0; l < GLP; ++l) {
          const float2 dd = make_float2(dout[l], dout[l]);
          const float4 wa = lds4g(w2 + l * GH + hc * 8);
          const float4 wc = lds4g(w2 + l * GH + hc * 8 + 4);
          ffma2(dh2[0], make_float2(wa.x, wa.y), dd);
          ffma2(dh2[1], make_float2(wa.z, wa.w), dd);
          ffma2(dh2[2], make_float2(wc.x, wc.y), dd);
          ffma2(dh2[3], make_float2(wc.z, wc.w), dd);
        }
        float hid[8], dpre[8];
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const float pre = kSaved ? prs[h] : pre2[h].x + pre2[h].y + b1[hc * 8 + h];
          float cdf, pdf;
          normal_cdf_pdf_fast(pre, cdf, pdf);
          hid[h] = pre * cdf;
          const float dhid = (h & 1) ? dh2[h >> 1].y : dh2[h >> 1].x;
          dpre[h] = dhid * (cdf + pre * pdf);
        }
#pragma unroll
        for (int h = 0; h < 8; ++h) {
          const float2 dd = make_float2(dpre[h], dpre[h]);
#pragma unroll
          for (int e4 = 0; e4 < GE / 4; ++e4) {
            const float4 w = lds4g(w1 + (hc * 8 + h) * GE + 4 * e4);
            ffma2(dx2[2 * e4], make_float2(w.x, w.y), dd);
            ffma2(dx2[2 * e4 + 1], make_float2(w.z, w.w), dd);
          }
        }
        *reinterpret_cast<float4*>(dps + tid * XS + hc * 8) = make_float4(dpre[0], dpre[1], dpre[2], dpre[3]);
        *reinterpret_cast<float4*>(dps + tid * XS + hc * 8 + 4) = make_float4(dpre[4], dpre[5], dpre[6], dpre[7]);
        *reinterpret_cast<float4*>(hs + tid * XS + hc * 8) = make_float4(hid[0], hid[1], hid[2], hid[3]);
        *reinterpret_cast<float4*>(hs + tid * XS + hc * 8 + 4) = make_float4(hid[4], hid[5], hid[6], hid[7]);
      }
      __syncthreads();
      if (m + 1 < p.n_gen) {   // phase 2 reads neither the weights nor dout: the next generator's land meanwhile
        request_weights(m + 1);
        load_dout(m + 1, dnext);
      }
      cp_async_commit();

      float* acc_m = wsc + (long long)m * kAccBlk;
#if PSF_GEN_MMA
      // ---- phase 2 on the tensor cores: the two parameter-gradient GEMMs (K = rows) as warp-level
      //      mma.sync.m16n8k8 TF32 with error compensation (3xTF32: a = a_hi + a_lo, a b ~ a_lo b_hi + a_hi b_lo +
      //      a_hi b_hi, fp32 accumulate; relative error ~1e-6, gradients only).  Every warp multiplies the 32 rows it
      //      staged itself; the warps' partial 32x32 / 16x32 results are summed through shared memory (fixed order)
      //      and added to the CTA's accumulator slice with one red.add per four elements. ------------------------
      {
        const int g = lane >> 2, t = lane & 3;
        const int r0 = warp * 32;
        // dW1[h][e] += dpre^T x:  M = h (2 tiles of 16), N = e (4 tiles of 8), K = this warp's rows (4 steps of 8)
        float d1[2][4][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) d1[i][j][k] = 0.f;
        // dW2[l][h] += dout^T hid:  M = l (16), N = h (4 tiles of 8)
        float d2[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int k = 0; k < 4; ++k) d2[j][k] = 0.f;
#pragma unroll 1
        for (int ks = 0; ks < 4; ++ks) {
          const int ra = (r0 + ks * 8 + t) * XS, rb = ra + 4 * XS;          // rows k = t and k = t + 4 of this step
          uint32_t bh[4][2], bl[4][2];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {                                    // B = x[k][e]
            split_tf32(xs[ra + nt * 8 + g], bh[nt][0], bl[nt][0]);
            split_tf32(xs[rb + nt * 8 + g], bh[nt][1], bl[nt][1]);
          }
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {                                    // A = dpre[k][h]^T
            uint32_t ah[4], al[4];
            split_tf32(dps[ra + mt * 16 + g], ah[0], al[0]);
            split_tf32(dps[ra + mt * 16 + g + 8], ah[1], al[1]);
            split_tf32(dps[rb + mt * 16 + g], ah[2], al[2]);
            split_tf32(dps[rb + mt * 16 + g + 8], ah[3], al[3]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              mma_tf32(d1[mt][nt], al, bh[nt]);
              mma_tf32(d1[mt][nt], ah, bl[nt]);
              mma_tf32(d1[mt][nt], ah, bh[nt]);
            }
          }
          const int da = (r0 + ks * 8 + t) * DS, db_ = da + 4 * DS;
          uint32_t ah[4], al[4];                                              // A = dout[k][l]^T
          split_tf32(dos[da + g], ah[0], al[0]);
          split_tf32(dos[da + g + 8], ah[1], al[1]);
          split_tf32(dos[db_ + g], ah[2], al[2]);
          split_tf32(dos[db_ + g + 8], ah[3], al[3]);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {                                    // B = hid[k][h]
            uint32_t hh[2], hl[2];
            split_tf32(hs[ra + nt * 8 + g], hh[0], hl[0]);
            split_tf32(hs[rb + nt * 8 + g], hh[1], hl[1]);
            mma_tf32(d2[nt], al, hh);
            mma_tf32(d2[nt], ah, hl);
            mma_tf32(d2[nt], ah, hh);
          }
        }
        // bias gradients: column sums over this warp's rows (lane = column)
        float sb1 = 0.f, sb2 = 0.f;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
          sb1 += dps[(r0 + r) * XS + lane];
          sb2 += dos[(r0 + r) * DS + (lane & (GLP - 1))];
        }
        red_add1(acc_m + GH * GE + lane, sb1);
        if (lane < GLP) red_add1(acc_m + GH * GE + GH + GLP * GH + lane, sb2);
        // cross-warp reduction through shared memory (the dpre / hid tiles are dead once every warp is here):
        // partial[warp][48 rows: 32 of dW1, 16 of dW2][PS], then one red.add.v4 per 4 elements of the CTA's sums
        constexpr int PS = 40;                       // padded row: the fragment stores are at most 2-way conflicted
        float* part = dps;                           // dps and hs are contiguous: 2 * kTB * XS >= (kTB / 32) * 48 * PS
        static_assert(2 * kTB * XS >= (kTB / 32) * 48 * PS, "partial sums do not fit the dead tiles");
        __syncthreads();
        float* mine = part + warp * 48 * PS;
        // accumulator fragments: c0 (g, 2t), c1 (g, 2t+1), c2 (g+8, 2t), c3 (g+8, 2t+1)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            *reinterpret_cast<float2*>(mine + (mt * 16 + g) * PS + nt * 8 + 2 * t) = make_float2(d1[mt][nt][0], d1[mt][nt][1]);
            *reinterpret_cast<float2*>(mine + (mt * 16 + g + 8) * PS + nt * 8 + 2 * t) = make_float2(d1[mt][nt][2], d1[mt][nt][3]);
          }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          *reinterpret_cast<float2*>(mine + (32 + g) * PS + nt * 8 + 2 * t) = make_float2(d2[nt][0], d2[nt][1]);
          *reinterpret_cast<float2*>(mine + (32 + g + 8) * PS + nt * 8 + 2 * t) = make_float2(d2[nt][2], d2[nt][3]);
        }
        __syncthreads();
        for (int u = tid; u < 48 * 8; u += kTB) {    // float4 units: row u / 8, columns 4 (u % 8) ..
          const int row = u >> 3, col = (u & 7) * 4;
          float4 sum = lds4g(part + row * PS + col);
#pragma unroll
          for (int w = 1; w < kTB / 32; ++w) {
            const float4 v = lds4g(part + (w * 48 + row) * PS + col);
            sum.x += v.x;
            sum.y += v.y;
            sum.z += v.z;
            sum.w += v.w;
          }
          // rows 0..31 are dW1[h][e]; rows 32..47 are dW2[l][h], stored after dW1 and db1 in the accumulator block
          float* dst = row < 32 ? acc_m + row * GE + col : acc_m + GH * GE + GH + (row - 32) * GH + col;
          red_add4(dst, sum.x, sum.y, sum.z, sum.w);
        }
      }
#else
      // ---- phase 2a: dW1[h][e] += sum_rows dpre[row][h] * x[row][e]  (+ db1) -----------------
      {
        const int kg = lane & (KG1 - 1);                            // row group: rows rr*KG1 + kg
        const int t16 = lane / KG1 + warp * (32 / KG1);             // 16 output tiles of 8 x 8
        const int hb = t16 >> 2, eb = t16 & 3;
        float2 v[32];  // v[i * 4 + j] = outputs (i, 2j), (i, 2j + 1) of the 8 x 8 tile
        float db[8];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) db[i] = 0.f;
#pragma unroll kP2Unroll
        for (int rr = 0; rr < kTB / KG1; ++rr) {
          const int r = rr * KG1 + kg;
          const float4 d0 = lds4g(dps + r * XS + hb * 8), d1 = lds4g(dps + r * XS + hb * 8 + 4);
          const float4 a0 = lds4g(xs + r * XS + eb * 8), a1 = lds4g(xs + r * XS + eb * 8 + 4);
          const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
          const float2 aa[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y),
                                make_float2(a1.z, a1.w)};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 di = make_float2(dd[i], dd[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) ffma2(v[i * 4 + j], aa[j], di);
            db[i] += dd[i];
          }
        }
        if constexpr (KG1 == 8) {
          rs_step2<16>(v, 4, lane & 4);
          rs_step2<8>(v, 2, lane & 2);
          rs_step2<4>(v, 1, lane & 1);
          const int i = ((lane & 4) ? 4 : 0) + ((lane & 2) ? 2 : 0) + (lane & 1);   // tile row this lane ends up owning
          float* dst = acc_m + (hb * 8 + i) * GE + eb * 8;
          red_add4(dst, v[0].x, v[0].y, v[1].x, v[1].y);
          red_add4(dst + 4, v[2].x, v[2].y, v[3].x, v[3].y);
          // bias gradient: every lane group holds the column sums of dpre; the eb == 0 tiles write them
          rs_step<4>(db, 4, lane & 4);
          rs_step<2>(db, 2, lane & 2);
          rs_step<1>(db, 1, lane & 1);
          if (eb == 0) red_add1(acc_m + GH * GE + hb * 8 + i, db[0]);
        } else {  // four row groups: two reduce-scatter stages leave two tile rows per lane
          rs_step2<16>(v, 2, lane & 2);
          rs_step2<8>(v, 1, lane & 1);
          const int i = ((lane & 2) ? 4 : 0) + ((lane & 1) ? 2 : 0);
          float* dst = acc_m + (hb * 8 + i) * GE + eb * 8;
          red_add4(dst, v[0].x, v[0].y, v[1].x, v[1].y);
          red_add4(dst + 4, v[2].x, v[2].y, v[3].x, v[3].y);
          red_add4(dst + GE, v[4].x, v[4].y, v[5].x, v[5].y);
          red_add4(dst + GE + 4, v[6].x, v[6].y, v[7].x, v[7].y);
          rs_step<4>(db, 2, lane & 2);
          rs_step<2>(db, 1, lane & 1);
          if (eb == 0) {
            red_add1(acc_m + GH * GE + hb * 8 + i, db[0]);
            red_add1(acc_m + GH * GE + hb * 8 + i + 1, db[1]);
          }
        }
      }
      // ---- phase 2b: dW2[l][h] += sum_rows dout[row][l] * hid[row][h]  (+ db2) ----------------
      {
        const int kg = lane & (KG2 - 1);                            // row group: rows rr*KG2 + kg
        const int t8 = lane / KG2 + warp * (32 / KG2);              // 8 output tiles of 8 x 8
        const int lb = t8 >> 2, hb = t8 & 3;
        float2 v[32];  // v[i * 4 + j] = outputs (i, 2j), (i, 2j + 1) of the 8 x 8 tile
        float db[8];
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) db[i] = 0.f;
#pragma unroll kP2Unroll
        for (int rr = 0; rr < kTB / KG2; ++rr) {
          const int r = rr * KG2 + kg;
          const float4 d0 = lds4g(dos + r * DS + lb * 8), d1 = lds4g(dos + r * DS + lb * 8 + 4);
          const float4 a0 = lds4g(hs + r * XS + hb * 8), a1 = lds4g(hs + r * XS + hb * 8 + 4);
          const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
          const float2 aa[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y),
                                make_float2(a1.z, a1.w)};
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 di = make_float2(dd[i], dd[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) ffma2(v[i * 4 + j], aa[j], di);
            db[i] += dd[i];
          }
        }
        if constexpr (KG2 == 16) {
          rs_step2<16>(v, 8, lane & 8);
          rs_step2<8>(v, 4, lane & 4);
          rs_step2<4>(v, 2, lane & 2);
          rs_step2<2>(v, 1, lane & 1);
          const int o = ((lane & 8) ? 32 : 0) + ((lane & 4) ? 16 : 0) + ((lane & 2) ? 8 : 0) + ((lane & 1) ? 4 : 0);
          red_add4(acc_m + GH * GE + GH + (lb * 8 + (o >> 3)) * GH + hb * 8 + (o & 7), v[0].x, v[0].y, v[1].x, v[1].y);
          rs_step<4>(db, 8, lane & 8);
          rs_step<2>(db, 4, lane & 4);
          rs_step<1>(db, 2, lane & 2);
          db[0] += __shfl_xor_sync(0xffffffffu, db[0], 1);
          if (hb == 0 && !(lane & 1)) {
            const int l = lb * 8 + ((lane & 8) ? 4 : 0) + ((lane & 4) ? 2 : 0) + ((lane & 2) ? 1 : 0);
            red_add1(acc_m + GH * GE + GH + GLP * GH + l, db[0]);
          }
        } else {  // eight row groups: three stages leave one tile row per lane
          rs_step2<16>(v, 4, lane & 4);
          rs_step2<8>(v, 2, lane & 2);
          rs_step2<4>(v, 1, lane & 1);
          const int i = ((lane & 4) ? 4 : 0) + ((lane & 2) ? 2 : 0) + (lane & 1);
          float* dst = acc_m + GH * GE + GH + (lb * 8 + i) * GH + hb * 8;
          red_add4(dst, v[0].x, v[0].y, v[1].x, v[1].y);
          red_add4(dst + 4, v[2].x, v[2].y, v[3].x, v[3].y);
          rs_step<4>(db, 4, lane & 4);
          rs_step<2>(db, 2, lane & 2);
          rs_step<1>(db, 1, lane & 1);
          if (hb == 0) red_add1(acc_m + GH * GE + GH + GLP * GH + lb * 8 + i, db[0]);
        }
      }
#endif
    }
    if (live) {
      float4* dst = reinterpret_cast<float4*>(p.dx + row * GE);
#pragma unroll
      for (int i = 0; i < GE / 4; ++i) dst[i] = make_float4(dx2[2 * i].x, dx2[2 * i].y, dx2[2 * i + 1].x, dx2[2 * i + 1].y);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward, tensor-core version: every GEMM-shaped piece of the backward pass runs as warp-level
// mma.sync.m16n8k8 TF32 with 3xTF32 error compensation (gradients only; relative error ~1e-6):
//   pre  [rows x h] = x    [rows x e] W1^T      dhid [rows x h] = dout [rows x l] W2
//   dx   [rows x e] += dpre[rows x h] W1        dW1 += dpre^T x,  dW2 += dout^T hid   (K = rows)
// A warp owns 32 rows of the 128-row tile for all of them, so the only CTA-wide synchronisation is the weight
// staging and the cross-warp sum of the parameter-gradient partials.  pre and dhid come out in the same
// accumulator-fragment layout, so gelu / gelu' / dpre are elementwise on registers.
// ---------------------------------------------------------------------------------------------
constexpr int kMT = 128;                                  // rows per tile == threads per CTA
constexpr int kWP = GE + 4;                               // padded weight row (conflict-free B-fragment loads)
constexpr int kBlkMma = GH * kWP + GH + GLP * kWP;        // w1p[h][e], b1[h], w2p[l][h]
constexpr int kSmemBwdMma = (3 * kMT * XS + kMT * DS + kBlkMma) * 4;

__global__ void __launch_bounds__(kMT, 3) psf_gen_bwd_mma_kernel(const GenParams p) {
  extern __shared__ __align__(16) float gsm[];
  float* xs = gsm;                  // [kMT][XS]  x rows of the tile
  float* dps = xs + kMT * XS;       // [kMT][XS]  dpre
  float* hs = dps + kMT * XS;       // [kMT][XS]  hidden activations
  float* dos = hs + kMT * XS;       // [kMT][DS]  dout (taps padded with zeros)
  float* wb = dos + kMT * DS;       // weights of the current generator
  const float* w1 = wb;             // [GH][kWP]
  const float* b1 = wb + GH * kWP;  // [GH]
  const float* w2 = b1 + GH;        // [GLP][kWP], rows >= Lm are zero

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int r0 = warp * 32;
  float* wsc = p.ws + (long long)blockIdx.x * p.n_gen * kAccBlk;
  const long long ntiles = (p.rows + kMT - 1) / kMT;

  auto request_weights = [&](int m) {   // 16-byte cp.async pieces into padded rows
    const float *W1, *b1g, *W2, *b2g;
    int Lm;
    gen_ptrs(p, m, W1, b1g, W2, b2g, Lm);
    for (int c = tid; c < GH * 8; c += kMT) cp_async16(wb + (c >> 3) * kWP + 4 * (c & 7), W1 + 4 * c);
    if (tid < GH / 4) cp_async16(wb + GH * kWP + 4 * tid, b1g + 4 * tid);
    float* w2d = wb + GH * kWP + GH;
    for (int c = tid; c < GLP * 8; c += kMT) {
      const int l = c >> 3;
      if (l < Lm) cp_async16(w2d + l * kWP + 4 * (c & 7), W2 + 4 * c);
      else *reinterpret_cast<float4*>(w2d + l * kWP + 4 * (c & 7)) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long row = tile * kMT + tid;
    const bool live = row < p.rows;
    {
      float4 xin[GE / 4];
#pragma unroll
      for (int i = 0; i < GE / 4; ++i)
        xin[i] = live ? __ldg(reinterpret_cast<const float4*>(p.x + row * GE) + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      __syncthreads();  // the previous tile is completely done with the staged tiles and the weights
#pragma unroll
      for (int i = 0; i < GE / 4; ++i) *reinterpret_cast<float4*>(xs + tid * XS + 4 * i) = xin[i];
    }
    request_weights(0);
    cp_async_commit();
    const long long nat = (live && p.n_gen > p.n_W) ? natural_row(p, row) : 0;
    auto load_dout = [&](int mm, float (&d)[GLP]) {
      if (mm < p.n_W) {
        const float* src = p.Wt + (long long)mm * p.sWf + row;
#pragma unroll
        for (int l = 0; l < GLP; ++l) d[l] = (live && l < p.L) ? __ldcs(src + (long long)l * p.sWt) : 0.f;
      } else {
        const float* src = p.V + nat * p.C;
#pragma unroll
        for (int c = 0; c < GLP; ++c) d[c] = (live && c < p.C) ? __ldcs(src + c) : 0.f;
      }
    };
    float dnext[GLP];
    load_dout(0, dnext);
    float dxa[2][4][4];   // dx fragments of this warp's rows: [16-row tile][8-column tile][c0..c3]
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) dxa[i][j][k] = 0.f;

#pragma unroll 1
    for (int m = 0; m < p.n_gen; ++m) {
      __syncthreads();  // the reduction of the previous generator has finished reading the partials in dps / hs
#pragma unroll
      for (int i = 0; i < GLP / 4; ++i)
        *reinterpret_cast<float4*>(dos + tid * DS + 4 * i) =
            make_float4(dnext[4 * i], dnext[4 * i + 1], dnext[4 * i + 2], dnext[4 * i + 3]);
      if (m + 1 < p.n_gen) load_dout(m + 1, dnext);   // streamed from HBM: a whole generator ahead
      cp_async_wait<0>();
      __syncthreads();  // weights, dout (and x) are visible

      // ---- phase 1: pre = x W1^T, dhid = dout W2, gelu / gelu' / dpre elementwise, stage hid and dpre ----
#pragma unroll 1
      for (int mt = 0; mt < 2; ++mt) {
        const int ra = (r0 + 16 * mt + g) * XS, rb = ra + 8 * XS;
        uint32_t xh[4][4], xl[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          split_tf32(xs[ra + 8 * ks + t], xh[ks][0], xl[ks][0]);
          split_tf32(xs[rb + 8 * ks + t], xh[ks][1], xl[ks][1]);
          split_tf32(xs[ra + 8 * ks + t + 4], xh[ks][2], xl[ks][2]);
          split_tf32(xs[rb + 8 * ks + t + 4], xh[ks][3], xl[ks][3]);
        }
        const int da = (r0 + 16 * mt + g) * DS, db_ = da + 8 * DS;
        uint32_t oh[2][4], ol[2][4];
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          split_tf32(dos[da + 8 * ks + t], oh[ks][0], ol[ks][0]);
          split_tf32(dos[db_ + 8 * ks + t], oh[ks][1], ol[ks][1]);
          split_tf32(dos[da + 8 * ks + t + 4], oh[ks][2], ol[ks][2]);
          split_tf32(dos[db_ + 8 * ks + t + 4], oh[ks][3], ol[ks][3]);
        }
        // the four 8-column tiles are independent accumulators: their MMAs are interleaved so that the three
        // dependent products of one tile are four instructions apart (the warp has few others to hide behind)
        float pre[4][4], dh[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const float bia0 = b1[8 * nt + 2 * t], bia1 = b1[8 * nt + 2 * t + 1];
          pre[nt][0] = pre[nt][2] = bia0;
          pre[nt][1] = pre[nt][3] = bia1;
          dh[nt][0] = dh[nt][1] = dh[nt][2] = dh[nt][3] = 0.f;
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {   // B[k = e][n = h] = W1[h][e]
          uint32_t bh[4][2], bl[4][2];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            split_tf32(w1[(8 * nt + g) * kWP + 8 * ks + t], bh[nt][0], bl[nt][0]);
            split_tf32(w1[(8 * nt + g) * kWP + 8 * ks + t + 4], bh[nt][1], bl[nt][1]);
          }
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_tf32(pre[nt], xl[ks], bh[nt]);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_tf32(pre[nt], xh[ks], bl[nt]);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_tf32(pre[nt], xh[ks], bh[nt]);
        }
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {   // B[k = l][n = h] = W2[l][h]
          uint32_t bh[4][2], bl[4][2];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            split_tf32(w2[(8 * ks + t) * kWP + 8 * nt + g], bh[nt][0], bl[nt][0]);
            split_tf32(w2[(8 * ks + t + 4) * kWP + 8 * nt + g], bh[nt][1], bl[nt][1]);
          }
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_tf32(dh[nt], ol[ks], bh[nt]);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_tf32(dh[nt], oh[ks], bl[nt]);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_tf32(dh[nt], oh[ks], bh[nt]);
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          float hid[4], dpre[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float cdf, pdf;
            normal_cdf_pdf_fast(pre[nt][i], cdf, pdf);
            hid[i] = pre[nt][i] * cdf;
            dpre[i] = dh[nt][i] * (cdf + pre[nt][i] * pdf);
          }
          // fragment (c0, c1) = row g, columns 2t, 2t+1 of the 8-column tile; (c2, c3) = row g + 8
          *reinterpret_cast<float2*>(hs + ra + 8 * nt + 2 * t) = make_float2(hid[0], hid[1]);
          *reinterpret_cast<float2*>(hs + rb + 8 * nt + 2 * t) = make_float2(hid[2], hid[3]);
          *reinterpret_cast<float2*>(dps + ra + 8 * nt + 2 * t) = make_float2(dpre[0], dpre[1]);
          *reinterpret_cast<float2*>(dps + rb + 8 * nt + 2 * t) = make_float2(dpre[2], dpre[3]);
        }
      }
      __syncwarp();
      // ---- dx += dpre W1  (K = h) --------------------------------------------------------------
#pragma unroll 1
      for (int mt = 0; mt < 2; ++mt) {
        const int ra = (r0 + 16 * mt + g) * XS, rb = ra + 8 * XS;
        uint32_t ph[4][4], pl[4][4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          split_tf32(dps[ra + 8 * ks + t], ph[ks][0], pl[ks][0]);
          split_tf32(dps[rb + 8 * ks + t], ph[ks][1], pl[ks][1]);
          split_tf32(dps[ra + 8 * ks + t + 4], ph[ks][2], pl[ks][2]);
          split_tf32(dps[rb + 8 * ks + t + 4], ph[ks][3], pl[ks][3]);
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {   // B[k = h][n = e] = W1[h][e]
          uint32_t bh[4][2], bl[4][2];
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) {
            split_tf32(w1[(8 * ks + t) * kWP + 8 * nt + g], bh[nt][0], bl[nt][0]);
            split_tf32(w1[(8 * ks + t + 4) * kWP + 8 * nt + g], bh[nt][1], bl[nt][1]);
          }
          if (mt == 0) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma_tf32(dxa[0][nt], pl[ks], bh[nt]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma_tf32(dxa[0][nt], ph[ks], bl[nt]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma_tf32(dxa[0][nt], ph[ks], bh[nt]);
          } else {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma_tf32(dxa[1][nt], pl[ks], bh[nt]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma_tf32(dxa[1][nt], ph[ks], bl[nt]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) mma_tf32(dxa[1][nt], ph[ks], bh[nt]);
          }
        }
      }

      // ---- phase 2: dW1 += dpre^T x, dW2 += dout^T hid over this warp's rows (K = rows) ------------
      float* acc_m = wsc + (long long)m * kAccBlk;
      float d1[2][4][4];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int k = 0; k < 4; ++k) d1[i][j][k] = 0.f;
      float d2[4][4];
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int k = 0; k < 4; ++k) d2[j][k] = 0.f;
#pragma unroll 1
      for (int ks = 0; ks < 4; ++ks) {
        const int ra = (r0 + ks * 8 + t) * XS, rb = ra + 4 * XS;          // rows k = t and k = t + 4 of this step
        uint32_t bh[4][2], bl[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {                                    // B = x[k][e]
          split_tf32(xs[ra + nt * 8 + g], bh[nt][0], bl[nt][0]);
          split_tf32(xs[rb + nt * 8 + g], bh[nt][1], bl[nt][1]);
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {                                    // A = dpre[k][h]^T
          uint32_t ah[4], al[4];
          split_tf32(dps[ra + mt * 16 + g], ah[0], al[0]);
          split_tf32(dps[ra + mt * 16 + g + 8], ah[1], al[1]);
          split_tf32(dps[rb + mt * 16 + g], ah[2], al[2]);
          split_tf32(dps[rb + mt * 16 + g + 8], ah[3], al[3]);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_tf32(d1[mt][nt], al, bh[nt]);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_tf32(d1[mt][nt], ah, bl[nt]);
#pragma unroll
          for (int nt = 0; nt < 4; ++nt) mma_tf32(d1[mt][nt], ah, bh[nt]);
        }
        const int da = (r0 + ks * 8 + t) * DS, db_ = da + 4 * DS;
        uint32_t ah[4], al[4];                                              // A = dout[k][l]^T
        split_tf32(dos[da + g], ah[0], al[0]);
        split_tf32(dos[da + g + 8], ah[1], al[1]);
        split_tf32(dos[db_ + g], ah[2], al[2]);
        split_tf32(dos[db_ + g + 8], ah[3], al[3]);
        uint32_t hh[4][2], hl[4][2];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {                                    // B = hid[k][h]
          split_tf32(hs[ra + nt * 8 + g], hh[nt][0], hl[nt][0]);
          split_tf32(hs[rb + nt * 8 + g], hh[nt][1], hl[nt][1]);
        }
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_tf32(d2[nt], al, hh[nt]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_tf32(d2[nt], ah, hl[nt]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_tf32(d2[nt], ah, hh[nt]);
      }
      {  // bias gradients: column sums over this warp's rows (lane = column)
        float sb1 = 0.f, sb2 = 0.f;
#pragma unroll 8
        for (int r = 0; r < 32; ++r) {
          sb1 += dps[(r0 + r) * XS + lane];
          sb2 += dos[(r0 + r) * DS + (lane & (GLP - 1))];
        }
        red_add1(acc_m + GH * GE + lane, sb1);
        if (lane < GLP) red_add1(acc_m + GH * GE + GH + GLP * GH + lane, sb2);
      }
      // cross-warp sum of the partials through shared memory (dps / hs are dead once every warp is here)
      constexpr int PS = 40;
      float* part = dps;
      static_assert(2 * kMT * XS >= (kMT / 32) * 48 * PS, "partial sums do not fit the dead tiles");
      __syncthreads();
      if (m + 1 < p.n_gen) request_weights(m + 1);   // nobody reads the weights any more
      cp_async_commit();
      float* mine = part + warp * 48 * PS;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          *reinterpret_cast<float2*>(mine + (mt * 16 + g) * PS + nt * 8 + 2 * t) = make_float2(d1[mt][nt][0], d1[mt][nt][1]);
          *reinterpret_cast<float2*>(mine + (mt * 16 + g + 8) * PS + nt * 8 + 2 * t) = make_float2(d1[mt][nt][2], d1[mt][nt][3]);
        }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        *reinterpret_cast<float2*>(mine + (32 + g) * PS + nt * 8 + 2 * t) = make_float2(d2[nt][0], d2[nt][1]);
        *reinterpret_cast<float2*>(mine + (32 + g + 8) * PS + nt * 8 + 2 * t) = make_float2(d2[nt][2], d2[nt][3]);
      }
      __syncthreads();
      for (int u = tid; u < 48 * 8; u += kMT) {
        const int prow = u >> 3, col = (u & 7) * 4;
        float4 sum = lds4g(part + prow * PS + col);
#pragma unroll
        for (int w = 1; w < kMT / 32; ++w) {
          const float4 v = lds4g(part + (w * 48 + prow) * PS + col);
          sum.x += v.x;
          sum.y += v.y;
          sum.z += v.z;
          sum.w += v.w;
        }
        float* dst = prow < 32 ? acc_m + prow * GE + col : acc_m + GH * GE + GH + (prow - 32) * GH + col;
        red_add4(dst, sum.x, sum.y, sum.z, sum.w);
      }
    }
    // dx fragments -> global: rows r0 + 16 mt + g (+ 8), columns 8 nt + 2t, 2t + 1
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const long long ga = tile * kMT + r0 + 16 * mt + g, gb = ga + 8;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        if (ga < p.rows) *reinterpret_cast<float2*>(p.dx + ga * GE + 8 * nt + 2 * t) = make_float2(dxa[mt][nt][0], dxa[mt][nt][1]);
        if (gb < p.rows) *reinterpret_cast<float2*>(p.dx + gb * GE + 8 * nt + 2 * t) = make_float2(dxa[mt][nt][2], dxa[mt][nt][3]);
      }
    }
  }
}

// sum the per-CTA slices in a fixed order and scatter into the parameter-gradient tensors
struct GenGrads {
  float *W1, *b1, *W2, *b2;     // stacked [n_W, ...]
  float *gW1, *gb1, *gW2, *gb2; // the V generator (generator index n_W), may be null
};
__global__ void psf_gen_reduce_kernel(const float* __restrict__ ws, int nslices, int n_W, int n_gen, int L, int C,
                                      GenGrads g, int dw1_transposed = 0) {
  const int total = n_gen * kAccBlk;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int c = 0; c < nslices; ++c) s += ws[(long long)c * total + idx];
    const int m = idx / kAccBlk;
    int o = idx - m * kAccBlk;
    const bool isg = m >= n_W;
    float* dW1 = isg ? g.gW1 : g.W1 + (long long)m * GH * GE;
    float* db1 = isg ? g.gb1 : g.b1 + m * GH;
    float* dW2 = isg ? g.gW2 : g.W2 + (long long)m * L * GH;
    float* db2 = isg ? g.gb2 : g.b2 + m * L;
    const int Lm = isg ? C : L;
    if (o < GH * GE) {
      dW1[dw1_transposed ? (o & (GH - 1)) * GE + (o >> 5) : o] = s;   // slices hold [e][h] when transposed
    } else if ((o -= GH * GE) < GH) {
      db1[o] = s;
    } else if ((o -= GH) < GLP * GH) {
      if (o / GH < Lm) dW2[o] = s;
    } else {
      o -= GLP * GH;
      if (o < Lm) db2[o] = s;
    }
  }
}

std::mutex g_gen_mu;
bool g_gen_attr[PSF_MAX_DEVICES] = {};

// PSF_GEN_TC=0 in the environment keeps the FP32-FMA / mma.sync kernels (A/B timing, parity cross-checks)
bool gen_use_tc() {
  static const bool on = [] {
    const char* e = getenv("PSF_GEN_TC");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

// PSF_GEN_BWD=1 selects the first tcgen05 backward kernel (parameter gradients on FP32 FMA / mma.sync)
int gen_bwd_version() {
  static const int v = [] {
    const char* e = getenv("PSF_GEN_BWD");
    return e != nullptr ? atoi(e) : 2;
  }();
  return v;
}

int gen_grid(int per_sm) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms * per_sm;
}

#ifndef PSF_GEN_TB
#define PSF_GEN_TB 128
#endif
constexpr int kTB = PSF_GEN_TB;   // backward: rows per tile == threads per CTA; 384 / kTB CTAs per SM (tools/gen_bench.py)
constexpr int kSmemBwd = (3 * kTB * XS + kTB * DS + kBlkBwd) * 4;

int gen_check(int64_t rows, int n_W, int E, int H, int L) {
  if (E != GE || H != GH || L < 1 || L > GLP || n_W < 1 || rows < 1) return PSF_ERR_UNSUPPORTED;
  if (n_W * kBlkFwd * 4 > 110 * 1024) return PSF_ERR_UNSUPPORTED;  // forward keeps every generator's weights resident
  const int dev = psf_current_device();
  if (dev < 0) return PSF_ERR_UNSUPPORTED;
  std::lock_guard<std::mutex> lock(g_gen_mu);
  if (!g_gen_attr[dev]) {
    if (cudaFuncSetAttribute(psf_gen_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024) != cudaSuccess ||
        cudaFuncSetAttribute(psf_gen_bwd_kernel<kTB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBwd) != cudaSuccess ||
        cudaFuncSetAttribute(psf_gen_bwd_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBwdMma) != cudaSuccess ||
        cudaFuncSetAttribute(psf_gen_bwd_kernel<kTB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBwd) != cudaSuccess) {
      cudaGetLastError();
      return PSF_ERR_UNSUPPORTED;
    }
    g_gen_attr[dev] = true;
  }
  return PSF_OK;
}

}  // namespace

extern "C" {

static size_t gen_slices_bytes(int n_W) {
  return psf_round256((size_t)gen_grid(384 / kTB) * (size_t)(n_W + 1) * kAccBlk * sizeof(float));
}
size_t psf_generators_workspace_bytes(int n_W) {
  if (n_W < 1) return 0;
  // per-CTA parameter-gradient slices, then the weight images of the tcgen05 kernel
  return gen_slices_bytes(n_W) + psf_round256(psf_gen_tc_bwd_image_bytes(n_W + 1));
}

static int gen_fill(GenParams& p, const psf_gen_desc* d, const float* x, const psf_mlp_params& fs, const psf_mlp_params& g) {
  if (d == nullptr || x == nullptr || fs.W1 == nullptr || fs.b1 == nullptr || fs.W2 == nullptr || fs.b2 == nullptr)
    return PSF_ERR_ARG;
  const bool has_g = d->C > 0;
  if (has_g && (g.W1 == nullptr || g.b1 == nullptr || g.W2 == nullptr || g.b2 == nullptr)) return PSF_ERR_ARG;
  if (d->C < 0 || d->C > GLP) return PSF_ERR_UNSUPPORTED;
  const int Q = d->interleave < 1 ? 1 : d->interleave;
  if ((has_g || Q > 1) && (d->n < 1 || d->rows % d->n != 0 || (Q > 1 && d->n % (4 * (int64_t)Q) != 0))) return PSF_ERR_ARG;
  int rc = gen_check(d->rows < 1 ? 1 : d->rows, d->n_W + (has_g ? 1 : 0), d->E, d->H, d->L);
  if (rc) return rc;
  if (reinterpret_cast<uintptr_t>(x) & 15) return PSF_ERR_ARG;
  p.x = x;
  p.W1 = fs.W1;
  p.b1 = fs.b1;
  p.W2 = fs.W2;
  p.b2 = fs.b2;
  p.gW1 = g.W1;
  p.gb1 = g.b1;
  p.gW2 = g.W2;
  p.gb2 = g.b2;
  p.rows = d->rows;
  p.sWf = d->stride_factor;
  p.sWt = d->stride_tap;
  p.n_W = d->n_W;
  p.L = d->L;
  p.n_gen = d->n_W + (has_g ? 1 : 0);
  p.C = d->C;
  p.n = (int)(d->n < 1 ? 1 : d->n);
  p.Q = Q;
  return PSF_OK;
}

int psf_generators_fwd(const psf_gen_desc* d, const float* x, psf_mlp_params fs, psf_mlp_params g, float* W_out,
                       float* V_out, float* pre_save, psf_stream_t stream) {
  GenParams p = {};
  int rc = gen_fill(p, d, x, fs, g);
  if (rc) return rc;
  if (W_out == nullptr || (d->C > 0 && V_out == nullptr)) return PSF_ERR_ARG;
  if (d->rows == 0) return PSF_OK;
  p.Wt = W_out;
  p.V = V_out;
  p.pre = pre_save;
  if (gen_use_tc() && pre_save == nullptr) {   // tcgen05 / tensor-memory kernel (psf_gen_tc_sm100.cu)
    rc = psf_gen_tc_fwd(p, (cudaStream_t)stream);
    if (rc != PSF_ERR_UNSUPPORTED) return rc;
  }
  const long long ntiles = (d->rows + 2 * kFT - 1) / (2 * kFT);
  long long grid = gen_grid(2);
  if (grid > ntiles) grid = ntiles;
  psf_gen_fwd_kernel<<<(unsigned)grid, kFT, (size_t)p.n_gen * kBlkFwd * 4, (cudaStream_t)stream>>>(p);
  return (int)cudaGetLastError();
}

int psf_generators_bwd(const psf_gen_desc* d, const float* x, psf_mlp_params fs, psf_mlp_params g, const float* dW_in,
                       const float* dV_in, const float* pre_saved, float* dx, psf_mlp_grads dfs, psf_mlp_grads dg,
                       void* workspace, size_t workspace_bytes, psf_stream_t stream) {
  GenParams p = {};
  int rc = gen_fill(p, d, x, fs, g);
  if (rc) return rc;
  if (dW_in == nullptr || dx == nullptr || dfs.W1 == nullptr || dfs.b1 == nullptr || dfs.W2 == nullptr || dfs.b2 == nullptr)
    return PSF_ERR_ARG;
  if (d->C > 0 && (dV_in == nullptr || dg.W1 == nullptr || dg.b1 == nullptr || dg.W2 == nullptr || dg.b2 == nullptr))
    return PSF_ERR_ARG;
  if (reinterpret_cast<uintptr_t>(dx) & 15) return PSF_ERR_ARG;
  {  // the weights are staged with 16-byte cp.async copies
    const void* ptrs[6] = {fs.W1, fs.b1, fs.W2, d->C > 0 ? g.W1 : nullptr, d->C > 0 ? g.b1 : nullptr,
                           d->C > 0 ? g.W2 : nullptr};
    for (int i = 0; i < 6; ++i)
      if (reinterpret_cast<uintptr_t>(ptrs[i]) & 15) return PSF_ERR_ARG;
  }
  if (workspace == nullptr || workspace_bytes < psf_generators_workspace_bytes(d->n_W)) return PSF_ERR_WORKSPACE;
  cudaStream_t st = (cudaStream_t)stream;
  const long long ntiles = (d->rows + kTB - 1) / kTB;
  long long grid = gen_grid(384 / kTB);
  if (grid > ntiles) grid = ntiles;
  if (grid < 1) grid = 1;
  PSF_CUDA_TRY(cudaMemsetAsync(workspace, 0, (size_t)grid * p.n_gen * kAccBlk * sizeof(float), st));
  p.Wt = const_cast<float*>(dW_in);
  p.V = const_cast<float*>(dV_in);
  p.dx = dx;
  p.ws = (float*)workspace;
  p.pre = const_cast<float*>(pre_saved);
  if (d->rows > 0 && pre_saved == nullptr && gen_use_tc()) {   // tcgen05 / tensor-memory kernels
    int tc_grid = 0, transposed = 1;
    void* images = (char*)workspace + gen_slices_bytes(d->n_W);
    rc = psf_gen_tc_prep(p, images, st);
    if (rc) return rc;
    rc = gen_bwd_version() == 1 ? PSF_ERR_UNSUPPORTED : psf_gen_tc_bwd2(p, images, &tc_grid, st);   // all products on tcgen05
    if (rc == PSF_ERR_UNSUPPORTED) {                              // ragged tiles: parameter gradients on the SIMT pipes
      transposed = 0;
      rc = psf_gen_tc_bwd(p, images, &tc_grid, st);
    }
    if (rc != PSF_ERR_UNSUPPORTED) {
      if (rc) return rc;
      GenGrads gg = {dfs.W1, dfs.b1, dfs.W2, dfs.b2, dg.W1, dg.b1, dg.W2, dg.b2};
      psf_gen_reduce_kernel<<<(p.n_gen * kAccBlk + 255) / 256, 256, 0, st>>>((const float*)workspace, tc_grid, p.n_W, p.n_gen,
                                                                             p.L, p.C, gg, transposed);
      return (int)cudaGetLastError();
    }
  }
  if (d->rows > 0) {
    if (pre_saved != nullptr)
      psf_gen_bwd_kernel<kTB, true><<<(unsigned)grid, kTB, kSmemBwd, st>>>(p);
    else if (kBwdAllMma && kTB == kMT)
      psf_gen_bwd_mma_kernel<<<(unsigned)grid, kMT, kSmemBwdMma, st>>>(p);
    else
      psf_gen_bwd_kernel<kTB, false><<<(unsigned)grid, kTB, kSmemBwd, st>>>(p);
    PSF_CUDA_TRY(cudaGetLastError());
  }
  GenGrads gg = {dfs.W1, dfs.b1, dfs.W2, dfs.b2, dg.W1, dg.b1, dg.W2, dg.b2};
  psf_gen_reduce_kernel<<<(p.n_gen * kAccBlk + 255) / 256, 256, 0, st>>>((const float*)workspace, (int)grid, p.n_W, p.n_gen,
                                                                         p.L, p.C, gg);
  return (int)cudaGetLastError();
}

}  // extern "C"
