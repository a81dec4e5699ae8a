// Epilogue of the tcgen05 GEMM kernels (gemm_tc.cu, gemm_tc2.cu): TMEM -> registers -> fused bias / GELU / dGELU /
// residual -> global, 32 columns of one row at a time, global operands software-pipelined one chunk ahead.
#pragma once
#include <cuda_fp16.h>

#include "epilogue.cuh"
#include "tc_common.cuh"

#ifndef BFK_GELU_H2
#define BFK_GELU_H2 1      // A/B knob: 0 = fp32 value + derivative in the GELU_*_GRAD epilogues
#endif

namespace bfk {

constexpr int TC_EPI_WARPS = 8;   // generic epilogue: two warps per TMEM lane quarter, alternating 32-column chunks
// Kernels compiled for one specialised EpiMode need far fewer registers and run four warps per lane quarter: the fused
// epilogues are instruction-bound (GELU / dGELU cost 10-25 instructions per element), more warps = more issue slots used.
#ifndef BFK_RES_EW
#define BFK_RES_EW 8       // epilogue warps of the bias + fp32-residual mode: its 2 x 32 operand registers (residual double buffer) spill at the
                           // 112-register cap of 16 warps (104 B, measured 37.7 us on 12544x768x768); 8 warps: 166 registers, 35.6 us, +1.3 % on the step
#endif
__host__ __device__ constexpr int epi_warps(int mode) { return mode == EPI_GENERIC ? 8 : (mode == EPI_RES_F32 ? BFK_RES_EW : 16); }

// ---- fast epilogue math for bf16 outputs (the fp32 parity mode uses the exact functions of epilogue.cuh) ----
__device__ __forceinline__ float tanh_fast(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float k = 0.7978845608028654f, c = 0.044715f;
  return 0.5f * x * (1.f + tanh_fast(k * (x + c * x * x * x)));
}
__device__ __forceinline__ float dgelu_tanh_fast(float x) {
  const float k = 0.7978845608028654f, c = 0.044715f;
  const float t = tanh_fast(k * (x + c * x * x * x));
  return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * k * (1.f + 3.f * c * x * x);
}

// erf-GELU for bf16 outputs: Phi(x) = 0.5 (1 + tanh(u(x))) holds exactly for u = atanh(erf(x / sqrt 2)); the odd polynomial
// u(x) = x (c0 + c1 x^2 + c2 x^4) fitted to it on |x| <= 6 (tools/fit_gelu_erf.py) leaves |x Phi - gelu(x)| < 3.0e-5 and
// |d/dx - gelu'(x)| < 1.3e-4, both far below the bf16 output resolution and tanh.approx's own 2^-11: ONE MUFU and ~14
// FMA/ALU per element for value + derivative instead of ex2 + rcp + ~25 (Abramowitz-Stegun 7.1.26 form before).
// The polynomial is monotone up to |x| = 8.3; x is clamped to +-8 inside it (tanh is exactly +-1 there).
constexpr float GE_C0 = 0.7974584707823228f, GE_C1 = 0.037050345100392894f, GE_C2 = -0.000358732361145582f;
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float xc = fminf(fmaxf(x, -8.f), 8.f), x2 = xc * xc;
  const float t = tanh_fast(xc * fmaf(x2, fmaf(x2, GE_C2, GE_C1), GE_C0));
  return x * fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float dgelu_erf_fast(float x) {
  const float xc = fminf(fmaxf(x, -8.f), 8.f), x2 = xc * xc;
  const float t = tanh_fast(xc * fmaf(x2, fmaf(x2, GE_C2, GE_C1), GE_C0));
  const float du = fmaf(x2, fmaf(x2, 5.f * GE_C2, 3.f * GE_C1), GE_C0);
  return fmaf(0.5f * xc * fmaf(-t, t, 1.f), du, fmaf(0.5f, t, 0.5f));
}

// value and derivative together (the GELU_*_GRAD forward epilogues): tanh / exp evaluated once
__device__ __forceinline__ void gelu_tanh_both_fast(float x, float& g, float& d) {
  const float k = 0.7978845608028654f, c = 0.044715f;
  const float x2 = x * x;
  const float t = tanh_fast(k * x * fmaf(c, x2, 1.f));
  const float hp = fmaf(0.5f, t, 0.5f);
  g = x * hp;
  d = fmaf(0.5f * x * (1.f - t * t), k * fmaf(3.f * c, x2, 1.f), hp);
}
__device__ __forceinline__ void gelu_erf_both_fast(float x, float& g, float& d) {
  const float xc = fminf(fmaxf(x, -8.f), 8.f), x2 = xc * xc;
  const float t = tanh_fast(xc * fmaf(x2, fmaf(x2, GE_C2, GE_C1), GE_C0));
  const float hp = fmaf(0.5f, t, 0.5f);
  g = x * hp;
  d = fmaf(0.5f * xc * fmaf(-t, t, 1.f), fmaf(x2, fmaf(x2, 5.f * GE_C2, 3.f * GE_C1), GE_C0), hp);
}

__device__ __forceinline__ uint32_t pack_bf16_f(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// ---- packed half-precision GELU value + derivative (the GELU_*_GRAD forward epilogues) --------------------------------------
// Two elements per instruction (HFMA2 / HMUL2 / one tanh.approx.f16x2 per PAIR): ~10 instructions per element instead of ~18 in
// fp32, which is what decides whether the epilogue of a K = 768 tile hides behind the next tile's main loop (3.2 us per tile).
// fp16 (11-bit mantissa) intermediates, bf16 results: simulated against the exact functions (tools/sim_gelu_h2.py) the rms error
// of the stored bf16 value / derivative is 1.02x / 1.05x that of rounding the exact result to bf16.
__device__ __forceinline__ __half2 h2_tanh(__half2 x) {
  uint32_t y, xi = *reinterpret_cast<uint32_t*>(&x);
  asm("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(xi));
  return *reinterpret_cast<__half2*>(&y);
}
__device__ __forceinline__ uint32_t h2_to_bf16x2(__half2 h) {
  const float2 f = __half22float2(h);
  return pack_bf16_f(f.x, f.y);
}
template <bool ERF>
__device__ __forceinline__ void gelu_both_h2(float x0, float x1, uint32_t& g_out, uint32_t& d_out) {
  const __half2 x = __floats2half2_rn(x0, x1);
  const __half2 k8 = __float2half2_rn(8.f), half = __float2half2_rn(0.5f), one = __float2half2_rn(1.f);
  const __half2 xc = __hmin2(__hmax2(x, __hneg2(k8)), k8);
  const __half2 x2 = __hmul2(xc, xc);
  __half2 p, du;
  if constexpr (ERF) {
    p = __hfma2(x2, __hfma2(x2, __float2half2_rn(GE_C2), __float2half2_rn(GE_C1)), __float2half2_rn(GE_C0));
    du = __hfma2(x2, __hfma2(x2, __float2half2_rn(5.f * GE_C2), __float2half2_rn(3.f * GE_C1)), __float2half2_rn(GE_C0));
  } else {
    constexpr float k = 0.7978845608028654f, c = 0.044715f;
    p = __hfma2(x2, __float2half2_rn(k * c), __float2half2_rn(k));
    du = __hfma2(x2, __float2half2_rn(3.f * k * c), __float2half2_rn(k));
  }
  const __half2 t = h2_tanh(__hmul2(xc, p));
  const __half2 hp = __hfma2(t, half, half);
  const __half2 g = __hmul2(x, hp);
  const __half2 omt2 = __hfma2(__hneg2(t), t, one);
  const __half2 d = __hfma2(__hmul2(__hmul2(xc, half), omt2), du, hp);
  g_out = h2_to_bf16x2(g);
  d_out = h2_to_bf16x2(d);
}

// 256-bit global accesses (LDG.256 / STG.256): one full 32-byte sector per thread and instruction
struct alignas(32) V8 { uint32_t w[8]; };
__device__ __forceinline__ V8 ldg256(const void* p) {
  V8 v;
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v.w[0]), "=r"(v.w[1]), "=r"(v.w[2]), "=r"(v.w[3]), "=r"(v.w[4]), "=r"(v.w[5]), "=r"(v.w[6]), "=r"(v.w[7])
               : "l"(p));
  return v;
}
__device__ __forceinline__ void stg256(void* p, const V8& v) {
  asm volatile("st.global.L1::no_allocate.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v.w[0]), "r"(v.w[1]), "r"(v.w[2]), "r"(v.w[3]),
               "r"(v.w[4]), "r"(v.w[5]), "r"(v.w[6]), "r"(v.w[7])
               : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Global operands of one 32-column chunk of one row (fp32 residual: 128 B, or bf16 pre-activation: 64 B)
struct EpiOperands { V8 w[4]; };

__device__ __forceinline__ bool epi_fast(const EpiParams& ep, int nc) {
  return (ep.vec_ok & 1) && !ep.accumulate && (nc + 32 <= ep.N) && ep.act < BFK_ACT_GELU_TANH_GRAD;
}
// Issue the global loads of chunk [nc, nc+32) of row m (nothing waits on them here).
__device__ __forceinline__ void epi_load(const EpiParams& ep, int m, int nc, EpiOperands& o) {
  if (m >= ep.M || !epi_fast(ep, nc)) return;
  if (ep.residual) {
    const int rm = ep.res_row_mod > 0 ? (m % ep.res_row_mod) : m;
    const float* rp = ep.residual + (long long)rm * ep.ldr + nc;
#pragma unroll
    for (int i = 0; i < 4; ++i) o.w[i] = ldg256(rp + 8 * i);
  } else if (ep.act == BFK_ACT_DGELU_TANH || ep.act == BFK_ACT_DGELU_ERF) {
    const __nv_bfloat16* hp = reinterpret_cast<const __nv_bfloat16*>(ep.aux) + (long long)m * ep.ld_aux + nc;
#pragma unroll
    for (int i = 0; i < 2; ++i) o.w[i] = ldg256(hp + 16 * i);
  }
}

// One 32-column chunk [nc, nc+32) of row m: TMEM -> registers -> fused epilogue -> global.  `o` holds the
// operands loaded earlier by epi_load (software-pipelined one chunk ahead); bias comes from shared memory.
__device__ __forceinline__ void epi_chunk32(const EpiParams& ep, int m, int nc, uint32_t taddr, const float* __restrict__ bias_s,
                                            const EpiOperands& o) {
  const bool row_ok = m < ep.M;
  uint32_t r[32];
  if (!(ep.vec_ok & 8)) {          // bit 3 (BFK_GEMM_DBG=8, measurement only): skip the TMEM read
    tc_ld32(taddr, r);
    tc_wait_ld();
  } else {
#pragma unroll
    for (int t = 0; t < 32; ++t) r[t] = 0u;
  }
  if (!row_ok) return;
  if (!epi_fast(ep, nc)) {
#pragma unroll
    for (int t = 0; t < 32; ++t)
      if (nc + t < ep.N) epi_store(ep, m, nc + t, __uint_as_float(r[t]));
    return;
  }
  const bool dgelu = ep.act == BFK_ACT_DGELU_TANH || ep.act == BFK_ACT_DGELU_ERF;
  float v[32];
#pragma unroll
  for (int t = 0; t < 32; ++t) v[t] = __uint_as_float(r[t]) * ep.alpha;
  if (ep.bias) {
#pragma unroll
    for (int t = 0; t < 32; ++t) v[t] += bias_s[t];
  }
  const bool out_bf16 = ep.out_dtype == BFK_BF16;
  if (ep.act == BFK_ACT_GELU_TANH || ep.act == BFK_ACT_GELU_ERF) {
    if (ep.aux) {
      if (ep.aux_dtype == BFK_BF16) {
        __nv_bfloat16* ap = reinterpret_cast<__nv_bfloat16*>(ep.aux) + (long long)m * ep.ld_aux + nc;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          V8 q;
#pragma unroll
          for (int t = 0; t < 8; ++t) q.w[t] = pack_bf16(v[i * 16 + 2 * t], v[i * 16 + 2 * t + 1]);
          stg256(ap + 16 * i, q);
        }
      } else {
        float* ap = reinterpret_cast<float*>(ep.aux) + (long long)m * ep.ld_aux + nc;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          V8 q;
#pragma unroll
          for (int t = 0; t < 8; ++t) q.w[t] = __float_as_uint(v[i * 8 + t]);
          stg256(ap + 8 * i, q);
        }
      }
    }
    if (ep.act == BFK_ACT_GELU_TANH) {
      if (out_bf16) {
#pragma unroll
        for (int t = 0; t < 32; ++t) v[t] = gelu_tanh_fast(v[t]);
      } else {
#pragma unroll
        for (int t = 0; t < 32; ++t) v[t] = gelu_tanh(v[t]);
      }
    } else if (out_bf16) {
#pragma unroll
      for (int t = 0; t < 32; ++t) v[t] = gelu_erf_fast(v[t]);
    } else {
#pragma unroll
      for (int t = 0; t < 32; ++t) v[t] = gelu_erf(v[t]);
    }
  } else if (dgelu) {
    float h[32];
#pragma unroll
    for (int t = 0; t < 16; ++t) {
      const uint32_t hw = o.w[t >> 3].w[t & 7];
      const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw));
      h[2 * t] = f.x; h[2 * t + 1] = f.y;
    }
    // one uniform branch per variant (a per-element select would execute the exact AND the fast path)
    if (ep.act == BFK_ACT_DGELU_TANH) {
      if (out_bf16) {
#pragma unroll
        for (int t = 0; t < 32; ++t) v[t] *= dgelu_tanh_fast(h[t]);
      } else {
#pragma unroll
        for (int t = 0; t < 32; ++t) v[t] *= dgelu_tanh(h[t]);
      }
    } else if (out_bf16) {
#pragma unroll
      for (int t = 0; t < 32; ++t) v[t] *= dgelu_erf_fast(h[t]);
    } else {
#pragma unroll
      for (int t = 0; t < 32; ++t) v[t] *= dgelu_erf(h[t]);
    }
  }
  if (ep.residual) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int t = 0; t < 8; ++t) v[i * 8 + t] += __uint_as_float(o.w[i].w[t]);
  }
  if (ep.vec_ok & 4) return;       // bit 2 (BFK_GEMM_DBG=4, measurement only): skip the output stores
  if (out_bf16) {
    __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(ep.D) + (long long)m * ep.ldd + nc;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      V8 q;
#pragma unroll
      for (int t = 0; t < 8; ++t) q.w[t] = pack_bf16(v[i * 16 + 2 * t], v[i * 16 + 2 * t + 1]);
      stg256(dp + 16 * i, q);
    }
  } else {
    float* dp = reinterpret_cast<float*>(ep.D) + (long long)m * ep.ldd + nc;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      V8 q;
#pragma unroll
      for (int t = 0; t < 8; ++t) q.w[t] = __float_as_uint(v[i * 8 + t]);
      stg256(dp + 8 * i, q);
    }
  }
}

// ---- specialised tile epilogues ----------------------------------------------------------------------------
// One warp's share of a 128 x BN accumulator tile: row m (TMEM lane), chunks c = half, half + 2, ... of 32 columns.
// MODE is a compile-time EpiMode, the chunk loop is fully unrolled: no per-chunk flag tests, no register copies
// between code paths (the generic epi_chunk32 spends ~60% of its instructions on those), operands of the next chunk
// are in flight while the current one is processed.  bias_s is this tile's bias in shared memory (zeros if none).
// `sub` = this warp's index within its lane quarter, CS = warps per quarter: chunks c = sub, sub + CS, ...; NCH = chunks per warp.
// The operands of the first chunk are loaded by epi_tile_preload BEFORE the caller waits for the accumulator.
template <int MODE> struct EpiOps {
  static constexpr bool RES = MODE == EPI_RES_F32;
  static constexpr bool DG = MODE == EPI_DGELU_TANH_BF16 || MODE == EPI_DGELU_ERF_BF16 || MODE == EPI_MULAUX_BF16;
  static constexpr int NOP = RES ? 4 : (DG ? 2 : 1);       // 32-byte operand vectors per chunk
  V8 a[NOP], b[NOP];
};
template <int MODE>
__device__ __forceinline__ void epi_load_ops(const EpiParams& ep, int m, int nc, V8 (&op)[EpiOps<MODE>::NOP]) {
  if constexpr (EpiOps<MODE>::RES) {
    const float* rrow = ep.residual + (long long)(ep.res_row_mod > 0 ? (m % ep.res_row_mod) : m) * ep.ldr;
#pragma unroll
    for (int i = 0; i < 4; ++i) if (nc + 8 * i < ep.N) op[i] = ldg256(rrow + nc + 8 * i);       // N % 16 == 0: a vector is all in or all out
  } else if constexpr (EpiOps<MODE>::DG) {
    const __nv_bfloat16* hrow = reinterpret_cast<const __nv_bfloat16*>(ep.aux) + (long long)m * ep.ld_aux;
#pragma unroll
    for (int i = 0; i < 2; ++i) if (nc + 16 * i < ep.N) op[i] = ldg256(hrow + nc + 16 * i);
  }
}
template <int MODE>
__device__ __forceinline__ void epi_tile_preload(const EpiParams& ep, int m, int n0, int sub, EpiOps<MODE>& ops) {
  if constexpr (EpiOps<MODE>::RES || EpiOps<MODE>::DG) {
    if (m < ep.M && n0 + sub * 32 < ep.N) epi_load_ops<MODE>(ep, m, n0 + sub * 32, ops.a);
  }
}

template <int MODE, int NCH, int CS>
__device__ __forceinline__ void epi_tile(const EpiParams& ep, int m, int n0, int sub, uint32_t taddr, const float* __restrict__ bias_s,
                                         EpiOps<MODE>& ops) {
  constexpr bool RES = MODE == EPI_RES_F32;
  constexpr bool DG = EpiOps<MODE>::DG;
  constexpr bool GE = MODE == EPI_GELU_TANH_BF16 || MODE == EPI_GELU_ERF_BF16;
  constexpr bool GG = MODE == EPI_GELUG_TANH_BF16 || MODE == EPI_GELUG_ERF_BF16;
  constexpr bool OUT_BF16 = MODE != EPI_STORE_F32 && MODE != EPI_RES_F32;
  const bool row_ok = m < ep.M;
  constexpr int NOP = EpiOps<MODE>::NOP;
  // chunk i of this warp: columns [nc, nc + 32); `op` holds its operands, `nxt` receives those of chunk i + 1
  auto chunk = [&](int i, V8 (&op)[NOP], V8 (&nxt)[NOP]) {
    const int c = sub + CS * i, nc = n0 + c * 32;
    if (nc >= ep.N) return;                                  // warp-uniform
    if ((RES || DG) && i + 1 < NCH && row_ok && nc + CS * 32 < ep.N) epi_load_ops<MODE>(ep, m, nc + CS * 32, nxt);
    uint32_t r[32];
    tc_ld32(taddr + (uint32_t)(c * 32), r);
    tc_wait_ld();
    if (!row_ok) return;
    if constexpr (MODE == EPI_STORE_BF16 || MODE == EPI_STORE_F32) {
      if (nc + 32 > ep.N) {                                  // ragged last chunk of the matrix: element-wise
#pragma unroll
        for (int t = 0; t < 32; ++t) {
          if (nc + t < ep.N) {
            const float x = __uint_as_float(r[t]) + bias_s[c * 32 + t];
            if constexpr (MODE == EPI_STORE_BF16) reinterpret_cast<__nv_bfloat16*>(ep.D)[(long long)m * ep.ldd + nc + t] = __float2bfloat16_rn(x);
            else reinterpret_cast<float*>(ep.D)[(long long)m * ep.ldd + nc + t] = x;
          }
        }
        return;
      }
    }
    float v[32];
    if constexpr (DG) {
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const uint32_t hw = op[t >> 3].w[t & 7];
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw));
        if constexpr (MODE == EPI_MULAUX_BF16) {
          v[2 * t] = __uint_as_float(r[2 * t]) * f.x;
          v[2 * t + 1] = __uint_as_float(r[2 * t + 1]) * f.y;
        } else if constexpr (MODE == EPI_DGELU_TANH_BF16) {
          v[2 * t] = __uint_as_float(r[2 * t]) * dgelu_tanh_fast(f.x);
          v[2 * t + 1] = __uint_as_float(r[2 * t + 1]) * dgelu_tanh_fast(f.y);
        } else {
          v[2 * t] = __uint_as_float(r[2 * t]) * dgelu_erf_fast(f.x);
          v[2 * t + 1] = __uint_as_float(r[2 * t + 1]) * dgelu_erf_fast(f.y);
        }
      }
    } else {
#pragma unroll
      for (int t = 0; t < 32; ++t) v[t] = __uint_as_float(r[t]) + bias_s[c * 32 + t];
    }
    if constexpr (RES) {
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int t = 0; t < 8; ++t) v[q * 8 + t] += __uint_as_float(op[q].w[t]);
    }
    if constexpr (GE) {
      if (ep.aux) {                                          // pre-activation for the backward pass
        __nv_bfloat16* ap = reinterpret_cast<__nv_bfloat16*>(ep.aux) + (long long)m * ep.ld_aux + nc;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          V8 w;
#pragma unroll
          for (int t = 0; t < 8; ++t) w.w[t] = pack_bf16(v[q * 16 + 2 * t], v[q * 16 + 2 * t + 1]);
          if (nc + 16 * q < ep.N) stg256(ap + 16 * q, w);
        }
      }
#pragma unroll
      for (int t = 0; t < 32; ++t) v[t] = MODE == EPI_GELU_TANH_BF16 ? gelu_tanh_fast(v[t]) : gelu_erf_fast(v[t]);
    }
    if constexpr (GG) {                                      // value -> D, derivative -> aux, 16 columns at a time
      __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(ep.D) + (long long)m * ep.ldd + nc;
      __nv_bfloat16* ap = reinterpret_cast<__nv_bfloat16*>(ep.aux) + (long long)m * ep.ld_aux + nc;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        V8 wg, wd;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
#if BFK_GELU_H2
          gelu_both_h2<MODE == EPI_GELUG_ERF_BF16>(v[q * 16 + 2 * t], v[q * 16 + 2 * t + 1], wg.w[t], wd.w[t]);
#else
          float g0, d0, g1, d1;
          if constexpr (MODE == EPI_GELUG_TANH_BF16) {
            gelu_tanh_both_fast(v[q * 16 + 2 * t], g0, d0);
            gelu_tanh_both_fast(v[q * 16 + 2 * t + 1], g1, d1);
          } else {
            gelu_erf_both_fast(v[q * 16 + 2 * t], g0, d0);
            gelu_erf_both_fast(v[q * 16 + 2 * t + 1], g1, d1);
          }
          wg.w[t] = pack_bf16(g0, g1);
          wd.w[t] = pack_bf16(d0, d1);
#endif
        }
        if (nc + 16 * q < ep.N) {          // ragged last chunk (N % 32 == 16, e.g. the so400m MLP width 4304): first half only
          stg256(dp + 16 * q, wg);
          stg256(ap + 16 * q, wd);
        }
      }
    } else if constexpr (OUT_BF16) {
      __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(ep.D) + (long long)m * ep.ldd + nc;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        V8 w;
#pragma unroll
        for (int t = 0; t < 8; ++t) w.w[t] = pack_bf16(v[q * 16 + 2 * t], v[q * 16 + 2 * t + 1]);
        if (nc + 16 * q < ep.N) stg256(dp + 16 * q, w);
      }
    } else {
      float* dp = reinterpret_cast<float*>(ep.D) + (long long)m * ep.ldd + nc;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        V8 w;
#pragma unroll
        for (int t = 0; t < 8; ++t) w.w[t] = __float_as_uint(v[q * 8 + t]);
        if (nc + 8 * q < ep.N) stg256(dp + 8 * q, w);
      }
    }
  };
  if constexpr (NCH == 1) {
    chunk(0, ops.a, ops.b);
  } else {
    static_assert(NCH % 2 == 0, "chunks per warp must be 1 or even");
#pragma unroll 1
    for (int i = 0; i < NCH; i += 2) {                       // operand buffers alternate by name: no register copies
      chunk(i, ops.a, ops.b);
      chunk(i + 1, ops.b, ops.a);
    }
  }
}

// ---- TMA-store variant of the bf16-output tile epilogues ---------------------------------------------------------------------
// A register store writes 32 bytes per thread and ROW: 32 sectors in 32 different lines per warp instruction, i.e. 32 LSU
// wavefronts.  With two bf16 outputs (value + derivative) that is 4096 wavefronts = 2.2 us per 128 x 256 tile and SM, which no
// longer hides behind the 3.2 us main loop of a K = 768 tile (ncu: tensor pipe 46 % busy against 79 % for the single-output
// store).  Here a warp stages 32 rows x 16 columns per output in shared memory (32-byte rows, SWIZZLE_32B: conflict-free
// 16-byte stores), one lane issues a TMA bulk-tensor store per box, and the copy engine writes the rows; rows / columns beyond
// the matrix are clipped by the tensor map.  `stage` = this warp's 2 KB: two 1 KB regions (value | derivative for the
// dual-output modes, ping-pong for the others).
constexpr int TS_STAGE_BYTES = 2048;
// Measured on 12544 x 3072 x 768 (profiles/r02_gemm_tma_store.txt): value + derivative 68 -> 58-63 us, GELU 52 -> 46 us, dGELU 77 -> 73 us;
// the plain store (44 -> 46 us) and the aux multiply (56 us both ways) do not gain and keep their register stores.
__host__ __device__ constexpr bool epi_mode_can_tma_store(int mode) {
  return mode == EPI_GELU_TANH_BF16 || mode == EPI_GELU_ERF_BF16 || mode == EPI_DGELU_TANH_BF16 ||
         mode == EPI_DGELU_ERF_BF16 || mode == EPI_GELUG_TANH_BF16 || mode == EPI_GELUG_ERF_BF16;
}
__device__ __forceinline__ void ts_stage_half(uint32_t region, int lane, const uint32_t (&w)[8]) {
  const uint32_t row = region + (uint32_t)lane * 32u, sw = ((uint32_t)lane >> 2) & 1u;     // SWIZZLE_32B: 16-byte chunk ^= bit 7 of the offset
  sts128(row + ((0u ^ sw) << 4), w[0], w[1], w[2], w[3]);
  sts128(row + ((1u ^ sw) << 4), w[4], w[5], w[6], w[7]);
}
// DBL: 4 KB of staging per warp (value + derivative modes of the CTA-pair kernel): both outputs ping-pong between two regions each
template <int MODE, int NCH, int CS, bool DBL = false>
__device__ __forceinline__ void epi_tile_ts(const EpiParams& ep, int m, int mrow0, int n0, int sub, uint32_t taddr, const float* __restrict__ bias_s,
                                            EpiOps<MODE>& ops, uint32_t stage, const CUtensorMap* map_d, const CUtensorMap* map_x, int lane) {
  constexpr bool DG = EpiOps<MODE>::DG;
  constexpr bool GE = MODE == EPI_GELU_TANH_BF16 || MODE == EPI_GELU_ERF_BF16;
  constexpr bool GG = MODE == EPI_GELUG_TANH_BF16 || MODE == EPI_GELUG_ERF_BF16;
  constexpr int NOP = EpiOps<MODE>::NOP;
  const bool row_ok = m < ep.M;
  int phase = 0;      // single-output modes: region of the next half chunk (ping-pong)
  auto chunk = [&](int i, V8 (&op)[NOP], V8 (&nxt)[NOP]) {
    const int c = sub + CS * i, nc = n0 + c * 32;
    if (nc >= ep.N) return;                                  // warp-uniform
    if (DG && i + 1 < NCH && row_ok && nc + CS * 32 < ep.N) epi_load_ops<MODE>(ep, m, nc + CS * 32, nxt);
    uint32_t r[32];
    tc_ld32(taddr + (uint32_t)(c * 32), r);
    tc_wait_ld();
    float v[32];
    if constexpr (DG) {
#pragma unroll
      for (int t = 0; t < 16; ++t) {
        const uint32_t hw = op[t >> 3].w[t & 7];
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&hw));
        if constexpr (MODE == EPI_MULAUX_BF16) {
          v[2 * t] = __uint_as_float(r[2 * t]) * f.x;
          v[2 * t + 1] = __uint_as_float(r[2 * t + 1]) * f.y;
        } else if constexpr (MODE == EPI_DGELU_TANH_BF16) {
          v[2 * t] = __uint_as_float(r[2 * t]) * dgelu_tanh_fast(f.x);
          v[2 * t + 1] = __uint_as_float(r[2 * t + 1]) * dgelu_tanh_fast(f.y);
        } else {
          v[2 * t] = __uint_as_float(r[2 * t]) * dgelu_erf_fast(f.x);
          v[2 * t + 1] = __uint_as_float(r[2 * t + 1]) * dgelu_erf_fast(f.y);
        }
      }
    } else {
#pragma unroll
      for (int t = 0; t < 32; ++t) v[t] = __uint_as_float(r[t]) + bias_s[c * 32 + t];
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      if (nc + 16 * q >= ep.N) break;                        // N % 32 == 16: the second half of the last chunk does not exist
      uint32_t wd[8], wx[8];
#pragma unroll
      for (int t = 0; t < 8; ++t) {
        const float a = v[q * 16 + 2 * t], b = v[q * 16 + 2 * t + 1];
        if constexpr (GG) {
          gelu_both_h2<MODE == EPI_GELUG_ERF_BF16>(a, b, wd[t], wx[t]);
        } else if constexpr (GE) {
          wx[t] = pack_bf16(a, b);                            // pre-activation (only stored when ep.aux is set)
          wd[t] = MODE == EPI_GELU_TANH_BF16 ? pack_bf16(gelu_tanh_fast(a), gelu_tanh_fast(b)) : pack_bf16(gelu_erf_fast(a), gelu_erf_fast(b));
        } else {
          wd[t] = pack_bf16(a, b);
        }
      }
      const bool two = GG || (GE && ep.aux != nullptr);      // warp-uniform
      // the region(s) about to be overwritten were handed to the copy engine one (two regions: two) group(s) ago
      if (lane == 0) { if (two && !DBL) bulk_wait_read<0>(); else bulk_wait_read<1>(); }
      __syncwarp();
      const uint32_t reg_d = stage + ((two && !DBL) ? 0u : (uint32_t)(phase * 1024)), reg_x = stage + (DBL ? 2048u + (uint32_t)(phase * 1024) : 1024u);
      ts_stage_half(reg_d, lane, wd);
      if (two) ts_stage_half(reg_x, lane, wx);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_store_2d(map_d, reg_d, nc + 16 * q, mrow0);
        if (two) tma_store_2d(map_x, reg_x, nc + 16 * q, mrow0);
        bulk_commit();
      }
      phase ^= 1;
    }
  };
  if constexpr (NCH == 1) {
    chunk(0, ops.a, ops.b);
  } else {
#pragma unroll 1
    for (int i = 0; i < NCH; i += 2) {
      chunk(i, ops.a, ops.b);
      chunk(i + 1, ops.b, ops.a);
    }
  }
}

// ---- TMA-load variant: the epilogue OPERAND (fp32 residual / bf16 aux factor) of a 32 x 32 chunk arrives as one TMA box ----------
// Same reasoning as for the stores: a per-thread 32-byte load touches one line per ROW, 32 LSU wavefronts per warp instruction
// (128 per chunk for the fp32 residual).  Here one lane requests the box of the NEXT chunk right after the warp has pulled the
// current one out of shared memory (swizzled, conflict-free 16-byte reads), so the copy engine does the row gather.  One
// mbarrier and one staging buffer per warp (2 KB bf16 / 4 KB fp32); rows / columns beyond the matrix arrive as zeros.
struct TlState { uint32_t parity; bool pending; };
__host__ __device__ constexpr bool epi_mode_can_tma_load(int mode) { return mode == EPI_RES_F32 || mode == EPI_MULAUX_BF16; }
template <int MODE> __host__ __device__ constexpr int tl_box_bytes() { return MODE == EPI_RES_F32 ? 4096 : 2048; }
template <int MODE>
__device__ __forceinline__ void tl_request(uint32_t stage, uint32_t bar, const CUtensorMap* map, int col, int row) {
  mbar_expect_tx(bar, tl_box_bytes<MODE>());
  tma_load_2d(stage, map, bar, col, row);
}
template <int MODE, int NCH, int CS>
__device__ __forceinline__ void epi_tile_tl(const EpiParams& ep, int m, int mrow0, int n0, int sub, uint32_t taddr, const float* __restrict__ bias_s,
                                            uint32_t stage, uint32_t bar, TlState& st, const CUtensorMap* map_op, int lane,
                                            bool has_next, int next_mrow0, int next_n0) {
  constexpr bool RES = MODE == EPI_RES_F32;
  const bool row_ok = m < ep.M;
  // this warp's first chunk of the tile: requested by the previous tile's epilogue (or the kernel prologue); rarely not
  if (n0 + sub * 32 < ep.N && !st.pending) {
    if (lane == 0) tl_request<MODE>(stage, bar, map_op, n0 + sub * 32, mrow0);
    st.pending = true;
  }
#pragma unroll 1
  for (int i = 0; i < NCH; ++i) {
    const int c = sub + CS * i, nc = n0 + c * 32;
    if (nc >= ep.N) break;                                   // warp-uniform
    mbar_wait(bar, st.parity);
    st.parity ^= 1u;
    uint32_t op[RES ? 32 : 16];
    if constexpr (RES) {
      const uint32_t rowa = stage + (uint32_t)lane * 128u, sw = (uint32_t)lane & 7u;        // SWIZZLE_128B
#pragma unroll
      for (int q = 0; q < 8; ++q) lds128(rowa + (((uint32_t)q ^ sw) << 4), op[4 * q], op[4 * q + 1], op[4 * q + 2], op[4 * q + 3]);
    } else {
      const uint32_t rowa = stage + (uint32_t)lane * 64u, sw = ((uint32_t)lane >> 1) & 3u;  // SWIZZLE_64B
#pragma unroll
      for (int q = 0; q < 4; ++q) lds128(rowa + (((uint32_t)q ^ sw) << 4), op[4 * q], op[4 * q + 1], op[4 * q + 2], op[4 * q + 3]);
    }
    __syncwarp();
    // the buffer is free: request the next chunk of this tile, else this warp's first chunk of the CTA's next tile
    st.pending = false;
    {
      int ncol = -1, nrow = 0;
      if (i + 1 < NCH && nc + CS * 32 < ep.N) { ncol = nc + CS * 32; nrow = mrow0; }
      else if (has_next && next_n0 + sub * 32 < ep.N) { ncol = next_n0 + sub * 32; nrow = next_mrow0; }
      if (ncol >= 0) {
        if (lane == 0) tl_request<MODE>(stage, bar, map_op, ncol, nrow);
        st.pending = true;
      }
    }
    uint32_t r[32];
    tc_ld32(taddr + (uint32_t)(c * 32), r);
    tc_wait_ld();
    if (!row_ok) continue;
    if constexpr (RES) {
      float* dp = reinterpret_cast<float*>(ep.D) + (long long)m * ep.ldd + nc;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        V8 w;
#pragma unroll
        for (int t = 0; t < 8; ++t) w.w[t] = __float_as_uint(__uint_as_float(r[q * 8 + t]) + bias_s[c * 32 + q * 8 + t] + __uint_as_float(op[q * 8 + t]));
        if (nc + 8 * q < ep.N) stg256(dp + 8 * q, w);
      }
    } else {
      __nv_bfloat16* dp = reinterpret_cast<__nv_bfloat16*>(ep.D) + (long long)m * ep.ldd + nc;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        V8 w;
#pragma unroll
        for (int t = 0; t < 8; ++t) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&op[q * 8 + t]));
          w.w[t] = pack_bf16(__uint_as_float(r[q * 16 + 2 * t]) * f.x, __uint_as_float(r[q * 16 + 2 * t + 1]) * f.y);
        }
        if (nc + 16 * q < ep.N) stg256(dp + 16 * q, w);
      }
    }
  }
  // a tile in which this warp owns no chunk still has to line up the next tile's first box
  if (!st.pending && has_next && next_n0 + sub * 32 < ep.N) {
    if (lane == 0) tl_request<MODE>(stage, bar, map_op, next_n0 + sub * 32, next_mrow0);
    st.pending = true;
  }
}

// L2 prefetch of the operands the epilogue of (row m, chunk nc) will read (next work item of this CTA)
__device__ __forceinline__ void epi_prefetch(const EpiParams& ep, int m, int nc) {
  if (m >= ep.M || nc + 32 > ep.N) return;
  if (ep.residual) {
    const int rm = ep.res_row_mod > 0 ? (m % ep.res_row_mod) : m;
    prefetch_l2(ep.residual + (long long)rm * ep.ldr + nc);
  }
  if (act_reads_aux(ep.act))
    prefetch_l2(reinterpret_cast<const __nv_bfloat16*>(ep.aux) + (long long)m * ep.ld_aux + nc);
}


}  // namespace bfk
