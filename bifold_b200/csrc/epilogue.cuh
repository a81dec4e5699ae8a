// Fused GEMM epilogue shared by the SIMT (fp32 parity mode) and tcgen05 (bf16) GEMM kernels:
//   out = act(alpha*acc + bias[n]) (+ residual[m', n]);  optional pre-activation store / dGELU multiply.
#pragma once
#include "common.cuh"

namespace bfk {

struct EpiParams {
  int M, N;
  void* D; long long ldd; int out_dtype;
  const float* bias;
  int act;
  void* aux; long long ld_aux; int aux_dtype;
  const float* residual; long long ldr; int res_row_mod;
  float alpha;
  int accumulate;
  int vec_ok;  // host-verified: D/aux/residual/bias pointers and leading dims allow 32-byte vector access
  int tma_store;   // CTA-pair kernel, bf16-output specialised modes: outputs leave through shared memory + TMA bulk-tensor stores
  int tma_load;    // CTA-pair kernel, residual / aux-multiply modes: the epilogue operand arrives through TMA + shared memory
  int mode;    // EpiMode: compile-time specialised tcgen05 epilogue for the hot combinations (tc_epilogue.cuh), 0 = generic
};

// Specialised epilogues of the tcgen05 kernels.  All require vec_ok, alpha == 1, no accumulate; all but the two plain
// stores also N % 32 == 0.
enum EpiMode : int {
  EPI_GENERIC = 0,
  EPI_STORE_BF16,        // D(bf16) = acc (+ bias)
  EPI_STORE_F32,         // D(fp32) = acc (+ bias)
  EPI_RES_F32,           // D(fp32) = acc (+ bias) + residual(fp32)
  EPI_GELU_TANH_BF16,    // aux(bf16, optional) = acc + bias;  D(bf16) = gelu_tanh(acc + bias)
  EPI_GELU_ERF_BF16,
  EPI_DGELU_TANH_BF16,   // D(bf16) = acc * gelu_tanh'(aux(bf16))
  EPI_DGELU_ERF_BF16,
  EPI_GELUG_TANH_BF16,   // v = acc + bias;  aux(bf16) = gelu_tanh'(v);  D(bf16) = gelu_tanh(v)
  EPI_GELUG_ERF_BF16,
  EPI_MULAUX_BF16,       // D(bf16) = acc * aux(bf16)
};
__host__ __device__ inline bool act_reads_aux(int act) { return act == BFK_ACT_DGELU_TANH || act == BFK_ACT_DGELU_ERF || act == BFK_ACT_MUL_AUX; }
__host__ __device__ inline bool act_is_gelu_fwd(int act) {
  return act == BFK_ACT_GELU_TANH || act == BFK_ACT_GELU_ERF || act == BFK_ACT_GELU_TANH_GRAD || act == BFK_ACT_GELU_ERF_GRAD;
}

__device__ __forceinline__ float epi_value(const EpiParams& p, int m, int n, float acc) {
  float v = acc * p.alpha;
  if (p.bias) v += __ldg(p.bias + n);
  if (act_is_gelu_fwd(p.act)) {
    const bool tanh_form = p.act == BFK_ACT_GELU_TANH || p.act == BFK_ACT_GELU_TANH_GRAD;
    if (p.aux) {
      float sv = v;     // GELU_*: the pre-activation;  GELU_*_GRAD: the derivative at the pre-activation
      if (p.act == BFK_ACT_GELU_TANH_GRAD || p.act == BFK_ACT_GELU_ERF_GRAD) sv = tanh_form ? dgelu_tanh(v) : dgelu_erf(v);
      if (p.aux_dtype == BFK_BF16) ((__nv_bfloat16*)p.aux)[(long long)m * p.ld_aux + n] = __float2bfloat16_rn(sv);
      else ((float*)p.aux)[(long long)m * p.ld_aux + n] = sv;
    }
    v = tanh_form ? gelu_tanh(v) : gelu_erf(v);
  } else if (act_reads_aux(p.act)) {
    float h = (p.aux_dtype == BFK_BF16) ? __bfloat162float(((const __nv_bfloat16*)p.aux)[(long long)m * p.ld_aux + n])
                                        : ((const float*)p.aux)[(long long)m * p.ld_aux + n];
    v *= (p.act == BFK_ACT_MUL_AUX) ? h : ((p.act == BFK_ACT_DGELU_TANH) ? dgelu_tanh(h) : dgelu_erf(h));
  }
  if (p.residual) {
    int rm = p.res_row_mod > 0 ? (m % p.res_row_mod) : m;
    v += __ldg(p.residual + (long long)rm * p.ldr + n);
  }
  return v;
}

__device__ __forceinline__ void epi_store(const EpiParams& p, int m, int n, float acc) {
  float v = epi_value(p, m, n, acc);
  if (p.out_dtype == BFK_BF16) {
    ((__nv_bfloat16*)p.D)[(long long)m * p.ldd + n] = __float2bfloat16_rn(v);
  } else {
    float* d = (float*)p.D + (long long)m * p.ldd + n;
    *d = p.accumulate ? (*d + v) : v;
  }
}

// Row chunk of 8 consecutive columns [n0, n0+8) of row m held in registers (tcgen05 epilogue).
// Vector path when the chunk is fully inside N and the host verified alignment.
__device__ __forceinline__ void epi_store8(const EpiParams& p, int m, int n0, const float (&acc)[8]) {
  if (m >= p.M || n0 >= p.N) return;
  if (!p.vec_ok || n0 + 8 > p.N || p.act >= BFK_ACT_GELU_TANH_GRAD) {   // the GRAD / MUL_AUX acts only have the scalar form here
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (n0 + j < p.N) epi_store(p, m, n0 + j, acc[j]);
    return;
  }
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = acc[j] * p.alpha;
  if (p.bias) {
    float4 b0 = __ldg((const float4*)(p.bias + n0)), b1 = __ldg((const float4*)(p.bias + n0 + 4));
    v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w;
    v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
  }
  if (p.act == BFK_ACT_GELU_TANH || p.act == BFK_ACT_GELU_ERF) {
    if (p.aux) {
      if (p.aux_dtype == BFK_BF16) {
        __nv_bfloat162 h[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
        *(uint4*)((__nv_bfloat16*)p.aux + (long long)m * p.ld_aux + n0) = *(uint4*)h;
      } else {
        float* a = (float*)p.aux + (long long)m * p.ld_aux + n0;
        *(float4*)a = make_float4(v[0], v[1], v[2], v[3]);
        *(float4*)(a + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (p.act == BFK_ACT_GELU_TANH) ? gelu_tanh(v[j]) : gelu_erf(v[j]);
  } else if (p.act == BFK_ACT_DGELU_TANH || p.act == BFK_ACT_DGELU_ERF) {
    float h[8];
    if (p.aux_dtype == BFK_BF16) {
      uint4 raw = *(const uint4*)((const __nv_bfloat16*)p.aux + (long long)m * p.ld_aux + n0);
      const __nv_bfloat162* hp = (const __nv_bfloat162*)&raw;
#pragma unroll
      for (int j = 0; j < 4; ++j) { float2 f = __bfloat1622float2(hp[j]); h[2 * j] = f.x; h[2 * j + 1] = f.y; }
    } else {
      const float* a = (const float*)p.aux + (long long)m * p.ld_aux + n0;
      float4 a0 = *(const float4*)a, a1 = *(const float4*)(a + 4);
      h[0] = a0.x; h[1] = a0.y; h[2] = a0.z; h[3] = a0.w; h[4] = a1.x; h[5] = a1.y; h[6] = a1.z; h[7] = a1.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= (p.act == BFK_ACT_DGELU_TANH) ? dgelu_tanh(h[j]) : dgelu_erf(h[j]);
  }
  if (p.residual) {
    int rm = p.res_row_mod > 0 ? (m % p.res_row_mod) : m;
    const float* r = p.residual + (long long)rm * p.ldr + n0;
    float4 r0 = __ldg((const float4*)r), r1 = __ldg((const float4*)(r + 4));
    v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
    v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
  }
  if (p.out_dtype == BFK_BF16) {
    __nv_bfloat162 o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    *(uint4*)((__nv_bfloat16*)p.D + (long long)m * p.ldd + n0) = *(uint4*)o;
  } else {
    float* d = (float*)p.D + (long long)m * p.ldd + n0;
    if (p.accumulate) {
      float4 d0 = *(float4*)d, d1 = *(float4*)(d + 4);
      v[0] += d0.x; v[1] += d0.y; v[2] += d0.z; v[3] += d0.w;
      v[4] += d1.x; v[5] += d1.y; v[6] += d1.z; v[7] += d1.w;
    }
    *(float4*)d = make_float4(v[0], v[1], v[2], v[3]);
    *(float4*)(d + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// Host side: fill EpiParams from the public args and check alignment for the vector path.
inline bool epi_from_args(const bfk_gemm_args* a, EpiParams& p) {
  p.M = a->M; p.N = a->N;
  p.D = a->D; p.ldd = a->ldd; p.out_dtype = a->out_dtype;
  p.bias = a->bias; p.act = a->act;
  p.aux = a->aux; p.ld_aux = a->ld_aux;
  p.aux_dtype = act_reads_aux(a->act) ? a->in_dtype : a->out_dtype;
  p.residual = a->residual; p.ldr = a->ldr; p.res_row_mod = a->res_row_mod;
  p.alpha = a->alpha; p.accumulate = a->accumulate;
  auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31u) == 0; };
  size_t es = a->out_dtype == BFK_BF16 ? 2 : 4, as = p.aux_dtype == BFK_BF16 ? 2 : 4;
  bool ok = al32(a->D) && (a->ldd * es) % 32 == 0;
  if (a->bias) ok = ok && al32(a->bias);
  if (a->aux) ok = ok && al32(a->aux) && (a->ld_aux * as) % 32 == 0;
  if (a->residual) ok = ok && al32(a->residual) && (a->ldr * 4) % 32 == 0;
  p.vec_ok = ok ? 1 : 0;
  p.tma_store = 0;
  p.tma_load = 0;
  p.mode = EPI_GENERIC;
  if (ok && a->alpha == 1.f && !a->accumulate && a->act == BFK_ACT_NONE && !a->residual) {
    p.mode = a->out_dtype == BFK_BF16 ? EPI_STORE_BF16 : EPI_STORE_F32;   // plain stores also take a ragged last chunk
  } else if (ok && a->alpha == 1.f && !a->accumulate && a->N % 16 == 0) {   // N % 32 == 16: the tile epilogues guard the last half chunk
    const bool bf = a->out_dtype == BFK_BF16;
    if (false) {}
    else if (a->act == BFK_ACT_NONE && a->residual && !bf) p.mode = EPI_RES_F32;
    else if (a->act == BFK_ACT_GELU_TANH && !a->residual && bf && p.aux_dtype == BFK_BF16) p.mode = EPI_GELU_TANH_BF16;
    else if (a->act == BFK_ACT_GELU_ERF && !a->residual && bf && p.aux_dtype == BFK_BF16) p.mode = EPI_GELU_ERF_BF16;
    else if (a->act == BFK_ACT_DGELU_TANH && !a->residual && !a->bias && bf && p.aux_dtype == BFK_BF16) p.mode = EPI_DGELU_TANH_BF16;
    else if (a->act == BFK_ACT_DGELU_ERF && !a->residual && !a->bias && bf && p.aux_dtype == BFK_BF16) p.mode = EPI_DGELU_ERF_BF16;
    else if (a->act == BFK_ACT_GELU_TANH_GRAD && !a->residual && bf && a->aux && p.aux_dtype == BFK_BF16) p.mode = EPI_GELUG_TANH_BF16;
    else if (a->act == BFK_ACT_GELU_ERF_GRAD && !a->residual && bf && a->aux && p.aux_dtype == BFK_BF16) p.mode = EPI_GELUG_ERF_BF16;
    else if (a->act == BFK_ACT_MUL_AUX && !a->residual && !a->bias && bf && p.aux_dtype == BFK_BF16) p.mode = EPI_MULAUX_BF16;
  }
  return true;
}

}  // namespace bfk
