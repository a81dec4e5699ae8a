// LayerNorm forward / backward: one warp per row, fp32 statistics, float4 accesses.
// HBM-bound: forward reads 4D bytes and writes 2D (bf16) or 4D per row; backward reads dy, x (+dres)
// and writes dx (+ a bf16 copy).
#include <stdlib.h>
#include "common.cuh"

namespace bfk {

constexpr int LN_WARPS = 8;
constexpr int LN_MAXC = 12;  // float4 chunks per lane: D <= 1536 (kernels are templated on the actual chunk count)

template <typename T> __device__ __forceinline__ float4 load4(const T* p, int c);
template <> __device__ __forceinline__ float4 load4<float>(const float* p, int c) {
  return reinterpret_cast<const float4*>(p)[c];
}
template <> __device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p, int c) {
  uint2 u = reinterpret_cast<const uint2*>(p)[c];
  float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.x));
  float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
template <typename T> __device__ __forceinline__ void store4(T* p, int c, float4 v);
template <> __device__ __forceinline__ void store4<float>(float* p, int c, float4 v) {
  reinterpret_cast<float4*>(p)[c] = v;
}
template <> __device__ __forceinline__ void store4<__nv_bfloat16>(__nv_bfloat16* p, int c, float4 v) {
  __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&p0);
  u.y = *reinterpret_cast<uint32_t*>(&p1);
  reinterpret_cast<uint2*>(p)[c] = u;
}

template <typename TY, int LN_MAXC, bool LORA = false>
__global__ void __launch_bounds__(LN_WARPS * 32) ln_fwd_kernel(bfk_layernorm_fwd_args a) {
  pdl_trigger();   // the GEMM behind this pass may set itself up while the last rows are normalised (it waits for this grid to finish)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * LN_WARPS + warp;
  if (row >= a.M) return;
  const int nc = a.D >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(a.x + (long long)row * a.ldx);
  float4 xv[LN_MAXC];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXC; ++i) {
    int c = lane + 32 * i;
    if (c < nc) {
      xv[i] = x4[c];
      s += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
    }
  }
  const float mean = warp_sum(s) / (float)a.D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < LN_MAXC; ++i) {
    int c = lane + 32 * i;
    if (c < nc) {
      float d0 = xv[i].x - mean, d1 = xv[i].y - mean, d2 = xv[i].z - mean, d3 = xv[i].w - mean;
      q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)a.D + a.eps);
  if (lane == 0) {
    if (a.mean) a.mean[row] = mean;
    if (a.rstd) a.rstd[row] = rstd;
  }
  const float4* g4 = reinterpret_cast<const float4*>(a.gamma);
  const float4* b4 = reinterpret_cast<const float4*>(a.beta);
  TY* y = reinterpret_cast<TY*>(a.y) + (long long)row * a.ldy;
  TY* y2 = a.y2 ? reinterpret_cast<TY*>(a.y2) + (long long)row * a.ldy2 : nullptr;
  const float keep2 = a.y2_dropout_p > 0.f ? 1.f / (1.f - a.y2_dropout_p) : 1.f;
  const uint64_t seed2 = a.y2_seed + (a.seed_ptr ? (uint64_t)*a.seed_ptr * 4096ull : 0ull);   // replay counter: same stride as the host step counter
  float lacc[LORA ? 16 : 1];
  if constexpr (LORA) {
#pragma unroll
    for (int j = 0; j < 16; ++j) lacc[j] = 0.f;
  }
#pragma unroll
  for (int i = 0; i < LN_MAXC; ++i) {
    int c = lane + 32 * i;
    if (c < nc) {
      float4 g = __ldg(g4 + c), b = __ldg(b4 + c);
      float o0 = (xv[i].x - mean) * rstd * g.x + b.x, o1 = (xv[i].y - mean) * rstd * g.y + b.y;
      float o2 = (xv[i].z - mean) * rstd * g.z + b.z, o3 = (xv[i].w - mean) * rstd * g.w + b.w;
      if constexpr (sizeof(TY) == 4) {
        reinterpret_cast<float4*>(y)[c] = make_float4(o0, o1, o2, o3);
      } else {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(o0, o1), p1 = __floats2bfloat162_rn(o2, o3);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&p0);
        u.y = *reinterpret_cast<uint32_t*>(&p1);
        reinterpret_cast<uint2*>(y)[c] = u;
        // the dropped copy starts from the ROUNDED value (what a separate dropout pass over y would read)
        const float2 r0 = __bfloat1622float2(p0), r1 = __bfloat1622float2(p1);
        o0 = r0.x; o1 = r0.y; o2 = r1.x; o3 = r1.y;
      }
      if (y2) {
        if (a.y2_dropout_p > 0.f) {
          o0 = hash_uniform(seed2, row, 4 * c + 0) < a.y2_dropout_p ? 0.f : o0 * keep2;
          o1 = hash_uniform(seed2, row, 4 * c + 1) < a.y2_dropout_p ? 0.f : o1 * keep2;
          o2 = hash_uniform(seed2, row, 4 * c + 2) < a.y2_dropout_p ? 0.f : o2 * keep2;
          o3 = hash_uniform(seed2, row, 4 * c + 3) < a.y2_dropout_p ? 0.f : o3 * keep2;
        }
        store4<TY>(y2, c, make_float4(o0, o1, o2, o3));
      }
      if constexpr (LORA) {
        // u += drop(y) . A^T on the values as stored (bf16): what a GEMM over the stored rows would read
        const float2 q0 = __bfloat1622float2(__floats2bfloat162_rn(o0, o1)), q1 = __bfloat1622float2(__floats2bfloat162_rn(o2, o3));
        const uint2* A2 = reinterpret_cast<const uint2*>(a.lora_a) + c;     // row j starts at j * (D / 4)
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint2 aw = __ldg(A2 + (long long)j * nc);
          const float2 a0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw.x));
          const float2 a1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&aw.y));
          lacc[j] = fmaf(q0.x, a0.x, fmaf(q0.y, a0.y, fmaf(q1.x, a1.x, fmaf(q1.y, a1.y, lacc[j]))));
        }
      }
    }
  }
  if constexpr (LORA) {
    float mine = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float t = warp_sum(lacc[j]);
      if (lane == j) mine = t;
    }
    if (lane < 16) reinterpret_cast<__nv_bfloat16*>(a.lora_u)[(long long)row * a.ld_u + lane] = __float2bfloat16_rn(mine);
  }
}

// LoRA down-projection u[m, 0..16) = sum_k h[m, k] * A[j, k] over the bf16 rows the LayerNorm pass has just written (still in
// L2): warp-level tensor-core MMA (m16n8k16), 16 rows per CTA, the K range split over the 4 warps.  Every thread loads 16
// contiguous bytes of "its" row per 32-wide K block (whole 32-byte sectors per row); the k order inside an MMA is a free
// permutation as long as both operands use the same one, so the same 16-byte pattern is used for A's rows.
// Replaces 16 per-row dot products inside ln_fwd_kernel (96 L1 loads + ~600 FMA/convert instructions per lane and row).
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int NB>   // NB = 32-wide K blocks per warp (D = 128 * NB)
__global__ void __launch_bounds__(128) lora_down_mma_kernel(const __nv_bfloat16* __restrict__ h, long long ldh, int M,
                                                            const __nv_bfloat16* __restrict__ A, int D,
                                                            __nv_bfloat16* __restrict__ u, long long ldu) {
  __shared__ float red[4][16][17];
  pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int row0 = blockIdx.x * 16;
  const int r0 = min(row0 + g, M - 1), r1 = min(row0 + g + 8, M - 1);      // clamped loads, predicated stores
  const uint4* h0 = reinterpret_cast<const uint4*>(h + (long long)r0 * ldh) + t;
  const uint4* h1 = reinterpret_cast<const uint4*>(h + (long long)r1 * ldh) + t;
  const uint4* A0 = reinterpret_cast<const uint4*>(A + (long long)g * D) + t;
  const uint4* A1 = reinterpret_cast<const uint4*>(A + (long long)(g + 8) * D) + t;
  uint4 ra[NB], rb[NB], b0[NB], b1[NB];
#pragma unroll
  for (int i = 0; i < NB; ++i) {           // K block (warp + 4 i): uint4 index 4 * block + t
    const int o = 4 * (warp + 4 * i);
    ra[i] = __ldg(h0 + o); rb[i] = __ldg(h1 + o);
  }
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const int o = 4 * (warp + 4 * i);
    b0[i] = __ldg(A0 + o); b1[i] = __ldg(A1 + o);
  }
  float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    mma16816(c0, ra[i].x, rb[i].x, ra[i].y, rb[i].y, b0[i].x, b0[i].y);
    mma16816(c1, ra[i].x, rb[i].x, ra[i].y, rb[i].y, b1[i].x, b1[i].y);
    mma16816(c0, ra[i].z, rb[i].z, ra[i].w, rb[i].w, b0[i].z, b0[i].w);
    mma16816(c1, ra[i].z, rb[i].z, ra[i].w, rb[i].w, b1[i].z, b1[i].w);
  }
  red[warp][g][2 * t] = c0[0]; red[warp][g][2 * t + 1] = c0[1];
  red[warp][g + 8][2 * t] = c0[2]; red[warp][g + 8][2 * t + 1] = c0[3];
  red[warp][g][8 + 2 * t] = c1[0]; red[warp][g][8 + 2 * t + 1] = c1[1];
  red[warp][g + 8][8 + 2 * t] = c1[2]; red[warp][g + 8][8 + 2 * t + 1] = c1[3];
  __syncthreads();
  // 128 threads: row = tid / 8, column pair = tid % 8; fixed summation order over the 4 K slices
  const int r = threadIdx.x >> 3, cp = (threadIdx.x & 7) * 2;
  if (row0 + r < M) {
    const float s0 = (red[0][r][cp] + red[1][r][cp]) + (red[2][r][cp] + red[3][r][cp]);
    const float s1 = (red[0][r][cp + 1] + red[1][r][cp + 1]) + (red[2][r][cp + 1] + red[3][r][cp + 1]);
    *reinterpret_cast<__nv_bfloat162*>(u + (long long)(row0 + r) * ldu + cp) = __floats2bfloat162_rn(s0, s1);
  }
}

// raw (as stored) 4-element chunk: loads are issued in one batch per operand with no control flow in between, so a
// warp has every byte of its row in flight at once; conversion happens afterwards
template <typename T> struct Raw4;
template <> struct Raw4<float> { using type = float4; };
template <> struct Raw4<__nv_bfloat16> { using type = uint2; };
__device__ __forceinline__ float4 cvt4(float4 v) { return v; }
__device__ __forceinline__ float4 cvt4(uint2 u) {
  float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.x));
  float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}

template <typename TDY, typename TDX2, bool PARAM_GRADS, int LN_MAXC>
__global__ void __launch_bounds__(LN_WARPS * 32, (PARAM_GRADS || LN_MAXC > 6) ? 1 : 2) ln_bwd_kernel(bfk_layernorm_bwd_args a, int rows_per_block) {
  using RawDY = typename Raw4<TDY>::type;
  __shared__ float red[PARAM_GRADS ? LN_WARPS : 1][PARAM_GRADS ? 132 : 1];
  pdl_trigger();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nc = a.D >> 2;
  const float4* g4 = reinterpret_cast<const float4*>(a.gamma);
  float4 dg[LN_MAXC], db[LN_MAXC];
  if constexpr (PARAM_GRADS) {
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i) dg[i] = db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const int r_begin = blockIdx.x * rows_per_block;
  const int r_end = min(a.M, r_begin + rows_per_block);
  const float keep2 = a.dy2_dropout_p > 0.f ? 1.f / (1.f - a.dy2_dropout_p) : 1.f;
  const uint64_t seed2 = a.dy2_seed + (a.seed_ptr ? (uint64_t)*a.seed_ptr * 4096ull : 0ull);
  for (int row = r_begin + warp; row < r_end; row += LN_WARPS) {
    const float4* x4 = reinterpret_cast<const float4*>(a.x + (long long)row * a.ldx);
    const RawDY* dy = reinterpret_cast<const RawDY*>(reinterpret_cast<const TDY*>(a.dy) + (long long)row * a.ld_dy);
    const RawDY* dy2 = a.dy2 ? reinterpret_cast<const RawDY*>(reinterpret_cast<const TDY*>(a.dy2) + (long long)row * a.ld_dy2) : nullptr;
    const float4* dres = a.dres ? reinterpret_cast<const float4*>(a.dres + (long long)row * a.ld_dres) : nullptr;
    // ---- load phase: every operand of the row, back to back ----
    float4 xv[LN_MAXC], rv[LN_MAXC];
    RawDY dv[LN_MAXC], ev[LN_MAXC];
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i) if (lane + 32 * i < nc) xv[i] = x4[lane + 32 * i];
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i) if (lane + 32 * i < nc) dv[i] = dy[lane + 32 * i];
    if (dy2) {
#pragma unroll
      for (int i = 0; i < LN_MAXC; ++i) if (lane + 32 * i < nc) ev[i] = dy2[lane + 32 * i];
    }
    if (dres) {
#pragma unroll
      for (int i = 0; i < LN_MAXC; ++i) if (lane + 32 * i < nc) rv[i] = dres[lane + 32 * i];
    }
    const float mean = a.mean[row], rstd = a.rstd[row];
    // ---- statistics ----
    float4 gv[LN_MAXC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i) {
      const int c = lane + 32 * i;
      if (c < nc) {
        float4 d = cvt4(dv[i]);
        const float4 g = __ldg(g4 + c);
        if (dy2) {
          float4 e = cvt4(ev[i]);
          if (a.dy2_dropout_p > 0.f) {
            e.x = hash_uniform(seed2, row, 4 * c + 0) < a.dy2_dropout_p ? 0.f : e.x * keep2;
            e.y = hash_uniform(seed2, row, 4 * c + 1) < a.dy2_dropout_p ? 0.f : e.y * keep2;
            e.z = hash_uniform(seed2, row, 4 * c + 2) < a.dy2_dropout_p ? 0.f : e.z * keep2;
            e.w = hash_uniform(seed2, row, 4 * c + 3) < a.dy2_dropout_p ? 0.f : e.w * keep2;
          }
          d.x += e.x; d.y += e.y; d.z += e.z; d.w += e.w;
        }
        // xv becomes x-hat in place
        xv[i] = make_float4((xv[i].x - mean) * rstd, (xv[i].y - mean) * rstd, (xv[i].z - mean) * rstd, (xv[i].w - mean) * rstd);
        if constexpr (PARAM_GRADS) {
          dg[i].x += d.x * xv[i].x; dg[i].y += d.y * xv[i].y; dg[i].z += d.z * xv[i].z; dg[i].w += d.w * xv[i].w;
          db[i].x += d.x; db[i].y += d.y; db[i].z += d.z; db[i].w += d.w;
        }
        gv[i] = make_float4(d.x * g.x, d.y * g.y, d.z * g.z, d.w * g.w);
        s1 += (gv[i].x + gv[i].y) + (gv[i].z + gv[i].w);
        s2 += (gv[i].x * xv[i].x + gv[i].y * xv[i].y) + (gv[i].z * xv[i].z + gv[i].w * xv[i].w);
      }
    }
    const float c1 = warp_sum(s1) / (float)a.D, c2 = warp_sum(s2) / (float)a.D;
    float* dx = a.dx + (long long)row * a.ld_dx;
    TDX2* dx2 = a.dx2 ? reinterpret_cast<TDX2*>(a.dx2) + (long long)row * a.ld_dx2 : nullptr;
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i) {
      const int c = lane + 32 * i;
      if (c < nc) {
        float4 o = make_float4(rstd * (gv[i].x - c1 - xv[i].x * c2), rstd * (gv[i].y - c1 - xv[i].y * c2),
                               rstd * (gv[i].z - c1 - xv[i].z * c2), rstd * (gv[i].w - c1 - xv[i].w * c2));
        if (dres) { o.x += rv[i].x; o.y += rv[i].y; o.z += rv[i].z; o.w += rv[i].w; }
        store4<float>(dx, c, o);
        if (dx2) store4<TDX2>(dx2, c, o);
      }
    }
  }
  if constexpr (PARAM_GRADS) {
    // block reduce of per-warp column sums, 128 columns (one float4 chunk per lane) at a time
    float* pg = a.partials + (long long)blockIdx.x * 2 * a.D;
#pragma unroll
    for (int i = 0; i < LN_MAXC; ++i) {
      int c = lane + 32 * i;
      if (32 * i >= nc) break;
      for (int pass = 0; pass < 2; ++pass) {
        float4 v = pass == 0 ? dg[i] : db[i];
        red[warp][lane * 4 + 0] = v.x; red[warp][lane * 4 + 1] = v.y;
        red[warp][lane * 4 + 2] = v.z; red[warp][lane * 4 + 3] = v.w;
        __syncthreads();
        if (warp == 0 && c < nc) {
          float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int w = 0; w < LN_WARPS; ++w) {
            t.x += red[w][lane * 4 + 0]; t.y += red[w][lane * 4 + 1];
            t.z += red[w][lane * 4 + 2]; t.w += red[w][lane * 4 + 3];
          }
          reinterpret_cast<float4*>(pg + pass * a.D)[c] = t;
        }
        __syncthreads();
      }
    }
  }
}

// 32 columns x 8 row-slices per block; fixed summation order (deterministic)
__global__ void __launch_bounds__(256) ln_bwd_reduce(const float* __restrict__ partials, int nb, int D,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
  __shared__ float red[8][33];
  const int col = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + col;
  float s = 0.f;
  if (n < 2 * D) {
    // four independent chains (fixed assignment b -> chain, fixed combination order): the loads of a slice are in flight together
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    int b = slice;
    for (; b + 24 < nb; b += 32) {
      s0 += partials[(long long)b * 2 * D + n];
      s1 += partials[(long long)(b + 8) * 2 * D + n];
      s2 += partials[(long long)(b + 16) * 2 * D + n];
      s3 += partials[(long long)(b + 24) * 2 * D + n];
    }
    for (; b < nb; b += 8) s0 += partials[(long long)b * 2 * D + n];
    s = (s0 + s1) + (s2 + s3);
  }
  red[slice][col] = s;
  __syncthreads();
  if (slice == 0 && n < 2 * D) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][col];
    if (n < D) dgamma[n] = t; else dbeta[n - D] = t;
  }
}

static int ln_bwd_blocks(int M, int& rows_per_block, bool param_grads = true) {
  if (!param_grads) {   // no per-block partial sums to amortise: one row per warp, blocks scheduled dynamically
    rows_per_block = LN_WARPS;
    return (M + LN_WARPS - 1) / LN_WARPS;
  }
  int target = sm_count() * 2;
  rows_per_block = (M + target - 1) / target;
  rows_per_block = ((rows_per_block + LN_WARPS - 1) / LN_WARPS) * LN_WARPS;
  return (M + rows_per_block - 1) / rows_per_block;
}

}  // namespace bfk

namespace bfk {
static bool lora_down_fused_forced() {          // BFK_LORA_DOWN_FUSED=1: measurement / test knob, keeps the in-LayerNorm dot products
  static const bool v = [] { const char* e = getenv("BFK_LORA_DOWN_FUSED"); return e && e[0] == '1'; }();
  return v;
}
}  // namespace bfk

extern "C" int32_t bfk_layernorm_fwd(const bfk_layernorm_fwd_args* a, bfk_stream_t s_) {
  using namespace bfk;
  BFK_CHECK_ARG(a && a->x && a->y && a->gamma && a->beta, "bfk_layernorm_fwd: null pointer");
  BFK_CHECK_ARG(!a->y2 || a->ldy2 % 4 == 0, "bfk_layernorm_fwd: ldy2 %% 4 != 0");
  BFK_CHECK_ARG(!a->lora_a || (a->y_dtype == BFK_BF16 && a->lora_u && (reinterpret_cast<uintptr_t>(a->lora_a) & 7u) == 0),
                "bfk_layernorm_fwd: the fused LoRA down-projection needs bf16 y, lora_u and an 8-byte aligned lora_a");
  BFK_CHECK_ARG(a->M > 0 && a->D > 0 && a->D % 4 == 0 && a->D <= 128 * LN_MAXC && a->ldx % 4 == 0 && a->ldy % 4 == 0,
                "bfk_layernorm_fwd: need D%%4==0, D<=%d, ld%%4==0 (D=%d)", 128 * LN_MAXC, a->D);
  cudaStream_t s = to_stream(s_);
  int grid = (a->M + LN_WARPS - 1) / LN_WARPS;
  // LoRA down-projection: a second, tensor-core launch over the rows just written (D = 768 / 1152, 16-byte aligned rows);
  // other shapes keep the per-row dot products inside the LayerNorm kernel
  const void* lsrc = a->y2 ? a->y2 : a->y;
  const long long lld = a->y2 ? a->ldy2 : a->ldy;
  const bool lora_mma = a->lora_a && a->y_dtype == BFK_BF16 && (a->D == 768 || a->D == 1152) && lld % 8 == 0 && a->ld_u % 2 == 0 &&
                        ((reinterpret_cast<uintptr_t>(lsrc) | reinterpret_cast<uintptr_t>(a->lora_a)) & 15u) == 0 &&
                        (reinterpret_cast<uintptr_t>(a->lora_u) & 3u) == 0 && !lora_down_fused_forced();
  const bool lora_fused = a->lora_a && !lora_mma;
#define LN_FWD(NC)                                                                          \
  do {                                                                                      \
    if (a->y_dtype == BFK_BF16 && lora_fused) ln_fwd_kernel<__nv_bfloat16, NC, true><<<grid, LN_WARPS * 32, 0, s>>>(*a); \
    else if (a->y_dtype == BFK_BF16) ln_fwd_kernel<__nv_bfloat16, NC><<<grid, LN_WARPS * 32, 0, s>>>(*a); \
    else ln_fwd_kernel<float, NC><<<grid, LN_WARPS * 32, 0, s>>>(*a);                         \
  } while (0)
  if (a->D <= 256) LN_FWD(2); else if (a->D <= 768) LN_FWD(6); else if (a->D <= 1152) LN_FWD(9); else LN_FWD(12);
#undef LN_FWD
  if (lora_mma) {
    const int g16 = (a->M + 15) / 16;
    if (a->D == 768)
      lora_down_mma_kernel<6><<<g16, 128, 0, s>>>((const __nv_bfloat16*)lsrc, lld, a->M, (const __nv_bfloat16*)a->lora_a, a->D, (__nv_bfloat16*)a->lora_u, a->ld_u);
    else
      lora_down_mma_kernel<9><<<g16, 128, 0, s>>>((const __nv_bfloat16*)lsrc, lld, a->M, (const __nv_bfloat16*)a->lora_a, a->D, (__nv_bfloat16*)a->lora_u, a->ld_u);
  }
  BFK_CHECK_LAUNCH("bfk_layernorm_fwd");
  return BFK_OK;
}

extern "C" size_t bfk_layernorm_bwd_workspace_floats(int32_t M, int32_t D) {
  int rpb;
  return (size_t)bfk::ln_bwd_blocks(M, rpb) * 2 * (size_t)D;
}

extern "C" int32_t bfk_layernorm_bwd(const bfk_layernorm_bwd_args* a, bfk_stream_t s_) {
  using namespace bfk;
  BFK_CHECK_ARG(a && a->dy && a->x && a->mean && a->rstd && a->gamma && a->dx, "bfk_layernorm_bwd: null pointer");
  BFK_CHECK_ARG(a->M > 0 && a->D % 4 == 0 && a->D <= 128 * LN_MAXC && a->ldx % 4 == 0 && a->ld_dy % 4 == 0 &&
                    a->ld_dx % 4 == 0 && (!a->dres || a->ld_dres % 4 == 0) && (!a->dx2 || a->ld_dx2 % 4 == 0) && (!a->dy2 || a->ld_dy2 % 4 == 0),
                "bfk_layernorm_bwd: need D%%4==0, D<=%d, ld%%4==0", 128 * LN_MAXC);
  const bool pg = a->dgamma != nullptr;
  BFK_CHECK_ARG(!pg || (a->dbeta && a->partials), "bfk_layernorm_bwd: dgamma needs dbeta and partials");
  cudaStream_t s = to_stream(s_);
  int rpb;
  int grid = ln_bwd_blocks(a->M, rpb, pg);
  const bool dyb = a->dy_dtype == BFK_BF16, d2b = a->dx2 && a->dx2_dtype == BFK_BF16;
#define LAUNCH(TDY, TDX2, PG)                                                                      \
  do {                                                                                             \
    if (a->D <= 256) ln_bwd_kernel<TDY, TDX2, PG, 2><<<grid, LN_WARPS * 32, 0, s>>>(*a, rpb);        \
    else if (a->D <= 768) ln_bwd_kernel<TDY, TDX2, PG, 6><<<grid, LN_WARPS * 32, 0, s>>>(*a, rpb);   \
    else if (a->D <= 1152) ln_bwd_kernel<TDY, TDX2, PG, 9><<<grid, LN_WARPS * 32, 0, s>>>(*a, rpb);  \
    else ln_bwd_kernel<TDY, TDX2, PG, 12><<<grid, LN_WARPS * 32, 0, s>>>(*a, rpb);                   \
  } while (0)
  if (dyb) {
    if (d2b) { if (pg) LAUNCH(__nv_bfloat16, __nv_bfloat16, true); else LAUNCH(__nv_bfloat16, __nv_bfloat16, false); }
    else     { if (pg) LAUNCH(__nv_bfloat16, float, true); else LAUNCH(__nv_bfloat16, float, false); }
  } else {
    if (d2b) { if (pg) LAUNCH(float, __nv_bfloat16, true); else LAUNCH(float, __nv_bfloat16, false); }
    else     { if (pg) LAUNCH(float, float, true); else LAUNCH(float, float, false); }
  }
#undef LAUNCH
  if (pg) ln_bwd_reduce<<<(2 * a->D + 31) / 32, 256, 0, s>>>(a->partials, grid, a->D, a->dgamma, a->dbeta);
  BFK_CHECK_LAUNCH("bfk_layernorm_bwd");
  return BFK_OK;
}
