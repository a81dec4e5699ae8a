// Heatmap tail: collapsed ConvDecoder (bifold/models/modules.py:33-74 is linear), composite
// x16 bilinear upsample, sigmoid, nn.BCELoss, first-max argmax and the get_action gate.
// All HBM-bound: the forward heatmap kernel writes heads*H*W*4 B per sample (+ reads as much target).
#include <math.h>
#include <vector>

#include "common.cuh"

namespace bfk {

// order-preserving float -> uint32 map (handles negatives), used for packed argmax keys
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
__device__ __forceinline__ unsigned long long pack_key(float v, uint32_t idx) {
  return ((unsigned long long)f2ord(v) << 32) | (unsigned long long)(0xFFFFFFFFu - idx);
}

__device__ __forceinline__ float exp2f_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float4 ldg_stream4(const float* p) {   // read-once data: do not keep it in L1
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream4(float* p, const float4& v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- collapsed decoder ---------------------------------------------------------------------------
// out[c] = sum_r v[r] * W[r*C + c]   (thread per column)
__device__ void vecmat(const float* __restrict__ v_s, const float* __restrict__ W, int R, int C, float* __restrict__ out_s) {
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;   // 4 independent chains, 8 loads in flight
    int r = 0;
#pragma unroll 2
    for (; r + 4 <= R; r += 4) {
      s0 = fmaf(v_s[r], W[(long long)r * C + c], s0);
      s1 = fmaf(v_s[r + 1], W[(long long)(r + 1) * C + c], s1);
      s2 = fmaf(v_s[r + 2], W[(long long)(r + 2) * C + c], s2);
      s3 = fmaf(v_s[r + 3], W[(long long)(r + 3) * C + c], s3);
    }
    for (; r < R; ++r) s0 = fmaf(v_s[r], W[(long long)r * C + c], s0);
    out_s[c] = (s0 + s1) + (s2 + s3);
  }
}
// out[r] = sum_c W[r*C + c] * g[c] + extra * bias[r]   (warp per row)
__device__ void matvec(const float* __restrict__ W, const float* __restrict__ g_s, int R, int C, float extra,
                       const float* __restrict__ bias, float* __restrict__ out_s) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = warp; r < R; r += nw) {
    float s = 0.f;
    for (int c = lane; c < C; c += 32) s = fmaf(W[(long long)r * C + c], g_s[c], s);
    s = warp_sum(s);
    if (lane == 0) out_s[r] = s + extra * bias[r];
  }
}
__device__ float block_dot(const float* __restrict__ a_s, const float* __restrict__ b, int n, float* red_s) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s = fmaf(a_s[i], b[i], s);
  s = warp_sum(s);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red_s[threadIdx.x >> 5] = s;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red_s[w];
  __syncthreads();
  return t;
}

// grid: heads; smem: 5*D floats
__global__ void __launch_bounds__(1024) collapse_fwd_kernel(bfk_decoder_collapse_args a) {
  extern __shared__ float sm[];
  __shared__ float red[32];
  const int h = blockIdx.x, D = a.D, c1 = D / 2, c2 = D / 4;
  float *v5 = sm, *v4 = sm + D, *v3 = sm + 2 * D, *v2 = sm + 3 * D, *we = sm + 4 * D;
  const float* const* W = a.W + h * 5;
  const float* const* b = a.b + h * 5;
  for (int c = threadIdx.x; c < c2; c += blockDim.x) v5[c] = W[4][c];
  __syncthreads();
  vecmat(v5, W[3], c2, c2, v4); __syncthreads();
  vecmat(v4, W[2], c2, c1, v3); __syncthreads();
  vecmat(v3, W[1], c1, c1, v2); __syncthreads();
  vecmat(v2, W[0], c1, D, we);  __syncthreads();
  float be = block_dot(v2, b[0], c1, red) + block_dot(v3, b[1], c1, red) + block_dot(v4, b[2], c2, red) +
             block_dot(v5, b[3], c2, red) + b[4][0];
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    a.w_eff[(long long)h * D + c] = we[c];
    float* ch = a.chain + (long long)h * 4 * D;
    ch[c] = c < c1 ? v2[c] : 0.f;
    ch[D + c] = c < c1 ? v3[c] : 0.f;
    ch[2 * D + c] = c < c2 ? v4[c] : 0.f;
    ch[3 * D + c] = c < c2 ? v5[c] : 0.f;
  }
  if (threadIdx.x == 0) a.b_eff[h] = be;
}

// outer product dW[r][c] = u[r]*g[c]
__device__ void outer(float* __restrict__ dW, const float* __restrict__ u_s, const float* __restrict__ g_s, int R, int C) {
  for (long long i = threadIdx.x; i < (long long)R * C; i += blockDim.x) dW[i] = u_s[i / C] * g_s[i % C];
}

// grid: heads; smem: 7*D floats
__global__ void __launch_bounds__(1024) collapse_bwd_kernel(bfk_decoder_collapse_args a) {
  extern __shared__ float sm[];
  const int h = blockIdx.x, D = a.D, c1 = D / 2, c2 = D / 4;
  float *v2 = sm, *v3 = sm + D, *v4 = sm + 2 * D, *v5 = sm + 3 * D, *ga = sm + 4 * D, *gb = sm + 5 * D;
  const float* const* W = a.W + h * 5;
  const float* const* b = a.b + h * 5;
  float* const* dW = a.dW + h * 5;
  float* const* db = a.db + h * 5;
  const float g_b = a.d_b_eff[h];
  const float* ch = a.chain + (long long)h * 4 * D;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    v2[c] = ch[c]; v3[c] = ch[D + c]; v4[c] = ch[2 * D + c]; v5[c] = ch[3 * D + c];
    ga[c] = a.d_w_eff[(long long)h * D + c];
  }
  __syncthreads();
  // layer 1: W1 (c1 x D)
  outer(dW[0], v2, ga, c1, D);
  for (int r = threadIdx.x; r < c1; r += blockDim.x) db[0][r] = g_b * v2[r];
  matvec(W[0], ga, c1, D, g_b, b[0], gb);  // dv2
  __syncthreads();
  // layer 2: W2 (c1 x c1)
  outer(dW[1], v3, gb, c1, c1);
  for (int r = threadIdx.x; r < c1; r += blockDim.x) db[1][r] = g_b * v3[r];
  matvec(W[1], gb, c1, c1, g_b, b[1], ga);  // dv3
  __syncthreads();
  // layer 3: W3 (c2 x c1)
  outer(dW[2], v4, ga, c2, c1);
  for (int r = threadIdx.x; r < c2; r += blockDim.x) db[2][r] = g_b * v4[r];
  matvec(W[2], ga, c2, c1, g_b, b[2], gb);  // dv4
  __syncthreads();
  // layer 4: W4 (c2 x c2)
  outer(dW[3], v5, gb, c2, c2);
  for (int r = threadIdx.x; r < c2; r += blockDim.x) db[3][r] = g_b * v5[r];
  matvec(W[3], gb, c2, c2, g_b, b[3], ga);  // dv5
  __syncthreads();
  // layer 5: W5 (1 x c2)
  for (int c = threadIdx.x; c < c2; c += blockDim.x) dW[4][c] = ga[c];
  if (threadIdx.x == 0) db[4][0] = g_b;
}

// ---- head projection -------------------------------------------------------------------------------
constexpr int HP_TOK_PER_WARP = 4;
// warp per token (HP_TOK_PER_WARP tokens in sequence); w_eff staged in smem.  grid: ceil(B*Np/(8*HP_TOK_PER_WARP))
__global__ void __launch_bounds__(256) head_project_fwd_kernel(bfk_head_project_args a) {
  extern __shared__ __align__(16) float w_s[];  // [heads][D]
  for (int i = threadIdx.x; i < a.heads * a.D; i += blockDim.x) w_s[i] = a.w_eff[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int it = 0; it < HP_TOK_PER_WARP; ++it) {   // the staged w_eff is reused for several tokens
    const long long tok = ((long long)blockIdx.x * HP_TOK_PER_WARP + it) * 8 + warp;
    if (tok >= (long long)a.B * a.Np) return;
    const int b = (int)(tok / a.Np), n = (int)(tok % a.Np);
    const float* xr = a.X + (long long)b * a.x_batch_stride + (long long)(a.row0 + n) * a.D;
    float acc[8] = {};
    for (int c = lane; c < a.D / 4; c += 32) {
      const float4 x = ldg_stream4(xr + 4 * c);
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        if (h < a.heads) {
          const float4 w = *reinterpret_cast<const float4*>(w_s + h * a.D + c * 4);
          acc[h] += x.x * w.x + x.y * w.y + x.z * w.z + x.w * w.w;
        }
      }
    }
#pragma unroll
    for (int h = 0; h < 8; ++h) {
      if (h < a.heads) {
        float s = warp_sum(acc[h]);
        if (lane == 0) a.grid[((long long)b * a.heads + h) * a.Np + n] = s + a.b_eff[h];
      }
    }
  }
}

// dX rows: grid (Nf, B)
template <typename T2>
__global__ void head_project_bwd_dx_kernel(bfk_head_project_args a) {
  const int r = blockIdx.x, b = blockIdx.y;
  const int n = r - a.row0;
  float4* o4 = reinterpret_cast<float4*>(a.dX + ((long long)b * a.Nf + r) * a.D);
  T2* o2 = a.dX2 ? reinterpret_cast<T2*>(a.dX2) + ((long long)b * a.Nf + r) * a.D : nullptr;
  float g[8];
  const bool live = n >= 0 && n < a.Np;
#pragma unroll
  for (int h = 0; h < 8; ++h) g[h] = (live && h < a.heads) ? a.dgrid[((long long)b * a.heads + h) * a.Np + n] : 0.f;
  for (int c = threadIdx.x; c < a.D / 4; c += blockDim.x) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        if (h < a.heads) {
          float4 w = __ldg(reinterpret_cast<const float4*>(a.w_eff + (long long)h * a.D) + c);
          v.x = fmaf(g[h], w.x, v.x); v.y = fmaf(g[h], w.y, v.y); v.z = fmaf(g[h], w.z, v.z); v.w = fmaf(g[h], w.w, v.w);
        }
      }
    }
    o4[c] = v;
    if (o2) {
      if constexpr (sizeof(T2) == 4) {
        reinterpret_cast<float4*>(o2)[c] = v;
      } else {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v.x, v.y), p1 = __floats2bfloat162_rn(v.z, v.w);
        uint2 u; u.x = *reinterpret_cast<uint32_t*>(&p0); u.y = *reinterpret_cast<uint32_t*>(&p1);
        reinterpret_cast<uint2*>(o2)[c] = u;
      }
    }
  }
}

constexpr int HP_CHUNK = 64;  // tokens per first-stage block
// grid (ceil(D/128), chunks): partials[chunk][h][d]
__global__ void head_project_bwd_dw_stage1(bfk_head_project_args a) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= a.D) return;
  const long long t0 = (long long)blockIdx.y * HP_CHUNK, t1 = min((long long)a.B * a.Np, t0 + HP_CHUNK);
  float acc[8] = {};
  for (long long t = t0; t < t1; ++t) {
    int b = (int)(t / a.Np), n = (int)(t % a.Np);
    float x = a.X[(long long)b * a.x_batch_stride + (long long)(a.row0 + n) * a.D + d];
#pragma unroll
    for (int h = 0; h < 8; ++h)
      if (h < a.heads) acc[h] = fmaf(a.dgrid[((long long)b * a.heads + h) * a.Np + n], x, acc[h]);
  }
#pragma unroll
  for (int h = 0; h < 8; ++h)
    if (h < a.heads) a.partials[((long long)blockIdx.y * a.heads + h) * a.D + d] = acc[h];
}
__global__ void head_project_bwd_dw_stage2(bfk_head_project_args a, int chunks) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.heads * a.D) return;
  float s = 0.f;
  for (int c = 0; c < chunks; ++c) s += a.partials[(long long)c * a.heads * a.D + t];
  a.d_w_eff[t] = s;
}
// grid: heads
__global__ void __launch_bounds__(256) head_project_bwd_db_kernel(bfk_head_project_args a) {
  __shared__ float red[8];
  const int h = blockIdx.x;
  float s = 0.f;
  for (long long t = threadIdx.x; t < (long long)a.B * a.Np; t += blockDim.x) {
    int b = (int)(t / a.Np), n = (int)(t % a.Np);
    s += a.dgrid[((long long)b * a.heads + h) * a.Np + n];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    a.d_b_eff[h] = t;
  }
}

// ---- heatmap forward: x16 upsample + sigmoid (+ BCE partials, + argmax) ---------------------------
constexpr int HM_ROWS = 32;  // backward: output rows per CTA
constexpr int HM_THREADS = 256;
constexpr int HM_MAXIN = HM_ROWS / 16 + 4;  // input rows feeding a backward band
constexpr int HF_ROWS = 32;  // forward: output rows per CTA
constexpr int HF_MAXIN = HF_ROWS / 16 + 4;
constexpr int HF_BATCH = 4;  // rows of one thread whose target loads are in flight together

// One pixel: p = sigmoid(z) and (LOSS) its nn.BCELoss term in units of ln 2, both from e = 2^(-|z| log2 e):
//   p = z >= 0 ? 1/(1+e) : e/(1+e);   -[t log p + (1-t) log(1-p)] = max(z,0) - t z + log(1 + e)          (3 MUFU: ex2, rcp, lg2)
// The closed form is what torch computes from its ROUNDED p up to the rounding of 1 - p, i.e. to <= 6e-8 / (1 - p) absolute:
// below 3e-5 for z <= BCE_ZFAST.  Above it (and where exp(-z) overflows: p == 0, log p == -inf) torch's result is dominated
// by that rounding and by its -100 clamp, so those pixels take the literal path: logs of the rounded p, clamped.
constexpr float BCE_ZFAST = 6.f;
__device__ __forceinline__ float sigmoid_e(float z, float& e) {      // p = sigmoid(z), e = 2^(-|z| log2 e)
  e = exp2f_fast(-fabsf(z) * 1.4426950408889634f);
  const float r = __fdividef(1.f, 1.f + e);
  return z >= 0.f ? r : e * r;
}
// four pixels: probabilities and (LOSS) the sum of their BCE terms in units of ln 2; ONE (almost never divergent) branch per four
template <bool LOSS>
__device__ __forceinline__ float4 sigmoid_bce4(const float4& z, const float4& t, float& loss_l2) {
  float4 e, p;
  p.x = sigmoid_e(z.x, e.x); p.y = sigmoid_e(z.y, e.y); p.z = sigmoid_e(z.z, e.z); p.w = sigmoid_e(z.w, e.w);
  if (LOSS) {
    const float zmax = fmaxf(fmaxf(z.x, z.y), fmaxf(z.z, z.w)), zmin = fminf(fminf(z.x, z.y), fminf(z.z, z.w));
    if (zmax <= BCE_ZFAST && zmin >= -80.f) {
      const float lin = (fmaxf(z.x, 0.f) - t.x * z.x) + (fmaxf(z.y, 0.f) - t.y * z.y) + (fmaxf(z.z, 0.f) - t.z * z.z) + (fmaxf(z.w, 0.f) - t.w * z.w);
      // log(1+ex) + log(1+ey) + ... = log2 of the product: two MUFU instead of four (each factor is in (1, 2])
      loss_l2 += fmaf(lin, 1.4426950408889634f, __log2f((1.f + e.x) * (1.f + e.y)) + __log2f((1.f + e.z) * (1.f + e.w)));
    } else {
      const float cl = -144.26950408889634f;   // -100 / ln 2
      loss_l2 -= t.x * fmaxf(__log2f(p.x), cl) + (1.f - t.x) * fmaxf(__log2f(1.f - p.x), cl);
      loss_l2 -= t.y * fmaxf(__log2f(p.y), cl) + (1.f - t.y) * fmaxf(__log2f(1.f - p.y), cl);
      loss_l2 -= t.z * fmaxf(__log2f(p.z), cl) + (1.f - t.z) * fmaxf(__log2f(1.f - p.z), cl);
      loss_l2 -= t.w * fmaxf(__log2f(p.w), cl) + (1.f - t.w) * fmaxf(__log2f(1.f - p.w), cl);
    }
  }
  return p;
}

// grid: (bands, heads, B).  smem: tmp HF_MAXIN*W floats, taps (3 weights + start per output index), grid g*g
template <bool STREAM>
__global__ void __launch_bounds__(HM_THREADS, 4) heatmap_fwd_kernel(bfk_heatmap_fwd_args a) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float red_f[HM_THREADS / 32];
  __shared__ unsigned long long red_k[HM_THREADS / 32];
  const int g = a.g, W = 16 * g, H = W;
  const int band = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int y0 = band * HF_ROWS, y1 = min(H, y0 + HF_ROWS);
  // thread -> fixed float4 column, rows y0+grp, y0+grp+groups, ...
  const int W4 = W / 4;
  const int groups = blockDim.x / W4;
  const int grp = threadIdx.x / W4, x = (threadIdx.x % W4) * 4;
  const bool has_t = a.target != nullptr;
  const float* tgt = has_t ? a.target[h] + (long long)b * H * W : nullptr;
  const float* tmask = (has_t && a.target_mask && ((a.mask_heads >> h) & 1u)) ? a.target_mask + (long long)b * H * W : nullptr;
  auto load_t = [&](int y) -> float4 {
    float4 tv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_t && grp < groups && y < y1) {
      const long long off = (long long)y * W + x;
      tv = STREAM ? ldg_stream4(tgt + off) : __ldg(reinterpret_cast<const float4*>(tgt + off));
    }
    return tv;      // raw target: nothing may depend on it before the set-up below is done (the optional mask is applied at the use)
  };
  // the targets do not depend on anything this CTA computes: their loads go out FIRST and fly during the set-up below
  float4 tv[HF_BATCH];
#pragma unroll
  for (int u = 0; u < HF_BATCH; ++u) tv[u] = load_t(y0 + grp + u * groups);

  float* tmp_s = sm;                                   // [HF_MAXIN][W], 16-byte aligned
  float* tapw_s = tmp_s + HF_MAXIN * W;                // [W][3]
  int* taps_s = reinterpret_cast<int*>(tapw_s + 3 * W);   // [W]
  float* grid_s = reinterpret_cast<float*>(taps_s + W);   // [g][g]
  const float* gsrc = a.grid + ((long long)b * a.heads + h) * g * g;
  for (int i = threadIdx.x; i < g * g; i += blockDim.x) grid_s[i] = gsrc[i];
  for (int i = threadIdx.x; i < W; i += blockDim.x) taps_s[i] = a.tap_start[i];
  for (int i = threadIdx.x; i < 3 * W; i += blockDim.x) tapw_s[i] = a.tap_w[i];
  __syncthreads();
  const int i_lo = taps_s[y0];
  const int i_hi = min(g, taps_s[y1 - 1] + 3);  // exclusive
  // horizontal pass for the needed input rows
  for (int t = threadIdx.x; t < (i_hi - i_lo) * W; t += blockDim.x) {
    int ii = t / W, xx = t % W;
    int js = taps_s[xx];
    const float* w = tapw_s + 3 * xx;
    const float* row = grid_s + (i_lo + ii) * g;
    float v = w[0] * row[js];
    if (js + 1 < g) v = fmaf(w[1], row[js + 1], v);
    if (js + 2 < g) v = fmaf(w[2], row[js + 2], v);
    tmp_s[ii * W + xx] = v;
  }
  __syncthreads();
  float* heat = a.heat + ((long long)h * a.B + b) * H * W;
  float loss = 0.f;            // in units of ln 2
  float best = -INFINITY;
  uint32_t best_idx = 0xFFFFFFFFu;
  const int nin = i_hi - i_lo;
  if (grp < groups) {
    for (int yb = y0 + grp; yb < y1; yb += HF_BATCH * groups) {
      float4 tn[HF_BATCH];                                   // the next batch's targets fly while this batch is computed
#pragma unroll
      for (int u = 0; u < HF_BATCH; ++u) tn[u] = load_t(yb + (HF_BATCH + u) * groups);
#pragma unroll
      for (int u = 0; u < HF_BATCH; ++u) {
        const int y = yb + u * groups;
        if (y < y1) {
          const long long off = (long long)y * W + x;
          const int is = taps_s[y] - i_lo;
          const float* wy = tapw_s + 3 * y;
          float4 r0 = *reinterpret_cast<const float4*>(tmp_s + is * W + x);
          float4 z = make_float4(wy[0] * r0.x, wy[0] * r0.y, wy[0] * r0.z, wy[0] * r0.w);
          if (is + 1 < nin) {
            float4 r1 = *reinterpret_cast<const float4*>(tmp_s + (is + 1) * W + x);
            z.x = fmaf(wy[1], r1.x, z.x); z.y = fmaf(wy[1], r1.y, z.y); z.z = fmaf(wy[1], r1.z, z.z); z.w = fmaf(wy[1], r1.w, z.w);
          }
          if (is + 2 < nin) {
            float4 r2 = *reinterpret_cast<const float4*>(tmp_s + (is + 2) * W + x);
            z.x = fmaf(wy[2], r2.x, z.x); z.y = fmaf(wy[2], r2.y, z.y); z.z = fmaf(wy[2], r2.z, z.z); z.w = fmaf(wy[2], r2.w, z.w);
          }
          float4 p;
          if (has_t) {
            float4 t = tv[u];
            if (tmask) {
              const float4 mv = __ldg(reinterpret_cast<const float4*>(tmask + off));
              t.x *= mv.x; t.y *= mv.y; t.z *= mv.z; t.w *= mv.w;
            }
            p = sigmoid_bce4<true>(z, t, loss);
          } else {
            p = sigmoid_bce4<false>(z, z, loss);
          }
          if (STREAM) stg_stream4(heat + off, p); else *reinterpret_cast<float4*>(heat + off) = p;
          if (a.argmax_key) {
            const uint32_t idx = (uint32_t)off;
            if (p.x > best) { best = p.x; best_idx = idx; }
            if (p.y > best) { best = p.y; best_idx = idx + 1; }
            if (p.z > best) { best = p.z; best_idx = idx + 2; }
            if (p.w > best) { best = p.w; best_idx = idx + 3; }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < HF_BATCH; ++u) tv[u] = tn[u];
    }
  }
  loss *= 0.6931471805599453f;   // back to natural log
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (has_t) {
    loss = warp_sum(loss);
    if (lane == 0) red_f[warp] = loss;
  }
  if (a.argmax_key) {
    unsigned long long k = best_idx == 0xFFFFFFFFu ? 0ull : pack_key(best, best_idx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
      k = other > k ? other : k;
    }
    if (lane == 0) red_k[warp] = k;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (has_t) {
      float s = 0.f;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red_f[w];
      a.partials[((long long)h * a.B + b) * gridDim.x + band] = s;
    }
    if (a.argmax_key) {
      unsigned long long k = 0ull;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) k = red_k[w] > k ? red_k[w] : k;
      atomicMax(a.argmax_key + (long long)h * a.B + b, k);
    }
  }
}

// single block.  loss[h] = sum(partials[h]) / count for h < heads; optionally loss[heads] = sum_h loss[h]
// (left-to-right, the order of bifold/losses/__init__.py:84-88).  Fixed assignment + fixed tree => deterministic.
__global__ void __launch_bounds__(256) loss_finalize_kernel(const float* __restrict__ partials, int heads, int n_per_head,
                                                            float inv_count, float* __restrict__ loss, int write_total) {
  __shared__ float red[8];
  float total = 0.f;
  for (int h = 0; h < heads; ++h) {
    float s = 0.f;
    for (int i = threadIdx.x; i < n_per_head; i += 256) s += partials[(long long)h * n_per_head + i];
    s = warp_sum(s);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += red[w];
      t *= inv_count;
      loss[h] = t;
      total = h == 0 ? t : total + t;
    }
  }
  if (threadIdx.x == 0 && write_total) loss[heads] = total;
}

// ---- heatmap backward: dlogit -> dgrid through the transposed composite upsample -----------------------
// Stage 1, grid (bands, heads, B): a CTA streams HM_ROWS rows of (heat, target | dheat) once (float4, coalesced),
// folds them vertically into the <= HM_MAXIN grid rows they touch (registers), then horizontally into grid
// columns (shared memory) and writes one partial [HM_MAXIN][g] block.  Stage 2 sums the bands of a map in
// fixed order (deterministic).  HBM-bound: reads heads*H*W*4 B (x2 with targets) per sample.
constexpr int HB_THREADS = 256;
__global__ void __launch_bounds__(HB_THREADS) heatmap_bwd_stage1(bfk_heatmap_bwd_args a, float* __restrict__ partials) {
  extern __shared__ __align__(16) float sm[];
  const int g = a.g, W = 16 * g, H = W, W4 = W / 4;
  const int band = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int y0 = band * HM_ROWS, y1 = min(H, y0 + HM_ROWS);
  const int i_lo = a.tap_start[y0];
  const int groups = HB_THREADS / W4;                 // row groups working in parallel (>= 2 for W <= 448)
  const int grp = threadIdx.x / W4, x4 = threadIdx.x % W4;
  float* tmp = sm;                                     // [groups][HM_MAXIN][W]
  const float* heat = a.heat + ((long long)h * a.B + b) * H * W;
  const float* tgt = nullptr; const float* tmask = nullptr; const float* dheat = nullptr;
  float gs = 0.f;
  if (a.mode == 0) {
    tgt = a.target[h] + (long long)b * H * W;
    if (a.target_mask && ((a.mask_heads >> h) & 1u)) tmask = a.target_mask + (long long)b * H * W;
    gs = (a.gscale[h] + (a.gscale_has_total ? a.gscale[a.heads] : 0.f)) * a.gscale_mul;
  } else {
    dheat = a.dheat[h] ? a.dheat[h] + (long long)b * H * W : nullptr;
  }
  float4 acc[HM_MAXIN];
#pragma unroll
  for (int ii = 0; ii < HM_MAXIN; ++ii) acc[ii] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (grp < groups) {
    constexpr int RB = 2;   // rows whose loads are issued together
    for (int yb = y0 + grp; yb < y1; yb += groups * RB) {
      float4 pv[RB], qv[RB];
#pragma unroll
      for (int k = 0; k < RB; ++k) {
        const int y = yb + k * groups;
        if (y < y1) {
          const long long off = (long long)y * W + x4 * 4;
          pv[k] = ldg_stream4(heat + off);
          if (a.mode == 0) {
            qv[k] = ldg_stream4(tgt + off);
            if (tmask) {
              const float4 m = ldg_stream4(tmask + off);
              qv[k].x *= m.x; qv[k].y *= m.y; qv[k].z *= m.z; qv[k].w *= m.w;
            }
          } else if (dheat) {
            qv[k] = ldg_stream4(dheat + off);
          }
        }
      }
#pragma unroll
      for (int k = 0; k < RB; ++k) {
      const int y = yb + k * groups;
      if (y >= y1) break;
      const float4 p = pv[k];
      float4 d;
      if (a.mode == 0) {
        const float4 t = qv[k];
        d = make_float4(gs * (p.x - t.x), gs * (p.y - t.y), gs * (p.z - t.z), gs * (p.w - t.w));
      } else if (dheat) {
        const float4 u = qv[k];
        d = make_float4(u.x * p.x * (1.f - p.x), u.y * p.y * (1.f - p.y), u.z * p.z * (1.f - p.z), u.w * p.w * (1.f - p.w));
      } else {
        d = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      const int is = a.tap_start[y] - i_lo;              // uniform across the CTA for a given y: no divergence
      const float w0 = a.tap_w[3 * y], w1 = a.tap_w[3 * y + 1], w2 = a.tap_w[3 * y + 2];
#define BFK_ACC3(I)                                                                                              \
      acc[I].x = fmaf(w0, d.x, acc[I].x); acc[I].y = fmaf(w0, d.y, acc[I].y);                                    \
      acc[I].z = fmaf(w0, d.z, acc[I].z); acc[I].w = fmaf(w0, d.w, acc[I].w);                                    \
      acc[I + 1].x = fmaf(w1, d.x, acc[I + 1].x); acc[I + 1].y = fmaf(w1, d.y, acc[I + 1].y);                    \
      acc[I + 1].z = fmaf(w1, d.z, acc[I + 1].z); acc[I + 1].w = fmaf(w1, d.w, acc[I + 1].w);                    \
      acc[I + 2].x = fmaf(w2, d.x, acc[I + 2].x); acc[I + 2].y = fmaf(w2, d.y, acc[I + 2].y);                    \
      acc[I + 2].z = fmaf(w2, d.z, acc[I + 2].z); acc[I + 2].w = fmaf(w2, d.w, acc[I + 2].w);
      switch (is) {   // taps beyond the grid carry weight 0 and land in accumulator rows that stage 2 never reads
        case 0: BFK_ACC3(0) break;
        case 1: BFK_ACC3(1) break;
        case 2: BFK_ACC3(2) break;
        default: BFK_ACC3(3) break;
      }
#undef BFK_ACC3
      }
    }
#pragma unroll
    for (int ii = 0; ii < HM_MAXIN; ++ii)
      *reinterpret_cast<float4*>(tmp + ((long long)(grp * HM_MAXIN + ii)) * W + x4 * 4) = acc[ii];
  }
  __syncthreads();
  // combine the row groups (fixed order) into group 0
  for (int t = threadIdx.x; t < HM_MAXIN * W; t += HB_THREADS) {
    float v = tmp[t];
    for (int gq = 1; gq < groups; ++gq) v += tmp[(long long)gq * HM_MAXIN * W + t];
    tmp[t] = v;
  }
  __syncthreads();
  // horizontal fold: out[ii][j] = sum_x Cx[x][j] * tmp[ii][x];  x in the support of j
  float* out = partials + ((((long long)b * a.heads + h) * gridDim.x + band) * HM_MAXIN) * g;
  for (int t = threadIdx.x; t < HM_MAXIN * g; t += HB_THREADS) {
    const int ii = t / g, j = t % g;
    const int xlo = max(0, 16 * j - 32), xhi = min(W, 16 * j + 48);
    float sacc = 0.f;
    for (int x = xlo; x < xhi; ++x) {
      const int k = j - a.tap_start[x];
      if (k >= 0 && k < 3) sacc = fmaf(a.tap_w[3 * x + k], tmp[ii * W + x], sacc);
    }
    out[t] = sacc;
  }
}
// grid (heads, B): dgrid[b,h,i,j] = sum over bands of the partial rows that map to grid row i
__global__ void heatmap_bwd_stage2(bfk_heatmap_bwd_args a, const float* __restrict__ partials, int bands) {
  const int g = a.g, H = 16 * g;
  const int h = blockIdx.x, b = blockIdx.y;
  float* out = a.dgrid + ((long long)b * a.heads + h) * g * g;
  for (int t = threadIdx.x; t < g * g; t += blockDim.x) {
    const int i = t / g, j = t % g;
    float s = 0.f;
    for (int band = 0; band < bands; ++band) {
      const int i_lo = a.tap_start[band * HM_ROWS];
      const int ii = i - i_lo;
      if (ii >= 0 && ii < HM_MAXIN) s += partials[((((long long)b * a.heads + h) * bands + band) * HM_MAXIN + ii) * g + j];
    }
    out[t] = s;
  }
  (void)H;
}

// ---- standalone BCE on probabilities ---------------------------------------------------------------------
__global__ void __launch_bounds__(256) bce_prob_stage1(const float* __restrict__ p, const float* __restrict__ t,
                                                       const float* __restrict__ tm, long long n, float* __restrict__ partials) {
  __shared__ float red[8];
  float s = 0.f;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    float pv = p[i], tv = t[i];
    if (tm) tv *= tm[i];
    s -= tv * fmaxf(logf(pv), -100.f) + (1.f - tv) * fmaxf(log1pf(-pv), -100.f);
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = 0.f;
    for (int w = 0; w < 8; ++w) r += red[w];
    partials[blockIdx.x] = r;
  }
}
__global__ void bce_prob_bwd_kernel(const float* __restrict__ p, const float* __restrict__ t, const float* __restrict__ tm,
                                    long long n, const float* __restrict__ gscale, float* __restrict__ dp) {
  const float gsc = gscale[0];
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    float pv = p[i], tv = t[i];
    if (tm) tv *= tm[i];
    dp[i] = gsc * (pv - tv) / fmaxf(pv * (1.f - pv), 1e-12f);
  }
}

// ---- argmax over maps ------------------------------------------------------------------------------------
// grid: (chunks, n_maps)
__global__ void __launch_bounds__(256) argmax2d_kernel(const float* __restrict__ maps, int hw, unsigned long long* __restrict__ key) {
  __shared__ unsigned long long red[8];
  const float* m = maps + (long long)blockIdx.y * hw;
  const int per = (hw + gridDim.x - 1) / gridDim.x;
  const int i0 = blockIdx.x * per, i1 = min(hw, i0 + per);
  float best = -INFINITY; uint32_t bi = 0xFFFFFFFFu;
  for (int i = i0 + threadIdx.x; i < i1; i += 256) {
    float v = m[i];
    if (bi == 0xFFFFFFFFu || v > best) { best = v; bi = i; }
  }
  unsigned long long k = bi == 0xFFFFFFFFu ? 0ull : pack_key(best, bi);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
    k = other > k ? other : k;
  }
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = k;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long r = 0ull;
    for (int w = 0; w < 8; ++w) r = red[w] > r ? red[w] : r;
    atomicMax(key + blockIdx.y, r);
  }
}

// ---- get_action tail: decode, snap to mask, gate -------------------------------------------------------------
// grid: B; block 256
__global__ void __launch_bounds__(256) action_kernel(bfk_action_args a) {
  __shared__ unsigned long long red[8];
  __shared__ int px[8][2];     // per head [row, col]
  __shared__ float conf_s[8];
  const int b = blockIdx.x, H = a.H, W = a.W;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int h = 0; h < a.heads; ++h) {
    unsigned long long k = a.key[(long long)h * a.B + b];
    uint32_t idx = 0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull);
    if (k == 0ull || idx >= (uint32_t)(H * W)) idx = 0;  // all-NaN map
    int u = idx / W, v = idx % W;
    bool is_pick = false;
    for (int t = 0; t < a.n_arms; ++t) is_pick |= (a.pick_heads[t] == h);
    if (is_pick && a.mask) {
      // nearest mask pixel (squared Euclidean), first in row-major order among ties
      const float* m = a.mask + (long long)b * H * W;
      unsigned long long best = ~0ull;
      for (int i = threadIdx.x; i < H * W; i += 256) {
        if (m[i] != 0.f) {
          int du = i / W - u, dv = i % W - v;
          unsigned long long cand = ((unsigned long long)(uint32_t)(du * du + dv * dv) << 32) | (uint32_t)i;
          best = cand < best ? cand : best;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
      }
      if (lane == 0) red[warp] = best;
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned long long r = ~0ull;
        for (int w = 0; w < 8; ++w) r = red[w] < r ? red[w] : r;
        if (r != ~0ull) { uint32_t i = (uint32_t)(r & 0xFFFFFFFFull); u = i / W; v = i % W; }
        px[h][0] = u; px[h][1] = v;
        conf_s[h] = a.heat[((long long)h * a.B + b) * H * W + (long long)u * W + v];
      }
      __syncthreads();
    } else if (threadIdx.x == 0) {
      px[h][0] = u; px[h][1] = v;
      conf_s[h] = a.heat[((long long)h * a.B + b) * H * W + (long long)u * W + v];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    bool keep[2] = {true, true};
    if (a.n_arms == 2) {
      float c0 = conf_s[a.pick_heads[0]], c1 = conf_s[a.pick_heads[1]];
      int am = c1 > c0 ? 1 : 0;  // np.argmax: first max
      keep[0] = c0 >= a.threshold || am == 0;
      keep[1] = c1 >= a.threshold || am == 1;
    }
    for (int h = 0; h < a.heads; ++h) {
      bool k = true;
      for (int t = 0; t < a.n_arms; ++t)
        if (a.pick_heads[t] == h || a.place_heads[t] == h) k = keep[t];
      int* o = a.out_xy + ((long long)h * a.B + b) * 2;
      o[0] = k ? px[h][1] : -1;   // x = column
      o[1] = k ? px[h][0] : -1;   // y = row
      a.out_conf[(long long)h * a.B + b] = conf_s[h];
    }
  }
}

}  // namespace bfk

using namespace bfk;

template <typename K>
static int opt_in_smem(K kernel, size_t bytes, const char* name) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("%s: smem %zu: %s", name, bytes, cudaGetErrorString(e)); return BFK_E_CUDA; }
  }
  return BFK_OK;
}

extern "C" int32_t bfk_decoder_collapse_fwd(const bfk_decoder_collapse_args* a, bfk_stream_t s_) {
  BFK_CHECK_ARG(a && a->W && a->b && a->w_eff && a->b_eff && a->chain && a->heads > 0 && a->D >= 4, "bfk_decoder_collapse_fwd: bad arguments");
  collapse_fwd_kernel<<<a->heads, 1024, 5 * a->D * sizeof(float), to_stream(s_)>>>(*a);
  BFK_CHECK_LAUNCH("bfk_decoder_collapse_fwd");
  return BFK_OK;
}
extern "C" int32_t bfk_decoder_collapse_bwd(const bfk_decoder_collapse_args* a, bfk_stream_t s_) {
  BFK_CHECK_ARG(a && a->W && a->b && a->chain && a->d_w_eff && a->d_b_eff && a->dW && a->db, "bfk_decoder_collapse_bwd: bad arguments");
  collapse_bwd_kernel<<<a->heads, 1024, 6 * a->D * sizeof(float), to_stream(s_)>>>(*a);
  BFK_CHECK_LAUNCH("bfk_decoder_collapse_bwd");
  return BFK_OK;
}

extern "C" int32_t bfk_head_project_fwd(const bfk_head_project_args* a, bfk_stream_t s_) {
  BFK_CHECK_ARG(a && a->X && a->w_eff && a->b_eff && a->grid && a->heads > 0 && a->heads <= 8 && a->D % 4 == 0, "bfk_head_project_fwd: bad arguments");
  size_t smem = (size_t)a->heads * a->D * sizeof(float);
  if (int rc = opt_in_smem(head_project_fwd_kernel, smem, "head_project_fwd")) return rc;
  long long toks = (long long)a->B * a->Np;
  head_project_fwd_kernel<<<(unsigned)((toks + 8 * HP_TOK_PER_WARP - 1) / (8 * HP_TOK_PER_WARP)), 256, smem, to_stream(s_)>>>(*a);
  BFK_CHECK_LAUNCH("bfk_head_project_fwd");
  return BFK_OK;
}
extern "C" size_t bfk_head_project_bwd_workspace_floats(int32_t B, int32_t heads, int32_t Np, int32_t D) {
  long long chunks = ((long long)B * Np + HP_CHUNK - 1) / HP_CHUNK;
  return (size_t)chunks * heads * D;
}
extern "C" int32_t bfk_head_project_bwd(const bfk_head_project_args* a, bfk_stream_t s_) {
  BFK_CHECK_ARG(a && a->X && a->w_eff && a->dgrid && a->dX && a->d_w_eff && a->d_b_eff && a->partials, "bfk_head_project_bwd: null pointer");
  BFK_CHECK_ARG(a->heads > 0 && a->heads <= 8 && a->D % 4 == 0 && a->Nf >= a->row0 + a->Np, "bfk_head_project_bwd: bad shape");
  cudaStream_t s = to_stream(s_);
  int thr = std::min(256, std::max(32, a->D / 4));
  if (a->dX2 && a->dX2_dtype == BFK_BF16) head_project_bwd_dx_kernel<__nv_bfloat16><<<dim3(a->Nf, a->B), thr, 0, s>>>(*a);
  else head_project_bwd_dx_kernel<float><<<dim3(a->Nf, a->B), thr, 0, s>>>(*a);
  int chunks = (int)(((long long)a->B * a->Np + HP_CHUNK - 1) / HP_CHUNK);
  head_project_bwd_dw_stage1<<<dim3((a->D + 127) / 128, chunks), 128, 0, s>>>(*a);
  head_project_bwd_dw_stage2<<<(a->heads * a->D + 127) / 128, 128, 0, s>>>(*a, chunks);
  head_project_bwd_db_kernel<<<a->heads, 256, 0, s>>>(*a);
  BFK_CHECK_LAUNCH("bfk_head_project_bwd");
  return BFK_OK;
}

extern "C" int32_t bfk_upsample16_taps(int32_t g, int32_t* tap_start, float* tap_w) {
  BFK_CHECK_ARG(g > 0 && tap_start && tap_w, "bfk_upsample16_taps: bad arguments");
  // compose four x2 bilinear stages (align_corners=False) as dense row operators in double
  std::vector<std::vector<double>> cur(g, std::vector<double>(g, 0.0));
  for (int i = 0; i < g; ++i) cur[i][i] = 1.0;
  int n = g;
  for (int stage = 0; stage < 4; ++stage) {
    std::vector<std::vector<double>> nxt(2 * n, std::vector<double>(g, 0.0));
    for (int d = 0; d < 2 * n; ++d) {
      double src = (d + 0.5) / 2.0 - 0.5;
      if (src < 0) src = 0;
      int i0 = (int)floor(src);
      int i1 = i0 + 1 < n ? i0 + 1 : n - 1;
      double l = src - i0;
      for (int c = 0; c < g; ++c) nxt[d][c] = (1.0 - l) * cur[i0][c] + l * cur[i1][c];
    }
    cur.swap(nxt);
    n *= 2;
  }
  for (int y = 0; y < n; ++y) {
    int first = -1, last = -1;
    for (int c = 0; c < g; ++c)
      if (cur[y][c] != 0.0) { if (first < 0) first = c; last = c; }
    if (first < 0 || last - first > 2) { set_error("bfk_upsample16_taps: row %d needs %d taps", y, last - first + 1); return BFK_E_UNSUPPORTED; }
    tap_start[y] = first;
    for (int t = 0; t < 3; ++t) tap_w[3 * y + t] = (first + t < g) ? (float)cur[y][first + t] : 0.f;
  }
  return BFK_OK;
}

extern "C" int32_t bfk_heatmap_bands(int32_t g) { return (16 * g + HF_ROWS - 1) / HF_ROWS; }     // forward bands (loss partials)
static int heatmap_bwd_bands(int g) { return (16 * g + HM_ROWS - 1) / HM_ROWS; }

extern "C" int32_t bfk_heatmap_fwd(const bfk_heatmap_fwd_args* a, bfk_stream_t s_) {
  BFK_CHECK_ARG(a && a->grid && a->tap_start && a->tap_w && a->heat && a->B > 0 && a->heads > 0 && a->g > 0, "bfk_heatmap_fwd: bad arguments");
  BFK_CHECK_ARG(!a->target || (a->loss && a->partials), "bfk_heatmap_fwd: target needs loss and partials");
  cudaStream_t s = to_stream(s_);
  const int W = 16 * a->g, bands = bfk_heatmap_bands(a->g);
  if (a->argmax_key) cudaMemsetAsync(a->argmax_key, 0, sizeof(unsigned long long) * a->heads * a->B, s);
  size_t smem = ((size_t)HF_MAXIN * W + 3 * (size_t)W + (size_t)W + (size_t)a->g * a->g) * sizeof(float);
  bfk_heatmap_fwd_args k = *a;
  static int variant = -1;
  if (variant < 0) { const char* e = getenv("BFK_HM_VARIANT"); variant = e ? atoi(e) : 0; }
  dim3 grid(bands, a->heads, a->B);
  const int threads = ((HM_THREADS / (W / 4)) * (W / 4) + 31) / 32 * 32;   // whole rows of float4 columns (+ idle lanes to a full warp)
  if (variant == 1) heatmap_fwd_kernel<true><<<grid, threads, smem, s>>>(k);    // streaming (L1 no-allocate) accesses
  else heatmap_fwd_kernel<false><<<grid, threads, smem, s>>>(k);
  if (a->target) {
    float inv = 1.f / ((float)a->B * (float)W * (float)W);
    loss_finalize_kernel<<<1, 256, 0, s>>>(a->partials, a->heads, a->B * bands, inv, a->loss, 1);
  }
  BFK_CHECK_LAUNCH("bfk_heatmap_fwd");
  return BFK_OK;
}

extern "C" size_t bfk_heatmap_bwd_workspace_floats(int32_t B, int32_t heads, int32_t g) {
  return (size_t)B * heads * heatmap_bwd_bands(g) * HM_MAXIN * g;
}

extern "C" int32_t bfk_heatmap_bwd(const bfk_heatmap_bwd_args* a, bfk_stream_t s_) {
  BFK_CHECK_ARG(a && a->heat && a->tap_start && a->tap_w && a->dgrid && a->partials && a->g > 0 && 16 * a->g <= 448,
                "bfk_heatmap_bwd: bad arguments (W <= 448, partials required)");
  BFK_CHECK_ARG(a->mode == 0 ? (a->target && a->gscale) : (a->dheat != nullptr), "bfk_heatmap_bwd: mode inputs missing");
  const int W = 16 * a->g, bands = heatmap_bwd_bands(a->g);
  const int groups = HB_THREADS / (W / 4);
  size_t smem = (size_t)groups * HM_MAXIN * W * sizeof(float);
  if (int rc = opt_in_smem(heatmap_bwd_stage1, smem, "heatmap_bwd")) return rc;
  cudaStream_t s = to_stream(s_);
  heatmap_bwd_stage1<<<dim3(bands, a->heads, a->B), HB_THREADS, smem, s>>>(*a, a->partials);
  heatmap_bwd_stage2<<<dim3(a->heads, a->B), 256, 0, s>>>(*a, a->partials, bands);
  BFK_CHECK_LAUNCH("bfk_heatmap_bwd");
  return BFK_OK;
}

extern "C" int32_t bfk_bce_prob(const float* p, const float* t, const float* t_mask, int64_t n, float* loss,
                                float* partials, bfk_stream_t s_) {
  BFK_CHECK_ARG(p && t && loss && partials && n > 0, "bfk_bce_prob: bad arguments");
  cudaStream_t s = to_stream(s_);
  int blocks = (int)std::min<long long>(1024, (n + 255) / 256);
  bce_prob_stage1<<<blocks, 256, 0, s>>>(p, t, t_mask, n, partials);
  loss_finalize_kernel<<<1, 256, 0, s>>>(partials, 1, blocks, 1.f / (float)n, loss, 0);
  BFK_CHECK_LAUNCH("bfk_bce_prob");
  return BFK_OK;
}
extern "C" int32_t bfk_bce_prob_bwd(const float* p, const float* t, const float* t_mask, int64_t n, const float* gscale,
                                    float* dp, bfk_stream_t s_) {
  BFK_CHECK_ARG(p && t && gscale && dp && n > 0, "bfk_bce_prob_bwd: bad arguments");
  int blocks = (int)std::min<long long>((long long)sm_count() * 8, (n + 255) / 256);
  bce_prob_bwd_kernel<<<blocks, 256, 0, to_stream(s_)>>>(p, t, t_mask, n, gscale, dp);
  BFK_CHECK_LAUNCH("bfk_bce_prob_bwd");
  return BFK_OK;
}

extern "C" int32_t bfk_argmax2d(const float* maps, int32_t n_maps, int32_t hw, unsigned long long* key, bfk_stream_t s_) {
  BFK_CHECK_ARG(maps && key && n_maps > 0 && hw > 0, "bfk_argmax2d: bad arguments");
  cudaStream_t s = to_stream(s_);
  cudaMemsetAsync(key, 0, sizeof(unsigned long long) * n_maps, s);
  int chunks = std::max(1, std::min(64, hw / 4096));
  argmax2d_kernel<<<dim3(chunks, n_maps), 256, 0, s>>>(maps, hw, key);
  BFK_CHECK_LAUNCH("bfk_argmax2d");
  return BFK_OK;
}

extern "C" int32_t bfk_action(const bfk_action_args* a, bfk_stream_t s_) {
  BFK_CHECK_ARG(a && a->heat && a->key && a->out_xy && a->out_conf && a->heads > 0 && a->heads <= 8 && (a->n_arms == 1 || a->n_arms == 2), "bfk_action: bad arguments");
  action_kernel<<<a->B, 256, 0, to_stream(s_)>>>(*a);
  BFK_CHECK_LAUNCH("bfk_action");
  return BFK_OK;
}
