// Heatmap tail: collapsed ConvDecoder (bifold/models/modules.py:33-74 is linear), composite
// x16 bilinear upsample, sigmoid, nn.BCELoss, first-max argmax and the get_action gate.
// All HBM-bound: the forward heatmap kernel writes heads*H*W*4 B per sample (+ reads as much target).
#include <math.h>
#include <algorithm>
#include <map>
#include <mutex>
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
__device__ __forceinline__ float lg2f_fast(float x) {     // normal positive inputs only (no denormal rescue code)
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
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

// Fast path (D = 128 NV, heads <= 4): persistent warps, w_eff held in REGISTERS (4 heads x NV float4 per lane), the next
// token's row prefetched while the current one is reduced: NV (x2) 16-byte loads in flight per lane.
template <int NV, bool PREFETCH>
__global__ void __launch_bounds__(128, PREFETCH ? 3 : 2) head_project_fwd_v2_kernel(bfk_head_project_args a) {
  const int lane = threadIdx.x & 31;
  const long long gw = (long long)blockIdx.x * 4 + (threadIdx.x >> 5), nw = (long long)gridDim.x * 4;
  const long long toks = (long long)a.B * a.Np;
  float4 w[4][NV];
#pragma unroll
  for (int h = 0; h < 4; ++h)
#pragma unroll
    for (int v = 0; v < NV; ++v)
      w[h][v] = h < a.heads ? __ldg(reinterpret_cast<const float4*>(a.w_eff + (long long)h * a.D) + v * 32 + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
  float be[4];
#pragma unroll
  for (int h = 0; h < 4; ++h) be[h] = h < a.heads ? a.b_eff[h] : 0.f;
  auto row_of = [&](long long tok) {
    const int b = (int)(tok / a.Np), n = (int)(tok % a.Np);
    return a.X + (long long)b * a.x_batch_stride + (long long)(a.row0 + n) * a.D + lane * 4;
  };
  float4 x[NV], xn[NV];
  long long tok = gw;
  if (tok < toks) {
    const float* r = row_of(tok);
#pragma unroll
    for (int v = 0; v < NV; ++v) x[v] = ldg_stream4(r + v * 128);
  }
  for (; tok < toks; tok += nw) {
    const long long nxt = tok + nw;
    if (PREFETCH && nxt < toks) {
      const float* r = row_of(nxt);
#pragma unroll
      for (int v = 0; v < NV; ++v) xn[v] = ldg_stream4(r + v * 128);
    }
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int h = 0; h < 4; ++h)
        acc[h] = fmaf(x[v].x, w[h][v].x, fmaf(x[v].y, w[h][v].y, fmaf(x[v].z, w[h][v].z, fmaf(x[v].w, w[h][v].w, acc[h]))));
#pragma unroll
    for (int h = 0; h < 4; ++h) acc[h] = warp_sum(acc[h]);
    const int b = (int)(tok / a.Np), n = (int)(tok % a.Np);
    if (lane < a.heads) {
      const float s = lane == 0 ? acc[0] : lane == 1 ? acc[1] : lane == 2 ? acc[2] : acc[3];
      const float bb = lane == 0 ? be[0] : lane == 1 ? be[1] : lane == 2 ? be[2] : be[3];
      a.grid[((long long)b * a.heads + lane) * a.Np + n] = s + bb;
    }
    if (PREFETCH) {
#pragma unroll
      for (int v = 0; v < NV; ++v) x[v] = xn[v];
    } else if (nxt < toks) {
      const float* r = row_of(nxt);
#pragma unroll
      for (int v = 0; v < NV; ++v) x[v] = ldg_stream4(r + v * 128);
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

// ---- heatmap forward: x16 upsample + sigmoid (+ BCE, + argmax) --------------------------------------------
// A map (H = W = 16 g) is cut into bands of 16 g / bands rows (default 2 bands: measured best of 1/2/4/8 at B = 64 and 512); one CTA per (band, head, sample).  Threads own a
// fixed float4 column and walk down the rows of the band, so the set-up (taps, grid, horizontal pass) is paid
// once per 4 g rows.  The last CTA to finish (device counter, self-resetting) reduces the per-band BCE partial
// sums and argmax keys in a FIXED order: one launch, deterministic results.
constexpr int HM_BANDS_DEFAULT = 2;   // forward bands per map (BFK_HM_BANDS overrides; must divide 16)
constexpr int HM_THREADS = 256;
constexpr int HF_BATCH = 4;   // rows of one thread whose target loads are in flight together
constexpr int HM_RED = 128;   // logical width of the final loss reduction (independent of the block size)

// nn.BCELoss computes -[t log p + (1-t) log(1-p)] from the ROUNDED p with both logs clamped at -100.  The closed (softplus) form
// below equals it up to the rounding of 1 - p, i.e. to <= 6e-8 / (1 - p) absolute: below 3e-5 for z <= BCE_ZFAST.  Above that (and
// for very negative z) torch's result is dominated by that rounding and by its clamp, so those pixels take the literal path.
constexpr float BCE_ZFAST = 6.f, BCE_ZLOW = -40.f;
// four pixels: probabilities and (LOSS) the sum of their BCE terms in units of ln 2; ONE (almost never divergent) branch per four.
//   e = 2^(-z log2 e) = exp(-z);  p = 1 / (1 + e)  (e = inf for z < -88: p = 0);  BCE = (1 - t) z + log(1 + e)    [softplus form]
// For z in [BCE_ZLOW, BCE_ZFAST] the four 1 + e are <= 2.4e17, so two products of two stay finite: two lg2 for four pixels.
template <bool LOSS>
__device__ __forceinline__ float4 sigmoid_bce4(const float4& z, const float4& t, float& loss_l2) {
  const float kl = 1.4426950408889634f;
  const float dx = 1.f + exp2f_fast(-z.x * kl), dy = 1.f + exp2f_fast(-z.y * kl), dz = 1.f + exp2f_fast(-z.z * kl), dw = 1.f + exp2f_fast(-z.w * kl);
  const float4 p = make_float4(__fdividef(1.f, dx), __fdividef(1.f, dy), __fdividef(1.f, dz), __fdividef(1.f, dw));
  if (LOSS) {
    const float zmax = fmaxf(fmaxf(z.x, z.y), fmaxf(z.z, z.w)), zmin = fminf(fminf(z.x, z.y), fminf(z.z, z.w));
    if (zmax <= BCE_ZFAST && zmin >= BCE_ZLOW) {
      const float lin = fmaf(1.f - t.x, z.x, fmaf(1.f - t.y, z.y, fmaf(1.f - t.z, z.z, (1.f - t.w) * z.w)));
      loss_l2 += fmaf(lin, kl, lg2f_fast(dx * dy) + lg2f_fast(dz * dw));
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

// taps of output index i as one 16-byte smem word: {w0, w1, w2, bits(first input index)}
__device__ __forceinline__ void stage_taps(float4* tap_s, const int32_t* __restrict__ tap_start, const float* __restrict__ tap_w, int W) {
  for (int i = threadIdx.x; i < W; i += blockDim.x)
    tap_s[i] = make_float4(tap_w[3 * i], tap_w[3 * i + 1], tap_w[3 * i + 2], __int_as_float(tap_start[i]));
}
// whole-block sum of one value per logical lane < HM_RED (fixed tree: identical for every block size >= HM_RED)
__device__ __forceinline__ float red128(float s, float* red_s) {
  s = warp_sum(s);
  __syncthreads();
  if (threadIdx.x < HM_RED && (threadIdx.x & 31) == 0) red_s[threadIdx.x >> 5] = s;
  __syncthreads();
  return (red_s[0] + red_s[1]) + (red_s[2] + red_s[3]);
}
__device__ __forceinline__ unsigned long long* hm_key_partials(float* partials, int n_maps, int bands) {
  return reinterpret_cast<unsigned long long*>(partials + (((long long)n_maps * bands + 1) & ~1ll));
}

// Final reduction over the bands (all threads of ONE block): loss[h] = sum(partials[h]) / count, loss[heads] = their sum in
// head order (bifold/losses/__init__.py:84-88); argmax_key[m] = max over the bands of the per-band keys.
__device__ void hm_finalize(const bfk_heatmap_fwd_args& a, int bands, bool loss, bool keys, float* red_s) {
  const int n_maps = a.heads * a.B;
  if (loss) {
    const int n = a.B * bands;
    const float W = 16.f * (float)a.g, inv = 1.f / ((float)a.B * W * W);
    float total = 0.f;
    for (int h = 0; h < a.heads; ++h) {
      float s = 0.f;
      if (threadIdx.x < HM_RED)
        for (int i = threadIdx.x; i < n; i += HM_RED) s += __ldcg(a.partials + (long long)h * n + i);
      const float t = red128(s, red_s) * inv;
      if (threadIdx.x == 0) a.loss[h] = t;
      total = h == 0 ? t : total + t;
    }
    if (threadIdx.x == 0) a.loss[a.heads] = total;
  }
  if (keys) {
    const unsigned long long* kp = hm_key_partials(a.partials, n_maps, bands);
    for (int m = threadIdx.x; m < n_maps; m += blockDim.x) {
      unsigned long long k = 0ull;
      for (int q = 0; q < bands; ++q) { const unsigned long long o = __ldcg(kp + (long long)m * bands + q); k = o > k ? o : k; }
      a.argmax_key[m] = k;
    }
  }
}
__global__ void __launch_bounds__(HM_RED) heatmap_finalize_kernel(bfk_heatmap_fwd_args a, int bands, int loss, int keys) {
  __shared__ float red_s[4];
  hm_finalize(a, bands, loss != 0, keys != 0, red_s);
}

// grid: (bands, heads, B).  smem: taps [W] float4, tmp [(nin_max + 2)][W], grid [g*g]
template <bool LOSS, bool MASK, bool ARGMAX, int MINB>
__global__ void __launch_bounds__(HM_THREADS, MINB) heatmap_fwd_kernel(bfk_heatmap_fwd_args a, int nin_max) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float red_f[HM_THREADS / 32];
  __shared__ unsigned long long red_k[HM_THREADS / 32];
  __shared__ int last_s;
  const int g = a.g, W = 16 * g, bands = gridDim.x, R = W / bands;
  const int band = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int y0 = band * R, y1 = y0 + R;
  const int W4 = W / 4;
  const int groups = blockDim.x / W4;                 // rows worked on in parallel; thread -> fixed float4 column
  const int grp = threadIdx.x / W4, x = (threadIdx.x % W4) * 4;
  const bool active = grp < groups;
  const long long map_off = (long long)b * W * W;
  const float* tgt = LOSS ? a.target[h] + map_off : nullptr;
  const bool masked = MASK && ((a.mask_heads >> h) & 1u);
  const float* tmask = a.target_mask + map_off;        // dereferenced only if `masked`
  // thread-relative addressing: 32-bit element offsets from per-thread base pointers (row r of this thread = r * stride)
  const int nrows = active ? (R - grp + groups - 1) / groups : 0;
  const int stride = groups * W, off0 = (y0 + grp) * W + x;
  const float* tg = tgt + off0;
  // the targets do not depend on anything this CTA computes: their loads go out FIRST and fly during the set-up below
  float4 tv[HF_BATCH];
  if (LOSS) {
#pragma unroll
    for (int u = 0; u < HF_BATCH; ++u) tv[u] = u < nrows ? ldg_stream4(tg + u * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float4* tap_s = reinterpret_cast<float4*>(sm);        // [W]
  float* tmp_s = sm + 4 * W;                            // [(nin_max + 2)][W]
  float* grid_s = tmp_s + (nin_max + 2) * W;            // [g][g]
  const float* gsrc = a.grid + ((long long)b * a.heads + h) * g * g;
  for (int i = threadIdx.x; i < g * g; i += blockDim.x) grid_s[i] = gsrc[i];
  stage_taps(tap_s, a.tap_start, a.tap_w, W);
  __syncthreads();
  const int i_lo = __float_as_int(tap_s[y0].w);
  const int nin = min(g, __float_as_int(tap_s[y1 - 1].w) + 3) - i_lo;     // <= nin_max (checked on the host)
  // horizontal pass for the input rows this band needs (+ two zero rows: taps past the grid carry weight 0)
  for (int xx = threadIdx.x; xx < W; xx += blockDim.x) {
    const float4 tp = tap_s[xx];
    const int js = __float_as_int(tp.w), j1 = min(js + 1, g - 1), j2 = min(js + 2, g - 1);
    for (int ii = 0; ii < nin; ++ii) {
      const float* row = grid_s + (i_lo + ii) * g;
      tmp_s[ii * W + xx] = fmaf(tp.z, row[j2], fmaf(tp.y, row[j1], tp.x * row[js]));
    }
    tmp_s[nin * W + xx] = 0.f;
    tmp_s[(nin + 1) * W + xx] = 0.f;
  }
  __syncthreads();
  float* hp = a.heat + ((long long)h * a.B + b) * W * W + off0;
  const float4* tp = tap_s + y0 + grp;
  const float* tmx = tmp_s + x - i_lo * W;       // + (first input row of an output row) * W
  const float* tmk = tmask + off0;
  float loss = 0.f;            // in units of ln 2
  float best = -INFINITY;
  uint32_t best_idx = 0xFFFFFFFFu;
  float4 tn[HF_BATCH];
  // one batch of HF_BATCH rows: the other buffer's targets (rows r + HF_BATCH ..) are requested first and fly during the maths
  auto batch = [&](int r, const float4 (&tc)[HF_BATCH], float4 (&tnext)[HF_BATCH]) {
    if (LOSS) {
#pragma unroll
      for (int u = 0; u < HF_BATCH; ++u)
        tnext[u] = r + HF_BATCH + u < nrows ? ldg_stream4(tg + (r + HF_BATCH + u) * stride) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < HF_BATCH; ++u) {
      if (r + u < nrows) {
        const int o = (r + u) * stride;
        const float4 wy = tp[(r + u) * groups];
        const float* rp = tmx + __float_as_int(wy.w) * W;
        const float4 r0 = *reinterpret_cast<const float4*>(rp);
        const float4 r1 = *reinterpret_cast<const float4*>(rp + W);
        const float4 r2 = *reinterpret_cast<const float4*>(rp + 2 * W);
        float4 z;
        z.x = fmaf(wy.z, r2.x, fmaf(wy.y, r1.x, wy.x * r0.x));
        z.y = fmaf(wy.z, r2.y, fmaf(wy.y, r1.y, wy.x * r0.y));
        z.z = fmaf(wy.z, r2.z, fmaf(wy.y, r1.z, wy.x * r0.z));
        z.w = fmaf(wy.z, r2.w, fmaf(wy.y, r1.w, wy.x * r0.w));
        float4 t = LOSS ? tc[u] : z;
        if (masked) {
          const float4 mv = __ldg(reinterpret_cast<const float4*>(tmk + o));
          t.x *= mv.x; t.y *= mv.y; t.z *= mv.z; t.w *= mv.w;
        }
        const float4 p = sigmoid_bce4<LOSS>(z, t, loss);
        *reinterpret_cast<float4*>(hp + o) = p;
        if (ARGMAX) {
          const float m = fmaxf(fmaxf(p.x, p.y), fmaxf(p.z, p.w));
          if (m > best) {      // first maximum: strict, rows visited in increasing order
            best = m;
            best_idx = (uint32_t)(off0 + o) + (p.x == m ? 0u : p.y == m ? 1u : p.z == m ? 2u : 3u);
          }
        }
      }
    }
  };
  for (int r = 0; r < nrows; r += 2 * HF_BATCH) {
    batch(r, tv, tn);
    if (r + HF_BATCH < nrows) batch(r + HF_BATCH, tn, tv);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  if (LOSS) {
    loss = warp_sum(loss * 0.6931471805599453f);   // back to natural log
    if (lane == 0) red_f[warp] = loss;
  }
  if (ARGMAX) {
    unsigned long long k = best_idx == 0xFFFFFFFFu ? 0ull : pack_key(best, best_idx);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, k, o);
      k = other > k ? other : k;
    }
    if (lane == 0) red_k[warp] = k;
  }
  __syncthreads();
  const long long map = (long long)h * a.B + b;
  if (threadIdx.x == 0) {
    if (LOSS) {
      float s = 0.f;
      for (int w = 0; w < nwarp; ++w) s += red_f[w];
      a.partials[map * bands + band] = s;
    }
    if (ARGMAX) {
      unsigned long long k = 0ull;
      for (int w = 0; w < nwarp; ++w) k = red_k[w] > k ? red_k[w] : k;
      if (a.partials) hm_key_partials(a.partials, a.heads * a.B, bands)[map * bands + band] = k;
      else atomicMax(a.argmax_key + map, k);             // no workspace: keys pre-zeroed by the launcher
    }
    last_s = 0;
    if (a.counters) {
      __threadfence();
      const unsigned total = gridDim.x * gridDim.y * gridDim.z;
      last_s = atomicAdd(reinterpret_cast<unsigned*>(a.counters), 1u) == total - 1u;
    }
  }
  __syncthreads();
  if (last_s) {        // every other CTA's partials are visible: reduce them here instead of in a second launch
    __threadfence();
    hm_finalize(a, bands, LOSS, ARGMAX && a.partials != nullptr, red_f);
    if (threadIdx.x == 0) a.counters[0] = 0;
  }
}

// ---- heatmap backward: dlogit -> dgrid through the transposed composite upsample -----------------------
// grid (bands, heads, B): a CTA streams 16 g / bands rows of (heat, target | dheat) once (float4, coalesced), folds them
// vertically into the grid rows they touch (a sliding window of three accumulator rows in registers, flushed to
// shared memory as the window moves), combines the row groups, folds horizontally through a transposed tap table
// and writes one partial [nin][g] block.  The last CTA of a map (per-map device counter, self-resetting) sums the
// bands in fixed order (deterministic).  HBM-bound: reads heads*H*W*4 B (x2 with targets) per sample.
constexpr int HB_BANDS_DEFAULT = 1;  // backward bands per map (BFK_HB_BANDS overrides; measured best of 1/2/4): the folds after the streaming loop are paid once per band
constexpr int HB_SPAN = 80;   // output indices that can touch one grid column: [16 j - 32, 16 j + 48)

__device__ void hb_sum_bands(const bfk_heatmap_bwd_args& a, const float* __restrict__ partials, int nin_max, int bands, int b, int h,
                             const float4* tap_s) {
  const int g = a.g, R = 16 * g / bands;
  float* out = a.dgrid + ((long long)b * a.heads + h) * g * g;
  for (int t = threadIdx.x; t < g * g; t += blockDim.x) {
    const int i = t / g, j = t % g;
    float s = 0.f;
    for (int band = 0; band < bands; ++band) {
      const int i_lo = __float_as_int(tap_s[band * R].w);
      const int nin = min(g, __float_as_int(tap_s[band * R + R - 1].w) + 3) - i_lo;
      const int ii = i - i_lo;
      if (ii >= 0 && ii < nin)
        s += __ldcg(partials + ((((long long)b * a.heads + h) * bands + band) * nin_max + ii) * g + j);
    }
    out[t] = s;
  }
}
__global__ void heatmap_bwd_stage2(bfk_heatmap_bwd_args a, int nin_max, int bands) {
  extern __shared__ __align__(16) float sm[];
  float4* tap_s = reinterpret_cast<float4*>(sm);
  stage_taps(tap_s, a.tap_start, a.tap_w, 16 * a.g);
  __syncthreads();
  hb_sum_bands(a, a.partials, nin_max, bands, blockIdx.y, blockIdx.x, tap_s);
}

// smem: taps [W] float4, wT [g][HB_SPAN], tmp [groups][nin_max][W]
template <int MODE, bool MASK>
__global__ void __launch_bounds__(HM_THREADS, 3) heatmap_bwd_stage1(bfk_heatmap_bwd_args a, int nin_max) {
  extern __shared__ __align__(16) float sm[];
  __shared__ int last_s;
  const int g = a.g, W = 16 * g, bands = gridDim.x, R = W / bands, W4 = W / 4;
  const int band = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int y0 = band * R, y1 = y0 + R;
  const int groups = blockDim.x / W4;
  const int grp = threadIdx.x / W4, x = (threadIdx.x % W4) * 4;
  const bool active = grp < groups;
  float4* tap_s = reinterpret_cast<float4*>(sm);        // [W]
  float* wT = sm + 4 * W;                               // [g][HB_SPAN]
  float* tmp = wT + g * HB_SPAN;                        // [groups][nin_max][W]
  const long long map_off = (long long)b * W * W;
  const float* heat = a.heat + ((long long)h * a.B + b) * W * W;
  const float* src = nullptr;                           // target (MODE 0) or upstream gradient (MODE 1)
  float gs = 0.f;
  if (MODE == 0) {
    src = a.target[h] + map_off;
    gs = (a.gscale[h] + (a.gscale_has_total ? a.gscale[a.heads] : 0.f)) * a.gscale_mul;
  } else {
    src = a.dheat[h] ? a.dheat[h] + map_off : nullptr;
  }
  const bool masked = MASK && MODE == 0 && ((a.mask_heads >> h) & 1u);
  const float* tmask = a.target_mask + map_off;
  // thread-relative addressing: 32-bit element offsets from per-thread base pointers (row r of this thread = r * stride)
  const int nrows = (active && src != nullptr) ? (R - grp + groups - 1) / groups : 0;
  const int stride = groups * W, off0 = (y0 + grp) * W + x;
  const float* hp = heat + off0;
  const float* sp = src + off0;
  const float* tmk = tmask + off0;
  constexpr int RB = 2;   // rows whose loads are in flight together (x2: the next batch is requested before this one is used)
  float4 pv[RB], qv[RB];
#pragma unroll
  for (int k = 0; k < RB; ++k) {                         // in flight during the set-up
    if (k < nrows) { pv[k] = ldg_stream4(hp + k * stride); qv[k] = ldg_stream4(sp + k * stride); }
    else { pv[k] = make_float4(0.f, 0.f, 0.f, 0.f); qv[k] = pv[k]; }
  }
  stage_taps(tap_s, a.tap_start, a.tap_w, W);
  __syncthreads();
  // transposed taps: wT[j][q] = weight of output index x = 16 j - 32 + q on grid column j (0 outside its support):
  // zero the table, then every output index scatters its (at most three) weights
  for (int t = threadIdx.x; t < g * HB_SPAN; t += blockDim.x) wT[t] = 0.f;
  __syncthreads();
  for (int xx = threadIdx.x; xx < W; xx += blockDim.x) {
    const float4 tp = tap_s[xx];
    const int js = __float_as_int(tp.w);
    const int q = xx + 32 - 16 * js;                       // position of xx in the span of column js
    if (q >= 0 && q < HB_SPAN) wT[js * HB_SPAN + q] = tp.x;
    if (js + 1 < g && q - 16 >= 0 && q - 16 < HB_SPAN) wT[(js + 1) * HB_SPAN + q - 16] = tp.y;
    if (js + 2 < g && q - 32 >= 0 && q - 32 < HB_SPAN) wT[(js + 2) * HB_SPAN + q - 32] = tp.z;
  }
  const int i_lo = __float_as_int(tap_s[y0].w);
  const int nin = min(g, __float_as_int(tap_s[y1 - 1].w) + 3) - i_lo;
  if (active) {
    float* mine = tmp + grp * nin_max * W + x;
    const float4* tp = tap_s + y0 + grp;
    float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0;
    int cur = 0;                                         // window = accumulator rows cur, cur+1, cur+2
    for (int r = 0; r < nrows; r += RB) {
      float4 pn[RB], qn[RB];
#pragma unroll
      for (int k = 0; k < RB; ++k) {
        if (r + RB + k < nrows) { pn[k] = ldg_stream4(hp + (r + RB + k) * stride); qn[k] = ldg_stream4(sp + (r + RB + k) * stride); }
        else { pn[k] = make_float4(0.f, 0.f, 0.f, 0.f); qn[k] = pn[k]; }
      }
#pragma unroll
      for (int k = 0; k < RB; ++k) {
        if (r + k < nrows) {
          const float4 p = pv[k];
          float4 q = qv[k];
          float4 d;
          if (MODE == 0) {
            if (masked) {
              const float4 m = __ldg(reinterpret_cast<const float4*>(tmk + (r + k) * stride));
              q.x *= m.x; q.y *= m.y; q.z *= m.z; q.w *= m.w;
            }
            d = make_float4(gs * (p.x - q.x), gs * (p.y - q.y), gs * (p.z - q.z), gs * (p.w - q.w));
            // torch: dL/dp = (p - t) / max(p (1 - p), 1e-12) times dp/dz = p (1 - p): equal to (p - t) unless p (1 - p) < 1e-12
            // (p rounded to 1, or below 1e-12), where the product collapses towards 0 — reproduced on this rare path
            if (fmaxf(fmaxf(p.x, p.y), fmaxf(p.z, p.w)) == 1.f || fminf(fminf(p.x, p.y), fminf(p.z, p.w)) < 1.0001e-12f) {
              d.x *= fminf(p.x * (1.f - p.x) * 1e12f, 1.f); d.y *= fminf(p.y * (1.f - p.y) * 1e12f, 1.f);
              d.z *= fminf(p.z * (1.f - p.z) * 1e12f, 1.f); d.w *= fminf(p.w * (1.f - p.w) * 1e12f, 1.f);
            }
          } else {
            d = make_float4(q.x * p.x * (1.f - p.x), q.y * p.y * (1.f - p.y), q.z * p.z * (1.f - p.z), q.w * p.w * (1.f - p.w));
          }
          const float4 wy = tp[(r + k) * groups];
          const int is = __float_as_int(wy.w) - i_lo;
          while (cur < is) {                             // the window moves down: retire its first row
            if (cur < nin) *reinterpret_cast<float4*>(mine + cur * W) = a0;
            a0 = a1; a1 = a2; a2 = make_float4(0.f, 0.f, 0.f, 0.f);
            ++cur;
          }
          a0.x = fmaf(wy.x, d.x, a0.x); a0.y = fmaf(wy.x, d.y, a0.y); a0.z = fmaf(wy.x, d.z, a0.z); a0.w = fmaf(wy.x, d.w, a0.w);
          a1.x = fmaf(wy.y, d.x, a1.x); a1.y = fmaf(wy.y, d.y, a1.y); a1.z = fmaf(wy.y, d.z, a1.z); a1.w = fmaf(wy.y, d.w, a1.w);
          a2.x = fmaf(wy.z, d.x, a2.x); a2.y = fmaf(wy.z, d.y, a2.y); a2.z = fmaf(wy.z, d.z, a2.z); a2.w = fmaf(wy.z, d.w, a2.w);
        }
      }
#pragma unroll
      for (int k = 0; k < RB; ++k) { pv[k] = pn[k]; qv[k] = qn[k]; }
    }
    // retire the window and zero-fill the rows this thread never reached
    if (cur < nin) *reinterpret_cast<float4*>(mine + cur * W) = a0;
    if (cur + 1 < nin) *reinterpret_cast<float4*>(mine + (cur + 1) * W) = a1;
    if (cur + 2 < nin) *reinterpret_cast<float4*>(mine + (cur + 2) * W) = a2;
    for (int r = cur + 3; r < nin; ++r) *reinterpret_cast<float4*>(mine + r * W) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  // combine the row groups (fixed order) into group 0
  for (int t = threadIdx.x; t < nin * W; t += blockDim.x) {
    float v = tmp[t];
    for (int gq = 1; gq < groups; ++gq) v += tmp[(long long)gq * nin_max * W + t];
    tmp[t] = v;
  }
  __syncthreads();
  // horizontal fold: out[ii][j] = sum_q wT[j][q] * tmp[ii][16 j - 32 + q]; two lanes per output (even / odd q)
  float* out = a.partials + ((((long long)b * a.heads + h) * bands + band) * nin_max) * g;
  const int n_out = nin * g;
  for (int t0 = 0; t0 < 2 * n_out; t0 += blockDim.x) {
    const int t = t0 + threadIdx.x;
    const bool valid = t < 2 * n_out;
    const int o = valid ? t >> 1 : 0, half = t & 1;
    const int ii = o / g, j = o % g;
    const int qlo = max(0, 32 - 16 * j), qhi = min(HB_SPAN, W + 32 - 16 * j);
    const float* trow = tmp + ii * W + 16 * j - 32;
    const float* wrow = wT + j * HB_SPAN;
    float s = 0.f;
    for (int q = qlo + half; q < qhi; q += 2) s = fmaf(wrow[q], trow[q], s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (valid && half == 0) out[o] = s;
  }
  if (a.counters) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      last_s = atomicAdd(reinterpret_cast<unsigned*>(a.counters) + 1 + b * a.heads + h, 1u) == (unsigned)(bands - 1);
    }
    __syncthreads();
    if (last_s) {
      __threadfence();
      hb_sum_bands(a, a.partials, nin_max, bands, b, h, tap_s);
      if (threadIdx.x == 0) a.counters[1 + b * a.heads + h] = 0;
    }
  }
}

// single block: loss[0] = sum(partials) / count.  Fixed assignment + fixed tree => deterministic.
__global__ void __launch_bounds__(256) loss_finalize_kernel(const float* __restrict__ partials, int n, float inv_count, float* __restrict__ loss) {
  __shared__ float red[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) s += partials[i];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    loss[0] = t * inv_count;
  }
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
  const long long toks = (long long)a->B * a->Np;
  cudaStream_t s = to_stream(s_);
  static int v2 = -1;
  if (v2 < 0) { const char* e = getenv("BFK_HEAD_PROJECT_V2"); v2 = e ? atoi(e) : 1; }
  if (v2 && a->heads <= 4 && (a->D == 768 || a->D == 1152) && (reinterpret_cast<uintptr_t>(a->X) & 15) == 0 && a->x_batch_stride % 4 == 0) {
    // persistent warps, ~8 tokens each when there are enough tokens
    const int grid = (int)std::max<long long>(1, std::min<long long>((long long)sm_count() * (a->D == 768 ? 3 : 2), (toks + 3) / 4));
    if (a->D == 768) head_project_fwd_v2_kernel<6, true><<<grid, 128, 0, s>>>(*a);
    else head_project_fwd_v2_kernel<9, false><<<grid, 128, 0, s>>>(*a);
  } else {
    size_t smem = (size_t)a->heads * a->D * sizeof(float);
    if (int rc = opt_in_smem(head_project_fwd_kernel, smem, "head_project_fwd")) return rc;
    head_project_fwd_kernel<<<(unsigned)((toks + 8 * HP_TOK_PER_WARP - 1) / (8 * HP_TOK_PER_WARP)), 256, smem, s>>>(*a);
  }
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

static int env_bands(const char* name, int dflt) {
  const char* e = getenv(name);
  const int v = e ? atoi(e) : dflt;
  return (v == 1 || v == 2 || v == 4 || v == 8 || v == 16) ? v : dflt;
}
static int hm_bands() { static int v = env_bands("BFK_HM_BANDS", HM_BANDS_DEFAULT); return v; }
static int hb_bands() { static int v = env_bands("BFK_HB_BANDS", HB_BANDS_DEFAULT); return v; }
extern "C" int32_t bfk_heatmap_bands(int32_t g) { (void)g; return hm_bands(); }

// largest number of grid rows one band of 4 g output rows touches (from the host taps; cached per g)
static int band_nin_max(int g, int bands) {
  static std::mutex mu;
  static std::map<int, int> cache;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find(g * 64 + bands);
  if (it != cache.end()) return it->second;
  std::vector<int32_t> ts(16 * g);
  std::vector<float> tw(48 * g);
  if (bfk_upsample16_taps(g, ts.data(), tw.data()) != BFK_OK) return -1;
  const int R = 16 * g / bands;
  int m = 0;
  for (int band = 0; band < bands; ++band)
    m = std::max(m, std::min(g, ts[band * R + R - 1] + 3) - ts[band * R]);
  cache[g * 64 + bands] = m;
  return m;
}
static int hm_threads(int W) {     // whole rows of float4 columns (+ idle lanes up to a full warp)
  const int W4 = W / 4;
  return ((HM_THREADS / W4) * W4 + 31) / 32 * 32;
}

extern "C" size_t bfk_heatmap_fwd_workspace_floats(int32_t B, int32_t heads, int32_t g) {
  (void)g;
  const size_t n = (size_t)B * heads * hm_bands();
  return ((n + 1) & ~(size_t)1) + 2 * n;      // BCE partial sums, then the per-band argmax keys (u64)
}

template <bool LOSS, bool MASK, bool ARGMAX>
static void launch_heatmap_fwd(const bfk_heatmap_fwd_args& k, int nin_max, dim3 grid, int threads, size_t smem, cudaStream_t s) {
  static int minb = -1;
  if (minb < 0) { const char* e = getenv("BFK_HM_MINB"); minb = e ? atoi(e) : 4; }
  if (minb == 3) heatmap_fwd_kernel<LOSS, MASK, ARGMAX, 3><<<grid, threads, smem, s>>>(k, nin_max);
  else heatmap_fwd_kernel<LOSS, MASK, ARGMAX, 4><<<grid, threads, smem, s>>>(k, nin_max);
}

extern "C" int32_t bfk_heatmap_fwd(const bfk_heatmap_fwd_args* a, bfk_stream_t s_) {
  BFK_CHECK_ARG(a && a->grid && a->tap_start && a->tap_w && a->heat && a->B > 0 && a->heads > 0 && a->g > 0 && a->g <= 32,
                "bfk_heatmap_fwd: bad arguments (g <= 32)");
  BFK_CHECK_ARG(!a->target || (a->loss && a->partials), "bfk_heatmap_fwd: target needs loss and partials");
  cudaStream_t s = to_stream(s_);
  const int W = 16 * a->g;
  const int bands = hm_bands();
  const int nin_max = band_nin_max(a->g, bands);
  BFK_CHECK_ARG(nin_max > 0, "bfk_heatmap_fwd: no 3-tap composite for g = %d", a->g);
  const bool loss = a->target != nullptr, amax = a->argmax_key != nullptr;
  const bool mask = loss && a->target_mask && a->mask_heads;
  if (amax && !a->partials) cudaMemsetAsync(a->argmax_key, 0, sizeof(unsigned long long) * a->heads * a->B, s);
  const size_t smem = ((size_t)4 * W + (size_t)(nin_max + 2) * W + (size_t)a->g * a->g) * sizeof(float);
  const dim3 grid(bands, a->heads, a->B);
  const int threads = hm_threads(W);
  if (loss) {
    if (mask) { if (amax) launch_heatmap_fwd<true, true, true>(*a, nin_max, grid, threads, smem, s); else launch_heatmap_fwd<true, true, false>(*a, nin_max, grid, threads, smem, s); }
    else      { if (amax) launch_heatmap_fwd<true, false, true>(*a, nin_max, grid, threads, smem, s); else launch_heatmap_fwd<true, false, false>(*a, nin_max, grid, threads, smem, s); }
  } else {
    if (amax) launch_heatmap_fwd<false, false, true>(*a, nin_max, grid, threads, smem, s); else launch_heatmap_fwd<false, false, false>(*a, nin_max, grid, threads, smem, s);
  }
  const bool keys = amax && a->partials;
  if (!a->counters && (loss || keys)) heatmap_finalize_kernel<<<1, HM_RED, 0, s>>>(*a, bands, loss, keys);
  BFK_CHECK_LAUNCH("bfk_heatmap_fwd");
  return BFK_OK;
}

extern "C" size_t bfk_heatmap_bwd_workspace_floats(int32_t B, int32_t heads, int32_t g) {
  const int nin_max = std::max(band_nin_max(g, hb_bands()), 1);
  return (size_t)B * heads * hb_bands() * nin_max * g;
}

template <int MODE, bool MASK>
static int launch_heatmap_bwd(const bfk_heatmap_bwd_args& k, int nin_max, dim3 grid, int threads, size_t smem, cudaStream_t s) {
  if (int rc = opt_in_smem(heatmap_bwd_stage1<MODE, MASK>, smem, "heatmap_bwd")) return rc;
  heatmap_bwd_stage1<MODE, MASK><<<grid, threads, smem, s>>>(k, nin_max);
  return BFK_OK;
}

extern "C" int32_t bfk_heatmap_bwd(const bfk_heatmap_bwd_args* a, bfk_stream_t s_) {
  BFK_CHECK_ARG(a && a->heat && a->tap_start && a->tap_w && a->dgrid && a->partials && a->g > 0 && a->g <= 32,
                "bfk_heatmap_bwd: bad arguments (g <= 32, partials required)");
  BFK_CHECK_ARG(a->mode == 0 ? (a->target && a->gscale) : (a->dheat != nullptr), "bfk_heatmap_bwd: mode inputs missing");
  const int W = 16 * a->g;
  const int bands = hb_bands();
  const int nin_max = band_nin_max(a->g, bands);
  BFK_CHECK_ARG(nin_max > 0, "bfk_heatmap_bwd: no 3-tap composite for g = %d", a->g);
  const int threads = hm_threads(W), groups = HM_THREADS / (W / 4);
  const size_t smem = ((size_t)4 * W + (size_t)a->g * HB_SPAN + (size_t)groups * nin_max * W) * sizeof(float);
  cudaStream_t s = to_stream(s_);
  const dim3 grid(bands, a->heads, a->B);
  int rc;
  if (a->mode != 0) rc = launch_heatmap_bwd<1, false>(*a, nin_max, grid, threads, smem, s);
  else if (a->target_mask && a->mask_heads) rc = launch_heatmap_bwd<0, true>(*a, nin_max, grid, threads, smem, s);
  else rc = launch_heatmap_bwd<0, false>(*a, nin_max, grid, threads, smem, s);
  if (rc) return rc;
  if (!a->counters) heatmap_bwd_stage2<<<dim3(a->heads, a->B), 256, (size_t)4 * W * sizeof(float), s>>>(*a, nin_max, bands);
  BFK_CHECK_LAUNCH("bfk_heatmap_bwd");
  return BFK_OK;
}

extern "C" int32_t bfk_bce_prob(const float* p, const float* t, const float* t_mask, int64_t n, float* loss,
                                float* partials, bfk_stream_t s_) {
  BFK_CHECK_ARG(p && t && loss && partials && n > 0, "bfk_bce_prob: bad arguments");
  cudaStream_t s = to_stream(s_);
  int blocks = (int)std::min<long long>(1024, (n + 255) / 256);
  bce_prob_stage1<<<blocks, 256, 0, s>>>(p, t, t_mask, n, partials);
  loss_finalize_kernel<<<1, 256, 0, s>>>(partials, blocks, 1.f / (float)n, loss);
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
