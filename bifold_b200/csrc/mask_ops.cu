// Mask branch of PickPlaceConvDecoder (compute_mask=True, bifold/models/pickplace.py:66-69,78-101) and the mask losses
// (bifold/losses/__init__.py:112-155): elementwise HBM-bound kernels, float4 where the layout allows it, deterministic
// two-stage reductions (fixed assignment of elements to threads, fixed tree).
#include <algorithm>

#include "common.cuh"

namespace bfk {

constexpr int MK_THREADS = 256;

// ---- pick heatmaps gated by the mask heatmap ----------------------------------------------------------------
// heat_out[h] = heat_raw[h] * mask (pick heads) | heat_raw[h] (other heads); n4 = B*H*W/4 float4 per head
__global__ void __launch_bounds__(MK_THREADS) mask_gate_fwd_kernel(bfk_mask_gate_args a, long long n4) {
  const float4* m4 = reinterpret_cast<const float4*>(a.mask_p);
  for (long long i = blockIdx.x * (long long)MK_THREADS + threadIdx.x; i < n4; i += (long long)gridDim.x * MK_THREADS) {
    const float4 m = m4[i];
    for (int h = 0; h < a.heads; ++h) {
      float4 p = reinterpret_cast<const float4*>(a.heat_raw)[(long long)h * n4 + i];
      if ((a.pick_heads >> h) & 1u) { p.x *= m.x; p.y *= m.y; p.z *= m.z; p.w *= m.w; }
      reinterpret_cast<float4*>(a.heat_out)[(long long)h * n4 + i] = p;
    }
  }
}
// dheat_raw[h] = dheat_out[h] * mask (pick) | dheat_out[h];  dmask = dheat_out[heads] (if present) + sum_pick dheat_out[h] * heat_raw[h]
// (the pick products feed the mask only when the mask is not detached: pickplace.py:80-101)
__global__ void __launch_bounds__(MK_THREADS) mask_gate_bwd_kernel(bfk_mask_gate_bwd_args a, long long n4) {
  const float4* m4 = reinterpret_cast<const float4*>(a.mask_p);
  for (long long i = blockIdx.x * (long long)MK_THREADS + threadIdx.x; i < n4; i += (long long)gridDim.x * MK_THREADS) {
    const float4 m = m4[i];
    float4 dm = a.has_dmask ? reinterpret_cast<const float4*>(a.dheat_out)[(long long)a.heads * n4 + i] : make_float4(0.f, 0.f, 0.f, 0.f);
    for (int h = 0; h < a.heads; ++h) {
      float4 d = reinterpret_cast<const float4*>(a.dheat_out)[(long long)h * n4 + i];
      if ((a.pick_heads >> h) & 1u) {
        if (!a.detach_mask) {
          const float4 p = reinterpret_cast<const float4*>(a.heat_raw)[(long long)h * n4 + i];
          dm.x = fmaf(d.x, p.x, dm.x); dm.y = fmaf(d.y, p.y, dm.y); dm.z = fmaf(d.z, p.z, dm.z); dm.w = fmaf(d.w, p.w, dm.w);
        }
        d.x *= m.x; d.y *= m.y; d.z *= m.z; d.w *= m.w;
      }
      reinterpret_cast<float4*>(a.dheat_raw)[(long long)h * n4 + i] = d;
    }
    reinterpret_cast<float4*>(a.dmask)[i] = dm;
  }
}

__global__ void __launch_bounds__(MK_THREADS) add_inplace_kernel(float* __restrict__ dst, const float* __restrict__ src, long long n) {
  for (long long i = blockIdx.x * (long long)MK_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * MK_THREADS) dst[i] += src[i];
}

// block sum of up to three values (fixed tree); result valid in thread 0
__device__ __forceinline__ void block_sum3(float& a, float& b, float& c, float (*red)[3]) {
  a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) { red[warp][0] = a; red[warp][1] = b; red[warp][2] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    a = b = c = 0.f;
    for (int w = 0; w < MK_THREADS / 32; ++w) { a += red[w][0]; b += red[w][1]; c += red[w][2]; }
  }
}

// ---- Dice loss (bifold/losses/__init__.py:126-136) -------------------------------------------------------------
constexpr int DICE_CHUNKS = 16;
// grid (DICE_CHUNKS, B): partials[b][chunk] = {sum p t, sum p, sum t} over the chunk
__global__ void __launch_bounds__(MK_THREADS) dice_stage1(const float* __restrict__ p, const float* __restrict__ t, long long n,
                                                          float* __restrict__ partials) {
  __shared__ float red[MK_THREADS / 32][3];
  const int b = blockIdx.y;
  const long long per = (n + DICE_CHUNKS - 1) / DICE_CHUNKS, i0 = blockIdx.x * per, i1 = min(n, i0 + per);
  float spt = 0.f, sp = 0.f, st = 0.f;
  for (long long i = i0 + threadIdx.x; i < i1; i += MK_THREADS) {
    const float pv = p[b * n + i], tv = t[b * n + i];
    spt = fmaf(pv, tv, spt); sp += pv; st += tv;
  }
  block_sum3(spt, sp, st, red);
  if (threadIdx.x == 0) {
    float* o = partials + ((long long)b * DICE_CHUNKS + blockIdx.x) * 3;
    o[0] = spt; o[1] = sp; o[2] = st;
  }
}
// single block: sums[b] = {sum p t, sum p, sum t};  loss = sum_b 1 - (2 spt + 1) / (sp + st + 1)
__global__ void __launch_bounds__(MK_THREADS) dice_stage2(const float* __restrict__ partials, int B, float* __restrict__ sums,
                                                          float* __restrict__ loss) {
  __shared__ float red[MK_THREADS / 32][3];
  float l = 0.f, z0 = 0.f, z1 = 0.f;
  for (int b = threadIdx.x; b < B; b += MK_THREADS) {
    float spt = 0.f, sp = 0.f, st = 0.f;
    for (int c = 0; c < DICE_CHUNKS; ++c) {
      const float* o = partials + ((long long)b * DICE_CHUNKS + c) * 3;
      spt += o[0]; sp += o[1]; st += o[2];
    }
    sums[b * 3] = spt; sums[b * 3 + 1] = sp; sums[b * 3 + 2] = st;
    l += 1.f - (2.f * spt + 1.f) / (sp + st + 1.f);
  }
  block_sum3(l, z0, z1, red);
  if (threadIdx.x == 0) loss[0] = l;
}
// dp_i = -g * (2 t_i D - N) / D^2,  N = 2 spt + 1,  D = sp + st + 1
__global__ void __launch_bounds__(MK_THREADS) dice_bwd_kernel(const float* __restrict__ t, long long n, const float* __restrict__ sums,
                                                             const float* __restrict__ g, float* __restrict__ dp) {
  const int b = blockIdx.y;
  const float N = 2.f * sums[b * 3] + 1.f, D = sums[b * 3 + 1] + sums[b * 3 + 2] + 1.f;
  const float c0 = -g[0] * 2.f / D, c1 = g[0] * N / (D * D);
  for (long long i = blockIdx.x * (long long)MK_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * MK_THREADS)
    dp[b * n + i] = fmaf(c0, t[b * n + i], c1);
}

// ---- Focal loss (bifold/losses/__init__.py:139-155) ------------------------------------------------------------
__device__ __forceinline__ float focal_term(float p, float t, float alpha, float gamma, float* dterm) {
  const float ce = -(t * fmaxf(logf(p), -100.f) + (1.f - t) * fmaxf(log1pf(-p), -100.f));
  const float m = 1.f - (p * t + (1.f - p) * (1.f - t));                   // 1 - p_t
  const float at = alpha >= 0.f ? alpha * t + (1.f - alpha) * (1.f - t) : 1.f;
  const float mg = gamma == 2.f ? m * m : powf(m, gamma);
  if (dterm) {
    const float dce = (p - t) / fmaxf(p * (1.f - p), 1e-12f);               // torch binary_cross_entropy backward
    const float dmg = gamma == 2.f ? 2.f * m : (m > 0.f ? gamma * powf(m, gamma - 1.f) : 0.f);
    *dterm = at * (dce * mg + ce * dmg * (1.f - 2.f * t));
  }
  return at * ce * mg;
}
__global__ void __launch_bounds__(MK_THREADS) focal_stage1(const float* __restrict__ p, const float* __restrict__ t, long long n,
                                                           float alpha, float gamma, float* __restrict__ partials) {
  __shared__ float red[MK_THREADS / 32][3];
  float s = 0.f, z0 = 0.f, z1 = 0.f;
  for (long long i = blockIdx.x * (long long)MK_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * MK_THREADS)
    s += focal_term(p[i], t[i], alpha, gamma, nullptr);
  block_sum3(s, z0, z1, red);
  if (threadIdx.x == 0) partials[blockIdx.x] = s;
}
__global__ void __launch_bounds__(MK_THREADS) focal_stage2(const float* __restrict__ partials, int n, float scale, float* __restrict__ loss) {
  __shared__ float red[MK_THREADS / 32][3];
  float s = 0.f, z0 = 0.f, z1 = 0.f;
  for (int i = threadIdx.x; i < n; i += MK_THREADS) s += partials[i];
  block_sum3(s, z0, z1, red);
  if (threadIdx.x == 0) loss[0] = s * scale;
}
__global__ void __launch_bounds__(MK_THREADS) focal_bwd_kernel(const float* __restrict__ p, const float* __restrict__ t, long long n,
                                                              float alpha, float gamma, float scale, const float* __restrict__ g,
                                                              float* __restrict__ dp) {
  const float gs = g[0] * scale;
  for (long long i = blockIdx.x * (long long)MK_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * MK_THREADS) {
    float d;
    focal_term(p[i], t[i], alpha, gamma, &d);
    dp[i] = gs * d;
  }
}

}  // namespace bfk

using namespace bfk;

static int ew_blocks(long long n) { return (int)std::max<long long>(1, std::min<long long>((long long)sm_count() * 8, (n + MK_THREADS - 1) / MK_THREADS)); }

extern "C" int32_t bfk_mask_gate_fwd(const bfk_mask_gate_args* a, bfk_stream_t s) {
  BFK_CHECK_ARG(a && a->heat_raw && a->mask_p && a->heat_out && a->B > 0 && a->heads > 0 && a->heads <= 8, "bfk_mask_gate_fwd: bad arguments");
  const long long n = (long long)a->B * a->H * a->W;
  BFK_CHECK_ARG(n % 4 == 0, "bfk_mask_gate_fwd: B*H*W must be a multiple of 4");
  mask_gate_fwd_kernel<<<ew_blocks(n / 4), MK_THREADS, 0, to_stream(s)>>>(*a, n / 4);
  BFK_CHECK_LAUNCH("bfk_mask_gate_fwd");
  return BFK_OK;
}
extern "C" int32_t bfk_mask_gate_bwd(const bfk_mask_gate_bwd_args* a, bfk_stream_t s) {
  BFK_CHECK_ARG(a && a->dheat_out && a->heat_raw && a->mask_p && a->dheat_raw && a->dmask && a->B > 0 && a->heads > 0 && a->heads <= 8,
                "bfk_mask_gate_bwd: bad arguments");
  const long long n = (long long)a->B * a->H * a->W;
  BFK_CHECK_ARG(n % 4 == 0, "bfk_mask_gate_bwd: B*H*W must be a multiple of 4");
  mask_gate_bwd_kernel<<<ew_blocks(n / 4), MK_THREADS, 0, to_stream(s)>>>(*a, n / 4);
  BFK_CHECK_LAUNCH("bfk_mask_gate_bwd");
  return BFK_OK;
}
extern "C" int32_t bfk_add_inplace(float* dst, const float* src, int64_t n, bfk_stream_t s) {
  BFK_CHECK_ARG(dst && src && n > 0, "bfk_add_inplace: bad arguments");
  add_inplace_kernel<<<ew_blocks(n), MK_THREADS, 0, to_stream(s)>>>(dst, src, n);
  BFK_CHECK_LAUNCH("bfk_add_inplace");
  return BFK_OK;
}

extern "C" size_t bfk_dice_loss_workspace_floats(int32_t B) { return (size_t)B * DICE_CHUNKS * 3; }
extern "C" int32_t bfk_dice_loss_fwd(const float* p, const float* t, int32_t B, int64_t n, float* sums, float* loss, float* partials,
                                     bfk_stream_t s) {
  BFK_CHECK_ARG(p && t && sums && loss && partials && B > 0 && n > 0, "bfk_dice_loss_fwd: bad arguments");
  dice_stage1<<<dim3(DICE_CHUNKS, B), MK_THREADS, 0, to_stream(s)>>>(p, t, n, partials);
  dice_stage2<<<1, MK_THREADS, 0, to_stream(s)>>>(partials, B, sums, loss);
  BFK_CHECK_LAUNCH("bfk_dice_loss_fwd");
  return BFK_OK;
}
extern "C" int32_t bfk_dice_loss_bwd(const float* t, int32_t B, int64_t n, const float* sums, const float* g, float* dp, bfk_stream_t s) {
  BFK_CHECK_ARG(t && sums && g && dp && B > 0 && n > 0, "bfk_dice_loss_bwd: bad arguments");
  const int bx = (int)std::max<long long>(1, std::min<long long>(64, (n + MK_THREADS - 1) / MK_THREADS));
  dice_bwd_kernel<<<dim3(bx, B), MK_THREADS, 0, to_stream(s)>>>(t, n, sums, g, dp);
  BFK_CHECK_LAUNCH("bfk_dice_loss_bwd");
  return BFK_OK;
}

extern "C" int32_t bfk_focal_loss_fwd(const float* p, const float* t, int64_t n, float alpha, float gamma, float scale, float* loss,
                                      float* partials, bfk_stream_t s) {
  BFK_CHECK_ARG(p && t && loss && partials && n > 0, "bfk_focal_loss_fwd: bad arguments");
  const int blocks = (int)std::min<long long>(1024, (n + MK_THREADS - 1) / MK_THREADS);
  focal_stage1<<<blocks, MK_THREADS, 0, to_stream(s)>>>(p, t, n, alpha, gamma, partials);
  focal_stage2<<<1, MK_THREADS, 0, to_stream(s)>>>(partials, blocks, scale, loss);
  BFK_CHECK_LAUNCH("bfk_focal_loss_fwd");
  return BFK_OK;
}
extern "C" int32_t bfk_focal_loss_bwd(const float* p, const float* t, int64_t n, float alpha, float gamma, float scale, const float* g,
                                      float* dp, bfk_stream_t s) {
  BFK_CHECK_ARG(p && t && g && dp && n > 0, "bfk_focal_loss_bwd: bad arguments");
  focal_bwd_kernel<<<ew_blocks(n), MK_THREADS, 0, to_stream(s)>>>(p, t, n, alpha, gamma, scale, g, dp);
  BFK_CHECK_LAUNCH("bfk_focal_loss_bwd");
  return BFK_OK;
}
