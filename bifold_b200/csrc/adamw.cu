// Fused Adam/AdamW over a flat fp32 buffer + bf16 shadow weights; fp32 -> bf16/fp32 strided casts.
// HBM-bound: 28 B per parameter (read p,g,m,v; write p,m,v) + 2 B for the shadow.
#include "common.cuh"

namespace bfk {

__global__ void adamw_advance_kernel(int32_t* step_dev, const float* lr_dev, float beta1, float beta2, float* dyn) {
  const int step = ++(*step_dev);
  dyn[0] = *lr_dev;
  dyn[1] = (float)(1.0 - pow((double)beta1, (double)step));
  dyn[2] = (float)sqrt(1.0 - pow((double)beta2, (double)step));
}

__global__ void __launch_bounds__(256) adamw_kernel(bfk_adamw_args a, float bc1, float bc2_sqrt) {
  const long long n4 = a.n >> 2;
  if (a.dyn) { a.lr = a.dyn[0]; bc1 = a.dyn[1]; bc2_sqrt = a.dyn[2]; }   // device-resident schedule (CUDA-graph replay)
  const float step_size = a.lr / bc1;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n4; i += (long long)gridDim.x * 256) {
    float4 p = reinterpret_cast<float4*>(a.p)[i];
    float4 g = reinterpret_cast<const float4*>(a.g)[i];
    float4 m = reinterpret_cast<float4*>(a.m)[i];
    float4 v = reinterpret_cast<float4*>(a.v)[i];
    float* pp = &p.x; float* gg = &g.x; float* mm = &m.x; float* vv = &v.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gk = gg[k] * a.grad_scale;
      if (a.decoupled) pp[k] *= 1.f - a.lr * a.weight_decay;
      else if (a.weight_decay != 0.f) gk = fmaf(a.weight_decay, pp[k], gk);
      mm[k] = mm[k] + (gk - mm[k]) * (1.f - a.beta1);           // torch: exp_avg.lerp_(grad, 1-beta1)
      vv[k] = vv[k] * a.beta2 + (1.f - a.beta2) * gk * gk;
      float denom = sqrtf(vv[k]) / bc2_sqrt + a.eps;
      pp[k] -= step_size * (mm[k] / denom);
    }
    reinterpret_cast<float4*>(a.p)[i] = p;
    reinterpret_cast<float4*>(a.m)[i] = m;
    reinterpret_cast<float4*>(a.v)[i] = v;
    if (a.shadow_bf16) {
      __nv_bfloat162 s0 = __floats2bfloat162_rn(p.x, p.y), s1 = __floats2bfloat162_rn(p.z, p.w);
      uint2 u; u.x = *reinterpret_cast<uint32_t*>(&s0); u.y = *reinterpret_cast<uint32_t*>(&s1);
      reinterpret_cast<uint2*>(a.shadow_bf16)[i] = u;
    }
  }
  // tail (n % 4)
  if (blockIdx.x == 0 && threadIdx.x < (a.n & 3)) {
    long long i = (n4 << 2) + threadIdx.x;
    float gk = a.g[i] * a.grad_scale, pk = a.p[i];
    if (a.decoupled) pk *= 1.f - a.lr * a.weight_decay;
    else if (a.weight_decay != 0.f) gk = fmaf(a.weight_decay, pk, gk);
    float mk = a.m[i] + (gk - a.m[i]) * (1.f - a.beta1);
    float vk = a.v[i] * a.beta2 + (1.f - a.beta2) * gk * gk;
    pk -= step_size * (mk / (sqrtf(vk) / bc2_sqrt + a.eps));
    a.p[i] = pk; a.m[i] = mk; a.v[i] = vk;
    if (a.shadow_bf16) reinterpret_cast<__nv_bfloat16*>(a.shadow_bf16)[i] = __float2bfloat16_rn(pk);
  }
}

template <typename T>
__global__ void cast_kernel(const float* __restrict__ src, T* __restrict__ dst, long long rows, long long cols,
                            long long ld_src, long long ld_dst) {
  const long long total = rows * cols;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    long long r = i / cols, c = i % cols;
    dst[r * ld_dst + c] = from_f32<T>(src[r * ld_src + c]);
  }
}

__global__ void __launch_bounds__(256) multi_cast_kernel(const bfk_cast_segment* __restrict__ segs, int nseg, long long total_chunks) {
  for (long long chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
    int lo = 0, hi = nseg - 1;                    // last segment whose first_chunk <= chunk (block-uniform search)
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (segs[mid].first_chunk <= chunk) lo = mid; else hi = mid - 1;
    }
    const bfk_cast_segment sg = segs[lo];
    const long long n = (long long)sg.rows * sg.cols;
    const long long e0 = (chunk - sg.first_chunk) * 1024 + threadIdx.x * 4;
    if (e0 >= n) continue;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (e0 + 4 <= n && (reinterpret_cast<uintptr_t>(sg.src + e0) & 15u) == 0) {
      const float4 f = *reinterpret_cast<const float4*>(sg.src + e0);
      v[0] = f.x; v[1] = f.y; v[2] = f.z; v[3] = f.w;
    } else {
      for (int j = 0; j < 4; ++j) if (e0 + j < n) v[j] = sg.src[e0 + j];
    }
    if (e0 + 4 <= n && (sg.cols & 3) == 0 && (sg.ld_dst & 3) == 0) {   // the 4 elements share a row: one packed store
      const long long r = e0 / sg.cols, c = e0 % sg.cols;
      if (sg.dst_dtype == BFK_BF16) {
        __nv_bfloat16* d = reinterpret_cast<__nv_bfloat16*>(sg.dst) + r * sg.ld_dst + c;
        if ((reinterpret_cast<uintptr_t>(d) & 7u) == 0) {
          const __nv_bfloat162 p0 = __floats2bfloat162_rn(v[0] * sg.scale, v[1] * sg.scale), p1 = __floats2bfloat162_rn(v[2] * sg.scale, v[3] * sg.scale);
          uint2 u;
          u.x = *reinterpret_cast<const uint32_t*>(&p0);
          u.y = *reinterpret_cast<const uint32_t*>(&p1);
          *reinterpret_cast<uint2*>(d) = u;
          continue;
        }
      } else {
        float* d = reinterpret_cast<float*>(sg.dst) + r * sg.ld_dst + c;
        if ((reinterpret_cast<uintptr_t>(d) & 15u) == 0) {
          *reinterpret_cast<float4*>(d) = make_float4(v[0] * sg.scale, v[1] * sg.scale, v[2] * sg.scale, v[3] * sg.scale);
          continue;
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long e = e0 + j;
      if (e < n) {
        const long long r = e / sg.cols, c = e % sg.cols;
        const float x = v[j] * sg.scale;
        if (sg.dst_dtype == BFK_BF16) reinterpret_cast<__nv_bfloat16*>(sg.dst)[r * sg.ld_dst + c] = __float2bfloat16_rn(x);
        else reinterpret_cast<float*>(sg.dst)[r * sg.ld_dst + c] = x;
      }
    }
  }
}

}  // namespace bfk

using namespace bfk;

extern "C" int32_t bfk_multi_cast(const bfk_cast_segment* segs, int32_t nseg, int64_t total_chunks, bfk_stream_t s_) {
  BFK_CHECK_ARG(segs && nseg > 0 && total_chunks > 0, "bfk_multi_cast: bad arguments");
  const int blocks = (int)std::min<long long>((long long)sm_count() * 16, total_chunks);
  multi_cast_kernel<<<blocks, 256, 0, to_stream(s_)>>>(segs, nseg, total_chunks);
  BFK_CHECK_LAUNCH("bfk_multi_cast");
  return BFK_OK;
}

extern "C" int32_t bfk_adamw(const bfk_adamw_args* a, bfk_stream_t s_) {
  BFK_CHECK_ARG(a && a->p && a->g && a->m && a->v && a->n > 0 && (a->step >= 1 || a->dyn), "bfk_adamw: bad arguments");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15u) == 0; };
  BFK_CHECK_ARG(al16(a->p) && al16(a->g) && al16(a->m) && al16(a->v) && (!a->shadow_bf16 || (reinterpret_cast<uintptr_t>(a->shadow_bf16) & 7u) == 0),
                "bfk_adamw: buffers must be 16-byte aligned");
  const int hstep = a->step >= 1 ? a->step : 1;
  double bc1 = 1.0 - pow((double)a->beta1, (double)hstep);
  double bc2 = 1.0 - pow((double)a->beta2, (double)hstep);
  int blocks = (int)std::min<long long>((long long)sm_count() * 8, ((a->n >> 2) + 255) / 256 + 1);
  adamw_kernel<<<blocks, 256, 0, to_stream(s_)>>>(*a, (float)bc1, (float)sqrt(bc2));
  BFK_CHECK_LAUNCH("bfk_adamw");
  return BFK_OK;
}

extern "C" int32_t bfk_adamw_advance(int32_t* step_dev, const float* lr_dev, float beta1, float beta2, float* dyn, bfk_stream_t s_) {
  BFK_CHECK_ARG(step_dev && lr_dev && dyn, "bfk_adamw_advance: null pointer");
  adamw_advance_kernel<<<1, 1, 0, to_stream(s_)>>>(step_dev, lr_dev, beta1, beta2, dyn);
  BFK_CHECK_LAUNCH("bfk_adamw_advance");
  return BFK_OK;
}

extern "C" int32_t bfk_cast_f32_to(const float* src, void* dst, int32_t dst_dtype, int64_t rows, int64_t cols,
                                   int64_t ld_src, int64_t ld_dst, bfk_stream_t s_) {
  BFK_CHECK_ARG(src && dst && rows > 0 && cols > 0, "bfk_cast_f32_to: bad arguments");
  int blocks = (int)std::min<long long>((long long)sm_count() * 8, (rows * cols + 255) / 256);
  if (dst_dtype == BFK_BF16) cast_kernel<__nv_bfloat16><<<blocks, 256, 0, to_stream(s_)>>>(src, (__nv_bfloat16*)dst, rows, cols, ld_src, ld_dst);
  else cast_kernel<float><<<blocks, 256, 0, to_stream(s_)>>>(src, (float*)dst, rows, cols, ld_src, ld_dst);
  BFK_CHECK_LAUNCH("bfk_cast_f32_to");
  return BFK_OK;
}

extern "C" int32_t bfk_cast_bf16(const float* src, void* dst, int64_t n, bfk_stream_t s_) {
  return bfk_cast_f32_to(src, dst, BFK_BF16, 1, n, n, n, s_);
}
