// Cross attention of the alternate fusion module (bifold/models/fusion.py:67-125, torch.nn.MultiheadAttention with the image
// tokens as queries and the text (+ context) tokens as keys / values), on the packed projection [rows, 3*H*hd] = [q | k | v]
// of ALL fusion tokens: queries are rows [q0, q0 + nq) of every sample, keys / values rows [k0, k0 + nk).
// SIMT, fp32 accumulation, any head_dim <= 96, deterministic (no atomics): nk is 65 (single frame) to a few hundred tokens,
// one warp per query row (forward, dQ) or per key row (dK / dV).  No shipped config selects this module; it shares nothing
// with the tcgen05 flash kernels of the default ConcatTransformer path.
// Optional additive key bias: the reference passes MultiheadAttention a FLOAT attn_mask built by
// masks[:, :Nc].unsqueeze(1).repeat(heads, Nq, 1) (fusion.py:107-113) - it is ADDED to the scores and its batch axis is tiled,
// so problem (b, h) reads bias row (b * H + h) % B.
#include "common.cuh"

namespace bfk {

constexpr int XA_WARPS = 4;
constexpr int XA_MAXD = 3;   // head_dim <= 96

template <typename T>
__device__ __forceinline__ float xa_dot(const float* __restrict__ a_s, const T* __restrict__ b, int hd) {
  float s = 0.f;
  for (int d = 0; d < hd; ++d) s = fmaf(a_s[d], to_f32<T>(b[d]), s);
  return s;
}
__device__ __forceinline__ const float* xa_bias_row(const bfk_cross_attention_args& a, int b, int h) {
  return a.key_bias ? a.key_bias + (long long)((b * a.H + h) % a.B) * a.nk : nullptr;
}

// grid (ceil(nq / XA_WARPS), H, B); smem per warp: hd + nk floats
template <typename T>
__global__ void __launch_bounds__(XA_WARPS * 32) xattn_fwd_kernel(bfk_cross_attention_args a) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * XA_WARPS + warp, h = blockIdx.y, b = blockIdx.z;
  if (i >= a.nq) return;
  const int hd = a.hd, HD = a.H * a.hd, nk = a.nk;
  float* q_s = smem + warp * (nk + hd);
  float* p_s = q_s + hd;
  const T* base = reinterpret_cast<const T*>(a.qkv) + (long long)b * a.Nf * a.ld_qkv;
  const T* q = base + (long long)(a.q0 + i) * a.ld_qkv + h * hd;
  for (int d = lane; d < hd; d += 32) q_s[d] = to_f32<T>(q[d]);
  __syncwarp();
  const float* bias = xa_bias_row(a, b, h);
  float mx = -INFINITY;
  for (int j = lane; j < nk; j += 32) {
    float s = a.scale * xa_dot<T>(q_s, base + (long long)(a.k0 + j) * a.ld_qkv + HD + h * hd, hd);
    if (bias) s += bias[j];
    p_s[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < nk; j += 32) {
    const float e = __expf(p_s[j] - mx);
    p_s[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  float acc[XA_MAXD] = {};
  const T* v = base + (long long)a.k0 * a.ld_qkv + 2 * HD + h * hd;
  for (int j = 0; j < nk; ++j) {
    const float p = p_s[j];
#pragma unroll
    for (int t = 0; t < XA_MAXD; ++t) {
      const int d = lane + 32 * t;
      if (d < hd) acc[t] = fmaf(p, to_f32<T>(v[(long long)j * a.ld_qkv + d]), acc[t]);
    }
  }
  T* o = reinterpret_cast<T*>(a.out) + ((long long)b * a.Nf + a.q0 + i) * a.ld_out + h * hd;
#pragma unroll
  for (int t = 0; t < XA_MAXD; ++t) {
    const int d = lane + 32 * t;
    if (d < hd) o[d] = from_f32<T>(acc[t] * inv);
  }
  if (lane == 0 && a.lse) a.lse[((long long)b * a.H + h) * a.nq + i] = mx + __logf(sum);
}

// delta[b,h,i] = sum_d dout[i,d] * out[i,d] over the query rows
template <typename T>
__global__ void __launch_bounds__(XA_WARPS * 32) xattn_delta_kernel(bfk_cross_attention_args a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * XA_WARPS + warp, h = blockIdx.y, b = blockIdx.z;
  if (i >= a.nq) return;
  const long long row = (long long)b * a.Nf + a.q0 + i;
  const T* o = reinterpret_cast<const T*>(a.out) + row * a.ld_out + h * a.hd;
  const T* dO = reinterpret_cast<const T*>(a.dout) + row * a.ld_dout + h * a.hd;
  float s = 0.f;
  for (int d = lane; d < a.hd; d += 32) s = fmaf(to_f32<T>(o[d]), to_f32<T>(dO[d]), s);
  s = warp_sum(s);
  if (lane == 0) a.delta[((long long)b * a.H + h) * a.nq + i] = s;
}

// dQ: warp per query row; smem per warp: 2 hd + nk floats
template <typename T>
__global__ void __launch_bounds__(XA_WARPS * 32) xattn_bwd_dq_kernel(bfk_cross_attention_args a) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * XA_WARPS + warp, h = blockIdx.y, b = blockIdx.z;
  if (i >= a.nq) return;
  const int hd = a.hd, HD = a.H * a.hd, nk = a.nk;
  float* q_s = smem + warp * (nk + 2 * hd);
  float* do_s = q_s + hd;
  float* ds_s = do_s + hd;
  const T* base = reinterpret_cast<const T*>(a.qkv) + (long long)b * a.Nf * a.ld_qkv;
  const long long row = (long long)b * a.Nf + a.q0 + i;
  const T* q = base + (long long)(a.q0 + i) * a.ld_qkv + h * hd;
  const T* dO = reinterpret_cast<const T*>(a.dout) + row * a.ld_dout + h * hd;
  for (int d = lane; d < hd; d += 32) { q_s[d] = to_f32<T>(q[d]); do_s[d] = to_f32<T>(dO[d]); }
  __syncwarp();
  const float* bias = xa_bias_row(a, b, h);
  const float lse = a.lse[((long long)b * a.H + h) * a.nq + i], delta = a.delta[((long long)b * a.H + h) * a.nq + i];
  for (int j = lane; j < nk; j += 32) {
    const T* kj = base + (long long)(a.k0 + j) * a.ld_qkv + HD + h * hd;
    const T* vj = kj + HD;
    const float p = __expf(a.scale * xa_dot<T>(q_s, kj, hd) + (bias ? bias[j] : 0.f) - lse);
    ds_s[j] = p * (xa_dot<T>(do_s, vj, hd) - delta);
  }
  __syncwarp();
  float acc[XA_MAXD] = {};
  const T* k = base + (long long)a.k0 * a.ld_qkv + HD + h * hd;
  for (int j = 0; j < nk; ++j) {
    const float ds = ds_s[j];
#pragma unroll
    for (int t = 0; t < XA_MAXD; ++t) {
      const int d = lane + 32 * t;
      if (d < hd) acc[t] = fmaf(ds, to_f32<T>(k[(long long)j * a.ld_qkv + d]), acc[t]);
    }
  }
  T* dq = reinterpret_cast<T*>(a.dqkv) + row * a.ld_dqkv + h * hd;
#pragma unroll
  for (int t = 0; t < XA_MAXD; ++t) {
    const int d = lane + 32 * t;
    if (d < hd) dq[d] = from_f32<T>(acc[t] * a.scale);
  }
}

// dK, dV: warp per key row; smem per warp: 2 hd + 2 nq floats
template <typename T>
__global__ void __launch_bounds__(XA_WARPS * 32) xattn_bwd_dkv_kernel(bfk_cross_attention_args a) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * XA_WARPS + warp, h = blockIdx.y, b = blockIdx.z;
  if (j >= a.nk) return;
  const int hd = a.hd, HD = a.H * a.hd, nq = a.nq;
  float* k_s = smem + warp * (2 * nq + 2 * hd);
  float* v_s = k_s + hd;
  float* p_s = v_s + hd;
  float* ds_s = p_s + nq;
  const T* base = reinterpret_cast<const T*>(a.qkv) + (long long)b * a.Nf * a.ld_qkv;
  const T* kj = base + (long long)(a.k0 + j) * a.ld_qkv + HD + h * hd;
  const T* vj = kj + HD;
  for (int d = lane; d < hd; d += 32) { k_s[d] = to_f32<T>(kj[d]); v_s[d] = to_f32<T>(vj[d]); }
  __syncwarp();
  const float* bias = xa_bias_row(a, b, h);
  const float bj = bias ? bias[j] : 0.f;
  const float* lse = a.lse + ((long long)b * a.H + h) * nq;
  const float* delta = a.delta + ((long long)b * a.H + h) * nq;
  const T* qbase = base + (long long)a.q0 * a.ld_qkv + h * hd;
  const T* dObase = reinterpret_cast<const T*>(a.dout) + ((long long)b * a.Nf + a.q0) * a.ld_dout + h * hd;
  for (int i = lane; i < nq; i += 32) {
    const float p = __expf(a.scale * xa_dot<T>(k_s, qbase + (long long)i * a.ld_qkv, hd) + bj - lse[i]);
    p_s[i] = p;
    ds_s[i] = p * (xa_dot<T>(v_s, dObase + (long long)i * a.ld_dout, hd) - delta[i]);
  }
  __syncwarp();
  float ak[XA_MAXD] = {}, av[XA_MAXD] = {};
  for (int i = 0; i < nq; ++i) {
    const float p = p_s[i], ds = ds_s[i];
#pragma unroll
    for (int t = 0; t < XA_MAXD; ++t) {
      const int d = lane + 32 * t;
      if (d < hd) {
        av[t] = fmaf(p, to_f32<T>(dObase[(long long)i * a.ld_dout + d]), av[t]);
        ak[t] = fmaf(ds, to_f32<T>(qbase[(long long)i * a.ld_qkv + d]), ak[t]);
      }
    }
  }
  T* dk = reinterpret_cast<T*>(a.dqkv) + ((long long)b * a.Nf + a.k0 + j) * a.ld_dqkv + HD + h * hd;
  T* dv = dk + HD;
#pragma unroll
  for (int t = 0; t < XA_MAXD; ++t) {
    const int d = lane + 32 * t;
    if (d < hd) { dk[d] = from_f32<T>(ak[t] * a.scale); dv[d] = from_f32<T>(av[t]); }
  }
}

// attention weights averaged over the heads (MultiheadAttention(need_weights=True, average_attn_weights=True), fusion.py:117-124):
// block per (query row, sample); warp w recomputes the rows of heads w, w + 4, ... from the stored log-sum-exp and adds them into
// its own shared-memory row; the four rows are combined in a fixed order.  smem: 4 * (hd + nk) floats.
template <typename T>
__global__ void __launch_bounds__(XA_WARPS * 32) xattn_weights_kernel(bfk_cross_attention_args a, float* __restrict__ weights) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x, b = blockIdx.y;
  const int hd = a.hd, HD = a.H * a.hd, nk = a.nk;
  float* q_s = smem + warp * (nk + hd);
  float* w_s = q_s + hd;
  for (int j = lane; j < nk; j += 32) w_s[j] = 0.f;
  const T* base = reinterpret_cast<const T*>(a.qkv) + (long long)b * a.Nf * a.ld_qkv;
  for (int h = warp; h < a.H; h += XA_WARPS) {
    __syncwarp();
    const T* q = base + (long long)(a.q0 + i) * a.ld_qkv + h * hd;
    for (int d = lane; d < hd; d += 32) q_s[d] = to_f32<T>(q[d]);
    __syncwarp();
    const float* bias = xa_bias_row(a, b, h);
    const float lse = a.lse[((long long)b * a.H + h) * a.nq + i];
    for (int j = lane; j < nk; j += 32) {
      const float s = a.scale * xa_dot<T>(q_s, base + (long long)(a.k0 + j) * a.ld_qkv + HD + h * hd, hd) + (bias ? bias[j] : 0.f);
      w_s[j] += __expf(s - lse);
    }
  }
  __syncthreads();
  const float inv = 1.f / (float)a.H;
  for (int j = threadIdx.x; j < nk; j += XA_WARPS * 32) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < XA_WARPS; ++w) t += smem[w * (nk + hd) + hd + j];
    weights[((long long)b * a.nq + i) * nk + j] = t * inv;
  }
}

template <typename K>
static int xa_smem(K kernel, size_t bytes, const char* name) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("%s: smem %zu: %s", name, bytes, cudaGetErrorString(e)); return BFK_E_CUDA; }
  }
  return BFK_OK;
}
static bool xa_args_ok(const bfk_cross_attention_args* a, bool bwd) {
  if (!a || !a->qkv || !a->out || !a->lse) return false;
  if (a->B <= 0 || a->H <= 0 || a->hd <= 0 || a->hd > 32 * XA_MAXD || a->nq <= 0 || a->nk <= 0) return false;
  if (a->q0 < 0 || a->k0 < 0 || a->q0 + a->nq > a->Nf || a->k0 + a->nk > a->Nf) return false;
  if (a->dtype != BFK_BF16 && a->dtype != BFK_F32) return false;
  if (bwd && (!a->dout || !a->dqkv || !a->delta)) return false;
  return true;
}

template <typename T>
static int xattn_fwd_launch(const bfk_cross_attention_args* a, cudaStream_t s) {
  const size_t smem = (size_t)XA_WARPS * (a->nk + a->hd) * sizeof(float);
  if (int rc = xa_smem(xattn_fwd_kernel<T>, smem, "bfk_cross_attention_fwd")) return rc;
  dim3 grid((a->nq + XA_WARPS - 1) / XA_WARPS, a->H, a->B);
  xattn_fwd_kernel<T><<<grid, XA_WARPS * 32, smem, s>>>(*a);
  BFK_CHECK_LAUNCH("bfk_cross_attention_fwd");
  return BFK_OK;
}
template <typename T>
static int xattn_bwd_launch(const bfk_cross_attention_args* a, cudaStream_t s) {
  dim3 gq((a->nq + XA_WARPS - 1) / XA_WARPS, a->H, a->B), gk((a->nk + XA_WARPS - 1) / XA_WARPS, a->H, a->B);
  const size_t s1 = (size_t)XA_WARPS * (a->nk + 2 * a->hd) * sizeof(float), s2 = (size_t)XA_WARPS * (2 * a->nq + 2 * a->hd) * sizeof(float);
  if (int rc = xa_smem(xattn_bwd_dq_kernel<T>, s1, "bfk_cross_attention_bwd")) return rc;
  if (int rc = xa_smem(xattn_bwd_dkv_kernel<T>, s2, "bfk_cross_attention_bwd")) return rc;
  xattn_delta_kernel<T><<<gq, XA_WARPS * 32, 0, s>>>(*a);
  xattn_bwd_dq_kernel<T><<<gq, XA_WARPS * 32, s1, s>>>(*a);
  xattn_bwd_dkv_kernel<T><<<gk, XA_WARPS * 32, s2, s>>>(*a);
  BFK_CHECK_LAUNCH("bfk_cross_attention_bwd");
  return BFK_OK;
}

template <typename T>
static int xattn_weights_launch(const bfk_cross_attention_args* a, float* weights, cudaStream_t s) {
  const size_t smem = (size_t)XA_WARPS * (a->nk + a->hd) * sizeof(float);
  if (int rc = xa_smem(xattn_weights_kernel<T>, smem, "bfk_cross_attention_weights")) return rc;
  xattn_weights_kernel<T><<<dim3(a->nq, a->B), XA_WARPS * 32, smem, s>>>(*a, weights);
  BFK_CHECK_LAUNCH("bfk_cross_attention_weights");
  return BFK_OK;
}

}  // namespace bfk

extern "C" int32_t bfk_cross_attention_weights(const bfk_cross_attention_args* a, float* weights, bfk_stream_t s_) {
  using namespace bfk;
  BFK_CHECK_ARG(xa_args_ok(a, false) && weights, "bfk_cross_attention_weights: bad arguments");
  return a->dtype == BFK_BF16 ? xattn_weights_launch<__nv_bfloat16>(a, weights, to_stream(s_)) : xattn_weights_launch<float>(a, weights, to_stream(s_));
}

extern "C" int32_t bfk_cross_attention_fwd(const bfk_cross_attention_args* a, bfk_stream_t s_) {
  using namespace bfk;
  BFK_CHECK_ARG(xa_args_ok(a, false), "bfk_cross_attention_fwd: bad arguments (null pointer, head_dim > 96, or row ranges outside the sample)");
  return a->dtype == BFK_BF16 ? xattn_fwd_launch<__nv_bfloat16>(a, to_stream(s_)) : xattn_fwd_launch<float>(a, to_stream(s_));
}
extern "C" int32_t bfk_cross_attention_bwd(const bfk_cross_attention_args* a, bfk_stream_t s_) {
  using namespace bfk;
  BFK_CHECK_ARG(xa_args_ok(a, true), "bfk_cross_attention_bwd: bad arguments (null pointer, head_dim > 96, or row ranges outside the sample)");
  return a->dtype == BFK_BF16 ? xattn_bwd_launch<__nv_bfloat16>(a, to_stream(s_)) : xattn_bwd_launch<float>(a, to_stream(s_));
}
