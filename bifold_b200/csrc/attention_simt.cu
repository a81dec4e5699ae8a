// SIMT attention (fp32 accumulate, any head_dim/sequence length): the fp32 "parity mode" kernel and
// the checker of the tensor-core path.  One warp per query row (forward, dQ) or per key row (dK/dV);
// deterministic, no atomics.  Query-row mask semantics of bifold/models/transformer.py:33-35.
#include "common.cuh"

namespace bfk {

constexpr int AT_WARPS = 4;
constexpr int AT_MAXD = 3;  // head_dim <= 96

template <typename T>
__device__ __forceinline__ float dot_row(const float* __restrict__ a_s, const T* __restrict__ b, int hd) {
  float s = 0.f;
  for (int d = 0; d < hd; ++d) s = fmaf(a_s[d], to_f32<T>(b[d]), s);
  return s;
}

// grid: (ceil(N/AT_WARPS), H, B)
template <typename T>
__global__ void __launch_bounds__(AT_WARPS * 32) attn_fwd_simt(bfk_attention_args a) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * AT_WARPS + warp, h = blockIdx.y, b = blockIdx.z;
  if (i >= a.N) return;
  const int N = a.N, hd = a.hd, HD = a.H * a.hd;
  float* q_s = smem + warp * (N + hd);
  float* p_s = q_s + hd;
  const T* base = reinterpret_cast<const T*>(a.qkv) + (long long)b * N * a.ld_qkv;
  const T* q = base + (long long)i * a.ld_qkv + h * hd;
  for (int d = lane; d < hd; d += 32) q_s[d] = to_f32<T>(q[d]);
  __syncwarp();
  const bool masked = a.row_mask && a.row_mask[(long long)b * N + i] == 0.f;
  float mx = -INFINITY;
  for (int j = lane; j < N; j += 32) {
    float s = masked ? a.mask_fill : a.scale * dot_row<T>(q_s, base + (long long)j * a.ld_qkv + HD + h * hd, hd);
    p_s[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < N; j += 32) {
    float e = __expf(p_s[j] - mx);
    p_s[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  float acc[AT_MAXD] = {};
  const T* v = base + 2 * HD + h * hd;
  for (int j = 0; j < N; ++j) {
    const float p = p_s[j];
#pragma unroll
    for (int t = 0; t < AT_MAXD; ++t) {
      int d = lane + 32 * t;
      if (d < hd) acc[t] = fmaf(p, to_f32<T>(v[(long long)j * a.ld_qkv + d]), acc[t]);
    }
  }
  T* o = reinterpret_cast<T*>(a.out) + ((long long)b * N + i) * a.ld_out + h * hd;
#pragma unroll
  for (int t = 0; t < AT_MAXD; ++t) {
    int d = lane + 32 * t;
    if (d < hd) o[d] = from_f32<T>(acc[t] * inv);
  }
  if (lane == 0 && a.lse) a.lse[((long long)b * a.H + h) * N + i] = mx + __logf(sum);
}

// delta[b,h,i] = sum_d dout[i,d]*out[i,d];  grid: (ceil(N/AT_WARPS), H, B)
template <typename T>
__global__ void __launch_bounds__(AT_WARPS * 32) attn_delta_simt(bfk_attention_args a) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * AT_WARPS + warp, h = blockIdx.y, b = blockIdx.z;
  if (i >= a.N) return;
  const T* o = reinterpret_cast<const T*>(a.out) + ((long long)b * a.N + i) * a.ld_out + h * a.hd;
  const T* dO = reinterpret_cast<const T*>(a.dout) + ((long long)b * a.N + i) * a.ld_dout + h * a.hd;
  float s = 0.f;
  for (int d = lane; d < a.hd; d += 32) s = fmaf(to_f32<T>(o[d]), to_f32<T>(dO[d]), s);
  s = warp_sum(s);
  if (lane == 0) a.delta[((long long)b * a.H + h) * a.N + i] = s;
}

// dQ: warp per query row.
template <typename T>
__global__ void __launch_bounds__(AT_WARPS * 32) attn_bwd_dq_simt(bfk_attention_args a) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * AT_WARPS + warp, h = blockIdx.y, b = blockIdx.z;
  if (i >= a.N) return;
  const int N = a.N, hd = a.hd, HD = a.H * a.hd;
  float* q_s = smem + warp * (N + 2 * hd);
  float* do_s = q_s + hd;
  float* ds_s = do_s + hd;
  const T* base = reinterpret_cast<const T*>(a.qkv) + (long long)b * N * a.ld_qkv;
  T* dq = reinterpret_cast<T*>(a.dqkv) + ((long long)b * N + i) * a.ld_dqkv + h * hd;
  const bool masked = a.row_mask && a.row_mask[(long long)b * N + i] == 0.f;
  if (masked) {  // masked_fill_ cuts the gradient to q (and k) for this row
    for (int d = lane; d < hd; d += 32) dq[d] = from_f32<T>(0.f);
    return;
  }
  const T* q = base + (long long)i * a.ld_qkv + h * hd;
  const T* dO = reinterpret_cast<const T*>(a.dout) + ((long long)b * N + i) * a.ld_dout + h * hd;
  for (int d = lane; d < hd; d += 32) { q_s[d] = to_f32<T>(q[d]); do_s[d] = to_f32<T>(dO[d]); }
  __syncwarp();
  const float lse = a.lse[((long long)b * a.H + h) * N + i], delta = a.delta[((long long)b * a.H + h) * N + i];
  for (int j = lane; j < N; j += 32) {
    const T* kj = base + (long long)j * a.ld_qkv + HD + h * hd;
    const T* vj = base + (long long)j * a.ld_qkv + 2 * HD + h * hd;
    float p = __expf(a.scale * dot_row<T>(q_s, kj, hd) - lse);
    float dp = dot_row<T>(do_s, vj, hd);
    ds_s[j] = p * (dp - delta);
  }
  __syncwarp();
  float acc[AT_MAXD] = {};
  const T* k = base + HD + h * hd;
  for (int j = 0; j < N; ++j) {
    const float ds = ds_s[j];
#pragma unroll
    for (int t = 0; t < AT_MAXD; ++t) {
      int d = lane + 32 * t;
      if (d < hd) acc[t] = fmaf(ds, to_f32<T>(k[(long long)j * a.ld_qkv + d]), acc[t]);
    }
  }
#pragma unroll
  for (int t = 0; t < AT_MAXD; ++t) {
    int d = lane + 32 * t;
    if (d < hd) dq[d] = from_f32<T>(acc[t] * a.scale);
  }
}

// dK, dV: warp per key row.
template <typename T>
__global__ void __launch_bounds__(AT_WARPS * 32) attn_bwd_dkv_simt(bfk_attention_args a) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = blockIdx.x * AT_WARPS + warp, h = blockIdx.y, b = blockIdx.z;
  if (j >= a.N) return;
  const int N = a.N, hd = a.hd, HD = a.H * a.hd;
  float* k_s = smem + warp * (2 * N + 2 * hd);
  float* v_s = k_s + hd;
  float* p_s = v_s + hd;
  float* ds_s = p_s + N;
  const T* base = reinterpret_cast<const T*>(a.qkv) + (long long)b * N * a.ld_qkv;
  const T* kj = base + (long long)j * a.ld_qkv + HD + h * hd;
  const T* vj = base + (long long)j * a.ld_qkv + 2 * HD + h * hd;
  for (int d = lane; d < hd; d += 32) { k_s[d] = to_f32<T>(kj[d]); v_s[d] = to_f32<T>(vj[d]); }
  __syncwarp();
  const float* lse = a.lse + ((long long)b * a.H + h) * N;
  const float* delta = a.delta + ((long long)b * a.H + h) * N;
  const T* dObase = reinterpret_cast<const T*>(a.dout) + (long long)b * N * a.ld_dout + h * hd;
  for (int i = lane; i < N; i += 32) {
    const bool masked = a.row_mask && a.row_mask[(long long)b * N + i] == 0.f;
    const T* qi = base + (long long)i * a.ld_qkv + h * hd;
    // masked rows: every score equals mask_fill => p is exactly 1/N (lse would lose precision at -1e5)
    float p = masked ? 1.f / (float)N : __expf(a.scale * dot_row<T>(k_s, qi, hd) - lse[i]);
    float dp = dot_row<T>(v_s, dObase + (long long)i * a.ld_dout, hd);
    p_s[i] = p;
    ds_s[i] = masked ? 0.f : p * (dp - delta[i]);
  }
  __syncwarp();
  float ak[AT_MAXD] = {}, av[AT_MAXD] = {};
  for (int i = 0; i < N; ++i) {
    const float p = p_s[i], ds = ds_s[i];
#pragma unroll
    for (int t = 0; t < AT_MAXD; ++t) {
      int d = lane + 32 * t;
      if (d < hd) {
        av[t] = fmaf(p, to_f32<T>(dObase[(long long)i * a.ld_dout + d]), av[t]);
        ak[t] = fmaf(ds, to_f32<T>(base[(long long)i * a.ld_qkv + h * hd + d]), ak[t]);
      }
    }
  }
  T* dk = reinterpret_cast<T*>(a.dqkv) + ((long long)b * N + j) * a.ld_dqkv + HD + h * hd;
  T* dv = dk + HD;
#pragma unroll
  for (int t = 0; t < AT_MAXD; ++t) {
    int d = lane + 32 * t;
    if (d < hd) { dk[d] = from_f32<T>(ak[t] * a.scale); dv[d] = from_f32<T>(av[t]); }
  }
}

template <typename K>
static int set_smem(K kernel, size_t bytes, const char* name) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("%s: smem %zu: %s", name, bytes, cudaGetErrorString(e)); return BFK_E_CUDA; }
  }
  return BFK_OK;
}

template <typename T>
static int attn_fwd_simt_launch(const bfk_attention_args* a, cudaStream_t s) {
  size_t smem = (size_t)AT_WARPS * (a->N + a->hd) * sizeof(float);
  int rc = set_smem(attn_fwd_simt<T>, smem, "attention_fwd_simt");
  if (rc) return rc;
  dim3 grid((a->N + AT_WARPS - 1) / AT_WARPS, a->H, a->B);
  attn_fwd_simt<T><<<grid, AT_WARPS * 32, smem, s>>>(*a);
  BFK_CHECK_LAUNCH("attention_fwd_simt");
  return BFK_OK;
}

template <typename T>
static int attn_bwd_simt_launch(const bfk_attention_args* a, cudaStream_t s) {
  dim3 grid((a->N + AT_WARPS - 1) / AT_WARPS, a->H, a->B);
  attn_delta_simt<T><<<grid, AT_WARPS * 32, 0, s>>>(*a);
  size_t smem1 = (size_t)AT_WARPS * (a->N + 2 * a->hd) * sizeof(float);
  size_t smem2 = (size_t)AT_WARPS * (2 * a->N + 2 * a->hd) * sizeof(float);
  int rc = set_smem(attn_bwd_dq_simt<T>, smem1, "attention_bwd_dq_simt");
  if (rc) return rc;
  rc = set_smem(attn_bwd_dkv_simt<T>, smem2, "attention_bwd_dkv_simt");
  if (rc) return rc;
  attn_bwd_dq_simt<T><<<grid, AT_WARPS * 32, smem1, s>>>(*a);
  attn_bwd_dkv_simt<T><<<grid, AT_WARPS * 32, smem2, s>>>(*a);
  BFK_CHECK_LAUNCH("attention_bwd_simt");
  return BFK_OK;
}

int attention_fwd_simt(const bfk_attention_args* a, cudaStream_t s) {
  return a->dtype == BFK_BF16 ? attn_fwd_simt_launch<__nv_bfloat16>(a, s) : attn_fwd_simt_launch<float>(a, s);
}
int attention_bwd_simt(const bfk_attention_args* a, cudaStream_t s) {
  return a->dtype == BFK_BF16 ? attn_bwd_simt_launch<__nv_bfloat16>(a, s) : attn_bwd_simt_launch<float>(a, s);
}

}  // namespace bfk
