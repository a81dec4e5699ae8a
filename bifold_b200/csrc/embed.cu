// Bandwidth-bound glue kernels around the towers: im2col for the patch-embedding GEMM, text
// embedding gather, LoRA rank-r side path, fusion-token assembly and its backward.
#include "common.cuh"

namespace bfk {

// ---- im2col (HF:178-179) -------------------------------------------------------------------------
// One thread per 4 consecutive pixels of a patch row (p % 4 == 0).
template <typename T>
__global__ void im2col_kernel(const float* __restrict__ rgb, int B, int C, int H, int W, int p, T* __restrict__ out,
                              long long ld) {
  const int g = W / p, gh = H / p;
  const long long total = (long long)B * C * H * (W / 4);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    int x4 = (int)(t % (W / 4));
    long long r = t / (W / 4);
    int y = (int)(r % H); r /= H;
    int c = (int)(r % C);
    int b = (int)(r / C);
    float4 v = reinterpret_cast<const float4*>(rgb)[t];
    int x = x4 * 4, py = y / p, i = y % p, px = x / p, j = x % p;
    long long row = ((long long)b * gh + py) * g + px;
    T* o = out + row * ld + ((long long)c * p + i) * p + j;
    o[0] = from_f32<T>(v.x); o[1] = from_f32<T>(v.y); o[2] = from_f32<T>(v.z); o[3] = from_f32<T>(v.w);
  }
}

// ---- text embedding (HF:217-224) -------------------------------------------------------------------
__global__ void text_embed_kernel(const long long* __restrict__ ids, int L, int D, int vocab,
                                  const float* __restrict__ tok, const float* __restrict__ pos, float* __restrict__ out) {
  const int row = blockIdx.x;  // b*L + t
  long long id = ids[row];
  id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
  const float4* tr = reinterpret_cast<const float4*>(tok + id * D);
  const float4* pr = reinterpret_cast<const float4*>(pos + (long long)(row % L) * D);
  float4* o = reinterpret_cast<float4*>(out + (long long)row * D);
  for (int c = threadIdx.x; c < D / 4; c += blockDim.x) {
    float4 a = __ldg(tr + c), b = __ldg(pr + c);
    o[c] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
  }
}

// ---- LoRA down projection: h[m, D+j] = sum_k drop(h[m,k]) Acat[j,k], j < 16 ------------------------
// Acat is staged transposed in shared memory as fp32 rows of 20 floats (16 used; the pad makes the
// float4 reads of consecutive k conflict-free).  One warp handles LR_ROWS rows per pass so every
// shared-memory read of A is reused LR_ROWS times.
constexpr int LR_ROWS = 4;
constexpr int LR_PAD = 20;

template <typename T>
__device__ __forceinline__ void stage_acat(float* A_s, const T* __restrict__ Acat, int D) {
  for (int i = threadIdx.x; i < 16 * D; i += blockDim.x) {
    int j = i / D, k = i % D;
    A_s[k * LR_PAD + j] = to_f32<T>(Acat[i]);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) lora_down_kernel(T* __restrict__ h, long long ldh, int M, int D,
                                                        const T* __restrict__ Acat, float p_drop, uint64_t seed) {
  extern __shared__ __align__(16) float A_s[];  // [D][LR_PAD]
  stage_acat<T>(A_s, Acat, D);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  for (int row0 = (blockIdx.x * 8 + warp) * LR_ROWS; row0 < M; row0 += gridDim.x * 8 * LR_ROWS) {
    float acc[LR_ROWS][16];
#pragma unroll
    for (int r = 0; r < LR_ROWS; ++r)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[r][j] = 0.f;
    for (int k = lane; k < D; k += 32) {
      float x[LR_ROWS];
#pragma unroll
      for (int r = 0; r < LR_ROWS; ++r) {
        const int row = row0 + r;
        float v = row < M ? to_f32<T>(h[(long long)row * ldh + k]) : 0.f;
        if (p_drop > 0.f) v = hash_uniform(seed, row, k) < p_drop ? 0.f : v * keep_scale;
        x[r] = v;
      }
      const float4* a4 = reinterpret_cast<const float4*>(A_s + k * LR_PAD);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 a = a4[q];
#pragma unroll
        for (int r = 0; r < LR_ROWS; ++r) {
          acc[r][q * 4 + 0] = fmaf(x[r], a.x, acc[r][q * 4 + 0]);
          acc[r][q * 4 + 1] = fmaf(x[r], a.y, acc[r][q * 4 + 1]);
          acc[r][q * 4 + 2] = fmaf(x[r], a.z, acc[r][q * 4 + 2]);
          acc[r][q * 4 + 3] = fmaf(x[r], a.w, acc[r][q * 4 + 3]);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < LR_ROWS; ++r) {
      float mine = 0.f;
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float t = warp_sum(acc[r][j]);
        if (lane == j) mine = t;
      }
      if (lane < 16 && row0 + r < M) h[(long long)(row0 + r) * ldh + D + lane] = from_f32<T>(mine);
    }
  }
}

// dh[m,k] += mask(m,k) * sum_j dh[m,D+j] Acat[j,k]
template <typename T>
__global__ void __launch_bounds__(256) lora_down_bwd_kernel(T* __restrict__ dh, long long ldh, int M, int D,
                                                            const T* __restrict__ Acat, float p_drop, uint64_t seed) {
  extern __shared__ __align__(16) float A_s[];
  stage_acat<T>(A_s, Acat, D);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  for (int row0 = (blockIdx.x * 8 + warp) * LR_ROWS; row0 < M; row0 += gridDim.x * 8 * LR_ROWS) {
    float du[LR_ROWS][16];
#pragma unroll
    for (int r = 0; r < LR_ROWS; ++r)
#pragma unroll
      for (int j = 0; j < 16; ++j) du[r][j] = row0 + r < M ? to_f32<T>(dh[(long long)(row0 + r) * ldh + D + j]) : 0.f;
    for (int k = lane; k < D; k += 32) {
      const float4* a4 = reinterpret_cast<const float4*>(A_s + k * LR_PAD);
      float s[LR_ROWS];
#pragma unroll
      for (int r = 0; r < LR_ROWS; ++r) s[r] = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float4 a = a4[q];
#pragma unroll
        for (int r = 0; r < LR_ROWS; ++r)
          s[r] += du[r][q * 4] * a.x + du[r][q * 4 + 1] * a.y + du[r][q * 4 + 2] * a.z + du[r][q * 4 + 3] * a.w;
      }
#pragma unroll
      for (int r = 0; r < LR_ROWS; ++r) {
        const int row = row0 + r;
        if (row < M) {
          float v = s[r];
          if (p_drop > 0.f) v = hash_uniform(seed, row, k) < p_drop ? 0.f : v * keep_scale;
          T* e = dh + (long long)row * ldh + k;
          *e = from_f32<T>(to_f32<T>(*e) + v);
        }
      }
    }
  }
}

// ---- inverted-dropout copy (LoRA input dropout): dst = mask(seed,row,col) ? src/(1-p) : 0 -----------------
// 8 elements (16 bytes of bf16) per thread when the layout allows it.
template <typename T>
__global__ void dropout_copy_kernel(const T* __restrict__ src, long long lds, T* __restrict__ dst, long long ldd, int M, int D,
                                    float p_drop, uint64_t seed, int vec) {
  const float keep_scale = 1.f / (1.f - p_drop);
  if (vec) {
    const int dv = D / 8;
    const long long total = (long long)M * dv;
    for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
      const int row = (int)(i / dv), col = (int)(i % dv) * 8;
      const uint4 u = *reinterpret_cast<const uint4*>(src + (long long)row * lds + col);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
      __nv_bfloat162 o[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float2 f = __bfloat1622float2(h[t]);
        f.x = hash_uniform(seed, row, col + 2 * t) < p_drop ? 0.f : f.x * keep_scale;
        f.y = hash_uniform(seed, row, col + 2 * t + 1) < p_drop ? 0.f : f.y * keep_scale;
        o[t] = __floats2bfloat162_rn(f.x, f.y);
      }
      *reinterpret_cast<uint4*>(dst + (long long)row * ldd + col) = *reinterpret_cast<uint4*>(o);
    }
    return;
  }
  const long long total = (long long)M * D;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
    const int row = (int)(i / D), col = (int)(i % D);
    const float v = to_f32<T>(src[(long long)row * lds + col]);
    dst[(long long)row * ldd + col] = from_f32<T>(hash_uniform(seed, row, col) < p_drop ? 0.f : v * keep_scale);
  }
}

// ---- tall-skinny Gram: out[i,j] = alpha * sum_m P[m,i] * drop(Q[m,j]) -------------------------------
// stage 1: block = 8 warps over SG_ROWS rows x (32 lanes * VC columns); every lane streams 16-byte
// pieces of Q rows (8 loads in flight per warp), P rows are broadcast from shared memory, the 8 per-warp
// register tiles are combined by a fixed smem tree; stage 2 sums the row-chunk partials in fixed order.
constexpr int SG_ROWS = 256;
template <typename T> struct SgVec;
template <> struct SgVec<float> { static constexpr int VC = 4; };
template <> struct SgVec<__nv_bfloat16> { static constexpr int VC = 8; };
template <typename T>
__device__ __forceinline__ void sg_load(const T* p, bool vec, int nvalid, float (&q)[SgVec<T>::VC]) {
  constexpr int VC = SgVec<T>::VC;
  if (vec && nvalid >= VC) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    if constexpr (sizeof(T) == 4) {
      q[0] = __uint_as_float(u.x); q[1] = __uint_as_float(u.y); q[2] = __uint_as_float(u.z); q[3] = __uint_as_float(u.w);
    } else {
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) { float2 f = __bfloat1622float2(h[t]); q[2 * t] = f.x; q[2 * t + 1] = f.y; }
    }
  } else {
#pragma unroll
    for (int t = 0; t < VC; ++t) q[t] = t < nvalid ? to_f32<T>(p[t]) : 0.f;
  }
}
template <typename T>
__global__ void __launch_bounds__(256) skinny_gram_stage1(const T* __restrict__ P, long long ldp, int ni,
                                                          const T* __restrict__ Q, long long ldq, int nj, int M,
                                                          float p_drop, uint64_t seed, float* __restrict__ partials, int vec) {
  constexpr int VC = SgVec<T>::VC;
  extern __shared__ __align__(16) float sg_smem[];
  float* P_s = sg_smem;                      // [SG_ROWS][16]
  float* red = sg_smem + SG_ROWS * 16;       // [4 warps][32 lanes][16*VC] tree-reduction scratch
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int j = (blockIdx.x * 32 + lane) * VC;
  const int r0 = blockIdx.y * SG_ROWS, r1 = min(M, r0 + SG_ROWS);
  const float keep_scale = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
  for (int t = threadIdx.x; t < SG_ROWS * 16; t += 256) {
    int rr = t / 16, i = t % 16;
    P_s[t] = (r0 + rr < r1 && i < ni) ? to_f32<T>(P[(long long)(r0 + rr) * ldp + i]) : 0.f;
  }
  __syncthreads();
  float acc[16][VC];
#pragma unroll
  for (int i = 0; i < 16; ++i)
#pragma unroll
    for (int t = 0; t < VC; ++t) acc[i][t] = 0.f;
  const int nvalid = nj - j;   // <= 0: lane has no columns
  if (nvalid > 0) {
#pragma unroll 4
    for (int rr = warp; rr < r1 - r0; rr += 8) {
      float q[VC];
      sg_load<T>(Q + (long long)(r0 + rr) * ldq + j, vec != 0, nvalid, q);
      if (p_drop > 0.f) {
#pragma unroll
        for (int t = 0; t < VC; ++t) q[t] = hash_uniform(seed, r0 + rr, j + t) < p_drop ? 0.f : q[t] * keep_scale;
      }
      const float4* p4 = reinterpret_cast<const float4*>(P_s + rr * 16);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4 pv = p4[g];
#pragma unroll
        for (int t = 0; t < VC; ++t) {
          acc[g * 4 + 0][t] = fmaf(pv.x, q[t], acc[g * 4 + 0][t]);
          acc[g * 4 + 1][t] = fmaf(pv.y, q[t], acc[g * 4 + 1][t]);
          acc[g * 4 + 2][t] = fmaf(pv.z, q[t], acc[g * 4 + 2][t]);
          acc[g * 4 + 3][t] = fmaf(pv.w, q[t], acc[g * 4 + 3][t]);
        }
      }
    }
  }
  // fixed tree: warps 4-7 -> 0-3, warps 2-3 -> 0-1, warp 1 -> 0
#pragma unroll
  for (int stride = 4; stride >= 1; stride >>= 1) {
    __syncthreads();
    if (warp >= stride && warp < 2 * stride) {
      float* dst = red + ((warp - stride) * 32 + lane) * (16 * VC);
#pragma unroll
      for (int i = 0; i < 16; ++i)
#pragma unroll
        for (int t = 0; t < VC; ++t) dst[i * VC + t] = acc[i][t];
    }
    __syncthreads();
    if (warp < stride) {
      const float* src = red + (warp * 32 + lane) * (16 * VC);
#pragma unroll
      for (int i = 0; i < 16; ++i)
#pragma unroll
        for (int t = 0; t < VC; ++t) acc[i][t] += src[i * VC + t];
    }
  }
  if (warp == 0 && nvalid > 0) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float* o = partials + ((long long)blockIdx.y * 16 + i) * nj + j;
#pragma unroll
      for (int t = 0; t < VC; ++t)
        if (t < nvalid) o[t] = acc[i][t];
    }
  }
}
__global__ void skinny_gram_stage2(const float* __restrict__ partials, int nb, int ni, int nj, float alpha,
                                   float* __restrict__ out, long long ldo, int transpose_out) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= ni * nj) return;
  int i = t / nj, j = t % nj;
  float s = 0.f;
  for (int b = 0; b < nb; ++b) s += partials[((long long)b * 16 + i) * nj + j];
  s *= alpha;
  if (transpose_out) out[(long long)j * ldo + i] = s; else out[(long long)i * ldo + j] = s;
}

// ---- fusion-token assembly (bifold/models/siglip.py:62-68,91-130; fusion.py:46-56) ---------------------
// grid: (Nf, B); block: D/4 threads (capped at 256, loops)
__global__ void assemble_fwd_kernel(bfk_assemble_args a) {
  const int r = blockIdx.x, b = blockIdx.y;
  const int L1 = a.L + 1, P1 = a.Np + 1;
  const int Nf = L1 + (a.T + 1) * P1;
  const float* src = nullptr;   // tower feature row (or token)
  const float* pos = nullptr;   // context positional embedding row
  int type;
  float mask = 1.f;
  if (r < L1) {
    type = 0;
    src = r == 0 ? a.text_token : a.txt + ((long long)b * a.L + (r - 1)) * a.D;
  } else {
    type = 1;
    int q = r - L1, blk = q / P1, w = q % P1;
    if (blk < a.T) {
      src = w == 0 ? a.image_token : a.ctx + (((long long)b * a.T + blk) * a.Np + (w - 1)) * a.D;
      pos = a.ctx_pos + (long long)q * a.D;
      mask = (float)a.ctx_mask[(long long)b * a.T + blk];
    } else {
      src = w == 0 ? a.image_token : a.img + ((long long)b * a.Np + (w - 1)) * a.D;
    }
  }
  const float4* s4 = reinterpret_cast<const float4*>(src);
  const float4* p4 = reinterpret_cast<const float4*>(pos);
  const float4* e4 = reinterpret_cast<const float4*>(a.type_emb + (long long)type * a.D);
  float4* o4 = reinterpret_cast<float4*>(a.X + ((long long)b * Nf + r) * a.D);
  for (int c = threadIdx.x; c < a.D / 4; c += blockDim.x) {
    float4 v = s4[c], e = __ldg(e4 + c);
    v.x += e.x; v.y += e.y; v.z += e.z; v.w += e.w;
    if (pos) { float4 q = __ldg(p4 + c); v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w; }
    o4[c] = v;
  }
  if (a.mask_out && threadIdx.x == 0) a.mask_out[(long long)b * Nf + r] = mask;
}

// feature gradients: pure row gather.  grid: (Nf, B)
__global__ void assemble_bwd_gather_kernel(bfk_assemble_bwd_args a) {
  const int r = blockIdx.x, b = blockIdx.y;
  const int L1 = a.L + 1, P1 = a.Np + 1;
  const int Nf = L1 + (a.T + 1) * P1;
  float* dst = nullptr;
  if (r < L1) {
    if (r > 0 && a.dtxt) dst = a.dtxt + ((long long)b * a.L + (r - 1)) * a.D;
  } else {
    int q = r - L1, blk = q / P1, w = q % P1;
    if (w > 0) {
      if (blk < a.T) { if (a.dctx) dst = a.dctx + (((long long)b * a.T + blk) * a.Np + (w - 1)) * a.D; }
      else if (a.dimg) dst = a.dimg + ((long long)b * a.Np + (w - 1)) * a.D;
    }
  }
  if (!dst) return;
  const float4* s4 = reinterpret_cast<const float4*>(a.dX + ((long long)b * Nf + r) * a.D);
  float4* d4 = reinterpret_cast<float4*>(dst);
  for (int c = threadIdx.x; c < a.D / 4; c += blockDim.x) d4[c] = s4[c];
}

// parameter gradients: column-parallel, deterministic sums over (b, r).
// small jobs: grid (ceil(D/128), n_jobs) with jobs 0 text_token, 1 image_token, 2.. ctx_pos rows
__global__ void assemble_bwd_param_kernel(bfk_assemble_bwd_args a) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= a.D) return;
  const int L1 = a.L + 1, P1 = a.Np + 1;
  const int Nf = L1 + (a.T + 1) * P1;
  const int job = blockIdx.y;
  float s = 0.f;
  auto at = [&](int b, int r) { return a.dX[((long long)b * Nf + r) * a.D + d]; };
  if (job == 0) {
    for (int b = 0; b < a.B; ++b) s += at(b, 0);
    a.d_text_token[d] = s;
  } else if (job == 1) {
    for (int b = 0; b < a.B; ++b)
      for (int blk = 0; blk <= a.T; ++blk) s += at(b, L1 + blk * P1);
    a.d_image_token[d] = s;
  } else {
    const int q = job - 2;  // context position row
    for (int b = 0; b < a.B; ++b) s += at(b, L1 + q);
    a.d_ctx_pos[(long long)q * a.D + d] = s;
  }
}

// type-embedding gradient: sum over all rows of a type, split AS_SPLIT ways, then reduced.
constexpr int AS_SPLIT = 64;
// grid (ceil(D/128), 2, AS_SPLIT)
__global__ void assemble_bwd_type_stage1(bfk_assemble_bwd_args a, float* __restrict__ partials) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= a.D) return;
  const int L1 = a.L + 1, P1 = a.Np + 1;
  const int Nf = L1 + (a.T + 1) * P1;
  const int type = blockIdx.y;
  const int rlo = type == 0 ? 0 : L1, rn = type == 0 ? L1 : Nf - L1;
  const long long total = (long long)a.B * rn;
  const long long per = (total + AS_SPLIT - 1) / AS_SPLIT;
  const long long t0 = per * blockIdx.z, t1 = min(total, t0 + per);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;   // independent chains: several loads in flight
  auto at = [&](long long t) { int b = (int)(t / rn), r = rlo + (int)(t % rn); return a.dX[((long long)b * Nf + r) * a.D + d]; };
  long long t = t0;
  for (; t + 4 <= t1; t += 4) { s0 += at(t); s1 += at(t + 1); s2 += at(t + 2); s3 += at(t + 3); }
  for (; t < t1; ++t) s0 += at(t);
  partials[((long long)type * AS_SPLIT + blockIdx.z) * a.D + d] = (s0 + s1) + (s2 + s3);
}
__global__ void assemble_bwd_type_stage2(const float* __restrict__ partials, int D, float* __restrict__ d_type_emb) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= 2 * D) return;
  const int type = t / D, d = t % D;
  float s = 0.f;
  for (int k = 0; k < AS_SPLIT; ++k) s += partials[((long long)type * AS_SPLIT + k) * D + d];
  d_type_emb[t] = s;
}

}  // namespace bfk

using namespace bfk;

extern "C" int32_t bfk_im2col(const float* rgb, int32_t B, int32_t C, int32_t H, int32_t W, int32_t p, void* out,
                              int64_t ld_out, int32_t out_dtype, bfk_stream_t s_) {
  BFK_CHECK_ARG(rgb && out && B > 0 && p > 0 && p % 4 == 0 && H % p == 0 && W % p == 0, "bfk_im2col: bad shape");
  long long total = (long long)B * C * H * (W / 4);
  int grid = (int)std::min<long long>((total + 255) / 256, (long long)sm_count() * 16);
  if (out_dtype == BFK_BF16) im2col_kernel<__nv_bfloat16><<<grid, 256, 0, to_stream(s_)>>>(rgb, B, C, H, W, p, (__nv_bfloat16*)out, ld_out);
  else im2col_kernel<float><<<grid, 256, 0, to_stream(s_)>>>(rgb, B, C, H, W, p, (float*)out, ld_out);
  BFK_CHECK_LAUNCH("bfk_im2col");
  return BFK_OK;
}

extern "C" int32_t bfk_text_embed(const int64_t* ids, int32_t B, int32_t L, int32_t D, int32_t vocab,
                                  const float* tok_emb, const float* pos_emb, float* out, bfk_stream_t s_) {
  BFK_CHECK_ARG(ids && tok_emb && pos_emb && out && B > 0 && L > 0 && D % 4 == 0, "bfk_text_embed: bad arguments");
  text_embed_kernel<<<B * L, std::min(256, std::max(32, D / 4)), 0, to_stream(s_)>>>((const long long*)ids, L, D, vocab, tok_emb, pos_emb, out);
  BFK_CHECK_LAUNCH("bfk_text_embed");
  return BFK_OK;
}

template <typename K>
static int lora_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) { set_error("lora: smem %zu: %s", bytes, cudaGetErrorString(e)); return BFK_E_CUDA; }
  }
  return BFK_OK;
}

extern "C" int32_t bfk_lora_down(void* h, int64_t ldh, int32_t dtype, int32_t M, int32_t D, const void* Acat,
                                 float dropout_p, uint64_t seed, bfk_stream_t s_) {
  BFK_CHECK_ARG(h && Acat && M > 0 && D > 0 && ldh >= D + 16, "bfk_lora_down: bad arguments");
  size_t smem = (size_t)LR_PAD * D * sizeof(float);
  int grid = std::min((M + 8 * LR_ROWS - 1) / (8 * LR_ROWS), sm_count() * 3);
  if (dtype == BFK_BF16) {
    if (int rc = lora_smem(lora_down_kernel<__nv_bfloat16>, smem)) return rc;
    lora_down_kernel<__nv_bfloat16><<<grid, 256, smem, to_stream(s_)>>>((__nv_bfloat16*)h, ldh, M, D, (const __nv_bfloat16*)Acat, dropout_p, seed);
  } else {
    if (int rc = lora_smem(lora_down_kernel<float>, smem)) return rc;
    lora_down_kernel<float><<<grid, 256, smem, to_stream(s_)>>>((float*)h, ldh, M, D, (const float*)Acat, dropout_p, seed);
  }
  BFK_CHECK_LAUNCH("bfk_lora_down");
  return BFK_OK;
}

extern "C" int32_t bfk_lora_down_bwd(void* dh, int64_t ldh, int32_t dtype, int32_t M, int32_t D, const void* Acat,
                                     float dropout_p, uint64_t seed, bfk_stream_t s_) {
  BFK_CHECK_ARG(dh && Acat && M > 0 && D > 0 && ldh >= D + 16, "bfk_lora_down_bwd: bad arguments");
  size_t smem = (size_t)LR_PAD * D * sizeof(float);
  int grid = std::min((M + 8 * LR_ROWS - 1) / (8 * LR_ROWS), sm_count() * 3);
  if (dtype == BFK_BF16) {
    if (int rc = lora_smem(lora_down_bwd_kernel<__nv_bfloat16>, smem)) return rc;
    lora_down_bwd_kernel<__nv_bfloat16><<<grid, 256, smem, to_stream(s_)>>>((__nv_bfloat16*)dh, ldh, M, D, (const __nv_bfloat16*)Acat, dropout_p, seed);
  } else {
    if (int rc = lora_smem(lora_down_bwd_kernel<float>, smem)) return rc;
    lora_down_bwd_kernel<float><<<grid, 256, smem, to_stream(s_)>>>((float*)dh, ldh, M, D, (const float*)Acat, dropout_p, seed);
  }
  BFK_CHECK_LAUNCH("bfk_lora_down_bwd");
  return BFK_OK;
}

extern "C" int32_t bfk_dropout_copy(const void* src, int64_t lds, void* dst, int64_t ldd, int32_t dtype, int32_t M, int32_t D,
                                    float dropout_p, uint64_t seed, bfk_stream_t s_) {
  BFK_CHECK_ARG(src && dst && M > 0 && D > 0 && dropout_p >= 0.f && dropout_p < 1.f, "bfk_dropout_copy: bad arguments");
  const int vec = (dtype == BFK_BF16 && D % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0 &&
                   (reinterpret_cast<uintptr_t>(dst) & 15u) == 0) ? 1 : 0;
  int blocks = (int)std::min<long long>((long long)sm_count() * 16, ((long long)M * D / (vec ? 8 : 1) + 255) / 256);
  if (dtype == BFK_BF16) dropout_copy_kernel<__nv_bfloat16><<<blocks, 256, 0, to_stream(s_)>>>((const __nv_bfloat16*)src, lds, (__nv_bfloat16*)dst, ldd, M, D, dropout_p, seed, vec);
  else dropout_copy_kernel<float><<<blocks, 256, 0, to_stream(s_)>>>((const float*)src, lds, (float*)dst, ldd, M, D, dropout_p, seed, 0);
  BFK_CHECK_LAUNCH("bfk_dropout_copy");
  return BFK_OK;
}

extern "C" size_t bfk_skinny_gram_workspace_floats(int32_t M, int32_t ni, int32_t nj) {
  (void)ni;
  return (size_t)((M + SG_ROWS - 1) / SG_ROWS) * 16 * (size_t)nj;
}

extern "C" int32_t bfk_skinny_gram(const void* P, int64_t ldp, int32_t ni, const void* Q, int64_t ldq, int32_t nj,
                                   int32_t dtype, int32_t M, float alpha, float* out, int64_t ldo,
                                   int32_t transpose_out, float dropout_p, uint64_t seed, float* partials,
                                   bfk_stream_t s_) {
  BFK_CHECK_ARG(P && Q && out && partials && ni > 0 && ni <= 16 && nj > 0 && M > 0, "bfk_skinny_gram: bad arguments");
  int nb = (M + SG_ROWS - 1) / SG_ROWS;
  cudaStream_t s = to_stream(s_);
  const size_t es = dtype == BFK_BF16 ? 2 : 4;
  const int vc = (int)(16 / es);
  const int vec = ((reinterpret_cast<uintptr_t>(Q) & 15u) == 0 && (ldq * es) % 16 == 0) ? 1 : 0;
  dim3 grid((nj + 32 * vc - 1) / (32 * vc), nb);
  const size_t smem = ((size_t)SG_ROWS * 16 + (size_t)4 * 32 * 16 * vc) * sizeof(float);
  if (dtype == BFK_BF16) {
    if (int rc = lora_smem(skinny_gram_stage1<__nv_bfloat16>, smem)) return rc;
    skinny_gram_stage1<__nv_bfloat16><<<grid, 256, smem, s>>>((const __nv_bfloat16*)P, ldp, ni, (const __nv_bfloat16*)Q, ldq, nj, M, dropout_p, seed, partials, vec);
  } else {
    if (int rc = lora_smem(skinny_gram_stage1<float>, smem)) return rc;
    skinny_gram_stage1<float><<<grid, 256, smem, s>>>((const float*)P, ldp, ni, (const float*)Q, ldq, nj, M, dropout_p, seed, partials, vec);
  }
  skinny_gram_stage2<<<(ni * nj + 127) / 128, 128, 0, s>>>(partials, nb, ni, nj, alpha, out, ldo, transpose_out);
  BFK_CHECK_LAUNCH("bfk_skinny_gram");
  return BFK_OK;
}

extern "C" int32_t bfk_assemble_fwd(const bfk_assemble_args* a, bfk_stream_t s_) {
  BFK_CHECK_ARG(a && a->txt && a->img && a->text_token && a->image_token && a->type_emb && a->X, "bfk_assemble_fwd: null pointer");
  BFK_CHECK_ARG(a->D % 4 == 0 && a->B > 0 && (a->T == 0 || (a->ctx && a->ctx_pos && a->ctx_mask)), "bfk_assemble_fwd: bad arguments");
  int Nf = (a->L + 1) + (a->T + 1) * (a->Np + 1);
  assemble_fwd_kernel<<<dim3(Nf, a->B), std::min(256, std::max(32, a->D / 4)), 0, to_stream(s_)>>>(*a);
  BFK_CHECK_LAUNCH("bfk_assemble_fwd");
  return BFK_OK;
}

extern "C" int32_t bfk_assemble_bwd(const bfk_assemble_bwd_args* a, bfk_stream_t s_) {
  BFK_CHECK_ARG(a && a->dX && a->d_text_token && a->d_image_token && a->d_type_emb, "bfk_assemble_bwd: null pointer");
  BFK_CHECK_ARG(a->D % 4 == 0 && a->B > 0 && (a->T == 0 || a->d_ctx_pos), "bfk_assemble_bwd: bad arguments");
  int Nf = (a->L + 1) + (a->T + 1) * (a->Np + 1);
  cudaStream_t s = to_stream(s_);
  if (a->dtxt || a->dimg || a->dctx)
    assemble_bwd_gather_kernel<<<dim3(Nf, a->B), std::min(256, std::max(32, a->D / 4)), 0, s>>>(*a);
  BFK_CHECK_ARG(a->partials, "bfk_assemble_bwd: partials workspace (2*64*D floats) required");
  int jobs = 2 + a->T * (a->Np + 1);
  assemble_bwd_param_kernel<<<dim3((a->D + 127) / 128, jobs), 128, 0, s>>>(*a);
  assemble_bwd_type_stage1<<<dim3((a->D + 127) / 128, 2, AS_SPLIT), 128, 0, s>>>(*a, a->partials);
  assemble_bwd_type_stage2<<<(2 * a->D + 127) / 128, 128, 0, s>>>(a->partials, a->D, a->d_type_emb);
  BFK_CHECK_LAUNCH("bfk_assemble_bwd");
  return BFK_OK;
}
