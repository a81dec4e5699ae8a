// SIMT GEMM with the fused epilogue: the fp32 "parity mode" GEMM (and the checker of the tcgen05 path).
// D[M,N] = epi(A[M,K] * B[N,K]^T), any combination of operand transposes, fp32 accumulate.
#include "epilogue.cuh"

namespace bfk {

constexpr int SB_M = 64, SB_N = 64, SB_K = 16;

template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const T* __restrict__ A, long long lda, int ta,
                                                        const T* __restrict__ B, long long ldb, int tb, int K,
                                                        EpiParams p) {
  __shared__ float As[SB_K][SB_M + 4];
  __shared__ float Bs[SB_K][SB_N + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int m0 = blockIdx.y * SB_M, n0 = blockIdx.x * SB_N;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += SB_K) {
#pragma unroll
    for (int i = tid; i < SB_M * SB_K; i += 256) {
      int mm, kk;
      if (ta) { kk = i / SB_M; mm = i % SB_M; } else { mm = i / SB_K; kk = i % SB_K; }
      int m = m0 + mm, k = k0 + kk;
      float v = 0.f;
      if (m < p.M && k < K) v = to_f32<T>(ta ? A[(long long)k * lda + m] : A[(long long)m * lda + k]);
      As[kk][mm] = v;
    }
#pragma unroll
    for (int i = tid; i < SB_N * SB_K; i += 256) {
      int nn, kk;
      if (tb) { kk = i / SB_N; nn = i % SB_N; } else { nn = i / SB_K; kk = i % SB_K; }
      int n = n0 + nn, k = k0 + kk;
      float v = 0.f;
      if (n < p.N && k < K) v = to_f32<T>(tb ? B[(long long)k * ldb + n] : B[(long long)n * ldb + k]);
      Bs[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SB_K; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n < p.N) epi_store(p, m, n, acc[i][j]);
    }
  }
}

int gemm_simt(const bfk_gemm_args* a, cudaStream_t s) {
  EpiParams p;
  epi_from_args(a, p);
  dim3 grid((a->N + SB_N - 1) / SB_N, (a->M + SB_M - 1) / SB_M);
  if (a->in_dtype == BFK_BF16)
    gemm_simt_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)a->A, a->lda, a->trans_a,
                                                         (const __nv_bfloat16*)a->B, a->ldb, a->trans_b, a->K, p);
  else
    gemm_simt_kernel<float><<<grid, 256, 0, s>>>((const float*)a->A, a->lda, a->trans_a, (const float*)a->B, a->ldb,
                                                 a->trans_b, a->K, p);
  BFK_CHECK_LAUNCH("gemm_simt");
  return BFK_OK;
}

// ---- column sums (bias gradients) --------------------------------------------------------------
constexpr int CS_ROWS = 128;  // rows per first-stage block

template <typename T>
__global__ void colsum_stage1(const T* __restrict__ X, int M, int N, long long ld, float* __restrict__ partials) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int r0 = blockIdx.y * CS_ROWS, r1 = min(M, r0 + CS_ROWS);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;   // independent chains keep several loads in flight
  int r = r0;
#pragma unroll 2
  for (; r + 4 <= r1; r += 4) {
    s0 += to_f32<T>(X[(long long)r * ld + n]);
    s1 += to_f32<T>(X[(long long)(r + 1) * ld + n]);
    s2 += to_f32<T>(X[(long long)(r + 2) * ld + n]);
    s3 += to_f32<T>(X[(long long)(r + 3) * ld + n]);
  }
  for (; r < r1; ++r) s0 += to_f32<T>(X[(long long)r * ld + n]);
  partials[(long long)blockIdx.y * N + n] = (s0 + s1) + (s2 + s3);
}
// Vector variant: a thread owns 16 bytes of a row (4 fp32 / 8 bf16 columns); a block = 32 column groups x 8 row slices, every thread
// keeps four rows (64 bytes) in flight; the slices are combined in shared memory in a fixed order (deterministic).  The scalar
// kernel (8-16 bytes in flight per thread) reached 3.5-3.8 TB/s on the bias-gradient shapes.
template <typename T>
__global__ void __launch_bounds__(256) colsum_stage1_vec(const T* __restrict__ X, int M, int N, long long ld, float* __restrict__ partials) {
  constexpr int V = 16 / sizeof(T);
  __shared__ float red[8][32 * V + 4];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = (blockIdx.x * 32 + tx) * V;
  const int r0 = blockIdx.y * CS_ROWS, r1 = min(M, r0 + CS_ROWS);
  float s[V];
#pragma unroll
  for (int v = 0; v < V; ++v) s[v] = 0.f;
  if (n < N) {
    auto add = [&](const uint4& u) {
      if constexpr (sizeof(T) == 4) {
        s[0] += __uint_as_float(u.x); s[1] += __uint_as_float(u.y); s[2] += __uint_as_float(u.z); s[3] += __uint_as_float(u.w);
      } else {
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
          s[2 * j] += f.x; s[2 * j + 1] += f.y;
        }
      }
    };
    int r = r0 + ty;
    for (; r + 24 < r1; r += 32) {          // rows r, r+8, r+16, r+24 of this slice: four loads in flight
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = __ldg(reinterpret_cast<const uint4*>(X + (long long)(r + 8 * k) * ld + n));
#pragma unroll
      for (int k = 0; k < 4; ++k) add(u[k]);
    }
    for (; r < r1; r += 8) add(__ldg(reinterpret_cast<const uint4*>(X + (long long)r * ld + n)));
  }
#pragma unroll
  for (int v = 0; v < V; ++v) red[ty][tx * V + v] = s[v];
  __syncthreads();
  for (int c = threadIdx.x; c < 32 * V; c += 256) {
    const int col = blockIdx.x * 32 * V + c;
    if (col < N) {
      float t = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) t += red[k][c];
      partials[(long long)blockIdx.y * N + col] = t;
    }
  }
}
// 32 columns x 8 row-slices per block; fixed summation order (deterministic)
__global__ void __launch_bounds__(256) colsum_stage2(const float* __restrict__ partials, int nb, int N, float* __restrict__ out) {
  __shared__ float red[8][33];
  const int col = threadIdx.x & 31, slice = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + col;
  float s = 0.f;
  if (n < N)
    for (int b = slice; b < nb; b += 8) s += partials[(long long)b * N + n];
  red[slice][col] = s;
  __syncthreads();
  if (slice == 0 && n < N) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][col];
    out[n] = t;
  }
}

}  // namespace bfk

extern "C" size_t bfk_colsum_workspace_floats(int32_t M, int32_t N) {
  return (size_t)((M + bfk::CS_ROWS - 1) / bfk::CS_ROWS) * (size_t)N;
}

extern "C" int32_t bfk_colsum(const void* X, int32_t dtype, int32_t M, int32_t N, int64_t ld, float* out,
                              float* partials, bfk_stream_t s_) {
  using namespace bfk;
  BFK_CHECK_ARG(X && out && partials && M > 0 && N > 0, "bfk_colsum: bad arguments");
  cudaStream_t s = to_stream(s_);
  int nb = (M + CS_ROWS - 1) / CS_ROWS;
  const int V = dtype == BFK_BF16 ? 8 : 4;
  if (N % V == 0 && ld % V == 0 && (reinterpret_cast<uintptr_t>(X) & 15u) == 0) {
    dim3 grid((N / V + 31) / 32, nb);
    if (dtype == BFK_BF16) colsum_stage1_vec<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)X, M, N, ld, partials);
    else colsum_stage1_vec<float><<<grid, 256, 0, s>>>((const float*)X, M, N, ld, partials);
  } else {
    dim3 grid((N + 127) / 128, nb);
    if (dtype == BFK_BF16) colsum_stage1<__nv_bfloat16><<<grid, 128, 0, s>>>((const __nv_bfloat16*)X, M, N, ld, partials);
    else colsum_stage1<float><<<grid, 128, 0, s>>>((const float*)X, M, N, ld, partials);
  }
  colsum_stage2<<<(N + 31) / 32, 256, 0, s>>>(partials, nb, N, out);
  BFK_CHECK_LAUNCH("bfk_colsum");
  return BFK_OK;
}
