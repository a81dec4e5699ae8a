// LoRA factor gradients of one tower layer in three launches (bf16 operands, tcgen05):
//   P1 = dqkv^T u   [3D, 16]   split-K partials   (rows [0,D) x cols [0,r) -> dB_q, rows [2D,3D) x cols [8,8+r) -> dB_v)
//   P2 = du^T  x    [16, D]    split-K partials   (rows [0,r) -> dA_q, rows [8,8+r) -> dA_v)
//   one reduction kernel sums the partials in a fixed order and writes the four gradients.
// peft LoRA on q_proj / v_proj:  y = W h + (alpha/r) B (A drop(h));  u = A drop(h) for q (cols 0..r-1) and v (cols 8..8+r-1).
#include "common.cuh"

namespace bfk {

int gemm_tc_wgrad_partials(const bfk_gemm_args* a, cudaStream_t s, int* splits);   // gemm_tc.cu
bool gemm_tc_supported(const bfk_gemm_args* a, const char** why);

__global__ void __launch_bounds__(256) lora_grads_reduce_kernel(const float* __restrict__ p1, int s1, const float* __restrict__ p2, int s2,
                                                                int D, int r, float scale, float* __restrict__ dBq,
                                                                float* __restrict__ dBv, float* __restrict__ dAq,
                                                                float* __restrict__ dAv) {
  const int per = D * r;                       // elements of each of the four outputs
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= 4 * per) return;
  const int which = i / per, e = i % per;
  float acc = 0.f;
  if (which < 2) {                             // dB[d, j]  <-  P1[(which ? 2D : 0) + d][(which ? 8 : 0) + j]
    const int d = e / r, j = e % r;
    const long long off = (long long)((which ? 2 * D : 0) + d) * 16 + (which ? 8 : 0) + j, stride = (long long)3 * D * 16;
    for (int k = 0; k < s1; ++k) acc += p1[k * stride + off];
    (which ? dBv : dBq)[e] = scale * acc;
  } else {                                     // dA[j, c]  <-  P2[(which == 3 ? 8 : 0) + j][c]
    const int j = e / D, c = e % D;
    const long long off = (long long)((which == 3 ? 8 : 0) + j) * D + c, stride = (long long)16 * D;
    for (int k = 0; k < s2; ++k) acc += p2[k * stride + off];
    (which == 3 ? dAv : dAq)[e] = acc;
  }
}

}  // namespace bfk

using namespace bfk;

extern "C" size_t bfk_lora_grads_workspace_bytes(int32_t D) {
  return ((size_t)16 * 3 * D * 16 + (size_t)32 * 16 * D) * sizeof(float);   // up to 16 / 32 K-splits
}

extern "C" int32_t bfk_lora_grads(const bfk_lora_grads_args* a, bfk_stream_t s_) {
  BFK_CHECK_ARG(a && a->dqkv && a->u && a->du && a->x && a->dBq && a->dBv && a->dAq && a->dAv && a->workspace,
                "bfk_lora_grads: null pointer");
  BFK_CHECK_ARG(a->M > 0 && a->D > 0 && a->r > 0 && a->r <= 8, "bfk_lora_grads: need 0 < r <= 8 (M=%d D=%d r=%d)", a->M, a->D, a->r);
  BFK_CHECK_ARG((size_t)a->workspace_bytes >= bfk_lora_grads_workspace_bytes(a->D), "bfk_lora_grads: workspace too small");
  cudaStream_t s = to_stream(s_);
  float* p1 = (float*)a->workspace;
  float* p2 = p1 + (size_t)16 * 3 * a->D * 16;
  bfk_gemm_args g = {};
  g.in_dtype = BFK_BF16; g.out_dtype = BFK_F32; g.alpha = 1.f; g.trans_a = 1; g.trans_b = 1; g.backend = BFK_BACKEND_TCGEN05;
  // P1 [3D, 16] = dqkv^T u
  g.M = 3 * a->D; g.N = 16; g.K = a->M;
  g.A = a->dqkv; g.lda = a->ld_dqkv; g.B = a->u; g.ldb = a->ld_u;
  g.D = p1; g.ldd = 16; g.workspace = p1; g.workspace_bytes = (int64_t)16 * 3 * a->D * 16 * 4; g.split_k = 0;
  const char* why = "";
  if (!gemm_tc_supported(&g, &why)) { set_error("bfk_lora_grads: %s", why); return BFK_E_UNSUPPORTED; }
  int s1 = 1, s2 = 1;
  if (int rc = gemm_tc_wgrad_partials(&g, s, &s1)) return rc;
  // P2 [16, D] = du^T x
  g.M = 16; g.N = a->D; g.K = a->M;
  g.A = a->du; g.lda = a->ld_du; g.B = a->x; g.ldb = a->ld_x;
  g.D = p2; g.ldd = a->D; g.workspace = p2; g.workspace_bytes = (int64_t)32 * 16 * a->D * 4; g.split_k = 32;
  if (!gemm_tc_supported(&g, &why)) { set_error("bfk_lora_grads: %s", why); return BFK_E_UNSUPPORTED; }
  if (int rc = gemm_tc_wgrad_partials(&g, s, &s2)) return rc;
  const int total = 4 * a->D * a->r;
  lora_grads_reduce_kernel<<<(total + 255) / 256, 256, 0, s>>>(p1, s1, p2, s2, a->D, a->r, a->scale, a->dBq, a->dBv, a->dAq, a->dAv);
  BFK_CHECK_LAUNCH("bfk_lora_grads");
  return BFK_OK;
}
