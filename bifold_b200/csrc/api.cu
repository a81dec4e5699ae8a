// C-ABI glue: version / error reporting and backend dispatch for GEMM and attention.
#include <stdarg.h>

#include "common.cuh"

namespace bfk {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      (void)cudaGetLastError();
      n = 148;
    }
  }
  return n;
}

int gemm_simt(const bfk_gemm_args* a, cudaStream_t s);
int gemm_tc(const bfk_gemm_args* a, cudaStream_t s);          // tcgen05 + TMA (gemm_tc.cu)
bool gemm_tc_supported(const bfk_gemm_args* a, const char** why);
int attention_fwd_simt(const bfk_attention_args* a, cudaStream_t s);
int attention_bwd_simt(const bfk_attention_args* a, cudaStream_t s);
int attention_fwd_tc(const bfk_attention_args* a, cudaStream_t s);   // tensor-core path (attention_tc.cu)
int attention_bwd_tc(const bfk_attention_args* a, cudaStream_t s);
bool attention_tc_supported(const bfk_attention_args* a, bool bwd, const char** why);

}  // namespace bfk

using namespace bfk;

extern "C" int32_t bfk_version(void) { return 100; }
extern "C" const char* bfk_last_error(void) { return g_err; }
extern "C" int32_t bfk_device_sm_count(void) { return sm_count(); }

extern "C" int32_t bfk_gemm(const bfk_gemm_args* a, bfk_stream_t s_) {
  BFK_CHECK_ARG(a && a->A && a->B && a->D, "bfk_gemm: null pointer");
  BFK_CHECK_ARG(a->M > 0 && a->N > 0 && a->K > 0, "bfk_gemm: bad shape %d x %d x %d", a->M, a->N, a->K);
  BFK_CHECK_ARG(a->in_dtype == BFK_F32 || a->in_dtype == BFK_BF16, "bfk_gemm: bad in_dtype");
  BFK_CHECK_ARG(a->out_dtype == BFK_F32 || a->out_dtype == BFK_BF16, "bfk_gemm: bad out_dtype");
  BFK_CHECK_ARG(!a->accumulate || a->out_dtype == BFK_F32, "bfk_gemm: accumulate needs fp32 output");
  BFK_CHECK_ARG(a->act >= BFK_ACT_NONE && a->act <= BFK_ACT_MUL_AUX, "bfk_gemm: bad act %d", a->act);
  BFK_CHECK_ARG(!(a->act == BFK_ACT_DGELU_TANH || a->act == BFK_ACT_DGELU_ERF || a->act == BFK_ACT_MUL_AUX ||
                  a->act == BFK_ACT_GELU_TANH_GRAD || a->act == BFK_ACT_GELU_ERF_GRAD) || a->aux, "bfk_gemm: this act needs aux");
  cudaStream_t s = to_stream(s_);
  if (a->backend == BFK_BACKEND_SIMT || a->in_dtype == BFK_F32) {
    BFK_CHECK_ARG(a->backend != BFK_BACKEND_TCGEN05, "bfk_gemm: tcgen05 backend needs bf16 operands");
    return gemm_simt(a, s);
  }
  const char* why = "";
  if (gemm_tc_supported(a, &why)) return gemm_tc(a, s);
  if (a->backend == BFK_BACKEND_TCGEN05) {
    set_error("bfk_gemm: tcgen05 backend unsupported for this problem: %s", why);
    return BFK_E_UNSUPPORTED;
  }
  return gemm_simt(a, s);
}

static int check_attention(const bfk_attention_args* a, bool bwd) {
  BFK_CHECK_ARG(a && a->qkv && a->out, "bfk_attention: null pointer");
  BFK_CHECK_ARG(a->B > 0 && a->N > 0 && a->H > 0 && a->hd > 0 && a->hd <= 96, "bfk_attention: bad shape (hd <= 96)");
  BFK_CHECK_ARG(a->dtype == BFK_F32 || a->dtype == BFK_BF16, "bfk_attention: bad dtype");
  if (bwd) BFK_CHECK_ARG(a->dout && a->dqkv && a->lse && a->delta, "bfk_attention_bwd: null pointer");
  return BFK_OK;
}

extern "C" int32_t bfk_attention_fwd(const bfk_attention_args* a, bfk_stream_t s_) {
  if (int rc = check_attention(a, false)) return rc;
  cudaStream_t s = to_stream(s_);
  const char* why = "";
  if (a->backend != BFK_BACKEND_SIMT && a->dtype == BFK_BF16 && attention_tc_supported(a, false, &why)) return attention_fwd_tc(a, s);
  if (a->backend >= BFK_BACKEND_TCGEN05) { set_error("bfk_attention_fwd: tensor-core backend unsupported: %s", why); return BFK_E_UNSUPPORTED; }
  return attention_fwd_simt(a, s);
}

extern "C" int32_t bfk_attention_bwd(const bfk_attention_args* a, bfk_stream_t s_) {
  if (int rc = check_attention(a, true)) return rc;
  cudaStream_t s = to_stream(s_);
  const char* why = "";
  if (a->backend != BFK_BACKEND_SIMT && a->dtype == BFK_BF16 && attention_tc_supported(a, true, &why)) return attention_bwd_tc(a, s);
  if (a->backend >= BFK_BACKEND_TCGEN05) { set_error("bfk_attention_bwd: tensor-core backend unsupported: %s", why); return BFK_E_UNSUPPORTED; }
  return attention_bwd_simt(a, s);
}
