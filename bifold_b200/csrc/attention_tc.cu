// Tensor-core attention (placeholder until the tcgen05 kernels land): reports "unsupported" so the
// dispatcher uses the SIMT kernels.
#include "common.cuh"
namespace bfk {
bool attention_tc_supported(const bfk_attention_args*, bool, const char** why) { *why = "not built"; return false; }
int attention_fwd_tc(const bfk_attention_args*, cudaStream_t) { set_error("attention_tc: not built"); return BFK_E_UNSUPPORTED; }
int attention_bwd_tc(const bfk_attention_args*, cudaStream_t) { set_error("attention_tc: not built"); return BFK_E_UNSUPPORTED; }
}  // namespace bfk
