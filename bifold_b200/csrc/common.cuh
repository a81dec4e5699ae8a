// Shared helpers for the bifold_b200 kernels (sm_100a).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/bifold_b200.h"

namespace bfk {

void set_error(const char* fmt, ...);

#define BFK_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      ::bfk::set_error(__VA_ARGS__);             \
      return BFK_E_INVALID;                      \
    }                                            \
  } while (0)

#define BFK_CHECK_LAUNCH(name)                                                       \
  do {                                                                               \
    cudaError_t e__ = cudaPeekAtLastError();                                         \
    if (e__ != cudaSuccess) {                                                        \
      ::bfk::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));      \
      return BFK_E_CUDA;                                                             \
    }                                                                                \
  } while (0)

inline cudaStream_t to_stream(bfk_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
int sm_count();

// ---- element access templated on storage type ------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Programmatic dependent launch (PDL).  A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while
// its predecessor in the stream still runs: it sets itself up (barriers, TMEM, tensor-map prefetch) and then blocks in pdl_wait()
// until the predecessor has COMPLETED and its writes are visible; nothing the predecessor produces may be touched before.
// pdl_trigger() in the predecessor lets the dependent's CTAs be scheduled as soon as every CTA of the predecessor has started
// and SM resources free up (without it: when the predecessor's last CTA exits).  Both are no-ops in ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// BFK_PDL=1: the tcgen05 GEMM launches carry the attribute (host side)
inline bool pdl_enabled() {
  static const bool v = [] { const char* e = getenv("BFK_PDL"); return e && e[0] == '1'; }();
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ---- activations -------------------------------------------------------------------------------
// torch gelu(approximate="tanh"): 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k = 0.7978845608028654f, c = 0.044715f;
  return 0.5f * x * (1.f + tanhf(k * (x + c * x * x * x)));
}
__device__ __forceinline__ float dgelu_tanh(float x) {
  const float k = 0.7978845608028654f, c = 0.044715f;
  float t = tanhf(k * (x + c * x * x * x));
  return 0.5f * (1.f + t) + 0.5f * x * (1.f - t * t) * k * (1.f + 3.f * c * x * x);
}
// torch gelu(): 0.5 x (1 + erf(x / sqrt(2)))
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.7071067811865476f)); }
__device__ __forceinline__ float dgelu_erf(float x) {
  return 0.5f * (1.f + erff(x * 0.7071067811865476f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}

// Counter-based keep-mask for LoRA dropout: hash of (seed, row, col / 4) -> four 16-bit uniforms in [0,1), one per column of
// the aligned group of four (kernels that walk rows four columns at a time evaluate the 64-bit mix once per group: the
// compiler merges the identical sub-expressions).  Drop probability resolution: 2^-16.
__device__ __forceinline__ uint64_t hash_mix(uint64_t seed, uint32_t row, uint32_t col4) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (((uint64_t)row << 32) | col4);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float hash_uniform(uint64_t seed, uint32_t row, uint32_t col) {
  const uint64_t z = hash_mix(seed, row, col >> 2);
  return (float)((uint32_t)(z >> (16 * (col & 3))) & 0xFFFFu) * (1.0f / 65536.0f);
}

}  // namespace bfk
