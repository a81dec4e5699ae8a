// Device helpers shared by the tcgen05 attention kernels (attention_tc.cu: generic flash kernels; attention_v2.cu: persistent
// whole-K/V-resident kernels for short sequences): UMMA descriptors of the 128-byte-swizzled tiles, P-tile stores, row stores.
#pragma once
#include "tc_common.cuh"

namespace bfk {

constexpr uint32_t AT_ATOM = 128 * 128;  // one 128-row x 128-byte swizzle atom column
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;

// K-major operand (rows x head_dim or rows x 128 keys), atoms stacked every AT_ATOM bytes; k-th UMMA_K slice
__device__ __forceinline__ uint64_t kmajor_desc(uint32_t tile, int k, uint32_t atom = AT_ATOM) {
  return make_sdesc(tile + (uint32_t)(k >> 2) * atom + (uint32_t)(k & 3) * 32u, 16, 1024);
}
// MN-major view of a [128 rows][NA*64 cols] tile: reduction over rows (slice k = rows 16k..16k+15), MN = columns
__device__ __forceinline__ uint64_t mnmajor_desc(uint32_t tile, int k, uint32_t atom = AT_ATOM) {
  return make_sdesc(tile + (uint32_t)k * 2048u, atom, 1024);
}
template <int NA>
__device__ __forceinline__ void load_tile(uint32_t dst, const CUtensorMap* map, uint32_t bar, int slot, int row0, int b,
                                          uint32_t atom = AT_ATOM) {
#pragma unroll
  for (int a = 0; a < NA; ++a) tma_load_4d(dst + a * atom, map, bar, 64 * a, slot, row0, b);
}
// write 8 bf16 (one 16-byte chunk `c16` of 16 per 128-key row) of row r into a K-major swizzled [128 x 128] tile
__device__ __forceinline__ void store_p_chunk(uint32_t tile, int r, int c16, const float (&v)[8]) {
  __nv_bfloat162 h0 = __floats2bfloat162_rn(v[0], v[1]), h1 = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 h2 = __floats2bfloat162_rn(v[4], v[5]), h3 = __floats2bfloat162_rn(v[6], v[7]);
  uint32_t addr = tile + (uint32_t)(c16 >> 3) * AT_ATOM + (uint32_t)r * 128u + (uint32_t)(((c16 & 7) ^ (r & 7)) << 4);
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(*reinterpret_cast<uint32_t*>(&h0)),
               "r"(*reinterpret_cast<uint32_t*>(&h1)), "r"(*reinterpret_cast<uint32_t*>(&h2)),
               "r"(*reinterpret_cast<uint32_t*>(&h3))
               : "memory");
}
// store 16 consecutive head-dim columns [d0, d0+16) of one row as bf16 (16-byte pieces, clipped at hd)
__device__ __forceinline__ void store_row16(__nv_bfloat16* dst, int d0, int hd, const uint32_t (&r)[16], float mul) {
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    if (d0 + half * 8 + 8 <= hd) {
      __nv_bfloat162 h[4];
#pragma unroll
      for (int t = 0; t < 4; ++t)
        h[t] = __floats2bfloat162_rn(__uint_as_float(r[half * 8 + 2 * t]) * mul, __uint_as_float(r[half * 8 + 2 * t + 1]) * mul);
      *reinterpret_cast<uint4*>(dst + d0 + half * 8) = *reinterpret_cast<uint4*>(h);
    }
  }
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// width of key block j: full 128, or the valid tail rounded up to a multiple of 16 (MMA N / K granularity)
__device__ __forceinline__ int kv_cols(int N, int j, int bw = 128) {
  const int rem = N - j * bw;
  return rem >= bw ? bw : ((rem + 15) & ~15);
}


// 4-D bf16 tensor map {hd, slots, N, B} over a [B*N, ld] row-major matrix whose row holds `slots` runs of hd elements; box =
// {64, 1, box_rows, 1} with 128-byte swizzle (zero fill beyond hd and beyond the sample's last token).  attention_tc.cu.
bool make_map4(CUtensorMap* out, const void* ptr, int hd, int slots, int N, int B, long long ld, int box_rows);

}  // namespace bfk
