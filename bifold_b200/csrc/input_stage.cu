// GPU-side input stage: the two per-sample host loops of the reference's Processor that dominate the
// host->device bytes of a training batch, as HBM-bound kernels.
//   gauss_targets : bifold/data/processor.py:405-417 (_make_gaussmap, strategy "first") — 8 integers per bimanual sample
//                   instead of four H x W fp32 maps over PCIe (51 MB -> 2 KB at B = 64, 224 x 224)
//   normalize_u8  : bifold/data/processor.py:143-148 (_process_rgb -> SigLIP image processor: rescale 1/255, normalise with
//                   mean/std, HWC -> CHW) — uint8 frames over PCIe instead of fp32 (4x fewer bytes)
#include "common.cuh"

namespace bfk {

// One block per map, P candidate points per map.  exp(-((x-cx)^2 + (y-cy)^2) / (2 s^2)) is separable: P (H + W) double-precision
// exponentials per map, P double multiply-adds per pixel, rounded once to fp32 — the reference evaluates the map in float64 and
// casts (torch.FloatTensor), so the fp32 values agree to the last bit except for double-rounding ties.
//   one valid point  : strategy "first" (processor.py:412-417), no normalisation
//   several points   : strategy "gmm" (processor.py:418-425): sum of Gaussians / (2 pi s^2), divided by the map maximum
// A point with x == INT32_MIN is padding; a map whose first point is padding is "no action for this arm": all zeros
// (processor.py:296, _make_gaussmap()).
__global__ void __launch_bounds__(256) gauss_targets_kernel(const int32_t* __restrict__ centers, int P, int H, int W, double inv_two_sigma2,
                                                            double norm, float* __restrict__ out) {
  extern __shared__ double s_e[];       // [P][W] ex, then [P][H] ey
  __shared__ double s_red[8];
  __shared__ int s_np;
  const long long map = blockIdx.x;
  const int32_t* c = centers + map * 2 * P;
  float* o = out + map * (long long)H * W;
  if (threadIdx.x == 0) {
    int np = 0;
    while (np < P && c[2 * np] != INT32_MIN) ++np;
    s_np = np;
  }
  __syncthreads();
  const int np = s_np;
  double* ex = s_e;
  double* ey = s_e + (size_t)P * W;
  for (int i = threadIdx.x; i < np * (W + H); i += blockDim.x) {
    const int pt = i / (W + H), r = i - pt * (W + H);
    const double d = r < W ? (double)(r - c[2 * pt]) : (double)(r - W - c[2 * pt + 1]);
    const double e = exp(-(d * d) * inv_two_sigma2);
    if (r < W) ex[pt * W + r] = e; else ey[pt * H + r - W] = e;
  }
  __syncthreads();
  auto value = [&](int y, int x) -> double {
    if (np == 1) return ey[y] * ex[x];
    double v = 0.0;
    for (int pt = 0; pt < np; ++pt) v += ey[pt * H + y] * ex[pt * W + x] * norm;
    return v;
  };
  double inv_max = 1.0;
  if (np > 1) {                          // gmm: normalise by the maximum of the map
    double m = 0.0;
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) m = fmax(m, value(i / W, i % W));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = s_red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, s_red[w]);
    inv_max = m;                         // divide (not multiply by a reciprocal): same rounding as gauss_map /= gauss_map.max()
  }
  const bool vec = (W & 3) == 0;
  const int W4 = W >> 2;
  if (vec) {
    for (int i = threadIdx.x; i < H * W4; i += blockDim.x) {
      const int y = i / W4, x = (i - y * W4) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (np == 1) v = make_float4((float)value(y, x), (float)value(y, x + 1), (float)value(y, x + 2), (float)value(y, x + 3));
      else if (np > 1) v = make_float4((float)(value(y, x) / inv_max), (float)(value(y, x + 1) / inv_max), (float)(value(y, x + 2) / inv_max),
                                       (float)(value(y, x + 3) / inv_max));
      __stcs(reinterpret_cast<float4*>(o) + i, v);
    }
  } else {
    for (int i = threadIdx.x; i < H * W; i += blockDim.x) {
      const int y = i / W, x = i - y * W;
      o[i] = np == 0 ? 0.f : (np == 1 ? (float)value(y, x) : (float)(value(y, x) / inv_max));
    }
  }
}

// uint8 [n, H, W, 3] (PIL / numpy frame layout) -> fp32 [n, 3, H, W]: (x * (1/255) - mean) / std.
// One thread handles 4 consecutive pixels of one frame: 12 bytes in (three 32-bit loads), one float4 per channel out.
__global__ void __launch_bounds__(256) normalize_u8_kernel(const uint8_t* __restrict__ in, long long n_quads, int HW, float mean, float inv_std,
                                                           float* __restrict__ out) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_quads) return;
  const int qpf = HW >> 2;                        // quads per frame
  const long long f = q / qpf;
  const int p = (int)(q - f * qpf) * 4;
  const uint32_t* src = reinterpret_cast<const uint32_t*>(in + (f * HW + p) * 3);
  const uint32_t w0 = __ldcs(src), w1 = __ldcs(src + 1), w2 = __ldcs(src + 2);
  uint8_t b[12];
#pragma unroll
  for (int i = 0; i < 4; ++i) { b[i] = (w0 >> (8 * i)) & 255u; b[4 + i] = (w1 >> (8 * i)) & 255u; b[8 + i] = (w2 >> (8 * i)) & 255u; }
  const float k = 1.0f / 255.0f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float4 v;
    v.x = ((float)b[c] * k - mean) * inv_std;
    v.y = ((float)b[3 + c] * k - mean) * inv_std;
    v.z = ((float)b[6 + c] * k - mean) * inv_std;
    v.w = ((float)b[9 + c] * k - mean) * inv_std;
    *reinterpret_cast<float4*>(out + (f * 3 + c) * HW + p) = v;
  }
}

}  // namespace bfk

using namespace bfk;

extern "C" int32_t bfk_gauss_targets(const int32_t* centers, int32_t n_maps, int32_t n_points, int32_t H, int32_t W, float sigma, float* out,
                                     bfk_stream_t s_) {
  BFK_CHECK_ARG(centers && out, "bfk_gauss_targets: null pointer");
  BFK_CHECK_ARG(n_maps > 0 && n_points > 0 && H > 0 && W > 0 && sigma > 0.f, "bfk_gauss_targets: bad shape / sigma");
  const size_t smem = (size_t)n_points * (H + W) * sizeof(double);
  BFK_CHECK_ARG(smem <= 48 * 1024, "bfk_gauss_targets: n_points * (H + W) too large (%d x %d)", n_points, H + W);
  const double s = (double)sigma;
  gauss_targets_kernel<<<n_maps, 256, smem, to_stream(s_)>>>(centers, n_points, H, W, 1.0 / (2.0 * s * s),
                                                              1.0 / (2.0 * 3.141592653589793 * s * s), out);
  BFK_CHECK_LAUNCH("bfk_gauss_targets");
  return BFK_OK;
}

extern "C" int32_t bfk_normalize_u8(const uint8_t* frames, int64_t n_frames, int32_t H, int32_t W, float mean, float std_, float* out,
                                    bfk_stream_t s_) {
  BFK_CHECK_ARG(frames && out, "bfk_normalize_u8: null pointer");
  BFK_CHECK_ARG(n_frames > 0 && H > 0 && W > 0 && std_ > 0.f, "bfk_normalize_u8: bad shape / std");
  BFK_CHECK_ARG(((long long)H * W) % 4 == 0, "bfk_normalize_u8: H*W must be a multiple of 4");
  BFK_CHECK_ARG((reinterpret_cast<uintptr_t>(frames) & 3u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0, "bfk_normalize_u8: misaligned buffers");
  const long long quads = n_frames * ((long long)H * W / 4);
  normalize_u8_kernel<<<(unsigned)((quads + 255) / 256), 256, 0, to_stream(s_)>>>(frames, quads, H * W, mean, 1.0f / std_, out);
  BFK_CHECK_LAUNCH("bfk_normalize_u8");
  return BFK_OK;
}
