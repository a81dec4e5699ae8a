// Throughput of a few instruction classes per SM and clock on the target GPU (one CTA of NW warps per SM, all SMs).
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters, long long* clk) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = 0.001f * (threadIdx.x + i + 1);
  uint32_t acc = 0;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
      if (OP == 3) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7])); acc ^= r; }
      if (OP == 4) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 5) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
      if (OP == 6) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]));
    }
  }
  long long t1 = clock64();
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)acc;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
template <int OP> void run(const char* name, int nw) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out; long long* clk; cudaMalloc(&out, sms * 1024 * 4); cudaMalloc(&clk, sms * 8);
  const int iters = 4096;
  k<OP><<<sms, 32 * nw>>>(out, iters, clk); cudaDeviceSynchronize();
  k<OP><<<sms, 32 * nw>>>(out, iters, clk); cudaDeviceSynchronize();
  long long h[256]; cudaMemcpy(h, clk, sms * 8, cudaMemcpyDeviceToHost);
  double c = 0; for (int i = 0; i < sms; ++i) c += h[i]; c /= sms;
  printf("%-10s warps/SM %2d : %.2f lane-ops / clk / SM\n", name, nw, (double)iters * 8 * 32 * nw / c);
  cudaFree(out); cudaFree(clk);
}
int main() {
  for (int nw : {4, 8, 16}) {
    run<0>("ex2", nw); run<1>("rcp", nw); run<4>("lg2", nw); run<5>("tanh", nw); run<2>("ffma", nw); run<3>("cvt.bf16x2", nw); run<6>("fmax", nw);
  }
  return 0;
}
