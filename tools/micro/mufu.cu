// MUFU.EX2 / FFMA / tcgen05.ld throughput probes (per-SM rates), sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ex2(float* out, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fma(float* out, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = 0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(0.999f), "f"(0.001f));
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  for (int warps = 4; warps <= 32; warps *= 2) {
    for (int which = 0; which < 2; ++which) {
      const int iters = 20000;
      if (which == 0) k_ex2<<<148, warps * 32>>>(d, 100); else k_fma<<<148, warps * 32>>>(d, 100);
      cudaDeviceSynchronize();
      cudaEventRecord(e0);
      if (which == 0) k_ex2<<<148, warps * 32>>>(d, iters); else k_fma<<<148, warps * 32>>>(d, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      double ops_per_sm = (double)warps * 32 * 8 * iters;
      double cycles = ms * 1e-3 * clk * 1e3;
      printf("%s warps/SM=%2d: %.2f ops/clk/SM (at %d MHz nominal), %.3f ms\n", which == 0 ? "ex2" : "fma", warps, ops_per_sm / cycles, clk / 1000, ms);
    }
  }
  return 0;
}
