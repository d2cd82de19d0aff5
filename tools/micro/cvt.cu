// Which pipe do the 16-bit pack conversions run on? Per-SM rates of cvt.rn.{f16x2,bf16x2}.f32 alone and mixed with
// ex2.approx (2 ex2 : 1 cvt, the softmax ratio). sm_100a.   nvcc -arch=sm_100a -O3 -o cvt cvt.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters) {
  float a[8];
  unsigned acc = 0;
  for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      if (MODE == 0 || MODE == 3 || MODE == 4 || MODE == 5) {
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i + 1]));
      }
      unsigned r = 0;
      if (MODE == 1 || MODE == 3) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[i + 1]));
      if (MODE == 2 || MODE == 4) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[i + 1]));
      if (MODE == 5) {   // integer round-to-nearest-up bf16 pack: 2 IADD + 1 PRMT
        unsigned x = __float_as_uint(a[i]) + 0x8000u, y = __float_as_uint(a[i + 1]) + 0x8000u;
        asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(x), "r"(y));
      }
      acc ^= r;
    }
  }
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + __uint_as_float(acc);
}
template <int MODE> void run(const char* name, float* d, int clk) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int warps = 16, iters = 20000;
  k<MODE><<<148, warps * 32>>>(d, 100);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<148, warps * 32>>>(d, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double groups_per_sm = (double)warps * 32 * 4 * iters;   // one group = (2 ex2) and/or (1 pack)
  double cycles = ms * 1e-3 * clk * 1e3;
  printf("%-44s %.2f groups/clk/SM  (%.3f ms)\n", name, groups_per_sm / cycles, ms);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  run<0>("2 x ex2", d, clk);
  run<1>("1 x cvt.rn.f16x2.f32", d, clk);
  run<2>("1 x cvt.rn.bf16x2.f32", d, clk);
  run<3>("2 x ex2 + cvt.rn.f16x2.f32", d, clk);
  run<4>("2 x ex2 + cvt.rn.bf16x2.f32", d, clk);
  run<5>("2 x ex2 + integer bf16 pack (2 IADD + PRMT)", d, clk);
  return 0;
}
