// Optimizer step on the flat ScaleKD arenas: global-norm clip + AdamW as TWO launches over one contiguous fp32 buffer
// instead of hundreds of per-parameter kernels. The reference builds torch.optim.AdamW over student + loss parameters
// (train/distillation_module.py:440-502, config/config.yaml:25-30) and lets Lightning clip the global gradient norm to 1.0
// (train.py:267-268); this is the same arithmetic for the loss-module parameters, with the student's share of the norm
// passed in. (the step immediately after the hot path.)
#include "common.cuh"
#include "../../include/b200_distill.h"

namespace b200 {

__global__ void __launch_bounds__(256) sqnorm_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  float acc = 0.f;
  const long long n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x4 + i);
    acc += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n & 3)) {
    const float v = x[(n4 << 2) + threadIdx.x];
    acc += v * v;
  }
  acc = warp_sum(acc);
  __shared__ float red[8];
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 8) {
    float v = red[threadIdx.x];
    v += __shfl_xor_sync(0xffu, v, 4);
    v += __shfl_xor_sync(0xffu, v, 2);
    v += __shfl_xor_sync(0xffu, v, 1);
    if (threadIdx.x == 0) atomicAdd(out, v);
  }
}

// torch.optim.AdamW (decoupled weight decay, no amsgrad) on gradients scaled by the clip coefficient
// min(1, max_norm / (sqrt(sq[0] + sq[1]) + 1e-6)) -- torch.nn.utils.clip_grad_norm_'s formula; max_norm <= 0: no clipping.
__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, long long n,
             float lr, float beta1, float beta2, float eps, float wd, float inv_bc1, float inv_sqrt_bc2,
             const float* __restrict__ sq, const float* __restrict__ sq_extra, float max_norm) {
  pdl_trigger();
  pdl_wait();
  float coef = 1.f;
  if (max_norm > 0.f) {
    const float tot = sq[0] + (sq_extra != nullptr ? sq_extra[0] : 0.f);
    coef = fminf(1.f, max_norm / (sqrtf(tot) + 1e-6f));
  }
  const float decay = 1.f - lr * wd;
  const float step = lr * inv_bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * coef;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    p[i] = p[i] * decay - step * mi / (sqrtf(vi) * inv_sqrt_bc2 + eps);
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_sqnorm_f32(const float* x, long long n, float* out_accum, void* stream) {
  B200_CHECK_ARG(x && out_accum && n > 0, "bad args");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "x must be 16-byte aligned");
  long long blocks = cdiv(n / 4 + 1, 256);
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  B200_CUDA_OK(launch_pdl(sqnorm_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), x, n, out_accum));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1,
                               float beta2, float eps, float weight_decay, int step, const float* grad_sqnorm,
                               const float* extra_sqnorm, float max_grad_norm, void* stream) {
  B200_CHECK_ARG(p && g && m && v && n > 0 && step >= 1, "bad args");
  B200_CHECK_ARG(max_grad_norm <= 0.f || grad_sqnorm != nullptr, "clipping needs the squared gradient norm");
  const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
  long long blocks = cdiv(n, 256);
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  B200_CUDA_OK(launch_pdl(adamw_kernel, dim3((unsigned)blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), p, g, m, v, n,
                          lr, beta1, beta2, eps, weight_decay, (float)(1.0 / bc1), (float)(1.0 / sqrt(bc2)), grad_sqnorm,
                          extra_sqnorm, max_grad_norm));
  B200_LAUNCH_OK();
  return 0;
}
