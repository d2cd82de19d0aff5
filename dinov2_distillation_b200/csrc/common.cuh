// Shared host/device helpers for libb200distill.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

namespace b200 {

void set_error(const std::string& msg);
void count_launch(int n = 1);
int sm_count();
// dispatch options (core.cu): read by the launchers, set through b200_set_option()
enum { OPT_PDL = 0, OPT_ATTN_TC_FWD, OPT_ATTN_TC_FWD_LONG, OPT_ATTN_TC_BWD, OPT_ATTN_BWD_FUSED, OPT_ATTN_TC_BWD_LONG,
       OPT_GEMM_V2, OPT_GEMM_BN, OPT_GEMM_2CTA, OPT_GEMM_INPLACE_RED, OPT_GEMM_EW, OPT_GEMM_DBG, OPT_ATTN_PROBE_SKIP, OPT_GEMM_LN,
       STAT_ATTN_TC_BWD, STAT_ATTN_MMA_BWD, STAT_ATTN_TC_FWD, STAT_ATTN_MMA_FWD, OPT_ATTN_PP_FWD, STAT_ATTN_PP_FWD, OPT_GEMM_AUX_DEEP, OPT_COUNT };
int option(int id);
void bump_stat(int id);
// per-launch timing hooks (no-ops unless b200_profile_enable(1)); cat: 0 GEMM, 1 attention fwd, 2 attention bwd
int prof_begin(cudaStream_t st);
void prof_end(int idx, cudaStream_t st, double flops, int cat, double bytes = 0.0);   // bytes: algorithmic operand + result traffic

#define B200_CHECK_ARG(cond, msg)                                                      \
  do {                                                                                 \
    if (!(cond)) {                                                                     \
      ::b200::set_error(std::string(__func__) + ": " + (msg) + " [" #cond "]");        \
      return -1;                                                                       \
    }                                                                                  \
  } while (0)

#define B200_CUDA_OK(expr)                                                                              \
  do {                                                                                                  \
    cudaError_t _e = (expr);                                                                            \
    if (_e != cudaSuccess) {                                                                            \
      ::b200::set_error(std::string(__func__) + ": " #expr " -> " + cudaGetErrorString(_e));            \
      return -2;                                                                                        \
    }                                                                                                   \
  } while (0)

#define B200_LAUNCH_OK()                                                                                \
  do {                                                                                                  \
    cudaError_t _e = cudaGetLastError();                                                                \
    if (_e != cudaSuccess) {                                                                            \
      ::b200::set_error(std::string(__func__) + ": kernel launch -> " + cudaGetErrorString(_e));        \
      return -3;                                                                                        \
    }                                                                                                   \
    ::b200::count_launch();                                                                             \
  } while (0)

#define B200_TRY(expr)        \
  do {                        \
    int _r = (expr);          \
    if (_r != 0) return _r;   \
  } while (0)

static inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(h);
}

// 16-bit packing with a runtime element type: fp16 (ScaleKD projector forward operands) or bf16 (everything else)
__device__ __forceinline__ uint32_t pack16(float a, float b, int fp16) {
  if (fp16) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
  }
  return pack_bf16(a, b);
}
__device__ __forceinline__ float2 unpack16(uint32_t u, int fp16) {
  if (fp16) return __half22float2(*reinterpret_cast<__half2*>(&u));
  return unpack_bf16(u);
}
__device__ __forceinline__ void store16(__nv_bfloat16* dst, float v, int fp16) {
  if (fp16) *reinterpret_cast<__half*>(dst) = __float2half_rn(v);
  else *dst = __float2bfloat16(v);
}
__device__ __forceinline__ float load16(const __nv_bfloat16* src, int fp16) {
  if (fp16) return __half2float(*reinterpret_cast<const __half*>(src));
  return __bfloat162float(*src);
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_erf(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}

// erf to ~2e-7 absolute (Abramowitz-Stegun 7.1.26) with the two transcendental steps on the MUFU approximations
// (rcp.approx / ex2.approx: relative error ~2^-22, far below the 16-bit rounding of every consumer): 2 MUFU + 9 FMA-class
// instructions, against ~40 for erff(). The GEMM epilogue is instruction bound, so this is what GELU costs.
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float erf_fast(float x) {
  const float ax = fabsf(x);
  const float t = rcp_approx(fmaf(0.3275911f, ax, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = ex2_approx(-1.4426950408889634f * ax * ax);
  return copysignf(fmaf(-poly, e, 1.0f), x);
}
__device__ __forceinline__ float gelu_fast(float x) { return 0.5f * x * (1.0f + erf_fast(x * 0.70710678118654752f)); }
__device__ __forceinline__ float dgelu_fast(float x) {
  const float cdf = 0.5f * (1.0f + erf_fast(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * ex2_approx(-0.72134752044448170f * x * x);
  return cdf + x * pdf;
}

// ---- programmatic dependent launch (PDL): a kernel launched through launch_pdl() may be scheduled while its predecessor
// in the stream is still draining; it MUST call pdl_wait() before its first access to global memory the predecessor may
// have written (or may still read). pdl_trigger() lets the NEXT kernel start being scheduled; single-wave / persistent
// kernels call it early, multi-wave kernels leave it to the implicit trigger at exit (early dependents would only sit
// on SM resources their own later waves need). B200_PDL=0 turns the launch attribute off (plain stream order).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
}

// Multi-job parameter preparation (elementwise.cu: param_prep_kernel)
enum { PREP_CAST16 = 0, PREP_SPLIT3_RIGHT = 1, PREP_COPY32 = 2, PREP_TRANSPOSE16 = 3, PREP_TRANSPOSE32 = 4 };
constexpr int PREP_MAX_JOBS = 20;
struct PrepJob {
  const float* src;
  void* dst;
  int type, rows, cols, fp16;   // flat jobs cover rows*cols elements; transposes read [rows, cols]
  long long out_ld;             // transposes: output row pitch (elements)
};
struct PrepJobs {
  PrepJob j[PREP_MAX_JOBS];
  int n;
  void add(int type, const float* src, void* dst, int rows, int cols, int fp16 = 0, long long out_ld = 0) {
    if (n < PREP_MAX_JOBS) j[n] = PrepJob{src, dst, type, rows, cols, fp16, out_ld};
    ++n;
  }
};
int launch_param_prep(const PrepJobs& jobs, cudaStream_t st);

// Dual-format variants of three elementwise entries (elementwise.cu): the ScaleKD projector forward saves each activation
// a wgrad GEMM will need in fp16 (forward operand) AND bf16 (backward operand) from the pass that produces it --
// tcgen05 kind::f16 cannot mix the two formats in one product, and a separate conversion pass costs a launch each.
int cast_f32_f16_dual(const float* x, void* y_f16, void* y_bf16, long long n, void* stream);
int layernorm_fwd_dual(const float* x, const float* w, const float* b, float eps, float* y_f32, void* y16, void* y16_alt,
                       float* mean, float* rstd, int rows, int D, int in_period, int in_pad, int y16_is_fp16,
                       void* stream);
// NCHW fp32 -> bf16 tokens [B*HW, C] + 3-term fp16 split [B*HW, 3C] ([hi|hi|lo]) in one pass (elementwise.cu)
int tokenize_split3(const float* x, void* xt_bf16, void* xt3_fp16, int B, int C, int HW, void* stream);
// b200_bn_finalize that also bumps nn.BatchNorm2d.num_batches_tracked (int64 device scalar, may be NULL)
int bn_finalize_counted(const float* sums, float* mean, float* rstd, float* running_mean, float* running_var,
                        float momentum, float eps, int M, int D, long long* num_batches_tracked, void* stream);
// b200_bn_relu_pos_bwd_apply that also adds sums2 (d beta | d gamma) into the BN affine gradient buffers
int bn_relu_pos_bwd_apply_acc(const float* dz, const float* y, const float* mean, const float* rstd, const float* w,
                              const float* b, const float* sums2, void* dy_bf16, int use_batch_stats, int M, int D,
                              float* acc_bn_b, float* acc_bn_w, void* stream);
int bn_relu_pos_fwd_dual(const float* y, const float* mean, const float* rstd, const float* w, const float* b,
                         const float* pos, float* z_f32, void* z16, void* z16_alt, int M, int D, int HW,
                         int z16_is_fp16, void* stream);

// Bump allocator over a caller-provided workspace (256-byte aligned carve-outs).
struct Arena {
  uint8_t* base;
  size_t cap;
  size_t off;
  Arena(void* p, size_t bytes) : base(static_cast<uint8_t*>(p)), cap(bytes), off(0) {}
  void* take(size_t bytes) {
    size_t a = (off + 255) & ~size_t(255);
    off = a + bytes;
    return base ? base + a : nullptr;
  }
  template <typename T>
  T* take_n(size_t n) { return static_cast<T*>(take(n * sizeof(T))); }
  bool ok() const { return base == nullptr || off <= cap; }
  size_t used() const { return (off + 255) & ~size_t(255); }
};

}  // namespace b200
