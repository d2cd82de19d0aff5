// ScaleKD mimicking terms (losses/scalekd.py:67-92 get_spat_loss, :95-127 get_freq_loss) as single-pass
// token-major kernels: one warp per token row, warp-shuffle reductions, fp32 throughout.
//
// Frequency term: the reference computes idct2(zero_dc(dct2(x))) per (b, channel) plane. Zeroing the DC coefficient of
// an orthogonal-up-to-scaling separable DCT-II removes exactly the plane's mean, so the term equals the spatial term
// applied to x - mean_{H,W}(x).  b200_dct_zero_dc_idct is the explicit transform kept as a cross-check.
#include "common.cuh"
#include "../../include/b200_distill.h"

namespace b200 {

int zero_f32(float* p, long long n, cudaStream_t st);

// ws layout (floats): KD_SLOTS x [loss sum, sim sum] (blocks spread their atomics over the slots: a thousand blocks on
// one address serialise for longer than the kernel reads its 50 MB), then ms[B*D], mt[B*D], md[B*D]
constexpr int KD_SLOTS = 32;
static inline long long ws_acc_floats() { return 2 * KD_SLOTS; }

// mean over the HW tokens of each image: out[b, c] = 1/HW sum_p x[b, skip + p, c]
// (32 x 16 threads, eight rows in flight per thread: one block per (image, 128 columns) has to cover the latency itself)
__global__ void __launch_bounds__(512)
token_mean_kernel(const float* __restrict__ x, float* __restrict__ out, int HW, int D, long long bstride, int skip) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  __shared__ float4 sh[16][32];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int b = blockIdx.y;
  float4 a = make_float4(0, 0, 0, 0);
  if (c < D) {
    const float* base = x + (long long)b * bstride + (long long)skip * D + c;
#pragma unroll 8
    for (int p = threadIdx.y; p < HW; p += 16) {
      const float4 v = *reinterpret_cast<const float4*>(base + (long long)p * D);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
  }
  sh[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < D) {
#pragma unroll
    for (int k = 1; k < 16; ++k) {
      const float4 t = sh[k][threadIdx.x];
      a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
    }
    const float inv = 1.0f / (float)HW;
    *reinterpret_cast<float4*>(out + (long long)b * D + c) = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
  }
}

struct RowStats { float ss, tt, st; };

__device__ __forceinline__ RowStats row_stats(const float* __restrict__ s, const float* __restrict__ t,
                                              const float* __restrict__ ms, const float* __restrict__ mt, int D,
                                              int lane) {
  float ss = 0.f, tt = 0.f, st = 0.f;
  for (int c = lane * 4; c < D; c += 128) {
    float4 a = *reinterpret_cast<const float4*>(s + c);
    float4 b = *reinterpret_cast<const float4*>(t + c);
    if (ms) {
      const float4 m = *reinterpret_cast<const float4*>(ms + c);
      const float4 n = *reinterpret_cast<const float4*>(mt + c);
      a.x -= m.x; a.y -= m.y; a.z -= m.z; a.w -= m.w;
      b.x -= n.x; b.y -= n.y; b.z -= n.z; b.w -= n.w;
    }
    ss += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
    tt += (b.x * b.x + b.y * b.y) + (b.z * b.z + b.w * b.w);
    st += (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
  }
  RowStats r;
  r.ss = warp_sum(ss);
  r.tt = warp_sum(tt);
  r.st = warp_sum(st);
  return r;
}

__global__ void __launch_bounds__(256)
kd_loss_fwd_kernel(const float* __restrict__ S, const float* __restrict__ T, int rows, int HW, int D, int Nt,
                   int t_skip, const float* __restrict__ ms, const float* __restrict__ mt, float* __restrict__ acc) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  float loss = 0.f, sim = 0.f;
  // two rows per trip: the loads of both are issued before either row's three warp reductions
  const int step = gridDim.x * wpb;
  for (int r = blockIdx.x * wpb + wid; r < rows; r += 2 * step) {
    RowStats q[2];
    const int r2 = r + step;
    const bool two = r2 < rows;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const int rr = (k == 0 || two) ? (k == 0 ? r : r2) : r;
      const int b = rr / HW, p = rr - b * HW;
      q[k] = row_stats(S + (long long)rr * D, T + ((long long)b * Nt + t_skip + p) * D, ms ? ms + (long long)b * D : nullptr,
                       ms ? mt + (long long)b * D : nullptr, D, lane);
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      if (k == 1 && !two) break;
      // F.normalize(eps=1e-12) then MSE(sum) and cosine_similarity(eps=1e-8) on the normalised vectors
      const float ns = fmaxf(sqrtf(q[k].ss), 1e-12f), nt = fmaxf(sqrtf(q[k].tt), 1e-12f);
      const float hs2 = q[k].ss / (ns * ns), ht2 = q[k].tt / (nt * nt), hst = q[k].st / (ns * nt);
      loss += hs2 + ht2 - 2.f * hst;
      sim += hst / (fmaxf(sqrtf(hs2), 1e-8f) * fmaxf(sqrtf(ht2), 1e-8f));
    }
  }
  __shared__ float sl[8], sm[8];
  if (lane == 0) { sl[wid] = loss; sm[wid] = sim; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, c = 0.f;
    for (int k = 0; k < wpb; ++k) { a += sl[k]; c += sm[k]; }
    float* slot = acc + 2 * (blockIdx.x % KD_SLOTS);
    atomicAdd(slot, a);
    atomicAdd(slot + 1, c);
  }
}

__global__ void kd_loss_finalize_kernel(const float* __restrict__ acc, float* __restrict__ out, float loss_scale,
                                        float sim_scale) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  float l = 0.f, c = 0.f;
  for (int k = 0; k < KD_SLOTS; ++k) { l += acc[2 * k]; c += acc[2 * k + 1]; }
  out[0] = l * loss_scale;
  out[1] = c * sim_scale;
}

__global__ void __launch_bounds__(256)
kd_loss_bwd_kernel(const float* __restrict__ S, const float* __restrict__ T, int rows, int HW, int D, int Nt,
                   int t_skip, const float* __restrict__ ms, const float* __restrict__ mt,
                   const float* __restrict__ g_loss, const float* __restrict__ g_sim, float loss_scale, float sim_scale,
                   float* __restrict__ dS) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const float gl = g_loss[0] * loss_scale, gs = (g_sim != nullptr ? g_sim[0] : 0.0f) * sim_scale;
  for (int r = blockIdx.x * wpb + wid; r < rows; r += gridDim.x * wpb) {
    const int b = r / HW, p = r - b * HW;
    const float* s = S + (long long)r * D;
    const float* t = T + ((long long)b * Nt + t_skip + p) * D;
    const float* msb = ms ? ms + (long long)b * D : nullptr;
    const float* mtb = ms ? mt + (long long)b * D : nullptr;
    const RowStats q = row_stats(s, t, msb, mtb, D, lane);
    const float ns = fmaxf(sqrtf(q.ss), 1e-12f), nt = fmaxf(sqrtf(q.tt), 1e-12f);
    const float ins = 1.f / ns, intt = 1.f / nt;
    const float hs2 = q.ss * ins * ins, hst = q.st * ins * intt;
    // d/ds sum (s^ - t^)^2 = 2/ns [ (s^ - t^) - s^ (|s^|^2 - s^.t^) ] ;  d/ds cos = 1/ns [ t^ - cos s^ ]
    const float k1 = hs2 - hst;
    float* o = dS + (long long)r * D;
    for (int c = lane * 4; c < D; c += 128) {
      float4 a = *reinterpret_cast<const float4*>(s + c);
      float4 bt = *reinterpret_cast<const float4*>(t + c);
      if (msb) {
        const float4 m = *reinterpret_cast<const float4*>(msb + c);
        const float4 n = *reinterpret_cast<const float4*>(mtb + c);
        a.x -= m.x; a.y -= m.y; a.z -= m.z; a.w -= m.w;
        bt.x -= n.x; bt.y -= n.y; bt.z -= n.z; bt.w -= n.w;
      }
      const float sh[4] = {a.x * ins, a.y * ins, a.z * ins, a.w * ins};
      const float th[4] = {bt.x * intt, bt.y * intt, bt.z * intt, bt.w * intt};
      float g[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        g[e] = gl * 2.f * ins * ((sh[e] - th[e]) - sh[e] * k1) + gs * ins * (th[e] - hst * sh[e]);
      }
      *reinterpret_cast<float4*>(o + c) = make_float4(g[0], g[1], g[2], g[3]);
    }
  }
}

// dS[b, p, c] -= md[b, c]
__global__ void __launch_bounds__(256)
sub_token_mean_kernel(float* __restrict__ dS, const float* __restrict__ md, long long rows, int HW, int D) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  const int D4 = D >> 2;
  const long long n4 = rows * D4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / D4;
    const int c4 = (int)(i - r * D4);
    const long long b = r / HW;
    float4 v = reinterpret_cast<float4*>(dS)[i];
    const float4 m = reinterpret_cast<const float4*>(md + b * D)[c4];
    v.x -= m.x; v.y -= m.y; v.z -= m.z; v.w -= m.w;
    reinterpret_cast<float4*>(dS)[i] = v;
  }
}

// explicit separable DCT-II (x2 per axis, unnormalised) -> zero DC -> inverse; 8 channels of one image per block.
__global__ void __launch_bounds__(256)
dct_zero_dc_idct_kernel(const float* __restrict__ x, float* __restrict__ y, int R, int D, long long x_bs,
                        long long x_ts) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  extern __shared__ float smem[];
  const int RR = R * R;
  float* A = smem;             // [RR][8]
  float* Bf = A + RR * 8;      // [RR][8]
  float* Cf = Bf + RR * 8;     // [R][R] forward  Cf[k][n] = 2 cos(pi (2n+1) k / 2R)
  float* Ci = Cf + RR;         // [R][R] inverse  Ci[k][n] = w_k / R cos(pi (2n+1) k / 2R), w_0 = 1/2
  const int b = blockIdx.y, c0 = blockIdx.x * 8;
  const float PI = 3.14159265358979323846f;
  for (int i = threadIdx.x; i < RR; i += blockDim.x) {
    const int k = i / R, n = i - k * R;
    const float cv = cosf(PI * (2 * n + 1) * k / (2.f * R));
    Cf[i] = 2.f * cv;
    Ci[i] = (k == 0 ? 0.5f : 1.f) / (float)R * cv;
  }
  for (int i = threadIdx.x; i < RR * 8; i += blockDim.x) {
    const int tkn = i >> 3, c = i & 7;
    A[i] = (c0 + c < D) ? x[(long long)b * x_bs + (long long)tkn * x_ts + c0 + c] : 0.f;
  }
  __syncthreads();
  // pass 1: Bf[i][k] = sum_j A[i][j] Cf[k][j]
  for (int i = threadIdx.x; i < RR * 8; i += blockDim.x) {
    const int c = i & 7, ik = i >> 3, ii = ik / R, k = ik - ii * R;
    float acc = 0.f;
    for (int j = 0; j < R; ++j) acc += A[(ii * R + j) * 8 + c] * Cf[k * R + j];
    Bf[i] = acc;
  }
  __syncthreads();
  // pass 2: A[k2][k] = sum_i Bf[i][k] Cf[k2][i]
  for (int i = threadIdx.x; i < RR * 8; i += blockDim.x) {
    const int c = i & 7, kk = i >> 3, k2 = kk / R, k = kk - k2 * R;
    float acc = 0.f;
    for (int ii = 0; ii < R; ++ii) acc += Bf[(ii * R + k) * 8 + c] * Cf[k2 * R + ii];
    A[i] = (kk == 0) ? 0.f : acc;  // zero the DC coefficient
  }
  __syncthreads();
  // pass 3: Bf[k2][j] = sum_k A[k2][k] Ci[k][j]
  for (int i = threadIdx.x; i < RR * 8; i += blockDim.x) {
    const int c = i & 7, kj = i >> 3, k2 = kj / R, j = kj - k2 * R;
    float acc = 0.f;
    for (int k = 0; k < R; ++k) acc += A[(k2 * R + k) * 8 + c] * Ci[k * R + j];
    Bf[i] = acc;
  }
  __syncthreads();
  // pass 4: out[i][j] = sum_k2 Bf[k2][j] Ci[k2][i]
  for (int i = threadIdx.x; i < RR * 8; i += blockDim.x) {
    const int c = i & 7, ij = i >> 3, ii = ij / R, j = ij - ii * R;
    float acc = 0.f;
    for (int k2 = 0; k2 < R; ++k2) acc += Bf[(k2 * R + j) * 8 + c] * Ci[k2 * R + ii];
    if (c0 + c < D) y[((long long)b * RR + ij) * D + c0 + c] = acc;
  }
}

}  // namespace b200

using namespace b200;

extern "C" long long b200_kd_loss_ws_floats(int B, int HW, int D) {
  (void)HW;
  return ws_acc_floats() + 3LL * B * D;
}

extern "C" int b200_kd_loss_fwd(const float* S, const float* T, int B, int HW, int D, int Nt, int t_skip, int freq,
                                float alpha, float* out, float* ws, void* stream) {
  B200_CHECK_ARG(S && T && out && ws && B > 0 && HW > 0 && D > 0, "bad args");
  B200_CHECK_ARG(D % 4 == 0, "D must be a multiple of 4");
  B200_CHECK_ARG(Nt >= HW + t_skip, "teacher token count too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B200_TRY(zero_f32(ws, ws_acc_floats(), st));
  float* ms = ws + ws_acc_floats();
  float* mt = ms + (long long)B * D;
  if (freq) {
    dim3 grid((unsigned)cdiv(D, 128), (unsigned)B);
    B200_CUDA_OK(launch_pdl(token_mean_kernel, dim3(grid), dim3(dim3(32, 16)), 0, st, S, ms, HW, D, (long long)HW * D, 0));
    B200_LAUNCH_OK();
    B200_CUDA_OK(launch_pdl(token_mean_kernel, dim3(grid), dim3(dim3(32, 16)), 0, st, T, mt, HW, D, (long long)Nt * D, t_skip));
    B200_LAUNCH_OK();
  }
  const int rows = B * HW;
  long long g = cdiv(rows, 8);
  if (g > (long long)sm_count() * 8) g = (long long)sm_count() * 8;
  B200_CUDA_OK(launch_pdl(kd_loss_fwd_kernel, dim3((unsigned)g), dim3(256), 0, st, S, T, rows, HW, D, Nt, t_skip, freq ? ms : nullptr,
                                                  freq ? mt : nullptr, ws));
  B200_LAUNCH_OK();
  B200_CUDA_OK(launch_pdl(kd_loss_finalize_kernel, dim3(1), dim3(1), 0, st, ws, out, alpha / (float)B, 1.0f / (float)rows));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_kd_loss_bwd(const float* S, const float* T, int B, int HW, int D, int Nt, int t_skip, int freq,
                                float alpha, const float* g_out, float* dS, int accumulate, float* ws, void* stream) {
  B200_CHECK_ARG(g_out != nullptr, "bad args");
  return b200_kd_loss_bwd_split(S, T, B, HW, D, Nt, t_skip, freq, alpha, g_out, g_out + 1, dS, accumulate, ws, stream);
}

extern "C" int b200_kd_loss_bwd_split(const float* S, const float* T, int B, int HW, int D, int Nt, int t_skip, int freq,
                                      float alpha, const float* g_loss, const float* g_sim, float* dS, int accumulate,
                                      float* ws, void* stream) {
  B200_CHECK_ARG(S && T && g_loss && dS && ws && B > 0 && HW > 0 && D > 0 && D % 4 == 0, "bad args");
  B200_CHECK_ARG(!accumulate, "accumulate is not supported (autograd sums the branches)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* ms = ws + ws_acc_floats();
  float* mt = ms + (long long)B * D;
  float* md = mt + (long long)B * D;
  const int rows = B * HW;
  long long g = cdiv(rows, 8);
  if (g > (long long)sm_count() * 8) g = (long long)sm_count() * 8;
  B200_CUDA_OK(launch_pdl(kd_loss_bwd_kernel, dim3((unsigned)g), dim3(256), 0, st, S, T, rows, HW, D, Nt, t_skip, freq ? ms : nullptr,
                                                  freq ? mt : nullptr, g_loss, g_sim, alpha / (float)B, 1.0f / (float)rows, dS));
  B200_LAUNCH_OK();
  if (freq) {
    dim3 grid((unsigned)cdiv(D, 128), (unsigned)B);
    B200_CUDA_OK(launch_pdl(token_mean_kernel, dim3(grid), dim3(dim3(32, 16)), 0, st, dS, md, HW, D, (long long)HW * D, 0));
    B200_LAUNCH_OK();
    long long n4 = (long long)rows * D / 4;
    long long gg = cdiv(n4, 256);
    if (gg > (long long)sm_count() * 8) gg = (long long)sm_count() * 8;
    B200_CUDA_OK(launch_pdl(sub_token_mean_kernel, dim3((unsigned)gg), dim3(256), 0, st, dS, md, rows, HW, D));
    B200_LAUNCH_OK();
  }
  return 0;
}

extern "C" int b200_dct_zero_dc_idct(const float* x, float* y, int B, int R, int D, long long x_bs, long long x_ts,
                                     void* stream) {
  B200_CHECK_ARG(x && y && B > 0 && R > 0 && R <= 64 && D > 0, "bad args");
  const size_t smem = (size_t(R) * R * 16 + size_t(R) * R * 2) * sizeof(float);
  B200_CHECK_ARG(smem <= 220 * 1024, "resolution too large");
  if (smem > 48 * 1024) {
    static size_t set_to = 0;
    if (set_to < smem) {
      B200_CUDA_OK(cudaFuncSetAttribute(dct_zero_dc_idct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      set_to = smem;
    }
  }
  dim3 grid((unsigned)cdiv(D, 8), (unsigned)B);
  B200_CUDA_OK(launch_pdl(dct_zero_dc_idct_kernel, dim3(grid), dim3(256), smem, static_cast<cudaStream_t>(stream), x, y, R, D, x_bs, x_ts));
  B200_LAUNCH_OK();
  return 0;
}
