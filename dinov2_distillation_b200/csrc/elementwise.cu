// Bandwidth-bound kernels of the distillation path: casts / layout changes, patch im2col, LayerNorm fwd/bwd,
// BatchNorm(+ReLU+pos) fwd/bwd, column reductions, SwiGLU gate.  All vectorised (128-bit) and coalesced on the
// contiguous channel dimension; reductions use warp shuffles + one atomic per block column.
#include "common.cuh"
#include <stdlib.h>
#include "../../include/b200_distill.h"

namespace b200 {

static inline int grid_for(long long work_items, int per_block, int max_blocks_per_sm = 8) {
  long long g = cdiv(work_items, per_block);
  long long cap = (long long)sm_count() * max_blocks_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------------------------------------ cast
// y: 16-bit copy in the format `fp16` selects; y_alt (optional): a second copy in the OTHER 16-bit format
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                     __nv_bfloat16* __restrict__ y_alt, long long n, int fp16) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    uint2 u;
    u.x = pack16(v.x, v.y, fp16);
    u.y = pack16(v.z, v.w, fp16);
    reinterpret_cast<uint2*>(y)[i] = u;
    if (y_alt) {
      u.x = pack16(v.x, v.y, !fp16);
      u.y = pack16(v.z, v.w, !fp16);
      reinterpret_cast<uint2*>(y_alt)[i] = u;
    }
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    store16(y + i, x[i], fp16);
    if (y_alt) store16(y_alt + i, x[i], !fp16);
  }
}

__global__ void cast_f16_bf16_kernel(const __half* __restrict__ x, __nv_bfloat16* __restrict__ y, long long n) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  const long long n8 = n >> 3;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += stride) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(x) + i);
    const uint32_t in[4] = {u.x, u.y, u.z, u.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&in[e]));
      o[e] = pack_bf16(f.x, f.y);
    }
    reinterpret_cast<uint4*>(y)[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
  for (long long i = (n8 << 3) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    y[i] = __float2bfloat16(__half2float(x[i]));
}

// 3-term split cast: x = hi + lo with hi = r16(x), lo = r16(x - hi) (r16 = round to fp16 or bf16).
// out row = [hi | hi | lo] (left operand) or [hi | lo | hi] (right operand): a plain 16-bit GEMM over the 3K-long
// contraction then yields hi*hi + hi*lo + lo*hi, i.e. ~2x the mantissa bits. Used for the one GEMM whose result feeds
// BatchNorm -> ReLU with no residual around it (conv1x1 of proj_student), where mask flips dominate the gradient error.
__global__ void split3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, long long rows, int K,
                              int right, int fp16) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  const int K4 = K >> 2;
  const long long n4 = rows * K4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / K4;
    const int c4 = (int)(i - r * K4);
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    const float f[4] = {v.x, v.y, v.z, v.w};
    float hi[4], lo[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      hi[e] = fp16 ? __half2float(__float2half_rn(f[e])) : __bfloat162float(__float2bfloat16(f[e]));
      lo[e] = f[e] - hi[e];
    }
    uint2 uh, ul;
    uh.x = pack16(hi[0], hi[1], fp16); uh.y = pack16(hi[2], hi[3], fp16);
    ul.x = pack16(lo[0], lo[1], fp16); ul.y = pack16(lo[2], lo[3], fp16);
    __nv_bfloat16* o = out + r * 3 * K + c4 * 4;
    *reinterpret_cast<uint2*>(o) = uh;
    *reinterpret_cast<uint2*>(o + K) = right ? ul : uh;
    *reinterpret_cast<uint2*>(o + 2 * K) = right ? uh : ul;
  }
}

// ------------------------------------------------------------------------------------------------ transposes
// out[c, r] = in[r, c] * row_scale[r]   (fp32 -> bf16), batch via blockIdx.z
template <typename TOut, bool ACC>
__global__ void transpose_kernel(const float* __restrict__ in, TOut* __restrict__ out, int rows, int cols,
                                 const float* __restrict__ row_scale, long long in_bstride, long long out_bstride,
                                 float* __restrict__ out2_f32, long long out_ld, int fp16) {
  pdl_wait();   // PDL: launched through launch_pdl(); multi-wave grid, no early trigger
  __shared__ float tile[32][33];
  in += (long long)blockIdx.z * in_bstride;
  out += (long long)blockIdx.z * out_bstride;
  if (out2_f32) out2_f32 += (long long)blockIdx.z * out_bstride;
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < rows && c < cols) {
      v = in[(long long)r * cols + c];
      if (row_scale) v *= __ldg(row_scale + r);
    }
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (r < rows && c < cols) {
      const float v = tile[threadIdx.x][j];
      const long long o = (long long)c * out_ld + r;
      if constexpr (sizeof(TOut) == 2) {
        store16(reinterpret_cast<__nv_bfloat16*>(out) + o, v, fp16);
        if (out2_f32) out2_f32[o] = v;
      } else {
        if (ACC) out[o] += v; else out[o] = v;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ student tokenisation
// NCHW fp32 [B, C, HW] -> token-major working copies of the projector input in ONE pass: xt bf16 [B*HW, C] (operand of
// the conv wgrad) and the 3-term fp16 split xt3 [B*HW, 3C] = [hi | hi | lo] (left operand of the split conv GEMM, see
// split3_kernel). Replaces transpose (bf16 + fp32 token copies) followed by split3 over the fp32 copy: 12 instead of
// 20 bytes of traffic per element. 64 channels x 32 positions per block; every store is a full 128-byte row segment.
__global__ void __launch_bounds__(256) tokenize_split3_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xt,
                                                             __nv_bfloat16* __restrict__ xt3, int C, int HW) {
  pdl_wait();   // multi-wave grid: no early trigger
  __shared__ float tile[64][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 64, p0 = blockIdx.x * 32;
  const float* src = x + (long long)b * C * HW;
#pragma unroll
  for (int j = threadIdx.y; j < 64; j += 8) {
    const int c = c0 + j, p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && p < HW) ? src[(long long)c * HW + p] : 0.f;
  }
  __syncthreads();
  const int c = c0 + 2 * threadIdx.x;
  if (c >= C) return;
#pragma unroll
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int p = p0 + j;
    if (p >= HW) continue;
    const float v0 = tile[2 * threadIdx.x][j], v1 = tile[2 * threadIdx.x + 1][j];
    const long long row = (long long)b * HW + p;
    *reinterpret_cast<uint32_t*>(xt + row * C + c) = pack_bf16(v0, v1);
    const float h0 = __half2float(__float2half_rn(v0)), h1 = __half2float(__float2half_rn(v1));
    const uint32_t hi = pack16(h0, h1, 1), lo = pack16(v0 - h0, v1 - h1, 1);
    __nv_bfloat16* o = xt3 + row * 3 * C + c;
    *reinterpret_cast<uint32_t*>(o) = hi;
    *reinterpret_cast<uint32_t*>(o + C) = hi;
    *reinterpret_cast<uint32_t*>(o + 2 * C) = lo;
  }
}

// Same for channel counts that are a multiple of 4 (every student tap of the model zoo): 128 channels x 32 positions per
// block, the tile kept position-major in shared memory so that a thread reads its four channels with one 16-byte load
// and writes 8-byte words -- each warp store is a 256-byte run of one token row (the 64-channel version above moves 4
// bytes per thread and store: 57-59 % of the copy bandwidth, measured with tools/hbm_bench.py).
__global__ void __launch_bounds__(256) tokenize_split3_wide_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xt,
                                                                  __nv_bfloat16* __restrict__ xt3, int C, int HW) {
  pdl_wait();   // multi-wave grid: no early trigger
  __shared__ __align__(16) float tile[32][132];   // [position][channel], pitch 132: 16-byte aligned rows
  const int b = blockIdx.z, c0 = blockIdx.y * 128, p0 = blockIdx.x * 32;
  const float* src = x + (long long)b * C * HW;
  const int p = p0 + threadIdx.x;
#pragma unroll 8
  for (int j = threadIdx.y; j < 128; j += 8) {
    const int c = c0 + j;
    tile[threadIdx.x][j] = (c < C && p < HW) ? src[(long long)c * HW + p] : 0.f;
  }
  __syncthreads();
  const int c = c0 + 4 * threadIdx.x;
  if (c >= C) return;
#pragma unroll
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int pp = p0 + j;
    if (pp >= HW) continue;
    const float4 v = *reinterpret_cast<const float4*>(&tile[j][4 * threadIdx.x]);
    const long long row = (long long)b * HW + pp;
    *reinterpret_cast<uint2*>(xt + row * C + c) = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    const float h0 = __half2float(__float2half_rn(v.x)), h1 = __half2float(__float2half_rn(v.y));
    const float h2 = __half2float(__float2half_rn(v.z)), h3 = __half2float(__float2half_rn(v.w));
    const uint2 hi = make_uint2(pack16(h0, h1, 1), pack16(h2, h3, 1));
    const uint2 lo = make_uint2(pack16(v.x - h0, v.y - h1, 1), pack16(v.z - h2, v.w - h3, 1));
    __nv_bfloat16* o = xt3 + row * 3 * C + c;
    *reinterpret_cast<uint2*>(o) = hi;
    *reinterpret_cast<uint2*>(o + C) = hi;
    *reinterpret_cast<uint2*>(o + 2 * C) = lo;
  }
}

// ------------------------------------------------------------------------------------------------ patch im2col
// One block per (image, patch row, channel), no staging: every thread produces output column PAIRS -- one aligned 8-byte
// load (two horizontally adjacent pixels; 14 is even, so a pair never leaves its patch row) and one 4-byte store. Lanes
// run over consecutive pairs of one patch's 196-column run of the channel (coalesced 128-byte stores; the block of channel
// 2 also writes the zero padding up to Kp); the loads of a warp touch ~5 image rows x 56 bytes whose sector remainders
// belong to the neighbouring patches handled by the same block, so they hit in L1 and HBM sees each pixel once. All
// loads of a thread are independent and there is no barrier: the earlier version staged the 3 x 14 x W strip in shared
// memory (load phase, __syncthreads, store phase; 2 blocks per SM at 518 pixels) and measured 1.3-2.1 TB/s.
__global__ void __launch_bounds__(256)
patch_im2col_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int H, int W, int Kp) {
  pdl_wait();   // PDL: launched through launch_pdl(); multi-wave grid, no early trigger
  const int Wp = W / 14, Hp = H / 14;
  const int c = blockIdx.x % 3;
  const int bp = blockIdx.x / 3;
  const int b = bp / Hp, ph = bp % Hp;
  const float* base = img + (((long long)b * 3 + c) * H + ph * 14) * W;
  uint2* obase = reinterpret_cast<uint2*>(out + ((long long)(b * Hp + ph) * Wp) * Kp + c * 196);
  const int n4 = c == 2 ? (Kp - 392) >> 2 : 49;   // 4-column groups of this block per patch (channel 2: + the zero padding)
  const int Kp4 = Kp >> 2;
  const int total = Wp * n4;
#pragma unroll 4
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int pw = idx / n4, k4 = idx - pw * n4;
    uint2 v = make_uint2(0u, 0u);
    if (k4 < 49) {   // two column pairs = two aligned 8-byte loads (a group of 4 may straddle two image rows: 14 = 3.5 x 4)
      const int ka = 2 * k4, kb = 2 * k4 + 1;
      const int ia = ka / 7, ja = ka - ia * 7, ib = kb / 7, jb = kb - ib * 7;
      const float2 pa = __ldg(reinterpret_cast<const float2*>(base + (long long)ia * W + pw * 14) + ja);
      const float2 pb = __ldg(reinterpret_cast<const float2*>(base + (long long)ib * W + pw * 14) + jb);
      v = make_uint2(pack_bf16(pa.x, pa.y), pack_bf16(pb.x, pb.y));
    }
    obase[(long long)pw * Kp4 + k4] = v;
  }
}

__global__ void write_cls_rows_kernel(float* __restrict__ x, const float* __restrict__ cls,
                                      const float* __restrict__ pos, int B, int N, int D) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, d = i - b * D;
  x[(long long)b * N * D + d] = cls[d] + pos[d];
}

// ------------------------------------------------------------------------------------------------ LayerNorm
constexpr int LN_MAX_V4 = 12;  // up to D = 1536 held in registers (one warp per row)

template <int NV>
__global__ void __launch_bounds__(256)
layernorm_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b, float eps,
                     float* __restrict__ y32, __nv_bfloat16* __restrict__ y16, __nv_bfloat16* __restrict__ y16_alt,
                     float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows, int D, int in_period,
                     int in_pad, int fp16) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const float inv_d = 1.0f / (float)D;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps_per_block) {
    long long ir = r;
    if (in_period > 0) ir = (long long)(r / in_period) * (in_period + in_pad) + in_pad + r % in_period;
    const float4* xr = reinterpret_cast<const float4*>(x + ir * D);
    float4 v[NV];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = i * 32 + lane;
      if (c4 * 4 < D) {
        v[i] = xr[c4];
        s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
      } else {
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    const float mean = warp_sum(s) * inv_d;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = i * 32 + lane;
      if (c4 * 4 < D) {
        const float a = v[i].x - mean, bq = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += (a * a + bq * bq) + (c * c + d * d);
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * inv_d + eps);
    if (lane == 0) {
      if (mean_out) mean_out[r] = mean;
      if (rstd_out) rstd_out[r] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = i * 32 + lane;
      if (c4 * 4 < D) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(w) + c4);
        const float4 be = __ldg(reinterpret_cast<const float4*>(b) + c4);
        float4 o;
        o.x = (v[i].x - mean) * rstd * g.x + be.x;
        o.y = (v[i].y - mean) * rstd * g.y + be.y;
        o.z = (v[i].z - mean) * rstd * g.z + be.z;
        o.w = (v[i].w - mean) * rstd * g.w + be.w;
        if (y32) reinterpret_cast<float4*>(y32 + (long long)r * D)[c4] = o;
        if (y16) {
          uint2 u;
          u.x = pack16(o.x, o.y, fp16);
          u.y = pack16(o.z, o.w, fp16);
          reinterpret_cast<uint2*>(y16 + (long long)r * D)[c4] = u;
        }
        if (y16_alt) {   // second copy in the other 16-bit format (bf16 operand of the wgrad GEMM)
          uint2 u;
          u.x = pack16(o.x, o.y, !fp16);
          u.y = pack16(o.z, o.w, !fp16);
          reinterpret_cast<uint2*>(y16_alt + (long long)r * D)[c4] = u;
        }
      }
    }
  }
}

// WG: the launch needs column partials (dgamma / dbeta and / or the column sums of the output). The plain form (re-used
// teacher blocks: no parameter gradients) drops those 9 * NV accumulator registers: 89 -> ~56 registers, four blocks per
// SM instead of two, i.e. twice the loads in flight for a kernel that is pure HBM traffic (3.1 -> ... TB/s, profiles/).
template <int NV, bool WG>
__global__ void __launch_bounds__(256, WG ? 2 : 4)
layernorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
                     const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ dres,
                     float* __restrict__ dx, __nv_bfloat16* __restrict__ dx16, float* __restrict__ dw,
                     float* __restrict__ db, float* __restrict__ dx_colsum, int rows, int D) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const float inv_d = 1.0f / (float)D;
  float4 aw[WG ? NV : 1], ab[WG ? NV : 1], ao[WG ? NV : 1];   // column partials: dw, db and (dx_colsum) the output itself
#pragma unroll
  for (int i = 0; i < (WG ? NV : 1); ++i) { aw[i] = make_float4(0, 0, 0, 0); ab[i] = make_float4(0, 0, 0, 0); ao[i] = make_float4(0, 0, 0, 0); }
  // The accumulating form keeps two blocks per SM (one atomic per column per block), so each warp keeps its NEXT row's
  // loads in flight while it reduces the current one (register double buffer); the plain form has the occupancy instead.
  const int r_first = blockIdx.x * warps_per_block + (threadIdx.x >> 5), r_step = gridDim.x * warps_per_block;
  float4 nd[WG ? NV : 1], nx[WG ? NV : 1];
  if (WG && r_first < rows) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = i * 32 + lane;
      if (c4 * 4 < D) {
        nd[i] = reinterpret_cast<const float4*>(dy + (long long)r_first * D)[c4];
        nx[i] = reinterpret_cast<const float4*>(x + (long long)r_first * D)[c4];
      }
    }
  }
  for (int r = r_first; r < rows; r += r_step) {
    const float4* dyr = reinterpret_cast<const float4*>(dy + (long long)r * D);
    const float4* xr = reinterpret_cast<const float4*>(x + (long long)r * D);
    const float mu = mean[r], rs = rstd[r];
    float4 g[NV], xh[NV];
    float4 cd[WG ? NV : 1], cx[WG ? NV : 1];
    if (WG) {
#pragma unroll
      for (int i = 0; i < NV; ++i) { cd[i] = nd[i]; cx[i] = nx[i]; }
      const int rn = r + r_step;
      if (rn < rows) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int c4 = i * 32 + lane;
          if (c4 * 4 < D) {
            nd[i] = reinterpret_cast<const float4*>(dy + (long long)rn * D)[c4];
            nx[i] = reinterpret_cast<const float4*>(x + (long long)rn * D)[c4];
          }
        }
      }
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = i * 32 + lane;
      if (c4 * 4 < D) {
        const float4 d4 = WG ? cd[i] : dyr[c4];
        const float4 x4 = WG ? cx[i] : xr[c4];
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(w) + c4);
        xh[i] = make_float4((x4.x - mu) * rs, (x4.y - mu) * rs, (x4.z - mu) * rs, (x4.w - mu) * rs);
        g[i] = make_float4(d4.x * w4.x, d4.y * w4.y, d4.z * w4.z, d4.w * w4.w);
        s1 += (g[i].x + g[i].y) + (g[i].z + g[i].w);
        s2 += (g[i].x * xh[i].x + g[i].y * xh[i].y) + (g[i].z * xh[i].z + g[i].w * xh[i].w);
        if (WG && dw) {
          aw[i].x += d4.x * xh[i].x; aw[i].y += d4.y * xh[i].y; aw[i].z += d4.z * xh[i].z; aw[i].w += d4.w * xh[i].w;
          ab[i].x += d4.x; ab[i].y += d4.y; ab[i].z += d4.z; ab[i].w += d4.w;
        }
      }
    }
    s1 = warp_sum(s1) * inv_d;
    s2 = warp_sum(s2) * inv_d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = i * 32 + lane;
      if (c4 * 4 < D) {
        float4 o;
        o.x = rs * (g[i].x - s1 - xh[i].x * s2);
        o.y = rs * (g[i].y - s1 - xh[i].y * s2);
        o.z = rs * (g[i].z - s1 - xh[i].z * s2);
        o.w = rs * (g[i].w - s1 - xh[i].w * s2);
        if (dres) {
          const float4 rr = reinterpret_cast<const float4*>(dres + (long long)r * D)[c4];
          o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
        }
        if (WG && dx_colsum) { ao[i].x += o.x; ao[i].y += o.y; ao[i].z += o.z; ao[i].w += o.w; }
        if (dx) reinterpret_cast<float4*>(dx + (long long)r * D)[c4] = o;
        if (dx16) {
          uint2 u;
          u.x = pack_bf16(o.x, o.y);
          u.y = pack_bf16(o.z, o.w);
          reinterpret_cast<uint2*>(dx16 + (long long)r * D)[c4] = u;
        }
      }
    }
  }
  if constexpr (!WG) return;
  if (dw) {
    // block reduction over warps through shared memory, then one atomic per column per block
    extern __shared__ float red[];  // [warps][2][D]  (D <= 1536, warps = 8 -> 96 KB max; sized by the launcher)
    const int wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = i * 32 + lane;
      if (c4 * 4 < D) {
        reinterpret_cast<float4*>(red + (long long)(wid * 2) * D)[c4] = aw[i];
        reinterpret_cast<float4*>(red + (long long)(wid * 2 + 1) * D)[c4] = ab[i];
      }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float sw = 0.f, sb = 0.f;
      for (int k = 0; k < warps_per_block; ++k) {
        sw += red[(long long)(k * 2) * D + c];
        sb += red[(long long)(k * 2 + 1) * D + c];
      }
      atomicAdd(dw + c, sw);
      atomicAdd(db + c, sb);
    }
    if (dx_colsum) __syncthreads();
  }
  if (dx_colsum) {
    // column sums of the OUTPUT (the bias gradient of the linear layer that produced this LayerNorm's input)
    extern __shared__ float red[];
    const int wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = i * 32 + lane;
      if (c4 * 4 < D) reinterpret_cast<float4*>(red + (long long)wid * D)[c4] = ao[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
      float so = 0.f;
      for (int k = 0; k < warps_per_block; ++k) so += red[(long long)k * D + c];
      atomicAdd(dx_colsum + c, so);
    }
  }
}

// Accumulating backward for wide rows (D >= 768: NV >= 6). The register form above keeps 9 * NV accumulator registers plus a
// double-buffered row per lane -- fine at NV = 3 (ViT-S), 1.2-2.4 KB of spills per thread at NV = 6 / 8 / 12 (measured:
// 17 % of the copy bandwidth at D = 1024). Here the row is read twice (statistics, then outputs: the second read is an
// L1 / L2 hit, HBM sees each byte once) so no row values live across the warp reduction, and the column partials live in
// per-warp shared-memory strips [warp][dw | db | colsum][D] that each lane updates at its own columns (no conflicts, no
// synchronisation until the final cross-warp reduction).
template <int NV>
__global__ void __launch_bounds__(256, 2)
layernorm_bwd_wide_kernel(const float* __restrict__ dy, const float* __restrict__ x, const float* __restrict__ w,
                          const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ dres,
                          float* __restrict__ dx, __nv_bfloat16* __restrict__ dx16, float* __restrict__ dw,
                          float* __restrict__ db, float* __restrict__ dx_colsum, int rows, int D) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float red[];   // [warps][K][D], K = 2 (dw, db) and / or 1 (colsum)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int warps_per_block = blockDim.x >> 5;
  const int K = (dw ? 2 : 0) + (dx_colsum ? 1 : 0);
  const float inv_d = 1.0f / (float)D;
  float4* s_w = reinterpret_cast<float4*>(red + (size_t)(wid * K) * D);
  float4* s_b = reinterpret_cast<float4*>(red + (size_t)(wid * K + 1) * D);
  float4* s_o = reinterpret_cast<float4*>(red + (size_t)(wid * K + (dw ? 2 : 0)) * D);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c4 = i * 32 + lane;
    if (c4 * 4 < D) {
      if (dw) { s_w[c4] = make_float4(0, 0, 0, 0); s_b[c4] = make_float4(0, 0, 0, 0); }
      if (dx_colsum) s_o[c4] = make_float4(0, 0, 0, 0);
    }
  }
  const int r_step = gridDim.x * warps_per_block;
  for (int r = blockIdx.x * warps_per_block + wid; r < rows; r += r_step) {
    const float4* dyr = reinterpret_cast<const float4*>(dy + (long long)r * D);
    const float4* xr = reinterpret_cast<const float4*>(x + (long long)r * D);
    const float mu = mean[r], rs = rstd[r];
    // pass 1: the two row sums (all 2 * NV loads of the lane are issued before the first is consumed)
    float s1 = 0.f, s2 = 0.f;
    {
      float4 d4[NV], x4[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c4 = i * 32 + lane;
        if (c4 * 4 < D) { d4[i] = dyr[c4]; x4[i] = xr[c4]; }
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c4 = i * 32 + lane;
        if (c4 * 4 < D) {
          const float4 w4 = __ldg(reinterpret_cast<const float4*>(w) + c4);
          const float gx = d4[i].x * w4.x, gy = d4[i].y * w4.y, gz = d4[i].z * w4.z, gw = d4[i].w * w4.w;
          s1 += (gx + gy) + (gz + gw);
          s2 += (gx * (x4[i].x - mu) + gy * (x4[i].y - mu)) + (gz * (x4[i].z - mu) + gw * (x4[i].w - mu));
        }
      }
    }
    s1 = warp_sum(s1) * inv_d;
    s2 = warp_sum(s2) * rs * inv_d;
    // pass 2: outputs and column partials (row re-read: cache hit)
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int c4 = i * 32 + lane;
      if (c4 * 4 < D) {
        const float4 d4 = dyr[c4], x4 = xr[c4];
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(w) + c4);
        const float4 xh = make_float4((x4.x - mu) * rs, (x4.y - mu) * rs, (x4.z - mu) * rs, (x4.w - mu) * rs);
        float4 o;
        o.x = rs * (d4.x * w4.x - s1 - xh.x * s2);
        o.y = rs * (d4.y * w4.y - s1 - xh.y * s2);
        o.z = rs * (d4.z * w4.z - s1 - xh.z * s2);
        o.w = rs * (d4.w * w4.w - s1 - xh.w * s2);
        if (dres) {
          const float4 rr = reinterpret_cast<const float4*>(dres + (long long)r * D)[c4];
          o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
        }
        if (dw) {
          float4 a = s_w[c4], bq = s_b[c4];
          a.x += d4.x * xh.x; a.y += d4.y * xh.y; a.z += d4.z * xh.z; a.w += d4.w * xh.w;
          bq.x += d4.x; bq.y += d4.y; bq.z += d4.z; bq.w += d4.w;
          s_w[c4] = a; s_b[c4] = bq;
        }
        if (dx_colsum) {
          float4 a = s_o[c4];
          a.x += o.x; a.y += o.y; a.z += o.z; a.w += o.w;
          s_o[c4] = a;
        }
        if (dx) reinterpret_cast<float4*>(dx + (long long)r * D)[c4] = o;
        if (dx16) {
          uint2 u;
          u.x = pack_bf16(o.x, o.y);
          u.y = pack_bf16(o.z, o.w);
          reinterpret_cast<uint2*>(dx16 + (long long)r * D)[c4] = u;
        }
      }
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float sw = 0.f, sb = 0.f, so = 0.f;
    for (int k = 0; k < warps_per_block; ++k) {
      if (dw) { sw += red[(size_t)(k * K) * D + c]; sb += red[(size_t)(k * K + 1) * D + c]; }
      if (dx_colsum) so += red[(size_t)(k * K + (dw ? 2 : 0)) * D + c];
    }
    if (dw) { atomicAdd(dw + c, sw); atomicAdd(db + c, sb); }
    if (dx_colsum) atomicAdd(dx_colsum + c, so);
  }
}

// ------------------------------------------------------------------------------------------------ column reductions
// generic: out[c] += sum_r f(x[r,c]); MODE 0: x  (fp32/bf16) ; MODE 1: x and x^2 (out[D + c])
// block (32, 8): threadIdx.x -> 4 consecutive columns, threadIdx.y -> row lane; blockIdx.y -> row chunk
template <typename T, bool SQ>
__global__ void __launch_bounds__(256)
colreduce_kernel(const T* __restrict__ x, long long ldx, float* __restrict__ out, int rows, int cols,
                 int rows_per_block) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  __shared__ float4 sh[8][32];
  __shared__ float4 sh2[8][32];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  float4 a = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
  if (c < cols) {
#pragma unroll 4
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {   // (independent loads: four rows in flight per thread)
      float4 v;
      if constexpr (sizeof(T) == 4) {
        v = *reinterpret_cast<const float4*>(x + (long long)r * ldx + c);
      } else {
        const uint2 u = *reinterpret_cast<const uint2*>(x + (long long)r * ldx + c);
        const float2 lo = unpack_bf16(u.x), hi = unpack_bf16(u.y);
        v = make_float4(lo.x, lo.y, hi.x, hi.y);
      }
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      if (SQ) { q.x += v.x * v.x; q.y += v.y * v.y; q.z += v.z * v.z; q.w += v.w * v.w; }
    }
  }
  sh[threadIdx.y][threadIdx.x] = a;
  if (SQ) sh2[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 t = sh[k][threadIdx.x];
      a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
      if (SQ) {
        const float4 u = sh2[k][threadIdx.x];
        q.x += u.x; q.y += u.y; q.z += u.z; q.w += u.w;
      }
    }
    // one 16-byte reduction per accumulator (a few hundred blocks meet on cols / 4 addresses) when `out` allows it
    // (gradient views of a flat arena are only 4-byte aligned)
    const bool al16 = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    if (al16) atomicAdd(reinterpret_cast<float4*>(out + c), a);
    else { atomicAdd(out + c, a.x); atomicAdd(out + c + 1, a.y); atomicAdd(out + c + 2, a.z); atomicAdd(out + c + 3, a.w); }
    if (SQ) {
      float* o2 = out + cols;
      if (al16) atomicAdd(reinterpret_cast<float4*>(o2 + c), q);
      else { atomicAdd(o2 + c, q.x); atomicAdd(o2 + c + 1, q.y); atomicAdd(o2 + c + 2, q.z); atomicAdd(o2 + c + 3, q.w); }
    }
  }
}

__global__ void zero_kernel(float* __restrict__ p, long long n) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    p[i] = 0.f;
}

__global__ void bn_finalize_kernel(const float* __restrict__ sums, float* __restrict__ mean, float* __restrict__ rstd,
                                   float* __restrict__ rmean, float* __restrict__ rvar, float momentum, float eps,
                                   int M, int D, long long* __restrict__ num_batches_tracked) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;   // nn.BatchNorm2d's step counter
  if (c >= D) return;
  if (sums != nullptr) {
    const float mu = sums[c] / (float)M;
    float var = sums[D + c] / (float)M - mu * mu;
    var = fmaxf(var, 0.f);
    mean[c] = mu;
    rstd[c] = rsqrtf(var + eps);
    if (rmean) {
      rmean[c] = (1.f - momentum) * rmean[c] + momentum * mu;
      const float unb = M > 1 ? var * ((float)M / (float)(M - 1)) : var;
      rvar[c] = (1.f - momentum) * rvar[c] + momentum * unb;
    }
  } else {
    mean[c] = rmean[c];
    rstd[c] = rsqrtf(rvar[c] + eps);
  }
}

__global__ void __launch_bounds__(256)
bn_relu_pos_fwd_kernel(const float* __restrict__ y, const float* __restrict__ mean, const float* __restrict__ rstd,
                       const float* __restrict__ w, const float* __restrict__ b, const float* __restrict__ pos,
                       float* __restrict__ z32, __nv_bfloat16* __restrict__ z16, __nv_bfloat16* __restrict__ z16_alt,
                       long long M, int D, int HW, int fp16) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  const int D4 = D >> 2;
  const long long n4 = M * D4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / D4;
    const int c4 = (int)(i - r * D4);
    const float4 v = reinterpret_cast<const float4*>(y)[i];
    const float4 mu = __ldg(reinterpret_cast<const float4*>(mean) + c4);
    const float4 rs = __ldg(reinterpret_cast<const float4*>(rstd) + c4);
    const float4 g = __ldg(reinterpret_cast<const float4*>(w) + c4);
    const float4 be = __ldg(reinterpret_cast<const float4*>(b) + c4);
    const float4 pe = __ldg(reinterpret_cast<const float4*>(pos + (r % HW) * D) + c4);
    float4 o;
    o.x = fmaxf((v.x - mu.x) * rs.x * g.x + be.x, 0.f) + pe.x;
    o.y = fmaxf((v.y - mu.y) * rs.y * g.y + be.y, 0.f) + pe.y;
    o.z = fmaxf((v.z - mu.z) * rs.z * g.z + be.z, 0.f) + pe.z;
    o.w = fmaxf((v.w - mu.w) * rs.w * g.w + be.w, 0.f) + pe.w;
    if (z32) reinterpret_cast<float4*>(z32)[i] = o;
    if (z16) {
      uint2 u;
      u.x = pack16(o.x, o.y, fp16);
      u.y = pack16(o.z, o.w, fp16);
      reinterpret_cast<uint2*>(z16)[i] = u;
    }
    if (z16_alt) {
      uint2 u;
      u.x = pack16(o.x, o.y, !fp16);
      u.y = pack16(o.z, o.w, !fp16);
      reinterpret_cast<uint2*>(z16_alt)[i] = u;
    }
  }
}

// Same column sums, with the pos_embed gradient in the same pass: dpos[p, :] += sum over images of dz[(b, p), :]
// (pos_embed is added after the ReLU, so dz reaches it unmasked). A block owns 8 positions x 128 columns for EVERY image
// -- thread (x, y): 4 columns of position p0 + y -- so the per-position sums need no atomics and dz is read once (the
// separate batch-sum pass read it a second time: 25 MB and a launch per projector).
__global__ void __launch_bounds__(256)
bn_bwd_reduce_pos_kernel(const float* __restrict__ dz, const float* __restrict__ y, const float* __restrict__ mean,
                         const float* __restrict__ rstd, const float* __restrict__ w, const float* __restrict__ b,
                         float* __restrict__ sums2, float* __restrict__ dpos, int B, int HW, int D) {
  pdl_trigger();
  pdl_wait();
  __shared__ float4 sh[8][32];
  __shared__ float4 sh2[8][32];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int p = blockIdx.y * 8 + threadIdx.y;
  // the images are split over gridDim.z blocks (more loads in flight than one block per position group gives); the
  // per-position sums of the z-slices meet in dpos through atomics (dpos is an accumulator, zeroed by the caller)
  const int per_z = (B + gridDim.z - 1) / gridDim.z;
  const int b0 = blockIdx.z * per_z, b1 = min(B, b0 + per_z);
  float4 a = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0), dp = make_float4(0, 0, 0, 0);
  if (c < D && p < HW) {
    const float4 mu = *reinterpret_cast<const float4*>(mean + c);
    const float4 rs = *reinterpret_cast<const float4*>(rstd + c);
    const float4 g = *reinterpret_cast<const float4*>(w + c);
    const float4 be = *reinterpret_cast<const float4*>(b + c);
#pragma unroll 4
    for (int i = b0; i < b1; ++i) {
      const long long off = ((long long)i * HW + p) * D + c;
      const float4 d4 = *reinterpret_cast<const float4*>(dz + off);
      const float4 y4 = *reinterpret_cast<const float4*>(y + off);
      const float4 yh = make_float4((y4.x - mu.x) * rs.x, (y4.y - mu.y) * rs.y, (y4.z - mu.z) * rs.z, (y4.w - mu.w) * rs.w);
      float4 dr;
      dr.x = (yh.x * g.x + be.x > 0.f) ? d4.x : 0.f;
      dr.y = (yh.y * g.y + be.y > 0.f) ? d4.y : 0.f;
      dr.z = (yh.z * g.z + be.z > 0.f) ? d4.z : 0.f;
      dr.w = (yh.w * g.w + be.w > 0.f) ? d4.w : 0.f;
      a.x += dr.x; a.y += dr.y; a.z += dr.z; a.w += dr.w;
      q.x += dr.x * yh.x; q.y += dr.y * yh.y; q.z += dr.z * yh.z; q.w += dr.w * yh.w;
      dp.x += d4.x; dp.y += d4.y; dp.z += d4.z; dp.w += d4.w;
    }
    float* o = dpos + (long long)p * D + c;
    if (gridDim.z == 1) {
      float4 old = *reinterpret_cast<float4*>(o);
      old.x += dp.x; old.y += dp.y; old.z += dp.z; old.w += dp.w;
      *reinterpret_cast<float4*>(o) = old;
    } else {
      atomicAdd(reinterpret_cast<float4*>(o), dp);
    }
  }
  sh[threadIdx.y][threadIdx.x] = a;
  sh2[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && c < D) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 t = sh[k][threadIdx.x], u = sh2[k][threadIdx.x];
      a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
      q.x += u.x; q.y += u.y; q.z += u.z; q.w += u.w;
    }
    atomicAdd(reinterpret_cast<float4*>(sums2 + c), a);
    atomicAdd(reinterpret_cast<float4*>(sums2 + D + c), q);
  }
}

// sums2[0:D] += sum_r dr ; sums2[D:2D] += sum_r dr * yhat ; dr = dz * (yhat*w+b > 0)
__global__ void __launch_bounds__(256)
bn_bwd_reduce_kernel(const float* __restrict__ dz, const float* __restrict__ y, const float* __restrict__ mean,
                     const float* __restrict__ rstd, const float* __restrict__ w, const float* __restrict__ b,
                     float* __restrict__ sums2, int M, int D, int rows_per_block) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  __shared__ float4 sh[8][32];
  __shared__ float4 sh2[8][32];
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float4 a = make_float4(0, 0, 0, 0), q = make_float4(0, 0, 0, 0);
  if (c < D) {
    const float4 mu = *reinterpret_cast<const float4*>(mean + c);
    const float4 rs = *reinterpret_cast<const float4*>(rstd + c);
    const float4 g = *reinterpret_cast<const float4*>(w + c);
    const float4 be = *reinterpret_cast<const float4*>(b + c);
#pragma unroll 4
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      const float4 d4 = *reinterpret_cast<const float4*>(dz + (long long)r * D + c);
      const float4 y4 = *reinterpret_cast<const float4*>(y + (long long)r * D + c);
      float4 yh = make_float4((y4.x - mu.x) * rs.x, (y4.y - mu.y) * rs.y, (y4.z - mu.z) * rs.z, (y4.w - mu.w) * rs.w);
      float4 dr;
      dr.x = (yh.x * g.x + be.x > 0.f) ? d4.x : 0.f;
      dr.y = (yh.y * g.y + be.y > 0.f) ? d4.y : 0.f;
      dr.z = (yh.z * g.z + be.z > 0.f) ? d4.z : 0.f;
      dr.w = (yh.w * g.w + be.w > 0.f) ? d4.w : 0.f;
      a.x += dr.x; a.y += dr.y; a.z += dr.z; a.w += dr.w;
      q.x += dr.x * yh.x; q.y += dr.y * yh.y; q.z += dr.z * yh.z; q.w += dr.w * yh.w;
    }
  }
  sh[threadIdx.y][threadIdx.x] = a;
  sh2[threadIdx.y][threadIdx.x] = q;
  __syncthreads();
  if (threadIdx.y == 0 && c < D) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      const float4 t = sh[k][threadIdx.x], u = sh2[k][threadIdx.x];
      a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
      q.x += u.x; q.y += u.y; q.z += u.z; q.w += u.w;
    }
    atomicAdd(reinterpret_cast<float4*>(sums2 + c), a);
    atomicAdd(reinterpret_cast<float4*>(sums2 + D + c), q);
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float* __restrict__ dz, const float* __restrict__ y, const float* __restrict__ mean,
                    const float* __restrict__ rstd, const float* __restrict__ w, const float* __restrict__ b,
                    const float* __restrict__ sums2, __nv_bfloat16* __restrict__ dy16, int batch_stats, long long M,
                    int D, float* __restrict__ acc_b, float* __restrict__ acc_w, int rows_per_block) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  // optional: the BN affine gradients are the two column sums this pass already reads (d beta = sums2[0:D],
  // d gamma = sums2[D:2D]): block (0, 0) adds them to the gradient buffers (was two axpy launches)
  if (acc_b != nullptr && blockIdx.x == 0 && blockIdx.y == 0)
    for (int c = threadIdx.y * 32 + threadIdx.x; c < D; c += 256) { acc_b[c] += sums2[c]; acc_w[c] += sums2[D + c]; }
  // thread (x, y): 4 columns, every 8th row of this block's row range -- the per-column constants are loaded once and
  // there is no per-element index division
  const int c = (blockIdx.x * 32 + threadIdx.x) * 4;
  if (c >= D) return;
  const float invM = 1.0f / (float)M;
  const float4 mu = *reinterpret_cast<const float4*>(mean + c);
  const float4 rs = *reinterpret_cast<const float4*>(rstd + c);
  const float4 g = *reinterpret_cast<const float4*>(w + c);
  const float4 be = *reinterpret_cast<const float4*>(b + c);
  float4 s1 = make_float4(0, 0, 0, 0), s2 = make_float4(0, 0, 0, 0);
  if (batch_stats) {
    s1 = *reinterpret_cast<const float4*>(sums2 + c);
    s2 = *reinterpret_cast<const float4*>(sums2 + D + c);
  }
  const float gg[4] = {g.x, g.y, g.z, g.w}, bb[4] = {be.x, be.y, be.z, be.w}, rr[4] = {rs.x, rs.y, rs.z, rs.w};
  const float mm[4] = {mu.x, mu.y, mu.z, mu.w};
  const float a1[4] = {s1.x * invM, s1.y * invM, s1.z * invM, s1.w * invM};
  const float a2[4] = {s2.x * invM, s2.y * invM, s2.z * invM, s2.w * invM};
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  const long long r1 = (r0 + rows_per_block < M) ? r0 + rows_per_block : M;
#pragma unroll 4
  for (long long r = r0 + threadIdx.y; r < r1; r += 8) {
    const float4 d4 = *reinterpret_cast<const float4*>(dz + r * D + c);
    const float4 y4 = *reinterpret_cast<const float4*>(y + r * D + c);
    const float dd[4] = {d4.x, d4.y, d4.z, d4.w}, yy[4] = {y4.x, y4.y, y4.z, y4.w};
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float yh = (yy[e] - mm[e]) * rr[e];
      const float dr = (yh * gg[e] + bb[e] > 0.f) ? dd[e] : 0.f;
      o[e] = gg[e] * rr[e] * (dr - a1[e] - yh * a2[e]);
    }
    uint2 u;
    u.x = pack_bf16(o[0], o[1]);
    u.y = pack_bf16(o[2], o[3]);
    *reinterpret_cast<uint2*>(dy16 + r * D + c) = u;
  }
}

__global__ void batch_sum_kernel(const float* __restrict__ x, float* __restrict__ out, int B, long long n,
                                 int accumulate) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  const long long n4 = n >> 2;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 a = make_float4(0, 0, 0, 0);
    for (int b = 0; b < B; ++b) {
      const float4 v = reinterpret_cast<const float4*>(x + (long long)b * n)[i];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    float4* o = reinterpret_cast<float4*>(out) + i;
    if (accumulate) { const float4 p = *o; a.x += p.x; a.y += p.y; a.z += p.z; a.w += p.w; }
    *o = a;
  }
}

__global__ void swiglu_kernel(const __nv_bfloat16* __restrict__ x12, __nv_bfloat16* __restrict__ out, long long rows,
                              int H) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  const int H8 = H >> 3;
  const long long n = rows * H8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / H8;
    const int c8 = (int)(i - r * H8);
    const uint4 a = *reinterpret_cast<const uint4*>(x12 + r * 2 * H + c8 * 8);
    const uint4 g = *reinterpret_cast<const uint4*>(x12 + r * 2 * H + H + c8 * 8);
    const uint32_t au[4] = {a.x, a.y, a.z, a.w}, gu[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x1 = unpack_bf16(au[e]), x2 = unpack_bf16(gu[e]);
      const float s0 = x1.x / (1.f + __expf(-x1.x)) * x2.x;
      const float s1 = x1.y / (1.f + __expf(-x1.y)) * x2.y;
      o[e] = pack_bf16(s0, s1);
    }
    *reinterpret_cast<uint4*>(out + r * H + c8 * 8) = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Block (32, 8): lane x owns 4 consecutive elements, slice y sums the images b = y, y + 8, ... with up to eight 8-byte loads
// in flight, the eight partial sums meet in shared memory. (One thread per 4 elements summing all B images alone -- the
// earlier form -- has n / 4 = 24 576 threads at the cfg2 shape: too few bytes in flight, 0.75 TB/s under ncu.)
__global__ void __launch_bounds__(256)
batch_sum_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ o32, __nv_bfloat16* __restrict__ o16, int B,
                      long long n) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  __shared__ float4 red[8][32];
  const long long n4 = n >> 2;
  const long long i = (long long)blockIdx.x * 32 + threadIdx.x;
  float4 a = make_float4(0, 0, 0, 0);
  if (i < n4) {
#pragma unroll 8
    for (int b = threadIdx.y; b < B; b += 8) {
      const uint2 u = reinterpret_cast<const uint2*>(x + (long long)b * n)[i];
      const float2 lo = unpack_bf16(u.x), hi = unpack_bf16(u.y);
      a.x += lo.x; a.y += lo.y; a.z += hi.x; a.w += hi.y;
    }
  }
  red[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y != 0 || i >= n4) return;
#pragma unroll
  for (int k = 1; k < 8; ++k) {
    const float4 r = red[k][threadIdx.x];
    a.x += r.x; a.y += r.y; a.z += r.z; a.w += r.w;
  }
  if (o32) reinterpret_cast<float4*>(o32)[i] = a;
  if (o16) {
    uint2 u;
    u.x = pack_bf16(a.x, a.y);
    u.y = pack_bf16(a.z, a.w);
    reinterpret_cast<uint2*>(o16)[i] = u;
  }
}

__global__ void axpy_kernel(const float* __restrict__ x, float* __restrict__ y, float a, long long n) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    y[i] += a * x[i];
}

__global__ void swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ x12, const __nv_bfloat16* __restrict__ dout,
                                  __nv_bfloat16* __restrict__ dx12, long long rows, int H) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  const int H8 = H >> 3;
  const long long n = rows * H8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / H8;
    const int c8 = (int)(i - r * H8);
    const uint4 a = *reinterpret_cast<const uint4*>(x12 + r * 2 * H + c8 * 8);
    const uint4 g = *reinterpret_cast<const uint4*>(x12 + r * 2 * H + H + c8 * 8);
    const uint4 d = *reinterpret_cast<const uint4*>(dout + r * H + c8 * 8);
    const uint32_t au[4] = {a.x, a.y, a.z, a.w}, gu[4] = {g.x, g.y, g.z, g.w}, du[4] = {d.x, d.y, d.z, d.w};
    uint32_t o1[4], o2[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x1 = unpack_bf16(au[e]), x2 = unpack_bf16(gu[e]), dd = unpack_bf16(du[e]);
      const float s0 = 1.f / (1.f + __expf(-x1.x)), s1 = 1.f / (1.f + __expf(-x1.y));
      const float silu0 = x1.x * s0, silu1 = x1.y * s1;
      const float ds0 = s0 * (1.f + x1.x * (1.f - s0)), ds1 = s1 * (1.f + x1.y * (1.f - s1));
      o1[e] = pack_bf16(dd.x * x2.x * ds0, dd.y * x2.y * ds1);
      o2[e] = pack_bf16(dd.x * silu0, dd.y * silu1);
    }
    *reinterpret_cast<uint4*>(dx12 + r * 2 * H + c8 * 8) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
    *reinterpret_cast<uint4*>(dx12 + r * 2 * H + H + c8 * 8) = make_uint4(o2[0], o2[1], o2[2], o2[3]);
  }
}

template <int NV>
static int launch_ln_fwd(const float* x, const float* w, const float* b, float eps, float* y32, void* y16, void* y16_alt,
                         float* mean, float* rstd, int rows, int D, int in_period, int in_pad, int fp16, cudaStream_t st) {
  constexpr int bps = 12;    // blocks per SM cap (12: best of a 4..32 scan at the teacher shape)
  constexpr int wpb = 8;     // warps per block
  const int grid = grid_for(rows, wpb, bps);
  B200_CUDA_OK(launch_pdl(layernorm_fwd_kernel<NV>, dim3(grid), dim3(32 * wpb), 0, st, x, w, b, eps, y32,
                          static_cast<__nv_bfloat16*>(y16), static_cast<__nv_bfloat16*>(y16_alt), mean, rstd, rows, D,
                          in_period, in_pad, fp16));
  B200_LAUNCH_OK();
  return 0;
}

template <int NV>
static int launch_ln_bwd(const float* dy, const float* x, const float* w, const float* mean, const float* rstd,
                         const float* dres, float* dx, void* dx16, float* dw, float* db, float* dx_colsum, int rows,
                         int D, cudaStream_t st) {
  const bool wg = dw != nullptr || dx_colsum != nullptr;
  if (wg && NV >= 6) {   // wide rows: shared-memory column partials, two passes over the row
    auto kern = layernorm_bwd_wide_kernel<NV>;
    const size_t smem = size_t(8) * ((dw ? 2 : 0) + (dx_colsum ? 1 : 0)) * D * sizeof(float);
    static bool set = false;
    if (!set) { B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); set = true; }
    const int bps = smem > 110 * 1024 ? 1 : 2;
    const int grid = grid_for(rows, 32, bps);
    B200_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(256), smem, st, dy, x, w, mean, rstd, dres, dx,
                            static_cast<__nv_bfloat16*>(dx16), dw, db, dx_colsum, rows, D));
    B200_LAUNCH_OK();
    return 0;
  }
  auto kern = wg ? layernorm_bwd_kernel<NV, true> : layernorm_bwd_kernel<NV, false>;
  size_t smem = dw ? size_t(8) * 2 * D * sizeof(float) : (dx_colsum ? size_t(8) * D * sizeof(float) : 0);
  if (smem > 48 * 1024) {
    static bool set = false;
    if (!set) { B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); set = true; }
  }
  constexpr int rpb = 8;    // rows per block (plain form)
  constexpr int bps = 8;    // blocks per SM cap
  const int grid = wg ? grid_for(rows, 32, 2) : grid_for(rows, rpb, bps);
  B200_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(256), smem, st, dy, x, w, mean, rstd, dres, dx,
                          static_cast<__nv_bfloat16*>(dx16), dw, db, dx_colsum, rows, D));
  B200_LAUNCH_OK();
  return 0;
}

template <typename T, bool SQ>
static int launch_colreduce(const T* x, long long ldx, float* out, int rows, int cols, cudaStream_t st) {
  const int gx = (int)cdiv(cols, 128);
  int gy = (int)cdiv((long long)sm_count() * 4, gx);
  int rpb = (int)cdiv(rows, gy);
  if (rpb < 64) rpb = 64;
  gy = (int)cdiv(rows, rpb);
  B200_CUDA_OK(launch_pdl(colreduce_kernel<T, SQ>, dim3(dim3(gx, gy)), dim3(dim3(32, 8)), 0, st, x, ldx, out, rows, cols, rpb));
  B200_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------ bilinear resize (tokens)
// F.interpolate(mode='bilinear', align_corners=False) of models/model_zoo.py:121-126, applied to TOKEN-major maps
// [B, h*w, D] -> [B, H*W, D]. ModelWrapper resizes the student map and ScaleKD's proj_student starts with a 1x1 conv
// (losses/scalekd.py:199): both are linear and act on different axes, so conv1x1(resize(x)) == resize(conv1x1(x)) (the
// four tap weights sum to one, so the bias commutes too). The projector runs the conv GEMM on the raw map and resizes
// its D-channel token output with this kernel.
// PyTorch's source index: src = max(0, (dst + 0.5) * in / out - 0.5); i0 = floor(src); i1 = i0 + (i0 < in - 1); l = src - i0.
struct BilinearTap { int i0, i1; float l0, l1; };
__device__ __forceinline__ BilinearTap bilinear_tap(int dst, int in, int out) {
  const float scale = (float)in / (float)out;
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  BilinearTap t;
  t.i0 = (int)src;
  if (t.i0 > in - 1) t.i0 = in - 1;
  t.i1 = t.i0 + (t.i0 < in - 1 ? 1 : 0);
  t.l1 = src - (float)t.i0;
  t.l0 = 1.0f - t.l1;
  return t;
}

__global__ void __launch_bounds__(256) bilinear_tokens_fwd_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                                 int B, int h, int w, int H, int W, int D) {
  pdl_trigger();
  pdl_wait();
  const int D4 = D >> 2;
  const long long n = (long long)B * H * W * D4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % D4);
    long long r = i / D4;
    const int ox = (int)(r % W); r /= W;
    const int oy = (int)(r % H);
    const int b = (int)(r / H);
    const BilinearTap ty = bilinear_tap(oy, h, H), tx = bilinear_tap(ox, w, W);
    const float4* s = reinterpret_cast<const float4*>(src + (long long)b * h * w * D) + c4;
    const float4 a = __ldg(s + (long long)(ty.i0 * w + tx.i0) * D4), bq = __ldg(s + (long long)(ty.i0 * w + tx.i1) * D4);
    const float4 c = __ldg(s + (long long)(ty.i1 * w + tx.i0) * D4), d = __ldg(s + (long long)(ty.i1 * w + tx.i1) * D4);
    float4 o;
    o.x = ty.l0 * (tx.l0 * a.x + tx.l1 * bq.x) + ty.l1 * (tx.l0 * c.x + tx.l1 * d.x);
    o.y = ty.l0 * (tx.l0 * a.y + tx.l1 * bq.y) + ty.l1 * (tx.l0 * c.y + tx.l1 * d.y);
    o.z = ty.l0 * (tx.l0 * a.z + tx.l1 * bq.z) + ty.l1 * (tx.l0 * c.z + tx.l1 * d.z);
    o.w = ty.l0 * (tx.l0 * a.w + tx.l1 * bq.w) + ty.l1 * (tx.l0 * c.w + tx.l1 * d.w);
    reinterpret_cast<float4*>(dst)[i] = o;
  }
}

// adjoint: d_src[b, (i,j), :] = sum over output pixels of weight((oy,ox) -> (i,j)) * d_dst[b, (oy,ox), :]. One block per
// (image, source pixel): the row / column weights of every output coordinate onto (i, j) go to shared memory first
// (each source pixel is touched by a short contiguous range of outputs), then each thread gathers its 4 channels.
// bf16 in (the BN backward's output), fp32 accumulation, bf16 out (operand of the conv wgrad / dgrad GEMMs).
constexpr int BILINEAR_MAX_OUT = 256;
__global__ void __launch_bounds__(128) bilinear_tokens_bwd_kernel(const __nv_bfloat16* __restrict__ d_dst,
                                                                 __nv_bfloat16* __restrict__ d_src, int h, int w, int H,
                                                                 int W, int D) {
  pdl_trigger();
  pdl_wait();
  __shared__ float wy[BILINEAR_MAX_OUT], wx[BILINEAR_MAX_OUT];
  __shared__ int range[4];
  const int q = blockIdx.x, b = blockIdx.y;
  const int i = q / w, j = q - i * w;
  for (int o = threadIdx.x; o < H; o += blockDim.x) {
    const BilinearTap t = bilinear_tap(o, h, H);
    wy[o] = (t.i0 == i ? t.l0 : 0.f) + (t.i1 == i ? t.l1 : 0.f);
  }
  for (int o = threadIdx.x; o < W; o += blockDim.x) {
    const BilinearTap t = bilinear_tap(o, w, W);
    wx[o] = (t.i0 == j ? t.l0 : 0.f) + (t.i1 == j ? t.l1 : 0.f);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int lo = H, hi = -1;
    for (int o = 0; o < H; ++o) if (wy[o] != 0.f) { lo = o < lo ? o : lo; hi = o; }
    range[0] = lo; range[1] = hi;
    lo = W; hi = -1;
    for (int o = 0; o < W; ++o) if (wx[o] != 0.f) { lo = o < lo ? o : lo; hi = o; }
    range[2] = lo; range[3] = hi;
  }
  __syncthreads();
  const int y0 = range[0], y1 = range[1], x0 = range[2], x1 = range[3];
  const int D4 = D >> 2;
  const __nv_bfloat16* g = d_dst + (long long)b * H * W * D;
  for (int c4 = threadIdx.x; c4 < D4; c4 += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int oy = y0; oy <= y1; ++oy) {
      const float a = wy[oy];
      if (a == 0.f) continue;
      for (int ox = x0; ox <= x1; ++ox) {
        const float wgt = a * wx[ox];
        if (wgt == 0.f) continue;
        const uint2 u = __ldg(reinterpret_cast<const uint2*>(g + (long long)(oy * W + ox) * D) + c4);
        const float2 f0 = unpack_bf16(u.x), f1 = unpack_bf16(u.y);
        acc.x = fmaf(wgt, f0.x, acc.x); acc.y = fmaf(wgt, f0.y, acc.y);
        acc.z = fmaf(wgt, f1.x, acc.z); acc.w = fmaf(wgt, f1.y, acc.w);
      }
    }
    uint2 o;
    o.x = pack_bf16(acc.x, acc.y);
    o.y = pack_bf16(acc.z, acc.w);
    reinterpret_cast<uint2*>(d_src + ((long long)b * h * w + q) * D)[c4] = o;
  }
}


// ------------------------------------------------------------------------------------------------ window token order
// WindowMultiheadPosAttention.separate_tokens (losses/scalekd.py:326-335) cuts the H x W token grid into win_h x win_w
// windows. Rows of a token-major 16-bit matrix [B*H*W, ld] are moved between raster order and WINDOW-major order
// (window index row-major over the windows, then row-major inside the window), so that each window's tokens are a
// contiguous run of rows and the attention kernels can take a window as one sequence. One warp per row, 16-byte copies.
__global__ void __launch_bounds__(256) window_rows16_kernel(const __nv_bfloat16* __restrict__ src,
                                                           __nv_bfloat16* __restrict__ dst, long long rows, int H, int W,
                                                           int win_h, int win_w, int cols, long long ld, int to_raster) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int hh = H / win_h, ww = W / win_w, HW = H * W;
  for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows;
       r += (long long)gridDim.x * (blockDim.x >> 5)) {
    const long long b = r / HW;
    const int t = (int)(r - b * HW);            // window-major index: win * (hh*ww) + ly * ww + lx
    const int win = t / (hh * ww), l = t - win * (hh * ww);
    const int wy = win / win_w, wx = win - wy * win_w, ly = l / ww, lx = l - ly * ww;
    const long long raster = b * HW + (long long)(wy * hh + ly) * W + wx * ww + lx;
    const uint4* s = reinterpret_cast<const uint4*>(src + (to_raster ? r : raster) * ld);
    uint4* d = reinterpret_cast<uint4*>(dst + (to_raster ? raster : r) * ld);
    for (int c = lane; c < cols / 8; c += 32) d[c] = s[c];
  }
}

// ------------------------------------------------------------------------------------------------ parameter prep
// One launch for all the per-step working copies of a projector's fp32 master parameters (casts, the 3-term split of
// the conv weight, transposed bf16 copies for the dgrad GEMMs, bias concatenation, pos_embed to token-major): these were
// 10 (forward) and 7 (backward) launches of a few microseconds each. blockIdx.y selects the job.
__global__ void __launch_bounds__(256) param_prep_kernel(const PrepJobs jobs) {
  pdl_trigger();   // PDL (common.cuh): launched through launch_pdl()
  pdl_wait();
  const PrepJob& jb = jobs.j[blockIdx.y];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  if (jb.type == PREP_CAST16 || jb.type == PREP_COPY32) {
    const long long n = (long long)jb.rows * jb.cols;
    const long long n4 = n >> 2;
    const float4* src4 = reinterpret_cast<const float4*>(jb.src);
    for (long long i = (long long)blockIdx.x * 256 + tid; i < n4; i += (long long)gridDim.x * 256) {
      const float4 v = __ldg(src4 + i);
      if (jb.type == PREP_COPY32) {
        reinterpret_cast<float4*>(jb.dst)[i] = v;
      } else {
        uint2 u;
        u.x = pack16(v.x, v.y, jb.fp16);
        u.y = pack16(v.z, v.w, jb.fp16);
        reinterpret_cast<uint2*>(jb.dst)[i] = u;
      }
    }
    if (blockIdx.x == 0 && tid < (int)(n & 3)) {   // tail (n not a multiple of 4)
      const long long i = (n4 << 2) + tid;
      if (jb.type == PREP_COPY32) static_cast<float*>(jb.dst)[i] = jb.src[i];
      else store16(static_cast<__nv_bfloat16*>(jb.dst) + i, jb.src[i], jb.fp16);
    }
  } else if (jb.type == PREP_SPLIT3_RIGHT) {
    const int K = jb.cols, K4 = K >> 2;
    const long long n4 = (long long)jb.rows * K4;
    __nv_bfloat16* out = static_cast<__nv_bfloat16*>(jb.dst);
    for (long long i = (long long)blockIdx.x * 256 + tid; i < n4; i += (long long)gridDim.x * 256) {
      const long long r = i / K4;
      const int c4 = (int)(i - r * K4);
      const float4 v = __ldg(reinterpret_cast<const float4*>(jb.src) + i);
      const float f[4] = {v.x, v.y, v.z, v.w};
      float hi[4], lo[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        hi[e] = jb.fp16 ? __half2float(__float2half_rn(f[e])) : __bfloat162float(__float2bfloat16(f[e]));
        lo[e] = f[e] - hi[e];
      }
      uint2 uh, ul;
      uh.x = pack16(hi[0], hi[1], jb.fp16); uh.y = pack16(hi[2], hi[3], jb.fp16);
      ul.x = pack16(lo[0], lo[1], jb.fp16); ul.y = pack16(lo[2], lo[3], jb.fp16);
      __nv_bfloat16* o = out + r * 3 * K + c4 * 4;
      *reinterpret_cast<uint2*>(o) = uh;          // [hi | lo | hi]: right operand of the split product
      *reinterpret_cast<uint2*>(o + K) = ul;
      *reinterpret_cast<uint2*>(o + 2 * K) = uh;
    }
  } else {
    // PREP_TRANSPOSE16 / PREP_TRANSPOSE32: out[c * out_ld + r] = in[r * cols + c], 32 x 32 tiles through shared memory
    __shared__ float tile[32][33];
    const int tiles_c = (jb.cols + 31) >> 5, tiles_r = (jb.rows + 31) >> 5;
    for (int t = blockIdx.x; t < tiles_c * tiles_r; t += gridDim.x) {
      const int r0 = (t / tiles_c) << 5, c0 = (t % tiles_c) << 5;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int r = r0 + threadIdx.y + 8 * k, c = c0 + threadIdx.x;
        tile[threadIdx.y + 8 * k][threadIdx.x] = (r < jb.rows && c < jb.cols) ? __ldg(jb.src + (long long)r * jb.cols + c) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int c = c0 + threadIdx.y + 8 * k, r = r0 + threadIdx.x;
        if (c < jb.cols && r < jb.rows) {
          const float v = tile[threadIdx.x][threadIdx.y + 8 * k];
          if (jb.type == PREP_TRANSPOSE32) static_cast<float*>(jb.dst)[(long long)c * jb.out_ld + r] = v;
          else store16(static_cast<__nv_bfloat16*>(jb.dst) + (long long)c * jb.out_ld + r, v, jb.fp16);
        }
      }
      __syncthreads();
    }
  }
}

int launch_param_prep(const PrepJobs& jobs, cudaStream_t st) {
  if (jobs.n <= 0) return 0;
  B200_CHECK_ARG(jobs.n <= PREP_MAX_JOBS, "too many prep jobs");
  for (int i = 0; i < jobs.n; ++i) {
    const PrepJob& j = jobs.j[i];
    B200_CHECK_ARG(j.src && j.dst && j.rows > 0 && j.cols > 0, "bad prep job");
    if (j.type == PREP_CAST16 || j.type == PREP_COPY32 || j.type == PREP_SPLIT3_RIGHT)
      B200_CHECK_ARG((reinterpret_cast<uintptr_t>(j.src) & 15) == 0 && (reinterpret_cast<uintptr_t>(j.dst) & 15) == 0,
                     "prep job alignment");
    if (j.type == PREP_SPLIT3_RIGHT) B200_CHECK_ARG(j.cols % 4 == 0, "split3 needs K % 4 == 0");
  }
  B200_CUDA_OK(launch_pdl(param_prep_kernel, dim3(dim3(192, (unsigned)jobs.n)), dim3(dim3(32, 8)), 0, st, jobs));
  B200_LAUNCH_OK();
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_cast_f32_bf16(const float* x, void* y, long long n, void* stream) {
  B200_CHECK_ARG(x && y && n >= 0, "bad args");
  if (n == 0) return 0;
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 7) == 0, "alignment");
  B200_CUDA_OK(launch_pdl(cast_f32_bf16_kernel, dim3(grid_for(n / 4 + 1, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      x, static_cast<__nv_bfloat16*>(y), static_cast<__nv_bfloat16*>(nullptr), n, 0));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_cast_f32_f16(const float* x, void* y, long long n, void* stream) {
  return b200::cast_f32_f16_dual(x, y, nullptr, n, stream);
}

// fp32 -> fp16 and (optionally) bf16 in one pass
int b200::cast_f32_f16_dual(const float* x, void* y_f16, void* y_bf16, long long n, void* stream) {
  B200_CHECK_ARG(x && y_f16 && n >= 0, "bad args");
  if (n == 0) return 0;
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y_f16) & 7) == 0 &&
                     (reinterpret_cast<uintptr_t>(y_bf16) & 7) == 0, "alignment");
  B200_CUDA_OK(launch_pdl(cast_f32_bf16_kernel, dim3(grid_for(n / 4 + 1, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
      x, static_cast<__nv_bfloat16*>(y_f16), static_cast<__nv_bfloat16*>(y_bf16), n, 1));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_cast_f16_bf16(const void* x, void* y, long long n, void* stream) {
  B200_CHECK_ARG(x && y && n >= 0, "bad args");
  if (n == 0) return 0;
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0, "alignment");
  B200_CUDA_OK(launch_pdl(cast_f16_bf16_kernel, dim3(grid_for(n / 8 + 1, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __half*>(x), static_cast<__nv_bfloat16*>(y), n));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_split3_16(const float* x, void* out, long long rows, int K, int right_operand, int out_is_fp16,
                              void* stream) {
  B200_CHECK_ARG(x && out && rows > 0 && K > 0 && K % 4 == 0, "bad args");
  B200_CUDA_OK(launch_pdl(split3_kernel, dim3(grid_for(rows * K / 4, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      x, static_cast<__nv_bfloat16*>(out), rows, K, right_operand ? 1 : 0, out_is_fp16 ? 1 : 0));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_transpose_f32_bf16(const float* in, void* out, int rows, int cols, const float* row_scale,
                                       void* stream) {
  B200_CHECK_ARG(in && out && rows > 0 && cols > 0, "bad args");
  dim3 grid((unsigned)cdiv(cols, 32), (unsigned)cdiv(rows, 32), 1);
  B200_CUDA_OK(launch_pdl(transpose_kernel<__nv_bfloat16, false>, dim3(grid), dim3(dim3(32, 8)), 0, static_cast<cudaStream_t>(stream), 
      in, static_cast<__nv_bfloat16*>(out), rows, cols, row_scale, 0, 0, nullptr, rows, 0));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_transpose_f32_bf16_ld(const float* in, void* out, int rows, int cols, long long out_ld,
                                          const float* row_scale, void* stream) {
  B200_CHECK_ARG(in && out && rows > 0 && cols > 0 && out_ld >= rows, "bad args");
  dim3 grid((unsigned)cdiv(cols, 32), (unsigned)cdiv(rows, 32), 1);
  B200_CUDA_OK(launch_pdl(transpose_kernel<__nv_bfloat16, false>, dim3(grid), dim3(dim3(32, 8)), 0, static_cast<cudaStream_t>(stream), 
      in, static_cast<__nv_bfloat16*>(out), rows, cols, row_scale, 0, 0, nullptr, out_ld, 0));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_nchw_to_tokens(const float* x, void* tok_bf16, float* tok_f32, int B, int C, int HW,
                                   int tok16_is_fp16, void* stream) {
  B200_CHECK_ARG(x && (tok_bf16 || tok_f32) && B > 0 && C > 0 && HW > 0, "bad args");
  dim3 grid((unsigned)cdiv(HW, 32), (unsigned)cdiv(C, 32), (unsigned)B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (tok_bf16) {
    B200_CUDA_OK(launch_pdl(transpose_kernel<__nv_bfloat16, false>, dim3(grid), dim3(dim3(32, 8)), 0, st, 
        x, static_cast<__nv_bfloat16*>(tok_bf16), C, HW, nullptr, (long long)C * HW, (long long)C * HW, tok_f32, C, tok16_is_fp16));
  } else {
    B200_CUDA_OK(launch_pdl(transpose_kernel<float, false>, dim3(grid), dim3(dim3(32, 8)), 0, st, x, tok_f32, C, HW, nullptr, (long long)C * HW,
                                                                  (long long)C * HW, nullptr, C, 0));
  }
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_tokens_to_nchw(const float* tok, float* x, int B, int C, int HW, int accumulate, void* stream) {
  B200_CHECK_ARG(tok && x && B > 0 && C > 0 && HW > 0, "bad args");
  dim3 grid((unsigned)cdiv(C, 32), (unsigned)cdiv(HW, 32), (unsigned)B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (accumulate) {
    B200_CUDA_OK(launch_pdl(transpose_kernel<float, true>, dim3(grid), dim3(dim3(32, 8)), 0, st, tok, x, HW, C, nullptr, (long long)C * HW,
                                                                 (long long)C * HW, nullptr, HW, 0));
  } else {
    B200_CUDA_OK(launch_pdl(transpose_kernel<float, false>, dim3(grid), dim3(dim3(32, 8)), 0, st, tok, x, HW, C, nullptr, (long long)C * HW,
                                                                  (long long)C * HW, nullptr, HW, 0));
  }
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_bilinear_tokens_fwd(const float* src, float* dst, int B, int h, int w, int H, int W, int D,
                                        void* stream) {
  B200_CHECK_ARG(src && dst && B > 0 && h > 0 && w > 0 && H > 0 && W > 0 && D > 0 && D % 4 == 0, "bad args");
  B200_CUDA_OK(launch_pdl(bilinear_tokens_fwd_kernel, dim3(grid_for((long long)B * H * W * D / 4, 256)), dim3(256), 0,
                          static_cast<cudaStream_t>(stream), src, dst, B, h, w, H, W, D));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_bilinear_tokens_bwd(const void* d_dst_bf16, void* d_src_bf16, int B, int h, int w, int H, int W,
                                        int D, void* stream) {
  B200_CHECK_ARG(d_dst_bf16 && d_src_bf16 && B > 0 && h > 0 && w > 0 && H > 0 && W > 0 && D > 0 && D % 4 == 0, "bad args");
  B200_CHECK_ARG(H <= BILINEAR_MAX_OUT && W <= BILINEAR_MAX_OUT, "output grid larger than 256 per side");
  B200_CUDA_OK(launch_pdl(bilinear_tokens_bwd_kernel, dim3((unsigned)(h * w), (unsigned)B), dim3(128), 0,
                          static_cast<cudaStream_t>(stream), static_cast<const __nv_bfloat16*>(d_dst_bf16),
                          static_cast<__nv_bfloat16*>(d_src_bf16), h, w, H, W, D));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_window_rows16(const void* src, void* dst, long long rows, int H, int W, int win_h, int win_w,
                                  int cols, long long ld, int to_raster, void* stream) {
  B200_CHECK_ARG(src && dst && src != dst && rows > 0 && H > 0 && W > 0 && win_h > 0 && win_w > 0, "bad args");
  B200_CHECK_ARG(H % win_h == 0 && W % win_w == 0 && rows % ((long long)H * W) == 0, "windows must tile the grid");
  B200_CHECK_ARG(cols > 0 && cols % 8 == 0 && ld % 8 == 0 && cols <= ld, "cols and ld must be multiples of 8");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0, "alignment");
  B200_CUDA_OK(launch_pdl(window_rows16_kernel, dim3(grid_for(rows, 8, 8)), dim3(256), 0, static_cast<cudaStream_t>(stream),
                          static_cast<const __nv_bfloat16*>(src), static_cast<__nv_bfloat16*>(dst), rows, H, W, win_h, win_w,
                          cols, ld, to_raster));
  B200_LAUNCH_OK();
  return 0;
}

int b200::tokenize_split3(const float* x, void* xt_bf16, void* xt3_fp16, int B, int C, int HW, void* stream) {
  B200_CHECK_ARG(x && xt_bf16 && xt3_fp16 && B > 0 && C > 0 && HW > 0 && C % 2 == 0, "bad args");
  const bool wide = C % 4 == 0 && (reinterpret_cast<uintptr_t>(xt_bf16) & 7) == 0 && (reinterpret_cast<uintptr_t>(xt3_fp16) & 7) == 0;
  if (wide) {
    dim3 grid((unsigned)cdiv(HW, 32), (unsigned)cdiv(C, 128), (unsigned)B);
    B200_CUDA_OK(launch_pdl(tokenize_split3_wide_kernel, grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream), x,
                            static_cast<__nv_bfloat16*>(xt_bf16), static_cast<__nv_bfloat16*>(xt3_fp16), C, HW));
    B200_LAUNCH_OK();
    return 0;
  }
  dim3 grid((unsigned)cdiv(HW, 32), (unsigned)cdiv(C, 64), (unsigned)B);
  B200_CUDA_OK(launch_pdl(tokenize_split3_kernel, grid, dim3(32, 8), 0, static_cast<cudaStream_t>(stream), x,
                          static_cast<__nv_bfloat16*>(xt_bf16), static_cast<__nv_bfloat16*>(xt3_fp16), C, HW));
  B200_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------------------------------------ teacher-feature cache
// the frozen teacher's output for a sample does not depend on the student, so a deterministic input pipeline
// (no augmentation, or a fixed augmentation per sample id) can keep it in HBM -- 180 GB hold 0.9 M images of vits14 @224
// in bf16 -- and skip models/backbones/dinov2.py:27-46 from the second epoch on. Rows of the pool are whole images
// ([N, D] tokens, cls row included or not as the caller lays them out); slots[b] picks the row of batch item b.
// One block per (image, 8 KB chunk): 16-byte loads, 8- or 16-byte stores.
template <bool POOL_BF16>
__global__ void __launch_bounds__(256)
feature_cache_store_kernel(const float* __restrict__ feat, long long f_bs, long long f_ts, const long long* __restrict__ slots,
                           void* __restrict__ pool, int N, int D) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const long long slot = slots[b];
  if (slot < 0) return;
  const int D4 = D >> 2;
  const long long n4 = (long long)N * D4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const int t = (int)(i / D4), c4 = (int)(i - (long long)t * D4);
    const float4 v = reinterpret_cast<const float4*>(feat + (long long)b * f_bs + (long long)t * f_ts)[c4];
    if (POOL_BF16) {
      reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(pool) + slot * N * D)[i] = make_uint2(pack_bf16(v.x, v.y), pack_bf16(v.z, v.w));
    } else {
      reinterpret_cast<float4*>(static_cast<float*>(pool) + slot * N * D)[i] = v;
    }
  }
}
template <bool POOL_BF16>
__global__ void __launch_bounds__(256)
feature_cache_load_kernel(const void* __restrict__ pool, const long long* __restrict__ slots, float* __restrict__ out, int N,
                          int D) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.y;
  const long long slot = slots[b];
  if (slot < 0) return;
  const long long n4 = (long long)N * D >> 2;
  float4* o = reinterpret_cast<float4*>(out + (long long)b * N * D);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    if (POOL_BF16) {
      const uint2 u = reinterpret_cast<const uint2*>(static_cast<const __nv_bfloat16*>(pool) + slot * N * D)[i];
      const float2 a = unpack_bf16(u.x), c = unpack_bf16(u.y);
      o[i] = make_float4(a.x, a.y, c.x, c.y);
    } else {
      o[i] = reinterpret_cast<const float4*>(static_cast<const float*>(pool) + slot * N * D)[i];
    }
  }
}

extern "C" int b200_feature_cache_store(const float* feat, long long f_bs, long long f_ts, const long long* slots, void* pool,
                                        int pool_is_bf16, int B, int N, int D, void* stream) {
  B200_CHECK_ARG(feat && slots && pool && B > 0 && N > 0 && D > 0 && D % 4 == 0, "bad args");
  B200_CHECK_ARG(f_bs % 4 == 0 && f_ts % 4 == 0 && (reinterpret_cast<uintptr_t>(feat) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(pool) & 15) == 0, "feature rows and the pool must be 16-byte aligned");
  const long long n4 = (long long)N * D / 4;
  dim3 grid((unsigned)(cdiv(n4, 2048) < 1 ? 1 : cdiv(n4, 2048)), (unsigned)B);
  if (pool_is_bf16) B200_CUDA_OK(launch_pdl(feature_cache_store_kernel<true>, grid, dim3(256), 0, static_cast<cudaStream_t>(stream), feat, f_bs, f_ts, slots, pool, N, D));
  else B200_CUDA_OK(launch_pdl(feature_cache_store_kernel<false>, grid, dim3(256), 0, static_cast<cudaStream_t>(stream), feat, f_bs, f_ts, slots, pool, N, D));
  B200_LAUNCH_OK();
  return 0;
}
extern "C" int b200_feature_cache_load(const void* pool, int pool_is_bf16, const long long* slots, float* out, int B, int N, int D,
                                       void* stream) {
  B200_CHECK_ARG(pool && slots && out && B > 0 && N > 0 && D > 0 && ((long long)N * D) % 4 == 0, "bad args");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(pool) & 15) == 0, "alignment");
  const long long n4 = (long long)N * D / 4;
  dim3 grid((unsigned)(cdiv(n4, 2048) < 1 ? 1 : cdiv(n4, 2048)), (unsigned)B);
  if (pool_is_bf16) B200_CUDA_OK(launch_pdl(feature_cache_load_kernel<true>, grid, dim3(256), 0, static_cast<cudaStream_t>(stream), pool, slots, out, N, D));
  else B200_CUDA_OK(launch_pdl(feature_cache_load_kernel<false>, grid, dim3(256), 0, static_cast<cudaStream_t>(stream), pool, slots, out, N, D));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_patch_im2col(const float* img, void* out, int B, int H, int W, int Kp, void* stream) {
  B200_CHECK_ARG(img && out && B > 0, "bad args");
  B200_CHECK_ARG(H % 14 == 0 && W % 14 == 0 && H > 0 && W > 0, "image size must be a multiple of the 14-pixel patch");
  B200_CHECK_ARG(Kp >= 588 && Kp % 8 == 0, "Kp must be >= 588 and a multiple of 8");
  B200_CUDA_OK(launch_pdl(patch_im2col_kernel, dim3(B * (H / 14) * 3), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      img, static_cast<__nv_bfloat16*>(out), H, W, Kp));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_write_cls_rows(float* x, const float* cls, const float* pos, int B, int N, int D, void* stream) {
  B200_CHECK_ARG(x && cls && pos && B > 0 && N > 0 && D > 0, "bad args");
  B200_CUDA_OK(launch_pdl(write_cls_rows_kernel, dim3((unsigned)cdiv((long long)B * D, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      x, cls, pos, B, N, D));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_layernorm_fwd(const float* x, const float* w, const float* b, float eps, float* y_f32,
                                  void* y_bf16, float* mean, float* rstd, int rows, int D, int in_period, int in_pad,
                                  int y16_is_fp16, void* stream) {
  return b200::layernorm_fwd_dual(x, w, b, eps, y_f32, y_bf16, nullptr, mean, rstd, rows, D, in_period, in_pad,
                                  y16_is_fp16, stream);
}

// y16_alt: optional second 16-bit copy of the output in the format y16 does NOT use
int b200::layernorm_fwd_dual(const float* x, const float* w, const float* b, float eps, float* y_f32, void* y16,
                             void* y16_alt, float* mean, float* rstd, int rows, int D, int in_period, int in_pad,
                             int y16_is_fp16, void* stream) {
  B200_CHECK_ARG(x && w && b && rows > 0, "bad args");
  B200_CHECK_ARG(D % 4 == 0 && D > 0 && D <= LN_MAX_V4 * 128, "D must be a multiple of 4 and <= 1536");
  B200_CHECK_ARG(y16_alt == nullptr || y16 != nullptr, "an alternate-format copy needs the primary 16-bit output");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nv = (int)cdiv(D, 128);
  if (nv <= 3) return launch_ln_fwd<3>(x, w, b, eps, y_f32, y16, y16_alt, mean, rstd, rows, D, in_period, in_pad, y16_is_fp16, st);
  if (nv <= 6) return launch_ln_fwd<6>(x, w, b, eps, y_f32, y16, y16_alt, mean, rstd, rows, D, in_period, in_pad, y16_is_fp16, st);
  if (nv <= 8) return launch_ln_fwd<8>(x, w, b, eps, y_f32, y16, y16_alt, mean, rstd, rows, D, in_period, in_pad, y16_is_fp16, st);
  return launch_ln_fwd<12>(x, w, b, eps, y_f32, y16, y16_alt, mean, rstd, rows, D, in_period, in_pad, y16_is_fp16, st);
}

extern "C" int b200_layernorm_bwd(const float* dy, const float* x, const float* w, const float* mean,
                                  const float* rstd, const float* dres, float* dx, void* dx_bf16, float* dw, float* db,
                                  int rows, int D, void* stream) {
  return b200_layernorm_bwd_colsum(dy, x, w, mean, rstd, dres, dx, dx_bf16, dw, db, nullptr, rows, D, stream);
}

extern "C" int b200_layernorm_bwd_colsum(const float* dy, const float* x, const float* w, const float* mean,
                                         const float* rstd, const float* dres, float* dx, void* dx_bf16, float* dw,
                                         float* db, float* dx_colsum, int rows, int D, void* stream) {
  B200_CHECK_ARG(dy && x && w && mean && rstd && rows > 0, "bad args");
  B200_CHECK_ARG((dw == nullptr) == (db == nullptr), "dw and db go together");
  B200_CHECK_ARG(D % 4 == 0 && D > 0 && D <= LN_MAX_V4 * 128, "D must be a multiple of 4 and <= 1536");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nv = (int)cdiv(D, 128);
  if (nv <= 3) return launch_ln_bwd<3>(dy, x, w, mean, rstd, dres, dx, dx_bf16, dw, db, dx_colsum, rows, D, st);
  if (nv <= 6) return launch_ln_bwd<6>(dy, x, w, mean, rstd, dres, dx, dx_bf16, dw, db, dx_colsum, rows, D, st);
  if (nv <= 8) return launch_ln_bwd<8>(dy, x, w, mean, rstd, dres, dx, dx_bf16, dw, db, dx_colsum, rows, D, st);
  return launch_ln_bwd<12>(dy, x, w, mean, rstd, dres, dx, dx_bf16, dw, db, dx_colsum, rows, D, st);
}

extern "C" int b200_bn_stats(const float* y, float* sums, int M, int D, void* stream) {
  B200_CHECK_ARG(y && sums && M > 0 && D > 0 && D % 4 == 0, "bad args");
  return launch_colreduce<float, true>(y, D, sums, M, D, static_cast<cudaStream_t>(stream));
}

extern "C" int b200_bn_finalize(const float* sums, float* mean, float* rstd, float* running_mean, float* running_var,
                                float momentum, float eps, int M, int D, void* stream) {
  return b200::bn_finalize_counted(sums, mean, rstd, running_mean, running_var, momentum, eps, M, D, nullptr, stream);
}

int b200::bn_finalize_counted(const float* sums, float* mean, float* rstd, float* running_mean, float* running_var,
                              float momentum, float eps, int M, int D, long long* num_batches_tracked, void* stream) {
  B200_CHECK_ARG(mean && rstd && D > 0, "bad args");
  B200_CHECK_ARG(sums || (running_mean && running_var), "need batch sums or running statistics");
  B200_CUDA_OK(launch_pdl(bn_finalize_kernel, dim3((unsigned)cdiv(D, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
      sums, mean, rstd, running_mean, running_var, momentum, eps, M, D, num_batches_tracked));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_bn_relu_pos_fwd(const float* y, const float* mean, const float* rstd, const float* w,
                                    const float* b, const float* pos, float* z_f32, void* z_bf16, int M, int D, int HW,
                                    int z16_is_fp16, void* stream) {
  return b200::bn_relu_pos_fwd_dual(y, mean, rstd, w, b, pos, z_f32, z_bf16, nullptr, M, D, HW, z16_is_fp16, stream);
}

int b200::bn_relu_pos_fwd_dual(const float* y, const float* mean, const float* rstd, const float* w, const float* b,
                               const float* pos, float* z_f32, void* z16, void* z16_alt, int M, int D, int HW,
                               int z16_is_fp16, void* stream) {
  B200_CHECK_ARG(y && mean && rstd && w && b && pos && (z_f32 || z16) && M > 0 && D % 4 == 0 && HW > 0, "bad args");
  B200_CUDA_OK(launch_pdl(bn_relu_pos_fwd_kernel, dim3(grid_for((long long)M * D / 4, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream),
      y, mean, rstd, w, b, pos, z_f32, static_cast<__nv_bfloat16*>(z16), static_cast<__nv_bfloat16*>(z16_alt), (long long)M, D, HW,
      z16_is_fp16));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_bn_relu_pos_bwd_reduce(const float* dz, const float* y, const float* mean, const float* rstd,
                                           const float* w, const float* b, float* sums2, float* dpos, int M, int D,
                                           int HW, void* stream) {
  B200_CHECK_ARG(dz && y && mean && rstd && w && b && sums2 && M > 0 && D % 4 == 0, "bad args");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int gx = (int)cdiv(D, 128);
  if (dpos != nullptr && HW > 0 && M % HW == 0 && M / HW >= 8) {   // enough images per position to keep a thread busy
    const int nb = M / HW;
    const long long blocks_xy = (long long)gx * cdiv(HW, 8);
    int gz = (int)cdiv((long long)sm_count() * 4, blocks_xy);    // ~4 blocks per SM, at least 8 images per block
    if (gz > nb / 8) gz = nb / 8;
    if (gz < 1) gz = 1;
    B200_CUDA_OK(launch_pdl(bn_bwd_reduce_pos_kernel, dim3(dim3(gx, (unsigned)cdiv(HW, 8), (unsigned)gz)), dim3(dim3(32, 8)), 0, st,
                            dz, y, mean, rstd, w, b, sums2, dpos, nb, HW, D));
    B200_LAUNCH_OK();
    return 0;
  }
  int gy = (int)cdiv((long long)sm_count() * 4, gx);
  int rpb = (int)cdiv(M, gy);
  if (rpb < 64) rpb = 64;
  gy = (int)cdiv(M, rpb);
  B200_CUDA_OK(launch_pdl(bn_bwd_reduce_kernel, dim3(dim3(gx, gy)), dim3(dim3(32, 8)), 0, st, dz, y, mean, rstd, w, b, sums2, M, D, rpb));
  B200_LAUNCH_OK();
  if (dpos) {
    B200_CHECK_ARG(M % HW == 0 && ((long long)HW * D) % 4 == 0, "M must be a multiple of HW");
    return b200_batch_sum(dz, dpos, M / HW, (long long)HW * D, stream);
  }
  return 0;
}

extern "C" int b200_bn_relu_pos_bwd_apply(const float* dz, const float* y, const float* mean, const float* rstd,
                                          const float* w, const float* b, const float* sums2, void* dy_bf16,
                                          int use_batch_stats, int M, int D, void* stream) {
  return b200::bn_relu_pos_bwd_apply_acc(dz, y, mean, rstd, w, b, sums2, dy_bf16, use_batch_stats, M, D, nullptr, nullptr,
                                         stream);
}

int b200::bn_relu_pos_bwd_apply_acc(const float* dz, const float* y, const float* mean, const float* rstd, const float* w,
                                    const float* b, const float* sums2, void* dy_bf16, int use_batch_stats, int M, int D,
                                    float* acc_bn_b, float* acc_bn_w, void* stream) {
  B200_CHECK_ARG(dz && y && mean && rstd && w && b && dy_bf16 && M > 0 && D % 4 == 0, "bad args");
  B200_CHECK_ARG(!use_batch_stats || sums2, "batch statistics need sums2");
  B200_CHECK_ARG((acc_bn_b == nullptr) == (acc_bn_w == nullptr) && (acc_bn_b == nullptr || sums2 != nullptr), "bad accumulators");
  const int gx = (int)cdiv(D, 128);
  int gy = (int)cdiv((long long)sm_count() * 6, gx);
  int rpb = (int)cdiv(M, gy);
  if (rpb < 32) rpb = 32;
  gy = (int)cdiv(M, rpb);
  B200_CUDA_OK(launch_pdl(bn_bwd_apply_kernel, dim3(gx, gy), dim3(32, 8), 0, static_cast<cudaStream_t>(stream),
      dz, y, mean, rstd, w, b, sums2, static_cast<__nv_bfloat16*>(dy_bf16), use_batch_stats, (long long)M, D, acc_bn_b, acc_bn_w,
      rpb));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_colsum(const void* x, int x_is_bf16, long long ldx, float* out, int rows, int cols, void* stream) {
  B200_CHECK_ARG(x && out && rows > 0 && cols > 0 && cols % 4 == 0 && ldx % 4 == 0, "bad args");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (x_is_bf16) return launch_colreduce<__nv_bfloat16, false>(static_cast<const __nv_bfloat16*>(x), ldx, out, rows, cols, st);
  return launch_colreduce<float, false>(static_cast<const float*>(x), ldx, out, rows, cols, st);
}

extern "C" int b200_batch_sum(const float* x, float* out, int B, long long n, void* stream) {
  B200_CHECK_ARG(x && out && B > 0 && n > 0 && n % 4 == 0, "bad args");
  B200_CUDA_OK(launch_pdl(batch_sum_kernel, dim3(grid_for(n / 4, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, out, B, n, 1));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_swiglu(const void* x12, void* out, int rows, int H, void* stream) {
  B200_CHECK_ARG(x12 && out && rows > 0 && H > 0 && H % 8 == 0, "bad args");
  B200_CUDA_OK(launch_pdl(swiglu_kernel, dim3(grid_for((long long)rows * H / 8, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(x12), static_cast<__nv_bfloat16*>(out), rows, H));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_batch_sum_bf16(const void* x, float* out_f32, void* out_bf16, int B, long long n, void* stream) {
  B200_CHECK_ARG(x && (out_f32 || out_bf16) && B > 0 && n > 0 && n % 4 == 0, "bad args");
  B200_CUDA_OK(launch_pdl(batch_sum_bf16_kernel, dim3((unsigned)cdiv(n / 4, 32)), dim3(32, 8), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(x), out_f32, static_cast<__nv_bfloat16*>(out_bf16), B, n));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_axpy(const float* x, float* y, float a, long long n, void* stream) {
  B200_CHECK_ARG(x && y && n > 0, "bad args");
  B200_CUDA_OK(launch_pdl(axpy_kernel, dim3(grid_for(n, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, y, a, n));
  B200_LAUNCH_OK();
  return 0;
}

extern "C" int b200_swiglu_bwd(const void* x12, const void* d_out, void* d_x12, int rows, int H, void* stream) {
  B200_CHECK_ARG(x12 && d_out && d_x12 && rows > 0 && H > 0 && H % 8 == 0, "bad args");
  B200_CUDA_OK(launch_pdl(swiglu_bwd_kernel, dim3(grid_for((long long)rows * H / 8, 256)), dim3(256), 0, static_cast<cudaStream_t>(stream), 
      static_cast<const __nv_bfloat16*>(x12), static_cast<const __nv_bfloat16*>(d_out),
      static_cast<__nv_bfloat16*>(d_x12), rows, H));
  B200_LAUNCH_OK();
  return 0;
}

namespace b200 {
int zero_f32(float* p, long long n, cudaStream_t st) {
  if (n <= 0) return 0;
  B200_CUDA_OK(launch_pdl(zero_kernel, dim3(grid_for(n, 1024, 4)), dim3(256), 0, st, p, n));
  B200_LAUNCH_OK();
  return 0;
}
}  // namespace b200
