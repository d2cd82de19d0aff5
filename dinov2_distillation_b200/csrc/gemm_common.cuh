// Shared pieces of the tcgen05 GEMM kernels (1-CTA and CTA-pair variants): parameters and the register epilogue.
#pragma once
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/b200_distill.h"

namespace b200 {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int A_TILE_BYTES = BM * BK * 2;
constexpr int EPI_WARPS = 16;                      // four warps per TMEM lane quadrant
constexpr int EPI_W = 16;                          // accumulator columns per tcgen05.ld (keeps the epilogue < 100 registers)
constexpr int GEMM_THREADS = 128 + EPI_WARPS * 32;   // warps 0-3: TMA / MMA / TMEM alloc / spare
constexpr int EPI_STAGE_BYTES = 0;

struct GemmParams {
  int M, N, K;
  int m_tiles, n_tiles, num_kb, kb_per_split, split_k, total_tiles;
  const float* bias;
  int act;
  const __nv_bfloat16* aux;
  long long ldaux;
  int aux_mode;
  const float* col_scale;
  const float* residual;
  long long ldres;
  int res_row_period;
  float* out_f32;
  long long ldo32;
  int atomic_add;
  __nv_bfloat16* out_bf16;
  long long ldo16;
  __nv_bfloat16* out_bf16_pre;
  long long ldo16_pre;
  int out_row_period, out_row_pad;
  int out_bp;                // fp32 output batched along columns with this period (3-D output tensor map); 0 = plain 2-D
  int vec_ok;  // all leading dims / pointers allow 16-byte vector access
  int out16_fp16, aux_fp16;  // 16-bit output / aux element type: 0 = bf16, 1 = fp16
  int aux_deep;              // gemm_v2: aux tiles requested two units ahead across tiles (option gemm_aux_deep)
  float* colsum;             // gemm_v2.cu: fp32 [N], += column sums of the 16-bit output as stored (or NULL)
  int pre_alt;               // 1: the second 16-bit output is the final value in the other format (gemm_v2.cu only)
  uint32_t idesc;            // tcgen05 instruction descriptor (operand formats, majors, tile shape)
  float algo_scale;          // profiling: algorithmic flops / executed flops
  int dbg;                   // B200_GEMM_DBG experiments: bit 0 = drain TMEM but skip the epilogue math and stores
  long long* dbg_buf;        // B200_GEMM_DBG bit 5: per-role wait-cycle counters of CTA 0 (gemm_v2.cu)
};

// Epilogue for 32 consecutive columns of one accumulator row (thread = row): 128-bit vector loads / stores along the
// row. (A shared-memory transposed, lane = column variant was measured 3x slower: the epilogue is instruction bound.)
template <int W>
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, bool row_ok, int row, int n0, float (&v)[W],
                                               bool first_split) {
  if (!row_ok || n0 >= p.N) return;
  const bool full = (n0 + W <= p.N) && p.vec_ok;
  if (p.bias != nullptr && first_split) {
    if (full) {
#pragma unroll
      for (int j = 0; j < W; j += 4) {
        float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + n0 + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (n0 + j < p.N) v[j] += __ldg(p.bias + n0 + j);
    }
  }
  long long orow = row;
  if (p.out_row_period > 0) {
    orow = (long long)(row / p.out_row_period) * (p.out_row_period + p.out_row_pad) + p.out_row_pad +
           row % p.out_row_period;
  }
  if (p.out_bf16_pre != nullptr) {
    __nv_bfloat16* dst = p.out_bf16_pre + orow * p.ldo16_pre + n0;
    if (full) {
#pragma unroll
      for (int j = 0; j < W; j += 8) {
        uint4 u;
        u.x = pack16(v[j], v[j + 1], p.out16_fp16); u.y = pack16(v[j + 2], v[j + 3], p.out16_fp16);
        u.z = pack16(v[j + 4], v[j + 5], p.out16_fp16); u.w = pack16(v[j + 6], v[j + 7], p.out16_fp16);
        *reinterpret_cast<uint4*>(dst + j) = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (n0 + j < p.N) store16(dst + j, v[j], p.out16_fp16);
    }
  }
  if (p.act == B200_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < W; ++j) v[j] = gelu_fast(v[j]);
  } else if (p.act == B200_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < W; ++j) v[j] = fmaxf(v[j], 0.0f);
  }
  if (p.aux_mode != B200_AUX_NONE) {
    const __nv_bfloat16* ax = p.aux + (long long)row * p.ldaux + n0;
    float a[W];
    if (full) {
#pragma unroll
      for (int j = 0; j < W; j += 8) {
        uint4 u = __ldg(reinterpret_cast<const uint4*>(ax + j));
        float2 f0 = unpack16(u.x, p.aux_fp16), f1 = unpack16(u.y, p.aux_fp16), f2 = unpack16(u.z, p.aux_fp16), f3 = unpack16(u.w, p.aux_fp16);
        a[j] = f0.x; a[j + 1] = f0.y; a[j + 2] = f1.x; a[j + 3] = f1.y;
        a[j + 4] = f2.x; a[j + 5] = f2.y; a[j + 6] = f3.x; a[j + 7] = f3.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) a[j] = (n0 + j < p.N) ? load16(ax + j, p.aux_fp16) : 0.0f;
    }
    if (p.aux_mode == B200_AUX_DGELU) {
#pragma unroll
      for (int j = 0; j < W; ++j) v[j] *= dgelu_fast(a[j]);
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j) v[j] = a[j] > 0.0f ? v[j] : 0.0f;
    }
  }
  if (p.col_scale != nullptr) {
    if (full) {
#pragma unroll
      for (int j = 0; j < W; j += 4) {
        float4 g = __ldg(reinterpret_cast<const float4*>(p.col_scale + n0 + j));
        v[j] *= g.x; v[j + 1] *= g.y; v[j + 2] *= g.z; v[j + 3] *= g.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (n0 + j < p.N) v[j] *= __ldg(p.col_scale + n0 + j);
    }
  }
  if (p.residual != nullptr && first_split) {
    const long long rr = p.res_row_period > 0 ? (row % p.res_row_period) : orow;
    const float* rs = p.residual + rr * p.ldres + n0;
    if (full) {
#pragma unroll
      for (int j = 0; j < W; j += 4) {
        float4 r = *reinterpret_cast<const float4*>(rs + j);
        v[j] += r.x; v[j + 1] += r.y; v[j + 2] += r.z; v[j + 3] += r.w;
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (n0 + j < p.N) v[j] += rs[j];
    }
  }
  if (p.out_f32 != nullptr) {
    float* dst = p.out_f32 + orow * p.ldo32 + n0;
    if (p.atomic_add) {
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (n0 + j < p.N) atomicAdd(dst + j, v[j]);
    } else if (full) {
#pragma unroll
      for (int j = 0; j < W; j += 4)
        *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (n0 + j < p.N) dst[j] = v[j];
    }
  }
  if (p.out_bf16 != nullptr) {
    __nv_bfloat16* dst = p.out_bf16 + orow * p.ldo16 + n0;
    if (full) {
#pragma unroll
      for (int j = 0; j < W; j += 8) {
        uint4 u;
        u.x = pack16(v[j], v[j + 1], p.out16_fp16); u.y = pack16(v[j + 2], v[j + 3], p.out16_fp16);
        u.z = pack16(v[j + 4], v[j + 5], p.out16_fp16); u.w = pack16(v[j + 6], v[j + 7], p.out16_fp16);
        *reinterpret_cast<uint4*>(dst + j) = u;
      }
    } else {
#pragma unroll
      for (int j = 0; j < W; ++j)
        if (n0 + j < p.N) store16(dst + j, v[j], p.out16_fp16);
    }
  }
}


int make_tensor_map_2d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0,
                       uint32_t b1);
// Kernel-internal measurement probes (mainloop-only / epilogue-only floors, wait-cycle counters: DESIGN.md section 4) are
// compiled in only with -DB200_GEMM_PROBES; the production build carries none of their branches.
#ifdef B200_GEMM_PROBES
constexpr bool kGemmProbes = true;
#else
constexpr bool kGemmProbes = false;
#endif
int make_tensor_map_ex(CUtensorMap* out, const void* ptr, int esize, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0,
                       uint32_t b1, int swizzle);
int make_tensor_map_3d(CUtensorMap* out, const void* ptr, int esize, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1,
                       uint64_t ld2, uint32_t b0, uint32_t b1, uint32_t b2, int swizzle);
// second-generation kernel (gemm_v2.cu): returns 1 when the shape is not handled (the caller falls through to the
// first-generation kernel), 0 on success, negative on error
int launch_gemm_v2(const b200_gemm_desc* d, GemmParams& p, cudaStream_t st);

}  // namespace b200
