// tcgen05 GEMM, second generation: the mainloop of gemm_tcgen05.cu / gemm_tcgen05_2cta.cu with an epilogue built for the
// small-K shapes of the ViT-S/B teacher, where the epilogue -- not the tensor pipe -- sets the pace.
//
//   warp 0        : TMA producer (one elected lane)
//   warp 1        : MMA issuer   (one elected lane; the leader CTA only in CTA-pair mode)
//   warp 2        : TMEM allocate / free
//   warps 4..     : epilogue, 8 or 16 warps (template EW; 16 when no operand tile has to be staged). Warp w drains TMEM
//                   lanes 32*(w%4)..+31 (hardware rule: warp id % 4) and every (EW/4)-th 32-column unit of the tile. Per unit: tcgen05.ld 32x32b.x32 (thread = row, 32 fp32 columns) ->
//                   bias / activation / aux / LayerScale / residual in registers -> 16-byte st.shared into a swizzled
//                   32-row staging tile -> ONE bulk tensor store (cp.async.bulk.tensor, or cp.reduce.async.bulk.tensor
//                   .add for in-place residual streams and split-K) issued by lane 0. Row-strided 16-byte global stores
//                   of the first generation (32 sectors per warp store) are gone, tails are clipped by the tensor map,
//                   and operands the epilogue reads (fp32 residual, 16-bit aux) arrive by TMA one unit ahead.
//
// Output modes (template EPI): 0 = 16-bit output (+ optional pre-activation copy, ReLU, dGELU/dReLU aux),
//                              1 = 16-bit output with GELU, 2 = fp32 output (LayerScale, residual, accumulate).
#include "gemm_common.cuh"

#include <stdlib.h>

namespace b200 {

// Epilogue warps per CTA (template EW): 8 warps with 8 KB of staging each (every path), or 16 warps with 4 KB each for
// the epilogues that need no operand tile (no separate residual / aux / pre-activation copy): four warps per scheduler
// instead of two hide the tcgen05.ld -> st.shared -> fence -> bulk-store latency chain that paces the K = 384 shapes.
constexpr int V2_MAX_EPI_WARPS = 16;   // (also the number of operand-tile barriers: one per warp, or two per warp of an 8-warp epilogue)
constexpr int V2_EPI_BYTES = 65536;

template <int BN, bool PAIR>
struct V2Cfg {
  static constexpr int B_ROWS = PAIR ? BN / 2 : BN;   // rows of B this CTA stages per k-block
  static constexpr int B_TILE_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int RING_BUDGET = 232448 - 1024 - 512 - V2_EPI_BYTES;
  static constexpr int STAGES = (RING_BUDGET / STAGE_BYTES) > 8 ? 8 : (RING_BUDGET / STAGE_BYTES);
  static constexpr int TMEM_STRIDE = (BN <= 128) ? 128 : 256;
  static constexpr int TMEM_COLS = 2 * TMEM_STRIDE;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + V2_EPI_BYTES + 1024 + 512;
};

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t v2_cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void v2_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// executed by both CTAs of a pair; clearing the peer bit makes the transaction bytes land on CTA 0's barrier
__device__ __forceinline__ void v2_tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* leader_bar, int c0,
                                                    int c1) {
  const uint32_t bar = smem_u32(leader_bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void v2_mma_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void v2_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void v2_mbar_arrive_cta(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
// accumulator hand-back: the TMEM reads were ordered by tcgen05.wait::ld + tcgen05.fence::before_thread_sync; a release
// at cluster scope would add a MEMBAR.ALL.CTA + ERRBAR per tile and warp (18 % of the stall samples, profiles/r02)
__device__ __forceinline__ void v2_mbar_arrive_cta_relaxed(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}\n"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
template <int kCols>
__device__ __forceinline__ void v2_tmem_alloc_pair(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void v2_tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
// shared -> global bulk tensor store / reduce-add (bulk async-group completion)
__device__ __forceinline__ void v2_tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void v2_tma_reduce_add_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void v2_tma_store_3d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void v2_tma_reduce_add_3d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void v2_tma_load_2d_s(uint32_t smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void v2_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void v2_bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void v2_bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ void v2_sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ uint4 v2_lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}

// GELU(x) = max(x, 0) - |x| * Phi(-|x|), Phi(-a) = 2^q(a) with q a degree-5 fit of log2(Phi(-a)) on [0, 8] (max abs error
// of the product 1.7e-5, relative 3.5e-4 where |gelu| > 1e-4; the leading coefficient is negative, so larger |x| decay
// to the exact limit). One MUFU (ex2) + 7 FMA-pipe instructions per element: the fc1 epilogue was MUFU bound with the
// two-transcendental erf form.
__device__ __forceinline__ float v2_phi_neg(float a) {
  float q = fmaf(-3.046068783e-04f, a, 5.602596781e-03f);
  q = fmaf(q, a, -4.712946504e-02f);
  q = fmaf(q, a, -4.664196592e-01f);
  q = fmaf(q, a, -1.147315491e+00f);
  q = fmaf(q, a, -1.000508484e+00f);
  return ex2_approx(q);
}
__device__ __forceinline__ float v2_gelu(float x) {
  const float a = fabsf(x);
  return fmaf(-a, v2_phi_neg(a), fmaxf(x, 0.0f));
}
// two elements per instruction where the ISA allows it (FFMA2: one issue slot for two fp32 FMAs -- the GELU epilogue is
// issue bound): the five Horner steps and the final fma run packed, abs / max / ex2 stay scalar
__device__ __forceinline__ float2 v2_gelu2(float x0, float x1) {
  const float2 a = make_float2(fabsf(x0), fabsf(x1));
  float2 q = __ffma2_rn(make_float2(-3.046068783e-04f, -3.046068783e-04f), a, make_float2(5.602596781e-03f, 5.602596781e-03f));
  q = __ffma2_rn(q, a, make_float2(-4.712946504e-02f, -4.712946504e-02f));
  q = __ffma2_rn(q, a, make_float2(-4.664196592e-01f, -4.664196592e-01f));
  q = __ffma2_rn(q, a, make_float2(-1.147315491e+00f, -1.147315491e+00f));
  q = __ffma2_rn(q, a, make_float2(-1.000508484e+00f, -1.000508484e+00f));
  const float2 e = make_float2(ex2_approx(q.x), ex2_approx(q.y));
  return __ffma2_rn(make_float2(-a.x, -a.y), e, make_float2(fmaxf(x0, 0.0f), fmaxf(x1, 0.0f)));
}
__device__ __forceinline__ float v2_dgelu(float x) {
  const float e = v2_phi_neg(fabsf(x));
  const float cdf = x > 0.0f ? 1.0f - e : e;
  const float pdf = 0.39894228040143268f * ex2_approx(-0.72134752044448170f * x * x);
  return fmaf(x, pdf, cdf);
}

// ------------------------------------------------------------------------------------------------ epilogue of one tile
// One warp: 32 rows (TMEM lanes of its quadrant) x the units u = half, half + NG, ... of the tile's BN / 32 column units
// (NG = EW / 4 warps share a quadrant).
template <int BN, int EPI, int EW>
__device__ __forceinline__ void v2_epilogue_tile(const GemmParams& p, const CUtensorMap* tmO, const CUtensorMap* tmX,
                                                 uint32_t taddr, int row0, int n0, uint32_t stage_smem, uint64_t* xbar,
                                                 uint32_t& xphase, int half, int lane, bool first_split, bool use_x_arg,
                                                 bool reduce_out, uint64_t* tempty, bool pair, uint32_t* out_toggle, uint32_t vec_smem,
                                                 uint32_t bias_row_smem = 0, bool deep_arg = false, int next_row0 = -1,
                                                 int next_n0 = 0) {
  constexpr int NG = EW / 4;                             // warps per TMEM quadrant
  constexpr int UNITS = (BN / 32 + NG - 1) / NG;         // units per warp (the last may be absent: BN = 192, NG = 4)
  constexpr bool PREFETCH = EW == 8;                     // 8 warps: next unit's tcgen05.ld in flight during this one
  constexpr bool F32 = EPI == 2;
  const bool use_x = EW == 8 && use_x_arg;               // 16 warps: 4 KB of staging per warp, no operand tile
  // Aux tiles two units ahead, across tiles (x relu mask / x gelu' of the dgrad-through-activation GEMMs; BN = 192: three
  // units per warp and tile). Requested one unit ahead, the 32 x 32 tile arrived ~540 cycles late in every unit (wait
  // cycles of a -DB200_GEMM_PROBES build: 2 600 cycles per unit against 1 600 without the aux operand). Two aux tiles
  // rotate over this warp's unit sequence g = 3 * tile + unit: buffer / barrier g & 1 (staging bytes 4096.. and 6144..,
  // the latter free because these GEMMs carry no bias vector; barriers xbar and xbar + 8), and the tile of unit g + 2 --
  // unit 2 of this tile, or unit 0 / 1 of the warp's NEXT tile -- is requested as soon as unit g has been read.
  // xphase: bit 0 / 1 = parity of the two barriers, bits 2.. = g.
  constexpr bool DEEP_OK = EW == 8 && !F32 && BN == 192;
  const bool deep = DEEP_OK && deep_arg && use_x;
  constexpr int ROWB = F32 ? 128 : 64;                   // staging row bytes (32 columns)
  constexpr int CHUNKS = ROWB / 16;
  constexpr uint32_t UNIT_BYTES = 32 * ROWB;
  const uint32_t sw = F32 ? (lane & 7) : ((lane >> 1) & 3);
  const uint32_t row_off = lane * ROWB;
  // staging tiles of this warp: F32: [0] = output, [1] = residual load. 16-bit: [0],[1] = outputs (alternating; [1] is
  // the pre-activation copy when one is requested), [2] = aux load.
  const uint32_t buf_x = stage_smem + (F32 ? 4096u : 4096u);
  const bool has_pre = !F32 && p.out_bf16_pre != nullptr;

  // valid units of this warp (columns beyond N are clipped; nu is warp-uniform)
  int nu = 0;
#pragma unroll
  for (int i = 0; i < UNITS; ++i)
    if ((half + NG * i) * 32 < BN && n0 + (half + NG * i) * 32 < p.N) nu = i + 1;
  if (nu == 0) {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) {
      if (pair) v2_mbar_arrive_cta_relaxed(tempty, 0);
      else mbar_arrive(tempty);
    }
    return;
  }

  uint32_t raw[PREFETCH ? 2 : 1][32];
  // residual rows repeat with period res_row_period (pos_embed under the patch-embed GEMM): 32-row units never straddle
  const int xrow = (F32 && p.res_row_period > 0) ? row0 % p.res_row_period : row0;
  if (use_x && !deep && lane == 0) {
    mbar_expect_tx(xbar, UNIT_BYTES);
    v2_tma_load_2d_s(buf_x, tmX, xbar, n0 + half * 32, xrow);
  }
  tmem_ld_32x32(taddr + half * 32, raw[0]);
#pragma unroll
  for (int i = 0; i < UNITS; ++i) {
    if (i >= nu) break;
    const int u = half + NG * i;
    const int c0 = n0 + u * 32;
    const bool next_ok = i + 1 < nu;
    tmem_ld_wait();
    if (PREFETCH && next_ok) {
      tmem_ld_32x32(taddr + (u + NG) * 32, raw[(i + 1) & 1]);
    } else if (!next_ok) {
      // all tcgen05.ld of this tile have completed: hand the accumulator buffer back to the MMA warp right away
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (pair) v2_mbar_arrive_cta_relaxed(tempty, 0);
        else mbar_arrive(tempty);
      }
    }
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(raw[PREFETCH ? (i & 1) : 0][j]);
    if (!PREFETCH && next_ok) tmem_ld_32x32(taddr + (u + NG) * 32, raw[0]);   // values copied out: refill
    if (kGemmProbes && (p.dbg & 1)) {   // B200_GEMM_DBG=1: mainloop-only floor (drain TMEM, no epilogue math, no stores)
      if (v[0] == 123.456f && p.out_f32) p.out_f32[0] = v[1];
      if (use_x) { mbar_wait(xbar, xphase); xphase ^= 1; __syncwarp();
        if (next_ok && lane == 0) { mbar_expect_tx(xbar, UNIT_BYTES); v2_tma_load_2d_s(buf_x, tmX, xbar, c0 + NG * 32, xrow); } }
      continue;
    }
    // (N is a multiple of 32 on this path -- launch_gemm_v2 sends ragged N to the first-generation kernel -- so there
    // is no per-column tail code: the if-converted tail branches were ~170 predicated-off issue slots per unit)
    if (p.bias != nullptr && first_split) {
      if (bias_row_smem != 0) {   // the whole bias vector sits in shared memory (row-panel kernels: every column tile passes)
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const uint4 q = v2_lds128(bias_row_smem + (c0 + j) * 4);
          v[j] += __uint_as_float(q.x); v[j + 1] += __uint_as_float(q.y);
          v[j + 2] += __uint_as_float(q.z); v[j + 3] += __uint_as_float(q.w);
        }
      } else if (vec_smem != 0) {   // staged in this warp's spare shared memory before the accumulator wait (broadcast reads)
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const uint4 q = v2_lds128(vec_smem + i * 128 + j * 4);
          v[j] += __uint_as_float(q.x); v[j + 1] += __uint_as_float(q.y);
          v[j + 2] += __uint_as_float(q.z); v[j + 3] += __uint_as_float(q.w);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c0 + j));
          v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
        }
      }
    }
    if constexpr (!F32) {
      // ---- 16-bit output
      uint32_t buf_o = stage_smem + ((*out_toggle & 1u) ? 2048u : 0u);
      if (has_pre) buf_o = stage_smem;
      // the staging tile about to be overwritten must have been read by its previous bulk store
      if (lane == 0) {
        if (has_pre) v2_bulk_wait_read<0>();
        else v2_bulk_wait_read<1>();
      }
      __syncwarp();
      if (has_pre && !p.pre_alt) {   // copy of (acc + bias) before the activation, same format
        const uint32_t bp = stage_smem + 2048u + row_off;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          const int j = c * 8;
          v2_sts128(bp + ((c ^ sw) << 4), pack16(v[j], v[j + 1], p.out16_fp16), pack16(v[j + 2], v[j + 3], p.out16_fp16),
                    pack16(v[j + 4], v[j + 5], p.out16_fp16), pack16(v[j + 6], v[j + 7], p.out16_fp16));
        }
      }
      if constexpr (EPI == 1) {
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float2 r = v2_gelu2(v[j], v[j + 1]);
          v[j] = r.x; v[j + 1] = r.y;
        }
      } else {
        if (p.act == B200_ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        }
        if (use_x) {
          const uint32_t xb_i = deep ? ((xphase >> 2) & 1u) : 0u;          // buffer / barrier of this unit
          uint64_t* xb = xbar + 8 * xb_i;
          const uint32_t bx_tile = buf_x + 2048u * xb_i;
          if (deep) {
            mbar_wait(xb, (xphase >> xb_i) & 1u);
            xphase = (xphase ^ (1u << xb_i)) + 4u;                         // flip the parity bit, g += 1
          } else {
            mbar_wait(xbar, xphase);
            xphase ^= 1;
          }
          const uint32_t bx = bx_tile + row_off;
          float a[32];
#pragma unroll
          for (int c = 0; c < CHUNKS; ++c) {
            const uint4 q = v2_lds128(bx + ((c ^ sw) << 4));
            const float2 f0 = unpack16(q.x, p.aux_fp16), f1 = unpack16(q.y, p.aux_fp16), f2 = unpack16(q.z, p.aux_fp16),
                         f3 = unpack16(q.w, p.aux_fp16);
            const int j = c * 8;
            a[j] = f0.x; a[j + 1] = f0.y; a[j + 2] = f1.x; a[j + 3] = f1.y;
            a[j + 4] = f2.x; a[j + 5] = f2.y; a[j + 6] = f3.x; a[j + 7] = f3.y;
          }
          __syncwarp();   // every lane has read the aux tile: it may be refilled
          if (lane == 0) {
            if (deep) {   // unit g + 2 into the tile just read: unit 2 of this tile, or unit i - 1 of the next
              if (i == 0) {
                mbar_expect_tx(xb, UNIT_BYTES);
                v2_tma_load_2d_s(bx_tile, tmX, xb, n0 + (half + 2 * NG) * 32, row0);
              } else if (next_row0 >= 0) {
                mbar_expect_tx(xb, UNIT_BYTES);
                v2_tma_load_2d_s(bx_tile, tmX, xb, next_n0 + (half + (i - 1) * NG) * 32, next_row0);
              }
            } else if (next_ok) {
              mbar_expect_tx(xbar, UNIT_BYTES);
              v2_tma_load_2d_s(buf_x, tmX, xbar, c0 + NG * 32, xrow);
            }
          }
          if (p.aux_mode == B200_AUX_DGELU) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] *= v2_dgelu(a[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = a[j] > 0.0f ? v[j] : 0.0f;
          }
        }
      }
      if (has_pre && p.pre_alt) {    // copy of the FINAL value in the other 16-bit format
        const uint32_t bp = stage_smem + 2048u + row_off;
        const int alt = !p.out16_fp16;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          const int j = c * 8;
          v2_sts128(bp + ((c ^ sw) << 4), pack16(v[j], v[j + 1], alt), pack16(v[j + 2], v[j + 3], alt),
                    pack16(v[j + 4], v[j + 5], alt), pack16(v[j + 6], v[j + 7], alt));
        }
      }
      const uint32_t bo = buf_o + row_off;
      if (kGemmProbes && (p.dbg & 16)) {
      } else if (p.out16_fp16) {
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          const int j = c * 8;
          v2_sts128(bo + ((c ^ sw) << 4), pack16(v[j], v[j + 1], 1), pack16(v[j + 2], v[j + 3], 1),
                    pack16(v[j + 4], v[j + 5], 1), pack16(v[j + 6], v[j + 7], 1));
        }
      } else {
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          const int j = c * 8;
          v2_sts128(bo + ((c ^ sw) << 4), pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]),
                    pack_bf16(v[j + 4], v[j + 5]), pack_bf16(v[j + 6], v[j + 7]));
        }
      }
      if (p.colsum != nullptr) {
        // column sums of the 16-bit output exactly as stored (a bias gradient: dW1's bias = column sums of dh): lane j
        // walks column j of the staged 32 x 32 tile (64-byte rows, 16-byte chunks swizzled by (row >> 1) & 3)
        __syncwarp();
        const int nrow = min(32, p.M - row0);
        const uint32_t cbase = buf_o + (lane & 7) * 2;
        const uint32_t chunk = lane >> 3;
        float cs = 0.0f;
#pragma unroll 8
        for (int r = 0; r < nrow; ++r) {
          uint16_t h;
          asm volatile("ld.shared.u16 %0, [%1];" : "=h"(h) : "r"(cbase + r * 64 + ((chunk ^ ((r >> 1) & 3)) << 4)));
          cs += p.out16_fp16 ? __half2float(__ushort_as_half(h)) : __bfloat162float(__ushort_as_bfloat16(h));
        }
        atomicAdd(p.colsum + c0 + lane, cs);
      }
      if (!(kGemmProbes && (p.dbg & 8))) fence_proxy_async();
      __syncwarp();
      if (lane == 0 && !(kGemmProbes && (p.dbg & 2))) {
        v2_tma_store_2d(tmO, buf_o, c0, row0);
        if (has_pre)
          v2_tma_store_2d(tmX, stage_smem + 2048u, c0, row0);
        v2_bulk_commit();
      }
      *out_toggle ^= 1u;
    } else {
      // ---- fp32 output
      if (p.col_scale != nullptr) {
        if (vec_smem != 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const uint4 q = v2_lds128(vec_smem + 512 + i * 128 + j * 4);
            v[j] *= __uint_as_float(q.x); v[j + 1] *= __uint_as_float(q.y);
            v[j + 2] *= __uint_as_float(q.z); v[j + 3] *= __uint_as_float(q.w);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(p.col_scale + c0 + j));
            v[j] *= g.x; v[j + 1] *= g.y; v[j + 2] *= g.z; v[j + 3] *= g.w;
          }
        }
      }
      if (use_x) {
        mbar_wait(xbar, xphase);
        xphase ^= 1;
        const uint32_t bx = buf_x + row_off;
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          const uint4 q = v2_lds128(bx + ((c ^ sw) << 4));
          const int j = c * 4;
          v[j] += __uint_as_float(q.x); v[j + 1] += __uint_as_float(q.y);
          v[j + 2] += __uint_as_float(q.z); v[j + 3] += __uint_as_float(q.w);
        }
        __syncwarp();
        if (next_ok && lane == 0) {
          mbar_expect_tx(xbar, UNIT_BYTES);
          v2_tma_load_2d_s(buf_x, tmX, xbar, c0 + NG * 32, xrow);
        }
      }
      if (lane == 0) v2_bulk_wait_read<0>();
      __syncwarp();
      const uint32_t bo = stage_smem + row_off;
      if (!(kGemmProbes && (p.dbg & 16))) {
#pragma unroll
        for (int c = 0; c < CHUNKS; ++c) {
          const int j = c * 4;
          v2_sts128(bo + ((c ^ sw) << 4), __float_as_uint(v[j]), __float_as_uint(v[j + 1]), __float_as_uint(v[j + 2]),
                    __float_as_uint(v[j + 3]));
        }
      } else if (v[3] == 123.456f) p.out_f32[1] = v[5];
      if (!(kGemmProbes && (p.dbg & 8))) fence_proxy_async();
      __syncwarp();
      if (lane == 0 && !(kGemmProbes && (p.dbg & 2))) {
        if (p.out_row_period > 0) {   // rows batched with a gap (cls rows of the token buffer): (column, row in batch, batch)
          const int bi = row0 / p.out_row_period;
          v2_tma_store_3d(tmO, stage_smem, c0, row0 - bi * p.out_row_period, bi);
        } else if (p.out_bp > 0) {   // columns batched with period out_bp: (column in batch, row, batch)
          const int bi = c0 / p.out_bp, cb = c0 - bi * p.out_bp;
          if (reduce_out) v2_tma_reduce_add_3d(tmO, stage_smem, cb, row0, bi);
          else v2_tma_store_3d(tmO, stage_smem, cb, row0, bi);
        } else if (reduce_out) {
          v2_tma_reduce_add_2d(tmO, stage_smem, c0, row0);
        } else {
          v2_tma_store_2d(tmO, stage_smem, c0, row0);
        }
        v2_bulk_commit();
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ kernel
template <int BN, bool PAIR, bool MN, int EPI, int EW>
__global__ void __launch_bounds__(128 + EW * 32, 1)
gemm_v2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmX, const GemmParams p,
               const int use_x, const int reduce_out) {
  using Cfg = V2Cfg<BN, PAIR>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int TILE_M = PAIR ? 256 : 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + V2_EPI_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* x_bar = tempty_bar + 2;                      // one per epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(x_bar + V2_MAX_EPI_WARPS);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? v2_cluster_ctarank() : 0u;
  const bool leader = rank == 0;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    if (use_x || (EPI != 2 && p.out_bf16_pre != nullptr)) tma_prefetch_desc(&tmX);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], (PAIR ? 2 : 1) * EW);
    }
    for (int s = 0; s < V2_MAX_EPI_WARPS; ++s) mbar_init(&x_bar[s], 1);   // (8-warp epilogues: [ew] and [ew + 8])
    fence_barrier_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) v2_tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
    else tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  }
  tc_fence_before();
  if constexpr (PAIR) v2_cluster_sync();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // PDL: everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the previous kernel's tail
  pdl_trigger();
  pdl_wait();

  const int tiles_mn = p.m_tiles * p.n_tiles;
  const int worker = PAIR ? (blockIdx.x >> 1) : blockIdx.x;
  const int workers = PAIR ? (gridDim.x >> 1) : gridDim.x;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      long long w_empty = 0;
      const long long t_begin = clock64();
      for (int t = worker; t < p.total_tiles; t += workers) {
        const int ks = t / tiles_mn;
        const int r = t - ks * tiles_mn;
        const int m0 = (r / p.n_tiles) * TILE_M + (int)rank * 128;
        const int n0 = (r % p.n_tiles) * BN + (int)rank * Cfg::B_ROWS * (PAIR ? 1 : 0);
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = (kGemmProbes && (p.dbg & 4)) ? kb0 : min(p.num_kb, kb0 + p.kb_per_split);   // B200_GEMM_DBG=4: epilogue-only floor
        for (int kb = kb0; kb < kb1; ++kb) {
          if (kGemmProbes && p.dbg_buf) { const long long c = clock64(); mbar_wait(&empty_bar[stage], phase ^ 1); w_empty += clock64() - c; }
          else mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + A_TILE_BYTES;
          if constexpr (PAIR) {
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            v2_tma_load_2d_pair(sA, &tmA, &full_bar[stage], kb * BK, m0);
            v2_tma_load_2d_pair(sB, &tmB, &full_bar[stage], kb * BK, n0);
          } else {
            mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            if constexpr (!MN) {
              tma_load_2d(sA, &tmA, &full_bar[stage], kb * BK, m0);
              tma_load_2d(sB, &tmB, &full_bar[stage], kb * BK, n0);
            } else {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_2d(sA + j * 8192, &tmA, &full_bar[stage], m0 + j * 64, kb * BK);
#pragma unroll
              for (int j = 0; j < BN / 64; ++j) tma_load_2d(sB + j * 8192, &tmB, &full_bar[stage], n0 + j * 64, kb * BK);
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
      if (kGemmProbes && p.dbg_buf && blockIdx.x == 0) { p.dbg_buf[0] = clock64() - t_begin; p.dbg_buf[1] = w_empty; }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      const uint32_t idesc = p.idesc;
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      long long w_full = 0, w_tempty = 0, n_t = 0;
      const long long t_begin = clock64();
      for (int t = worker; t < p.total_tiles; t += workers) {
        const int ks = t / tiles_mn;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = (kGemmProbes && (p.dbg & 4)) ? kb0 : min(p.num_kb, kb0 + p.kb_per_split);
        ++n_t;
        if (kGemmProbes && p.dbg_buf) { const long long c = clock64(); mbar_wait(&tempty_bar[as], aphase ^ 1); w_tempty += clock64() - c; }
        else mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * Cfg::TMEM_STRIDE;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (kGemmProbes && p.dbg_buf) { const long long c = clock64(); mbar_wait(&full_bar[stage], phase); w_full += clock64() - c; }
          else mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = MN ? make_smem_desc_sw128(a_addr + k * 2048, 8192, 1024)
                                   : make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db = MN ? make_smem_desc_sw128(b_addr + k * 2048, 8192, 1024)
                                   : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            const uint32_t acc = (kb > kb0 || k > 0) ? 1u : 0u;
            if constexpr (PAIR) v2_mma_pair(d_tmem, da, db, idesc, acc);
            else tc_mma_bf16(d_tmem, da, db, idesc, acc);
          }
          if constexpr (PAIR) v2_commit_pair(&empty_bar[stage]);
          else tc_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if constexpr (PAIR) v2_commit_pair(&tfull_bar[as]);
        else tc_commit(&tfull_bar[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
      if (kGemmProbes && p.dbg_buf && blockIdx.x == 0) {
        p.dbg_buf[2] = clock64() - t_begin; p.dbg_buf[3] = w_full; p.dbg_buf[4] = w_tempty; p.dbg_buf[5] = n_t;
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int quad = warp & 3;            // TMEM lanes 32*quad .. +31 (hardware: warp id % 4)
    const int half = ew >> 2;             // which of the EW / 4 interleaved unit sets
    const uint32_t stage_smem = smem_u32(epi_smem + ew * (V2_EPI_BYTES / EW));
    uint64_t* xbar = &x_bar[ew];
    uint32_t xphase = 0;
    uint32_t out_toggle = 0;
    int as = 0;
    uint32_t aphase = 0;
    long long w_tfull = 0;
    const long long t_begin = clock64();
    // aux tiles two units ahead (v2_epilogue_tile): every column tile full (three live units per warp), no bias vector in the
    // staging bytes the second aux tile uses, no split-K
    const bool deep = EW == 8 && EPI != 2 && BN == 192 && !PAIR && use_x != 0 && p.bias == nullptr && p.out_bf16_pre == nullptr &&
                      p.N % 192 == 0 && p.split_k == 1 && p.aux_deep != 0;
    auto tile_origin = [&](int t, int& m0, int& n0) {
      const int r = t % tiles_mn;
      m0 = (r / p.n_tiles) * TILE_M + (int)rank * 128;
      n0 = (r % p.n_tiles) * BN;
    };
    if (deep && worker < p.total_tiles && lane == 0) {   // units 0 and 1 of the first tile
      int m0, n0;
      tile_origin(worker, m0, n0);
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        mbar_expect_tx(xbar + 8 * i, 2048u);
        v2_tma_load_2d_s(stage_smem + 4096u + 2048u * i, &tmX, xbar + 8 * i, n0 + (half + 2 * i) * 32, m0 + quad * 32);
      }
    }
    for (int t = worker; t < p.total_tiles; t += workers) {
      const int ks = t / tiles_mn;
      const int r = t - ks * tiles_mn;
      const int m0 = (r / p.n_tiles) * TILE_M + (int)rank * 128;
      const int n0 = (r % p.n_tiles) * BN;
      int next_row0 = -1, next_n0 = 0;
      if (deep && t + workers < p.total_tiles) {
        int nm0;
        tile_origin(t + workers, nm0, next_n0);
        next_row0 = nm0 + quad * 32;
      }
      // per-column epilogue vectors (bias, LayerScale) of this warp's units -> its spare staging bytes, while the
      // accumulator is still being produced: first touch of a new column range is an L2 round trip per unit otherwise
      // (16-bit: bytes 6144.. are free; fp32: the residual tile's 4 KB when no residual is loaded)
      uint32_t vec_smem = 0;
      if (EW == 8 && (EPI != 2 || !use_x)) {
        vec_smem = stage_smem + (EPI != 2 ? 6144u : 4096u);
        const int ui = lane >> 3, c = n0 + (half + 2 * ui) * 32 + (lane & 7) * 4;
        if (ui < BN / 64 && c < p.N) {
          if (p.bias != nullptr) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + c));
            v2_sts128(vec_smem + ui * 128 + (lane & 7) * 16, __float_as_uint(b.x), __float_as_uint(b.y), __float_as_uint(b.z), __float_as_uint(b.w));
          }
          if (EPI == 2 && p.col_scale != nullptr) {
            const float4 g = __ldg(reinterpret_cast<const float4*>(p.col_scale + c));
            v2_sts128(vec_smem + 512 + ui * 128 + (lane & 7) * 16, __float_as_uint(g.x), __float_as_uint(g.y), __float_as_uint(g.z), __float_as_uint(g.w));
          }
        }
        __syncwarp();
      }
      if (kGemmProbes && p.dbg_buf) { const long long c = clock64(); mbar_wait(&tfull_bar[as], aphase); w_tfull += clock64() - c; }
      else mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * Cfg::TMEM_STRIDE;
      v2_epilogue_tile<BN, EPI, EW>(p, &tmO, &tmX, taddr, m0 + quad * 32, n0, stage_smem, xbar, xphase, half, lane, ks == 0,
                                use_x != 0, reduce_out != 0, &tempty_bar[as], PAIR, &out_toggle, vec_smem, 0u, deep, next_row0,
                                next_n0);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    const long long t_loop = clock64() - t_begin;
    // bulk stores read shared memory asynchronously: the CTA must not exit before they are done
    if (lane == 0) v2_bulk_wait_all();
    __syncwarp();
    if (kGemmProbes && p.dbg_buf && blockIdx.x == 0 && lane == 0 && (ew == 0 || ew == EW - 1)) {
      const int o = ew == 0 ? 6 : 9;
      p.dbg_buf[o] = t_loop; p.dbg_buf[o + 1] = w_tfull; p.dbg_buf[o + 2] = clock64() - t_begin;
    }
  }

  tc_fence_before();
  if constexpr (PAIR) v2_cluster_sync();
  else __syncthreads();
  if (warp == 2) {
    if constexpr (PAIR) v2_tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
    else tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------ LayerNorm-prologue GEMM
// C[M, N] = epilogue(LN(x)[M, K] W[N, K]^T) for K = D <= 384 (the ViT-S teacher's LN1 -> qkv and LN2 -> fc1: hub Block.forward,
// called via models/backbones/dinov2.py:32 and train/distillation_module.py:177). One CTA owns a 128-row panel:
//   * its epilogue warps read the panel's fp32 rows ONCE (one warp per row, row in registers, the arithmetic of
//     layernorm_fwd_kernel), normalise, and write the bf16 operand straight into shared memory in the 128-byte-swizzled
//     K-major layout the MMA descriptors expect -- the [M, D] bf16 round trip through HBM and the LayerNorm launch are gone,
//     and the A operand is staged once instead of once per column tile;
//   * the panel then stays resident while W streams through a TMA ring; CTAs run as pairs (cta_group::2: each CTA stages
//     half of every W tile, the leader issues 256 x BN MMAs for both), so each W tile leaves L2 once per 256 rows;
//   * accumulators double-buffered in TMEM, epilogue = v2_epilogue_tile (bias / GELU / pre-activation copy, bulk stores).
struct LnPrologue {
  const float* x; long long ldx;
  const float* w; const float* b; float eps;
  float* mean; float* rstd;   // optional outputs [M] (saved for the block's input-gradient pass)
};

constexpr int LNG_KB_MAX = 6;   // K <= 384

template <int BN>
struct LnCfg {
  static constexpr int A_BYTES = LNG_KB_MAX * A_TILE_BYTES;        // resident panel: 128 rows x 384 bf16
  static constexpr int B_STAGE = (BN / 2) * BK * 2;                // this CTA's half of a W tile (BN/2 rows x 64)
  static constexpr int BIAS_BYTES = 8192;                          // the bias vector of every column tile (N <= 2048)
  static constexpr int RING = 232448 - 1024 - 512 - V2_EPI_BYTES - A_BYTES - BIAS_BYTES;
  static constexpr int STAGES = (RING / B_STAGE) > 8 ? 8 : (RING / B_STAGE);
  static constexpr int TMEM_STRIDE = (BN <= 128) ? 128 : 256;
  static constexpr int TMEM_COLS = 2 * TMEM_STRIDE;
  static constexpr int SMEM_BYTES = A_BYTES + STAGES * B_STAGE + V2_EPI_BYTES + BIAS_BYTES + 1024 + 512;
  static_assert(STAGES >= 3, "W ring too shallow");
};

template <int BN, int EPI, int EW>
__global__ void __launch_bounds__(128 + EW * 32, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO,
               const __grid_constant__ CUtensorMap tmX, const GemmParams p, const LnPrologue ln) {
  using Cfg = LnCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::A_BYTES;
  uint8_t* epi_smem = sB + STAGES * Cfg::B_STAGE;
  float* s_bias = reinterpret_cast<float*>(epi_smem + V2_EPI_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + V2_EPI_BYTES + Cfg::BIAS_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* x_bar = tempty_bar + 2;                      // one per epilogue warp
  uint64_t* a_bar = x_bar + V2_MAX_EPI_WARPS;            // both CTAs' panels are normalised and staged
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = v2_cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmO);
    if (p.out_bf16_pre != nullptr) tma_prefetch_desc(&tmX);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * EW);
    }
    for (int s = 0; s < EW; ++s) mbar_init(&x_bar[s], 1);
    mbar_init(a_bar, 2 * EW);
    fence_barrier_init();
  }
  if (warp == 2) v2_tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  v2_cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  const int m0 = (int)(blockIdx.x >> 1) * 256 + (int)rank * 128;
  const int kblocks = p.K / BK;
#define LNG_STAMP(slot) do { if (kGemmProbes && p.dbg_buf && blockIdx.x == 0) p.dbg_buf[slot] = clock64(); } while (0)
  if (threadIdx.x == 0) LNG_STAMP(0);

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int nt = 0; nt < p.n_tiles; ++nt) {
        const int n0 = nt * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::B_STAGE);
          v2_tma_load_2d_pair(sB + stage * Cfg::B_STAGE, &tmB, &full_bar[stage], kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      mbar_wait(a_bar, 0);
      tc_fence_after();
      LNG_STAMP(2);
      const uint32_t a_addr = smem_u32(sA);
      for (int nt = 0; nt < p.n_tiles; ++nt) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        if (nt < 8) LNG_STAMP(8 + nt);
        const uint32_t d_tmem = tmem_base + as * Cfg::TMEM_STRIDE;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t b_addr = smem_u32(sB + stage * Cfg::B_STAGE);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + kb * A_TILE_BYTES + k * 32, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            v2_mma_pair(d_tmem, da, db, p.idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          v2_commit_pair(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        v2_commit_pair(&tfull_bar[as]);
        if (nt < 8) LNG_STAMP(16 + nt);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int ew = warp - 4;
    const int quad = warp & 3;
    const int half = ew >> 2;
    // ---- LayerNorm prologue: rows ew, ew + EW, ... of this CTA's panel (layernorm_fwd_kernel's arithmetic, row in registers)
    {
      constexpr int NV = (LNG_KB_MAX * BK) / 128;   // float4 per lane: 3 at K = 384
      constexpr int RB = 2;                         // rows per batch; the next batch's loads are in flight while this one is
      constexpr int NB = 128 / (EW * RB);           // normalised and stored (the prologue is load latency otherwise)
      const float inv_d = 1.0f / (float)p.K;
      const uint32_t a_base = smem_u32(sA);
      // the bias of every column tile -> shared memory (each epilogue unit would pay an L2 round trip for it otherwise)
      if (p.bias != nullptr) {
        for (int i = (ew * 32 + lane) * 4; i < p.N; i += EW * 32 * 4)
          *reinterpret_cast<float4*>(s_bias + i) = __ldg(reinterpret_cast<const float4*>(p.bias + i));
      }
      // gamma / beta of this lane's columns and the column part of its shared-memory addresses are row invariant: the
      // prologue is instruction bound (128 x K elements per CTA at ~8 issue slots each: the MMA cannot start before
      // the panel is complete), so everything that can leave the row loop does
      float4 g4[NV], b4[NV];
      uint32_t col_addr[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int c4 = i * 32 + lane;
        const int c = c4 * 4;
        if (c < p.K) {
          g4[i] = __ldg(reinterpret_cast<const float4*>(ln.w) + c4);
          b4[i] = __ldg(reinterpret_cast<const float4*>(ln.b) + c4);
        } else {
          g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          b4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        // element (r, c) of the K-major panel: k-block c / 64, 128-byte row r, 16-byte chunk swizzled by r & 7 -- and
        // r & 7 == ew & 7 for every row of this warp when EW is a multiple of 8 (rows ew, ew + EW, ...)
        col_addr[i] = a_base + (c >> 6) * A_TILE_BYTES + ((uint32_t)c & 7u) * 2;
      }
      float4 v[2][RB][NV];
      auto load_batch = [&](int bi, float4 (&dst)[RB][NV]) {
#pragma unroll
        for (int j = 0; j < RB; ++j) {
          const int row = m0 + ew + (bi * RB + j) * EW;
          const bool live = row < p.M;
          const float4* xr = reinterpret_cast<const float4*>(ln.x + (long long)row * ln.ldx);
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int c4 = i * 32 + lane;
            dst[j][i] = (live && c4 * 4 < p.K) ? xr[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      };
      load_batch(0, v[0]);
#pragma unroll
      for (int bi = 0; bi < NB; ++bi) {
        if (bi + 1 < NB) load_batch(bi + 1, v[(bi + 1) & 1]);
#pragma unroll
        for (int j = 0; j < RB; ++j) {
          const float4 (&vr)[NV] = v[bi & 1][j];
          const int r = ew + (bi * RB + j) * EW;
          const int row = m0 + r;
          const bool live = row < p.M;
          float sm = 0.f;
#pragma unroll
          for (int i = 0; i < NV; ++i) sm += (vr[i].x + vr[i].y) + (vr[i].z + vr[i].w);
          const float mean = warp_sum(sm) * inv_d;
          float q = 0.f;
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int c4 = i * 32 + lane;
            if (c4 * 4 < p.K) {
              const float a = vr[i].x - mean, bq = vr[i].y - mean, c = vr[i].z - mean, d = vr[i].w - mean;
              q += (a * a + bq * bq) + (c * c + d * d);
            }
          }
          const float rstd = rsqrtf(warp_sum(q) * inv_d + ln.eps);
          if (lane == 0 && live) {
            if (ln.mean) ln.mean[row] = mean;
            if (ln.rstd) ln.rstd[row] = rstd;
          }
          const uint32_t row_addr = (uint32_t)r * 128u;
#pragma unroll
          for (int i = 0; i < NV; ++i) {
            const int c = (i * 32 + lane) * 4;
            if (c < p.K) {
              uint32_t u0 = 0u, u1 = 0u;
              if (live) {
                // (layernorm_fwd_kernel's expression, operation for operation: a batch whose row count sends it to the
                // two-kernel path must produce the same bits -- test_cfg2_full_size_properties compares B = 64 with B = 8)
                u0 = pack_bf16((vr[i].x - mean) * rstd * g4[i].x + b4[i].x, (vr[i].y - mean) * rstd * g4[i].y + b4[i].y);
                u1 = pack_bf16((vr[i].z - mean) * rstd * g4[i].z + b4[i].z, (vr[i].w - mean) * rstd * g4[i].w + b4[i].w);
              }
              const uint32_t sw = EW % 8 == 0 ? (uint32_t)ew & 7u : (uint32_t)r & 7u;
              const uint32_t addr = col_addr[i] + row_addr + (((((uint32_t)c & 63u) >> 3) ^ sw) << 4);
              asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(u0), "r"(u1) : "memory");
            }
          }
        }
      }
      fence_proxy_async();   // generic-proxy writes -> visible to the tensor-core (async proxy) reads, of either CTA
      __syncwarp();
      if (lane == 0) v2_mbar_arrive_cta(a_bar, 0);
      if (ew == 0 && lane == 0) LNG_STAMP(1);
      asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");   // the bias row is staged by all epilogue warps
    }
    const uint32_t bias_row = p.bias != nullptr ? smem_u32(s_bias) : 0u;
    // ---- epilogue over the column tiles
    const uint32_t stage_smem = smem_u32(epi_smem + ew * (V2_EPI_BYTES / EW));
    uint64_t* xbar = &x_bar[ew];
    uint32_t xphase = 0;
    uint32_t out_toggle = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int nt = 0; nt < p.n_tiles; ++nt) {
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * Cfg::TMEM_STRIDE;
      if (ew == 0 && lane == 0 && nt < 8) LNG_STAMP(24 + nt);
      v2_epilogue_tile<BN, EPI, EW>(p, &tmO, &tmX, taddr, m0 + quad * 32, nt * BN, stage_smem, xbar, xphase, half, lane, true,
                                    false, false, &tempty_bar[as], true, &out_toggle, 0u, bias_row);
      if (ew == 0 && lane == 0 && nt < 8) LNG_STAMP(32 + nt);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
    if (lane == 0) v2_bulk_wait_all();
    __syncwarp();
    if (ew == 0 && lane == 0) LNG_STAMP(3);
  }

  tc_fence_before();
  v2_cluster_sync();
  if (warp == 2) v2_tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
  if (threadIdx.x == 0) LNG_STAMP(4);
}

// ------------------------------------------------------------------------------------------------ host side
template <int BN, bool PAIR, bool MN, int EPI, int EW>
static int v2_launch(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tx,
                     const GemmParams& p, int use_x, int reduce_out, cudaStream_t st) {
  using Cfg = V2Cfg<BN, PAIR>;
  auto kern = gemm_v2_kernel<BN, PAIR, MN, EPI, EW>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  int workers = PAIR ? sm_count() / 2 : sm_count();
  if (workers > p.total_tiles) workers = p.total_tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(PAIR ? workers * 2 : workers);
  cfg.blockDim = dim3(128 + EW * 32);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = PAIR ? 2 : 1;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attrs;
  cfg.numAttrs = 2;
  const int prof = prof_begin(st);
  if (kGemmProbes && (p.dbg & 32)) {   // wait-cycle counters of CTA 0 (eager launches only: synchronises and prints)
    static long long* buf = nullptr;
    if (!buf) cudaMalloc(&buf, 16 * sizeof(long long));
    cudaMemsetAsync(buf, 0, 16 * sizeof(long long), st);
    GemmParams q = p;
    q.dbg_buf = buf;
    B200_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, tx, q, use_x, reduce_out));
    cudaStreamSynchronize(st);
    long long h[16];
    cudaMemcpy(h, buf, sizeof h, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[gemm_v2 dbg] BN=%d pair=%d epi=%d EW=%d tiles/cta=%lld | producer: total %lld wait_empty %lld | mma: total %lld "
            "wait_full %lld wait_tempty %lld | epi w0: loop %lld wait_tfull %lld total %lld | epi wlast: loop %lld wait_tfull %lld total %lld\n",
            BN, (int)PAIR, EPI, EW, h[5], h[0], h[1], h[2], h[3], h[4], h[6], h[7], h[8], h[9], h[10], h[11]);
    B200_LAUNCH_OK();
    return 0;
  }
  B200_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, tx, p, use_x, reduce_out));
  {
    // algorithmic traffic: both operands once, the result once (a read-modify-write result twice), operand tiles of the
    // epilogue (fp32 residual / 16-bit aux) once, the second 16-bit copy once
    const double mn = (double)p.M * p.N;
    double bytes = 2.0 * ((double)p.M * p.K + (double)p.N * p.K) * p.algo_scale;
    if (EPI == 2) bytes += mn * 4.0 * ((reduce_out || p.split_k > 1) ? 2.0 : 1.0) + (use_x ? mn * 4.0 : 0.0);
    else bytes += mn * 2.0 * (p.out_bf16_pre != nullptr ? 2.0 : 1.0) + (use_x ? mn * 2.0 : 0.0);
    prof_end(prof, st, 2.0 * p.M * p.N * p.K * p.algo_scale, 0, bytes);
  }
  B200_LAUNCH_OK();
  return 0;
}

template <int BN, bool PAIR, bool MN, int EW>
static int v2_dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                           const CUtensorMap& tx, const GemmParams& p, int use_x, int reduce_out, cudaStream_t st) {
  if constexpr (MN) {
    return v2_launch<BN, PAIR, MN, 2, EW>(ta, tb, to, tx, p, use_x, reduce_out, st);
  } else {
    if (epi == 0) return v2_launch<BN, PAIR, MN, 0, EW>(ta, tb, to, tx, p, use_x, reduce_out, st);
    if (epi == 1) return v2_launch<BN, PAIR, MN, 1, EW>(ta, tb, to, tx, p, use_x, reduce_out, st);
    return v2_launch<BN, PAIR, MN, 2, EW>(ta, tb, to, tx, p, use_x, reduce_out, st);
  }
}

template <bool PAIR, bool MN, int EW>
static int v2_dispatch_bn(int bn, int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                          const CUtensorMap& tx, const GemmParams& p, int use_x, int reduce_out, cudaStream_t st) {
  if (bn == 256) return v2_dispatch_epi<256, PAIR, MN, EW>(epi, ta, tb, to, tx, p, use_x, reduce_out, st);
  if (bn == 192) return v2_dispatch_epi<192, PAIR, MN, EW>(epi, ta, tb, to, tx, p, use_x, reduce_out, st);
  return v2_dispatch_epi<128, PAIR, MN, EW>(epi, ta, tb, to, tx, p, use_x, reduce_out, st);
}

// tile width: fewest (waves x tile cost); wide tiles amortise the A traffic, narrow ones the wave quantisation
static int v2_pick_bn(int N, long long row_tiles, int workers) {
  const int forced = option(OPT_GEMM_BN);
  if (forced == 128 || forced == 192 || forced == 256) return forced;
  const int cands[3] = {256, 192, 128};
  int best = 128;
  double best_cost = 1e30;
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const long long tiles = cdiv(N, bn) * row_tiles;
    const double cost = double(cdiv(tiles, workers)) * (bn + 24);   // +24: per-tile fixed cost (pipeline fill / drain)
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

// Returns 1 when the problem is outside what this kernel handles (the caller falls back to the first-generation
// kernels), 0 on success, negative on error.
int launch_gemm_v2(const b200_gemm_desc* d, GemmParams& p, cudaStream_t st) {
  if (!option(OPT_GEMM_V2)) return 1;
  const bool mn = d->a_mn_major && d->b_mn_major;
  if (d->a_mn_major != d->b_mn_major) return 1;
  const bool f32 = d->out_f32 != nullptr;
  if (d->out_row_period > 0 || d->res_row_period > 0) {
    // row-remapped output / periodic residual (patch embed -> token buffer with cls gaps, + pos_embed): 3-D output map,
    // when the periods are multiples of the 32-row store unit and nothing else is going on
    const bool ok = f32 && !d->a_mn_major && !d->b_mn_major && d->split_k <= 1 && !d->atomic_add && d->out_batch_period <= 0 &&
                    (d->out_row_period <= 0 || (d->out_row_period % 32 == 0 && d->M % d->out_row_period == 0)) &&
                    (d->res_row_period <= 0 || (d->res_row_period % 32 == 0 && d->residual != nullptr)) &&
                    (d->out_row_period <= 0 || d->residual == nullptr || d->res_row_period > 0) &&
                    (const void*)d->residual != (const void*)d->out_f32;
    if (!ok) return 1;
  }
  if (d->out_batch_period > 0) {
    B200_CHECK_ARG(f32 && !d->residual && d->out_batch_period % 32 == 0 && d->N % d->out_batch_period == 0 &&
                       d->N % 32 == 0 && d->out_batch_stride % 4 == 0 && d->ldo32 % 4 == 0 && !d->a_mn_major && !d->b_mn_major,
                   "out_batch_period needs a plain fp32 output with 32-aligned batches");
  }
  if (f32 && d->out_bf16) return 1;
  if (f32 && (d->out_bf16_pre || d->aux_mode != 0 || d->act != 0)) return 1;
  if (!f32 && (d->col_scale || d->residual || d->atomic_add)) return 1;
  if (!f32 && d->out_bf16_pre && d->aux) return 1;
  if (!f32 && d->act == B200_ACT_GELU && d->aux) return 1;
  if (mn && !f32) return 1;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (f32 && (!al16(d->out_f32) || d->ldo32 % 4 != 0)) return 1;
  if (!f32 && (!al16(d->out_bf16) || d->ldo16 % 8 != 0)) return 1;
  if (d->residual && (!al16(d->residual) || d->ldres % 4 != 0)) return 1;
  if (d->aux && d->aux_mode && (!al16(d->aux) || d->ldaux % 8 != 0)) return 1;
  if (d->out_bf16_pre && (!al16(d->out_bf16_pre) || d->ldo16_pre % 8 != 0)) return 1;
  if (d->bias && !al16(d->bias)) return 1;
  if (d->col_scale && !al16(d->col_scale)) return 1;
  if (d->N % 32 != 0) return 1;   // ragged N: first-generation kernel (per-column tail handling)
  p.pre_alt = (d->out16_pre_alt && d->out_bf16_pre) ? 1 : 0;
  p.aux_deep = option(OPT_GEMM_AUX_DEEP);
  p.colsum = (!f32 && p.split_k == 1) ? d->out16_colsum : nullptr;
  if (d->out16_colsum != nullptr && p.colsum == nullptr) return 1;

  const int pair_mode = option(OPT_GEMM_2CTA);   // -1 auto, 0 never, 1 whenever possible
  bool pair = false;
  if (!mn && p.split_k == 1 && d->M >= 1024 && d->N >= 128) {
    // the CTA pair halves the B traffic per SM; it pays once the mainloop is long enough to be operand bound
    pair = pair_mode == 1 || (pair_mode < 0 && d->K >= 1024);
  }
  const int tile_m = pair ? 256 : 128;
  const int workers = pair ? sm_count() / 2 : sm_count();
  p.m_tiles = (int)cdiv(d->M, tile_m);
  const int bn = v2_pick_bn(d->N, (long long)p.m_tiles * p.split_k, workers);
  p.n_tiles = (int)cdiv(d->N, bn);
  p.total_tiles = p.m_tiles * p.n_tiles * p.split_k;
  p.idesc = make_idesc_bf16(tile_m, bn, mn, mn);
  if (d->a_is_fp16) p.idesc &= ~(7u << 7);
  if (d->b_is_fp16) p.idesc &= ~(7u << 10);

  const bool inplace = f32 && d->residual && (const void*)d->residual == (const void*)d->out_f32 && d->ldres == d->ldo32;
  const int inplace_red = option(OPT_GEMM_INPLACE_RED);
  int reduce_out = (f32 && d->atomic_add) ? 1 : 0;
  int use_x = 0;
  if (f32 && d->residual) {
    if (inplace && inplace_red && !d->atomic_add) reduce_out = 1;
    else use_x = 1;
  }
  if (!f32 && d->aux && d->aux_mode) use_x = 1;
  if (use_x && p.split_k > 1) return 1;

  CUtensorMap ta, tb, to, tx;
  const int bsw = pair ? bn / 2 : bn;
  if (!mn) {
    B200_TRY(make_tensor_map_2d(&ta, d->A, (uint64_t)d->K, (uint64_t)d->M, (uint64_t)d->lda, BK, BM));
    B200_TRY(make_tensor_map_2d(&tb, d->B, (uint64_t)d->K, (uint64_t)d->N, (uint64_t)d->ldb, BK, (uint32_t)bsw));
  } else {
    B200_TRY(make_tensor_map_2d(&ta, d->A, (uint64_t)d->M, (uint64_t)d->K, (uint64_t)d->lda, 64, BK));
    B200_TRY(make_tensor_map_2d(&tb, d->B, (uint64_t)d->N, (uint64_t)d->K, (uint64_t)d->ldb, 64, BK));
  }
  p.out_bp = d->out_batch_period > 0 ? d->out_batch_period : 0;
  if (f32 && p.out_bp > 0) {
    B200_TRY(make_tensor_map_3d(&to, d->out_f32, 4, (uint64_t)p.out_bp, (uint64_t)d->M, (uint64_t)(d->N / p.out_bp),
                                (uint64_t)d->ldo32, (uint64_t)d->out_batch_stride, 32, 32, 1, 128));
    tx = to;
  } else if (f32) {
    if (d->out_row_period > 0) {
      const long long rp = d->out_row_period, pad = d->out_row_pad;
      B200_TRY(make_tensor_map_3d(&to, d->out_f32 + pad * d->ldo32, 4, (uint64_t)d->N, (uint64_t)rp, (uint64_t)(d->M / rp),
                                  (uint64_t)d->ldo32, (uint64_t)((rp + pad) * d->ldo32), 32, 32, 1, 128));
    } else {
      B200_TRY(make_tensor_map_ex(&to, d->out_f32, 4, (uint64_t)d->N, (uint64_t)d->M, (uint64_t)d->ldo32, 32, 32, 128));
    }
    const uint64_t res_rows = d->res_row_period > 0 ? (uint64_t)d->res_row_period : (uint64_t)d->M;
    if (use_x) B200_TRY(make_tensor_map_ex(&tx, d->residual, 4, (uint64_t)d->N, res_rows, (uint64_t)d->ldres, 32, 32, 128));
    else tx = to;
  } else {
    B200_TRY(make_tensor_map_ex(&to, d->out_bf16, 2, (uint64_t)d->N, (uint64_t)d->M, (uint64_t)d->ldo16, 32, 32, 64));
    if (use_x) B200_TRY(make_tensor_map_ex(&tx, d->aux, 2, (uint64_t)d->N, (uint64_t)d->M, (uint64_t)d->ldaux, 32, 32, 64));
    else if (d->out_bf16_pre) B200_TRY(make_tensor_map_ex(&tx, d->out_bf16_pre, 2, (uint64_t)d->N, (uint64_t)d->M, (uint64_t)d->ldo16_pre, 32, 32, 64));
    else tx = to;
  }
  const int epi = f32 ? 2 : (d->act == B200_ACT_GELU ? 1 : 0);
  // 16 epilogue warps when the epilogue stages no operand tile (4 KB of staging per warp is then enough)
  const int ew_mode = option(OPT_GEMM_EW);
  const bool ew16 = ew_mode == 16 && !use_x && !pair;
  if (mn) {
    if (ew16) return v2_dispatch_bn<false, true, 16>(bn, epi, ta, tb, to, tx, p, use_x, reduce_out, st);
    return v2_dispatch_bn<false, true, 8>(bn, epi, ta, tb, to, tx, p, use_x, reduce_out, st);
  }
  if (pair) return v2_dispatch_bn<true, false, 8>(bn, epi, ta, tb, to, tx, p, use_x, reduce_out, st);
  if (ew16) return v2_dispatch_bn<false, false, 16>(bn, epi, ta, tb, to, tx, p, use_x, reduce_out, st);
  return v2_dispatch_bn<false, false, 8>(bn, epi, ta, tb, to, tx, p, use_x, reduce_out, st);
}

template <int BN, int EPI, int EW>
static int lng_launch(const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tx, const GemmParams& p,
                      const LnPrologue& ln, cudaStream_t st) {
  using Cfg = LnCfg<BN>;
  auto kern = gemm_ln_kernel<BN, EPI, EW>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * (unsigned)cdiv(p.M, 256));
  cfg.blockDim = dim3(128 + EW * 32);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = 2;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attrs;
  cfg.numAttrs = 2;
  const int prof = prof_begin(st);
  if (kGemmProbes && option(OPT_GEMM_DBG)) {   // clock stamps of CTA 0 (eager launches only: synchronises and prints)
    static long long* buf = nullptr;
    if (!buf) cudaMalloc(&buf, 48 * sizeof(long long));
    cudaMemsetAsync(buf, 0, 48 * sizeof(long long), st);
    GemmParams q = p;
    q.dbg_buf = buf;
    B200_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tb, to, tx, q, ln));
    cudaStreamSynchronize(st);
    long long h[48];
    cudaMemcpy(h, buf, sizeof h, cudaMemcpyDeviceToHost);
    fprintf(stderr, "[gemm_ln dbg] BN=%d epi=%d: ln_done %lld a_bar %lld epi_done %lld exit %lld |", BN, EPI, h[1] - h[0], h[2] - h[0],
            h[3] - h[0], h[4] - h[0]);
    for (int t = 0; t < p.n_tiles && t < 8; ++t)
      fprintf(stderr, " tile%d mma %lld-%lld epi %lld-%lld |", t, h[8 + t] - h[0], h[16 + t] - h[0], h[24 + t] - h[0], h[32 + t] - h[0]);
    fprintf(stderr, "\n");
    B200_LAUNCH_OK();
    return 0;
  }
  B200_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, tb, to, tx, p, ln));
  prof_end(prof, st, 2.0 * p.M * p.N * p.K, 0,
           4.0 * p.M * p.K + 2.0 * p.N * p.K + 2.0 * p.M * p.N * (p.out_bf16_pre != nullptr ? 2.0 : 1.0));
  B200_LAUNCH_OK();
  return 0;
}

template <int BN>
static int lng_dispatch(int epi, bool has_pre, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& tx,
                        const GemmParams& p, const LnPrologue& ln, cudaStream_t st) {
  (void)has_pre;
  if (epi == 1) return lng_launch<BN, 1, 16>(tb, to, tx, p, ln, st);
  return lng_launch<BN, 0, 16>(tb, to, tx, p, ln, st);
}

// Returns 1 when the problem is outside the kernel's envelope (the caller runs layernorm_fwd + the plain GEMM).
int launch_gemm_ln(const b200_gemm_desc* d, const float* x, long long ldx, const float* ln_w, const float* ln_b, float eps,
                   float* mean, float* rstd, cudaStream_t st) {
  if (!option(OPT_GEMM_LN)) return 1;
  if (d->K % BK != 0 || d->K > LNG_KB_MAX * BK || d->K % 128 != 0 || d->N % 32 != 0 || d->N > 2048 || d->M < 256) return 1;
  // a CTA owns a 128-row panel and walks every column tile of it: with fewer panels than half the SMs (cfg1: B = 2, five
  // panels) the column tiles serialise on a handful of SMs (measured 25 us against 12 us for layernorm + GEMM)
  if (cdiv(d->M, 128) * 2 < sm_count() && option(OPT_GEMM_LN) < 2) return 1;   // (gemm_ln = 2: any M, for the tests)
  if (d->a_mn_major || d->b_mn_major || d->a_is_fp16 || d->b_is_fp16 || d->split_k > 1 || d->atomic_add) return 1;
  if (d->out_f32 || d->residual || d->col_scale || d->aux || d->out16_colsum || d->out16_pre_alt || d->out16_is_fp16) return 1;
  if (d->out_row_period > 0 || d->res_row_period > 0 || d->out_batch_period > 0) return 1;
  if (d->out_bf16 == nullptr || (d->act != B200_ACT_NONE && d->act != B200_ACT_GELU)) return 1;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al16(x) || ldx % 4 != 0 || !al16(ln_w) || !al16(ln_b) || !al16(d->B) || d->ldb % 8 != 0) return 1;
  if (!al16(d->out_bf16) || d->ldo16 % 8 != 0 || (d->bias && !al16(d->bias))) return 1;
  if (d->out_bf16_pre && (!al16(d->out_bf16_pre) || d->ldo16_pre % 8 != 0)) return 1;
  // column tile: the kernel is paced by its epilogue (32-column units, EW / 4 warps per TMEM quadrant): fewest unit rounds
  // per warp first (256 columns = two full rounds per tile), least padded MMA work on ties
  int bn = 256;
  long long best = -1;
  for (int c : {256, 192, 128}) {
    const long long tiles = cdiv(d->N, c), rounds_per_tile = cdiv(c / 32, 4);
    long long rounds = (tiles - 1) * rounds_per_tile + cdiv(cdiv(d->N - (tiles - 1) * c, 32), 4);
    const long long cost = rounds * 1000 + tiles * c;
    if (best < 0 || cost < best) { best = cost; bn = c; }
  }
  GemmParams p{};
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.m_tiles = (int)cdiv(d->M, 256); p.n_tiles = (int)cdiv(d->N, bn); p.num_kb = d->K / BK; p.kb_per_split = p.num_kb; p.split_k = 1;
  p.total_tiles = p.m_tiles * p.n_tiles;
  p.bias = d->bias; p.act = d->act;
  p.out_bf16 = static_cast<__nv_bfloat16*>(d->out_bf16); p.ldo16 = d->ldo16;
  p.out_bf16_pre = static_cast<__nv_bfloat16*>(d->out_bf16_pre); p.ldo16_pre = d->ldo16_pre;
  p.vec_ok = 1;
  p.algo_scale = 1.0f;
  p.idesc = make_idesc_bf16(256, bn, false, false);
  p.dbg = kGemmProbes ? option(OPT_GEMM_DBG) : 0;
  LnPrologue ln{x, ldx, ln_w, ln_b, eps, mean, rstd};
  CUtensorMap tb, to, tx;
  B200_TRY(make_tensor_map_2d(&tb, d->B, (uint64_t)d->K, (uint64_t)d->N, (uint64_t)d->ldb, BK, (uint32_t)(bn / 2)));
  B200_TRY(make_tensor_map_ex(&to, d->out_bf16, 2, (uint64_t)d->N, (uint64_t)d->M, (uint64_t)d->ldo16, 32, 32, 64));
  if (d->out_bf16_pre) B200_TRY(make_tensor_map_ex(&tx, d->out_bf16_pre, 2, (uint64_t)d->N, (uint64_t)d->M, (uint64_t)d->ldo16_pre, 32, 32, 64));
  else tx = to;
  const int epi = d->act == B200_ACT_GELU ? 1 : 0;
  if (bn == 256) return lng_dispatch<256>(epi, d->out_bf16_pre != nullptr, tb, to, tx, p, ln, st);
  if (bn == 192) return lng_dispatch<192>(epi, d->out_bf16_pre != nullptr, tb, to, tx, p, ln, st);
  return lng_dispatch<128>(epi, d->out_bf16_pre != nullptr, tb, to, tx, p, ln, st);
}

}  // namespace b200
