// tcgen05 attention forward for head_dim 64: the 224-pixel teacher (N = 257, whole key range resident) and every
// 518-pixel sequence (N = 1 369 / 1 370, key blocks of 256 with online softmax) -- hub Attention.forward, reached through
// models/backbones/dinov2.py:32 and train/distillation_module.py:177 (bf16), and the ScaleKD cross-attention at 518
// pixels, losses/scalekd.py:299-314 (template FP16: fp16 operands, a bf16 copy of the output, batch-invariant query).
// Key ranges of 128-256 tokens go to attention_pp.cu (two tiles in flight) first.
//
// With the whole key range resident, a 128-query tile needs ONE QK^T and ONE PV: no online-softmax rescaling.
//   S[128, Nk] = Q K^T        tcgen05.mma  SS  (Q: smem K-major, K: smem K-major)   -> TMEM columns [0, 272)
//   P = exp2(scale*S - max)   8 softmax warps (two per TMEM lane quadrant, interleaved 32-column chunks), thread = row:
//                             the warp's whole share of the S row (<= 160 values) is pulled into registers with ONE
//                             round of tcgen05.ld -- S is then free, so QK^T of the next tile runs under this tile's
//                             exponentials -- partial row maxima meet in shared memory, bf16 pairs are written back
//                             with tcgen05.st                                       -> TMEM columns [288, 424)
//   O[128, 64] = P V          tcgen05.mma  TS  (P: TMEM, V: smem MN-major)          -> TMEM columns [448, 512)
//   O / rowsum -> bf16 -> swizzled smem -> one bulk tensor store per warp.
// Persistent CTAs walk (batch, head) units; K/V are loaded once per unit by TMA (3-D maps over the fused qkv tensor,
// rows past the sequence end zero-filled) and every 128-row query tile of the unit re-uses them. S, P and O live in
// disjoint TMEM columns, so QK^T of the next tile runs under the epilogue of the current one.
// A query count that leaves a short tail (Nq mod 128 <= 8 -- the teacher's 257 = 2*128 + 1) would waste a whole MMA
// tile on it; those rows are computed by an otherwise idle warp with mma.sync from the same shared-memory K/V.
//
//   warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM alloc   warp 3: tail rows   warps 4-11: softmax + epilogue
// Registers are re-balanced with setmaxnreg: 104 for warps 0-3, 200 for the softmax warps (the S row lives there).
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/b200_distill.h"

#include <stdlib.h>

namespace b200 {

int make_tensor_map_3d(CUtensorMap* out, const void* ptr, int esize, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1,
                       uint64_t ld2, uint32_t b0, uint32_t b1, uint32_t b2, int swizzle);

// per-phase clock stamps of CTA 0 are compiled in only with -DB200_ATTN_PROBES
#ifdef B200_ATTN_PROBES
constexpr bool kAttnProbes = true;
#else
constexpr bool kAttnProbes = false;
#endif
constexpr int ATC_THREADS = 384;
constexpr int ATC_SM_WARPS = 8;
constexpr int ATC_NKP_MAX = 272;                 // keys, padded to a multiple of 16
constexpr int ATC_KV_ROWS = 272;
constexpr int ATC_KV_BOX = 136;                  // two TMA boxes of 136 rows
constexpr int ATC_Q_BYTES = 128 * 128;           // 128 rows x 64 bf16
constexpr int ATC_KV_BYTES = ATC_KV_ROWS * 128;  // 34816
constexpr int ATC_TAIL_MAX = 8;                  // query rows left to the CUDA-core warp
constexpr uint32_t ATC_TMEM_S = 0, ATC_TMEM_P = 288, ATC_TMEM_O = 448;
// shared memory map (bytes from the 1024-aligned base)
constexpr int ATC_OFF_Q = 0;                                   // [2][16384]
constexpr int ATC_OFF_K = ATC_OFF_Q + 2 * ATC_Q_BYTES;         // [2][34816]
constexpr int ATC_OFF_V = ATC_OFF_K + 2 * ATC_KV_BYTES;        // [2][34816]
constexpr int ATC_OFF_O = ATC_OFF_V + 2 * ATC_KV_BYTES;        // [8][2048] output staging, one 32x32 tile per softmax warp
constexpr int ATC_OFF_PB = ATC_OFF_O + 8 * 2048;               // tail-row scratch
constexpr int ATC_OFF_X = ATC_OFF_PB + 13312;                  // (tail: [8][272] fp32 scores, [8][272] bf16 P, [8] 1/sum)                   // [2][2][128] fp32: partial row max / row sum per column half
constexpr int ATC_OFF_BAR = ATC_OFF_X + 2048;
constexpr int ATC_SMEM_BYTES = ATC_OFF_BAR + 256 + 1024;

struct AttnTcParams {
  int B, heads, Nq, Nk, nkp;
  int n_qt;                  // 128-row query tiles per (b, h) on the tensor path
  int tail0, tail_n;         // first tail row / number of tail rows (CUDA-core path)
  float scale_log2;
  float* lse;
  const __nv_bfloat16* q; long long q_bs, q_ts;
  __nv_bfloat16* o; long long o_bs, o_ts;
  uint32_t idesc_qk1, idesc_qk2, idesc_pv;
  int n1, n2;                // QK^T column split of the LAST key block (n1 <= 256, n2 = nkp - n1)
  int nblk;                  // key blocks per query tile: 1 (whole key range resident, Nk <= 272) or ceil(Nk / 256)
  int nk_last, nkp_last;     // keys / padded keys of the last block (all other blocks hold 256)
  uint32_t idesc_qk_full;    // QK^T of a full 256-key block
  long long* dbg;            // B200_ATTN_DBG=1: per-phase clock64 stamps of CTA 0 ([item][16])
  int fp16;                  // q / k / v / o are fp16 (ScaleKD projector forward precision); P is packed the same way
  int has_alt;               // a second copy of o in the other 16-bit format goes out through tmOalt
  int q_batched;             // 0: q is batch invariant (the self-query embedding)
};

__device__ __forceinline__ void atc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void atc_tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void atc_tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void atc_tma_store_3d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
#define ATC_STAMP(slot)                                                                        \
  do {                                                                                          \
    if (kAttnProbes && p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && item < 8) p.dbg[item * 16 + (slot)] = clock64(); \
  } while (0)

__device__ __forceinline__ void atc_ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void atc_ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(smem_u32(p)));
}
__device__ __forceinline__ void atc_mma_sync(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float atc_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// MULTI = false: the whole key range is resident (one block of up to 272 keys per query tile, 5 S chunks per softmax warp,
// the output is read once from TMEM). MULTI = true: key blocks of 256 with online softmax (4 chunks per warp plus the
// running 32-column output accumulator in registers).
// FP16: q / k / v / o (and P) are fp16 and a bf16 copy of o leaves through tmOalt -- the ScaleKD projector's forward operands
// (DESIGN.md section 3); otherwise bf16 in, bf16 out (the teacher).
template <bool MULTI, bool FP16>
__global__ void __launch_bounds__(ATC_THREADS, 1)
attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO,
                   const __grid_constant__ CUtensorMap tmOalt, const AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATC_OFF_BAR);
  uint64_t* q_full = bars;          // [2]
  uint64_t* q_empty = bars + 2;     // [2]
  uint64_t* kv_full = bars + 4;     // [2]
  uint64_t* kv_empty = bars + 6;    // [2]
  uint64_t* s_full = bars + 8;      // [1]
  uint64_t* o_full = bars + 9;      // [1]
  uint64_t* s_free = bars + 10;     // [1] S columns drained into registers
  uint64_t* p_full = bars + 11;     // [2] P published in two groups: chunks 0..3 (each warp's first two), then the rest
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
    if (FP16) tma_prefetch_desc(&tmOalt);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], p.tail_n > 0 ? 2 : 1);
    }
    mbar_init(s_full, 1);
    for (int i = 0; i < 2; ++i) mbar_init(&p_full[i], ATC_SM_WARPS);
    mbar_init(o_full, 1);
    mbar_init(s_free, ATC_SM_WARPS);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int n_units = p.B * p.heads;
  pdl_trigger();   // PDL: the prologue above overlapped the previous kernel's tail
  pdl_wait();

  if (warp < 4) {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    // K/V tiles: one load per (batch, head) unit when the whole key range fits (nblk == 1: every query tile of the unit
    // re-uses it), else one per (query tile, key block) through the same two-deep ring.
    if (elect_one()) {
      int qc = 0, kvc = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        const int b = u / p.heads, h = u - b * p.heads;
        for (int qt = 0; qt < p.n_qt; ++qt, ++qc) {
          const int qb = qc & 1;
          mbar_wait(&q_empty[qb], ((qc >> 1) & 1) ^ 1);
          mbar_expect_tx(&q_full[qb], ATC_Q_BYTES);
          tma_load_3d(smem + ATC_OFF_Q + qb * ATC_Q_BYTES, &tmQ, &q_full[qb], h * 64, qt * 128, p.q_batched ? b : 0);
          for (int j = 0; j < p.nblk; ++j) {
            if (p.nblk == 1 && qt > 0) break;
            const int kb = kvc & 1;
            mbar_wait(&kv_empty[kb], ((kvc >> 1) & 1) ^ 1);
            mbar_expect_tx(&kv_full[kb], 2 * ATC_KV_BYTES);
            uint8_t* sK = smem + ATC_OFF_K + kb * ATC_KV_BYTES;
            uint8_t* sV = smem + ATC_OFF_V + kb * ATC_KV_BYTES;
            const int r0 = j * 256;
            tma_load_3d(sK, &tmK, &kv_full[kb], h * 64, r0, b);
            tma_load_3d(sK + ATC_KV_BOX * 128, &tmK, &kv_full[kb], h * 64, r0 + ATC_KV_BOX, b);
            tma_load_3d(sV, &tmV, &kv_full[kb], h * 64, r0, b);
            tma_load_3d(sV + ATC_KV_BOX * 128, &tmV, &kv_full[kb], h * 64, r0 + ATC_KV_BOX, b);
            ++kvc;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      int item = 0, qc = 0, kvc = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
        for (int qt = 0; qt < p.n_qt; ++qt, ++qc) {
          const int qb = qc & 1;
          const uint32_t q_addr = smem_u32(smem + ATC_OFF_Q + qb * ATC_Q_BYTES);
          mbar_wait(&q_full[qb], (qc >> 1) & 1);
          for (int j = 0; j < p.nblk; ++j, ++item) {
            const bool kv_new = p.nblk > 1 || qt == 0;
            if (kv_new) mbar_wait(&kv_full[kvc & 1], (kvc >> 1) & 1);
            const int kb = (kv_new ? kvc : kvc - 1) & 1;
            const uint32_t k_addr = smem_u32(smem + ATC_OFF_K + kb * ATC_KV_BYTES);
            const uint32_t v_addr = smem_u32(smem + ATC_OFF_V + kb * ATC_KV_BYTES);
            const bool last_blk = j == p.nblk - 1;
            const int nkp = last_blk ? p.nkp_last : 256;
            if (kAttnProbes && p.dbg != nullptr && blockIdx.x == 0 && item < 8) p.dbg[item * 16 + 8] = clock64();
            if (item > 0) mbar_wait(s_free, (item - 1) & 1);   // the previous block's S row is in registers
            tc_fence_after();
            if (kAttnProbes && p.dbg != nullptr && blockIdx.x == 0 && item < 8) p.dbg[item * 16 + 9] = clock64();
            // S = Q K_j^T
            const uint32_t id1 = last_blk ? p.idesc_qk1 : p.idesc_qk_full;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t da = make_smem_desc_sw128(q_addr + ks * 32, 16, 1024);
              const uint64_t db = make_smem_desc_sw128(k_addr + ks * 32, 16, 1024);
              tc_mma_bf16(tmem_base + ATC_TMEM_S, da, db, id1, ks > 0 ? 1u : 0u);
              if (last_blk && p.n2 > 0) {
                const uint64_t db2 = make_smem_desc_sw128(k_addr + 256 * 128 + ks * 32, 16, 1024);
                tc_mma_bf16(tmem_base + ATC_TMEM_S + 256, da, db2, p.idesc_qk2, ks > 0 ? 1u : 0u);
              }
            }
            if (last_blk) tc_commit(&q_empty[qb]);
            tc_commit(s_full);
            // O_j = P V_j in two groups (k-steps 0..7 = keys 0..127, then the rest) as the softmax warps publish P: the
            // first group runs under the remaining exponentials. (Finer groups starve: this warp shares its scheduler
            // with two softmax warps that are always eligible.)
            {
              const int ksteps = nkp >> 4;
              const int split = ksteps < 8 ? ksteps : 8;
              mbar_wait(&p_full[0], item & 1);
              tc_fence_after();
              for (int ks = 0; ks < split; ++ks) {
                const uint64_t db = make_smem_desc_sw128(v_addr + ks * 2048, 8192, 1024);
                atc_mma_ts(tmem_base + ATC_TMEM_O, tmem_base + ATC_TMEM_P + ks * 8, db, p.idesc_pv, ks > 0 ? 1u : 0u);
              }
              if (kAttnProbes && p.dbg != nullptr && blockIdx.x == 0 && item < 8) p.dbg[item * 16 + 10] = clock64();
              mbar_wait(&p_full[1], item & 1);
              tc_fence_after();
              for (int ks = split; ks < ksteps; ++ks) {
                const uint64_t db = make_smem_desc_sw128(v_addr + ks * 2048, 8192, 1024);
                atc_mma_ts(tmem_base + ATC_TMEM_O, tmem_base + ATC_TMEM_P + ks * 8, db, p.idesc_pv, 1u);
              }
              if (kAttnProbes && p.dbg != nullptr && blockIdx.x == 0 && item < 8) p.dbg[item * 16 + 11] = clock64();
            }
            tc_commit(o_full);
            if (p.nblk > 1 || qt == p.n_qt - 1) tc_commit(&kv_empty[kb]);
            if (kv_new) ++kvc;
          }
        }
      }
    }
  } else if (warp == 3) {
    // ------------------------------------------------------------------ tail rows (Nq mod 128 <= 8) on the legacy MMA path
    // One warp, mma.sync m16n8k16 straight from the TMA-swizzled K/V tiles (ldmatrix with the 128B-swizzle address
    // math): ~700 instructions per unit where a scalar version needed ~4500 and became the critical path.
    if (p.tail_n > 0) {
      float* ssc = reinterpret_cast<float*>(smem + ATC_OFF_PB);                                 // [8][272] scaled scores
      __nv_bfloat16* sp = reinterpret_cast<__nv_bfloat16*>(smem + ATC_OFF_PB + 8 * 272 * 4);     // [8][272] probabilities
      float* stat = reinterpret_cast<float*>(smem + ATC_OFF_PB + 8 * 272 * 6);                   // [8] 1/sum
      const int g = lane >> 2, t = lane & 3, mi = lane >> 3, rr = lane & 7;
      int uc = 0;
      for (int u = blockIdx.x; u < n_units; u += gridDim.x, ++uc) {
        const int b = u / p.heads, h = u - b * p.heads;
        const int kb = uc & 1;
        const uint8_t* sK = smem + ATC_OFF_K + kb * ATC_KV_BYTES;
        const uint8_t* sV = smem + ATC_OFF_V + kb * ATC_KV_BYTES;
        mbar_wait(&kv_full[kb], (uc >> 1) & 1);
        if (kAttnProbes && p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && uc < 4) p.dbg[uc * 16 + 12] = clock64();
        // A fragments of the tail queries: fragment rows 0..7 = tail rows (zero beyond tail_n), rows 8..15 unused
        uint32_t qa[4][2];
        {
          const __nv_bfloat16* qrow = p.q + (long long)b * p.q_bs + (long long)(p.tail0 + g) * p.q_ts + h * 64;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            qa[ks][0] = g < p.tail_n ? __ldg(reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 2 * t)) : 0u;
            qa[ks][1] = g < p.tail_n ? __ldg(reinterpret_cast<const uint32_t*>(qrow + ks * 16 + 8 + 2 * t)) : 0u;
          }
        }
        // scores: two 8-key blocks per step
        for (int nb = 0; nb < (p.nkp >> 3); nb += 2) {
          float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const int row = (nb + (mi >> 1)) * 8 + rr, chunk = ks * 2 + (mi & 1);
            uint32_t bf[4];
            atc_ldsm_x4(bf, sK + row * 128 + ((chunk ^ (row & 7)) << 4));
            const uint32_t a[4] = {qa[ks][0], 0u, qa[ks][1], 0u};
            atc_mma_sync(c0, a, bf[0], bf[1]);
            atc_mma_sync(c1, a, bf[2], bf[3]);
          }
          *reinterpret_cast<float2*>(ssc + g * 272 + nb * 8 + 2 * t) = make_float2(c0[0] * p.scale_log2, c0[1] * p.scale_log2);
          *reinterpret_cast<float2*>(ssc + g * 272 + (nb + 1) * 8 + 2 * t) = make_float2(c1[0] * p.scale_log2, c1[1] * p.scale_log2);
        }
        __syncwarp();
        // softmax, one tail row at a time across the warp
        for (int r = 0; r < p.tail_n; ++r) {
          float mx = -INFINITY;
          for (int j = lane; j < p.Nk; j += 32) mx = fmaxf(mx, ssc[r * 272 + j]);
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          float sum = 0.f;
          for (int j = lane; j < p.nkp; j += 32) {
            const float e = j < p.Nk ? atc_ex2(ssc[r * 272 + j] - mx) : 0.f;
            sum += e;
            sp[r * 272 + j] = __float2bfloat16(e);
          }
          sum = warp_sum(sum);
          if (lane == 0) {
            stat[r] = 1.f / sum;
            if (p.lse != nullptr)
              p.lse[((long long)b * p.heads + h) * p.Nq + p.tail0 + r] = (mx + log2f(sum)) * 0.6931471805599453f;
          }
        }
        __syncwarp();
        // O = P V
        float oc[8][4];
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) { oc[nb][0] = oc[nb][1] = oc[nb][2] = oc[nb][3] = 0.f; }
        for (int ks = 0; ks < (p.nkp >> 4); ++ks) {
          uint32_t a[4] = {0u, 0u, 0u, 0u};
          if (g < p.tail_n) {
            a[0] = *reinterpret_cast<const uint32_t*>(sp + g * 272 + ks * 16 + 2 * t);
            a[2] = *reinterpret_cast<const uint32_t*>(sp + g * 272 + ks * 16 + 8 + 2 * t);
          }
#pragma unroll
          for (int nb = 0; nb < 8; nb += 2) {
            const int key = ks * 16 + (mi & 1) * 8 + rr, chunk = nb + (mi >> 1);
            uint32_t bf[4];
            atc_ldsm_x4_t(bf, sV + key * 128 + ((chunk ^ (key & 7)) << 4));
            atc_mma_sync(oc[nb], a, bf[0], bf[1]);
            atc_mma_sync(oc[nb + 1], a, bf[2], bf[3]);
          }
        }
        if (g < p.tail_n) {
          const float inv = stat[g];
          __nv_bfloat16* orow = p.o + (long long)b * p.o_bs + (long long)(p.tail0 + g) * p.o_ts + h * 64;
#pragma unroll
          for (int nb = 0; nb < 8; ++nb)
            *reinterpret_cast<uint32_t*>(orow + nb * 8 + 2 * t) = pack_bf16(oc[nb][0] * inv, oc[nb][1] * inv);
        }
        __syncwarp();
        if (kAttnProbes && p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && uc < 4) p.dbg[uc * 16 + 13] = clock64();
        if (lane == 0) mbar_arrive(&kv_empty[kb]);
      }
    }
  }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    // ------------------------------------------------------------------ softmax + epilogue (thread = query row)
    const int ew = warp - 4;
    const int quad = warp & 3;            // TMEM lanes 32*quad .. +31 (hardware: warp id % 4)
    const int hf = ew >> 2;               // column half: chunks hf, hf + 2, ... of 32 S columns; O columns 32*hf .. +31
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t t_s = tmem_base + lane_base + ATC_TMEM_S;
    const uint32_t t_p = tmem_base + lane_base + ATC_TMEM_P;
    const uint32_t t_o = tmem_base + lane_base + ATC_TMEM_O;
    const uint32_t stage = smem_u32(smem + ATC_OFF_O + ew * 2048);
    const uint32_t row_off = lane * 64, sw = (lane >> 1) & 3;   // 64-byte rows, SWIZZLE_64B
    float* xmax = reinterpret_cast<float*>(smem + ATC_OFF_X);    // [2 halves][128 rows]
    float* xsum = xmax + 256;                                    // [2 halves][128 rows]
    const int row_in_tile = quad * 32 + lane;
    constexpr int MAXC = MULTI ? 4 : 5;   // chunks per warp: 8 / 2 for 256-key blocks, ceil(9 / 2) for a resident range of 272
    int item = 0;
    for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
      const int b = u / p.heads, h = u - b * p.heads;
      for (int qt = 0; qt < p.n_qt; ++qt) {
        // online-softmax state of this thread's row over the key blocks (one block when the key range is resident):
        // running maximum of the raw scores, this warp's partial row sum, and its 32 output columns
        float m_run = -INFINITY, sum = 0.f;
        float o_acc[32];
        if constexpr (MULTI) {
#pragma unroll
          for (int j2 = 0; j2 < 32; ++j2) o_acc[j2] = 0.f;
        }
        for (int j = 0; j < p.nblk; ++j, ++item) {
          const bool last_blk = j == p.nblk - 1;
          const int nk = last_blk ? p.nk_last : 256;
          const int nkp = last_blk ? p.nkp_last : 256;
          const int n_chunks = (nkp + 31) >> 5;
          if (ew == 0) ATC_STAMP(0);
          mbar_wait(s_full, item & 1);
          tc_fence_after();
          if (ew == 0) ATC_STAMP(1);
          // the warp's share of the S row -> registers, then S is free for the next block's QK^T
          uint32_t sv[MAXC][32];
#pragma unroll
          for (int i = 0; i < MAXC; ++i) {
            const int c = hf + 2 * i;
            if (c < n_chunks) tmem_ld_32x32(t_s + c * 32, sv[i]);
          }
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(s_free);
          if (ew == 0) ATC_STAMP(2);
          // columns past the block's keys (zero-filled rows) must not take part: only the last chunk can hold any
          if ((nk & 31) != 0) {
#pragma unroll
            for (int i = 0; i < MAXC; ++i) {
              const int c = hf + 2 * i;
              if (c == n_chunks - 1) {
#pragma unroll
                for (int j2 = 0; j2 < 32; ++j2)
                  if (c * 32 + j2 >= nk) sv[i][j2] = 0xff800000u;   // -inf
              }
            }
          }
          // block maximum: own columns, then the partner warp's through shared memory; fold into the running maximum
          float mx = -INFINITY;
#pragma unroll
          for (int i = 0; i < MAXC; ++i) {
            if (hf + 2 * i < n_chunks) {
#pragma unroll
              for (int j2 = 0; j2 < 32; ++j2) mx = fmaxf(mx, __uint_as_float(sv[i][j2]));
            }
          }
          xmax[hf * 128 + row_in_tile] = mx;
          asm volatile("bar.sync %0, 64;" ::"r"(1 + quad) : "memory");
          mx = fmaxf(mx, xmax[(hf ^ 1) * 128 + row_in_tile]);
          // (no second barrier: the partner cannot reach its next xmax write before this warp has published P for this
          // block -- the o_full wait below needs every warp's P)
          const float m_new = fmaxf(m_run, mx);
          const float alpha = atc_ex2((m_run - m_new) * p.scale_log2);   // 0 on the first block (m_run = -inf)
          m_run = m_new;
          const float ms = m_new * p.scale_log2;
          if (ew == 0) ATC_STAMP(3);
          if (kAttnProbes && p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && item == 4) p.dbg[128 + ew * 2] = clock64();
          // P = exp2(scale * S - max) as bf16 pairs (tcgen05.st), published in two groups; partial row sum in fp32
          sum *= alpha;
#pragma unroll
          for (int i = 0; i < MAXC; ++i) {
            const int c = hf + 2 * i;
            if (c < n_chunks) {
              uint32_t pk[16];
#pragma unroll
              for (int j2 = 0; j2 < 16; ++j2) {
                const float e0 = atc_ex2(fmaf(__uint_as_float(sv[i][2 * j2]), p.scale_log2, -ms));
                const float e1 = atc_ex2(fmaf(__uint_as_float(sv[i][2 * j2 + 1]), p.scale_log2, -ms));
                sum += e0 + e1;
                pk[j2] = FP16 ? pack16(e0, e1, 1) : pack_bf16(e0, e1);
              }
              atc_tmem_st_32x16(t_p + c * 16, pk);
              const bool last = (i == MAXC - 1) || (c + 2 >= n_chunks);
              if (i == 1 && !last) {
                // first group (chunks 0..3) complete for this warp
                atc_tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[0]);
              }
              if (last) {
                if (last_blk) xsum[hf * 128 + row_in_tile] = sum;   // read by the partner after the o_full wait
                atc_tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                  if (i <= 1) mbar_arrive(&p_full[0]);   // short rows: everything is in the first group
                  mbar_arrive(&p_full[1]);
                }
              }
            }
          }
          if (hf >= n_chunks) {   // a short last block leaves this warp without columns: publish nothing, twice
            if (last_blk) xsum[hf * 128 + row_in_tile] = sum;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&p_full[0]); mbar_arrive(&p_full[1]); }
          }
          if (kAttnProbes && p.dbg != nullptr && blockIdx.x == 0 && lane == 0 && item == 4) p.dbg[128 + ew * 2 + 1] = clock64();
          if (ew == 0) ATC_STAMP(4);
          // this block's contribution to the warp's 32 output columns: o = alpha * o + P V_j
          mbar_wait(o_full, item & 1);
          tc_fence_after();
          if (ew == 0) ATC_STAMP(5);
          {
            uint32_t raw[32];
            tmem_ld_32x32(t_o + hf * 32, raw);
            tmem_ld_wait();
            if constexpr (MULTI) {
#pragma unroll
              for (int j2 = 0; j2 < 32; ++j2) o_acc[j2] = fmaf(o_acc[j2], alpha, __uint_as_float(raw[j2]));
            } else {
#pragma unroll
              for (int j2 = 0; j2 < 32; ++j2) o_acc[j2] = __uint_as_float(raw[j2]);
            }
          }
          tc_fence_before();   // (orders the O read before the next block's P publish, i.e. before PV overwrites O)
        }
        // epilogue: O / sum -> bf16 -> swizzled staging tile -> bulk tensor store
        sum += xsum[(hf ^ 1) * 128 + row_in_tile];
        const float inv = 1.f / sum;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncwarp();
        if (ew == 0) ATC_STAMP(7);
        const int row0 = qt * 128 + quad * 32;
        auto emit = [&](const bool f16, const CUtensorMap* tm) {
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int j2 = c * 8;
            const uint32_t w0 = f16 ? pack16(o_acc[j2] * inv, o_acc[j2 + 1] * inv, 1) : pack_bf16(o_acc[j2] * inv, o_acc[j2 + 1] * inv);
            const uint32_t w1 = f16 ? pack16(o_acc[j2 + 2] * inv, o_acc[j2 + 3] * inv, 1) : pack_bf16(o_acc[j2 + 2] * inv, o_acc[j2 + 3] * inv);
            const uint32_t w2 = f16 ? pack16(o_acc[j2 + 4] * inv, o_acc[j2 + 5] * inv, 1) : pack_bf16(o_acc[j2 + 4] * inv, o_acc[j2 + 5] * inv);
            const uint32_t w3 = f16 ? pack16(o_acc[j2 + 6] * inv, o_acc[j2 + 7] * inv, 1) : pack_bf16(o_acc[j2 + 6] * inv, o_acc[j2 + 7] * inv);
            const uint32_t addr = stage + row_off + ((c ^ sw) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            atc_tma_store_3d(tm, stage, h * 64 + hf * 32, row0, b);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        };
        emit(FP16, &tmO);
        if constexpr (FP16) {   // second copy in the other 16-bit format: the first must have left the staging tile
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          __syncwarp();
          emit(!FP16, &tmOalt);
        }
        if (hf == 0 && p.lse != nullptr && row0 + lane < p.Nq)
          p.lse[((long long)b * p.heads + h) * p.Nq + row0 + lane] = (m_run * p.scale_log2 + log2f(sum)) * 0.6931471805599453f;
        if (ew == 0) ATC_STAMP(6);
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

// Returns 1 when the problem is outside this kernel's envelope (the caller falls back to the mma.sync kernel).
int launch_attention_tc_fwd(const b200_attn_desc* d, cudaStream_t st) {
  if (!option(OPT_ATTN_TC_FWD)) return 1;
  if (d->hd != 64 || d->Nk < 33) return 1;
  if (d->Nk > ATC_NKP_MAX && !option(OPT_ATTN_TC_FWD_LONG)) return 1;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al16(d->q) || !al16(d->k) || !al16(d->v) || !al16(d->o) || (d->o_alt != nullptr && !al16(d->o_alt))) return 1;
  if (d->k_bs == 0 || d->v_bs == 0 || d->o_bs == 0) return 1;
  if (d->q_ts % 8 || d->k_ts % 8 || d->v_ts % 8 || d->o_ts % 8 || d->q_bs % 8 || d->k_bs % 8 || d->v_bs % 8 || d->o_bs % 8)
    return 1;
  AttnTcParams p{};
  p.B = d->B; p.heads = d->heads; p.Nq = d->Nq; p.Nk = d->Nk;
  // key blocks: the whole range when it fits the S columns (<= 272), else blocks of 256 with online softmax
  p.nblk = d->Nk <= ATC_NKP_MAX ? 1 : (d->Nk + 255) / 256;
  p.nk_last = d->Nk - (p.nblk - 1) * 256;
  p.nkp_last = (p.nk_last + 15) & ~15;
  p.nkp = p.nkp_last;
  const int rem = d->Nq % 128;
  const int fmt = d->qkvo_is_fp16 ? 0 : 1;
  p.fp16 = d->qkvo_is_fp16 ? 1 : 0;
  p.has_alt = d->o_alt != nullptr ? 1 : 0;
  if (p.fp16 != p.has_alt) return 1;   // two variants are built: bf16 -> bf16 (teacher) and fp16 -> fp16 + bf16 copy (projector)
  p.q_batched = d->q_bs != 0 ? 1 : 0;
  // (the tail-row warp reads bf16 q rows with a bf16 mma.sync and indexes q per batch item)
  if (p.nblk == 1 && rem > 0 && rem <= ATC_TAIL_MAX && fmt == 1 && p.q_batched) {
    p.n_qt = d->Nq / 128; p.tail0 = p.n_qt * 128; p.tail_n = rem;
  } else {
    p.n_qt = (d->Nq + 127) / 128; p.tail0 = 0; p.tail_n = 0;
  }
  if (p.n_qt == 0 || d->scale <= 0.f) return 1;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.lse = d->lse;
  p.q = static_cast<const __nv_bfloat16*>(d->q); p.q_bs = d->q_bs; p.q_ts = d->q_ts;
  p.o = static_cast<__nv_bfloat16*>(d->o); p.o_bs = d->o_bs; p.o_ts = d->o_ts;
  p.n1 = p.nkp_last > 256 ? 256 : p.nkp_last;
  p.n2 = p.nkp_last - p.n1;
  p.idesc_qk1 = make_idesc_16(128, p.n1, false, false, fmt);
  p.idesc_qk2 = p.n2 > 0 ? make_idesc_16(128, p.n2, false, false, fmt) : 0u;
  p.idesc_qk_full = make_idesc_16(128, 256, false, false, fmt);
  p.idesc_pv = make_idesc_16(128, 64, false, true, fmt);

  const uint64_t cols = (uint64_t)d->heads * 64;
  CUtensorMap tq, tk, tv, to, toa;
  B200_TRY(make_tensor_map_3d(&tq, d->q, 2, cols, (uint64_t)d->Nq, (uint64_t)(p.q_batched ? d->B : 1), (uint64_t)d->q_ts,
                              (uint64_t)(p.q_batched ? d->q_bs : (long long)d->Nq * d->q_ts), 64, 128, 1, 128));
  B200_TRY(make_tensor_map_3d(&tk, d->k, 2, cols, (uint64_t)d->Nk, (uint64_t)d->B, (uint64_t)d->k_ts, (uint64_t)d->k_bs, 64, ATC_KV_BOX, 1, 128));
  B200_TRY(make_tensor_map_3d(&tv, d->v, 2, cols, (uint64_t)d->Nk, (uint64_t)d->B, (uint64_t)d->v_ts, (uint64_t)d->v_bs, 64, ATC_KV_BOX, 1, 128));
  B200_TRY(make_tensor_map_3d(&to, d->o, 2, cols, (uint64_t)d->Nq, (uint64_t)d->B, (uint64_t)d->o_ts, (uint64_t)d->o_bs, 32, 32, 1, 64));
  toa = to;
  if (p.has_alt)
    B200_TRY(make_tensor_map_3d(&toa, d->o_alt, 2, cols, (uint64_t)d->Nq, (uint64_t)d->B, (uint64_t)d->o_ts, (uint64_t)d->o_bs, 32, 32, 1, 64));

  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(attn_tc_fwd_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM_BYTES));
    B200_CUDA_OK(cudaFuncSetAttribute(attn_tc_fwd_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM_BYTES));
    B200_CUDA_OK(cudaFuncSetAttribute(attn_tc_fwd_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM_BYTES));
    B200_CUDA_OK(cudaFuncSetAttribute(attn_tc_fwd_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM_BYTES));
    attr_set = true;
  }
  const int units = d->B * d->heads;
  const int grid = units < sm_count() ? units : sm_count();
  const int prof = prof_begin(st);
  static long long* dbg_buf = nullptr;
  const bool dbg_on = kAttnProbes;
  if (dbg_on && dbg_buf == nullptr) { cudaMalloc(&dbg_buf, 9 * 16 * sizeof(long long)); }
  if (dbg_on) cudaMemsetAsync(dbg_buf, 0, 9 * 16 * sizeof(long long), st);
  p.dbg = dbg_on ? dbg_buf : nullptr;
  auto kern = p.nblk > 1 ? (p.fp16 ? attn_tc_fwd_kernel<true, true> : attn_tc_fwd_kernel<true, false>)
                         : (p.fp16 ? attn_tc_fwd_kernel<false, true> : attn_tc_fwd_kernel<false, false>);
  B200_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(ATC_THREADS), ATC_SMEM_BYTES, st, tq, tk, tv, to, toa, p));
  {
    const double D = (double)d->heads * d->hd;   // q (once when batch invariant), k, v read; o (and its copy) written
    prof_end(prof, st, 4.0 * d->B * d->heads * (double)d->Nq * d->Nk * d->hd, 1,
             2.0 * D * ((d->q_bs != 0 ? d->B : 1) * (double)d->Nq + 2.0 * d->B * d->Nk + d->B * (double)d->Nq * (d->o_alt ? 2.0 : 1.0)));
  }
  B200_LAUNCH_OK();
  if (dbg_on) {
    long long h[9 * 16];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, dbg_buf, sizeof h, cudaMemcpyDeviceToHost);
    const long long t0 = h[0];
    for (int it = 0; it < 8; ++it) {
      if (h[it * 16] == 0) break;
      fprintf(stderr, "[attn_tc dbg] item %d:", it);
      for (int s2 = 0; s2 < 16; ++s2) fprintf(stderr, " %s%lld", s2 == 8 ? "| mma " : (s2 == 12 ? "| tail " : (s2 == 14 ? "| w8 " : "")), h[it * 16 + s2] ? h[it * 16 + s2] - t0 : -1);
      fprintf(stderr, "\n");
    }
    fprintf(stderr, "[attn_tc dbg] item 4 exp start/end per softmax warp:");
    for (int w = 0; w < 8; ++w) fprintf(stderr, " w%d %lld-%lld", w + 4, h[128 + 2 * w] - t0, h[129 + 2 * w] - t0);
    fprintf(stderr, "\n");
  }
  return 0;
}

}  // namespace b200
