// tcgen05 / TMEM attention backward for sequences that fit two 128-key blocks (Nq, Nk <= 256: every 224-pixel
// configuration) and head dims up to 64.
//   re-used teacher blocks : hub Attention under autograd, train/distillation_module.py:176-177 (bf16 q/k/v, head_dim 64)
//   ScaleKD cross-attention: losses/scalekd.py:299-314 under autograd (fp16 forward operands, head_dim 16 / 24 / 32 / 48)
//
// One persistent CTA walks (batch, head) units. Per unit, for each 128-key block j and each 64-query item i:
//   S^T [128 k, 64 q] = K_j Q_i^T          tcgen05.mma SS (both K-major)                    -> TMEM, two buffers
//   dP^T[128 k, 64 q] = V_j dO_i^T         tcgen05.mma SS
//   P^T = exp2(scale * S^T - lse_q), dS^T = P^T o (dP^T - delta_q)     8 softmax warps (two sets taking alternate
//        items), thread = key row; the 16-bit
//        results go back into the TMEM columns their fp32 inputs came from (tcgen05.st) and dS^T also into shared
//        memory (128-byte-swizzled rows of 64 queries)
//   dV_j += P^T  dO_i                      tcgen05.mma TS (A from TMEM, B = dO_i MN-major)   -> TMEM accumulator
//   dK_j += dS^T Q_i                       tcgen05.mma TS (B = Q_i MN-major)
//   dQ_m += dS   K_j   (per query pair m)  tcgen05.mma SS (A = staged dS^T read MN-major, B = K_j MN-major)
// Resident mode (Nq, Nk <= 256: every 224-pixel configuration): dQ of the whole unit (<= 256 x hd) stays in TMEM across
// the key blocks, nothing is reduced through memory. Streaming mode (longer sequences: the 518-pixel configurations, 1369
// tokens): one unit per (batch, head, key block); dV_j / dK_j stay in TMEM over all query items, the dQ contribution of
// every query pair is drained from one of two TMEM slots and added into an fp32 accumulator [B, Nq, heads*hd] with
// cp.reduce.async.bulk.tensor (.add.f32), and a small pass scales / converts it to bf16 afterwards.
// Operands arrive by TMA through 4-D maps (head_dim, heads, tokens, batch): boxes wider than the head (24 -> 32,
// 48 -> 64) and rows past the sequence are zero-filled by the hardware; outputs leave through swizzled staging tiles and
// bulk tensor stores (clipped the same way). fp16 forward operands (DUAL): S^T is recomputed from the fp16 q / k so that
// P matches the forward bit for bit, every gradient product uses the bf16 copies the forward saved (tcgen05 kind::f16
// cannot mix the two formats).
//
//   warp 0: TMA producer   warp 1: MMA issuer   warp 2: TMEM alloc   warps 4-11: softmax   warps 12-15: epilogue
#include "common.cuh"
#include "ptx.cuh"
#include "../../include/b200_distill.h"

#include <stdlib.h>

namespace b200 {

int make_tensor_map_4d(CUtensorMap* out, const void* ptr, const uint64_t (&dims)[4], const uint64_t (&ld)[3],
                       const uint32_t (&box)[4], int swizzle, int esize);

constexpr int ABT_THREADS = 512;
constexpr int ABT_SM_WARPS = 8;
constexpr uint32_t ABT_T_SDP = 0;      // [2 buffers][S^T 64 | dP^T 64]
constexpr uint32_t ABT_T_DV = 256, ABT_T_DK = 320, ABT_T_DQ = 384;   // dQ: two 64-column slots (query pairs 0 / 1)
constexpr int ABT_DS_BLOCK = 128 * 128;   // one item's dS^T: 128 keys x 64 queries (bf16)

template <int HDP, bool DUAL>
struct AbtCfg {
  static constexpr int ROWB = HDP * 2;                 // bytes per operand row = swizzle width (128 or 64)
  static constexpr int KTILE = 128 * ROWB;             // key-side tile (128 rows)
  static constexpr int QTILE = 64 * ROWB;              // query-side tile (64 rows)
  static constexpr int NT = DUAL ? 3 : 2;              // tiles per stage: [S operand | bf16 copy (DUAL) | V or dO]
  static constexpr int KVB = (DUAL && HDP == 64) ? 1 : 2;   // key-block buffers
  static constexpr int QST = (DUAL && HDP == 64) ? 3 : 4;   // query-item stages
  static constexpr int OFF_KV = 0;
  static constexpr int OFF_Q = OFF_KV + KVB * NT * KTILE;
  static constexpr int OFF_DS = OFF_Q + QST * NT * QTILE;
  static constexpr int OFF_OUT = OFF_DS + 4 * ABT_DS_BLOCK;       // [4 epilogue warps][4 KB: 32 rows x ROWB 16-bit, or 32 x 32 fp32]
  static constexpr int OFF_X = OFF_OUT + 4 * 4096;           // [8 softmax warps][lse2 64 | delta 64] fp32
  static constexpr int OFF_BAR = OFF_X + ABT_SM_WARPS * 512;
  static constexpr int SMEM = OFF_BAR + 256 + 1024;
  static constexpr uint32_t LAYOUT = HDP == 64 ? 2u : 4u;         // SWIZZLE_128B / SWIZZLE_64B
  static_assert(SMEM <= 232448, "shared memory budget");
};

struct AbtParams {
  int B, heads, Nq, Nk, hd;
  int n_kb, n_items;             // 128-key blocks, 64-query items per unit
  int stream, kb_total;          // streaming mode: unit = (batch, head, key block), n_kb = 1, kb_total blocks per (b, h)
  int q_batched;                 // 0: q (and its copies) are batch invariant
  float scale, scale_log2;
  const float* lse;
  const float* delta;
  float *dq_colsum, *dk_colsum, *dv_colsum;
  uint32_t idesc_s, idesc_dp, idesc_g, idesc_dq;
  long long* dbg;                // -DB200_ATTN_PROBES: clock stamps of CTA 0 ([item][16])
  int skip;                      // -DB200_ATTN_PROBES: work-skipping mask for floor measurements (results are garbage)
};

#ifdef B200_ATTN_PROBES
constexpr bool kAbtProbes = true;
#else
constexpr bool kAbtProbes = false;
#endif
#define ABT_SKIP(bit) (kAbtProbes && (p.skip & (bit)))
#define ABT_STAMP(item, slot)                                                                                    \
  do {                                                                                                            \
    if (kAbtProbes && p.dbg != nullptr && blockIdx.x == 0 && (item) < 48) p.dbg[(item) * 16 + (slot)] = clock64(); \
  } while (0)

__device__ __forceinline__ float abt_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// (unit, key block, query item) walk shared by every role
struct AbtIter {
  int u, j, i, n_units, n_kb, n_items, stride;
  __device__ __forceinline__ bool valid() const { return u < n_units; }
  __device__ __forceinline__ void next() {
    if (++i == n_items) {
      i = 0;
      if (++j == n_kb) { j = 0; u += stride; }
    }
  }
};

// Column sums over the 32 lanes of a warp for 32 per-lane values: afterwards lane c holds the total of column c
// (recursive halving: 31 shuffles instead of 160).
__device__ __forceinline__ float abt_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < off; ++k) {
      const float keep = up ? v[k + off] : v[k];
      const float send = up ? v[k] : v[k + off];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// Epilogue of one [32 rows x HDP] accumulator slice of this warp: TMEM -> registers -> (x scale) -> bf16 -> swizzled
// staging tile -> one bulk tensor store; optional column sums (bias gradients) added to `colsum` with one atomic per
// column. rows_valid: rows of this slice inside the tensor (the store clips the rest; the sums must skip them).
template <int HDP>
__device__ __forceinline__ void abt_store_slice(uint32_t taddr, float scale, const CUtensorMap* tm, uint32_t stage,
                                                int h, int row0, int b, int lane, float* colsum, int hd, int rows_valid,
                                                uint64_t* free_bar) {
  constexpr int ROWB = HDP * 2;
  uint32_t raw[HDP];
#pragma unroll
  for (int c = 0; c < HDP / 32; ++c) tmem_ld_32x32(taddr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&raw[c * 32]));
  tmem_ld_wait();
  if (free_bar != nullptr) {   // the accumulator columns are free as soon as they are in registers
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(free_bar);
  }
  // the previous bulk store of this warp must have finished reading the staging tile
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
#pragma unroll
  for (int c = 0; c < HDP / 8; ++c) {
    const uint32_t w0 = pack_bf16(__uint_as_float(raw[8 * c]) * scale, __uint_as_float(raw[8 * c + 1]) * scale);
    const uint32_t w1 = pack_bf16(__uint_as_float(raw[8 * c + 2]) * scale, __uint_as_float(raw[8 * c + 3]) * scale);
    const uint32_t w2 = pack_bf16(__uint_as_float(raw[8 * c + 4]) * scale, __uint_as_float(raw[8 * c + 5]) * scale);
    const uint32_t w3 = pack_bf16(__uint_as_float(raw[8 * c + 6]) * scale, __uint_as_float(raw[8 * c + 7]) * scale);
    const uint32_t sw = HDP == 64 ? (lane & 7) : ((lane >> 1) & 3);
    const uint32_t addr = stage + lane * ROWB + ((c ^ sw) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w0), "r"(w1), "r"(w2), "r"(w3) : "memory");
  }
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    tma_store_4d(tm, stage, 0, h, row0, b);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  if (colsum != nullptr) {
    const bool ok = lane < rows_valid;
#pragma unroll
    for (int c = 0; c < HDP / 32; ++c) {
      float v[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] = ok ? __uint_as_float(raw[c * 32 + k]) : 0.f;
      const float tot = abt_colsum32(v, lane);
      if (c * 32 + lane < hd) atomicAdd(colsum + h * hd + c * 32 + lane, tot * scale);
    }
  }
}

__device__ __forceinline__ void abt_tma_reduce_add_4d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}

// Streaming mode: one [32 rows x HDP] fp32 slice of a query pair's dQ contribution, TMEM -> registers -> swizzled fp32
// staging tile (32 columns at a time) -> bulk tensor reduce-add into the fp32 accumulator (rows / columns past the
// tensor are clipped by the map).
template <int HDP>
__device__ __forceinline__ void abt_reduce_slice(uint32_t taddr, const CUtensorMap* tm, uint32_t stage, int h, int row0,
                                                 int b, int lane, uint64_t* free_bar) {
  uint32_t raw[HDP];
#pragma unroll
  for (int c = 0; c < HDP / 32; ++c) tmem_ld_32x32(taddr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&raw[c * 32]));
  tmem_ld_wait();
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(free_bar);
#pragma unroll
  for (int t = 0; t < HDP / 32; ++t) {
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const uint32_t addr = stage + lane * 128 + ((c ^ (lane & 7)) << 4);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(raw[t * 32 + 4 * c]), "r"(raw[t * 32 + 4 * c + 1]),
                   "r"(raw[t * 32 + 4 * c + 2]), "r"(raw[t * 32 + 4 * c + 3]) : "memory");
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      abt_tma_reduce_add_4d(tm, stage, t * 32, h, row0, b);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    }
  }
}

template <int HDP, bool DUAL>
__global__ void __launch_bounds__(ABT_THREADS, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQs, const __grid_constant__ CUtensorMap tmKs,
                   const __grid_constant__ CUtensorMap tmQg, const __grid_constant__ CUtensorMap tmKg,
                   const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO,
                   const __grid_constant__ CUtensorMap tmdQ, const __grid_constant__ CUtensorMap tmdK,
                   const __grid_constant__ CUtensorMap tmdV, const AbtParams p) {
  using C = AbtCfg<HDP, DUAL>;
  constexpr int ROWB = C::ROWB, NT = C::NT, KVB = C::KVB, QST = C::QST;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* kv_full = bars;            // [2]
  uint64_t* kv_empty = bars + 2;       // [2]
  uint64_t* q_full = bars + 4;         // [4]
  uint64_t* q_empty = bars + 8;        // [4]
  uint64_t* s_full = bars + 12;        // [2] S^T / dP^T of an item are in TMEM
  uint64_t* p_ready = bars + 14;       // [2] P^T / dS^T written back (8 softmax warps)
  uint64_t* ds_free = bars + 16;       // [2] staged dS^T pair consumed by the dQ product
  uint64_t* dkv_full = bars + 18;      // dV / dK of a key block complete
  uint64_t* dkv_free = bars + 19;      // ... and drained (4 epilogue warps)
  uint64_t* dq_full = bars + 20;       // [2] (streaming mode: one per dQ slot; resident mode uses [0])
  uint64_t* dq_free = bars + 22;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmQs); tma_prefetch_desc(&tmKs); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    if (DUAL) { tma_prefetch_desc(&tmQg); tma_prefetch_desc(&tmKg); }
    tma_prefetch_desc(&tmdQ); tma_prefetch_desc(&tmdK); tma_prefetch_desc(&tmdV);
  }
  if (warp == 1 && elect_one()) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1); mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], ABT_SM_WARPS / 2);
      mbar_init(&ds_free[i], 1);
    }
    for (int i = 0; i < 4; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); }
    mbar_init(dkv_full, 1); mbar_init(dkv_free, 4);
    for (int i = 0; i < 2; ++i) { mbar_init(&dq_full[i], 1); mbar_init(&dq_free[i], 4); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();
  pdl_wait();

  AbtIter it0{(int)blockIdx.x, 0, 0, p.B * p.heads * (p.stream ? p.kb_total : 1), p.n_kb, p.n_items, (int)gridDim.x};
  // unit -> (batch, head) index and the key block an iterator position works on
  auto unit_bh = [&](int u) { return p.stream ? u / p.kb_total : u; };
  auto key_block = [&](const AbtIter& t) { return p.stream ? t.u % p.kb_total : t.j; };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int kvc = 0, ic = 0;
      for (AbtIter it = it0; it.valid(); it.next()) {
        const int bh = unit_bh(it.u), b = bh / p.heads, h = bh - b * p.heads;
        if (it.i == 0) {
          const int kb = kvc % KVB, jg = key_block(it);
          mbar_wait(&kv_empty[kb], ((kvc / KVB) & 1) ^ 1);
          mbar_expect_tx(&kv_full[kb], NT * C::KTILE);
          uint8_t* dst = smem + C::OFF_KV + kb * NT * C::KTILE;
          tma_load_4d(dst, &tmKs, &kv_full[kb], 0, h, jg * 128, b);
          if (DUAL) tma_load_4d(dst + C::KTILE, &tmKg, &kv_full[kb], 0, h, jg * 128, b);
          tma_load_4d(dst + (NT - 1) * C::KTILE, &tmV, &kv_full[kb], 0, h, jg * 128, b);
          ++kvc;
        }
        const int st = ic % QST;
        mbar_wait(&q_empty[st], ((ic / QST) & 1) ^ 1);
        mbar_expect_tx(&q_full[st], NT * C::QTILE);
        uint8_t* dst = smem + C::OFF_Q + st * NT * C::QTILE;
        const int bq = p.q_batched ? b : 0;
        tma_load_4d(dst, &tmQs, &q_full[st], 0, h, it.i * 64, bq);
        if (DUAL) tma_load_4d(dst + C::QTILE, &tmQg, &q_full[st], 0, h, it.i * 64, bq);
        tma_load_4d(dst + (NT - 1) * C::QTILE, &tmdO, &q_full[st], 0, h, it.i * 64, b);
        ++ic;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      AbtIter nx = it0, cur = it0;
      int s_ic = 0, s_kvc = 0;                  // S / dP issue: item and key-block counters
      int g_ic = 0, g_jc = 0, g_pc = 0, g_uc = 0;   // gradient products: item, key-block, pair and unit counters
      int issued = 0, graded = 0;
      const uint32_t ds_base = smem_u32(smem + C::OFF_DS);
      for (;;) {
        if (nx.valid() && issued - graded < 2 && (KVB > 1 || nx.i != 0 || issued == graded)) {
          // ---- S^T and dP^T of item `nx` into TMEM buffer s_ic & 1
          if (nx.i == 0) {
            mbar_wait(&kv_full[s_kvc % KVB], (s_kvc / KVB) & 1);
            ++s_kvc;
          }
          const int kb = (s_kvc - 1) % KVB, st = s_ic % QST, bf = s_ic & 1;
          ABT_STAMP(s_ic, 8);
          mbar_wait(&q_full[st], (s_ic / QST) & 1);
          tc_fence_after();
          ABT_STAMP(s_ic, 9);
          const uint32_t kv_addr = smem_u32(smem + C::OFF_KV + kb * NT * C::KTILE);
          const uint32_t q_addr = smem_u32(smem + C::OFF_Q + st * NT * C::QTILE);
          const uint32_t t_s = tmem_base + ABT_T_SDP + bf * 128, t_dp = t_s + 64;
#pragma unroll
          for (int ks = 0; ks < HDP / 16; ++ks) {
            const uint64_t da = make_smem_desc(kv_addr + ks * 32, 16, 8 * ROWB, C::LAYOUT);
            const uint64_t db = make_smem_desc(q_addr + ks * 32, 16, 8 * ROWB, C::LAYOUT);
            tc_mma_bf16(t_s, da, db, p.idesc_s, ks > 0 ? 1u : 0u);
          }
#pragma unroll
          for (int ks = 0; ks < HDP / 16; ++ks) {
            const uint64_t da = make_smem_desc(kv_addr + (NT - 1) * C::KTILE + ks * 32, 16, 8 * ROWB, C::LAYOUT);
            const uint64_t db = make_smem_desc(q_addr + (NT - 1) * C::QTILE + ks * 32, 16, 8 * ROWB, C::LAYOUT);
            if (!ABT_SKIP(4)) tc_mma_bf16(t_dp, da, db, p.idesc_dp, ks > 0 ? 1u : 0u);
          }
          tc_commit(&s_full[bf]);
          ++s_ic; ++issued;
          nx.next();
          continue;
        }
        if (!cur.valid()) break;
        // ---- gradient products of item `cur`
        {
          const int bf = g_ic & 1, st = g_ic % QST, kb = g_jc % KVB;
          const uint32_t kv_addr = smem_u32(smem + C::OFF_KV + kb * NT * C::KTILE);
          const uint32_t q_addr = smem_u32(smem + C::OFF_Q + st * NT * C::QTILE);
          const uint32_t t_p = tmem_base + ABT_T_SDP + bf * 128, t_ds = t_p + 64;
          ABT_STAMP(g_ic, 10);
          mbar_wait(&p_ready[bf], (g_ic >> 1) & 1);
          ABT_STAMP(g_ic, 11);
          if (cur.i == 0 && g_jc > 0) mbar_wait(dkv_free, (g_jc - 1) & 1);
          tc_fence_after();
          ABT_STAMP(g_ic, 12);
          // dV_j += P^T dO_i ; dK_j += dS^T Q_i     (A from TMEM: the warp with query columns 32*hf.. wrote its 16
          // packed columns at 32*hf, so k-step ks lives at column 32*(ks >> 1) + 8*(ks & 1))
          const uint32_t do_addr = q_addr + (NT - 1) * C::QTILE;
          const uint32_t qg_addr = q_addr + (DUAL ? C::QTILE : 0);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t db = make_smem_desc(do_addr + ks * 16 * ROWB, 8 * ROWB, 8 * ROWB, C::LAYOUT);
            if (!ABT_SKIP(2)) tc_mma_ts(tmem_base + ABT_T_DV, t_p + 32 * (ks >> 1) + 8 * (ks & 1), db, p.idesc_g, (cur.i > 0 || ks > 0) ? 1u : 0u);
          }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t db = make_smem_desc(qg_addr + ks * 16 * ROWB, 8 * ROWB, 8 * ROWB, C::LAYOUT);
            if (!ABT_SKIP(2)) tc_mma_ts(tmem_base + ABT_T_DK, t_ds + 32 * (ks >> 1) + 8 * (ks & 1), db, p.idesc_g, (cur.i > 0 || ks > 0) ? 1u : 0u);
          }
          tc_commit(&q_empty[st]);
          const bool last_item = cur.i == p.n_items - 1;
          if ((cur.i & 1) || last_item) {
            // dQ_m += dS_pair K_j over the 128 keys of the block (A: two staged blocks of 64 queries, MN-major)
            // resident: slot m accumulates over the unit's key blocks; streaming: slot = pair parity, drained per pair
            const int m = cur.i >> 1;
            const int slot = p.stream ? (g_pc & 1) : m;
            if (p.stream) {
              if (g_pc >= 2) { mbar_wait(&dq_free[slot], ((g_pc >> 1) - 1) & 1); tc_fence_after(); }
            } else if (cur.j == 0 && m == 0 && g_uc > 0) {
              mbar_wait(&dq_free[0], (g_uc - 1) & 1);
              tc_fence_after();
            }
            const bool dq_acc = !p.stream && cur.j > 0;
            const uint32_t a_addr = ds_base + (g_pc & 1) * 2 * ABT_DS_BLOCK;
            const uint32_t kg_addr = kv_addr + (DUAL ? C::KTILE : 0);
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
              const uint64_t da = make_smem_desc(a_addr + ks * 2048, ABT_DS_BLOCK, 1024, 2u);
              const uint64_t db = make_smem_desc(kg_addr + ks * 16 * ROWB, 8 * ROWB, 8 * ROWB, C::LAYOUT);
              if (!ABT_SKIP(1)) tc_mma_bf16(tmem_base + ABT_T_DQ + slot * 64, da, db, p.idesc_dq, (dq_acc || ks > 0) ? 1u : 0u);
            }
            if (p.stream) tc_commit(&dq_full[slot]);
            tc_commit(&ds_free[g_pc & 1]);
            ++g_pc;
          }
          if (last_item) {
            tc_commit(&kv_empty[kb]);
            tc_commit(dkv_full);
            ++g_jc;
            if (cur.j == p.n_kb - 1) {
              if (!p.stream) tc_commit(&dq_full[0]);
              ++g_uc;
            }
          }
          ABT_STAMP(g_ic, 13);
          ++g_ic; ++graded;
          cur.next();
        }
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // ------------------------------------------------------------------ softmax warps (thread = key row)
    // The two warps of a TMEM lane quadrant take alternate items (warp set hf <-> TMEM buffer hf), 64 query columns
    // each in two halves: while one set is in its exponentials the other is loading / storing / waiting, so the MUFU
    // pipe sees two items in flight instead of one lock-stepped phase sequence.
    const int ew = warp - 4;
    const int quad = warp & 3;            // TMEM lanes 32*quad .. +31 (hardware: warp id % 4)
    const int hf = ew >> 2;               // item parity / TMEM buffer served by this warp
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    float* xs = reinterpret_cast<float*>(smem + C::OFF_X + ew * 512);   // [lse2 64 | delta 64]
    const uint32_t xs_addr = smem_u32(xs);
    const uint32_t ds_base = smem_u32(smem + C::OFF_DS);
    const uint32_t t_sb = tmem_base + lane_base + ABT_T_SDP + hf * 128;
    const int row = quad * 32 + lane;     // key row inside the block
    // per-query statistics of the warp's NEXT item travel in registers (loaded one item ahead, latency hidden)
    float l0 = 0.f, l1 = 0.f, d0 = 0.f, d1 = 0.f;
    auto ld_stats = [&](const AbtIter& t) {
      if (!t.valid()) return;
      const long long o = (long long)unit_bh(t.u) * p.Nq;
      const int q0 = t.i * 64 + lane, q1 = q0 + 32;
      l0 = q0 < p.Nq ? __ldg(p.lse + o + q0) * 1.4426950408889634f : INFINITY;   // +inf: P = 0 past the last query
      l1 = q1 < p.Nq ? __ldg(p.lse + o + q1) * 1.4426950408889634f : INFINITY;
      d0 = q0 < p.Nq ? __ldg(p.delta + o + q0) : 0.f;
      d1 = q1 < p.Nq ? __ldg(p.delta + o + q1) : 0.f;
    };
    AbtIter mine = it0;
    if (hf) mine.next();
    ld_stats(mine);
    int ic = 0, pc = 0;
    for (AbtIter it = it0; it.valid(); it.next(), ++ic) {
      const bool pair_end = (it.i & 1) || it.i == p.n_items - 1;
      if ((ic & 1) != hf) {
        if (pair_end) ++pc;
        continue;
      }
      if (quad == 0 && lane == 0) ABT_STAMP(ic, 0);
      __syncwarp();
      xs[lane] = l0; xs[32 + lane] = l1; xs[64 + lane] = d0; xs[96 + lane] = d1;
      __syncwarp();
      mine.next(); mine.next();
      ld_stats(mine);
      if (quad == 0 && lane == 0) ABT_STAMP(ic, 1);
      mbar_wait(&s_full[hf], (ic >> 1) & 1);
      // both items of a pair wait for the dQ product that last read the pair's staging blocks
      mbar_wait(&ds_free[pc & 1], ((pc >> 1) & 1) ^ 1);
      tc_fence_after();
      if (quad == 0 && lane == 0) ABT_STAMP(ic, 2);
      const bool row_dead = key_block(it) * 128 + row >= p.Nk;   // rows past the last key (zero-filled K / V) take no part
      const uint32_t blk = ds_base + ((pc & 1) * 2 + (it.i & 1)) * ABT_DS_BLOCK + row * 128;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t sv[32], dv[32];
        tmem_ld_32x32(t_sb + half * 32, sv);
        tmem_ld_32x32(t_sb + 64 + half * 32, dv);
        tmem_ld_wait();
        if (half == 0 && quad == 0 && lane == 0) ABT_STAMP(ic, 3);
        uint32_t pk[16], dk[16];
        if (ABT_SKIP(8)) {
#pragma unroll
          for (int k = 0; k < 16; ++k) { pk[k] = sv[k]; dk[k] = dv[k]; }
        } else
#pragma unroll
        for (int g = 0; g < 8; ++g) {
          float4 l4, d4;
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(l4.x), "=f"(l4.y), "=f"(l4.z), "=f"(l4.w)
                       : "r"(xs_addr + (half * 32 + 4 * g) * 4));
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(d4.x), "=f"(d4.y), "=f"(d4.z), "=f"(d4.w)
                       : "r"(xs_addr + (64 + half * 32 + 4 * g) * 4));
          const float p0 = abt_ex2(fmaf(__uint_as_float(sv[4 * g]), p.scale_log2, -l4.x));
          const float p1 = abt_ex2(fmaf(__uint_as_float(sv[4 * g + 1]), p.scale_log2, -l4.y));
          const float p2 = abt_ex2(fmaf(__uint_as_float(sv[4 * g + 2]), p.scale_log2, -l4.z));
          const float p3 = abt_ex2(fmaf(__uint_as_float(sv[4 * g + 3]), p.scale_log2, -l4.w));
          pk[2 * g] = pack_bf16(p0, p1);
          pk[2 * g + 1] = pack_bf16(p2, p3);
          dk[2 * g] = pack_bf16(p0 * (__uint_as_float(dv[4 * g]) - d4.x), p1 * (__uint_as_float(dv[4 * g + 1]) - d4.y));
          dk[2 * g + 1] = pack_bf16(p2 * (__uint_as_float(dv[4 * g + 2]) - d4.z), p3 * (__uint_as_float(dv[4 * g + 3]) - d4.w));
        }
        if (row_dead) {
#pragma unroll
          for (int k = 0; k < 16; ++k) { pk[k] = 0u; dk[k] = 0u; }
        }
        tmem_st_32x16(t_sb + half * 32, pk);          // P^T over the first 16 columns of this half's S^T
        tmem_st_32x16(t_sb + 64 + half * 32, dk);     // dS^T likewise over dP^T
        // dS^T row -> the item's staging block (rows = keys, 128-byte swizzle, 64 queries per row)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          if (ABT_SKIP(16)) break;
          const uint32_t addr = blk + (((half * 4 + c) ^ (row & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(dk[4 * c]), "r"(dk[4 * c + 1]),
                       "r"(dk[4 * c + 2]), "r"(dk[4 * c + 3]) : "memory");
        }
      }
      if (quad == 0 && lane == 0) ABT_STAMP(ic, 5);
      tmem_st_wait();
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[hf]);
      if (quad == 0 && lane == 0) ABT_STAMP(ic, 6);
      if (pair_end) ++pc;
    }
  } else if (warp >= 12) {
    // ------------------------------------------------------------------ epilogue warps (thread = output row)
    const int quad = warp & 3;
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t stage = smem_u32(smem + C::OFF_OUT + quad * 4096);
    int jc = 0, uc = 0, pc = 0;
    for (AbtIter it = it0; it.valid(); it.next()) {
      const int bh = unit_bh(it.u), b = bh / p.heads, h = bh - b * p.heads;
      const bool pair_end = (it.i & 1) || it.i == p.n_items - 1;
      if (pair_end) {
        if (p.stream) {   // this pair's dQ contribution: TMEM slot -> fp32 reduce-add into the accumulator
          const int slot = pc & 1;
          mbar_wait(&dq_full[slot], (pc >> 1) & 1);
          tc_fence_after();
          abt_reduce_slice<HDP>(tmem_base + lane_base + ABT_T_DQ + slot * 64, &tmdQ, stage, h, (it.i >> 1) * 128 + quad * 32,
                                b, lane, &dq_free[slot]);
        }
        ++pc;
      }
      if (it.i != p.n_items - 1) continue;   // below: one visit per (unit, key block)
      const int row0 = key_block(it) * 128 + quad * 32;
      mbar_wait(dkv_full, jc & 1);
      tc_fence_after();
      if (quad == 0 && lane == 0) ABT_STAMP(jc, 14);
      abt_store_slice<HDP>(tmem_base + lane_base + ABT_T_DV, 1.f, &tmdV, stage, h, row0, b, lane, p.dv_colsum, p.hd,
                           p.Nk - row0, nullptr);
      abt_store_slice<HDP>(tmem_base + lane_base + ABT_T_DK, p.scale, &tmdK, stage, h, row0, b, lane, p.dk_colsum, p.hd,
                           p.Nk - row0, dkv_free);
      if (quad == 0 && lane == 0) ABT_STAMP(jc, 15);
      ++jc;
      if (!p.stream && it.j == p.n_kb - 1) {
        mbar_wait(&dq_full[0], uc & 1);
        tc_fence_after();
        const int n_pairs = (p.n_items + 1) >> 1;
        for (int m = 0; m < n_pairs; ++m) {
          const int qrow0 = m * 128 + quad * 32;
          abt_store_slice<HDP>(tmem_base + lane_base + ABT_T_DQ + m * 64, p.scale, &tmdQ, stage, h, qrow0, b, lane,
                               p.dq_colsum, p.hd, p.Nq - qrow0, m == n_pairs - 1 ? &dq_free[0] : nullptr);
        }
        ++uc;
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<512>(tmem_base);
}

template <int HDP, bool DUAL>
static int abt_launch(const CUtensorMap (&tm)[9], const AbtParams& p, int grid, cudaStream_t st) {
  using C = AbtCfg<HDP, DUAL>;
  auto kern = attn_bwd_tc_kernel<HDP, DUAL>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
    attr_set = true;
  }
  B200_CUDA_OK(launch_pdl(kern, dim3(grid), dim3(ABT_THREADS), C::SMEM, st, tm[0], tm[1], tm[2], tm[3], tm[4], tm[5],
                          tm[6], tm[7], tm[8], p));
  return 0;
}

// Streaming mode, last pass: dq = bf16(scale * accumulator), with the optional column sums (q-bias gradient).
// Thread = one 8-column group, walking the rows of its block: 32-byte loads / 16-byte stores, column partials in registers.
constexpr int ABT_FIN_ROWS = 64;
__global__ void __launch_bounds__(256)
attn_dq_finish_kernel(const float* __restrict__ acc, __nv_bfloat16* __restrict__ dq, long long dq_bs, long long dq_ts,
                      int Nq, int cols, float scale, float* __restrict__ colsum) {
  pdl_wait();
  const int ngroups = cols >> 3;
  const int ry = 256 / ngroups > 0 ? 256 / ngroups : 1;
  const int b = blockIdx.y;
  const int r_begin = blockIdx.x * ABT_FIN_ROWS, r_end = min(Nq, r_begin + ABT_FIN_ROWS);
  for (int cg = threadIdx.x % ngroups + (threadIdx.x / (ngroups * ry)) * ngroups; cg < ngroups; cg += ngroups) {
    // (ngroups <= 256 on this path: every column group has exactly one owner column of threads)
    const int ty = (threadIdx.x / ngroups) % ry;
    if (threadIdx.x >= ngroups * ry) break;
    float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = r_begin + ty; r < r_end; r += ry) {
      const float* src = acc + ((long long)b * Nq + r) * cols + cg * 8;
      const float4 a0 = *reinterpret_cast<const float4*>(src), a1 = *reinterpret_cast<const float4*>(src + 4);
      const float v[8] = {a0.x * scale, a0.y * scale, a0.z * scale, a0.w * scale, a1.x * scale, a1.y * scale, a1.z * scale, a1.w * scale};
      uint4 o;
      o.x = pack_bf16(v[0], v[1]); o.y = pack_bf16(v[2], v[3]); o.z = pack_bf16(v[4], v[5]); o.w = pack_bf16(v[6], v[7]);
      *reinterpret_cast<uint4*>(dq + (long long)b * dq_bs + (long long)r * dq_ts + cg * 8) = o;
#pragma unroll
      for (int e = 0; e < 8; ++e) cs[e] += v[e];
    }
    if (colsum != nullptr) {
#pragma unroll
      for (int e = 0; e < 8; ++e) atomicAdd(colsum + cg * 8 + e, cs[e]);
    }
  }
}

// Returns 1 when the problem is outside this kernel's envelope (the caller falls back to the mma.sync kernels), 0 when
// the backward was launched, negative on error. delta must already be in d->delta.
int launch_attention_tc_bwd(const b200_attn_desc* d, cudaStream_t st) {
  if (!option(OPT_ATTN_TC_BWD)) return 1;
  const bool dual = d->qkvo_is_fp16 != 0;
  const bool stream = d->Nq > 256 || d->Nk > 256;
  if (d->hd < 16 || d->hd > 64 || d->hd % 8 != 0) return 1;
  if (stream && (d->dq_accum == nullptr || !option(OPT_ATTN_TC_BWD_LONG) || d->heads * d->hd > 2048)) return 1;
  if (dual && (d->q_alt == nullptr || d->k_alt == nullptr || d->v_alt == nullptr)) return 1;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (!al16(d->q) || !al16(d->k) || !al16(d->v) || !al16(d->d_o) || !al16(d->dq) || !al16(d->dk) || !al16(d->dv)) return 1;
  if (dual && (!al16(d->q_alt) || !al16(d->k_alt) || !al16(d->v_alt))) return 1;
  if (stream && !al16(d->dq_accum)) return 1;
  const long long strides[] = {d->q_ts, d->k_ts, d->v_ts, d->do_ts, d->dq_ts, d->dk_ts, d->dv_ts,
                               d->q_bs, d->k_bs, d->v_bs, d->do_bs, d->dq_bs, d->dk_bs, d->dv_bs};
  for (long long s : strides)
    if (s % 8 != 0 || s < 0) return 1;
  if (d->k_bs == 0 || d->v_bs == 0 || d->do_bs == 0 || d->dq_bs == 0 || d->dk_bs == 0 || d->dv_bs == 0) return 1;
  const int hdp = d->hd <= 32 ? 32 : 64;
  const int cols = d->heads * d->hd;

  AbtParams p{};
  p.B = d->B; p.heads = d->heads; p.Nq = d->Nq; p.Nk = d->Nk; p.hd = d->hd;
  p.kb_total = (d->Nk + 127) / 128;
  p.stream = stream ? 1 : 0;
  p.n_kb = stream ? 1 : p.kb_total;
  p.n_items = (d->Nq + 63) / 64;
  p.q_batched = d->q_bs != 0 ? 1 : 0;
  p.scale = d->scale;
  p.scale_log2 = d->scale * 1.4426950408889634f;
  p.lse = d->lse; p.delta = d->delta;
  p.dq_colsum = stream ? nullptr : d->dq_colsum;   // (streaming: summed by the finishing pass)
  p.dk_colsum = d->dk_colsum; p.dv_colsum = d->dv_colsum;
  p.idesc_s = make_idesc_16(128, 64, false, false, dual ? 0 : 1);
  p.idesc_dp = make_idesc_16(128, 64, false, false, 1);
  p.idesc_g = make_idesc_16(128, hdp, false, true, 1);
  p.idesc_dq = make_idesc_16(128, hdp, true, true, 1);

  const int sw = hdp * 2;
  auto mk = [&](CUtensorMap* out, const void* ptr, int N, long long ts, long long bs, int box_rows) -> int {
    const bool batched = bs != 0;
    const uint64_t dims[4] = {(uint64_t)d->hd, (uint64_t)d->heads, (uint64_t)N, (uint64_t)(batched ? d->B : 1)};
    const uint64_t ld[3] = {(uint64_t)d->hd, (uint64_t)ts, (uint64_t)(batched ? bs : (long long)N * ts)};
    const uint32_t box[4] = {(uint32_t)hdp, 1u, (uint32_t)box_rows, 1u};
    return make_tensor_map_4d(out, ptr, dims, ld, box, sw, 2);
  };
  CUtensorMap tm[9];
  // a tensor-map encoding the driver rejects is not an error of the caller: use the fallback kernels
  const void* qg = dual ? d->q_alt : d->q;
  const void* kg = dual ? d->k_alt : d->k;
  const void* vg = dual ? d->v_alt : d->v;
  int bad = mk(&tm[0], d->q, d->Nq, d->q_ts, d->q_bs, 64) || mk(&tm[1], d->k, d->Nk, d->k_ts, d->k_bs, 128) ||
            mk(&tm[2], qg, d->Nq, d->q_ts, d->q_bs, 64) || mk(&tm[3], kg, d->Nk, d->k_ts, d->k_bs, 128) ||
            mk(&tm[4], vg, d->Nk, d->v_ts, d->v_bs, 128) || mk(&tm[5], d->d_o, d->Nq, d->do_ts, d->do_bs, 64) ||
            mk(&tm[7], d->dk, d->Nk, d->dk_ts, d->dk_bs, 32) || mk(&tm[8], d->dv, d->Nk, d->dv_ts, d->dv_bs, 32);
  if (!bad) {
    if (stream) {   // fp32 accumulator [B, Nq, heads*hd]: boxes of 32 columns x 32 rows
      const uint64_t dims[4] = {(uint64_t)d->hd, (uint64_t)d->heads, (uint64_t)d->Nq, (uint64_t)d->B};
      const uint64_t ld[3] = {(uint64_t)d->hd, (uint64_t)cols, (uint64_t)d->Nq * cols};
      const uint32_t box[4] = {32u, 1u, 32u, 1u};
      bad = make_tensor_map_4d(&tm[6], d->dq_accum, dims, ld, box, 128, 4);
    } else {
      bad = mk(&tm[6], d->dq, d->Nq, d->dq_ts, d->dq_bs, 32);
    }
  }
  if (bad) {
    static bool warned = false;
    if (!warned) { fprintf(stderr, "[b200] attention backward: tensor map rejected, using the mma.sync path\n"); warned = true; }
    return 1;
  }
  const int units = d->B * d->heads * (stream ? p.kb_total : 1);
  const int grid = units < sm_count() ? units : sm_count();
  static long long* dbg_buf = nullptr;
  if (kAbtProbes) {
    if (dbg_buf == nullptr) cudaMalloc(&dbg_buf, 48 * 16 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 48 * 16 * sizeof(long long), st);
    p.dbg = dbg_buf;
    p.skip = option(OPT_ATTN_PROBE_SKIP);
  }
  if (stream) B200_CUDA_OK(cudaMemsetAsync(d->dq_accum, 0, (size_t)d->B * d->Nq * cols * sizeof(float), st));
  const int prof = prof_begin(st);
  int r;
  if (hdp == 64) r = dual ? abt_launch<64, true>(tm, p, grid, st) : abt_launch<64, false>(tm, p, grid, st);
  else r = dual ? abt_launch<32, true>(tm, p, grid, st) : abt_launch<32, false>(tm, p, grid, st);
  if (r != 0) return r;
  {
    const double D = (double)d->heads * d->hd;   // q, dO read and dq written per query; k, v read and dk, dv written per key
    prof_end(prof, st, 8.0 * d->B * d->heads * (double)d->Nq * d->Nk * d->hd, 2,
             2.0 * D * d->B * (3.0 * d->Nq + 4.0 * d->Nk) * (d->qkvo_is_fp16 ? 1.5 : 1.0));
  }
  B200_LAUNCH_OK();
  if (stream) {
    dim3 g((unsigned)cdiv(d->Nq, ABT_FIN_ROWS), (unsigned)d->B);
    B200_CUDA_OK(launch_pdl(attn_dq_finish_kernel, g, dim3(256), 0, st, (const float*)d->dq_accum,
                            static_cast<__nv_bfloat16*>(d->dq), (long long)d->dq_bs, (long long)d->dq_ts, d->Nq, cols, d->scale,
                            d->dq_colsum));
    B200_LAUNCH_OK();
  }
  if (kAbtProbes) {
    static int printed = 0;
    cudaStreamSynchronize(st);
    if (printed++ == 3 && p.skip == 0) {   // (a warm launch)
      long long h[48 * 16];
      cudaMemcpy(h, dbg_buf, sizeof h, cudaMemcpyDeviceToHost);
      const long long t0 = h[0];
      fprintf(stderr, "[abt probes] softmax w0: start lse_staged s_full ld_done computed stored arrived | mma: sdp_begin q_full | grads_begin p_ready dkv_free done | epi: dkv_full dk_stored\n");
      for (int it = 0; it < 48; ++it) {
        if (h[it * 16] == 0 && it > 0) break;
        fprintf(stderr, "[abt probes] item %2d:", it);
        for (int s2 = 0; s2 < 16; ++s2) {
          if (s2 == 7) { fprintf(stderr, " |"); continue; }
          if (s2 == 10 || s2 == 14) fprintf(stderr, " |");
          fprintf(stderr, " %6lld", h[it * 16 + s2] ? h[it * 16 + s2] - t0 : -1);
        }
        fprintf(stderr, "\n");
      }
    }
  }
  return 0;
}

}  // namespace b200
