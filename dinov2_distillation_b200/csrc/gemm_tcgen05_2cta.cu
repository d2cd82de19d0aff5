// CTA-pair (cta_group::2) variant of the tcgen05 GEMM for K-major operands: C[M,N] = A[M,K] * B[N,K]^T.
//
// Two CTAs of a cluster (same TPC) compute one 256 x BN tile: each CTA stages its own 128 rows of A and its own BN/2
// rows of B, the leader issues tcgen05.mma.cta_group::2 (M = 256) which reads both CTAs' shared memory, and each CTA
// keeps its 128 x BN half of the accumulator in its own TMEM. Per CTA and k-block this moves 16 KB (A) + BN/2*128 B (B)
// instead of 16 KB + BN*128 B, which is what the small-K shapes of the ViT-S/B teacher need: with 128-row tiles they are
// bound by L2 -> SM operand traffic, not by the tensor pipe.
//
//   warp 0 (both CTAs): TMA producer  -- cp.async.bulk.tensor .cta_group::2, transaction bytes land on the LEADER's barrier
//   warp 1 (leader)   : MMA issuer    -- tcgen05.commit multicast releases the smem slot / publishes the accumulator in
//                                        both CTAs
//   warp 2 (both)     : TMEM alloc / free (cta_group::2)
//   warps 4..19 (both): epilogue of the CTA's own 128 rows; arrive on the leader's tmem_empty barrier
#include "gemm_common.cuh"

#include <stdlib.h>

namespace b200 {

template <int BN>
struct PairCfg {
  static constexpr int HALF_BN = BN / 2;
  static constexpr int B_TILE_BYTES = HALF_BN * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;   // per CTA
  static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;
  static constexpr int TMEM_STRIDE = (BN <= 128) ? 128 : 256;
  static constexpr int TMEM_COLS = 2 * TMEM_STRIDE;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* leader_bar, int c0,
                                                 int c1) {
  // executed by both CTAs; clearing the peer bit makes the transaction bytes land on CTA 0's barrier
  const uint32_t bar = smem_u32(leader_bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_mma_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  const uint16_t mask = 3;
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}\n"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_slot) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCols>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}

template <int BN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const GemmParams p) {
  using Cfg = PairCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);    // leader only: its own arrive.expect_tx covers the bytes of both CTAs
      mbar_init(&empty_bar[s], 1);   // one multicast tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 2 * EPI_WARPS);   // leader only: epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int pair = blockIdx.x >> 1;
  const int npairs = gridDim.x >> 1;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = pair; t < p.total_tiles; t += npairs) {
        const int m0 = (t / p.n_tiles) * 256 + (int)rank * 128;
        const int n0 = (t % p.n_tiles) * BN + (int)rank * Cfg::HALF_BN;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + A_TILE_BYTES;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
          tma_load_2d_pair(sA, &tmA, &full_bar[stage], kb * BK, m0);
          tma_load_2d_pair(sB, &tmB, &full_bar[stage], kb * BK, n0);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && elect_one()) {
      const uint32_t idesc = p.idesc;
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = pair; t < p.total_tiles; t += npairs) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * Cfg::TMEM_STRIDE;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            const uint64_t da = make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            tc_mma_pair(d_tmem, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          tc_commit_pair(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit_pair(&tfull_bar[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;
    const int part = (warp - 4) >> 2;
    int as = 0;
    uint32_t aphase = 0;
    for (int t = pair; t < p.total_tiles; t += npairs) {
      const int m0 = (t / p.n_tiles) * 256 + (int)rank * 128;
      const int n0 = (t % p.n_tiles) * BN;
      const int row0 = m0 + quad * 32;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * Cfg::TMEM_STRIDE;
#pragma unroll 1
      for (int c = part; c < BN / EPI_W; c += EPI_WARPS / 4) {
        uint32_t raw[EPI_W];
        tmem_ld_32x16(taddr + c * EPI_W, raw);
        tmem_ld_wait();
        float v[EPI_W];
#pragma unroll
        for (int j = 0; j < EPI_W; ++j) v[j] = __uint_as_float(raw[j]);
        if (p.dbg & 1) {
          if (v[0] == 123.456f && p.out_f32) p.out_f32[0] = v[1];
          continue;
        }
        epilogue_chunk<EPI_W>(p, row0 + lane < p.M, row0 + lane, n0 + c * EPI_W, v, true);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&tempty_bar[as], 0);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_pair<Cfg::TMEM_COLS>(tmem_base);
}

template <int BN>
static int launch_pair(const b200_gemm_desc* d, GemmParams& p, cudaStream_t st) {
  using Cfg = PairCfg<BN>;
  auto kern = gemm_tcgen05_pair_kernel<BN>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap ta, tb;
  B200_TRY(make_tensor_map_2d(&ta, d->A, (uint64_t)d->K, (uint64_t)d->M, (uint64_t)d->lda, BK, BM));
  B200_TRY(make_tensor_map_2d(&tb, d->B, (uint64_t)d->K, (uint64_t)d->N, (uint64_t)d->ldb, BK, (uint32_t)Cfg::HALF_BN));
  p.n_tiles = (int)cdiv(d->N, BN);
  p.m_tiles = (int)cdiv(d->M, 256);
  p.total_tiles = p.m_tiles * p.n_tiles;
  p.split_k = 1;
  p.kb_per_split = p.num_kb;
  p.idesc = make_idesc_bf16(256, BN, false, false);
  if (d->a_is_fp16) p.idesc &= ~(7u << 7);
  if (d->b_is_fp16) p.idesc &= ~(7u << 10);
  int pairs = sm_count() / 2;
  if (pairs > p.total_tiles) pairs = p.total_tiles;
  const int prof = prof_begin(st);
  kern<<<pairs * 2, GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(ta, tb, p);
  prof_end(prof, st, 2.0 * p.M * p.N * p.K * p.algo_scale, 0);
  B200_LAUNCH_OK();
  return 0;
}

int launch_gemm_2cta(const b200_gemm_desc* d, GemmParams& p, cudaStream_t st) {
  static int mode = -1;  // B200_GEMM_2CTA=0 disables the pair kernel (A/B measurements)
  if (mode < 0) {
    const char* e = getenv("B200_GEMM_2CTA");
    mode = (e && e[0] == '0') ? 0 : 1;
  }
  if (!mode || d->M < 1024 || d->N < 128) return 1;
  // fewest wasted columns / waves; prefer the widest tile on ties
  const int cands[3] = {256, 192, 128};
  int best = 256;
  double best_cost = 1e30;
  const long long mt = cdiv(d->M, 256);
  const int pairs = sm_count() / 2;
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const long long tiles = cdiv(d->N, bn) * mt;
    const double cost = double(cdiv(tiles, pairs)) * bn * (bn == 128 ? 1.08 : 1.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  static const int forced = gemm_env_int("B200_GEMM_BN", 0);
  if (forced == 128 || forced == 192 || forced == 256) best = forced;
  if (best == 256) return launch_pair<256>(d, p, st);
  if (best == 192) return launch_pair<192>(d, p, st);
  return launch_pair<128>(d, p, st);
}

}  // namespace b200
