// Persistent warp-specialised bf16 GEMM for sm_100a: TMA (128B swizzle) -> smem ring -> tcgen05.mma -> TMEM
// (double-buffered accumulators) -> tcgen05.ld epilogue with fused bias / activation / LayerScale / residual.
//
//   warp 0      : TMA producer (one elected lane)
//   warp 1      : MMA issuer   (one elected lane; tcgen05.commit releases smem slots / publishes accumulators)
//   warp 2      : TMEM allocate / free
//   warps 4..7  : epilogue (warp w reads TMEM lanes 32*(w%4) .. +31, one output row per thread)
//
// C[M,N] = A * B^T.  Operands may be K-major (contraction contiguous; the nn.Linear layout) or MN-major (output dim
// contiguous; used by wgrad dW = dY^T X where both operands are token-major).
#include "gemm_common.cuh"

#include <mutex>
#include <stdlib.h>
#include <unordered_map>

namespace b200 {

template <int BN>
struct TileCfg {
  static constexpr int B_TILE_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_TILE_BYTES + B_TILE_BYTES;
  static constexpr int STAGES = (BN == 128) ? 6 : (BN == 192 ? 5 : 4);
  static constexpr int TMEM_STRIDE = (BN <= 128) ? 128 : 256;
  static constexpr int TMEM_COLS = 2 * TMEM_STRIDE;
  static constexpr int SMEM_BYTES =
      STAGES * STAGE_BYTES + EPI_WARPS * EPI_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
  using Cfg = TileCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + EPI_WARPS * EPI_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], EPI_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int tiles_mn = p.m_tiles * p.n_tiles;

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int ks = t / tiles_mn;
        const int r = t - ks * tiles_mn;
        const int m0 = (r / p.n_tiles) * BM;
        const int n0 = (r % p.n_tiles) * BN;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sB = sA + A_TILE_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          if (!A_MN) {
            tma_load_2d(sA, &tmA, &full_bar[stage], kb * BK, m0);
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d(sA + j * 8192, &tmA, &full_bar[stage], m0 + j * 64, kb * BK);
          }
          if (!B_MN) {
            tma_load_2d(sB, &tmB, &full_bar[stage], kb * BK, n0);
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; ++j) tma_load_2d(sB + j * 8192, &tmB, &full_bar[stage], n0 + j * 64, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = p.idesc;
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const int ks = t / tiles_mn;
        const int kb0 = ks * p.kb_per_split;
        const int kb1 = min(p.num_kb, kb0 + p.kb_per_split);
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * Cfg::TMEM_STRIDE;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_TILE_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major: 16 elements = 32 bytes inside the 128B swizzle row; SBO = 8 rows * 128B.
            // MN-major: 16 contraction rows = 2 swizzle atoms of 1024B; LBO = next 64-wide column block (64 rows*128B).
            const uint64_t da = A_MN ? make_smem_desc_sw128(a_addr + k * 2048, 8192, 1024)
                                     : make_smem_desc_sw128(a_addr + k * 32, 16, 1024);
            const uint64_t db = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, 8192, 1024)
                                     : make_smem_desc_sw128(b_addr + k * 32, 16, 1024);
            tc_mma_bf16(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          tc_commit(&empty_bar[stage]);
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull_bar[as]);
        if (++as == 2) { as = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    const int quad = warp & 3;            // TMEM lanes 32*quad .. +31 (hardware: warp id % 4)
    const int part = (warp - 4) >> 2;     // which EPI_W-column blocks of the tile this warp drains
    int as = 0;
    uint32_t aphase = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const int ks = t / tiles_mn;
      const int r = t - ks * tiles_mn;
      const int m0 = (r / p.n_tiles) * BM;
      const int n0 = (r % p.n_tiles) * BN;
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
        if (kGemmProbes && (p.dbg & 1)) {
          if (v[0] == 123.456f && p.out_f32) p.out_f32[0] = v[1];
          continue;
        }
        epilogue_chunk<EPI_W>(p, row0 + lane < p.M, row0 + lane, n0 + c * EPI_W, v, ks == 0);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) { as = 0; aphase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, []() {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t d0, d1, ld;
  uint32_t b0, b1, esize, swizzle;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && ld == o.ld && b0 == o.b0 && b1 == o.b1 && esize == o.esize &&
           swizzle == o.swizzle;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= k.d0 * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= k.d1 * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= k.ld * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
    h ^= (uint64_t(k.b0) << 32 | k.b1) + (h << 6) + (h >> 2);
    h ^= (uint64_t(k.esize) << 32 | k.swizzle) + (h << 6) + (h >> 2);
    return h;
  }
};

// 2-D tensor map: dim0 (contiguous) x dim1 rows with row pitch ld elements of esize bytes (2: 16-bit, 4: fp32);
// box b0 x b1; swizzle 128 / 64 / 32 bytes. Encoded maps are cached by (pointer, shape, box).
int make_tensor_map_ex(CUtensorMap* out, const void* ptr, int esize, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0,
                       uint32_t b1, int swizzle) {
  static std::mutex mu;
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  MapKey key{ptr, d0, d1, ld, b0, b1, (uint32_t)esize, (uint32_t)swizzle};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) { set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)"); return -4; }
  cuuint64_t gdim[2] = {d0, d1};
  cuuint64_t gstride[1] = {ld * (uint64_t)esize};
  cuuint32_t box[2] = {b0, b1};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapSwizzle sw = swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d) ptr=%p esize=%d dims=(%llu,%llu) ld=%llu box=(%u,%u)",
             (int)r, ptr, esize, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)ld, b0, b1);
    set_error(buf);
    return -4;
  }
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

// 3-D tensor map (attention operands: head columns x tokens x batch). ld1 / ld2: element pitch of dim 1 / dim 2.
int make_tensor_map_3d(CUtensorMap* out, const void* ptr, int esize, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t ld1,
                       uint64_t ld2, uint32_t b0, uint32_t b1, uint32_t b2, int swizzle) {
  struct Key3 {
    const void* ptr; uint64_t d0, d1, d2, ld1, ld2; uint32_t b0, b1, b2, es_sw;
    bool operator==(const Key3& o) const {
      return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && ld1 == o.ld1 && ld2 == o.ld2 && b0 == o.b0 &&
             b1 == o.b1 && b2 == o.b2 && es_sw == o.es_sw;
    }
  };
  struct Key3Hash {
    size_t operator()(const Key3& k) const {
      size_t h = reinterpret_cast<size_t>(k.ptr);
      const uint64_t v[8] = {k.d0, k.d1, k.d2, k.ld1, k.ld2, (uint64_t(k.b0) << 32) | k.b1, k.b2, k.es_sw};
      for (uint64_t x : v) h ^= x * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
      return h;
    }
  };
  static std::mutex mu;
  static std::unordered_map<Key3, CUtensorMap, Key3Hash> cache;
  Key3 key{ptr, d0, d1, d2, ld1, ld2, b0, b1, b2, (uint32_t)(esize * 1000 + swizzle)};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) { set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)"); return -4; }
  if (d2 == 1 && ld2 == 0) ld2 = ld1 * d1;
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstride[2] = {ld1 * (uint64_t)esize, ld2 * (uint64_t)esize};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapSwizzle sw = swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                  const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[320];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(3d) failed (%d) ptr=%p dims=(%llu,%llu,%llu) ld=(%llu,%llu) box=(%u,%u,%u)",
             (int)r, ptr, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
             (unsigned long long)ld1, (unsigned long long)ld2, b0, b1, b2);
    set_error(buf);
    return -4;
  }
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

// Rank-4 tensor map over 16-bit (esize 2) or fp32 (esize 4) elements (attention backward: head_dim x heads x tokens x batch, so that a box wider than
// head_dim is zero-filled past the head instead of reading into the next one). ld[i]: element pitch of dim i+1.
// Returns -4 with the error text set when the driver rejects the encoding.
int make_tensor_map_4d(CUtensorMap* out, const void* ptr, const uint64_t (&dims)[4], const uint64_t (&ld)[3],
                       const uint32_t (&box)[4], int swizzle, int esize) {
  struct Key4 {
    const void* ptr; uint64_t d[4], l[3]; uint32_t b[4], sw;
    bool operator==(const Key4& o) const {
      if (ptr != o.ptr || sw != o.sw) return false;
      for (int i = 0; i < 4; ++i) if (d[i] != o.d[i] || b[i] != o.b[i]) return false;
      for (int i = 0; i < 3; ++i) if (l[i] != o.l[i]) return false;
      return true;
    }
  };
  struct Key4Hash {
    size_t operator()(const Key4& k) const {
      size_t h = reinterpret_cast<size_t>(k.ptr) ^ k.sw;
      for (int i = 0; i < 4; ++i) h ^= (k.d[i] * 0x9E3779B97F4A7C15ull + k.b[i]) + (h << 6) + (h >> 2);
      for (int i = 0; i < 3; ++i) h ^= k.l[i] * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
      return h;
    }
  };
  static std::mutex mu;
  static std::unordered_map<Key4, CUtensorMap, Key4Hash> cache;
  Key4 key{};
  key.ptr = ptr; key.sw = (uint32_t)(swizzle + 1000 * esize);
  for (int i = 0; i < 4; ++i) { key.d[i] = dims[i]; key.b[i] = box[i]; }
  for (int i = 0; i < 3; ++i) key.l[i] = ld[i];
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) { set_error("cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)"); return -4; }
  cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gstride[3] = {ld[0] * (uint64_t)esize, ld[1] * (uint64_t)esize, ld[2] * (uint64_t)esize};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = swizzle == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = fn(out, esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                  const_cast<void*>(ptr), gdim, gstride, bx, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[384];
    snprintf(buf, sizeof buf,
             "cuTensorMapEncodeTiled(4d) failed (%d) ptr=%p dims=(%llu,%llu,%llu,%llu) ld=(%llu,%llu,%llu) box=(%u,%u,%u,%u) sw=%d",
             (int)r, ptr, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
             (unsigned long long)dims[3], (unsigned long long)ld[0], (unsigned long long)ld[1], (unsigned long long)ld[2],
             box[0], box[1], box[2], box[3], swizzle);
    set_error(buf);
    return -4;
  }
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

// operand maps: 16-bit elements, 128-byte swizzle
int make_tensor_map_2d(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t ld, uint32_t b0,
                       uint32_t b1) {
  return make_tensor_map_ex(out, ptr, 2, d0, d1, ld, b0, b1, 128);
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  using Cfg = TileCfg<BN>;
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN>;
  static bool attr_set = false;
  if (!attr_set) {
    B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  const int grid = p.total_tiles < sm_count() ? p.total_tiles : sm_count();
  const int prof = prof_begin(st);
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(ta, tb, p);
  prof_end(prof, st, 2.0 * p.M * p.N * p.K * p.algo_scale, 0);
  B200_LAUNCH_OK();
  return 0;
}

template <int BN>
static int dispatch_major(bool a_mn, bool b_mn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                          cudaStream_t st) {
  if (!a_mn && !b_mn) return launch_gemm<BN, false, false>(ta, tb, p, st);
  if (a_mn && b_mn) return launch_gemm<BN, true, true>(ta, tb, p, st);
  if (a_mn) return launch_gemm<BN, true, false>(ta, tb, p, st);
  return launch_gemm<BN, false, true>(ta, tb, p, st);
}

static int pick_bn(int N, int m_tiles) {
  const int forced = option(OPT_GEMM_BN);
  if (forced == 128 || forced == 192 || forced == 256) return forced;
  // fewest wasted columns first; among equals prefer the widest tile (less smem traffic per MMA).
  const int cands[3] = {256, 192, 128};
  int best = 128;
  double best_cost = 1e30;
  const int sms = sm_count();
  for (int i = 0; i < 3; ++i) {
    const int bn = cands[i];
    const long long nt = cdiv(N, bn);
    const long long tiles = nt * m_tiles;
    const long long waves = cdiv(tiles, sms);
    // cost ~ waves * per-tile time (proportional to bn) ; small penalty for narrow tiles
    double cost = double(waves) * bn * (bn == 128 ? 1.06 : 1.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_gemm_bf16(const b200_gemm_desc* d, void* stream) {
  B200_CHECK_ARG(d != nullptr, "null descriptor");
  B200_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0, "empty problem");
  B200_CHECK_ARG(d->A && d->B, "null operand");
  B200_CHECK_ARG((reinterpret_cast<uintptr_t>(d->A) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->B) & 15) == 0,
                 "operands must be 16-byte aligned");
  B200_CHECK_ARG(d->lda % 8 == 0 && d->ldb % 8 == 0, "operand leading dims must be multiples of 8 elements");
  B200_CHECK_ARG(d->out_f32 || d->out_bf16, "no output");
  const int split = d->split_k < 1 ? 1 : d->split_k;
  if (split > 1) {
    B200_CHECK_ARG(d->out_f32 && d->atomic_add && !d->out_bf16 && !d->out_bf16_pre && d->act == 0 &&
                       d->aux_mode == 0 && d->col_scale == nullptr,
                   "split_k > 1 needs a plain atomic fp32 epilogue");
  }
  GemmParams p{};
  p.M = d->M; p.N = d->N; p.K = d->K;
  p.m_tiles = (int)cdiv(d->M, BM);
  p.num_kb = (int)cdiv(d->K, BK);
  int sk = split > p.num_kb ? p.num_kb : split;
  p.kb_per_split = (int)cdiv(p.num_kb, sk);
  p.split_k = (int)cdiv(p.num_kb, p.kb_per_split);
  const int bn = pick_bn(d->N, p.m_tiles * p.split_k);
  p.n_tiles = (int)cdiv(d->N, bn);
  p.total_tiles = p.m_tiles * p.n_tiles * p.split_k;
  p.bias = d->bias; p.act = d->act;
  p.aux = static_cast<const __nv_bfloat16*>(d->aux); p.ldaux = d->ldaux; p.aux_mode = d->aux ? d->aux_mode : 0;
  p.col_scale = d->col_scale;
  p.residual = d->residual; p.ldres = d->ldres; p.res_row_period = d->res_row_period;
  p.out_f32 = d->out_f32; p.ldo32 = d->ldo32; p.atomic_add = d->atomic_add;
  p.out_bf16 = static_cast<__nv_bfloat16*>(d->out_bf16); p.ldo16 = d->ldo16;
  p.out_bf16_pre = static_cast<__nv_bfloat16*>(d->out_bf16_pre); p.ldo16_pre = d->ldo16_pre;
  p.out_row_period = d->out_row_period; p.out_row_pad = d->out_row_pad;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  bool vec = true;
  if (p.bias) vec = vec && al16(p.bias);
  if (p.aux_mode) vec = vec && al16(p.aux) && p.ldaux % 8 == 0;
  if (p.col_scale) vec = vec && al16(p.col_scale);
  if (p.residual) vec = vec && al16(p.residual) && p.ldres % 4 == 0;
  if (p.out_f32) vec = vec && al16(p.out_f32) && p.ldo32 % 4 == 0;
  if (p.out_bf16) vec = vec && al16(p.out_bf16) && p.ldo16 % 8 == 0;
  if (p.out_bf16_pre) vec = vec && al16(p.out_bf16_pre) && p.ldo16_pre % 8 == 0;
  p.vec_ok = vec ? 1 : 0;
  p.algo_scale = d->algo_flops_scale > 0.f ? d->algo_flops_scale : 1.0f;
  p.dbg = kGemmProbes ? option(OPT_GEMM_DBG) : 0;
  p.dbg_buf = nullptr;
  p.pre_alt = 0;
  p.colsum = nullptr;
  p.out16_fp16 = d->out16_is_fp16 ? 1 : 0;
  p.aux_fp16 = d->aux_is_fp16 ? 1 : 0;
  // a_format / b_format: 0 = F16, 1 = BF16 (bits 7-9 / 10-12)
  p.idesc = make_idesc_bf16(BM, bn, d->a_mn_major != 0, d->b_mn_major != 0);
  if (d->a_is_fp16) p.idesc &= ~(7u << 7);
  if (d->b_is_fp16) p.idesc &= ~(7u << 10);

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    GemmParams p2 = p;
    const int rv = launch_gemm_v2(d, p2, st);
    if (rv <= 0) return rv;
  }
  B200_CHECK_ARG(d->out_batch_period <= 0, "out_batch_period is only implemented by the bulk-store kernel (see launch_gemm_v2)");
  B200_CHECK_ARG(d->out16_colsum == nullptr, "out16_colsum is only implemented by the bulk-store kernel (16-bit output, N % 32 == 0, no split-K)");
  B200_CHECK_ARG(!(d->out16_pre_alt && d->out_bf16_pre), "out16_pre_alt is only implemented by the bulk-store kernel (N % 32 == 0, no aux, 16-byte aligned outputs)");
  CUtensorMap ta, tb;
  if (!d->a_mn_major) {
    B200_TRY(make_tensor_map_2d(&ta, d->A, (uint64_t)d->K, (uint64_t)d->M, (uint64_t)d->lda, BK, BM));
  } else {
    B200_TRY(make_tensor_map_2d(&ta, d->A, (uint64_t)d->M, (uint64_t)d->K, (uint64_t)d->lda, 64, BK));
  }
  if (!d->b_mn_major) {
    B200_TRY(make_tensor_map_2d(&tb, d->B, (uint64_t)d->K, (uint64_t)d->N, (uint64_t)d->ldb, BK, (uint32_t)bn));
  } else {
    B200_TRY(make_tensor_map_2d(&tb, d->B, (uint64_t)d->N, (uint64_t)d->K, (uint64_t)d->ldb, 64, BK));
  }
  const bool amn = d->a_mn_major != 0, bmn = d->b_mn_major != 0;
  if (bn == 256) return dispatch_major<256>(amn, bmn, ta, tb, p, st);
  if (bn == 192) return dispatch_major<192>(amn, bmn, ta, tb, p, st);
  return dispatch_major<128>(amn, bmn, ta, tb, p, st);
}

namespace b200 {
int launch_gemm_ln(const b200_gemm_desc* d, const float* x, long long ldx, const float* ln_w, const float* ln_b, float eps,
                   float* mean, float* rstd, cudaStream_t st);   // gemm_v2.cu
}

extern "C" int b200_ln_gemm_bf16(const float* x, const float* ln_w, const float* ln_b, float eps, float* mean, float* rstd,
                                 void* xn_ws, const b200_gemm_desc* d, void* stream) {
  B200_CHECK_ARG(d != nullptr && x != nullptr && ln_w != nullptr && ln_b != nullptr, "null argument");
  B200_CHECK_ARG(d->M > 0 && d->N > 0 && d->K > 0 && d->B, "empty problem");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int r = launch_gemm_ln(d, x, d->K, ln_w, ln_b, eps, mean, rstd, st);
  if (r <= 0) return r;
  // outside the fused kernel's envelope: LayerNorm pass into the caller's bf16 workspace, then the plain GEMM
  B200_CHECK_ARG(xn_ws != nullptr, "the unfused path needs the [M, K] bf16 workspace");
  B200_TRY(b200_layernorm_fwd(x, ln_w, ln_b, eps, nullptr, xn_ws, mean, rstd, d->M, d->K, 0, 0, 0, stream));
  b200_gemm_desc g = *d;
  g.A = xn_ws;
  g.lda = d->K;
  return b200_gemm_bf16(&g, stream);
}
