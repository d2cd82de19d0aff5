// Flash-style attention forward / backward for the distillation path (no [B,h,N,N] tensor is ever materialised).
//   teacher self-attention  : hub Attention.forward, head_dim 64        (via models/backbones/dinov2.py:32)
//   ScaleKD cross-attention : losses/scalekd.py:299-314, head_dim 16/24/32/48/64/96, scale = hd^-0.5 * softmax_scale
// v1 data path: cp.async double-buffered K/V (or Q/dO) tiles in shared memory, ldmatrix fragments,
// mma.sync m16n8k16 bf16 -> fp32, online softmax in the exp2 domain. Token-major strided operands so q/k/v are read
// straight out of the fused qkv / kv GEMM outputs and the output lands head-merged.
// Backward is split in two kernels (dQ; dK+dV) so that no atomics or cross-warp transposes are needed.
#include "common.cuh"
#include "../../include/b200_distill.h"

#include <stdlib.h>

namespace b200 {

constexpr int ATT_BQ = 64;   // rows per CTA (4 warps x 16)
constexpr int ATT_BK = 64;   // columns per inner tile
constexpr int ATT_THREADS = 128;
constexpr int ATT_PAD = 8;   // bf16 elements of row padding (16 B) -> conflict-free ldmatrix

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(src_bytes)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <bool F16>
__device__ __forceinline__ void mma_16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  if constexpr (F16) mma_f16(c, a, b0, b1); else mma_bf16(c, a, b0, b1);
}
__device__ __forceinline__ uint32_t pack_half(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <bool F16>
__device__ __forceinline__ uint32_t pack_16(float a, float b) {
  if constexpr (F16) return pack_half(a, b); else return pack_bf16(a, b);
}
// packed fp16x2 -> packed bf16x2 (saved forward activations are fp16; gradient MMAs run in bf16 for range)
__device__ __forceinline__ uint32_t h2_to_bf2(uint32_t u) {
  const float2 f = __half22float2(*reinterpret_cast<__half2*>(&u));
  return pack_bf16(f.x, f.y);
}

// Load a [64][HDP] bf16 tile (row pitch HDP+PAD) of rows row0.. from a token-major strided tensor; rows >= nrows and
// columns >= hd are zero filled (cp.async src-size 0).
template <int HDP>
__device__ __forceinline__ void load_tile(__nv_bfloat16* s, const __nv_bfloat16* g, int row0, int nrows, long long ts,
                                          int hd) {
  constexpr int CH = HDP / 8;
  constexpr int PITCH = HDP + ATT_PAD;
  for (int i = threadIdx.x; i < 64 * CH; i += ATT_THREADS) {
    const int r = i / CH, c = i - r * CH;
    const int gr = row0 + r;
    const bool ok = (gr < nrows) && (c * 8 < hd);
    const __nv_bfloat16* src = ok ? g + (long long)gr * ts + c * 8 : g;
    cp_async16(s + r * PITCH + c * 8, src, ok ? 16 : 0);
  }
}

// A fragments (16 rows of this warp x HDP) from a row-major smem tile.
template <int HDP>
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[HDP / 16][4], const __nv_bfloat16* s, int warp, int lane) {
  constexpr int PITCH = HDP + ATT_PAD;
#pragma unroll
  for (int kb = 0; kb < HDP / 16; ++kb)
    ldsm_x4(a[kb], s + (warp * 16 + (lane & 15)) * PITCH + kb * 16 + (lane >> 4) * 8);
}

template <int HDP>
__device__ __forceinline__ void cvt_frags_h2b(uint32_t (&a)[HDP / 16][4]) {
#pragma unroll
  for (int kb = 0; kb < HDP / 16; ++kb)
#pragma unroll
    for (int e = 0; e < 4; ++e) a[kb][e] = h2_to_bf2(a[kb][e]);
}

// acc[16 x 64] (+)= A[16 x HDP] * T^T, T: smem tile [64][HDP] row-major (T rows are the output columns).
// F16: fp16 MMA (A and T fp16). CVT: T holds fp16 but the MMA is bf16 (A is a bf16 gradient) -> convert fragments.
template <int HDP, bool F16, bool CVT>
__device__ __forceinline__ void mma_a_tT(float (&acc)[8][4], const uint32_t (&a)[HDP / 16][4],
                                         const __nv_bfloat16* t, int lane) {
  constexpr int PITCH = HDP + ATT_PAD;
  const int mi = lane >> 3, rr = lane & 7;
#pragma unroll
  for (int kb = 0; kb < HDP / 16; ++kb) {
#pragma unroll
    for (int nb = 0; nb < 8; nb += 2) {
      uint32_t b[4];
      ldsm_x4(b, t + ((nb + (mi >> 1)) * 8 + rr) * PITCH + kb * 16 + (mi & 1) * 8);
      if constexpr (CVT) {
#pragma unroll
        for (int e = 0; e < 4; ++e) b[e] = h2_to_bf2(b[e]);
      }
      mma_16<F16>(acc[nb], a[kb], b[0], b[1]);
      mma_16<F16>(acc[nb + 1], a[kb], b[2], b[3]);
    }
  }
}

// out[16 x HDP] += P[16 x 64] * T, P given as C-layout fp32 registers (converted to 16-bit A fragments), T: [64][HDP].
template <int HDP, bool F16, bool CVT>
__device__ __forceinline__ void mma_p_t(float (&out)[HDP / 8][4], const float (&p)[8][4], const __nv_bfloat16* t,
                                        int lane) {
  constexpr int PITCH = HDP + ATT_PAD;
  const int mi = lane >> 3, rr = lane & 7;
#pragma unroll
  for (int kb = 0; kb < 4; ++kb) {
    uint32_t a[4];
    a[0] = pack_16<F16>(p[2 * kb][0], p[2 * kb][1]);
    a[1] = pack_16<F16>(p[2 * kb][2], p[2 * kb][3]);
    a[2] = pack_16<F16>(p[2 * kb + 1][0], p[2 * kb + 1][1]);
    a[3] = pack_16<F16>(p[2 * kb + 1][2], p[2 * kb + 1][3]);
#pragma unroll
    for (int nb = 0; nb < HDP / 8; nb += 2) {
      uint32_t b[4];
      ldsm_x4_t(b, t + (kb * 16 + (mi & 1) * 8 + rr) * PITCH + (nb + (mi >> 1)) * 8);
      if constexpr (CVT) {
#pragma unroll
        for (int e = 0; e < 4; ++e) b[e] = h2_to_bf2(b[e]);
      }
      mma_16<F16>(out[nb], a, b[0], b[1]);
      mma_16<F16>(out[nb + 1], a, b[2], b[3]);
    }
  }
}

struct AttnParams {
  const __nv_bfloat16 *q, *k, *v, *d_o;
  const __nv_bfloat16* o_in;
  __nv_bfloat16 *o, *o_alt, *dq, *dk, *dv;   // o_alt: optional copy of o in the other 16-bit format (forward)
  float *lse, *delta;
  float *dq_colsum, *dk_colsum, *dv_colsum;   // optional fp32 [heads*hd] accumulators (bias gradients of the q/k/v projections)
  long long q_bs, q_ts, k_bs, k_ts, v_bs, v_ts, o_bs, o_ts, do_bs, do_ts, dq_bs, dq_ts, dk_bs, dk_ts, dv_bs, dv_ts;
  int B, heads, Nq, Nk, hd;
  float scale, scale_log2;
  int half;
};

// ------------------------------------------------------------------------------------------------ forward
template <int HDP, bool HALF>
__global__ void __launch_bounds__(ATT_THREADS)
attn_fwd_kernel(const AttnParams p) {
  pdl_wait();   // PDL (common.cuh): launched through launch_pdl(); multi-wave grid, so no early pdl_trigger()
  constexpr int PITCH = HDP + ATT_PAD;
  constexpr int TILE = 64 * PITCH;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sK = sQ + TILE;       // [2][TILE]
  __nv_bfloat16* sV = sK + 2 * TILE;   // [2][TILE]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int q0 = blockIdx.x * ATT_BQ, h = blockIdx.y, b = blockIdx.z;
  const __nv_bfloat16* gq = p.q + (long long)b * p.q_bs + (long long)h * p.hd;
  const __nv_bfloat16* gk = p.k + (long long)b * p.k_bs + (long long)h * p.hd;
  const __nv_bfloat16* gv = p.v + (long long)b * p.v_bs + (long long)h * p.hd;

  load_tile<HDP>(sQ, gq, q0, p.Nq, p.q_ts, p.hd);
  load_tile<HDP>(sK, gk, 0, p.Nk, p.k_ts, p.hd);
  load_tile<HDP>(sV, gv, 0, p.Nk, p.v_ts, p.hd);
  cp_async_commit();

  const int n_tiles = (p.Nk + ATT_BK - 1) / ATT_BK;
  uint32_t qa[HDP / 16][4];
  float o[HDP / 8][4];
#pragma unroll
  for (int i = 0; i < HDP / 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;   // m: running max of the raw (sign-normalised) logits
  const float sc = fabsf(p.scale_log2);

  for (int it = 0; it < n_tiles; ++it) {
    const int buf = it & 1;
    if (it + 1 < n_tiles) {
      load_tile<HDP>(sK + (buf ^ 1) * TILE, gk, (it + 1) * ATT_BK, p.Nk, p.k_ts, p.hd);
      load_tile<HDP>(sV + (buf ^ 1) * TILE, gv, (it + 1) * ATT_BK, p.Nk, p.v_ts, p.hd);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (it == 0) {
      load_a_frags<HDP>(qa, sQ, warp, lane);
      if (p.scale_log2 < 0.f) {   // s' = -s, scale' = |scale|: same softmax
#pragma unroll
        for (int kb = 0; kb < HDP / 16; ++kb)
#pragma unroll
          for (int e = 0; e < 4; ++e) qa[kb][e] ^= 0x80008000u;
      }
    }

    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f; }
    mma_a_tT<HDP, HALF, false>(s, qa, sK + buf * TILE, lane);

    // Softmax in the exp2 domain on the RAW logits: running max m (raw), p = ex2(s * sc - m * sc) as one FFMA + MUFU per
    // element (sc = |scale| * log2 e > 0; a negative scale flips the sign of the Q fragments once, above). The key mask is
    // applied only on the tile that needs it (warp-uniform). This loop is issue bound for head dims 16-32: per element
    // FMNMX + FFMA + MUFU + FADD (+ half a pack), down from FMUL + ISETP + FSEL + FMNMX + FADD + MUFU + FADD.
    const int kbase = it * ATT_BK;
    if (kbase + ATT_BK > p.Nk) {
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (kbase + nb * 8 + 2 * t4 + (e & 1) >= p.Nk) s[nb][e] = -INFINITY;
      }
    }
    float mx0 = m0, mx1 = m1;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      mx0 = fmaxf(mx0, fmaxf(s[nb][0], s[nb][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nb][2], s[nb][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float a0 = ex2_fast((m0 - mx0) * sc), a1 = ex2_fast((m1 - mx1) * sc);
    m0 = mx0; m1 = mx1;
    const float ms0 = mx0 * sc, ms1 = mx1 * sc;
    float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      s[nb][0] = ex2_fast(fmaf(s[nb][0], sc, -ms0));
      s[nb][1] = ex2_fast(fmaf(s[nb][1], sc, -ms0));
      s[nb][2] = ex2_fast(fmaf(s[nb][2], sc, -ms1));
      s[nb][3] = ex2_fast(fmaf(s[nb][3], sc, -ms1));
      rs0 += s[nb][0] + s[nb][1];
      rs1 += s[nb][2] + s[nb][3];
    }
    l0 = l0 * a0 + rs0;
    l1 = l1 * a1 + rs1;
#pragma unroll
    for (int i = 0; i < HDP / 8; ++i) { o[i][0] *= a0; o[i][1] *= a0; o[i][2] *= a1; o[i][3] *= a1; }
    mma_p_t<HDP, HALF, false>(o, s, sV + buf * TILE, lane);
    __syncthreads();
  }
  l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
  l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
  l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
  const float inv0 = 1.f / l0, inv1 = 1.f / l1;
  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  __nv_bfloat16* go = p.o + (long long)b * p.o_bs + (long long)h * p.hd;
#pragma unroll
  for (int nb = 0; nb < HDP / 8; ++nb) {
    const int col = nb * 8 + 2 * t4;
    if (col < p.hd) {
      if (r0 < p.Nq) *reinterpret_cast<uint32_t*>(go + (long long)r0 * p.o_ts + col) = pack_16<HALF>(o[nb][0] * inv0, o[nb][1] * inv0);
      if (r1 < p.Nq) *reinterpret_cast<uint32_t*>(go + (long long)r1 * p.o_ts + col) = pack_16<HALF>(o[nb][2] * inv1, o[nb][3] * inv1);
    }
  }
  if (p.o_alt != nullptr) {
    __nv_bfloat16* ga = p.o_alt + (long long)b * p.o_bs + (long long)h * p.hd;
#pragma unroll
    for (int nb = 0; nb < HDP / 8; ++nb) {
      const int col = nb * 8 + 2 * t4;
      if (col < p.hd) {
        if (r0 < p.Nq) *reinterpret_cast<uint32_t*>(ga + (long long)r0 * p.o_ts + col) = pack_16<!HALF>(o[nb][0] * inv0, o[nb][1] * inv0);
        if (r1 < p.Nq) *reinterpret_cast<uint32_t*>(ga + (long long)r1 * p.o_ts + col) = pack_16<!HALF>(o[nb][2] * inv1, o[nb][3] * inv1);
      }
    }
  }
  if (p.lse != nullptr && t4 == 0) {
    float* lse = p.lse + ((long long)b * p.heads + h) * p.Nq;
    const float ln2 = 0.6931471805599453f;
    if (r0 < p.Nq) lse[r0] = (m0 * sc + log2f(l0)) * ln2;
    if (r1 < p.Nq) lse[r1] = (m1 * sc + log2f(l1)) * ln2;
  }
}

// ------------------------------------------------------------------------------------------------ backward: delta
// delta[b, h, q] = sum_d dO[b,q,h,d] * O[b,q,h,d]
template <bool HALF>
__global__ void attn_delta_kernel(const AttnParams p) {
  pdl_wait();   // PDL (common.cuh): launched through launch_pdl(); multi-wave grid, so no early pdl_trigger()
  const long long n = (long long)p.B * p.Nq * p.heads;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int h = (int)(i % p.heads);
    const long long bq = i / p.heads;
    const int q = (int)(bq % p.Nq);
    const int b = (int)(bq / p.Nq);
    const __nv_bfloat16* a = p.d_o + (long long)b * p.do_bs + (long long)q * p.do_ts + (long long)h * p.hd;
    const __nv_bfloat16* c = p.o_in + (long long)b * p.o_bs + (long long)q * p.o_ts + (long long)h * p.hd;
    float acc = 0.f;
    for (int d = 0; d < p.hd; d += 8) {
      const uint4 ua = *reinterpret_cast<const uint4*>(a + d);
      const uint4 uc = *reinterpret_cast<const uint4*>(c + d);
      const uint32_t xa[4] = {ua.x, ua.y, ua.z, ua.w}, xc[4] = {uc.x, uc.y, uc.z, uc.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 fa = unpack_bf16(xa[e]);
        const float2 fc = HALF ? __half22float2(*reinterpret_cast<const __half2*>(&xc[e])) : unpack_bf16(xc[e]);
        acc += fa.x * fc.x + fa.y * fc.y;
      }
    }
    p.delta[((long long)b * p.heads + h) * p.Nq + q] = acc;
  }
}

// ------------------------------------------------------------------------------------------------ backward: dQ
template <int HDP, bool HALF>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dq_kernel(const AttnParams p) {
  pdl_wait();   // PDL (common.cuh): launched through launch_pdl(); multi-wave grid, so no early pdl_trigger()
  constexpr int PITCH = HDP + ATT_PAD;
  constexpr int TILE = 64 * PITCH;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sdO = sQ + TILE;
  __nv_bfloat16* sK = sdO + TILE;      // [2]
  __nv_bfloat16* sV = sK + 2 * TILE;   // [2]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int q0 = blockIdx.x * ATT_BQ, h = blockIdx.y, b = blockIdx.z;
  const __nv_bfloat16* gq = p.q + (long long)b * p.q_bs + (long long)h * p.hd;
  const __nv_bfloat16* gdo = p.d_o + (long long)b * p.do_bs + (long long)h * p.hd;
  const __nv_bfloat16* gk = p.k + (long long)b * p.k_bs + (long long)h * p.hd;
  const __nv_bfloat16* gv = p.v + (long long)b * p.v_bs + (long long)h * p.hd;

  load_tile<HDP>(sQ, gq, q0, p.Nq, p.q_ts, p.hd);
  load_tile<HDP>(sdO, gdo, q0, p.Nq, p.do_ts, p.hd);
  load_tile<HDP>(sK, gk, 0, p.Nk, p.k_ts, p.hd);
  load_tile<HDP>(sV, gv, 0, p.Nk, p.v_ts, p.hd);
  cp_async_commit();

  const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
  const float* lse = p.lse + ((long long)b * p.heads + h) * p.Nq;
  const float* dl = p.delta + ((long long)b * p.heads + h) * p.Nq;
  const float log2e = 1.4426950408889634f;
  const float lse0 = r0 < p.Nq ? lse[r0] * log2e : INFINITY;
  const float lse1 = r1 < p.Nq ? lse[r1] * log2e : INFINITY;
  const float dl0 = r0 < p.Nq ? dl[r0] : 0.f;
  const float dl1 = r1 < p.Nq ? dl[r1] : 0.f;

  const int n_tiles = (p.Nk + ATT_BK - 1) / ATT_BK;
  uint32_t qa[HDP / 16][4], doa[HDP / 16][4];
  float dq[HDP / 8][4];
#pragma unroll
  for (int i = 0; i < HDP / 8; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }

  for (int it = 0; it < n_tiles; ++it) {
    const int buf = it & 1;
    if (it + 1 < n_tiles) {
      load_tile<HDP>(sK + (buf ^ 1) * TILE, gk, (it + 1) * ATT_BK, p.Nk, p.k_ts, p.hd);
      load_tile<HDP>(sV + (buf ^ 1) * TILE, gv, (it + 1) * ATT_BK, p.Nk, p.v_ts, p.hd);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (it == 0) {
      load_a_frags<HDP>(qa, sQ, warp, lane);
      load_a_frags<HDP>(doa, sdO, warp, lane);
    }
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    }
    mma_a_tT<HDP, HALF, false>(s, qa, sK + buf * TILE, lane);
    mma_a_tT<HDP, false, HALF>(dp, doa, sV + buf * TILE, lane);
    const int kbase = it * ATT_BK;
    // P = 2^(s * scale - lse) as one FFMA + MUFU (ex2.approx, like the fused kernel); the key mask only on the tail tile
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float l = (e < 2) ? lse0 : lse1;
        const float dlt = (e < 2) ? dl0 : dl1;
        const float pv = ex2_fast(fmaf(s[nb][e], p.scale_log2, -l));
        s[nb][e] = pv * (dp[nb][e] - dlt);
      }
    }
    if (kbase + ATT_BK > p.Nk) {   // warp-uniform
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (kbase + nb * 8 + 2 * t4 + (e & 1) >= p.Nk) s[nb][e] = 0.f;
      }
    }
    mma_p_t<HDP, false, HALF>(dq, s, sK + buf * TILE, lane);
    __syncthreads();
  }
  __nv_bfloat16* gdq = p.dq + (long long)b * p.dq_bs + (long long)h * p.hd;
#pragma unroll
  for (int nb = 0; nb < HDP / 8; ++nb) {
    const int col = nb * 8 + 2 * t4;
    if (col < p.hd) {
      if (r0 < p.Nq) *reinterpret_cast<uint32_t*>(gdq + (long long)r0 * p.dq_ts + col) = pack_bf16(dq[nb][0] * p.scale, dq[nb][1] * p.scale);
      if (r1 < p.Nq) *reinterpret_cast<uint32_t*>(gdq + (long long)r1 * p.dq_ts + col) = pack_bf16(dq[nb][2] * p.scale, dq[nb][3] * p.scale);
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward: dK, dV
template <int HDP, bool HALF>
__global__ void __launch_bounds__(ATT_THREADS)
attn_bwd_dkv_kernel(const AttnParams p) {
  pdl_wait();   // PDL (common.cuh): launched through launch_pdl(); multi-wave grid, so no early pdl_trigger()
  constexpr int PITCH = HDP + ATT_PAD;
  constexpr int TILE = 64 * PITCH;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* sV = sK + TILE;
  __nv_bfloat16* sQ = sV + TILE;        // [2]
  __nv_bfloat16* sdO = sQ + 2 * TILE;   // [2]
  float* sLse = reinterpret_cast<float*>(sdO + 2 * TILE);  // [2][64]
  float* sDl = sLse + 2 * 64;                              // [2][64]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int k0 = blockIdx.x * ATT_BK, h = blockIdx.y, b = blockIdx.z;
  const __nv_bfloat16* gq = p.q + (long long)b * p.q_bs + (long long)h * p.hd;
  const __nv_bfloat16* gdo = p.d_o + (long long)b * p.do_bs + (long long)h * p.hd;
  const __nv_bfloat16* gk = p.k + (long long)b * p.k_bs + (long long)h * p.hd;
  const __nv_bfloat16* gv = p.v + (long long)b * p.v_bs + (long long)h * p.hd;
  const float* lse = p.lse + ((long long)b * p.heads + h) * p.Nq;
  const float* dl = p.delta + ((long long)b * p.heads + h) * p.Nq;
  const float log2e = 1.4426950408889634f;

  auto load_q_side = [&](int buf, int qt) {
    load_tile<HDP>(sQ + buf * TILE, gq, qt * ATT_BQ, p.Nq, p.q_ts, p.hd);
    load_tile<HDP>(sdO + buf * TILE, gdo, qt * ATT_BQ, p.Nq, p.do_ts, p.hd);
    if (threadIdx.x < 64) {
      const int r = qt * ATT_BQ + threadIdx.x;
      sLse[buf * 64 + threadIdx.x] = r < p.Nq ? lse[r] * log2e : INFINITY;
      sDl[buf * 64 + threadIdx.x] = r < p.Nq ? dl[r] : 0.f;
    }
  };

  load_tile<HDP>(sK, gk, k0, p.Nk, p.k_ts, p.hd);
  load_tile<HDP>(sV, gv, k0, p.Nk, p.v_ts, p.hd);
  load_q_side(0, 0);
  cp_async_commit();

  const int kr0 = k0 + warp * 16 + g, kr1 = kr0 + 8;
  const bool kok0 = kr0 < p.Nk, kok1 = kr1 < p.Nk;
  const int n_tiles = (p.Nq + ATT_BQ - 1) / ATT_BQ;
  uint32_t ka[HDP / 16][4], va[HDP / 16][4];
  float dk[HDP / 8][4], dv[HDP / 8][4];
#pragma unroll
  for (int i = 0; i < HDP / 8; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  }

  for (int it = 0; it < n_tiles; ++it) {
    const int buf = it & 1;
    if (it + 1 < n_tiles) {
      load_q_side(buf ^ 1, it + 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (it == 0) {
      load_a_frags<HDP>(ka, sK, warp, lane);
      load_a_frags<HDP>(va, sV, warp, lane);
      if constexpr (HALF) cvt_frags_h2b<HDP>(va);  // V meets the bf16 gradient dO
    }
    float st[8][4], dpt[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
      dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
    }
    mma_a_tT<HDP, HALF, false>(st, ka, sQ + buf * TILE, lane);     // S^T[kv, q]
    mma_a_tT<HDP, false, false>(dpt, va, sdO + buf * TILE, lane);  // dP^T[kv, q]
    float pt[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int qc = nb * 8 + 2 * t4 + (e & 1);
        const float l = sLse[buf * 64 + qc];
        const float dlt = sDl[buf * 64 + qc];
        const bool ok = (e < 2) ? kok0 : kok1;
        const float pv = ok ? ex2_fast(fmaf(st[nb][e], p.scale_log2, -l)) : 0.f;
        pt[nb][e] = pv;
        st[nb][e] = pv * (dpt[nb][e] - dlt);
      }
    }
    mma_p_t<HDP, false, false>(dv, pt, sdO + buf * TILE, lane);  // dV += P^T dO
    mma_p_t<HDP, false, HALF>(dk, st, sQ + buf * TILE, lane);    // dK += dS^T Q
    __syncthreads();
  }
  __nv_bfloat16* gdk = p.dk + (long long)b * p.dk_bs + (long long)h * p.hd;
  __nv_bfloat16* gdv = p.dv + (long long)b * p.dv_bs + (long long)h * p.hd;
#pragma unroll
  for (int nb = 0; nb < HDP / 8; ++nb) {
    const int col = nb * 8 + 2 * t4;
    if (col < p.hd) {
      if (kok0) {
        *reinterpret_cast<uint32_t*>(gdk + (long long)kr0 * p.dk_ts + col) = pack_bf16(dk[nb][0] * p.scale, dk[nb][1] * p.scale);
        *reinterpret_cast<uint32_t*>(gdv + (long long)kr0 * p.dv_ts + col) = pack_bf16(dv[nb][0], dv[nb][1]);
      }
      if (kok1) {
        *reinterpret_cast<uint32_t*>(gdk + (long long)kr1 * p.dk_ts + col) = pack_bf16(dk[nb][2] * p.scale, dk[nb][3] * p.scale);
        *reinterpret_cast<uint32_t*>(gdv + (long long)kr1 * p.dv_ts + col) = pack_bf16(dv[nb][2], dv[nb][3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ backward: fused
// One CTA per (batch, head) when every key fits one pass (Nk <= 256: the 224-pixel configurations). S and dP are
// computed ONCE per (query block, key) instead of once in each of the dQ and dK/dV kernels, and nothing is reduced
// through atomics:
//   warp w owns keys [32w, 32w+32): S^T and dP^T for its keys against the current 64-query block (mma.sync), P^T and
//   dS^T in registers, dV += P^T dO and dK += dS^T Q accumulate in registers over the query blocks;
//   dS^T of all keys is staged (bf16) in shared memory, and after one barrier the 8 warps split the 64 x hd block of
//   dQ = dS K between them (contraction over ALL keys, so every dQ block is complete and written once).
constexpr int ATTF_KEYS = 256;
constexpr int ATTF_THREADS = 256;
constexpr int ATTF_DS_PITCH = 64 + 8;

template <int HDP>
__device__ __forceinline__ void load_rows(__nv_bfloat16* s, const __nv_bfloat16* g, int rows, int row0, int nrows,
                                          long long ts, int hd, int nthreads) {
  constexpr int CH = HDP / 8;
  constexpr int PITCH = HDP + ATT_PAD;
  for (int i = threadIdx.x; i < rows * CH; i += nthreads) {
    const int r = i / CH, c = i - r * CH;
    const int gr = row0 + r;
    const bool ok = (gr < nrows) && (c * 8 < hd);
    const __nv_bfloat16* src = ok ? g + (long long)gr * ts + c * 8 : g;
    cp_async16(s + r * PITCH + c * 8, src, ok ? 16 : 0);
  }
}

// A fragments of 16 rows starting at `row0` of a row-major smem tile
template <int HDP>
__device__ __forceinline__ void load_a_frags_at(uint32_t (&a)[HDP / 16][4], const __nv_bfloat16* s, int row0, int lane) {
  constexpr int PITCH = HDP + ATT_PAD;
#pragma unroll
  for (int kb = 0; kb < HDP / 16; ++kb)
    ldsm_x4(a[kb], s + (row0 + (lane & 15)) * PITCH + kb * 16 + (lane >> 4) * 8);
}

__device__ __forceinline__ void ldsm_x2_t(uint32_t (&r)[2], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];"
               : "=r"(r[0]), "=r"(r[1])
               : "r"((uint32_t)__cvta_generic_to_shared(p)));
}

// fp16 -> bf16 copy of a [rows][HDP] tile (same padded layout); every thread converts 8 elements at a time
template <int HDP>
__device__ __forceinline__ void tile_h2b(__nv_bfloat16* dst, const __nv_bfloat16* src, int rows, int nthreads) {
  constexpr int CH = HDP / 8;
  constexpr int PITCH = HDP + ATT_PAD;
  for (int i = threadIdx.x; i < rows * CH; i += nthreads) {
    const int r = i / CH, c = i - r * CH;
    uint4 u = *reinterpret_cast<const uint4*>(src + r * PITCH + c * 8);
    u.x = h2_to_bf2(u.x); u.y = h2_to_bf2(u.y); u.z = h2_to_bf2(u.z); u.w = h2_to_bf2(u.w);
    *reinterpret_cast<uint4*>(dst + r * PITCH + c * 8) = u;
  }
}

template <int HDP, bool HALF>
__global__ void __launch_bounds__(ATTF_THREADS, HDP == 16 ? 2 : 1)
attn_bwd_fused_kernel(const AttnParams p) {
  pdl_wait();   // PDL (common.cuh): launched through launch_pdl(); multi-wave grid, so no early pdl_trigger()
  constexpr int PITCH = HDP + ATT_PAD;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_raw);          // [256][PITCH]
  __nv_bfloat16* sV = sK + ATTF_KEYS * PITCH;                               // [256][PITCH]
  __nv_bfloat16* sQ = sV + ATTF_KEYS * PITCH;                               // [2][64][PITCH]
  __nv_bfloat16* sdO = sQ + 2 * 64 * PITCH;                                 // [2][64][PITCH]
  __nv_bfloat16* sdS = sdO + 2 * 64 * PITCH;                                // [256][72]  dS^T (key-major)
  float* sLse = reinterpret_cast<float*>(sdS + ATTF_KEYS * ATTF_DS_PITCH);  // [2][64]
  float* sDl = sLse + 2 * 64;                                               // [2][64]
  // fp16 forward operands (ScaleKD projector): bf16 copies of K and of the current Q block for the gradient products
  float* sSum = sDl + 2 * 64;                                               // [3][64] column sums of dQ / dK / dV
  __nv_bfloat16* sKb = HALF ? reinterpret_cast<__nv_bfloat16*>(sSum + 3 * 64) : sK;  // [256][PITCH]
  __nv_bfloat16* sQb = HALF ? sKb + ATTF_KEYS * PITCH : sQ;                           // [64][PITCH]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3, mi = lane >> 3, rr = lane & 7;
  const int h = blockIdx.x, b = blockIdx.y;
  const __nv_bfloat16* gq = p.q + (long long)b * p.q_bs + (long long)h * p.hd;
  const __nv_bfloat16* gdo = p.d_o + (long long)b * p.do_bs + (long long)h * p.hd;
  const __nv_bfloat16* gk = p.k + (long long)b * p.k_bs + (long long)h * p.hd;
  const __nv_bfloat16* gv = p.v + (long long)b * p.v_bs + (long long)h * p.hd;
  const float* lse = p.lse + ((long long)b * p.heads + h) * p.Nq;
  const float* dl = p.delta + ((long long)b * p.heads + h) * p.Nq;
  const float log2e = 1.4426950408889634f;

  auto load_q_side = [&](int buf, int qt) {
    load_rows<HDP>(sQ + buf * 64 * PITCH, gq, 64, qt * 64, p.Nq, p.q_ts, p.hd, ATTF_THREADS);
    load_rows<HDP>(sdO + buf * 64 * PITCH, gdo, 64, qt * 64, p.Nq, p.do_ts, p.hd, ATTF_THREADS);
    if (threadIdx.x < 64) {
      const int r = qt * 64 + threadIdx.x;
      sLse[buf * 64 + threadIdx.x] = r < p.Nq ? lse[r] * log2e : INFINITY;   // +inf: P = 0 for rows past Nq
      sDl[buf * 64 + threadIdx.x] = r < p.Nq ? dl[r] : 0.f;
    }
  };

  load_rows<HDP>(sK, gk, ATTF_KEYS, 0, p.Nk, p.k_ts, p.hd, ATTF_THREADS);
  load_rows<HDP>(sV, gv, ATTF_KEYS, 0, p.Nk, p.v_ts, p.hd, ATTF_THREADS);
  load_q_side(0, 0);
  cp_async_commit();
  const bool want_sums = p.dk_colsum != nullptr;   // the three go together
  if (want_sums && threadIdx.x < 3 * 64) sSum[threadIdx.x] = 0.f;   // (ordered before use by the barriers below)

  const int key0 = warp * 32;
  float dk[2][HDP / 8][4], dv[2][HDP / 8][4];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int i = 0; i < HDP / 8; ++i) {
      dk[m][i][0] = dk[m][i][1] = dk[m][i][2] = dk[m][i][3] = 0.f;
      dv[m][i][0] = dv[m][i][1] = dv[m][i][2] = dv[m][i][3] = 0.f;
    }

  const int n_qt = (p.Nq + 63) / 64;
  const int ksteps = (p.Nk + 15) / 16;       // contraction length of the dQ product (keys)
  float dq_cs[HDP / 16][2];                  // this lane's column partials of dQ over all query blocks
#pragma unroll
  for (int i = 0; i < HDP / 16; ++i) { dq_cs[i][0] = 0.f; dq_cs[i][1] = 0.f; }
  for (int it = 0; it < n_qt; ++it) {
    const int buf = it & 1;
    if (it + 1 < n_qt) {
      load_q_side(buf ^ 1, it + 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();   // (A) Q/dO block `it` (and K/V on the first pass) have landed; every warp is past dQ of block it-1
    const __nv_bfloat16* tQ = sQ + buf * 64 * PITCH;
    const __nv_bfloat16* tdO = sdO + buf * 64 * PITCH;
    if constexpr (HALF) {
      if (it == 0) tile_h2b<HDP>(sKb, sK, ATTF_KEYS, ATTF_THREADS);
      tile_h2b<HDP>(sQb, tQ, 64, ATTF_THREADS);
      __syncthreads();
    }
    const __nv_bfloat16* tQb = HALF ? sQb : tQ;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const int krow = key0 + m * 16;
      float st[8][4], dpt[8][4];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
        dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
      }
      {
        uint32_t ka[HDP / 16][4];
        load_a_frags_at<HDP>(ka, sK, krow, lane);
        mma_a_tT<HDP, HALF, false>(st, ka, tQ, lane);        // S^T[key, q] (fp16 operands when the forward ran fp16)
      }
      {
        uint32_t va[HDP / 16][4];
        load_a_frags_at<HDP>(va, sV, krow, lane);
        if constexpr (HALF) cvt_frags_h2b<HDP>(va);          // V meets the bf16 gradient dO
        mma_a_tT<HDP, false, false>(dpt, va, tdO, lane);     // dP^T[key, q]
      }
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const float2 l2 = *reinterpret_cast<const float2*>(sLse + buf * 64 + nb * 8 + 2 * t4);
        const float2 d2 = *reinterpret_cast<const float2*>(sDl + buf * 64 + nb * 8 + 2 * t4);
        const float p0 = ex2_fast(fmaf(st[nb][0], p.scale_log2, -l2.x));
        const float p1 = ex2_fast(fmaf(st[nb][1], p.scale_log2, -l2.y));
        const float p2 = ex2_fast(fmaf(st[nb][2], p.scale_log2, -l2.x));
        const float p3 = ex2_fast(fmaf(st[nb][3], p.scale_log2, -l2.y));
        st[nb][0] = p0; st[nb][1] = p1; st[nb][2] = p2; st[nb][3] = p3;            // P^T
        dpt[nb][0] = p0 * (dpt[nb][0] - d2.x); dpt[nb][1] = p1 * (dpt[nb][1] - d2.y);   // dS^T
        dpt[nb][2] = p2 * (dpt[nb][2] - d2.x); dpt[nb][3] = p3 * (dpt[nb][3] - d2.y);
      }
      if (krow + 16 > p.Nk) {   // warp-uniform: keys past Nk (zero-filled rows) take no part
        const bool kok0 = krow + g < p.Nk, kok1 = krow + g + 8 < p.Nk;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
          if (!kok0) { st[nb][0] = st[nb][1] = 0.f; dpt[nb][0] = dpt[nb][1] = 0.f; }
          if (!kok1) { st[nb][2] = st[nb][3] = 0.f; dpt[nb][2] = dpt[nb][3] = 0.f; }
        }
      }
      mma_p_t<HDP, false, false>(dv[m], st, tdO, lane);      // dV += P^T dO
      mma_p_t<HDP, false, false>(dk[m], dpt, tQb, lane);     // dK += dS^T Q
      // stage dS^T (bf16) for the dQ product: rows = keys, columns = the 64 queries of this block
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        const int col = nb * 8 + 2 * t4;
        *reinterpret_cast<uint32_t*>(sdS + (krow + g) * ATTF_DS_PITCH + col) = pack_bf16(dpt[nb][0], dpt[nb][1]);
        *reinterpret_cast<uint32_t*>(sdS + (krow + g + 8) * ATTF_DS_PITCH + col) = pack_bf16(dpt[nb][2], dpt[nb][3]);
      }
    }
    __syncthreads();   // (B) dS^T of every key is staged
    // dQ[64 q, hd] = dS[64 q, keys] K[keys, hd]: warp -> (16-row query tile, slice of the hd columns)
    {
      constexpr int NBW = HDP / 16;            // 8-column blocks per warp (the two warps of a query tile split hd)
      const int mt = warp & 3, nb0 = (warp >> 2) * NBW;
      float dq[NBW][4];
#pragma unroll
      for (int i = 0; i < NBW; ++i) { dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f; }
#pragma unroll 4
      for (int ks = 0; ks < ksteps; ++ks) {
        uint32_t a[4];
        // A = dS (row = query, k = key), read transposed from the key-major staging tile
        ldsm_x4_t(a, sdS + (ks * 16 + (mi >> 1) * 8 + rr) * ATTF_DS_PITCH + mt * 16 + (mi & 1) * 8);
        if constexpr (NBW == 1) {
          uint32_t bfr[2];
          ldsm_x2_t(bfr, sKb + (ks * 16 + (lane & 15)) * PITCH + nb0 * 8);
          mma_bf16(dq[0], a, bfr[0], bfr[1]);
        } else {
#pragma unroll
          for (int nb = 0; nb < NBW; nb += 2) {
            uint32_t bfr[4];
            ldsm_x4_t(bfr, sKb + (ks * 16 + (mi & 1) * 8 + rr) * PITCH + (nb0 + nb + (mi >> 1)) * 8);
            mma_bf16(dq[nb], a, bfr[0], bfr[1]);
            mma_bf16(dq[nb + 1], a, bfr[2], bfr[3]);
          }
        }
      }
      const int r0 = it * 64 + mt * 16 + g, r1 = r0 + 8;
      __nv_bfloat16* gdq = p.dq + (long long)b * p.dq_bs + (long long)h * p.hd;
#pragma unroll
      for (int nb = 0; nb < NBW; ++nb) {   // (rows past Nq carry dS = 0)
        dq_cs[nb][0] += dq[nb][0] + dq[nb][2];
        dq_cs[nb][1] += dq[nb][1] + dq[nb][3];
      }
#pragma unroll
      for (int nb = 0; nb < NBW; ++nb) {
        const int col = (nb0 + nb) * 8 + 2 * t4;
        if (col < p.hd) {
          if (r0 < p.Nq) *reinterpret_cast<uint32_t*>(gdq + (long long)r0 * p.dq_ts + col) = pack_bf16(dq[nb][0] * p.scale, dq[nb][1] * p.scale);
          if (r1 < p.Nq) *reinterpret_cast<uint32_t*>(gdq + (long long)r1 * p.dq_ts + col) = pack_bf16(dq[nb][2] * p.scale, dq[nb][3] * p.scale);
        }
      }
    }
  }
  if (want_sums) {
    // bias gradients of the q / k / v projections: column sums over this head's rows, reduced over the 8 lanes that
    // share a column pair, then over the warps in shared memory, then ONE global atomic per column per CTA
    constexpr int NBW = HDP / 16;
    const int nb0 = (warp >> 2) * NBW;
#pragma unroll
    for (int nb = 0; nb < NBW; ++nb) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float v = dq_cs[nb][e];
        v += __shfl_xor_sync(0xffffffffu, v, 4); v += __shfl_xor_sync(0xffffffffu, v, 8); v += __shfl_xor_sync(0xffffffffu, v, 16);
        if (g == 0) atomicAdd(&sSum[(nb0 + nb) * 8 + 2 * t4 + e], v * p.scale);
      }
    }
#pragma unroll
    for (int nb = 0; nb < HDP / 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float vk = (dk[0][nb][e] + dk[0][nb][e + 2]) + (dk[1][nb][e] + dk[1][nb][e + 2]);
        float vv = (dv[0][nb][e] + dv[0][nb][e + 2]) + (dv[1][nb][e] + dv[1][nb][e + 2]);
        vk += __shfl_xor_sync(0xffffffffu, vk, 4); vk += __shfl_xor_sync(0xffffffffu, vk, 8); vk += __shfl_xor_sync(0xffffffffu, vk, 16);
        vv += __shfl_xor_sync(0xffffffffu, vv, 4); vv += __shfl_xor_sync(0xffffffffu, vv, 8); vv += __shfl_xor_sync(0xffffffffu, vv, 16);
        if (g == 0) {
          atomicAdd(&sSum[64 + nb * 8 + 2 * t4 + e], vk * p.scale);
          atomicAdd(&sSum[128 + nb * 8 + 2 * t4 + e], vv);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x < 3 * 64) {
      const int which = threadIdx.x >> 6, col = threadIdx.x & 63;
      if (col < p.hd) {
        float* dst = which == 0 ? p.dq_colsum : (which == 1 ? p.dk_colsum : p.dv_colsum);
        atomicAdd(dst + h * p.hd + col, sSum[threadIdx.x]);
      }
    }
  }
  __nv_bfloat16* gdk = p.dk + (long long)b * p.dk_bs + (long long)h * p.hd;
  __nv_bfloat16* gdv = p.dv + (long long)b * p.dv_bs + (long long)h * p.hd;
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    const int kr0 = key0 + m * 16 + g, kr1 = kr0 + 8;
#pragma unroll
    for (int nb = 0; nb < HDP / 8; ++nb) {
      const int col = nb * 8 + 2 * t4;
      if (col < p.hd) {
        if (kr0 < p.Nk) {
          *reinterpret_cast<uint32_t*>(gdk + (long long)kr0 * p.dk_ts + col) = pack_bf16(dk[m][nb][0] * p.scale, dk[m][nb][1] * p.scale);
          *reinterpret_cast<uint32_t*>(gdv + (long long)kr0 * p.dv_ts + col) = pack_bf16(dv[m][nb][0], dv[m][nb][1]);
        }
        if (kr1 < p.Nk) {
          *reinterpret_cast<uint32_t*>(gdk + (long long)kr1 * p.dk_ts + col) = pack_bf16(dk[m][nb][2] * p.scale, dk[m][nb][3] * p.scale);
          *reinterpret_cast<uint32_t*>(gdv + (long long)kr1 * p.dv_ts + col) = pack_bf16(dv[m][nb][2], dv[m][nb][3]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ host
static int fill_params(const b200_attn_desc* d, AttnParams& p, bool bwd) {
  B200_CHECK_ARG(d != nullptr, "null descriptor");
  B200_CHECK_ARG(d->q && d->k && d->v && d->o, "null tensor");
  B200_CHECK_ARG(d->B > 0 && d->heads > 0 && d->Nq > 0 && d->Nk > 0, "empty problem");
  B200_CHECK_ARG(d->scale != 0.f && d->scale == d->scale, "softmax scale must be a non-zero number");
  B200_CHECK_ARG(d->hd % 8 == 0 && d->hd >= 8 && d->hd <= 96, "head_dim must be a multiple of 8 in [8, 96]");
  auto al = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  B200_CHECK_ARG(al(d->q) && al(d->k) && al(d->v) && al(d->o), "tensors must be 16-byte aligned");
  B200_CHECK_ARG(d->q_ts % 8 == 0 && d->k_ts % 8 == 0 && d->v_ts % 8 == 0 && d->o_ts % 2 == 0 && d->q_bs % 8 == 0 &&
                     d->k_bs % 8 == 0 && d->v_bs % 8 == 0 && d->o_bs % 2 == 0,
                 "strides must keep 16-byte alignment");
  p.q = static_cast<const __nv_bfloat16*>(d->q); p.k = static_cast<const __nv_bfloat16*>(d->k);
  p.v = static_cast<const __nv_bfloat16*>(d->v);
  p.o = static_cast<__nv_bfloat16*>(d->o); p.o_in = static_cast<const __nv_bfloat16*>(d->o);
  p.o_alt = bwd ? nullptr : static_cast<__nv_bfloat16*>(d->o_alt);
  p.lse = d->lse; p.delta = d->delta;
  p.q_bs = d->q_bs; p.q_ts = d->q_ts; p.k_bs = d->k_bs; p.k_ts = d->k_ts; p.v_bs = d->v_bs; p.v_ts = d->v_ts;
  p.o_bs = d->o_bs; p.o_ts = d->o_ts;
  p.B = d->B; p.heads = d->heads; p.Nq = d->Nq; p.Nk = d->Nk; p.hd = d->hd;
  p.scale = d->scale; p.scale_log2 = d->scale * 1.4426950408889634f;
  p.half = d->qkvo_is_fp16 ? 1 : 0;
  if (bwd) {
    B200_CHECK_ARG(d->d_o && d->delta && d->lse && d->dq && d->dk && d->dv, "backward needs d_o, lse, delta, dq, dk, dv");
    B200_CHECK_ARG(al(d->d_o) && al(d->dq) && al(d->dk) && al(d->dv), "tensors must be 16-byte aligned");
    B200_CHECK_ARG(d->do_ts % 8 == 0 && d->do_bs % 8 == 0 && d->o_ts % 8 == 0 && d->o_bs % 8 == 0, "strides");
    B200_CHECK_ARG(d->dq_ts % 2 == 0 && d->dk_ts % 2 == 0 && d->dv_ts % 2 == 0, "strides");
    p.d_o = static_cast<const __nv_bfloat16*>(d->d_o); p.do_bs = d->do_bs; p.do_ts = d->do_ts;
    p.dq = static_cast<__nv_bfloat16*>(d->dq); p.dq_bs = d->dq_bs; p.dq_ts = d->dq_ts;
    p.dk = static_cast<__nv_bfloat16*>(d->dk); p.dk_bs = d->dk_bs; p.dk_ts = d->dk_ts;
    p.dv = static_cast<__nv_bfloat16*>(d->dv); p.dv_bs = d->dv_bs; p.dv_ts = d->dv_ts;
    B200_CHECK_ARG((d->dq_colsum != nullptr) == (d->dk_colsum != nullptr) && (d->dk_colsum != nullptr) == (d->dv_colsum != nullptr),
                   "dq_colsum / dk_colsum / dv_colsum go together");
    p.dq_colsum = d->dq_colsum; p.dk_colsum = d->dk_colsum; p.dv_colsum = d->dv_colsum;
  }
  return 0;
}

template <typename K>
static int set_smem(K kern, size_t bytes) {
  if (bytes > 48 * 1024) B200_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

template <int HDP, bool HALF>
static int launch_fwd_t(const AttnParams& p, cudaStream_t st) {
  constexpr size_t smem = size_t(5) * 64 * (HDP + ATT_PAD) * 2;
  static bool once = false;
  if (!once) { B200_TRY(set_smem(attn_fwd_kernel<HDP, HALF>, smem)); once = true; }
  dim3 grid((unsigned)cdiv(p.Nq, ATT_BQ), (unsigned)p.heads, (unsigned)p.B);
  const int prof = prof_begin(st);
  B200_CUDA_OK(launch_pdl(attn_fwd_kernel<HDP, HALF>, dim3(grid), dim3(ATT_THREADS), smem, st, p));
  prof_end(prof, st, 4.0 * p.B * p.heads * (double)p.Nq * p.Nk * p.hd, 1);
  B200_LAUNCH_OK();
  return 0;
}
template <int HDP>
static int launch_fwd(const AttnParams& p, cudaStream_t st) {
  return p.half ? launch_fwd_t<HDP, true>(p, st) : launch_fwd_t<HDP, false>(p, st);
}

// delta = rowsum(dO o O): shared by the tcgen05 and the mma.sync backward
static int launch_delta(const AttnParams& p, cudaStream_t st) {
  const long long n = (long long)p.B * p.Nq * p.heads;
  long long gd = cdiv(n, 256);
  if (gd > (long long)sm_count() * 16) gd = (long long)sm_count() * 16;
  if (p.half) B200_CUDA_OK(launch_pdl(attn_delta_kernel<true>, dim3((unsigned)gd), dim3(256), 0, st, p));
  else B200_CUDA_OK(launch_pdl(attn_delta_kernel<false>, dim3((unsigned)gd), dim3(256), 0, st, p));
  B200_LAUNCH_OK();
  return 0;
}

template <int HDP, bool HALF>
static int launch_bwd_t(const AttnParams& p, cudaStream_t st) {
  constexpr size_t smem_dq = size_t(6) * 64 * (HDP + ATT_PAD) * 2;
  constexpr size_t smem_dkv = size_t(6) * 64 * (HDP + ATT_PAD) * 2 + 4 * 64 * sizeof(float);
  static bool once = false;
  if (!once) {
    B200_TRY(set_smem(attn_bwd_dq_kernel<HDP, HALF>, smem_dq));
    B200_TRY(set_smem(attn_bwd_dkv_kernel<HDP, HALF>, smem_dkv));
    once = true;
  }
  const int prof = prof_begin(st);
  if constexpr (HDP == 16 || HDP == 32 || HDP == 64) {
    if (option(OPT_ATTN_BWD_FUSED) && p.Nk <= ATTF_KEYS) {
      constexpr size_t smem_f = size_t(2 * ATTF_KEYS + 4 * 64) * (HDP + ATT_PAD) * 2 + size_t(ATTF_KEYS) * ATTF_DS_PITCH * 2 +
                                7 * 64 * sizeof(float) + (HALF ? size_t(ATTF_KEYS + 64) * (HDP + ATT_PAD) * 2 : 0);
      static bool once_f = false;
      if (!once_f) { B200_TRY(set_smem(attn_bwd_fused_kernel<HDP, HALF>, smem_f)); once_f = true; }
      dim3 gf((unsigned)p.heads, (unsigned)p.B);
      B200_CUDA_OK(launch_pdl(attn_bwd_fused_kernel<HDP, HALF>, dim3(gf), dim3(ATTF_THREADS), smem_f, st, p));
      prof_end(prof, st, 8.0 * p.B * p.heads * (double)p.Nq * p.Nk * p.hd, 2);
      B200_LAUNCH_OK();
      return 0;
    }
  }
  dim3 gq((unsigned)cdiv(p.Nq, ATT_BQ), (unsigned)p.heads, (unsigned)p.B);
  B200_CUDA_OK(launch_pdl(attn_bwd_dq_kernel<HDP, HALF>, dim3(gq), dim3(ATT_THREADS), smem_dq, st, p));
  B200_LAUNCH_OK();
  dim3 gk((unsigned)cdiv(p.Nk, ATT_BK), (unsigned)p.heads, (unsigned)p.B);
  B200_CUDA_OK(launch_pdl(attn_bwd_dkv_kernel<HDP, HALF>, dim3(gk), dim3(ATT_THREADS), smem_dkv, st, p));
  prof_end(prof, st, 8.0 * p.B * p.heads * (double)p.Nq * p.Nk * p.hd, 2);
  B200_LAUNCH_OK();
  if (p.dk_colsum != nullptr) {   // two-kernel path: the column sums come from a pass over the bf16 outputs
    const int cols = p.heads * p.hd;
    for (int bb = 0; bb < p.B; ++bb) {
      B200_TRY(b200_colsum(p.dq + (long long)bb * p.dq_bs, 1, p.dq_ts, p.dq_colsum, p.Nq, cols, st));
      B200_TRY(b200_colsum(p.dk + (long long)bb * p.dk_bs, 1, p.dk_ts, p.dk_colsum, p.Nk, cols, st));
      B200_TRY(b200_colsum(p.dv + (long long)bb * p.dv_bs, 1, p.dv_ts, p.dv_colsum, p.Nk, cols, st));
    }
  }
  return 0;
}
template <int HDP>
static int launch_bwd(const AttnParams& p, cudaStream_t st) {
  return p.half ? launch_bwd_t<HDP, true>(p, st) : launch_bwd_t<HDP, false>(p, st);
}

}  // namespace b200

using namespace b200;

namespace b200 {
int launch_attention_tc_fwd(const b200_attn_desc* d, cudaStream_t st);   // attention_tc.cu (tcgen05, head_dim 64)
int launch_attention_tc_bwd(const b200_attn_desc* d, cudaStream_t st);   // attention_bwd_tc.cu (tcgen05, Nq / Nk <= 256)
int launch_attention_pp_fwd(const b200_attn_desc* d, cudaStream_t st);   // attention_pp.cu (tcgen05, two tiles in flight, Nk <= 256)
}

extern "C" int b200_attention_fwd(const b200_attn_desc* d, void* stream) {
  AttnParams p{};
  B200_TRY(fill_params(d, p, false));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    const int r = launch_attention_pp_fwd(d, st);
    if (r == 0) bump_stat(STAT_ATTN_PP_FWD);
    if (r <= 0) return r;
  }
  {
    const int r = launch_attention_tc_fwd(d, st);
    if (r == 0) bump_stat(STAT_ATTN_TC_FWD);
    if (r <= 0) return r;
  }
  bump_stat(STAT_ATTN_MMA_FWD);
  if (p.hd <= 16) return launch_fwd<16>(p, st);
  if (p.hd <= 32) return launch_fwd<32>(p, st);
  if (p.hd <= 48) return launch_fwd<48>(p, st);
  if (p.hd <= 64) return launch_fwd<64>(p, st);
  return launch_fwd<96>(p, st);
}

extern "C" int b200_attention_bwd(const b200_attn_desc* d, void* stream) {
  AttnParams p{};
  B200_TRY(fill_params(d, p, true));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  B200_TRY(launch_delta(p, st));
  {
    const int r = launch_attention_tc_bwd(d, st);
    if (r == 0) bump_stat(STAT_ATTN_TC_BWD);
    if (r <= 0) return r;
  }
  bump_stat(STAT_ATTN_MMA_BWD);
  if (p.hd <= 16) return launch_bwd<16>(p, st);
  if (p.hd <= 32) return launch_bwd<32>(p, st);
  if (p.hd <= 48) return launch_bwd<48>(p, st);
  if (p.hd <= 64) return launch_bwd<64>(p, st);
  return launch_bwd<96>(p, st);
}
