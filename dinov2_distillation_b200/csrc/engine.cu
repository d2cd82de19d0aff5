// Composite entry points: one C call runs a whole teacher forward, one teacher block (forward / input-gradient), or a
// whole ScaleKD AttentionProjector forward / backward.  Only kernel sequencing lives here; every kernel is in
// gemm_tcgen05.cu / attention.cu / elementwise.cu / kd_loss.cu.
#include "common.cuh"
#include "../../include/b200_distill.h"

#include <math.h>
#include <string.h>

namespace b200 {

int zero_f32(float* p, long long n, cudaStream_t st);

typedef __nv_bfloat16 bf16;

static int auto_split(int M, int N, int K) {
  const long long tiles = cdiv(M, 128) * cdiv(N, 128);
  const long long kb = cdiv(K, 64);
  long long s = cdiv((long long)sm_count(), tiles);
  if (s > kb / 4) s = kb / 4;  // at least 4 k-blocks per split
  if (s < 1) s = 1;
  return (int)s;
}

struct Gemm {
  b200_gemm_desc d;
  Gemm(const void* A, long long lda, const void* B, long long ldb, int M, int N, int K) {
    memset(&d, 0, sizeof d);
    d.A = A; d.lda = lda; d.B = B; d.ldb = ldb; d.M = M; d.N = N; d.K = K; d.split_k = 1;
  }
  Gemm& bias(const float* b) { d.bias = b; return *this; }
  Gemm& act(int a) { d.act = a; return *this; }
  Gemm& aux(const void* a, long long ld, int mode) { d.aux = a; d.ldaux = ld; d.aux_mode = mode; return *this; }
  Gemm& col_scale(const float* g) { d.col_scale = g; return *this; }
  Gemm& residual(const float* r, long long ld, int period = 0) { d.residual = r; d.ldres = ld; d.res_row_period = period; return *this; }
  Gemm& out32(float* o, long long ld) { d.out_f32 = o; d.ldo32 = ld; return *this; }
  Gemm& out16(void* o, long long ld) { d.out_bf16 = o; d.ldo16 = ld; return *this; }
  Gemm& out16_pre(void* o, long long ld) { d.out_bf16_pre = o; d.ldo16_pre = ld; return *this; }
  Gemm& colsum16(float* cs) { d.out16_colsum = cs; return *this; }
  Gemm& out16_alt(void* o, long long ld) { d.out_bf16_pre = o; d.ldo16_pre = ld; d.out16_pre_alt = 1; return *this; }
  Gemm& row_map(int period, int pad) { d.out_row_period = period; d.out_row_pad = pad; return *this; }
  Gemm& fp16_operands() { d.a_is_fp16 = 1; d.b_is_fp16 = 1; return *this; }
  Gemm& out16_fp16() { d.out16_is_fp16 = 1; return *this; }
  Gemm& aux_fp16() { d.aux_is_fp16 = 1; return *this; }
  Gemm& algo_scale(float f) { d.algo_flops_scale = f; return *this; }
  Gemm& out_batched(int period, long long stride) { d.out_batch_period = period; d.out_batch_stride = stride; return *this; }
  Gemm& accumulate(int on) { d.atomic_add = on; return *this; }
  // wgrad form: both operands token-major (contraction over rows), fp32 atomic accumulate, split over the contraction
  Gemm& wgrad() {
    d.a_mn_major = 1; d.b_mn_major = 1; d.atomic_add = 1;
    d.split_k = auto_split(d.M, d.N, d.K);
    return *this;
  }
  int run(void* stream) { return b200_gemm_bf16(&d, stream); }
  // C = epilogue(LayerNorm(x) W^T): one kernel for K <= 384, else LayerNorm into xn_ws + this GEMM (b200_ln_gemm_bf16)
  int run_ln(const float* x, const float* w, const float* b, float eps, float* mean, float* rstd, void* xn_ws, void* stream) {
    return b200_ln_gemm_bf16(x, w, b, eps, mean, rstd, xn_ws, &d, stream);
  }
};

// ================================================================================================ teacher
struct VitWs {
  bf16 *xn, *qkv, *attn, *h, *h12, *d16, *dh, *dqkv, *dattn, *dh12;
  float *t32, *dmid, *delta, *lse;
  float* dq_acc;   // fp32 dQ accumulator of the long-sequence attention backward (N > 256)
};

static void carve_block_ws(Arena& a, const b200_vit_config* c, long long M, int B, int N, bool bwd, VitWs& w) {
  const int D = c->D, F = c->F;
  memset(&w, 0, sizeof w);
  if (!bwd) {
    w.xn = a.take_n<bf16>(M * D);
    w.qkv = a.take_n<bf16>(M * 3 * D);
    w.attn = a.take_n<bf16>(M * D);
    w.h = a.take_n<bf16>(M * F);
    if (c->swiglu) w.h12 = a.take_n<bf16>(M * 2 * F);
    w.lse = a.take_n<float>((long long)B * c->heads * N);
  } else {
    w.d16 = a.take_n<bf16>(M * D);
    w.dh = a.take_n<bf16>(M * F);
    if (c->swiglu) w.dh12 = a.take_n<bf16>(M * 2 * F);
    w.t32 = a.take_n<float>(M * D);
    w.dmid = a.take_n<float>(M * D);
    w.dattn = a.take_n<bf16>(M * D);
    w.dqkv = a.take_n<bf16>(M * 3 * D);
    w.delta = a.take_n<float>((long long)B * c->heads * N);
    if (c->swiglu) w.h = a.take_n<bf16>(M * F);
    w.dq_acc = N > 256 ? a.take_n<float>(M * D) : nullptr;
  }
}

struct VitSave {
  float *mean1, *rstd1, *mean2, *rstd2, *lse, *x_mid;
  bf16 *qkv, *attn, *h_pre;  // h_pre: [M, F] pre-GELU, or [M, 2F] w12 output for SwiGLU
};

static void carve_block_save(Arena& a, const b200_vit_config* c, long long M, int B, int N, VitSave& s) {
  const int D = c->D, F = c->F;
  s.mean1 = a.take_n<float>(M); s.rstd1 = a.take_n<float>(M);
  s.mean2 = a.take_n<float>(M); s.rstd2 = a.take_n<float>(M);
  s.lse = a.take_n<float>((long long)B * c->heads * N);
  s.x_mid = a.take_n<float>(M * D);
  s.qkv = a.take_n<bf16>(M * 3 * D);
  s.attn = a.take_n<bf16>(M * D);
  s.h_pre = a.take_n<bf16>(M * (c->swiglu ? 2 * F : F));
}

static int attn_desc_self(b200_attn_desc& ad, const b200_vit_config* c, const bf16* qkv, bf16* o, float* lse, int B,
                          int N) {
  const int D = c->D;
  memset(&ad, 0, sizeof ad);
  ad.q = qkv; ad.k = qkv + D; ad.v = qkv + 2 * D;
  ad.q_bs = ad.k_bs = ad.v_bs = (long long)N * 3 * D;
  ad.q_ts = ad.k_ts = ad.v_ts = 3 * D;
  ad.o = o; ad.o_bs = (long long)N * D; ad.o_ts = D;
  ad.lse = lse;
  ad.B = B; ad.heads = c->heads; ad.Nq = N; ad.Nk = N; ad.hd = D / c->heads;
  ad.scale = 1.0f / sqrtf((float)ad.hd);
  return 0;
}

static int vit_block_fwd_impl(const b200_vit_config* c, const b200_vit_block* k, const float* x, float* y, int B, int N,
                              VitSave* sv, VitWs& w, void* stream) {
  const int D = c->D, F = c->F;
  const long long M = (long long)B * N;
  const int Mi = (int)M;
  bf16* qkv = sv ? sv->qkv : w.qkv;
  bf16* attn = sv ? sv->attn : w.attn;
  float* lse = sv ? sv->lse : w.lse;
  float* x_mid = sv ? sv->x_mid : y;
  // attention half
  B200_TRY(Gemm(w.xn, D, k->qkv_w, D, Mi, 3 * D, D).bias(k->qkv_b).out16(qkv, 3 * D)
               .run_ln(x, k->ln1_w, k->ln1_b, c->ln_eps, sv ? sv->mean1 : nullptr, sv ? sv->rstd1 : nullptr, w.xn, stream));
  b200_attn_desc ad;
  attn_desc_self(ad, c, qkv, attn, lse, B, N);
  B200_TRY(b200_attention_fwd(&ad, stream));
  B200_TRY(Gemm(attn, D, k->proj_w, D, Mi, D, D).bias(k->proj_b).col_scale(k->ls1).residual(x, D).out32(x_mid, D).run(stream));
  // MLP half
  float* mean2 = sv ? sv->mean2 : nullptr;
  float* rstd2 = sv ? sv->rstd2 : nullptr;
  if (!c->swiglu) {
    Gemm g1(w.xn, D, k->fc1_w, D, Mi, F, D);
    g1.bias(k->fc1_b).act(B200_ACT_GELU).out16(w.h, F);
    if (sv) g1.out16_pre(sv->h_pre, F);
    B200_TRY(g1.run_ln(x_mid, k->ln2_w, k->ln2_b, c->ln_eps, mean2, rstd2, w.xn, stream));
  } else {
    bf16* h12 = sv ? sv->h_pre : w.h12;
    B200_TRY(Gemm(w.xn, D, k->fc1_w, D, Mi, 2 * F, D).bias(k->fc1_b).out16(h12, 2 * F)
                 .run_ln(x_mid, k->ln2_w, k->ln2_b, c->ln_eps, mean2, rstd2, w.xn, stream));
    B200_TRY(b200_swiglu(h12, w.h, Mi, F, stream));
  }
  B200_TRY(Gemm(w.h, F, k->fc2_w, F, Mi, D, F).bias(k->fc2_b).col_scale(k->ls2).residual(x_mid, D).out32(y, D).run(stream));
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" size_t b200_vit_block_ws_bytes(const b200_vit_config* c, int B, int N) {
  if (!c) return 0;
  Arena f(nullptr, 0), b(nullptr, 0);
  VitWs w;
  carve_block_ws(f, c, (long long)B * N, B, N, false, w);
  carve_block_ws(b, c, (long long)B * N, B, N, true, w);
  return (f.used() > b.used() ? f.used() : b.used()) + 256;
}

extern "C" size_t b200_vit_block_save_bytes(const b200_vit_config* c, int B, int N) {
  if (!c) return 0;
  Arena a(nullptr, 0);
  VitSave s;
  carve_block_save(a, c, (long long)B * N, B, N, s);
  return a.used() + 256;
}

extern "C" int b200_vit_block_fwd(const b200_vit_config* c, const b200_vit_block* blk, const float* x, float* y, int B,
                                  int N, void* save, void* ws, size_t ws_bytes, void* stream) {
  B200_CHECK_ARG(c && blk && x && y && ws && B > 0 && N > 0, "bad args");
  B200_CHECK_ARG(c->D % c->heads == 0, "D must be divisible by heads");
  B200_CHECK_ARG(save == nullptr || x != y, "in-place forward cannot save activations");
  Arena a(ws, ws_bytes);
  VitWs w;
  carve_block_ws(a, c, (long long)B * N, B, N, false, w);
  B200_CHECK_ARG(a.ok(), "workspace too small");
  VitSave sv;
  if (save) {
    Arena s(save, size_t(-1) >> 1);
    carve_block_save(s, c, (long long)B * N, B, N, sv);
  }
  return vit_block_fwd_impl(c, blk, x, y, B, N, save ? &sv : nullptr, w, stream);
}

extern "C" int b200_vit_block_bwd_input(const b200_vit_config* c, const b200_vit_block* k, const float* x,
                                        const float* dy, float* dx, int B, int N, const void* save, void* ws,
                                        size_t ws_bytes, void* stream) {
  B200_CHECK_ARG(c && k && x && dy && dx && save && ws && B > 0 && N > 0, "bad args");
  B200_CHECK_ARG(k->qkv_wT && k->proj_wT && k->fc1_wT && k->fc2_wT,
                 "block has no transposed weights (not in the re-used stage range)");
  const int D = c->D, F = c->F;
  const long long M = (long long)B * N;
  const int Mi = (int)M;
  Arena a(ws, ws_bytes);
  VitWs w;
  carve_block_ws(a, c, M, B, N, true, w);
  B200_CHECK_ARG(a.ok(), "workspace too small");
  VitSave sv;
  Arena s(const_cast<void*>(save), size_t(-1) >> 1);
  carve_block_save(s, c, M, B, N, sv);

  // MLP half: y = x_mid + ls2 * fc2(act(fc1(LN2(x_mid))))     (ls2 is folded into fc2_wT)
  B200_TRY(b200_cast_f32_bf16(dy, w.d16, M * D, stream));
  if (!c->swiglu) {
    B200_TRY(Gemm(w.d16, D, k->fc2_wT, D, Mi, F, D).aux(sv.h_pre, F, B200_AUX_DGELU).out16(w.dh, F).run(stream));
    B200_TRY(Gemm(w.dh, F, k->fc1_wT, F, Mi, D, F).out32(w.t32, D).run(stream));
  } else {
    B200_TRY(Gemm(w.d16, D, k->fc2_wT, D, Mi, F, D).out16(w.dh, F).run(stream));
    B200_TRY(b200_swiglu_bwd(sv.h_pre, w.dh, w.dh12, Mi, F, stream));
    B200_TRY(Gemm(w.dh12, 2 * F, k->fc1_wT, 2 * F, Mi, D, 2 * F).out32(w.t32, D).run(stream));
  }
  B200_TRY(b200_layernorm_bwd(w.t32, sv.x_mid, k->ln2_w, sv.mean2, sv.rstd2, dy, w.dmid, w.d16, nullptr, nullptr, Mi, D,
                              stream));
  // attention half: x_mid = x + ls1 * proj(attn(qkv(LN1(x))))  (ls1 is folded into proj_wT)
  B200_TRY(Gemm(w.d16, D, k->proj_wT, D, Mi, D, D).out16(w.dattn, D).run(stream));
  b200_attn_desc ad;
  attn_desc_self(ad, c, sv.qkv, sv.attn, sv.lse, B, N);
  ad.d_o = w.dattn; ad.do_bs = (long long)N * D; ad.do_ts = D;
  ad.delta = w.delta;
  ad.dq = w.dqkv; ad.dk = w.dqkv + D; ad.dv = w.dqkv + 2 * D;
  ad.dq_bs = ad.dk_bs = ad.dv_bs = (long long)N * 3 * D;
  ad.dq_ts = ad.dk_ts = ad.dv_ts = 3 * D;
  ad.dq_accum = w.dq_acc;
  B200_TRY(b200_attention_bwd(&ad, stream));
  B200_TRY(Gemm(w.dqkv, 3 * D, k->qkv_wT, 3 * D, Mi, D, 3 * D).out32(w.t32, D).run(stream));
  B200_TRY(b200_layernorm_bwd(w.t32, x, k->ln1_w, sv.mean1, sv.rstd1, w.dmid, dx, nullptr, nullptr, nullptr, Mi, D,
                              stream));
  return 0;
}

extern "C" size_t b200_vit_forward_ws_bytes(const b200_vit_config* c, int B, int H, int W) {
  if (!c) return 0;
  const int HW = (H / 14) * (W / 14), N = HW + 1;
  Arena a(nullptr, 0);
  a.take_n<bf16>((long long)B * HW * 592);
  a.take_n<float>((long long)B * N * c->D);
  VitWs w;
  carve_block_ws(a, c, (long long)B * N, B, N, false, w);
  return a.used() + 256;
}

extern "C" int b200_vit_forward(const b200_vit_config* c, const b200_vit_block* blocks, const void* patch_w, int Kp,
                                const float* patch_b, const float* cls, const float* pos, const float* norm_w,
                                const float* norm_b, const float* img, int B, int H, int W, float* out_tokens,
                                void* ws, size_t ws_bytes, void* stream) {
  B200_CHECK_ARG(c && blocks && patch_w && patch_b && cls && pos && norm_w && norm_b && img && out_tokens && ws,
                 "null argument");
  B200_CHECK_ARG(B > 0 && H % 14 == 0 && W % 14 == 0 && H > 0 && W > 0, "image size must be a multiple of 14");
  B200_CHECK_ARG(Kp == 592, "patch weight must be padded to 592 columns");
  const int D = c->D;
  const int HW = (H / 14) * (W / 14), N = HW + 1;
  const long long M = (long long)B * N;
  Arena a(ws, ws_bytes);
  bf16* patches = a.take_n<bf16>((long long)B * HW * Kp);
  float* x = a.take_n<float>(M * D);
  VitWs w;
  carve_block_ws(a, c, M, B, N, false, w);
  B200_CHECK_ARG(a.ok(), "workspace too small");

  B200_TRY(b200_patch_im2col(img, patches, B, H, W, Kp, stream));
  // x[b*N + 1 + p, :] = patches . Wp^T + bias + pos[1 + p, :]
  B200_TRY(Gemm(patches, Kp, patch_w, Kp, B * HW, D, Kp).bias(patch_b).residual(pos + D, D, HW).out32(x, D).row_map(HW, 1).run(stream));
  B200_TRY(b200_write_cls_rows(x, cls, pos, B, N, D, stream));
  for (int l = 0; l < c->L; ++l) B200_TRY(vit_block_fwd_impl(c, &blocks[l], x, x, B, N, nullptr, w, stream));
  B200_TRY(b200_layernorm_fwd(x, norm_w, norm_b, c->ln_eps, out_tokens, nullptr, nullptr, nullptr, (int)M, D, 0, 0, 0,
                              stream));
  return 0;
}

// ================================================================================================ ScaleKD projector
// Precision policy: the projector FORWARD runs fp16 operands with fp32 accumulation -- the reference's own
// default is fp16 autocast (train.py:263) and fp16's 11-bit mantissa keeps the ReLU masks (BN->ReLU, FFN) in step with
// the fp32 oracle, which bf16's 8 bits do not. Gradients have a wide dynamic range, so every BACKWARD operand is bf16:
// every saved activation a wgrad GEMM needs is written in BOTH formats by the forward pass that produces it (tcgen05
// kind::f16 cannot mix fp16 and bf16 operands in one product; a conversion pass per wgrad was 20 launches a step).
namespace b200 {

typedef __nv_bfloat16 h16;  // storage type for fp16 buffers (2 bytes; the kernels are told which format it holds)

// token count per image of the conv input: the raw student map when the resize is fused, else the teacher grid
static inline int proj_hw_in(const b200_projector_config* c) { return c->raw_h > 0 ? c->raw_h * c->raw_w : c->HW; }
static inline int proj_windows(const b200_projector_config* c) {
  return (c->win_h > 1 || c->win_w > 1) ? (c->win_h > 0 ? c->win_h : 1) * (c->win_w > 0 ? c->win_w : 1) : 1;
}
static inline long long proj_rows_max(const b200_projector_config* c, int B) {
  const int a = proj_hw_in(c);
  return (long long)B * (a > c->HW ? a : c->HW);
}

struct ProjSave {
  h16 *z, *qsrc, *q, *kv, *o, *g, *h;   // fp16 (forward operands; q/k/v/o also feed the attention backward)
  bf16 *zb, *qsrcb, *ob, *gb, *hb;      // bf16 copies of the wgrad operands, written by the same forward passes
  bf16 *qb, *kvb;                       // bf16 copies of q / [k|v]: gradient-product operands of the tcgen05 attention backward
  bf16 *w2T, *w1T, *wpT, *wkvT, *wqT, *wcT;   // training: transposed bf16 weights for the dgrad GEMMs, made by the
                                        // forward's parameter-prep launch (the masters do not change before backward)
  bf16* xt;                             // bf16 student tokens (only the conv wgrad reads them)
  float *y, *bn_mean, *bn_rstd, *lse, *f32, *mean1, *rstd1, *u32, *mean2, *rstd2;
};

static void carve_proj_save(Arena& a, const b200_projector_config* c, int B, ProjSave& s) {
  const long long M = (long long)B * c->HW;
  const int D = c->D;
  s.xt = a.take_n<bf16>((long long)B * proj_hw_in(c) * c->Cs);
  s.y = a.take_n<float>(M * D);
  s.bn_mean = a.take_n<float>(D);
  s.bn_rstd = a.take_n<float>(D);
  s.z = a.take_n<h16>(M * D);
  s.qsrc = a.take_n<h16>(M * D);
  s.q = a.take_n<h16>(M * D);
  s.kv = a.take_n<h16>(M * 2 * D);
  s.o = a.take_n<h16>(M * D);
  s.lse = a.take_n<float>((long long)B * c->heads * c->HW);
  s.f32 = a.take_n<float>(M * D);
  s.mean1 = a.take_n<float>(M); s.rstd1 = a.take_n<float>(M);
  s.g = a.take_n<h16>(M * D);
  s.h = a.take_n<h16>(M * 4 * D);
  s.u32 = a.take_n<float>(M * D);
  s.mean2 = a.take_n<float>(M); s.rstd2 = a.take_n<float>(M);
  s.zb = a.take_n<bf16>(M * D);
  s.qsrcb = a.take_n<bf16>(M * D);
  s.ob = a.take_n<bf16>(M * D);
  s.gb = a.take_n<bf16>(M * D);
  s.hb = a.take_n<bf16>(M * 4 * D);
  const long long Dl = D;
  s.w2T = a.take_n<bf16>(4 * Dl * Dl);
  s.w1T = a.take_n<bf16>(4 * Dl * Dl);
  s.wpT = a.take_n<bf16>(Dl * Dl);
  s.wkvT = a.take_n<bf16>(2 * Dl * Dl);
  s.wqT = a.take_n<bf16>(Dl * Dl);
  s.wcT = a.take_n<bf16>(Dl * c->Cs);
  s.qb = a.take_n<bf16>(M * D);
  s.kvb = a.take_n<bf16>(M * 2 * D);
}

struct ProjFwdWs {
  h16 *wc3, *wq, *wkv, *wp, *w1, *w2;   // fp16 working copies of the fp32 master weights (wc3: 3-term split)
  h16* xt3;                             // 3-term split of the student tokens [M_in, 3Cs]
  float *bkv, *sums, *pos_t, *z32, *g32;
  float* yraw;                          // conv output at the raw resolution [M_in, D] (fused resize only)
  h16 *qtmp, *kvtmp;                    // raster-order q / [k|v] before the move to window-major rows (windows only)
  bf16 *qtmpb, *kvtmpb;                 // ... and their bf16 copies
};
static void carve_proj_fwd_ws(Arena& a, const b200_projector_config* c, int B, ProjFwdWs& w) {
  const long long M = (long long)B * c->HW;
  const long long D = c->D;
  const long long M_in = (long long)B * proj_hw_in(c);
  w.wc3 = a.take_n<h16>(3 * D * c->Cs);
  w.xt3 = a.take_n<h16>(M_in * 3 * c->Cs);
  w.yraw = c->raw_h > 0 ? a.take_n<float>(M_in * D) : nullptr;
  w.qtmp = proj_windows(c) > 1 ? a.take_n<h16>(M * D) : nullptr;
  w.kvtmp = proj_windows(c) > 1 ? a.take_n<h16>(M * 2 * D) : nullptr;
  w.qtmpb = proj_windows(c) > 1 ? a.take_n<bf16>(M * D) : nullptr;
  w.kvtmpb = proj_windows(c) > 1 ? a.take_n<bf16>(M * 2 * D) : nullptr;
  w.wq = a.take_n<h16>(D * D);
  w.wkv = a.take_n<h16>(2 * D * D);
  w.wp = a.take_n<h16>(D * D);
  w.w1 = a.take_n<h16>(4 * D * D);
  w.w2 = a.take_n<h16>(4 * D * D);
  w.bkv = a.take_n<float>(2 * D);
  w.sums = a.take_n<float>(2 * D);
  w.pos_t = a.take_n<float>((long long)c->HW * D);
  w.z32 = a.take_n<float>(M * D);
  w.g32 = a.take_n<float>(M * D);
}

struct ProjBwdWs {
  bf16 *w2T, *w1T, *wpT, *wkvT, *wqT, *wcT;
  bf16 *du16, *dh16, *df16, *do16, *dq16, *dkv16, *dqs16, *dy16;
  bf16* dyraw16;                        // adjoint-resized dy [M_in, D] (fused resize only)
  bf16 *dq16r, *dkv16r;                 // dq / [dk|dv] moved back to raster order (windows only)
  float *du32, *dg32, *df32, *dz32, *dqs32, *sums2, *dpos_t, *delta, *dxt32;
  float* dq_acc;                        // fp32 dQ accumulator of the long-sequence attention backward (> 256 tokens)
};
static void carve_proj_bwd_ws(Arena& a, const b200_projector_config* c, int B, ProjBwdWs& w) {
  const long long M = (long long)B * c->HW;
  const long long D = c->D;
  w.w2T = a.take_n<bf16>(4 * D * D);
  w.w1T = a.take_n<bf16>(4 * D * D);
  w.wpT = a.take_n<bf16>(D * D);
  w.wkvT = a.take_n<bf16>(2 * D * D);
  w.wqT = a.take_n<bf16>(D * D);
  w.wcT = a.take_n<bf16>(D * c->Cs);
  w.du16 = a.take_n<bf16>(M * D);
  w.dh16 = a.take_n<bf16>(M * 4 * D);
  w.df16 = a.take_n<bf16>(M * D);
  w.do16 = a.take_n<bf16>(M * D);
  w.dq16 = a.take_n<bf16>(M * D);
  w.dkv16 = a.take_n<bf16>(M * 2 * D);
  w.dqs16 = a.take_n<bf16>((long long)c->HW * D);
  w.dy16 = a.take_n<bf16>(M * D);
  w.du32 = a.take_n<float>(M * D);
  w.dg32 = a.take_n<float>(M * D);
  w.df32 = a.take_n<float>(M * D);
  w.dz32 = a.take_n<float>(M * D);
  w.dqs32 = a.take_n<float>((long long)c->HW * D);
  w.sums2 = a.take_n<float>(2 * D + (long long)c->HW * D);   // [sums2 | dpos_t]: adjacent, zeroed by one launch
  w.dpos_t = w.sums2 ? w.sums2 + 2 * D : nullptr;
  w.delta = a.take_n<float>((long long)B * c->heads * c->HW);
  w.dxt32 = a.take_n<float>(proj_rows_max(c, B) * c->Cs);
  w.dyraw16 = c->raw_h > 0 ? a.take_n<bf16>((long long)B * proj_hw_in(c) * D) : nullptr;
  w.dq16r = proj_windows(c) > 1 ? a.take_n<bf16>(M * D) : nullptr;
  w.dkv16r = proj_windows(c) > 1 ? a.take_n<bf16>(M * 2 * D) : nullptr;
  w.dq_acc = c->HW / proj_windows(c) > 256 ? a.take_n<float>((long long)B * (c->HW / proj_windows(c)) * D) : nullptr;
}

static int check_proj_cfg(const b200_projector_config* c, int B) {
  B200_CHECK_ARG(c != nullptr && B > 0, "bad args");
  B200_CHECK_ARG(c->D > 0 && c->Cs > 0 && c->HW > 0 && c->heads > 0, "bad config");
  if (proj_windows(c) > 1)
    B200_CHECK_ARG(c->grid_h > 0 && c->grid_w > 0 && c->grid_h * c->grid_w == c->HW && c->grid_h % c->win_h == 0 &&
                       c->grid_w % c->win_w == 0,
                   "window_shapes must tile the grid_h x grid_w token grid (grid_h * grid_w == HW)");
  if (c->raw_h > 0)
    B200_CHECK_ARG(c->raw_w > 0 && c->grid_h > 0 && c->grid_w > 0 && c->grid_h * c->grid_w == c->HW &&
                       c->grid_h <= 256 && c->grid_w <= 256,
                   "fused resize needs raw_h x raw_w and grid_h x grid_w with grid_h * grid_w == HW (<= 256 per side)");
  B200_CHECK_ARG(c->D % c->heads == 0, "teacher_dims must be divisible by num_heads");
  const int hd = c->D / c->heads;
  B200_CHECK_ARG(hd % 8 == 0 && hd <= 96, "head_dim (teacher_dims / num_heads) must be a multiple of 8 and <= 96");
  B200_CHECK_ARG(c->D % 8 == 0 && c->Cs % 8 == 0, "channel counts must be multiples of 8");
  return 0;
}

// Attention descriptor of window `win` (of n_win; one window = the whole grid when there are none). q / k / v / o are
// window-major, so a window's tokens are `tpw` consecutive rows of every image; lse of launch `win` is [B, heads, tpw].
static void proj_attn_desc(b200_attn_desc& ad, const b200_projector_config* c, const ProjSave& s, bool ext_query,
                           int B, int win = 0) {
  const int D = c->D, HW = c->HW;
  const int tpw = HW / proj_windows(c);
  const long long r0 = (long long)win * tpw;
  memset(&ad, 0, sizeof ad);
  ad.q = s.q + r0 * D; ad.q_bs = ext_query ? (long long)HW * D : 0; ad.q_ts = D;
  ad.k = s.kv + r0 * 2 * D; ad.v = s.kv + r0 * 2 * D + D;
  ad.k_bs = ad.v_bs = (long long)HW * 2 * D; ad.k_ts = ad.v_ts = 2 * D;
  ad.o = s.o + r0 * D; ad.o_bs = (long long)HW * D; ad.o_ts = D;
  ad.lse = s.lse + (long long)win * B * c->heads * tpw;
  ad.B = B; ad.heads = c->heads; ad.Nq = tpw; ad.Nk = tpw; ad.hd = D / c->heads;
  ad.scale = c->softmax_scale / sqrtf((float)ad.hd);
  ad.qkvo_is_fp16 = 1;
}

// dW[n_out, x_cols] += dY^T X: token-major bf16 gradient times the bf16 copy of a saved activation
static int wgrad_tok(const bf16* dY, long long ld_dy, const bf16* X, long long rows, int x_cols, float* dW, int n_out,
                     void* stream) {
  return Gemm(dY, ld_dy, X, x_cols, n_out, x_cols, (int)rows).out32(dW, x_cols).wgrad().run(stream);
}

}  // namespace b200

extern "C" size_t b200_projector_ws_bytes(const b200_projector_config* c, int B) {
  if (!c || B <= 0) return 0;
  Arena f(nullptr, 0), b(nullptr, 0);
  ProjFwdWs wf;
  ProjBwdWs wb;
  carve_proj_fwd_ws(f, c, B, wf);
  carve_proj_bwd_ws(b, c, B, wb);
  return (f.used() > b.used() ? f.used() : b.used()) + 256;
}

extern "C" size_t b200_projector_save_bytes(const b200_projector_config* c, int B) {
  if (!c || B <= 0) return 0;
  Arena a(nullptr, 0);
  ProjSave s;
  carve_proj_save(a, c, B, s);
  return a.used() + 256;
}

// shared student tokens of one ScaleKD (both projectors read the same preds_S): [xt bf16 [M,Cs] | xt3 fp16 [M,3Cs]]
static size_t proj_tokens_xt3_offset(const b200_projector_config* c, int B) {
  const size_t xt = (size_t)B * proj_hw_in(c) * c->Cs * 2;
  return (xt + 255) & ~size_t(255);
}

extern "C" size_t b200_projector_tokens_bytes(const b200_projector_config* c, int B) {
  if (!c || B <= 0) return 0;
  return proj_tokens_xt3_offset(c, B) + (size_t)B * proj_hw_in(c) * c->Cs * 3 * 2 + 256;
}

extern "C" int b200_projector_tokenize(const b200_projector_config* c, const float* x, int B, void* tokens, void* ws,
                                       size_t ws_bytes, void* stream) {
  B200_TRY(check_proj_cfg(c, B));
  B200_CHECK_ARG(x && tokens, "null argument");
  (void)ws; (void)ws_bytes;   // (kept in the signature: the one-pass tokenisation needs no scratch)
  const long long M = (long long)B * proj_hw_in(c);
  bf16* xt = static_cast<bf16*>(tokens);
  h16* xt3 = reinterpret_cast<h16*>(static_cast<uint8_t*>(tokens) + proj_tokens_xt3_offset(c, B));
  (void)M;
  B200_TRY(tokenize_split3(x, xt, xt3, B, c->Cs, proj_hw_in(c), stream));
  return 0;
}

static int projector_fwd_impl(const b200_projector_config* c, const b200_projector_params* p, const float* x,
                              const float* query, int B, float* out, void* save, void* ws, size_t ws_bytes,
                              const void* tokens, void* stream) {
  B200_TRY(check_proj_cfg(c, B));
  B200_CHECK_ARG(p && x && out && save && ws, "null argument");
  B200_CHECK_ARG(query != nullptr || p->query_w != nullptr, "There is no query!");
  const int D = c->D, Cs = c->Cs, HW = c->HW;
  const long long M = (long long)B * HW;
  const int Mi = (int)M;
  const bool raw = c->raw_h > 0;
  const int hw_in = proj_hw_in(c);
  const long long M_in = (long long)B * hw_in;
  const bool ext = query != nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Arena sa(save, size_t(-1) >> 1);
  ProjSave s;
  carve_proj_save(sa, c, B, s);
  Arena wa(ws, ws_bytes);
  ProjFwdWs w;
  carve_proj_fwd_ws(wa, c, B, w);
  B200_CHECK_ARG(wa.ok(), "workspace too small");

  // per-step working copies of the fp32 master parameters, one launch: fp16 casts, 3-term split of the conv weight,
  // [k_b | v_b], pos_embed -> token-major
  {
    PrepJobs jobs{};
    jobs.add(PREP_SPLIT3_RIGHT, p->conv_w, w.wc3, D, Cs, 1);
    jobs.add(PREP_CAST16, p->q_w, w.wq, D, D, 1);
    jobs.add(PREP_CAST16, p->k_w, w.wkv, D, D, 1);
    jobs.add(PREP_CAST16, p->v_w, w.wkv + (long long)D * D, D, D, 1);
    jobs.add(PREP_CAST16, p->p_w, w.wp, D, D, 1);
    jobs.add(PREP_CAST16, p->ffn1_w, w.w1, 4 * D, D, 1);
    jobs.add(PREP_CAST16, p->ffn2_w, w.w2, D, 4 * D, 1);
    jobs.add(PREP_COPY32, p->k_b, w.bkv, 1, D);
    jobs.add(PREP_COPY32, p->v_b, w.bkv + D, 1, D);
    jobs.add(PREP_TRANSPOSE32, p->pos_embed, w.pos_t, D, HW, 0, D);   // [D, HW] -> [HW, D]
    if (c->training) {   // the backward's transposed bf16 weights ride along (one launch less per projector and step)
      jobs.add(PREP_TRANSPOSE16, p->ffn2_w, s.w2T, D, 4 * D, 0, D);        // [D,4D] -> [4D, D]
      jobs.add(PREP_TRANSPOSE16, p->ffn1_w, s.w1T, 4 * D, D, 0, 4 * D);    // [4D,D] -> [D, 4D]
      jobs.add(PREP_TRANSPOSE16, p->p_w, s.wpT, D, D, 0, D);
      jobs.add(PREP_TRANSPOSE16, p->k_w, s.wkvT, D, D, 0, 2 * D);          // [D, 2D]: [Wk^T | Wv^T]
      jobs.add(PREP_TRANSPOSE16, p->v_w, s.wkvT + D, D, D, 0, 2 * D);
      jobs.add(PREP_TRANSPOSE16, p->q_w, s.wqT, D, D, 0, D);
      jobs.add(PREP_TRANSPOSE16, p->conv_w, s.wcT, D, Cs, 0, D);           // [D,Cs] -> [Cs, D]
    }
    B200_TRY(launch_param_prep(jobs, st));
  }

  // proj_student: conv1x1 -> BN -> ReLU, + pos_embed            (losses/scalekd.py:199-201, :238)
  // the conv feeds BN -> ReLU with no residual around it: run it as a 3-term split fp16 product (K -> 3K)
  const h16* xt3 = w.xt3;
  if (tokens != nullptr) {
    xt3 = reinterpret_cast<const h16*>(static_cast<const uint8_t*>(tokens) + proj_tokens_xt3_offset(c, B));
  } else {
    B200_TRY(tokenize_split3(x, s.xt, w.xt3, B, Cs, hw_in, stream));
  }
  // fused ModelWrapper resize (models/model_zoo.py:121-126): conv at the raw resolution, then resize its output
  B200_TRY(Gemm(xt3, 3 * Cs, w.wc3, 3 * Cs, (int)M_in, D, 3 * Cs).fp16_operands().algo_scale(1.f / 3.f).bias(p->conv_b).out32(raw ? w.yraw : s.y, D).run(stream));
  if (raw) B200_TRY(b200_bilinear_tokens_fwd(w.yraw, s.y, B, c->raw_h, c->raw_w, c->grid_h, c->grid_w, D, stream));
  if (c->training) {
    B200_TRY(zero_f32(w.sums, 2 * D, st));
    B200_TRY(b200_bn_stats(s.y, w.sums, Mi, D, stream));
    B200_TRY(bn_finalize_counted(w.sums, s.bn_mean, s.bn_rstd, p->bn_running_mean, p->bn_running_var, c->bn_momentum,
                                 c->bn_eps, Mi, D, p->bn_num_batches_tracked, stream));
  } else {
    B200_TRY(b200_bn_finalize(nullptr, s.bn_mean, s.bn_rstd, p->bn_running_mean, p->bn_running_var, c->bn_momentum,
                              c->bn_eps, Mi, D, stream));
  }
  B200_TRY(bn_relu_pos_fwd_dual(s.y, s.bn_mean, s.bn_rstd, p->bn_w, p->bn_b, w.pos_t, w.z32, s.z, s.zb, Mi, D, HW, 1, stream));

  // cross attention: q from the query tokens, k/v from the student tokens      (losses/scalekd.py:299-316)
  const int Mq = ext ? Mi : HW;
  B200_TRY(cast_f32_f16_dual(ext ? query : p->query_w, s.qsrc, s.qsrcb, (long long)Mq * D, stream));
  const int n_win = proj_windows(c);
  // (q and [k|v] are also stored as bf16 -- the gradient products of the attention backward read those)
  {
    Gemm gq(s.qsrc, D, w.wq, D, Mq, D, D);
    gq.fp16_operands().bias(p->q_b).out16(n_win > 1 ? w.qtmp : s.q, D).out16_fp16();
    gq.out16_alt(n_win > 1 ? w.qtmpb : s.qb, D);
    B200_TRY(gq.run(stream));
    Gemm gkv(s.z, D, w.wkv, D, Mi, 2 * D, D);
    gkv.fp16_operands().bias(w.bkv).out16(n_win > 1 ? w.kvtmp : s.kv, 2 * D).out16_fp16();
    gkv.out16_alt(n_win > 1 ? w.kvtmpb : s.kvb, 2 * D);
    B200_TRY(gkv.run(stream));
  }
  if (n_win > 1) {   // separate_tokens (losses/scalekd.py:305-308): each window's tokens become consecutive rows
    B200_TRY(b200_window_rows16(w.qtmp, s.q, Mq, c->grid_h, c->grid_w, c->win_h, c->win_w, D, D, 0, stream));
    B200_TRY(b200_window_rows16(w.kvtmp, s.kv, M, c->grid_h, c->grid_w, c->win_h, c->win_w, 2 * D, 2 * D, 0, stream));
    B200_TRY(b200_window_rows16(w.qtmpb, s.qb, Mq, c->grid_h, c->grid_w, c->win_h, c->win_w, D, D, 0, stream));
    B200_TRY(b200_window_rows16(w.kvtmpb, s.kvb, M, c->grid_h, c->grid_w, c->win_h, c->win_w, 2 * D, 2 * D, 0, stream));
  }
  for (int win = 0; win < n_win; ++win) {   // (the output stays window-major, as in the reference: scalekd.py:313-314)
    b200_attn_desc ad;
    proj_attn_desc(ad, c, s, ext, B, win);
    ad.o_alt = s.ob + (long long)win * (HW / n_win) * D;
    B200_TRY(b200_attention_fwd(&ad, stream));
  }
  B200_TRY(Gemm(s.o, D, w.wp, D, Mi, D, D).fp16_operands().bias(p->p_b).residual(w.z32, D).out32(s.f32, D).run(stream));

  // norm -> FFN(ReLU) + residual -> norm_2                                      (losses/scalekd.py:243-245)
  B200_TRY(layernorm_fwd_dual(s.f32, p->ln1_w, p->ln1_b, c->ln_eps, w.g32, s.g, s.gb, s.mean1, s.rstd1, Mi, D, 0, 0, 1, stream));
  B200_TRY(Gemm(s.g, D, w.w1, D, Mi, 4 * D, D).fp16_operands().bias(p->ffn1_b).act(B200_ACT_RELU).out16(s.h, 4 * D).out16_fp16().out16_alt(s.hb, 4 * D).run(stream));
  B200_TRY(Gemm(s.h, 4 * D, w.w2, 4 * D, Mi, D, 4 * D).fp16_operands().bias(p->ffn2_b).residual(w.g32, D).out32(s.u32, D).run(stream));
  B200_TRY(b200_layernorm_fwd(s.u32, p->ln2_w, p->ln2_b, c->ln_eps, out, nullptr, s.mean2, s.rstd2, Mi, D, 0, 0, 0, stream));
  return 0;
}

extern "C" int b200_projector_fwd(const b200_projector_config* c, const b200_projector_params* p, const float* x,
                                  const float* query, int B, float* out, void* save, void* ws, size_t ws_bytes,
                                  void* stream) {
  return projector_fwd_impl(c, p, x, query, B, out, save, ws, ws_bytes, nullptr, stream);
}

extern "C" int b200_projector_fwd_tok(const b200_projector_config* c, const b200_projector_params* p, const float* x,
                                      const float* query, int B, float* out, void* save, void* ws, size_t ws_bytes,
                                      const void* tokens, void* stream) {
  return projector_fwd_impl(c, p, x, query, B, out, save, ws, ws_bytes, tokens, stream);
}

static int projector_bwd_impl(const b200_projector_config* c, const b200_projector_params* p,
                              const b200_projector_grads* g, const float* x, const float* query, const float* dout,
                              int B, float* dx, int dx_accumulate, float* dquery, const void* save, void* ws,
                              size_t ws_bytes, const void* tokens, void* stream) {
  B200_TRY(check_proj_cfg(c, B));
  B200_CHECK_ARG(p && g && dout && save && ws, "null argument");
  (void)x;
  const int D = c->D, Cs = c->Cs, HW = c->HW;
  const long long M = (long long)B * HW;
  const int Mi = (int)M;
  const bool ext = query != nullptr;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Arena sa(const_cast<void*>(save), size_t(-1) >> 1);
  ProjSave s;
  carve_proj_save(sa, c, B, s);
  Arena wa(ws, ws_bytes);
  ProjBwdWs w;
  carve_proj_bwd_ws(wa, c, B, w);
  B200_CHECK_ARG(wa.ok(), "workspace too small");

  // transposed bf16 weights for the dgrad GEMMs: made by the forward's prep launch when training, else here
  if (c->training) {
    w.w2T = s.w2T; w.w1T = s.w1T; w.wpT = s.wpT; w.wkvT = s.wkvT; w.wqT = s.wqT; w.wcT = s.wcT;
  } else {
    PrepJobs jobs{};
    jobs.add(PREP_TRANSPOSE16, p->ffn2_w, w.w2T, D, 4 * D, 0, D);        // [D,4D] -> [4D, D]
    jobs.add(PREP_TRANSPOSE16, p->ffn1_w, w.w1T, 4 * D, D, 0, 4 * D);    // [4D,D] -> [D, 4D]
    jobs.add(PREP_TRANSPOSE16, p->p_w, w.wpT, D, D, 0, D);
    jobs.add(PREP_TRANSPOSE16, p->k_w, w.wkvT, D, D, 0, 2 * D);          // [D, 2D]: [Wk^T | Wv^T]
    jobs.add(PREP_TRANSPOSE16, p->v_w, w.wkvT + D, D, D, 0, 2 * D);
    jobs.add(PREP_TRANSPOSE16, p->q_w, w.wqT, D, D, 0, D);
    jobs.add(PREP_TRANSPOSE16, p->conv_w, w.wcT, D, Cs, 0, D);           // [D,Cs] -> [Cs, D]
    B200_TRY(launch_param_prep(jobs, st));
  }

  // norm_2
  // (the column sums of du are the ffn2 bias gradient: produced by the same pass)
  B200_TRY(b200_layernorm_bwd_colsum(dout, s.u32, p->ln2_w, s.mean2, s.rstd2, nullptr, w.du32, w.du16, g->ln2_w, g->ln2_b,
                                     g->ffn2_b, Mi, D, stream));
  // FFN: u = g + W2 relu(W1 g + b1) + b2
  B200_TRY(wgrad_tok(w.du16, D, s.hb, M, 4 * D, g->ffn2_w, D, stream));
  // (ffn1's bias gradient = column sums of dh: accumulated by the same epilogue)
  B200_TRY(Gemm(w.du16, D, w.w2T, D, Mi, 4 * D, D).aux(s.h, 4 * D, B200_AUX_DRELU).aux_fp16().out16(w.dh16, 4 * D).colsum16(g->ffn1_b).run(stream));
  B200_TRY(wgrad_tok(w.dh16, 4 * D, s.gb, M, D, g->ffn1_w, 4 * D, stream));
  B200_TRY(Gemm(w.dh16, 4 * D, w.w1T, 4 * D, Mi, D, 4 * D).residual(w.du32, D).out32(w.dg32, D).run(stream));
  // norm
  B200_TRY(b200_layernorm_bwd_colsum(w.dg32, s.f32, p->ln1_w, s.mean1, s.rstd1, nullptr, w.df32, w.df16, g->ln1_w, g->ln1_b,
                                     g->p_b, Mi, D, stream));
  // attention output projection: f = Wp o + bp + z   (p_b gradient = column sums of df, fused above)
  B200_TRY(wgrad_tok(w.df16, D, s.ob, M, D, g->p_w, D, stream));
  B200_TRY(Gemm(w.df16, D, w.wpT, D, Mi, D, D).out16(w.do16, D).run(stream));
  // attention core (q/k/v/o fp16 from the forward; gradients bf16), one launch per window
  // q / k / v bias gradients = column sums of dq / dk / dv over every (image, token): produced inside the attention
  // backward (for a batch-invariant self query the sum over images of dq is the same quantity)
  const int n_win = proj_windows(c);
  for (int win = 0; win < n_win; ++win) {
    const int tpw = HW / n_win;
    const long long r0 = (long long)win * tpw;
    b200_attn_desc ad;
    proj_attn_desc(ad, c, s, ext, B, win);
    ad.d_o = w.do16 + r0 * D; ad.do_bs = (long long)HW * D; ad.do_ts = D;
    ad.delta = w.delta + (long long)win * B * c->heads * tpw;
    ad.dq = w.dq16 + r0 * D; ad.dq_bs = (long long)HW * D; ad.dq_ts = D;
    ad.dk = w.dkv16 + r0 * 2 * D; ad.dv = w.dkv16 + r0 * 2 * D + D;
    ad.dk_bs = ad.dv_bs = (long long)HW * 2 * D; ad.dk_ts = ad.dv_ts = 2 * D;
    ad.dq_colsum = g->q_b; ad.dk_colsum = g->k_b; ad.dv_colsum = g->v_b;
    ad.q_alt = s.qb + r0 * D; ad.k_alt = s.kvb + r0 * 2 * D; ad.v_alt = s.kvb + r0 * 2 * D + D;
    ad.dq_accum = w.dq_acc;
    B200_TRY(b200_attention_bwd(&ad, stream));
  }
  const bf16* dq16 = w.dq16;
  const bf16* dkv16 = w.dkv16;
  if (n_win > 1) {   // gradients of the window-major q / k / v back to the raster order their projections produced
    B200_TRY(b200_window_rows16(w.dq16, w.dq16r, M, c->grid_h, c->grid_w, c->win_h, c->win_w, D, D, 1, stream));
    B200_TRY(b200_window_rows16(w.dkv16, w.dkv16r, M, c->grid_h, c->grid_w, c->win_h, c->win_w, 2 * D, 2 * D, 1, stream));
    dq16 = w.dq16r;
    dkv16 = w.dkv16r;
  }
  // q path
  if (ext) {
    B200_TRY(wgrad_tok(dq16, D, s.qsrcb, M, D, g->q_w, D, stream));
    if (dquery) B200_TRY(Gemm(dq16, D, w.wqT, D, Mi, D, D).out32(dquery, D).run(stream));
  } else {
    B200_TRY(b200_batch_sum_bf16(dq16, w.dqs32, w.dqs16, B, (long long)HW * D, stream));
    B200_TRY(wgrad_tok(w.dqs16, D, s.qsrcb, HW, D, g->q_w, D, stream));
    if (g->query_w)
      B200_TRY(Gemm(w.dqs16, D, w.wqT, D, HW, D, D).residual(g->query_w, D).out32(g->query_w, D).run(stream));
  }
  // k / v path
  B200_TRY(wgrad_tok(dkv16, 2 * D, s.zb, M, D, g->k_w, D, stream));
  B200_TRY(wgrad_tok(dkv16 + D, 2 * D, s.zb, M, D, g->v_w, D, stream));
  B200_TRY(Gemm(dkv16, 2 * D, w.wkvT, 2 * D, Mi, D, 2 * D).residual(w.df32, D).out32(w.dz32, D).run(stream));
  // BN + ReLU + pos_embed
  B200_TRY(zero_f32(w.sums2, 2 * D + (long long)HW * D, st));   // sums2 and dpos_t
  B200_TRY(b200_bn_relu_pos_bwd_reduce(w.dz32, s.y, s.bn_mean, s.bn_rstd, p->bn_w, p->bn_b, w.sums2, w.dpos_t, Mi, D, HW, stream));
  B200_TRY(b200_tokens_to_nchw(w.dpos_t, g->pos_embed, 1, D, HW, 1, stream));
  // (BN affine gradients = sums2: added to g->bn_b / g->bn_w by the apply pass)
  B200_TRY(bn_relu_pos_bwd_apply_acc(w.dz32, s.y, s.bn_mean, s.bn_rstd, p->bn_w, p->bn_b, w.sums2, w.dy16, c->training ? 1 : 0, Mi, D,
                                     g->bn_b, g->bn_w, stream));
  // conv1x1
  // with batch statistics the column sums of dy vanish identically (sum_r yhat = 0): BN cancels the conv bias
  // (losses/scalekd.py:199-200), so its gradient is exactly zero; only the running-statistics (eval) path needs the sum.
  if (!c->training) B200_TRY(b200_colsum(w.dy16, 1, D, g->conv_b, Mi, D, stream));
  // fused resize: the conv saw the raw map, so its gradients need R^T dy (adjoint of the bilinear resize)
  const bf16* dyc = w.dy16;
  const int hw_in = proj_hw_in(c);
  const long long M_in = (long long)B * hw_in;
  if (c->raw_h > 0) {
    B200_TRY(b200_bilinear_tokens_bwd(w.dy16, w.dyraw16, B, c->raw_h, c->raw_w, c->grid_h, c->grid_w, D, stream));
    dyc = w.dyraw16;
  }
  B200_TRY(Gemm(dyc, D, tokens ? static_cast<const bf16*>(tokens) : s.xt, Cs, D, Cs, (int)M_in).out32(g->conv_w, Cs).wgrad().run(stream));
  if (dx) {
    if (hw_in % 32 == 0 && Cs % 8 == 0) {
      // dX in NCHW straight out of the GEMM: dx[b] [Cs, HW] = Wc^T [Cs, D] . dy_b^T -- operand roles swapped, columns
      // (b, hw) batched with period HW through a 3-D output tensor map; no token-major intermediate, no transpose
      B200_TRY(Gemm(w.wcT, D, dyc, D, Cs, (int)M_in, D).out32(dx, hw_in).out_batched(hw_in, (long long)Cs * hw_in).accumulate(dx_accumulate ? 1 : 0).run(stream));
    } else {
      B200_TRY(Gemm(dyc, D, w.wcT, D, (int)M_in, Cs, D).out32(w.dxt32, Cs).run(stream));
      B200_TRY(b200_tokens_to_nchw(w.dxt32, dx, B, Cs, hw_in, dx_accumulate, stream));
    }
  }
  return 0;
}

extern "C" int b200_projector_bwd(const b200_projector_config* c, const b200_projector_params* p,
                                  const b200_projector_grads* g, const float* x, const float* query, const float* dout,
                                  int B, float* dx, int dx_accumulate, float* dquery, const void* save, void* ws,
                                  size_t ws_bytes, void* stream) {
  return projector_bwd_impl(c, p, g, x, query, dout, B, dx, dx_accumulate, dquery, save, ws, ws_bytes, nullptr, stream);
}

extern "C" int b200_projector_bwd_tok(const b200_projector_config* c, const b200_projector_params* p,
                                      const b200_projector_grads* g, const float* x, const float* query,
                                      const float* dout, int B, float* dx, int dx_accumulate, float* dquery,
                                      const void* save, void* ws, size_t ws_bytes, const void* tokens, void* stream) {
  return projector_bwd_impl(c, p, g, x, query, dout, B, dx, dx_accumulate, dquery, save, ws, ws_bytes, tokens, stream);
}
