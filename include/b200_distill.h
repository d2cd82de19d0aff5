/*
 * b200_distill.h -- C ABI of libb200distill.so: hand-written sm_100a kernels for the distillation hot path of
 * ardaerendogru/dinov2_distillation (frozen DINOv2 ViT teacher forward + ScaleKD loss forward/backward).
 *
 * The reference has NO FFI of its own (it is 100% Python); its plug-in surface is Python duck typing:
 *   - teacher object   : models/backbones/dinov2.py:5-46  (DINOv2ViT.forward -> {'feature_map'}),
 *                        .model.blocks[i] used at train/distillation_module.py:169-177
 *   - loss object      : losses/scalekd.py:12-127 (ScaleKD), registered at train/distillation_module.py:11-13
 * The Python shells in dinov2_distillation_b200/{teacher,scalekd}.py mirror those two objects and call ONLY the entry
 * points below (ctypes, raw device pointers + sizes + a cudaStream_t passed as void*). No torch types cross this line.
 *
 * Conventions
 *   - every function returns 0 on success; on failure a negative code, message via b200_last_error().
 *   - all pointers are device pointers unless stated; `stream` is a cudaStream_t.
 *   - bf16 tensors are passed as void* (uint16 storage); fp32 as float*.
 *   - nothing allocates device memory: callers pass workspaces; sizes come from the *_bytes() queries.
 *   - "tokens" are row-major [rows, D] matrices (token-major), rows = batch * tokens.
 */
#ifndef B200_DISTILL_H
#define B200_DISTILL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_ABI_VERSION 3

const char* b200_last_error(void);
int b200_abi_version(void);
/* number of kernels this library has launched since load / since the last reset (bench.py's gpu_launches). */
long long b200_launch_count(void);
void b200_reset_launch_count(void);
/* Per-launch device timing of the dense kernels (CUDA events on the launch stream), for bench.py's roofline object.
 * Categories: 0 = tcgen05 GEMM (2MNK flops), 1 = attention forward (4 Nq Nk D), 2 = attention backward (8 Nq Nk D).
 * b200_profile_read synchronises the recorded events, fills sums per category and clears the records. */
void b200_profile_enable(int on);
int b200_profile_read(int n_cat, double* ms, double* flops, long long* launches);
/* what an event pair around ONE launch adds to its measured duration: a null kernel timed bracketed launch by launch (as
 * the records above are) and back to back; per-launch microseconds, the difference is the overhead */
int b200_profile_event_overhead(int reps, float* bracketed_us, float* back_to_back_us, void* stream);
/* the same records one by one, in launch order (returns how many were written, at most cap; clears them); bytes (may be
 * NULL): the algorithmic operand + result traffic of the launch, for the roofline that binds each shape */
int b200_profile_read_records(int cap, float* ms, double* flops, int* cat, double* bytes);
/* Dispatch options (ints; defaults select the production kernels). The launchers read this table, never the
 * environment; tests and A/B tools flip entries. Names: "pdl", "attn_tc_fwd", "attn_tc_fwd_long", "attn_tc_bwd",
 * "attn_bwd_fused", "attn_tc_bwd_long", "gemm_ln". b200_set_option returns -1 for an unknown name; b200_get_option returns -1 likewise. */
int b200_set_option(const char* name, int value);
int b200_get_option(const char* name);

/* ------------------------------------------------------------------------------------------------ GEMM (tcgen05)
 * C[M,N] = epilogue(A * B^T): bf16 operands staged by TMA (128B swizzle), tcgen05.mma with fp32 accumulators in TMEM.
 * Replaces every nn.Linear / 1x1 conv / patch-embed conv on the path:
 *   teacher qkv/proj/fc1/fc2/w12/w3 (facebookresearch/dinov2 layers, called via models/backbones/dinov2.py:32),
 *   ScaleKD proj_student conv1x1 (losses/scalekd.py:199), q/k/v/proj (losses/scalekd.py:277-282), FFN (:453-460),
 *   and their dgrad / wgrad in backward.
 * K-major operand  : X[r*ld + k]      (row r of the output dimension, contiguous along the contraction)
 * MN-major operand : X[k*ld + r]      (contiguous along the output dimension) -- used by wgrad (dW = dY^T X).
 * epilogue, per element (r,n), in this order:
 *   v = acc; v += bias[n]; v = act(v); v *= f(aux[r,n]); v *= col_scale[n]; v += residual[rr,n]; store.
 */
enum { B200_ACT_NONE = 0, B200_ACT_GELU = 1, B200_ACT_RELU = 2 };
enum { B200_AUX_NONE = 0, B200_AUX_DGELU = 1, B200_AUX_DRELU = 2 };

typedef struct b200_gemm_desc {
  const void* A; long long lda; int a_mn_major;
  const void* B; long long ldb; int b_mn_major;
  int M, N, K;
  int split_k;                 /* >=1; >1 requires atomic fp32 output and no act/aux/bf16 output */
  const float* bias;           /* [N] or NULL */
  int act;                     /* B200_ACT_* */
  const void* aux; long long ldaux; int aux_mode; /* bf16 [M,N] */
  const float* col_scale;      /* [N] or NULL (LayerScale gamma) */
  const float* residual; long long ldres; int res_row_period; /* fp32; row = r % period when period > 0 */
  float* out_f32; long long ldo32; int atomic_add;
  void* out_bf16; long long ldo16;
  void* out_bf16_pre; long long ldo16_pre; /* optional bf16 copy of (acc+bias) before act */
  int out_row_period, out_row_pad; /* period>0: out row = (r/period)*(period+pad) + pad + r%period */
  /* 16-bit element types: 0 = bf16 (default), 1 = fp16. The ScaleKD projector FORWARD runs fp16 operands (the
   * reference's own default is fp16 AMP, train.py:263); gradients stay bf16 for range. A and B may differ. */
  int a_is_fp16, b_is_fp16, out16_is_fp16, aux_is_fp16;
  /* profiling only: algorithmic flops = algo_flops_scale * 2MNK (0 means 1; 1/3 for the 3-term split product) */
  float algo_flops_scale;
  /* batched-column fp32 output: period P > 0 sends element (r, n) to out_f32[(n / P) * out_batch_stride + r * ldo32 +
   * n % P] -- with the operand roles swapped (A = W^T, B = token-major dY) the input gradient of the 1x1 conv lands
   * directly in NCHW (P = H*W, ldo32 = H*W, stride = C*H*W) with no transpose pass. P % 32 == 0, N % P == 0. */
  int out_batch_period; long long out_batch_stride;
  /* 1: out_bf16_pre receives the value AFTER the activation, in the 16-bit format out_bf16 does not use -- the ScaleKD
   * FFN saves relu(W1 g + b1) as fp16 (forward operand of W2) and as bf16 (operand of the W2 wgrad GEMM) from one
   * epilogue (tcgen05 kind::f16 cannot mix the formats in one product). Needs N % 32 == 0 and no aux operand. */
  int out16_pre_alt;
  /* optional fp32 [N]: += the column sums (over the M rows) of the 16-bit output exactly as stored -- the bias gradient
   * that belongs to an input-gradient GEMM (dh = (du W2) * relu'(h): ffn1's bias gradient is colsum(dh)), produced in
   * the epilogue instead of a second pass over the output. 16-bit output, N % 32 == 0, split_k == 1. */
  float* out16_colsum;
} b200_gemm_desc;

int b200_gemm_bf16(const b200_gemm_desc* d, void* stream);
/* LayerNorm in front of a GEMM: C = epilogue(LN(x) W^T), x fp32 [M, K] contiguous rows (K = the model width), the
 * pre-norm linears of hub Block.forward (norm1 -> attn.qkv, norm2 -> mlp.fc1; facebookresearch/dinov2 layers/block.py,
 * reached through models/backbones/dinov2.py:32 and train/distillation_module.py:177). d->A / d->lda are ignored.
 * K <= 384 with a 16-bit output runs ONE kernel (the panel is normalised into shared memory, no bf16 round trip);
 * anything else runs b200_layernorm_fwd into xn_ws (bf16 [M, K]) followed by b200_gemm_bf16. mean / rstd: optional
 * fp32 [M] outputs. */
int b200_ln_gemm_bf16(const float* x, const float* ln_w, const float* ln_b, float eps, float* mean, float* rstd,
                      void* xn_ws, const b200_gemm_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------------ bandwidth kernels */
/* fp32 -> bf16 / fp16 casts and fp16 -> bf16 (n elements). */
int b200_cast_f32_bf16(const float* x, void* y, long long n, void* stream);
int b200_cast_f32_f16(const float* x, void* y, long long n, void* stream);
int b200_cast_f16_bf16(const void* x, void* y, long long n, void* stream);
/* 3-term split cast, fp32 [rows, K] -> 16-bit [rows, 3K] = [hi | hi | lo] (left operand) or [hi | lo | hi] (right
 * operand), hi = r16(x), lo = r16(x - hi). A plain GEMM over the 3K contraction then carries ~2x the mantissa bits.
 * Used for proj_student's conv1x1, whose output feeds BatchNorm -> ReLU with no residual around it: there a flipped
 * ReLU mask is a first-order gradient error (precision policy). */
int b200_split3_16(const float* x, void* out, long long rows, int K, int right_operand, int out_is_fp16, void* stream);
/* out[c, r] = in[r, c] for in [rows, cols] fp32 -> bf16 (weight transposes for dgrad); optional per-row scale. */
int b200_transpose_f32_bf16(const float* in, void* out, int rows, int cols, const float* row_scale, void* stream);
/* same with an explicit output row pitch (elements), to write side-by-side blocks of a concatenated matrix. */
int b200_transpose_f32_bf16_ld(const float* in, void* out, int rows, int cols, long long out_ld,
                               const float* row_scale, void* stream);
/* NCHW fp32 [B,C,HW] <-> token-major [B*HW, C]. to_tokens: fp32 -> bf16 (+ optional fp32 copy); from_tokens: fp32 -> fp32. */
int b200_nchw_to_tokens(const float* x, void* tok_bf16, float* tok_f32, int B, int C, int HW, int tok16_is_fp16,
                        void* stream);
int b200_tokens_to_nchw(const float* tok, float* x, int B, int C, int HW, int accumulate, void* stream);

/* Bilinear resize (align_corners = False: F.interpolate of models/model_zoo.py:121-126) of token-major maps
 * fp32 [B, h*w, D] -> [B, H*W, D], and its adjoint on bf16 gradients [B, H*W, D] -> [B, h*w, D]. The projector uses
 * them to run proj_student's 1x1 conv on the RAW student map and resize afterwards (see b200_projector_config). */
int b200_bilinear_tokens_fwd(const float* src, float* dst, int B, int h, int w, int H, int W, int D, void* stream);
int b200_bilinear_tokens_bwd(const void* d_dst_bf16, void* d_src_bf16, int B, int h, int w, int H, int W, int D,
                             void* stream);

/* Token order of windowed attention (WindowMultiheadPosAttention.separate_tokens, losses/scalekd.py:326-335): rows of a
 * 16-bit token-major matrix [rows = B*H*W, ld] (cols <= ld copied) between raster order and window-major order
 * (to_raster = 0: raster -> window-major; 1: back). src != dst. */
int b200_window_rows16(const void* src, void* dst, long long rows, int H, int W, int win_h, int win_w, int cols,
                       long long ld, int to_raster, void* stream);

/* patch-embed im2col: images fp32 [B,3,H,W] -> bf16 [B*(H/14)*(W/14), Kp], column = c*196 + i*14 + j, zero padded to
 * Kp (>= 588, multiple of 8).  (hub PatchEmbed.proj, Conv2d(3,D,14,14), reached via models/backbones/dinov2.py:32) */
int b200_patch_im2col(const float* img, void* out, int B, int H, int W, int Kp, void* stream);
/* cls rows: x[b*N + 0, :] = cls[:] + pos[0, :]  */
int b200_write_cls_rows(float* x, const float* cls, const float* pos, int B, int N, int D, void* stream);

/* LayerNorm over the last dim (fp32 in). Any of y_f32 / y_bf16 / mean / rstd may be NULL.
 * row mapping: in row = (r/in_period)*(in_period+in_pad)+in_pad + r%in_period when in_period>0 (drop cls rows).
 * (hub Block.norm1/norm2/norm eps 1e-6; losses/scalekd.py:215-216 eps 1e-5) */
int b200_layernorm_fwd(const float* x, const float* w, const float* b, float eps, float* y_f32, void* y_bf16,
                       float* mean, float* rstd, int rows, int D, int in_period, int in_pad, int y16_is_fp16,
                       void* stream);
/* dx = LNbwd(dy) (+ dres if given). dw/db (fp32 [D]) accumulate with atomics when non-NULL. dx_bf16 optional copy. */
int b200_layernorm_bwd(const float* dy, const float* x, const float* w, const float* mean, const float* rstd,
                       const float* dres, float* dx, void* dx_bf16, float* dw, float* db, int rows, int D,
                       void* stream);
/* same; dx_colsum (fp32 [D], accumulated, may be NULL) additionally receives the column sums of the output dx: the bias
 * gradient of the linear layer feeding this LayerNorm (losses/scalekd.py:243-245: ffn2 / proj biases), which saves a
 * separate pass over dx. */
int b200_layernorm_bwd_colsum(const float* dy, const float* x, const float* w, const float* mean, const float* rstd,
                              const float* dres, float* dx, void* dx_bf16, float* dw, float* db, float* dx_colsum,
                              int rows, int D, void* stream);

/* BatchNorm2d on token-major y [M, D] (training statistics over M = B*H*W rows). losses/scalekd.py:200
 * stats: sums[0:D] = sum_r y, sums[D:2D] = sum_r y^2 (buffer must be zeroed).  */
int b200_bn_stats(const float* y, float* sums, int M, int D, void* stream);
/* mean/var -> (mean, rstd); update running stats (momentum; unbiased var) when running_* non-NULL. */
int b200_bn_finalize(const float* sums, float* mean, float* rstd, float* running_mean, float* running_var,
                     float momentum, float eps, int M, int D, void* stream);
/* z = relu((y-mean)*rstd*w + b) + pos[r % HW, :] ; z_f32 and z_bf16 outputs. */
int b200_bn_relu_pos_fwd(const float* y, const float* mean, const float* rstd, const float* w, const float* b,
                         const float* pos, float* z_f32, void* z_bf16, int M, int D, int HW, int z16_is_fp16,
                         void* stream);
/* backward of the above, pass 1: dr = dz * (z_pre > 0); sums2[0:D] += sum dr, sums2[D:2D] += sum dr*yhat;
 * dpos[r%HW,:] += dz (atomics).  pass 2: dy = w*rstd*(dr - sum_dr/M - yhat*sum_dr_yhat/M). */
int b200_bn_relu_pos_bwd_reduce(const float* dz, const float* y, const float* mean, const float* rstd, const float* w,
                                const float* b, float* sums2, float* dpos, int M, int D, int HW, void* stream);
int b200_bn_relu_pos_bwd_apply(const float* dz, const float* y, const float* mean, const float* rstd, const float* w,
                               const float* b, const float* sums2, void* dy_bf16, int use_batch_stats, int M, int D,
                               void* stream);

/* column sums: out[c] (+)= sum_r x[r, c];  x fp32 or bf16 */
int b200_colsum(const void* x, int x_is_bf16, long long ldx, float* out, int rows, int cols, void* stream);
/* out[i] += sum_b x[b*n + i]  (fp32, accumulating): pos_embed gradient, batch-invariant self query */
int b200_batch_sum(const float* x, float* out, int B, long long n, void* stream);
/* bf16 input variant: out_f32[i] = sum_b x[b*n+i] (written), optional bf16 copy. */
int b200_batch_sum_bf16(const void* x, float* out_f32, void* out_bf16, int B, long long n, void* stream);
/* y[i] += a * x[i] */
int b200_axpy(const float* x, float* y, float a, long long n, void* stream);
/* SwiGLU gate: out[r, j] = silu(x[r, j]) * x[r, H + j], x bf16 [rows, 2H] -> bf16 [rows, H]  (hub SwiGLUFFNFused) */
int b200_swiglu(const void* x12, void* out, int rows, int H, void* stream);
/* d_x12[r, j] = d_out[r,j] * x2 * silu'(x1) ; d_x12[r, H+j] = d_out[r,j] * silu(x1)   (all bf16) */
int b200_swiglu_bwd(const void* x12, const void* d_out, void* d_x12, int rows, int H, void* stream);
/* backward of GELU applied on the fly is fused into the GEMM epilogue (B200_AUX_DGELU). */

/* ------------------------------------------------------------------------------------------------ attention
 * softmax(scale * q k^T) v per (batch, head), flash style (no [B,h,N,N] tensor).  element (b, t, head, d) of X is
 * X[b*bs + t*ts + head*hd + d].  q batch stride may be 0 (batch-invariant self query, losses/scalekd.py:234).
 * hd in {16, 24, 32, 48, 64, 96}.  lse: fp32 [B, heads, Nq] (natural-log sum-exp of scaled logits).
 * (hub Attention.forward; losses/scalekd.py:299-314) */
typedef struct b200_attn_desc {
  const void* q; long long q_bs, q_ts;
  const void* k; long long k_bs, k_ts;
  const void* v; long long v_bs, v_ts;
  void* o; long long o_bs, o_ts;      /* bf16 */
  float* lse;
  int B, heads, Nq, Nk, hd;
  float scale;
  /* backward only */
  const void* d_o; long long do_bs, do_ts; /* bf16 */
  float* delta;                           /* fp32 [B, heads, Nq] workspace */
  void* dq; long long dq_bs, dq_ts;       /* bf16 (per-batch, even if q is batch invariant) */
  void* dk; long long dk_bs, dk_ts;       /* bf16 */
  void* dv; long long dv_bs, dv_ts;       /* bf16 */
  /* 1: q, k, v, o are fp16 (ScaleKD projector forward precision); gradients d_o/dq/dk/dv are always bf16 */
  int qkvo_is_fp16;
  /* backward only, optional (all three or none): fp32 [heads*hd], ACCUMULATED column sums of dq / dk / dv over
   * (batch, tokens) -- the bias gradients of the q / k / v projections (losses/scalekd.py:277-279) */
  float* dq_colsum; float* dk_colsum; float* dv_colsum;
  /* forward only, optional: a second copy of o (same strides) in the 16-bit format o does not use */
  void* o_alt;
  /* backward only, optional (all three or none), used when qkvo_is_fp16: bf16 copies of q / k / v with the strides of
   * the fp16 tensors. The tcgen05 backward needs them (kind::f16 cannot mix fp16 and bf16 operands in one product: the
   * scores are recomputed from the fp16 tensors, every gradient product reads the bf16 copies); without them the
   * mma.sync kernels run, which convert fragments in registers. */
  const void* q_alt; const void* k_alt; const void* v_alt;
  /* backward only, optional: fp32 workspace [B, Nq, heads*hd]. Sequences longer than 256 tokens run the tcgen05
   * backward one key block per work unit and add each unit's dQ contribution into this accumulator (bulk tensor
   * reduce-add); without it they run the two-kernel mma.sync backward. */
  float* dq_accum;
} b200_attn_desc;
int b200_attention_fwd(const b200_attn_desc* d, void* stream);
int b200_attention_bwd(const b200_attn_desc* d, void* stream);

/* ------------------------------------------------------------------------------------------------ ScaleKD loss terms
 * losses/scalekd.py:67-92 (get_spat_loss) and :95-127 (get_freq_loss).
 * S: fp32 tokens [B*HW, D]; T: fp32 [B, Nt, D] teacher tokens with t_skip leading rows per image dropped (cls).
 * freq != 0 subtracts the per-(b, channel) spatial mean from both (== idct(zeroDC(dct(.)))).
 * out[0] = alpha/B * sum (S^ - T^)^2 ; out[1] = mean cosine.   ws: fp32 workspace, b200_kd_loss_ws_floats() long. */
long long b200_kd_loss_ws_floats(int B, int HW, int D);
int b200_kd_loss_fwd(const float* S, const float* T, int B, int HW, int D, int Nt, int t_skip, int freq, float alpha,
                     float* out, float* ws, void* stream);
/* dS (+)= g_loss * d(out[0])/dS   (the similarity output carries no useful gradient; the reference's does, see
 * DESIGN.md -- g_sim is supported for completeness). */
int b200_kd_loss_bwd(const float* S, const float* T, int B, int HW, int D, int Nt, int t_skip, int freq, float alpha,
                     const float* g_out /* device [2] */, float* dS, int accumulate, float* ws, void* stream);
/* Same with the two upstream gradients as separate device scalars (g_sim may be NULL = 0): lets the Python shell hand
 * out the loss and the similarity as two autograd outputs without select / stack kernels in between. */
int b200_kd_loss_bwd_split(const float* S, const float* T, int B, int HW, int D, int Nt, int t_skip, int freq,
                           float alpha, const float* g_loss, const float* g_sim, float* dS, int accumulate, float* ws,
                           void* stream);
/* explicit separable 2-D DCT-II -> zero DC -> inverse, per (b, channel) plane, token-major in/out. Cross-check for the
 * mean-subtraction identity (losses/scalekd.py:337-428). R = H = W <= 64. */
int b200_dct_zero_dc_idct(const float* x, float* y, int B, int R, int D, long long x_bs, long long x_ts, void* stream);

/* ------------------------------------------------------------------------------------------------ input pipeline
 * datasets/augmentations.py:24-78 without RandAugment -- RandomResizedCrop (PIL's antialiased bicubic over
 * the crop box, restated integer for integer: bit-exact with the reference transforms) + horizontal flip + ToTensor +
 * Normalize + RandomErasing(value 0). The random parameters are drawn on the host (torchvision's get_params: same
 * distributions and draws as the reference) and passed in as device arrays.
 * pixels: decoded uint8 RGB, HWC, images packed back to back (offsets[b] = first byte of image b; hw[b] = {H, W});
 * crop[b] = {top, left, height, width} in the source; flip[b] != 0 mirrors the resized crop; erase[b] = {top, left,
 * height, width} in the OUTPUT (height 0 = no erasing); mean / std: HOST float[3]. out: fp32 [B, 3, S, S].
 * max_taps >= 2 * ceil(2 * max(largest crop side / S, 1)) + 1 (Pillow's ksize); ws: b200_augment_ws_bytes(). */
size_t b200_augment_ws_bytes(int B, int S, int max_taps);
int b200_augment_batch(const unsigned char* pixels, const long long* offsets, const int* hw, const int* crop,
                       const int* flip, const int* erase, float* out, int B, int S, int max_taps, const float* mean,
                       const float* std, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------ teacher-feature cache
 * SURVEY 8 f3. The frozen teacher's output (models/backbones/dinov2.py:27-46) depends only on the input image: with a
 * deterministic input pipeline it can be kept in HBM per sample and the teacher forward skipped on later epochs.
 * pool: [slots_total, N, D] bf16 (pool_is_bf16 = 1) or fp32; slots: device int64 [B], slot < 0 = skip this batch item.
 * store: feat fp32 [B, N, D] with batch / token strides f_bs / f_ts in elements (the teacher's strided token view, cls row
 * skipped by pointer offset, is passed as is). load: out fp32 [B, N, D] contiguous. */
int b200_feature_cache_store(const float* feat, long long f_bs, long long f_ts, const long long* slots, void* pool,
                             int pool_is_bf16, int B, int N, int D, void* stream);
int b200_feature_cache_load(const void* pool, int pool_is_bf16, const long long* slots, float* out, int B, int N, int D,
                            void* stream);

/* ------------------------------------------------------------------------------------------------ teacher (DINOv2 ViT)
 * Weight tables are host arrays of device pointers, filled by the Python shell from the hub-format state_dict. */
typedef struct b200_vit_block {
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
  const void* qkv_w;  const float* qkv_b;    /* bf16 [3D, D] */
  const void* proj_w; const float* proj_b;   /* bf16 [D, D]  */
  const float *ls1, *ls2;                    /* LayerScale gamma [D] */
  const void* fc1_w;  const float* fc1_b;    /* bf16 [F, D]   (vitg: w12 [2F, D]) */
  const void* fc2_w;  const float* fc2_b;    /* bf16 [D, F]   (vitg: w3) */
  /* transposed copies for the input-gradient pass (only for blocks re-used by _forward_specific_stage) */
  const void* qkv_wT;    /* bf16 [D, 3D]  */
  const void* proj_wT;   /* bf16 [D, D]   = (diag(ls1) proj_w)^T */
  const void* fc1_wT;    /* bf16 [D, F or 2F] */
  const void* fc2_wT;    /* bf16 [F, D]   = (diag(ls2) fc2_w)^T */
} b200_vit_block;

typedef struct b200_vit_config {
  int D, L, heads, F, swiglu; /* F = ffn hidden (after gating for swiglu) */
  float ln_eps;
} b200_vit_config;

/* whole teacher forward: images fp32 [B,3,H,W] -> normed tokens fp32 [B, N, D], N = 1 + (H/14)*(W/14).
 * pos: fp32 [N, D] (already interpolated for this grid); patch_w: bf16 [D, Kp]. */
size_t b200_vit_forward_ws_bytes(const b200_vit_config* c, int B, int H, int W);
int b200_vit_forward(const b200_vit_config* c, const b200_vit_block* blocks, const void* patch_w, int Kp,
                     const float* patch_b, const float* cls, const float* pos, const float* norm_w,
                     const float* norm_b, const float* img, int B, int H, int W, float* out_tokens, void* ws,
                     size_t ws_bytes, void* stream);

/* one block on tokens x fp32 [B, N, D] -> y fp32 (may alias x when save == NULL). `save` (b200_vit_block_save_bytes)
 * keeps what the input-gradient pass needs.  train/distillation_module.py:176-177 */
size_t b200_vit_block_ws_bytes(const b200_vit_config* c, int B, int N);
size_t b200_vit_block_save_bytes(const b200_vit_config* c, int B, int N);
int b200_vit_block_fwd(const b200_vit_config* c, const b200_vit_block* blk, const float* x, float* y, int B, int N,
                       void* save, void* ws, size_t ws_bytes, void* stream);
/* x: the block's forward input (needed by LayerNorm backward); dx may alias dy. */
int b200_vit_block_bwd_input(const b200_vit_config* c, const b200_vit_block* blk, const float* x, const float* dy,
                             float* dx, int B, int N, const void* save, void* ws, size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------ ScaleKD projector
 * AttentionProjector.forward / backward, losses/scalekd.py:177-245. Parameters are fp32 master copies (as held by the
 * nn.Module); grads are fp32 and ACCUMULATED (+=) so they can alias .grad buffers / a flat arena. */
typedef struct b200_projector_params {
  const float *conv_w, *conv_b;          /* [D, Cs], [D] */
  const float *bn_w, *bn_b;              /* [D] */
  float *bn_running_mean, *bn_running_var; /* [D], updated in training forward */
  const float* pos_embed;                /* [D, HW] (the module's [1,D,H,W]) */
  const float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *p_w, *p_b; /* [D,D], [D] */
  const float *ffn1_w, *ffn1_b, *ffn2_w, *ffn2_b; /* [4D,D],[4D],[D,4D],[D] */
  const float *ln1_w, *ln1_b, *ln2_w, *ln2_b;     /* norm, norm_2 */
  const float* query_w;                  /* [HW, D] or NULL (self_query=False) */
  long long* bn_num_batches_tracked;     /* int64 device scalar or NULL: += 1 per training forward (nn.BatchNorm2d) */
} b200_projector_params;

typedef struct b200_projector_grads {
  float *conv_w, *conv_b, *bn_w, *bn_b, *pos_embed;
  float *q_w, *q_b, *k_w, *k_b, *v_w, *v_b, *p_w, *p_b;
  float *ffn1_w, *ffn1_b, *ffn2_w, *ffn2_b, *ln1_w, *ln1_b, *ln2_w, *ln2_b, *query_w;
} b200_projector_grads;

typedef struct b200_projector_config {
  int Cs, D, HW, heads;
  float softmax_scale;   /* logits *= head_dim^-0.5 * softmax_scale */
  float bn_eps, bn_momentum, ln_eps;
  int training;          /* BN: batch statistics (1) or running statistics (0) */
  /* Fused ModelWrapper resize (models/model_zoo.py:118-128). raw_h > 0: x (and dx) are the student map
   * BEFORE the bilinear resize, fp32 NCHW [B, Cs, raw_h, raw_w]; the 1x1 conv runs at that resolution and its
   * D-channel output is resized to the teacher grid grid_h x grid_w (grid_h * grid_w == HW) -- equal to resizing first
   * (both maps are linear and the tap weights sum to one) at raw_h*raw_w / HW of the conv FLOPs and token traffic.
   * raw_h == 0: x is already [B, Cs, HW]. */
  int raw_h, raw_w, grid_h, grid_w;
  /* window_shapes of WindowMultiheadPosAttention (losses/scalekd.py:286, :305-314): win_h * win_w > 1 cuts the
   * grid_h x grid_w token grid into windows, attention runs inside each window, and -- as in the reference -- the
   * attention output STAYS in window-major token order. 0 or 1: no windows. Needs grid_h % win_h == 0 etc. */
  int win_h, win_w;
} b200_projector_config;

size_t b200_projector_ws_bytes(const b200_projector_config* c, int B);
size_t b200_projector_save_bytes(const b200_projector_config* c, int B);
/* x: fp32 NCHW [B, Cs, HW]; query: fp32 [B, HW, D] or NULL (then params.query_w); out: fp32 [B, HW, D]. */
int b200_projector_fwd(const b200_projector_config* c, const b200_projector_params* p, const float* x,
                       const float* query, int B, float* out, void* save, void* ws, size_t ws_bytes, void* stream);
/* dout fp32 [B,HW,D] -> grads (accumulated), dx fp32 NCHW (accumulated when dx_accumulate), dquery (written) or NULL */
int b200_projector_bwd(const b200_projector_config* c, const b200_projector_params* p, const b200_projector_grads* g,
                       const float* x, const float* query, const float* dout, int B, float* dx, int dx_accumulate,
                       float* dquery, const void* save, void* ws, size_t ws_bytes, void* stream);

/* Both projectors of one ScaleKD (projector_0 / projector_1, losses/scalekd.py:27-46) read the same preds_S. Tokenise it
 * once -- NCHW fp32 -> token-major bf16 + the 3-term fp16 split the conv1x1 consumes -- and hand the result to the
 * *_tok variants (tokens == NULL: identical to the plain entry points, which tokenise privately). The buffer must stay
 * alive until both backward calls have run. ws: b200_projector_ws_bytes() is enough. */
size_t b200_projector_tokens_bytes(const b200_projector_config* c, int B);
int b200_projector_tokenize(const b200_projector_config* c, const float* x, int B, void* tokens, void* ws,
                            size_t ws_bytes, void* stream);
int b200_projector_fwd_tok(const b200_projector_config* c, const b200_projector_params* p, const float* x,
                           const float* query, int B, float* out, void* save, void* ws, size_t ws_bytes,
                           const void* tokens, void* stream);
int b200_projector_bwd_tok(const b200_projector_config* c, const b200_projector_params* p,
                           const b200_projector_grads* g, const float* x, const float* query, const float* dout, int B,
                           float* dx, int dx_accumulate, float* dquery, const void* save, void* ws, size_t ws_bytes,
                           const void* tokens, void* stream);

/* ------------------------------------------------------------------------------------------------ optimizer step (next row f2)
 * Global-norm clip + AdamW over flat fp32 arenas (parameters, gradients, first / second moments): the arithmetic of
 * torch.optim.AdamW (train/distillation_module.py:440-502; config/config.yaml:25-30) after Lightning's
 * gradient_clip_val (train.py:267-268), for the loss-module parameters.
 * b200_sqnorm_f32: out_accum[0] += sum x^2.  b200_adamw_step: g is scaled by min(1, max_grad_norm / (sqrt(grad_sqnorm[0] +
 * extra_sqnorm[0]) + 1e-6)) (extra: e.g. the student's share of the global norm; NULL = 0; max_grad_norm <= 0: no clip);
 * step is the 1-based step count (bias correction). */
int b200_sqnorm_f32(const float* x, long long n, float* out_accum, void* stream);
int b200_adamw_step(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                    float eps, float weight_decay, int step, const float* grad_sqnorm, const float* extra_sqnorm,
                    float max_grad_norm, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200_DISTILL_H */
