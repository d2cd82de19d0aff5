"""Tensor-level wrappers over the C ABI (PyTorch is only used for device memory and streams).

Every function requires CUDA tensors and launches hand-written kernels from libb200distill.so on the current stream.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L

ACT = {"none": 0, "gelu": 1, "relu": 2}
AUX = {"none": 0, "dgelu": 1, "drelu": 2}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise L.B200Error("b200 ops need CUDA tensors: there is no CPU fallback")


def launch_count() -> int:
    return int(L.load().b200_launch_count())


def reset_launch_count() -> None:
    L.load().b200_reset_launch_count()


def gemm(a: torch.Tensor, b: torch.Tensor, *, a_mn_major: bool = False, b_mn_major: bool = False,
         bias: Optional[torch.Tensor] = None, act: str = "none", aux: Optional[torch.Tensor] = None,
         aux_mode: str = "none", col_scale: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
         res_row_period: int = 0, out_dtype: torch.dtype = torch.float32, out: Optional[torch.Tensor] = None,
         atomic_add: bool = False, split_k: int = 1, out_pre: bool = False, out_row_period: int = 0,
         out_row_pad: int = 0, out_batch_period: int = 0, out_alt: bool = False,
         out_colsum: Optional[torch.Tensor] = None):
    """C = epilogue(A @ B^T). K-major: a [M,K], b [N,K]. MN-major: a [K,M], b [K,N] (contraction over rows)."""
    _need_cuda(a, b, bias, aux, col_scale, residual, out)
    assert a.dtype in (torch.bfloat16, torch.float16) and b.dtype in (torch.bfloat16, torch.float16)
    assert a.stride(-1) == 1 and b.stride(-1) == 1
    if a_mn_major:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn_major:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    assert K == Kb, (a.shape, b.shape)
    d = L.GemmDesc()
    d.A, d.lda, d.a_mn_major = a.data_ptr(), a.stride(0), int(a_mn_major)
    d.B, d.ldb, d.b_mn_major = b.data_ptr(), b.stride(0), int(b_mn_major)
    d.M, d.N, d.K, d.split_k = M, N, K, split_k
    d.a_is_fp16, d.b_is_fp16 = int(a.dtype == torch.float16), int(b.dtype == torch.float16)
    d.bias = _p(bias)
    d.act = ACT[act]
    if aux is not None:
        assert aux.dtype in (torch.bfloat16, torch.float16)
        d.aux, d.ldaux, d.aux_mode = aux.data_ptr(), aux.stride(0), AUX[aux_mode]
        d.aux_is_fp16 = int(aux.dtype == torch.float16)
    d.col_scale = _p(col_scale)
    if residual is not None:
        assert residual.dtype == torch.float32
        d.residual, d.ldres, d.res_row_period = residual.data_ptr(), residual.stride(0), res_row_period
    out_rows = M if out_row_period <= 0 else (M // out_row_period) * (out_row_period + out_row_pad)
    if out_batch_period > 0:
        # columns batched with period P: out is [N / P, M, P] fp32 (element (r, n) -> out[n // P, r, n % P])
        assert out is not None and out.dtype == torch.float32 and out.dim() == 3 and out.is_contiguous()
        assert tuple(out.shape) == (N // out_batch_period, M, out_batch_period)
        d.out_batch_period, d.out_batch_stride = out_batch_period, out.stride(0)
        d.out_f32, d.ldo32, d.atomic_add = out.data_ptr(), out.stride(1), int(atomic_add)
        L.check(L.load().b200_gemm_bf16(C.byref(d), _stream()), "gemm_bf16")
        return out
    if out is None:
        out = torch.empty(out_rows, N, device=a.device, dtype=out_dtype)
    pre = None
    if out.dtype == torch.float32:
        d.out_f32, d.ldo32, d.atomic_add = out.data_ptr(), out.stride(0), int(atomic_add)
    else:
        assert out.dtype in (torch.bfloat16, torch.float16)
        d.out_bf16, d.ldo16 = out.data_ptr(), out.stride(0)
        d.out16_is_fp16 = int(out.dtype == torch.float16)
    if out_pre:
        pre = torch.empty(out_rows, N, device=a.device, dtype=out.dtype if out.dtype != torch.float32 else torch.bfloat16)
        d.out_bf16_pre, d.ldo16_pre = pre.data_ptr(), pre.stride(0)
    if out_alt:   # second copy of the final value in the other 16-bit format (shares the pre-activation slot)
        assert not out_pre and out.dtype in (torch.bfloat16, torch.float16)
        pre = torch.empty(out_rows, N, device=a.device,
                          dtype=torch.bfloat16 if out.dtype == torch.float16 else torch.float16)
        d.out_bf16_pre, d.ldo16_pre, d.out16_pre_alt = pre.data_ptr(), pre.stride(0), 1
    d.out_row_period, d.out_row_pad = out_row_period, out_row_pad
    if out_colsum is not None:   # fp32 [N], += column sums of the 16-bit output as stored
        assert out_colsum.dtype == torch.float32 and out_colsum.numel() == N and out_colsum.is_contiguous()
        d.out16_colsum = out_colsum.data_ptr()
    L.check(L.load().b200_gemm_bf16(C.byref(d), _stream()), "gemm_bf16")
    return (out, pre) if (out_pre or out_alt) else out


def bilinear_tokens(src: torch.Tensor, hw, HW) -> torch.Tensor:
    """fp32 tokens [B, h*w, D] -> [B, H*W, D]: F.interpolate(mode='bilinear', align_corners=False) on token-major maps."""
    _need_cuda(src)
    (h, w), (H, W) = hw, HW
    src = src.contiguous().float()
    B, n, D = src.shape
    assert n == h * w
    dst = torch.empty(B, H * W, D, device=src.device, dtype=torch.float32)
    L.check(L.load().b200_bilinear_tokens_fwd(src.data_ptr(), dst.data_ptr(), B, h, w, H, W, D, _stream()), "bilinear_fwd")
    return dst


def bilinear_tokens_adjoint(d_dst: torch.Tensor, hw, HW) -> torch.Tensor:
    """bf16 gradient tokens [B, H*W, D] -> [B, h*w, D] (transpose of `bilinear_tokens`)."""
    _need_cuda(d_dst)
    (h, w), (H, W) = hw, HW
    d_dst = d_dst.contiguous()
    assert d_dst.dtype == torch.bfloat16
    B, n, D = d_dst.shape
    assert n == H * W
    out = torch.empty(B, h * w, D, device=d_dst.device, dtype=torch.bfloat16)
    L.check(L.load().b200_bilinear_tokens_bwd(d_dst.data_ptr(), out.data_ptr(), B, h, w, H, W, D, _stream()), "bilinear_bwd")
    return out


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    x = x.contiguous()
    y = torch.empty_like(x, dtype=torch.bfloat16)
    L.check(L.load().b200_cast_f32_bf16(x.data_ptr(), y.data_ptr(), x.numel(), _stream()), "cast")
    return y


def transpose_bf16(w: torch.Tensor, row_scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [R, C] -> bf16 [C, R] (optionally rows scaled first)."""
    _need_cuda(w, row_scale)
    w = w.contiguous()
    R, Cc = w.shape
    out = torch.empty(Cc, R, device=w.device, dtype=torch.bfloat16)
    L.check(L.load().b200_transpose_f32_bf16(w.data_ptr(), out.data_ptr(), R, Cc, _p(row_scale), _stream()), "transpose")
    return out


def nchw_to_tokens(x: torch.Tensor, want_f32: bool = False):
    _need_cuda(x)
    x = x.contiguous()
    B, Cc = x.shape[0], x.shape[1]
    HW = x[0, 0].numel()
    t16 = torch.empty(B * HW, Cc, device=x.device, dtype=torch.bfloat16)
    t32 = torch.empty(B * HW, Cc, device=x.device, dtype=torch.float32) if want_f32 else None
    L.check(L.load().b200_nchw_to_tokens(x.data_ptr(), t16.data_ptr(), _p(t32), B, Cc, HW, 0, _stream()), "nchw_to_tokens")
    return (t16, t32) if want_f32 else t16


def tokens_to_nchw(tok: torch.Tensor, B: int, HW: int) -> torch.Tensor:
    _need_cuda(tok)
    Cc = tok.shape[-1]
    out = torch.empty(B, Cc, HW, device=tok.device, dtype=torch.float32)
    L.check(L.load().b200_tokens_to_nchw(tok.data_ptr(), out.data_ptr(), B, Cc, HW, 0, _stream()), "tokens_to_nchw")
    return out


def patch_im2col(img: torch.Tensor, Kp: int = 592) -> torch.Tensor:
    _need_cuda(img)
    img = img.contiguous()
    B, _, H, W = img.shape
    out = torch.empty(B * (H // 14) * (W // 14), Kp, device=img.device, dtype=torch.bfloat16)
    L.check(L.load().b200_patch_im2col(img.data_ptr(), out.data_ptr(), B, H, W, Kp, _stream()), "patch_im2col")
    return out


def ln_gemm(x: torch.Tensor, ln_w: torch.Tensor, ln_b: torch.Tensor, eps: float, w: torch.Tensor, *,
            bias: Optional[torch.Tensor] = None, act: str = "none", out_pre: bool = False, stats: bool = False):
    """C = epilogue(LayerNorm(x) @ w^T): x fp32 [M, K] contiguous, w bf16 [N, K] (b200_ln_gemm_bf16: one kernel for
    K <= 384, LayerNorm + GEMM otherwise). Returns (out bf16 [M, N], pre-activation copy or None, mean, rstd)."""
    _need_cuda(x, ln_w, ln_b, w, bias)
    assert x.dtype == torch.float32 and x.is_contiguous() and w.dtype == torch.bfloat16 and w.stride(-1) == 1
    M, K = x.shape
    N = w.shape[0]
    d = L.GemmDesc()
    d.B, d.ldb = w.data_ptr(), w.stride(0)
    d.M, d.N, d.K, d.split_k = M, N, K, 1
    d.bias = _p(bias)
    d.act = ACT[act]
    out = torch.empty(M, N, device=x.device, dtype=torch.bfloat16)
    d.out_bf16, d.ldo16 = out.data_ptr(), N
    pre = None
    if out_pre:
        pre = torch.empty(M, N, device=x.device, dtype=torch.bfloat16)
        d.out_bf16_pre, d.ldo16_pre = pre.data_ptr(), N
    mean = torch.empty(M, device=x.device, dtype=torch.float32) if stats else None
    rstd = torch.empty(M, device=x.device, dtype=torch.float32) if stats else None
    ws = torch.empty(M, K, device=x.device, dtype=torch.bfloat16)
    L.check(L.load().b200_ln_gemm_bf16(x.data_ptr(), ln_w.data_ptr(), ln_b.data_ptr(), float(eps), _p(mean), _p(rstd),
                                       ws.data_ptr(), C.byref(d), _stream()), "ln_gemm_bf16")
    return out, pre, mean, rstd


def layernorm_fwd(x, w, b, eps, want_f32=True, want_bf16=False, want_stats=False, in_period=0, in_pad=0, rows=None):
    _need_cuda(x, w, b)
    D = x.shape[-1]
    if rows is None:
        rows = x.numel() // D
    y32 = torch.empty(rows, D, device=x.device, dtype=torch.float32) if want_f32 else None
    y16 = torch.empty(rows, D, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    mean = torch.empty(rows, device=x.device, dtype=torch.float32) if want_stats else None
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32) if want_stats else None
    L.check(L.load().b200_layernorm_fwd(x.data_ptr(), w.data_ptr(), b.data_ptr(), float(eps), _p(y32), _p(y16), _p(mean),
                                        _p(rstd), rows, D, in_period, in_pad, 0, _stream()), "layernorm_fwd")
    return y32, y16, mean, rstd


def layernorm_bwd(dy, x, w, mean, rstd, dres=None, want_wgrad=True, want_bf16=False):
    _need_cuda(dy, x, w, mean, rstd, dres)
    D = x.shape[-1]
    rows = x.numel() // D
    dx = torch.empty(rows, D, device=x.device, dtype=torch.float32)
    dx16 = torch.empty(rows, D, device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    dw = torch.zeros(D, device=x.device, dtype=torch.float32) if want_wgrad else None
    db = torch.zeros(D, device=x.device, dtype=torch.float32) if want_wgrad else None
    L.check(L.load().b200_layernorm_bwd(dy.data_ptr(), x.data_ptr(), w.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                        _p(dres), dx.data_ptr(), _p(dx16), _p(dw), _p(db), rows, D, _stream()),
            "layernorm_bwd")
    return dx, dx16, dw, db


def _attn_desc(q, k, v, o, lse, heads, scale):
    # q: [B, Nq, heads*hd] view (possibly expanded over batch), k/v: [B, Nk, heads*hd] views
    B, Nq, Dm = q.shape
    Nk = k.shape[1]
    d = L.AttnDesc()
    d.q, d.q_bs, d.q_ts = q.data_ptr(), q.stride(0), q.stride(1)
    d.k, d.k_bs, d.k_ts = k.data_ptr(), k.stride(0), k.stride(1)
    d.v, d.v_bs, d.v_ts = v.data_ptr(), v.stride(0), v.stride(1)
    d.o, d.o_bs, d.o_ts = o.data_ptr(), o.stride(0), o.stride(1)
    d.lse = _p(lse)
    d.B, d.heads, d.Nq, d.Nk, d.hd = B, heads, Nq, Nk, Dm // heads
    d.scale = float(scale)
    d.qkvo_is_fp16 = int(q.dtype == torch.float16)
    return d


def attention_fwd(q, k, v, heads: int, scale: float, want_alt: bool = False):
    """q [B,Nq,D], k/v [B,Nk,D] bf16 or fp16 (last dim contiguous, any batch/token strides) -> o [B,Nq,D] in the input
    format, lse [B,h,Nq]. want_alt: also return a copy of o in the other 16-bit format (b200_attn_desc.o_alt)."""
    _need_cuda(q, k, v)
    B, Nq, Dm = q.shape
    o = torch.empty(B, Nq, Dm, device=q.device, dtype=q.dtype)
    lse = torch.empty(B, heads, Nq, device=q.device, dtype=torch.float32)
    d = _attn_desc(q, k, v, o, lse, heads, scale)
    o_alt = None
    if want_alt:
        o_alt = torch.empty(B, Nq, Dm, device=q.device,
                            dtype=torch.bfloat16 if q.dtype == torch.float16 else torch.float16)
        d.o_alt = o_alt.data_ptr()
    L.check(L.load().b200_attention_fwd(C.byref(d), _stream()), "attention_fwd")
    return (o, lse, o_alt) if want_alt else (o, lse)


def set_option(name: str, value: int) -> None:
    """Flip a dispatch option of the library (include/b200_distill.h: b200_set_option)."""
    L.check(L.load().b200_set_option(name.encode(), int(value)), "set_option")


def get_option(name: str) -> int:
    return int(L.load().b200_get_option(name.encode()))


def _bf16_copy_same_strides(t):
    """bf16 copy of a 16-bit tensor with t's strides (the *_alt operands of b200_attn_desc), or None when t's layout is
    not one a plain copy reproduces (the backward then runs its register-converting kernels)."""
    if t.is_contiguous():
        return t.to(torch.bfloat16)
    if t.stride(0) == 0 and t[0].is_contiguous():
        return t[:1].to(torch.bfloat16).expand_as(t)
    return None


def attention_bwd(q, k, v, o, lse, d_o, heads: int, scale: float, colsums=None, alts=None):
    """colsums: optional (dq_sum, dk_sum, dv_sum) fp32 [heads*hd] accumulators (bias gradients).
    alts: optional precomputed bf16 copies of fp16 (q, k, v) with the same strides (made here when omitted)."""
    _need_cuda(q, k, v, o, lse, d_o)
    B, Nq, Dm = q.shape
    Nk = k.shape[1]
    dq = torch.empty(B, Nq, Dm, device=q.device, dtype=torch.bfloat16)
    dk = torch.empty(B, Nk, Dm, device=q.device, dtype=torch.bfloat16)
    dv = torch.empty(B, Nk, Dm, device=q.device, dtype=torch.bfloat16)
    delta = torch.empty(B, heads, Nq, device=q.device, dtype=torch.float32)
    d = _attn_desc(q, k, v, o, lse, heads, scale)
    d.d_o, d.do_bs, d.do_ts = d_o.data_ptr(), d_o.stride(0), d_o.stride(1)
    d.delta = delta.data_ptr()
    d.dq, d.dq_bs, d.dq_ts = dq.data_ptr(), dq.stride(0), dq.stride(1)
    d.dk, d.dk_bs, d.dk_ts = dk.data_ptr(), dk.stride(0), dk.stride(1)
    d.dv, d.dv_bs, d.dv_ts = dv.data_ptr(), dv.stride(0), dv.stride(1)
    keep = None
    if q.dtype == torch.float16:
        keep = list(alts) if alts is not None else [_bf16_copy_same_strides(t) for t in (q, k, v)]
        if all(t is not None for t in keep):
            d.q_alt, d.k_alt, d.v_alt = (t.data_ptr() for t in keep)
    if colsums is not None:
        d.dq_colsum, d.dk_colsum, d.dv_colsum = (_p(t) for t in colsums)
    acc = None
    if max(Nq, Nk) > 256:   # long sequences: fp32 dQ accumulator of the tcgen05 backward
        acc = torch.empty(B, Nq, Dm, device=q.device, dtype=torch.float32)
        d.dq_accum = acc.data_ptr()
    L.check(L.load().b200_attention_bwd(C.byref(d), _stream()), "attention_bwd")
    del keep, acc
    return dq, dk, dv


def kd_loss_fwd(S: torch.Tensor, T_tokens: torch.Tensor, t_skip: int, freq: bool, alpha: float):
    """S fp32 [B,HW,D]; T_tokens fp32 [B,Nt,D] (contiguous). Returns (out[2] = (loss, similarity), ws)."""
    _need_cuda(S, T_tokens)
    B, HW, D = S.shape
    Nt = T_tokens.shape[1]
    ws = torch.empty(int(L.load().b200_kd_loss_ws_floats(B, HW, D)), device=S.device, dtype=torch.float32)
    out = torch.empty(2, device=S.device, dtype=torch.float32)
    L.check(L.load().b200_kd_loss_fwd(S.data_ptr(), T_tokens.data_ptr(), B, HW, D, Nt, t_skip, int(freq), float(alpha),
                                      out.data_ptr(), ws.data_ptr(), _stream()), "kd_loss_fwd")
    return out, ws


def kd_loss_bwd(S, T_tokens, t_skip, freq, alpha, g_out, ws):
    _need_cuda(S, T_tokens, g_out, ws)
    B, HW, D = S.shape
    Nt = T_tokens.shape[1]
    dS = torch.empty_like(S)
    L.check(L.load().b200_kd_loss_bwd(S.data_ptr(), T_tokens.data_ptr(), B, HW, D, Nt, t_skip, int(freq), float(alpha),
                                      g_out.data_ptr(), dS.data_ptr(), 0, ws.data_ptr(), _stream()), "kd_loss_bwd")
    return dS


def dct_zero_dc_idct(x_tokens: torch.Tensor, R: int) -> torch.Tensor:
    """Explicit DCT cross-check: x [B, R*R, D] fp32 contiguous -> same shape."""
    _need_cuda(x_tokens)
    B, HW, D = x_tokens.shape
    y = torch.empty_like(x_tokens)
    L.check(L.load().b200_dct_zero_dc_idct(x_tokens.data_ptr(), y.data_ptr(), B, R, D, HW * D, D, _stream()), "dct")
    return y
