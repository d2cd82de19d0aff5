"""Kernel-level parity on the GPU: each hand-written kernel against a plain PyTorch fp32 statement of the same op."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from dinov2_distillation_b200 import ops
    return ops


def rel_err(a, b):
    a = a.float()
    b = b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def bf(x):
    return x.to(torch.bfloat16)


@pytest.fixture(autouse=True)
def _seed():
    torch.manual_seed(0)


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 384, 384), (514, 1152, 384), (1000, 256, 592),
                                   (16448, 384, 1536), (300, 192, 128), (77, 1536, 384), (2048, 2048, 1024)])
def test_gemm_kmajor_plain(M, N, K):
    ops = _ops()
    a = bf(torch.randn(M, K, device="cuda"))
    b = bf(torch.randn(N, K, device="cuda") / math.sqrt(K))
    ref = a.float() @ b.float().t()
    out = ops.gemm(a, b)
    assert rel_err(out, ref) < 2e-3, rel_err(out, ref)
    out16 = ops.gemm(a, b, out_dtype=torch.bfloat16)
    assert rel_err(out16, ref) < 6e-3


@pytest.mark.parametrize("a_mn,b_mn", [(True, True), (True, False), (False, True)])
@pytest.mark.parametrize("M,N,K", [(384, 384, 4096), (128, 256, 128), (1024, 200, 333 * 8), (384, 1536, 1000)])
def test_gemm_mn_major(a_mn, b_mn, M, N, K):
    ops = _ops()
    a = bf(torch.randn(M, K, device="cuda"))
    b = bf(torch.randn(N, K, device="cuda") / math.sqrt(K))
    ref = a.float() @ b.float().t()
    aa = a.t().contiguous() if a_mn else a
    bb = b.t().contiguous() if b_mn else b
    out = ops.gemm(aa, bb, a_mn_major=a_mn, b_mn_major=b_mn)
    assert rel_err(out, ref) < 2e-3, rel_err(out, ref)


def test_gemm_split_k_atomic():
    ops = _ops()
    M, N, K = 384, 1536, 16384
    a = bf(torch.randn(K, M, device="cuda"))
    b = bf(torch.randn(K, N, device="cuda") / math.sqrt(K))
    ref = a.float().t() @ b.float()
    base = torch.randn(M, N, device="cuda")
    out = base.clone()
    ops.gemm(a, b, a_mn_major=True, b_mn_major=True, out=out, atomic_add=True, split_k=16)
    assert rel_err(out - base, ref) < 2e-3


def test_gemm_epilogues():
    ops = _ops()
    M, N, K = 1030, 768, 384
    a = bf(torch.randn(M, K, device="cuda"))
    b = bf(torch.randn(N, K, device="cuda") / math.sqrt(K))
    bias = torch.randn(N, device="cuda")
    gamma = torch.rand(N, device="cuda") + 0.1
    res = torch.randn(M, N, device="cuda")
    acc = a.float() @ b.float().t() + bias
    # bias + gelu, with pre-activation copy
    out, pre = ops.gemm(a, b, bias=bias, act="gelu", out_dtype=torch.bfloat16, out_pre=True)
    assert rel_err(out, torch.nn.functional.gelu(acc)) < 6e-3
    assert rel_err(pre, acc) < 6e-3
    # bias + relu
    out = ops.gemm(a, b, bias=bias, act="relu", out_dtype=torch.bfloat16)
    assert rel_err(out, torch.relu(acc)) < 6e-3
    # LayerScale + residual, in place on the residual stream
    x = res.clone()
    ops.gemm(a, b, bias=bias, col_scale=gamma, residual=x, out=x)
    assert rel_err(x, res + gamma * acc) < 2e-3
    # aux: multiply by gelu'(aux) / relu mask
    aux = bf(torch.randn(M, N, device="cuda"))
    out = ops.gemm(a, b, aux=aux, aux_mode="dgelu")
    xa = aux.float()
    dg = 0.5 * (1 + torch.erf(xa / math.sqrt(2))) + xa * torch.exp(-0.5 * xa * xa) / math.sqrt(2 * math.pi)
    assert rel_err(out, (a.float() @ b.float().t()) * dg) < 2e-3
    out = ops.gemm(a, b, aux=aux, aux_mode="drelu")
    assert rel_err(out, (a.float() @ b.float().t()) * (xa > 0)) < 2e-3
    # periodic residual + row remap (patch-embed form)
    period, pad = 103, 1
    pos = torch.randn(period, N, device="cuda")
    out = torch.zeros((M // period) * (period + pad), N, device="cuda")
    ops.gemm(a, b, bias=bias, residual=pos, res_row_period=period, out=out, out_row_period=period, out_row_pad=pad)
    ref = (acc.view(M // period, period, N) + pos).reshape(M // period, period, N)
    got = out.view(M // period, period + pad, N)[:, pad:, :]
    assert rel_err(got, ref) < 2e-3
    assert out.view(M // period, period + pad, N)[:, 0, :].abs().max().item() == 0.0


@pytest.mark.parametrize("M,N,K", [(1030, 768, 384), (2048, 512, 2048), (520, 200, 320), (16448, 1152, 384)])
def test_gemm_tma_epilogue_paths(M, N, K):
    """Every epilogue of the bulk-store GEMM (gemm_v2.cu): 16-bit and fp32 outputs, separate and in-place residual,
    16-bit aux by TMA, pre-activation copy, row / column tails, and the CTA-pair mainloop (M >= 1024 and K >= 1024)."""
    ops = _ops()
    a = bf(torch.randn(M, K, device="cuda"))
    b = bf(torch.randn(N, K, device="cuda") / math.sqrt(K))
    bias = torch.randn(N, device="cuda")
    gamma = torch.rand(N, device="cuda") + 0.1
    res = torch.randn(M, N, device="cuda")
    plain = a.float() @ b.float().t()
    acc = plain + bias
    # fp32: plain, bias, separate residual, in-place residual with LayerScale
    assert rel_err(ops.gemm(a, b), plain) < 2e-3
    out = torch.empty(M, N, device="cuda")
    ops.gemm(a, b, bias=bias, residual=res, out=out)
    assert rel_err(out, res + acc) < 2e-3
    ops.gemm(a, b, bias=bias, col_scale=gamma, residual=res, out=out)
    assert rel_err(out, res + gamma * acc) < 2e-3
    x = res.clone()
    ops.gemm(a, b, bias=bias, col_scale=gamma, residual=x, out=x)
    assert rel_err(x, res + gamma * acc) < 2e-3
    # accumulate into an existing fp32 buffer (split-K form with one split)
    y = res.clone()
    ops.gemm(a, b, out=y, atomic_add=True)
    assert rel_err(y - res, plain) < 2e-3
    # 16-bit: bias, relu, gelu (+ pre-activation copy), fp16 output
    assert rel_err(ops.gemm(a, b, bias=bias, out_dtype=torch.bfloat16), acc) < 6e-3
    assert rel_err(ops.gemm(a, b, bias=bias, act="relu", out_dtype=torch.bfloat16), torch.relu(acc)) < 6e-3
    o, pre = ops.gemm(a, b, bias=bias, act="gelu", out_dtype=torch.bfloat16, out_pre=True)
    assert rel_err(o, torch.nn.functional.gelu(acc)) < 6e-3
    assert rel_err(pre, acc) < 6e-3
    o, pre = ops.gemm(a, b, bias=bias, act="relu", out_dtype=torch.bfloat16, out_pre=True)
    assert rel_err(o, torch.relu(acc)) < 6e-3 and rel_err(pre, acc) < 6e-3
    ah, bh = a.to(torch.float16), b.to(torch.float16)
    o16 = ops.gemm(ah, bh, bias=bias, out_dtype=torch.float16)
    assert o16.dtype == torch.float16 and rel_err(o16, ah.float() @ bh.float().t() + bias) < 2e-3
    # final value in both 16-bit formats from one epilogue (the projector FFN saves relu(.) as fp16 and bf16)
    if N % 32 == 0:
        ref16 = torch.relu(ah.float() @ bh.float().t() + bias)
        o16, alt = ops.gemm(ah, bh, bias=bias, act="relu", out_dtype=torch.float16, out_alt=True)
        assert o16.dtype == torch.float16 and alt.dtype == torch.bfloat16
        assert rel_err(o16, ref16) < 2e-3 and rel_err(alt, ref16) < 6e-3
    # 16-bit output x f(aux): dGELU (bf16 aux) and dReLU (fp16 aux)
    aux = bf(torch.randn(M, N, device="cuda"))
    xa = aux.float()
    dg = 0.5 * (1 + torch.erf(xa / math.sqrt(2))) + xa * torch.exp(-0.5 * xa * xa) / math.sqrt(2 * math.pi)
    assert rel_err(ops.gemm(a, b, aux=aux, aux_mode="dgelu", out_dtype=torch.bfloat16), plain * dg) < 6e-3
    auxh = torch.randn(M, N, device="cuda").to(torch.float16)
    assert rel_err(ops.gemm(a, b, aux=auxh, aux_mode="drelu", out_dtype=torch.bfloat16), plain * (auxh.float() > 0)) < 6e-3


def test_gemm_batched_column_output_writes_nchw():
    """dX of the 1x1 conv straight into NCHW: A = W^T [C, D], B = token-major dY [B*HW, D], columns batched with
    period HW through the 3-D output tensor map (plain store and accumulate)."""
    ops = _ops()
    Bt, HW, Cc, D = 5, 256, 512, 384
    wT = bf(torch.randn(Cc, D, device="cuda") / math.sqrt(D))
    dy = bf(torch.randn(Bt * HW, D, device="cuda"))
    ref = torch.einsum("cd,bhd->bch", wT.float(), dy.float().view(Bt, HW, D))
    out = torch.empty(Bt, Cc, HW, device="cuda")
    ops.gemm(wT, dy, out=out, out_batch_period=HW)
    assert rel_err(out, ref) < 2e-3, rel_err(out, ref)
    base = torch.randn(Bt, Cc, HW, device="cuda")
    out2 = base.clone()
    ops.gemm(wT, dy, out=out2, out_batch_period=HW, atomic_add=True)
    assert rel_err(out2 - base, ref) < 2e-3


def test_gelu_epilogue_matches_erf_gelu():
    """The one-MUFU GELU of the 16-bit epilogue against erf-GELU over the whole useful range (identity GEMM)."""
    ops = _ops()
    K = 64
    x = torch.linspace(-12.0, 12.0, 4096 * K, device="cuda").reshape(4096, K)
    a = bf(x)
    eye = bf(torch.eye(K, device="cuda"))
    out = ops.gemm(a, eye, act="gelu", out_dtype=torch.bfloat16).float()
    ref = torch.nn.functional.gelu(a.float())
    # bf16 output rounding (2^-9 relative) dominates; the approximation itself is 1.7e-5 absolute
    assert ((out - ref).abs() <= 4e-3 * ref.abs() + 4e-5).all(), (out - ref).abs().max().item()


def test_gemm_odd_n_tail():
    ops = _ops()
    M, N, K = 200, 72, 64
    a = bf(torch.randn(M, K, device="cuda"))
    b = bf(torch.randn(N, K, device="cuda"))
    bias = torch.randn(N, device="cuda")
    out = ops.gemm(a, b, bias=bias)
    assert rel_err(out, a.float() @ b.float().t() + bias) < 2e-3


# ------------------------------------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("D", [384, 768, 1024, 1536])
def test_layernorm_fwd_bwd(D):
    ops = _ops()
    rows = 1031
    x = torch.randn(rows, D, device="cuda") * 2 + 0.5
    w = torch.randn(D, device="cuda")
    b = torch.randn(D, device="cuda")
    y32, y16, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-6, want_bf16=True, want_stats=True)
    ref = torch.nn.functional.layer_norm(x, (D,), w, b, 1e-6)
    assert (y32 - ref).abs().max().item() < 1e-4
    assert rel_err(y16, ref) < 5e-3
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    br = b.clone().requires_grad_(True)
    dy = torch.randn(rows, D, device="cuda")
    dres = torch.randn(rows, D, device="cuda")
    torch.nn.functional.layer_norm(xr, (D,), wr, br, 1e-6).backward(dy)
    dx, dx16, dw, db = ops.layernorm_bwd(dy, x, w, mean, rstd, dres=dres, want_bf16=True)
    assert rel_err(dx, xr.grad + dres) < 1e-5
    assert rel_err(dw, wr.grad) < 1e-4
    assert rel_err(db, br.grad) < 1e-4
    assert rel_err(dx16, xr.grad + dres) < 5e-3


def test_layernorm_drop_cls_rows():
    ops = _ops()
    B, N, D = 3, 17, 384
    x = torch.randn(B, N, D, device="cuda")
    w = torch.ones(D, device="cuda")
    b = torch.zeros(D, device="cuda")
    y32, _, _, _ = ops.layernorm_fwd(x, w, b, 1e-6, in_period=N - 1, in_pad=1, rows=B * (N - 1))
    ref = torch.nn.functional.layer_norm(x[:, 1:], (D,), w, b, 1e-6).reshape(-1, D)
    assert (y32 - ref).abs().max().item() < 1e-4


@pytest.mark.parametrize("M,K,N,act,pre", [(16448, 384, 1152, "none", False), (16384, 384, 1536, "gelu", True),
                                            (300, 384, 1536, "gelu", False), (1000, 256, 96, "none", True),
                                            (2000, 128, 640, "gelu", False), (4112, 768, 2304, "none", False)])
def test_ln_gemm_fused_prologue_and_fallback(M, K, N, act, pre):
    """b200_ln_gemm_bf16: LayerNorm fused in front of the GEMM as a shared-memory prologue (K <= 384: CTA-pair kernel,
    the panel is normalised once and stays resident; partial last pair / panel at M = 16448, 300, 1000) and the
    LayerNorm + GEMM fallback (K = 768). Must equal the two-kernel path (same per-row arithmetic, operation for operation:
    small batches are dispatched to it) and the fp32 reference."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(M + K + N)
    x = torch.randn(M, K, device="cuda", generator=g) * 2.0 + 0.3
    x[:, ::37] *= 8.0          # a few large-magnitude channels, like a ViT residual stream
    lw = 1.0 + 0.1 * torch.randn(K, device="cuda", generator=g)
    lb = 0.1 * torch.randn(K, device="cuda", generator=g)
    w = bf(torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K))
    bias = 0.1 * torch.randn(N, device="cuda", generator=g)
    ops.set_option("gemm_ln", 2)   # (2: the fused kernel also for row counts the dispatcher leaves to the two-kernel path)
    try:
        out, p, mean, rstd = ops.ln_gemm(x, lw, lb, 1e-6, w, bias=bias, act=act, out_pre=pre, stats=True)
    finally:
        ops.set_option("gemm_ln", 1)
    xn = torch.nn.functional.layer_norm(x, (K,), lw, lb, 1e-6)
    z = bf(xn).float() @ w.float().t() + bias
    ref = torch.nn.functional.gelu(z) if act == "gelu" else z
    assert rel_err(out, ref) < 6e-3, rel_err(out, ref)
    if pre:
        assert rel_err(p, z) < 6e-3, rel_err(p, z)
    assert (mean - x.mean(1)).abs().max().item() < 1e-4
    assert rel_err(rstd, (x.var(1, unbiased=False) + 1e-6).rsqrt()) < 1e-5
    # the unfused path (option off) runs layernorm_fwd + the plain GEMM: same operand bits, same products
    ops.set_option("gemm_ln", 0)
    try:
        out2, p2, _, _ = ops.ln_gemm(x, lw, lb, 1e-6, w, bias=bias, act=act, out_pre=pre, stats=True)
    finally:
        ops.set_option("gemm_ln", 1)
    assert rel_err(out, out2) < 2e-3, rel_err(out, out2)
    if pre:
        assert rel_err(p, p2) < 2e-3


# ------------------------------------------------------------------------------------------------ attention
def _attn_ref(q, k, v, heads, scale):
    B, Nq, D = q.shape
    Nk = k.shape[1]
    hd = D // heads
    qh = q.float().view(B, Nq, heads, hd).transpose(1, 2)
    kh = k.float().view(B, Nk, heads, hd).transpose(1, 2)
    vh = v.float().view(B, Nk, heads, hd).transpose(1, 2)
    p = torch.softmax(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    return (p @ vh).transpose(1, 2).reshape(B, Nq, D)


@pytest.mark.parametrize("hd,heads,Nq,Nk", [(64, 6, 257, 257), (16, 24, 256, 256), (24, 16, 256, 256), (32, 24, 100, 77),
                                            (48, 16, 64, 200), (96, 16, 130, 130), (64, 16, 1370, 1370)])
def test_attention_fwd_bwd(hd, heads, Nq, Nk):
    ops = _ops()
    B = 2
    D = hd * heads
    scale = 5.0 / math.sqrt(hd) if hd != 64 else 1.0 / math.sqrt(hd)
    q = bf(torch.randn(B, Nq, D, device="cuda") * 0.5)
    k = bf(torch.randn(B, Nk, D, device="cuda") * 0.5)
    v = bf(torch.randn(B, Nk, D, device="cuda"))
    o, lse = ops.attention_fwd(q, k, v, heads, scale)
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    ref = _attn_ref(qr, kr, vr, heads, scale)
    assert rel_err(o, ref) < 1e-2, rel_err(o, ref)
    d_o = bf(torch.randn(B, Nq, D, device="cuda"))
    ref.backward(d_o.float())
    dq, dk, dv = ops.attention_bwd(q, k, v, o, lse, d_o, heads, scale)
    assert rel_err(dv, vr.grad) < 2e-2, rel_err(dv, vr.grad)
    assert rel_err(dk, kr.grad) < 2e-2, rel_err(dk, kr.grad)
    assert rel_err(dq, qr.grad) < 2e-2, rel_err(dq, qr.grad)


@pytest.mark.parametrize("B,heads,Nq,Nk", [(3, 6, 256, 256), (5, 6, 257, 257), (2, 12, 200, 200), (2, 3, 264, 264),
                                           (2, 4, 130, 130), (1, 2, 300, 257), (2, 6, 64, 272), (70, 6, 257, 257),
                                           (2, 16, 1370, 1370), (1, 3, 200, 600), (2, 2, 520, 273), (1, 2, 1025, 1024)])
def test_attention_tc_head64(B, heads, Nq, Nk):
    """tcgen05 attention (attention_tc.cu): resident key range (Nk <= 272, one QK^T / PV per query tile) and key blocks
    of 256 with online softmax (518-pixel sequence lengths); full / partial query tiles, tail rows on mma.sync
    (Nq mod 128 <= 8), key padding to 16, ragged last key block, more units than SMs, strided q/k/v out of a fused qkv
    tensor, and the log-sum-exp output."""
    ops = _ops()
    hd = 64
    D = hd * heads
    scale = hd ** -0.5
    if Nq == Nk:
        qkv = bf(torch.randn(B, Nq, 3 * D, device="cuda") * 0.7)
        q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    else:
        q = bf(torch.randn(B, Nq, D, device="cuda") * 0.7)
        k = bf(torch.randn(B, Nk, D, device="cuda") * 0.7)
        v = bf(torch.randn(B, Nk, D, device="cuda"))
    o, lse = ops.attention_fwd(q, k, v, heads, scale)
    ref = _attn_ref(q.float(), k.float(), v.float(), heads, scale)
    assert rel_err(o, ref) < 6e-3, rel_err(o, ref)
    qh = q.float().reshape(B, Nq, heads, hd).transpose(1, 2)
    kh = k.float().reshape(B, Nk, heads, hd).transpose(1, 2)
    lse_ref = torch.logsumexp(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    assert (lse - lse_ref).abs().max().item() < 2e-3, (lse - lse_ref).abs().max().item()


@pytest.mark.parametrize("B,heads,Nq,Nk,shared_q", [(2, 16, 1369, 1369, False), (2, 16, 1369, 1369, True), (3, 4, 300, 273, True),
                                                    (2, 2, 520, 1024, False)])
def test_attention_tc_head64_fp16_both_formats_shared_query(B, heads, Nq, Nk, shared_q):
    """The long-sequence tcgen05 forward (attention_tc.cu) on the ScaleKD projector's operands at 518 pixels: fp16 q / k / v,
    both output formats, the batch-invariant self-query (batch stride 0), k / v sliced out of the fused [k|v] tensor."""
    ops = _ops()
    hd = 64
    D = hd * heads
    scale = 5.0 / math.sqrt(hd)
    kv = (torch.randn(B, Nk, 2 * D, device="cuda") * 0.5).half()
    k, v = kv[..., :D], kv[..., D:]
    q = (torch.randn(1 if shared_q else B, Nq, D, device="cuda") * 0.5).half()
    if shared_q:
        q = q.expand(B, Nq, D)
    ops.set_option("stat_attn_tc_fwd", 0)
    o, lse, o_alt = ops.attention_fwd(q, k, v, heads, scale, want_alt=True)
    assert ops.get_option("stat_attn_tc_fwd") == 1
    assert o.dtype == torch.float16 and o_alt.dtype == torch.bfloat16
    ref = _attn_ref(q, k, v, heads, scale)
    assert rel_err(o, ref) < 2e-3, rel_err(o, ref)
    assert rel_err(o_alt, ref) < 8e-3, rel_err(o_alt, ref)
    qh = q.float().reshape(B, Nq, heads, hd).transpose(1, 2)
    kh = k.float().reshape(B, Nk, heads, hd).transpose(1, 2)
    lse_ref = torch.logsumexp(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    assert (lse - lse_ref).abs().max().item() < 2e-3


@pytest.mark.parametrize("hd,heads,N", [(16, 24, 256), (24, 16, 256), (64, 4, 150), (96, 2, 70)])
def test_attention_fp16_forward_bf16_grads(hd, heads, N):
    """ScaleKD projector precision policy: q/k/v/o fp16, gradients bf16."""
    ops = _ops()
    B = 2
    D = hd * heads
    scale = 5.0 / math.sqrt(hd)
    q = (torch.randn(B, N, D, device="cuda") * 0.5).half()
    k = (torch.randn(B, N, D, device="cuda") * 0.5).half()
    v = torch.randn(B, N, D, device="cuda").half()
    o, lse = ops.attention_fwd(q, k, v, heads, scale)
    assert o.dtype == torch.float16
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    ref = _attn_ref(qr, kr, vr, heads, scale)
    assert rel_err(o, ref) < 2e-3, rel_err(o, ref)
    d_o = bf(torch.randn(B, N, D, device="cuda") * 1e-4)   # tiny gradients: would underflow in fp16
    ref.backward(d_o.float())
    dq, dk, dv = ops.attention_bwd(q, k, v, o, lse, d_o, heads, scale)
    assert dq.dtype == torch.bfloat16
    assert rel_err(dv, vr.grad) < 1e-2, rel_err(dv, vr.grad)
    assert rel_err(dk, kr.grad) < 1e-2, rel_err(dk, kr.grad)
    assert rel_err(dq, qr.grad) < 1e-2, rel_err(dq, qr.grad)


@pytest.mark.parametrize("dt", ["fp16", "bf16"])
@pytest.mark.parametrize("hd,heads,Nq,Nk,B", [(16, 24, 256, 256, 3), (24, 16, 256, 256, 3), (32, 24, 100, 77, 2),
                                              (48, 16, 64, 200, 2), (64, 6, 256, 256, 5), (24, 16, 64, 64, 9),
                                              (16, 24, 130, 16, 2), (64, 2, 300, 129, 2), (40, 3, 17, 5, 1),
                                              (24, 16, 256, 256, 40)])
def test_attention_pp_forward(dt, hd, heads, Nq, Nk, B):
    """Two-tile tcgen05 forward (attention_pp.cu): every projector head dim (16 / 24 / 32 / 48 / 64, padded to 32 / 64 by
    the TMA zero fill), fp16 and bf16, whole-grid and window-sized sequences, ragged query / key counts (partial tiles,
    key counts that are not a multiple of 16, a single 128-column half), odd tile counts per CTA, more units than SMs,
    both output formats and the log-sum-exp; the launch must have taken the new path."""
    ops = _ops()
    tdt = torch.float16 if dt == "fp16" else torch.bfloat16
    D = hd * heads
    scale = 5.0 / math.sqrt(hd)
    q = (torch.randn(B, Nq, D, device="cuda") * 0.5).to(tdt)
    k = (torch.randn(B, Nk, D, device="cuda") * 0.5).to(tdt)
    v = torch.randn(B, Nk, D, device="cuda").to(tdt)
    ops.set_option("stat_attn_pp_fwd", 0)
    ops.set_option("attn_pp_fwd", 2)   # (also the window-sized shapes the dispatcher leaves to the flash-style kernel)
    try:
        o, lse, o_alt = ops.attention_fwd(q, k, v, heads, scale, want_alt=True)
    finally:
        ops.set_option("attn_pp_fwd", 1)
    assert ops.get_option("stat_attn_pp_fwd") == 1
    assert o.dtype == tdt and o_alt.dtype != tdt
    ref = _attn_ref(q, k, v, heads, scale)
    tol = 2e-3 if dt == "fp16" else 8e-3
    assert rel_err(o, ref) < tol, rel_err(o, ref)
    assert rel_err(o_alt, ref) < 8e-3, rel_err(o_alt, ref)
    qh = q.float().reshape(B, Nq, heads, hd).transpose(1, 2)
    kh = k.float().reshape(B, Nk, heads, hd).transpose(1, 2)
    lse_ref = torch.logsumexp(qh @ kh.transpose(-1, -2) * scale, dim=-1)
    assert (lse - lse_ref).abs().max().item() < 2e-3, (lse - lse_ref).abs().max().item()
    # against the mma.sync kernel on the same inputs (same exp2-domain arithmetic): near bit-equal
    ops.set_option("attn_pp_fwd", 0)
    try:
        o2, lse2 = ops.attention_fwd(q, k, v, heads, scale)
    finally:
        ops.set_option("attn_pp_fwd", 1)
    assert rel_err(o, o2) < (1.5e-3 if dt == "fp16" else 6e-3)
    assert (lse - lse2).abs().max().item() < 1e-4


def test_attention_pp_batch_invariant_query_and_strided_kv():
    """The self-query embedding (scalekd.py:232-234) is one [HW, D] tensor for the whole batch (batch stride 0); k / v are
    column slices of the fused [k|v] projection output."""
    ops = _ops()
    B, N, heads, hd = 5, 256, 16, 24
    D = heads * hd
    kv = (torch.randn(B, N, 2 * D, device="cuda") * 0.5).half()
    k, v = kv[..., :D], kv[..., D:]
    qs = (torch.randn(1, N, D, device="cuda") * 0.5).half().expand(B, N, D)
    ops.set_option("stat_attn_pp_fwd", 0)
    o, lse = ops.attention_fwd(qs, k, v, heads, 3.0 / math.sqrt(hd))
    assert ops.get_option("stat_attn_pp_fwd") == 1
    assert rel_err(o, _attn_ref(qs, k, v, heads, 3.0 / math.sqrt(hd))) < 2e-3


def test_attention_strided_qkv_and_shared_query():
    ops = _ops()
    B, N, heads, hd = 3, 70, 6, 64
    D = heads * hd
    qkv = bf(torch.randn(B, N, 3 * D, device="cuda") * 0.5)
    q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
    o, _ = ops.attention_fwd(q, k, v, heads, hd ** -0.5)
    assert rel_err(o, _attn_ref(q, k, v, heads, hd ** -0.5)) < 1e-2
    qs = bf(torch.randn(1, N, D, device="cuda") * 0.5).expand(B, N, D)   # batch stride 0
    o, _ = ops.attention_fwd(qs, k, v, heads, hd ** -0.5)
    assert rel_err(o, _attn_ref(qs, k, v, heads, hd ** -0.5)) < 1e-2


# ------------------------------------------------------------------------------------------------ loss terms
def _ref_loss(S, T_map, alpha, freq):
    import torch.nn.functional as F
    N, C, H, W = T_map.shape
    s = S.permute(0, 2, 1).contiguous().view(N, C, H, W)
    t = T_map
    if freq:
        s = s - s.mean(dim=(2, 3), keepdim=True)
        t = t - t.mean(dim=(2, 3), keepdim=True)
    s = F.normalize(s, dim=1)
    t = F.normalize(t, dim=1)
    loss = ((s - t) ** 2).sum() / N * alpha
    sim = F.cosine_similarity(s, t, dim=1).mean()
    return loss, sim


@pytest.mark.parametrize("freq", [False, True])
@pytest.mark.parametrize("B,R,D", [(2, 16, 384), (3, 37, 1024)])
def test_kd_loss_fwd_bwd(freq, B, R, D):
    ops = _ops()
    HW = R * R
    S = torch.randn(B, HW, D, device="cuda")
    T = torch.randn(B, HW + 1, D, device="cuda") + 0.3
    T_map = T[:, 1:].reshape(B, R, R, D).permute(0, 3, 1, 2)
    out, ws = ops.kd_loss_fwd(S, T, 1, freq, 0.08)
    Sr = S.clone().requires_grad_(True)
    loss, sim = _ref_loss(Sr, T_map, 0.08, freq)
    assert abs(out[0].item() - loss.item()) / abs(loss.item()) < 1e-4
    assert abs(out[1].item() - sim.item()) < 1e-4
    g = torch.tensor([1.7, 0.0], device="cuda")
    (loss * 1.7).backward()
    dS = ops.kd_loss_bwd(S, T, 1, freq, 0.08, g, ws)
    assert rel_err(dS, Sr.grad) < 1e-3, rel_err(dS, Sr.grad)


def test_dct_identity_matches_mean_subtraction():
    ops = _ops()
    for R in (16, 37):
        x = torch.randn(2, R * R, 40, device="cuda")
        y = ops.dct_zero_dc_idct(x, R)
        ref = x - x.mean(dim=1, keepdim=True)
        assert (y - ref).abs().max().item() < 2e-4


# ------------------------------------------------------------------------------------------------ layout kernels
def test_im2col_matches_unfold():
    ops = _ops()
    img = torch.randn(2, 3, 56, 70, device="cuda")
    got = ops.patch_im2col(img).float()
    ref = torch.nn.functional.unfold(img, 14, stride=14).transpose(1, 2).reshape(-1, 588)
    assert (got[:, :588] - ref).abs().max().item() < 2e-2
    assert got[:, 588:].abs().max().item() == 0


def test_token_layout_roundtrip():
    ops = _ops()
    x = torch.randn(3, 72, 5, 7, device="cuda")
    t16, t32 = ops.nchw_to_tokens(x, want_f32=True)
    ref = x.flatten(2).transpose(1, 2).reshape(-1, 72)
    assert (t32 - ref).abs().max().item() == 0
    assert rel_err(t16, ref) < 5e-3
    back = ops.tokens_to_nchw(t32, 3, 35)
    assert (back.view_as(x) - x).abs().max().item() == 0
    w = torch.randn(96, 40, device="cuda")
    s = torch.rand(96, device="cuda")
    assert rel_err(ops.transpose_bf16(w, s), (w * s[:, None]).t()) < 5e-3


@pytest.mark.parametrize("adt,bdt", [(torch.float16, torch.float16)])
@pytest.mark.parametrize("mn", [False, True])
def test_gemm_fp16_operands(adt, bdt, mn):
    """Projector forward runs fp16 operands (A and B must share the format: tcgen05 kind::f16 traps on a mix)."""
    ops = _ops()
    M, N, K = 384, 512, 2048
    a = torch.randn(M, K, device="cuda").to(adt)
    b = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(bdt)
    ref = a.float() @ b.float().t()
    aa = a.t().contiguous() if mn else a
    bb = b.t().contiguous() if mn else b
    out = ops.gemm(aa, bb, a_mn_major=mn, b_mn_major=mn)
    assert rel_err(out, ref) < 1e-3, rel_err(out, ref)
    out16 = ops.gemm(aa, bb, a_mn_major=mn, b_mn_major=mn, out_dtype=torch.float16, act="relu")
    assert out16.dtype == torch.float16
    assert rel_err(out16, torch.relu(ref)) < 2e-3


# ------------------------------------------------------------------------------------------------ optimizer (next row f2)
@pytest.mark.parametrize("clip", [None, 1.0])
def test_arena_adamw_matches_torch_adamw_with_global_norm_clip(clip):
    """ArenaAdamW (two launches over flat arenas) == torch.optim.AdamW after torch.nn.utils.clip_grad_norm_, including
    the student's share of the global norm passed in as a device scalar."""
    from dinov2_distillation_b200.distributed import FlatGradArena
    from dinov2_distillation_b200.optim import ArenaAdamW
    torch.manual_seed(5)
    shapes = [(384, 512, 1, 1), (384,), (1, 384, 16, 16), (1536, 384), (7,)]
    ours = [torch.nn.Parameter(torch.randn(s, device="cuda") * 0.1) for s in shapes]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ours]
    student = torch.nn.Parameter(torch.randn(1000, device="cuda"))      # not in the arena: only its norm takes part
    arena = FlatGradArena(ours)
    opt = ArenaAdamW(arena, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.01, max_grad_norm=clip)
    topt = torch.optim.AdamW(ref, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.01)
    for it in range(4):
        arena.zero()
        scale = 10.0 if it % 2 == 0 else 0.01       # clipped and unclipped steps
        for p, r in zip(ours, ref):
            gr = torch.randn_like(p) * scale
            p.grad.copy_(gr)
            r.grad = gr.clone()
        student.grad = torch.randn_like(student) * scale
        extra = (student.grad ** 2).sum().reshape(1)
        if clip is not None:
            torch.nn.utils.clip_grad_norm_(ref + [student], clip)
        topt.step()
        opt.step(extra_sq_norm=extra if clip is not None else None)
        if clip is not None and it % 2 == 0:   # the arena's share of the (pre-clip) global norm, reduced on the device
            want = math.sqrt(sum((p_.grad.float() ** 2).sum().item() for p_ in ours))
            assert abs(opt.grad_norm().item() - want) <= 1e-4 * want
    for p, r in zip(ours, ref):
        assert rel_err(p.data, r.data) < 1e-6, rel_err(p.data, r.data)
        assert p.data.data_ptr() >= opt.flat_params.data_ptr()


@pytest.mark.parametrize("hw,HW", [((7, 7), (16, 16)), ((14, 14), (16, 16)), ((17, 17), (37, 37)), ((33, 33), (37, 37)),
                                   ((5, 9), (8, 8)), ((20, 20), (9, 9))])
def test_bilinear_token_resize_and_adjoint(hw, HW):
    """Token-major bilinear resize (align_corners=False, models/model_zoo.py:121-126) against oracle/resize_ref.py, and
    its adjoint both against the oracle's transpose and through the inner-product identity <R x, g> == <x, R^T g>."""
    from oracle import resize_ref
    ops = _ops()
    (h, w), (H, W) = hw, HW
    B, D = 3, 72
    g0 = torch.Generator().manual_seed(h * 1000 + H)
    x = torch.randn(B, D, h, w, generator=g0)
    ref = resize_ref.resize_bilinear(x.double(), HW).float()                       # [B, D, H, W]
    src = x.flatten(2).transpose(1, 2).contiguous().cuda()                          # [B, h*w, D]
    got = ops.bilinear_tokens(src, hw, HW)
    assert rel_err(got.cpu(), ref.flatten(2).transpose(1, 2)) < 1e-5     # fp32 source coordinates, like aten
    g = bf(torch.randn(B, H * W, D, generator=g0).cuda())
    adj = ops.bilinear_tokens_adjoint(g, hw, HW)
    Ry = resize_ref.resize_matrix(h, H).float()
    Rx = resize_ref.resize_matrix(w, W).float()
    gm = g.float().cpu().transpose(1, 2).reshape(B, D, H, W)
    adj_ref = torch.einsum("yh,bcyx,xw->bchw", Ry, gm, Rx).flatten(2).transpose(1, 2)
    assert rel_err(adj.cpu(), adj_ref) < 4e-3                                       # bf16 output rounding
    lhs = (got.double() * g.double()).sum().item()
    rhs = (src.double() * adj.double()).sum().item()
    assert abs(lhs - rhs) <= 2e-3 * max(abs(lhs), (got.double().norm() * g.double().norm()).item() * 1e-2)


def test_window_row_order_round_trip():
    """b200_window_rows16 against separate_tokens' view/permute (losses/scalekd.py:326-335) and back."""
    import ctypes as C
    from dinov2_distillation_b200 import _lib as L
    lib = L.load()
    B, H, W, D, wh, ww = 3, 8, 12, 40, 2, 3
    x = bf(torch.randn(B, H * W, D, device="cuda"))
    y = torch.empty_like(x)
    st = torch.cuda.current_stream().cuda_stream
    L.check(lib.b200_window_rows16(x.data_ptr(), y.data_ptr(), B * H * W, H, W, wh, ww, D, D, 0, st), "window_rows16")
    ref = x.view(B, wh, H // wh, ww, W // ww, D).permute(0, 1, 3, 2, 4, 5).reshape(B, H * W, D)
    assert torch.equal(y, ref)
    z = torch.zeros_like(x)
    L.check(lib.b200_window_rows16(y.data_ptr(), z.data_ptr(), B * H * W, H, W, wh, ww, D, D, 1, st), "window_rows16")
    assert torch.equal(z, x)


@pytest.mark.parametrize("B,period,N,K", [(5, 256, 384, 608), (3, 64, 160, 96)])
def test_gemm_row_remap_aligned_period_bulk_store(B, period, N, K):
    """Patch-embed form on the bulk-store kernel: token rows written around the cls gap through a 3-D output tensor map,
    pos_embed added as a residual that repeats every `period` rows (period a multiple of the 32-row store unit)."""
    ops = _ops()
    M, pad = B * period, 1
    a = bf(torch.randn(M, K, device="cuda"))
    b = bf(torch.randn(N, K, device="cuda") / math.sqrt(K))
    bias = torch.randn(N, device="cuda")
    pos = torch.randn(period, N, device="cuda")
    out = torch.full((B * (period + pad), N), 7.0, device="cuda")
    ops.gemm(a, b, bias=bias, residual=pos, res_row_period=period, out=out, out_row_period=period, out_row_pad=pad)
    ref = (a.float() @ b.float().t() + bias).view(B, period, N) + pos
    o3 = out.view(B, period + pad, N)
    assert rel_err(o3[:, pad:, :], ref) < 2e-3
    assert (o3[:, 0, :] == 7.0).all()        # cls rows untouched


@pytest.mark.parametrize("M,N,K", [(16384, 1536, 384), (1000, 96, 64)])
def test_gemm_epilogue_column_sums_of_16bit_output(M, N, K):
    """out16_colsum: the bias gradient that belongs to an input-gradient GEMM (ffn1: colsum((du W2) * relu'(h))),
    accumulated from the staged 16-bit tile -- equals the column sums of the stored output, rows past M excluded."""
    ops = _ops()
    a = bf(torch.randn(M, K, device="cuda"))
    b = bf(torch.randn(N, K, device="cuda") / math.sqrt(K))
    aux = torch.randn(M, N, device="cuda").to(torch.float16)
    acc0 = torch.randn(N, device="cuda")
    cs = acc0.clone()
    out = ops.gemm(a, b, aux=aux, aux_mode="drelu", out_dtype=torch.bfloat16, out_colsum=cs)
    assert rel_err(out, (a.float() @ b.float().t()) * (aux.float() > 0)) < 6e-3
    assert rel_err(cs - acc0, out.float().sum(0)) < 1e-4
    cs2 = torch.zeros(N, device="cuda")
    o2 = ops.gemm(a, b, out_dtype=torch.float16, out_colsum=cs2)
    assert rel_err(cs2, o2.float().sum(0)) < 1e-4


@pytest.mark.parametrize("M,N,K", [(16400, 1536, 384), (40000, 384, 256), (16384, 768, 128), (300, 192, 64)])
@pytest.mark.parametrize("mode", ["drelu", "dgelu"])
def test_gemm_aux_epilogue_shapes(M, N, K, mode):
    """dgrad-through-activation GEMMs (x relu mask / x gelu' from a 16-bit aux tile fetched by TMA per 32 x 32 unit) at the
    shapes of the step and around them: several tiles per CTA, a ragged last row tile, one tile per CTA, a single tile."""
    ops = _ops()
    a = bf(torch.randn(M, K, device="cuda"))
    b = bf(torch.randn(N, K, device="cuda") / math.sqrt(K))
    aux = torch.randn(M, N, device="cuda").to(torch.float16 if mode == "drelu" else torch.bfloat16)
    plain = a.float() @ b.float().t()
    xa = aux.float()
    if mode == "drelu":
        ref = plain * (xa > 0)
    else:
        ref = plain * (0.5 * (1 + torch.erf(xa / math.sqrt(2))) + xa * torch.exp(-0.5 * xa * xa) / math.sqrt(2 * math.pi))
    out = ops.gemm(a, b, aux=aux, aux_mode=mode, out_dtype=torch.bfloat16)
    assert rel_err(out, ref) < 6e-3, rel_err(out, ref)
    # twice in a row on the same buffers: barrier phases start clean in every launch
    out2 = ops.gemm(a, b, aux=aux, aux_mode=mode, out_dtype=torch.bfloat16)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("B,HW,D", [(16, 256, 384), (8, 1369, 256), (2, 49, 128)])
def test_batchnorm_relu_pos_kernels_vs_autograd(B, HW, D):
    """proj_student's BatchNorm2d -> ReLU, + pos_embed (losses/scalekd.py:199-201, :238) on token-major data: forward
    (stats, finalize, apply) and backward (reduce -- incl. the path that owns all images of a position and emits the
    pos_embed gradient from the same pass, and the fallback for few images -- and apply) against torch autograd."""
    import ctypes as C
    from dinov2_distillation_b200 import _lib as L
    lib = L.load()
    st = torch.cuda.current_stream().cuda_stream
    M = B * HW
    gen = torch.Generator(device="cuda").manual_seed(B * 1000 + HW)
    y = torch.randn(M, D, device="cuda", generator=gen)
    w = torch.rand(D, device="cuda", generator=gen) + 0.5
    b = torch.randn(D, device="cuda", generator=gen) * 0.3
    pos = torch.randn(HW, D, device="cuda", generator=gen)
    dz = torch.randn(M, D, device="cuda", generator=gen)
    eps = 1e-5
    # reference: autograd through batch-statistics BN
    yr, wr, br, pr = (t.clone().double().requires_grad_(True) for t in (y, w, b, pos))
    mu, var = yr.mean(0), yr.var(0, unbiased=False)
    z_ref = torch.relu((yr - mu) / torch.sqrt(var + eps) * wr + br) + pr.repeat(B, 1)
    z_ref.backward(dz.double())
    # ours
    sums = torch.zeros(2 * D, device="cuda")
    mean, rstd = torch.empty(D, device="cuda"), torch.empty(D, device="cuda")
    rm, rv = torch.zeros(D, device="cuda"), torch.ones(D, device="cuda")
    L.check(lib.b200_bn_stats(y.data_ptr(), sums.data_ptr(), M, D, st), "bn_stats")
    L.check(lib.b200_bn_finalize(sums.data_ptr(), mean.data_ptr(), rstd.data_ptr(), rm.data_ptr(), rv.data_ptr(), 0.1, eps, M, D, st), "fin")
    z = torch.empty(M, D, device="cuda")
    L.check(lib.b200_bn_relu_pos_fwd(y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), w.data_ptr(), b.data_ptr(), pos.data_ptr(),
                                     z.data_ptr(), None, M, D, HW, 0, st), "fwd")
    assert rel_err(z, z_ref.float()) < 1e-5
    assert rel_err(rm, 0.1 * mu.float()) < 1e-4
    sums2 = torch.zeros(2 * D, device="cuda")
    dpos = torch.zeros(HW, D, device="cuda")
    L.check(lib.b200_bn_relu_pos_bwd_reduce(dz.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), w.data_ptr(), b.data_ptr(),
                                            sums2.data_ptr(), dpos.data_ptr(), M, D, HW, st), "reduce")
    assert rel_err(dpos, pr.grad.float()) < 1e-5
    assert rel_err(sums2[:D], br.grad.float()) < 1e-4          # d beta
    assert rel_err(sums2[D:], wr.grad.float()) < 1e-4          # d gamma
    dy16 = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    L.check(lib.b200_bn_relu_pos_bwd_apply(dz.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(), w.data_ptr(), b.data_ptr(),
                                           sums2.data_ptr(), dy16.data_ptr(), 1, M, D, st), "apply")
    assert rel_err(dy16, yr.grad.float()) < 4e-3               # bf16 output
