"""Kernels for one `ncu --set full` capture (profiles/r02_*): python tools/prof_kernels.py dense|hbm
dense: LayerNorm-prologue GEMM and its unfused pair, tcgen05 attention backward (resident + streaming), attention forward.
hbm  : every bandwidth-bound kernel of the path at cfg2 (vits14, B=64 @224) and cfg4 (vitl14 widths, 37 x 37 tokens) sizes,
       launched in situ by a ScaleKD forward / backward and a re-used teacher block."""
import math
import sys
import warnings

import torch

sys.path.insert(0, ".")
warnings.simplefilter("ignore")
from dinov2_distillation_b200 import ops  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "dense"


def attn(B, heads, hd, N, half):
    D = heads * hd
    dt = torch.float16 if half else torch.bfloat16
    scale = (5.0 if hd != 64 else 1.0) / math.sqrt(hd)
    q = (torch.randn(B, N, D, device="cuda") * 0.5).to(dt)
    k = (torch.randn(B, N, D, device="cuda") * 0.5).to(dt)
    v = torch.randn(B, N, D, device="cuda").to(dt)
    d_o = torch.randn(B, N, D, device="cuda").bfloat16()
    o, lse = ops.attention_fwd(q, k, v, heads, scale)
    ops.attention_bwd(q, k, v, o, lse, d_o, heads, scale)


if mode == "dense":
    for M, K, N, act in ((16448, 384, 1152, "none"), (16448, 384, 1536, "gelu")):
        x = torch.randn(M, K, device="cuda")
        lw, lb = torch.ones(K, device="cuda"), torch.zeros(K, device="cuda")
        w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
        bias = torch.zeros(N, device="cuda")
        for on in (1, 0):
            ops.set_option("gemm_ln", on)
            for _ in range(2):
                ops.ln_gemm(x, lw, lb, 1e-6, w, bias=bias, act=act)
        ops.set_option("gemm_ln", 1)
    attn(64, 6, 64, 256, False)
    attn(64, 6, 64, 257, False)
    attn(64, 24, 16, 256, True)
    attn(64, 16, 24, 256, True)
    attn(8, 16, 64, 1369, False)
else:
    from dinov2_distillation_b200.scalekd import ScaleKD
    from dinov2_distillation_b200.teacher import DINOv2ViT
    for D, g, Cs, heads, B, name in ((384, 16, 512, 16, 64, "dinov2_vits14"), (1024, 37, 384, 16, 32, "dinov2_vitl14")):
        torch.manual_seed(0)
        m = ScaleKD(name="scalekd_res5", alpha=[0.08, 0.06], student_dims=Cs, teacher_dims=D, query_hw=[g, g], pos_hw=[g, g],
                    pos_dims=D, window_shapes=[1, 1], self_query=True, softmax_scale=[5.0, 5.0], num_heads=heads).cuda().train()
        S = torch.randn(B, Cs, g, g, device="cuda", requires_grad=True)
        T = torch.randn(B, D, g, g, device="cuda")
        for _ in range(2):
            m(S, T)["loss"].backward()
        t = DINOv2ViT(name, weights="synthetic").cuda().eval()
        f = torch.randn(B, g * g, D, device="cuda", requires_grad=True)
        blk = len(t.model.blocks) - 2
        for _ in range(2):
            t.model.blocks[blk](f).sum().backward()
        img = torch.randn(min(B, 16), 3, g * 14, g * 14, device="cuda")
        t(img)
torch.cuda.synchronize()
print("done", mode)
