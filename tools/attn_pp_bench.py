#!/usr/bin/env python
"""A/B of the two-tile tcgen05 attention forward (attention_pp.cu) against the paths it replaces, on the shapes of the
distillation step (device time via graph replay): projector cross-attention (fp16, head dims 16 / 24 / 32 / 48) and the
re-used teacher blocks (bf16, head_dim 64, N = 256)."""
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dinov2_distillation_b200 import ops
from tools.gemm_bench import bench


def main():
    shapes = [("projector res4 hd24 (cfg2)", 64, 16, 24, 256, torch.float16, 5.0, True),
              ("projector res5 hd16 (cfg2)", 64, 24, 16, 256, torch.float16, 5.0, True),
              ("projector hd48 (cfg3)", 32, 16, 48, 256, torch.float16, 5.0, True),
              ("projector hd32 (cfg3)", 32, 24, 32, 256, torch.float16, 5.0, True),
              ("stage blocks hd64 N=256", 64, 6, 64, 256, torch.bfloat16, 1.0, False),
              ("stage blocks vitb hd64 B=32", 32, 12, 64, 256, torch.bfloat16, 1.0, False),
              ("window 2x2 hd24 (64 tokens)", 256, 16, 24, 64, torch.float16, 5.0, True)]
    for name, B, heads, hd, N, dt, ss, alt in shapes:
        D = heads * hd
        q = (torch.randn(B, N, D, device="cuda") * 0.5).to(dt)
        kv = (torch.randn(B, N, 2 * D, device="cuda") * 0.5).to(dt)
        k, v = kv[..., :D], kv[..., D:]
        scale = ss / math.sqrt(hd)
        res = []
        for on in (1, 0):
            ops.set_option("attn_pp_fwd", on)
            ops.attention_fwd(q, k, v, heads, scale, want_alt=alt)
            res.append(bench(lambda: ops.attention_fwd(q, k, v, heads, scale, want_alt=alt)))
        ops.set_option("attn_pp_fwd", 1)
        fl = 4.0 * B * heads * N * N * hd
        exps = B * heads * N * N
        print(f"{name:30s} B={B} h={heads} hd={hd} N={N}: two-tile {res[0]:7.1f} us ({fl/res[0]/1e6:6.1f} TF/s, "
              f"{exps/res[0]/1e6/148/1.965e-3:5.2f} ex2/clk/SM) | previous path {res[1]:7.1f} us")


if __name__ == "__main__":
    main()
