#!/usr/bin/env python
"""A few launches of the attention backward at a projector / stage shape (profiling target)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dinov2_distillation_b200 import ops
hd = int(sys.argv[1]) if len(sys.argv) > 1 else 16
heads = {16: 24, 24: 16, 64: 6}[hd]
B, N = 64, 256
D = heads * hd
dt = torch.bfloat16 if hd == 64 else torch.float16
qkv = (torch.randn(B, N, 3 * D, device="cuda") * 0.5).to(dt)
q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
scale = (1.0 if hd == 64 else 5.0) / math.sqrt(hd)
o, lse = ops.attention_fwd(q, k, v, heads, scale)
d_o = torch.randn(B, N, D, device="cuda").bfloat16()
for _ in range(3):
    dq, dk, dv = ops.attention_bwd(q, k, v, o, lse, d_o, heads, scale)
torch.cuda.synchronize()
print("ok", float(dq.float().abs().mean()))
