#!/usr/bin/env python
"""A few launches of the attention forward at a teacher shape (profiling target)."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dinov2_distillation_b200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B, heads, hd = 64, 6, 64
D = heads * hd
qkv = (torch.randn(B, N, 3 * D, device="cuda") * 0.5).bfloat16()
q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
for _ in range(4):
    o, lse = ops.attention_fwd(q, k, v, heads, 1 / math.sqrt(hd))
torch.cuda.synchronize()
print("ok", float(o.float().abs().mean()))
