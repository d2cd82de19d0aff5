#!/usr/bin/env python
"""One GEMM of a teacher shape (debug / profiling target): python tools/gemm_once.py qkv|fc1|fc1gelu|proj|fc2"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dinov2_distillation_b200 import ops
which = sys.argv[1] if len(sys.argv) > 1 else "fc1"
M = 16448
N, K, kw = {"qkv": (1152, 384, {}), "fc1": (1536, 384, {}), "fc1gelu": (1536, 384, dict(act="gelu")),
            "proj": (384, 384, dict(res=True)), "fc2": (384, 1536, dict(res=True))}[which]
a = torch.randn(M, K, device="cuda").bfloat16()
b = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
bias = torch.randn(N, device="cuda")
if kw.get("res"):
    x = torch.randn(M, N, device="cuda")
    for _ in range(3):
        ops.gemm(a, b, bias=bias, residual=x, out=x)
else:
    for _ in range(3):
        ops.gemm(a, b, bias=bias, act=kw.get("act", "none"), out_dtype=torch.bfloat16)
torch.cuda.synchronize()
print("ok")
