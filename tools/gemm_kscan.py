#!/usr/bin/env python
"""GEMM time against K at the fc1 shape (M=16448, N=1536): separates the per-tile fixed cost from the K-proportional
mainloop cost. Run under B200_GEMM_DBG=0/1/4/5 and B200_GEMM_2CTA=0/1 to split mainloop, epilogue and pairing."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dinov2_distillation_b200 import ops
from gemm_bench import bench

M = int(os.environ.get("M", 16448)); N = int(os.environ.get("N", 1536))
for K in (128, 384, 768, 1536, 3072):
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    us = bench(lambda: ops.gemm(a, b, out=out))
    print(f"M={M} N={N} K={K:5d} {us:8.1f} us {2.0*M*N*K/us/1e6:7.1f} TF/s")
