#!/usr/bin/env python
"""dgrad-through-activation GEMMs (M = 16 384, K = 384 -> N = 1 536, x relu-mask / x gelu'(pre) from a 16-bit aux tile):
8 vs 16 epilogue warps (option gemm_ew), against the same GEMM without the aux operand."""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dinov2_distillation_b200 import ops
from tools.gemm_bench import bench
M, K, N = 16384, 384, 1536
a = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
aux = torch.randn(M, N, device="cuda").bfloat16()
cs = torch.zeros(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for ew in (8, 16):
    ops.set_option("gemm_ew", ew)
    for name, kw in (("plain bf16 out", {}), ("x relu mask", dict(aux=aux, aux_mode="drelu")), ("x gelu'", dict(aux=aux, aux_mode="dgelu")),
                     ("x relu mask + colsum", dict(aux=aux, aux_mode="drelu", out_colsum=cs))):
        us = bench(lambda: ops.gemm(a, w, out=out, **kw))
        print(f"EW={ew:2d} {name:24s}: {us:6.1f} us  {2.0 * M * N * K / us / 1e6:6.0f} TFLOP/s")
ops.set_option("gemm_ew", 16)
