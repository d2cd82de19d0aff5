#!/usr/bin/env python
"""LayerNorm forward / backward at the teacher shape (graph-replay device time). Env: B200_LN_BPS, B200_LN_WPB."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dinov2_distillation_b200 import ops
from gemm_bench import bench
M, D = 16448, 384
xs = [torch.randn(M, D, device="cuda") for _ in range(6)]    # rotate inputs: 25 MB each, like the residual stream
w, b = torch.rand(D, device="cuda") + 0.5, torch.randn(D, device="cuda")
i = [0]
def f():
    i[0] = (i[0] + 1) % len(xs)
    ops.layernorm_fwd(xs[i[0]], w, b, 1e-6, want_f32=False, want_bf16=True)
us = bench(f)
print(f"LN fwd [{M},{D}] fp32 -> bf16: {us:.1f} us  ({M*D*6/us/1e3:.0f} GB/s)")
