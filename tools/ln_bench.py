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
# backward, stage-block form: dy + x + dres (fp32) -> dx (fp32) + dx16, no dgamma/dbeta; 6 rotating buffer sets (cold)
sets = [(torch.randn(M, D, device="cuda"), torch.randn(M, D, device="cuda"), torch.randn(M, D, device="cuda")) for _ in range(6)]
mean, rstd = torch.randn(M, device="cuda"), torch.rand(M, device="cuda") + 0.5
j = [0]
def g():
    j[0] = (j[0] + 1) % len(sets)
    dy, x, dres = sets[j[0]]
    ops.layernorm_bwd(dy, x, w, mean, rstd, dres=dres, want_wgrad=False, want_bf16=True)
us = bench(g)
print(f"LN bwd (stage form) [{M},{D}]: {us:.1f} us  ({M*D*(12+6)/us/1e3:.0f} GB/s)")
def g2():
    j[0] = (j[0] + 1) % len(sets)
    dy, x, dres = sets[j[0]]
    ops.layernorm_bwd(dy, x, w, mean, rstd, want_wgrad=True, want_bf16=True)
us = bench(g2)
print(f"LN bwd (projector form, dgamma/dbeta) [{M},{D}]: {us:.1f} us  ({M*D*(8+6)/us/1e3:.0f} GB/s)")
