"""Clock-stamp probe of attn_bwd_tc_kernel (library built with B200_EXTRA_NVCC_FLAGS=-DB200_ATTN_PROBES).
usage: python tools/attn_bwd_probe.py [hd heads half]"""
import math
import sys

import torch

sys.path.insert(0, ".")
from dinov2_distillation_b200 import ops  # noqa: E402

hd = int(sys.argv[1]) if len(sys.argv) > 1 else 64
heads = int(sys.argv[2]) if len(sys.argv) > 2 else 6
half = len(sys.argv) > 3 and sys.argv[3] == "1"
B, N = 64, 256
D = heads * hd
dt = torch.float16 if half else torch.bfloat16
scale = (5.0 if hd != 64 else 1.0) / math.sqrt(hd)
q = (torch.randn(B, N, D, device="cuda") * 0.5).to(dt)
k = (torch.randn(B, N, D, device="cuda") * 0.5).to(dt)
v = torch.randn(B, N, D, device="cuda").to(dt)
d_o = torch.randn(B, N, D, device="cuda").bfloat16()
o, lse = ops.attention_fwd(q, k, v, heads, scale)
for _ in range(5):
    ops.attention_bwd(q, k, v, o, lse, d_o, heads, scale)
torch.cuda.synchronize()
