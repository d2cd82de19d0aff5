"""Clock stamps of the two-tile attention forward (needs a -DB200_ATTN_PROBES build:
B200_EXTRA_NVCC_FLAGS=-DB200_ATTN_PROBES python -m dinov2_distillation_b200.build --force)."""
import math, sys, torch
sys.path.insert(0, ".")
from dinov2_distillation_b200 import ops
for B, heads, hd, N, dt in ((64, 24, 16, 256, torch.float16), (64, 6, 64, 256, torch.bfloat16)):
    D = heads * hd
    q = (torch.randn(B, N, D, device="cuda") * 0.5).to(dt)
    kv = (torch.randn(B, N, 2 * D, device="cuda") * 0.5).to(dt)
    for _ in range(8):
        ops.attention_fwd(q, kv[..., :D], kv[..., D:], heads, 5.0 / math.sqrt(hd))
    torch.cuda.synchronize()
