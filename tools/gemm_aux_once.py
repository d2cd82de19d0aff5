"""One launch each of the plain and the x-relu-mask GEMM at M = 16 384, K = 384, N = 1 536 (for an ncu capture)."""
import math, sys, torch
sys.path.insert(0, ".")
from dinov2_distillation_b200 import ops
M, K, N = 16384, 384, 1536
a = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
aux = torch.randn(M, N, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
ops.set_option("gemm_ew", 8)
for _ in range(2):
    ops.gemm(a, w, out=out)
    ops.gemm(a, w, out=out, aux=aux, aux_mode="drelu")
torch.cuda.synchronize()
print("done")
