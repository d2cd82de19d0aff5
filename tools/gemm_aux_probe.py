"""Wait-cycle counters of the plain and the aux-epilogue GEMM (needs a -DB200_GEMM_PROBES build)."""
import math, sys, torch
sys.path.insert(0, ".")
from dinov2_distillation_b200 import ops
M, K, N = 16384, 384, 1536
a = torch.randn(M, K, device="cuda").bfloat16()
w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
aux = torch.randn(M, N, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
ops.set_option("gemm_ew", 8)
for _ in range(3):
    ops.gemm(a, w, out=out); ops.gemm(a, w, out=out, aux=aux, aux_mode="drelu")
torch.cuda.synchronize()
ops.set_option("gemm_dbg", 32)
for name, kw in (("plain", {}), ("x relu mask", dict(aux=aux, aux_mode="drelu")), ("x gelu'", dict(aux=aux, aux_mode="dgelu"))):
    print(name, file=sys.stderr, flush=True)
    for _ in range(2):
        ops.gemm(a, w, out=out, **kw)
    torch.cuda.synchronize()
