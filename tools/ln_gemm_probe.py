import math, sys, torch
sys.path.insert(0, ".")
from dinov2_distillation_b200 import ops
for dbg in (64, 64 + 1):
  ops.set_option("gemm_dbg", dbg)
  print("gemm_dbg", dbg, "(2: no TMA stores, 8: no fence.proxy.async, 16: no staging writes, 1: drain only)", file=sys.stderr)
  for M, K, N, act in ((16448, 384, 1152, "none"), (16448, 384, 1536, "gelu")):
    x = torch.randn(M, K, device="cuda")
    lw, lb = torch.ones(K, device="cuda"), torch.zeros(K, device="cuda")
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    bias = torch.zeros(N, device="cuda")
    for _ in range(2):
        ops.ln_gemm(x, lw, lb, 1e-6, w, bias=bias, act=act)
torch.cuda.synchronize()
