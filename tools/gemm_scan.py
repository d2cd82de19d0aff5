import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dinov2_distillation_b200 import ops
def bench(fn, iters=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for (N, K) in [(384, 384), (1152, 384), (1536, 384), (384, 1536)]:
    for M in [128, 1024, 4096, 8192, 16448, 32896]:
        a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        us = bench(lambda: ops.gemm(a, b, out=out))
        ref = bench(lambda: torch.matmul(a, b.t(), out=out))
        print(f"N={N:5d} K={K:5d} M={M:6d}: {us:7.1f} us  {2.0*M*N*K/us/1e6:7.1f} TF/s | cublas {ref:7.1f} us")
