"""LayerNorm-prologue GEMM (b200_ln_gemm_bf16) against layernorm_fwd + GEMM at the ViT-S teacher shapes."""
import math
import sys

import torch

sys.path.insert(0, ".")
from dinov2_distillation_b200 import ops  # noqa: E402


def bench(M, K, N, act, pre, iters=30):
    x = torch.randn(M, K, device="cuda")
    lw, lb = torch.ones(K, device="cuda"), torch.zeros(K, device="cuda")
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).bfloat16()
    bias = torch.zeros(N, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, on in (("fused", 1), ("ln + gemm", 0)):
        ops.set_option("gemm_ln", on)
        for _ in range(3):
            ops.ln_gemm(x, lw, lb, 1e-6, w, bias=bias, act=act, out_pre=pre)
        torch.cuda.synchronize()
        # ten calls in one CUDA graph: device time without the wrapper's host work; rotating inputs larger than L2 are
        # not needed for a relative comparison (both variants see the same warm L2)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(10):
                ops.ln_gemm(x, lw, lb, 1e-6, w, bias=bias, act=act, out_pre=pre)
        for cold in (False, True):
            ts = []
            for _ in range(iters):
                if cold:
                    flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                g.replay()
                b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b) * 1e2)
            ts.sort()
            print(f"M={M} K={K} N={N} act={act} pre={pre} {name:10s} {'cold-start' if cold else 'warm'}: median {ts[len(ts) // 2]:.1f} us/call"
                  f"  min {ts[0]:.1f}  ({2.0 * M * N * K / ts[len(ts) // 2] * 1e-6:.0f} TFLOP/s)")
    ops.set_option("gemm_ln", 1)


bench(16448, 384, 1152, "none", False)
bench(16448, 384, 1536, "gelu", False)
bench(16384, 384, 1536, "gelu", True)
