#!/usr/bin/env python
"""Micro-benchmark of the tcgen05 GEMM on the shapes of the distillation step (CUDA events, L2-cold by rotating buffers)."""
import math
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dinov2_distillation_b200 import ops

def bench(fn, iters=20):
    """Device time per call: the `iters` launches are replayed from ONE CUDA graph, so host launch cost is excluded."""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        fn()
        s.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(iters):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3  # us

def main():
    dev = "cuda"
    shapes = [
        ("square 8192", 8192, 8192, 8192, {}),
        ("square 4096", 4096, 4096, 4096, {}),
        ("vits qkv", 16448, 1152, 384, dict(bias=True, out="bf16")),
        ("vits proj +res", 16448, 384, 384, dict(bias=True, res=True)),
        ("vits fc1 gelu", 16448, 1536, 384, dict(bias=True, act="gelu", out="bf16")),
        ("vits fc1 noact", 16448, 1536, 384, dict(bias=True, out="bf16")),
        ("vits fc1 nobias", 16448, 1536, 384, dict(out="bf16")),
        ("vits fc2 +res", 16448, 384, 1536, dict(bias=True, res=True)),
        ("vitb qkv", 8224, 2304, 768, dict(bias=True, out="bf16")),
        ("vitb fc1 gelu", 8224, 3072, 768, dict(bias=True, act="gelu", out="bf16")),
        ("vitb fc2 +res", 8224, 768, 3072, dict(bias=True, res=True)),
        ("vitl fc1 gelu 518", 43840, 4096, 1024, dict(bias=True, act="gelu", out="bf16")),
        ("vitl fc2 +res 518", 43840, 1024, 4096, dict(bias=True, res=True)),
        ("wgrad ffn2 (MN)", 384, 1536, 16384, dict(wgrad=True)),
        ("wgrad conv (MN)", 384, 1024, 16384, dict(wgrad=True)),
        ("wgrad DxD (MN)", 384, 384, 16384, dict(wgrad=True)),
    ]
    filt = sys.argv[1] if len(sys.argv) > 1 else None
    for name, M, N, K, o in shapes:
        if filt and filt not in name:
            continue
        if o.get("wgrad"):
            a = torch.randn(K, M, device=dev).bfloat16()
            b = torch.randn(K, N, device=dev).bfloat16()
            out = torch.zeros(M, N, device=dev)
            tiles = math.ceil(M / 128) * math.ceil(N / 128)
            split = max(1, min(math.ceil(K / 64) // 4, math.ceil(148 / tiles)))
            fn = lambda: ops.gemm(a, b, a_mn_major=True, b_mn_major=True, out=out, atomic_add=True, split_k=split)
        else:
            a = torch.randn(M, K, device=dev).bfloat16()
            b = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
            bias = torch.randn(N, device=dev) if o.get("bias") else None
            res = torch.randn(M, N, device=dev) if o.get("res") else None
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if o.get("out") == "bf16" else torch.float32)
            if res is not None:
                out = res
            fn = lambda: ops.gemm(a, b, bias=bias, act=o.get("act", "none"), residual=res, out=out)
        us = bench(fn)
        tf = 2.0 * M * N * K / us / 1e6
        ref = float('nan') if filt else bench(lambda: torch.matmul(a.t() if o.get("wgrad") else a, b if o.get("wgrad") else b.t()))
        print(f"{name:22s} M={M:6d} N={N:5d} K={K:5d}  {us:8.1f} us  {tf:7.1f} TFLOP/s   (torch.matmul {ref:8.1f} us {2.0*M*N*K/ref/1e6:7.1f} TF/s)")

if __name__ == "__main__":
    main()
