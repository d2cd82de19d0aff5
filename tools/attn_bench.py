#!/usr/bin/env python
"""Micro-benchmark of the attention kernels on the shapes of the distillation step (device time via graph replay)."""
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dinov2_distillation_b200 import ops
from tools.gemm_bench import bench

def main():
    shapes = [("teacher vits14 @224", 64, 6, 64, 257, torch.bfloat16, 1.0), ("stage blocks", 64, 6, 64, 256, torch.bfloat16, 1.0),
              ("teacher vitb14 @224 B=32", 32, 12, 64, 257, torch.bfloat16, 1.0),
              ("teacher vitl14 @518 B=32", 32, 16, 64, 1370, torch.bfloat16, 1.0),
              ("projector res4 hd24", 64, 16, 24, 256, torch.float16, 5.0), ("projector res5 hd16", 64, 24, 16, 256, torch.float16, 5.0)]
    for name, B, heads, hd, N, dt, ss in shapes:
        D = heads * hd
        qkv = (torch.randn(B, N, 3 * D, device="cuda") * 0.5).to(dt)
        q, k, v = qkv[..., :D], qkv[..., D:2 * D], qkv[..., 2 * D:]
        scale = ss / math.sqrt(hd)
        o, lse = ops.attention_fwd(q, k, v, heads, scale)
        us = bench(lambda: ops.attention_fwd(q, k, v, heads, scale))
        fl = 4.0 * B * heads * N * N * hd
        d_o = torch.randn(B, N, D, device="cuda").bfloat16()
        usb = bench(lambda: ops.attention_bwd(q, k, v, o, lse, d_o, heads, scale))
        print(f"{name:28s} B={B} h={heads} hd={hd} N={N}: fwd {us:7.1f} us {fl/us/1e6:7.1f} TF/s | bwd {usb:7.1f} us {2*fl/usb/1e6:7.1f} TF/s")

if __name__ == "__main__":
    main()
