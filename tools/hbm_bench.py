#!/usr/bin/env python
"""Bandwidth-bound kernels of the path, timed alone (CUDA-graph replay, CUDA events, rotating buffers larger than the 126 MB
L2) at the cfg2 size (vits14: 64 x 256 tokens x 384) and the cfg4 size (vitl14 @518: 32 x 1369 tokens x 1024):
achieved = ALGORITHMIC bytes / device time against MEASURED_PEAKS.json's copy bandwidth (burst figure: kernels timed alone).
usage: python tools/hbm_bench.py > profiles/rNN_hbm_kernels.md"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dinov2_distillation_b200 import ops
from tools.gemm_bench import bench

try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    SRC = "MEASURED_PEAKS.json hbm_gbs"
except Exception:
    PEAK, SRC = 6545.0, "fallback (MEASURED_PEAKS.json absent)"


def rot(make, n):
    items = [make() for _ in range(n)]
    i = [0]

    def nxt():
        i[0] = (i[0] + 1) % n
        return items[i[0]]
    return nxt


rows_out = []


def report(name, size, us, nbytes, formula):
    gbps = nbytes / us * 1e-3
    rows_out.append(f"| `{name}` | {size} | {formula} | {nbytes / 1e6:.1f} | {us:.1f} | {gbps:.0f} | {gbps / PEAK * 100:.0f} % |")


for tag, B, HW, D in (("cfg2", 64, 256, 384), ("cfg4", 32, 1369, 1024)):
    M = B * HW
    E = M * D
    n = max(2, int(400e6 // (E * 4)) + 1)   # rotating sets: > 400 MB in flight
    w, b = torch.rand(D, device="cuda") + 0.5, torch.randn(D, device="cuda")
    size = f"{tag}: {M} x {D}"
    # LayerNorm forward fp32 -> bf16
    nx = rot(lambda: torch.randn(M, D, device="cuda"), n)
    us = bench(lambda: ops.layernorm_fwd(nx(), w, b, 1e-6, want_f32=False, want_bf16=True))
    report("layernorm_fwd_kernel", size, us, E * 6, "M·D·(4 + 2)")
    # LayerNorm backward, re-used teacher block form: dy, x, dres (fp32) -> dx fp32 + dx bf16
    mean, rstd = torch.randn(M, device="cuda"), torch.rand(M, device="cuda") + 0.5
    ns = rot(lambda: tuple(torch.randn(M, D, device="cuda") for _ in range(3)), max(2, n // 3 + 1))

    def lnb():
        dy, x, dres = ns()
        ops.layernorm_bwd(dy, x, w, mean, rstd, dres=dres, want_wgrad=False, want_bf16=True)
    us = bench(lnb)
    report("layernorm_bwd_kernel<.,0> (teacher block: +residual grad)", size, us, E * 18, "M·D·(12 + 4 + 2)")

    def lnb2():
        dy, x, _ = ns()
        ops.layernorm_bwd(dy, x, w, mean, rstd, want_wgrad=True, want_bf16=True)
    us = bench(lnb2)
    report("layernorm_bwd_kernel<.,1> (projector: dγ/dβ sums)", size, us, E * 14, "M·D·(8 + 4 + 2)")
    # ScaleKD loss terms: S tokens + T tokens (teacher layout: cls row skipped)
    nst = rot(lambda: (torch.randn(B, HW, D, device="cuda"), torch.randn(B, HW + 1, D, device="cuda")), max(2, n // 2 + 1))
    g_out = torch.tensor([1.0, 0.0], device="cuda")
    S0, T0 = nst()
    _, ws = ops.kd_loss_fwd(S0, T0, 1, False, 0.08)

    def kf():
        S, T = nst()
        ops.kd_loss_fwd(S, T, 1, False, 0.08)
    us = bench(kf)
    report("kd_loss_fwd_kernel (+ finalize)", size, us, E * 8, "2·B·HW·D·4")

    def kb():
        S, T = nst()
        ops.kd_loss_bwd(S, T, 1, False, 0.08, g_out, ws)
    us = bench(kb)
    report("kd_loss_bwd_kernel", size, us, E * 12, "3·B·HW·D·4")
    us = bench(lambda: ops.cast_bf16(nx()))
    report("cast_f32_bf16_kernel", size, us, E * 6, "M·D·(4 + 2)")
    del nx, ns, nst
    torch.cuda.empty_cache()

for tag, B, R in (("cfg2 (64 x 224²)", 64, 224), ("cfg4 (32 x 518²)", 32, 518)):
    n = max(2, int(400e6 // (B * 3 * R * R * 4)) + 1)
    ni = rot(lambda: torch.randn(B, 3, R, R, device="cuda"), n)
    us = bench(lambda: ops.patch_im2col(ni()))
    P = B * (R // 14) ** 2
    report("patch_im2col_kernel", tag, us, B * 3 * R * R * 4 + P * 592 * 2, "B·3·H·W·4 + B·HW·592·2")
    del ni
    torch.cuda.empty_cache()

from dinov2_distillation_b200.scalekd import ScaleKD  # noqa: E402
for tag, B, C, g, D, heads in (("cfg2 res5 (64 x 1024 x 16²)", 64, 1024, 16, 384, 24), ("cfg4 res5 (32 x 768 x 37²)", 32, 768, 37, 1024, 16)):
    m = ScaleKD(name="scalekd_res5", alpha=[0.08, 0.06], student_dims=C, teacher_dims=D, query_hw=[g, g], pos_hw=[g, g],
                pos_dims=D, window_shapes=[1, 1], self_query=True, softmax_scale=[5.0, 5.0], num_heads=heads).cuda().train()
    n = max(2, int(400e6 // (B * C * g * g * 4)) + 1)
    ni = rot(lambda: torch.randn(B, C, g, g, device="cuda"), n)
    us = bench(lambda: m.projector_0.tokenize(ni()))
    report("tokenize_split3_kernel (student map -> bf16 tokens + 3-term fp16 split)", tag, us, B * C * g * g * 12, "B·C·HW·(4 + 2 + 6)")
    del ni, m
    torch.cuda.empty_cache()

print(f"# Bandwidth-bound kernels, timed alone (tools/hbm_bench.py)\n")
print(f"Peak: {PEAK:.0f} GB/s ({SRC}; the burst copy figure, these kernels are timed alone). Device time per launch from a "
      f"CUDA-graph replay of 20 launches over rotating buffers (> 400 MB, L2-cold), CUDA events.\n")
print("| kernel | size | algorithmic bytes | MB | us | GB/s | of peak |")
print("|---|---|---|---:|---:|---:|---:|")
print("\n".join(rows_out))
