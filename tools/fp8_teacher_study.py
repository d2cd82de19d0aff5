#!/usr/bin/env python
"""SURVEY 8 f3, second half: would FP8 (e4m3) operands for the frozen teacher's linear layers keep the north-star gate
(teacher features: cosine >= 0.9999 per image against the fp32 reference)?  CPU emulation on the oracle: every
nn.Linear of the teacher (qkv / proj / fc1 / fc2 or w12 / w3) gets its weight rounded to e4m3 with a per-output-channel
scale and its input rounded to e4m3 with a per-token scale (the most favourable per-tensor-row scheme tcgen05
kind::f8f6f4 can consume without block scaling); accumulation, LayerNorm, softmax, residual stream stay fp32. The bf16
row shows the operand rounding the shipped kernels use. Synthetic seeded weights (no checkpoints offline), so the
absolute numbers are indicative; the ORDER OF MAGNITUDE of the fp8 error is what decides.
usage: python tools/fp8_teacher_study.py [--models dinov2_vits14,dinov2_vitg14] [--res 224]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from oracle import dinov2_ref as R

E4M3_MAX = 448.0


def q_e4m3(t, dim):
    s = t.abs().amax(dim=dim, keepdim=True).clamp_min(1e-12) / E4M3_MAX
    return (t / s).to(torch.float8_e4m3fn).to(torch.float32) * s


def make_linear(mode):
    real = F.linear

    def lin(x, w, b=None):
        if mode == "fp32":
            return real(x, w, b)
        if mode == "bf16":
            return real(x.bfloat16().float(), w.bfloat16().float(), b)
        if mode == "fp8_w":      # weights only (activations bf16): what a weight-only scheme would give
            return real(x.bfloat16().float(), q_e4m3(w, 1), b)
        return real(q_e4m3(x, -1), q_e4m3(w, 1), b)   # fp8: both operands
    return lin


def run(name, res, B):
    cfg = R.TEACHER_CFGS[name]
    sd = R.make_state_dict(cfg, seed=1)
    x = torch.randn(B, 3, res, res, generator=torch.Generator().manual_seed(0))
    outs = {}
    real = F.linear
    for mode in ("fp32", "bf16", "fp8_w", "fp8"):
        R.F.linear = make_linear(mode)
        try:
            with torch.no_grad():
                outs[mode] = R.get_intermediate_layers(sd, cfg, x)[0]
        finally:
            R.F.linear = real
    ref = outs["fp32"]
    for mode in ("bf16", "fp8_w", "fp8"):
        cos = F.cosine_similarity(outs[mode].flatten(1), ref.flatten(1), dim=1)
        tok = F.cosine_similarity(outs[mode], ref, dim=-1)
        print(f"{name} @{res} B={B} {mode:6s}: per-image cosine min {cos.min().item():.6f} mean {cos.mean().item():.6f} | "
              f"per-token cosine min {tok.min().item():.6f} | max-abs {((outs[mode] - ref).abs().max().item()):.4f} "
              f"| gate 0.9999: {'PASS' if cos.min().item() >= 0.9999 else 'FAIL'}")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--models", default="dinov2_vits14,dinov2_vitb14")
    ap.add_argument("--res", type=int, default=224)
    ap.add_argument("--batch", type=int, default=2)
    a = ap.parse_args()
    torch.manual_seed(0)
    for m in a.models.split(","):
        run(m, a.res, a.batch)
