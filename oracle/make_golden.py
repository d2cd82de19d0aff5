"""ORACLE tool: generate tests/golden/*.pt by EXECUTING THE UNMODIFIED REFERENCE (this container only).

    python -m oracle.make_golden

The reference has no golden vectors of its own, so these files are the pin: they hold inputs,
weights and the outputs/gradients that /root/reference's own `ScaleKD` and `DistillationModule._compute_losses`
produce on CPU fp32. The teacher used for the pipeline fixture is oracle/dinov2_ref.py (the hub model is not on disk).
"""
from __future__ import annotations

import os
import sys

import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import dinov2_ref, ref_shims  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def _grads(module: nn.Module):
    return {k: p.grad.detach().clone() for k, p in module.named_parameters() if p.grad is not None}


def scalekd_tiny(sk):
    torch.manual_seed(11)
    kw = dict(name="scalekd_res5", alpha=[0.08, 0.06], student_dims=32, teacher_dims=64, query_hw=[4, 4], pos_hw=[4, 4],
              pos_dims=64, window_shapes=[1, 1], self_query=True, softmax_scale=[5.0, 5.0], num_heads=4)
    m = sk.ScaleKD(**kw)
    # make every parameter non-trivial (BN/LN affine default to 1/0)
    g = torch.Generator().manual_seed(12)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    B = 2
    S = torch.randn(B, 32, 4, 4, generator=g, requires_grad=True)
    T = torch.randn(B, 64, 4, 4, generator=g) + 0.2
    sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
    m.train()
    out = m(S, T)
    out["loss"].backward()
    torch.save({"kwargs": kw, "state_dict": sd0, "preds_S": S.detach(), "preds_T": T,
                "out": {k: v.detach() for k, v in out.items()}, "grad_S": S.grad.detach(), "grads": _grads(m),
                "state_dict_after": {k: v.detach().clone() for k, v in m.state_dict().items()}},
               os.path.join(GOLDEN, "scalekd_tiny.pt"))
    print("scalekd_tiny:", {k: float(v) for k, v in out.items()})


def scalekd_windows(sk):
    """window_shapes = [2, 2] (losses/scalekd.py:305-308, :326-335) on an 8x8 grid: windowed cross attention whose
    output stays window-major. One self-query ScaleKD and one with external queries."""
    for tag, self_query in (("self", True), ("ext", False)):
        torch.manual_seed(31)
        kw = dict(name="scalekd_res5", alpha=[0.08, 0.06], student_dims=24, teacher_dims=64, query_hw=[8, 8],
                  pos_hw=[8, 8], pos_dims=64, window_shapes=[2, 2], self_query=self_query, softmax_scale=[5.0, 3.0],
                  num_heads=4)
        m = sk.ScaleKD(**kw)
        g = torch.Generator().manual_seed(32)
        with torch.no_grad():
            for n, p in m.named_parameters():
                if p.dim() == 1:
                    p.add_(0.1 * torch.randn(p.shape, generator=g))
        B = 3
        S = torch.randn(B, 24, 8, 8, generator=g, requires_grad=True)
        T = torch.randn(B, 64, 8, 8, generator=g) + 0.2
        qs = None if self_query else torch.randn(B, 64, 64, generator=g)
        qf = None if self_query else torch.randn(B, 64, 64, generator=g)
        sd0 = {k: v.detach().clone() for k, v in m.state_dict().items()}
        m.train()
        out = m(S, T, query_s=qs, query_f=qf)
        out["loss"].backward()
        torch.save({"kwargs": kw, "state_dict": sd0, "preds_S": S.detach(), "preds_T": T, "query_s": qs, "query_f": qf,
                    "out": {k: v.detach() for k, v in out.items()}, "grad_S": S.grad.detach(), "grads": _grads(m)},
                   os.path.join(GOLDEN, f"scalekd_win_{tag}.pt"))
        print(f"scalekd_win_{tag}:", {k: float(v) for k, v in out.items()})


def scalekd_cfg1(sk):
    """BASELINE.json configs[0] loss shapes: vits14 (D=384) + resnet_18 res5 (Cs=512), 16x16 grid, B=2, heads 24.
    Weights are NOT stored (35 MB): both sides rebuild them from torch.manual_seed(3) -- the constructor consumes the RNG
    identically (checked in tests). Stored: scalars, gradient norms and a strided sample of every gradient."""
    kw = dict(name="scalekd_res5", alpha=[0.08, 0.06], student_dims=512, teacher_dims=384, query_hw=[16, 16],
              pos_hw=[16, 16], pos_dims=384, window_shapes=[1, 1], self_query=True, softmax_scale=[5.0, 5.0],
              num_heads=24)
    torch.manual_seed(3)
    m = sk.ScaleKD(**kw)
    g = torch.Generator().manual_seed(4)
    S = torch.randn(2, 512, 16, 16, generator=g, requires_grad=True)
    T = torch.randn(2, 384, 16, 16, generator=g)
    m.train()
    out = m(S, T)
    out["loss"].backward()
    grads = _grads(m)
    torch.save({"kwargs": kw, "seed_model": 3, "seed_data": 4,
                "out": {k: v.detach() for k, v in out.items()},
                "grad_S_norm": S.grad.norm(), "grad_S_sample": S.grad.flatten()[::997].clone(),
                "grad_norms": {k: v.norm() for k, v in grads.items()},
                "grad_samples": {k: v.flatten()[::max(1, v.numel() // 64)][:64].clone() for k, v in grads.items()}},
               os.path.join(GOLDEN, "scalekd_cfg1.pt"))
    print("scalekd_cfg1:", {k: float(v) for k, v in out.items()})


def pipeline_tiny(sk, dm):
    """res4 -> teacher blocks -> res5 chaining through the reference's own DistillationModule._compute_losses."""
    cfg = dinov2_ref.VitCfg(64, 8, 4, 256)
    tsd = dinov2_ref.make_state_dict(cfg, seed=21, pos_grid=4)
    teacher = dinov2_ref.RefTeacher(cfg, tsd)
    torch.manual_seed(22)
    common = dict(alpha=[0.08, 0.06], teacher_dims=64, query_hw=[4, 4], pos_hw=[4, 4], pos_dims=64,
                  window_shapes=[1, 1], softmax_scale=[5.0, 5.0])
    specs = [
        {"type": "scalekd", "weight": 1, "kwargs": dict(common, name="scalekd_res4", student_dims=32, self_query=True, num_heads=4)},
        {"type": "scalekd", "weight": 1.0, "kwargs": dict(common, name="scalekd_res5", student_dims=48, self_query=False, num_heads=8)},
    ]
    mod = dm.DistillationModule.__new__(dm.DistillationModule)
    nn.Module.__init__(mod)
    mod.teacher = teacher
    mod.losses = nn.ModuleDict({s["kwargs"]["name"]: sk.ScaleKD(**s["kwargs"]) for s in specs})
    mod.loss_weights = {s["kwargs"]["name"]: s["weight"] for s in specs}
    g = torch.Generator().manual_seed(23)
    with torch.no_grad():
        for n, p in mod.losses.named_parameters():
            if p.dim() == 1:
                p.add_(0.1 * torch.randn(p.shape, generator=g))
    B = 2
    img = torch.randn(B, 3, 56, 56, generator=g)
    T = teacher(img)["feature_map"]
    feats = {"res4": torch.randn(B, 32, 4, 4, generator=g, requires_grad=True),
             "res5": torch.randn(B, 48, 4, 4, generator=g, requires_grad=True)}
    sd0 = {k: v.detach().clone() for k, v in mod.losses.state_dict().items()}
    mod.losses.train()
    out = mod._compute_losses({"student": feats, "teacher": T})
    out["loss"].backward()
    torch.save({"teacher_cfg": (64, 8, 4, 256), "teacher_sd": tsd, "specs": specs, "losses_sd": sd0, "img": img,
                "teacher_map": T.contiguous(), "feats": {k: v.detach() for k, v in feats.items()},
                "out": {k: (v.detach() if torch.is_tensor(v) else torch.tensor(v)) for k, v in out.items()},
                "grad_feats": {k: v.grad.detach() for k, v in feats.items()}, "grads": _grads(mod.losses)},
               os.path.join(GOLDEN, "pipeline_tiny.pt"))
    print("pipeline_tiny:", {k: float(v) for k, v in out.items()})


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    sk, dm = ref_shims.import_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    scalekd_tiny(sk)
    scalekd_windows(sk)
    scalekd_cfg1(sk)
    if dm is not None:
        pipeline_tiny(sk, dm)
    else:
        print("train.distillation_module not importable; pipeline fixture skipped")


if __name__ == "__main__":
    main()
