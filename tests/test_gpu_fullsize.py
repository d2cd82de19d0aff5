"""Full-size parity on the GPU. The CPU oracle is too slow for BASELINE.json's real sizes, so here the SAME oracle code
(oracle/*.py: fp32 restatement of the reference, pinned on the CPU against the unmodified reference and its golden
vectors) runs ON the GPU in fp32 with TF32 off -- stock ATen kernels as the checker -- against the CUDA path through the
C ABI, at: cfg2 as benchmarked (vits14 -> stdc_2 shapes, config.yaml losses, B = 64 @224), cfg3 / cfg4 shapes
(D = 768 / 1024, 37 x 37 tokens, head_dim 64), and the true widths of every teacher (vitl14 @518, vitg14 @224).

Gates are BASELINE.json's north_star values, unscaled: teacher cosine >= 0.9999 per image, each loss rel err <= 1e-3,
similarities abs 1e-3, dS and the flat projector gradient rel err <= 1e-2; per-tensor parameter gradients through
test_gpu_modules.check_param_grads (documented there). Every test prints its worst quantity.
"""
import os
import warnings

import pytest
import torch

from test_gpu_modules import GRAD_RTOL, LOSS_RTOL, SIM_ATOL, _teacher_pair, check_param_grads, rel

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_oracle_arithmetic():
    """The oracle's matmuls must be true fp32 on the GPU (the reference itself turns TF32 on, train.py:304: not inherited)."""
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _specs(D, g, defs):
    common = dict(alpha=[0.08, 0.06], teacher_dims=D, query_hw=[g, g], pos_hw=[g, g], pos_dims=D, window_shapes=[1, 1],
                  softmax_scale=[5.0, 5.0])
    return [{"type": "scalekd", "weight": 1.0, "kwargs": dict(common, name=n, student_dims=cs, self_query=sq, num_heads=h)}
            for n, cs, h, sq in defs], common


def _run_pipeline_vs_gpu_oracle(teacher_name, size, B, defs, seed=3):
    """Teacher forward + _compute_losses + backward: CUDA path vs the oracle port on the same device. Returns the worst
    loss / similarity / gradient errors after asserting the north-star gates."""
    warnings.simplefilter("ignore")
    from dinov2_distillation_b200 import distill
    from oracle import dinov2_ref, scalekd_ref
    t, cfg, tsd = _teacher_pair(teacher_name, seed=1)
    g = size // 14
    specs, common = _specs(cfg.dim, g, defs)
    torch.manual_seed(seed)
    step = distill.DistillationStep(None, t, specs)
    sds = {n: {k: v.detach().clone().cuda() for k, v in m.state_dict().items()} for n, m in step.losses.items()}
    step = step.cuda().train()
    gen = torch.Generator().manual_seed(2)
    img = torch.randn(B, 3, size, size, generator=gen).cuda()
    feats = {n.split("_")[1]: torch.randn(B, cs, g, g, generator=gen).cuda() for n, cs, _, _ in defs}
    # ---- oracle on the GPU (fp32, TF32 off)
    tsd_d = {k: v.cuda() for k, v in tsd.items()}
    with torch.no_grad():
        T_ref = dinov2_ref.teacher_feature_map(tsd_d, cfg, img)
    for sd in sds.values():
        for v in sd.values():
            if v.is_floating_point():
                v.requires_grad_(True)
    losses = {s["kwargs"]["name"]: dict(sd=sds[s["kwargs"]["name"]], weight=s["weight"], alpha=common["alpha"], hw=(g, g),
                                        num_heads=s["kwargs"]["num_heads"], softmax_scale=common["softmax_scale"])
              for s in specs}
    blocks = [lambda x, i=i: dinov2_ref.block(tsd_d, i, x, cfg) for i in range(cfg.depth)]
    fr = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
    ref = scalekd_ref.compute_losses(losses, fr, T_ref, blocks)
    ref["loss"].backward()
    # ---- CUDA path
    fc = {k: v.clone().requires_grad_(True) for k, v in feats.items()}
    T = t(img)["feature_map"]
    cos = torch.nn.functional.cosine_similarity(T.float().flatten(1), T_ref.flatten(1), dim=1)
    out = step._compute_losses({"student": fc, "teacher": T})
    out["loss"].backward()
    torch.cuda.synchronize()
    # ---- gates
    worst = {"teacher_cos_min": cos.min().item(), "teacher_max_abs": (T.float() - T_ref).abs().max().item()}
    assert cos.min().item() >= 0.9999, cos.min().item()
    lerr = {k: abs(out[k].item() - v.item()) / abs(v.item()) for k, v in ref.items() if not k.endswith("similarity")}
    serr = {k: abs(out[k].item() - v.item()) for k, v in ref.items() if k.endswith("similarity")}
    worst["loss_rel"] = max(lerr.items(), key=lambda kv: kv[1])
    worst["sim_abs"] = max(serr.items(), key=lambda kv: kv[1])
    derr = {k: rel(fc[k].grad, fr[k].grad) for k in fc}
    worst["dS_rel"] = max(derr.items(), key=lambda kv: kv[1])
    print(f"{teacher_name} @{size} B={B}: {worst}")
    assert worst["loss_rel"][1] <= LOSS_RTOL, lerr
    assert worst["sim_abs"][1] <= SIM_ATOL, serr
    assert worst["dS_rel"][1] <= GRAD_RTOL, derr
    flat = check_param_grads({f"{n}.{k}": (p.grad, sds[n][k].grad) for n, m in step.losses.items()
                              for k, p in m.named_parameters() if sds[n][k].grad is not None})
    worst["param_flat_rel"] = flat
    return worst


def test_cfg2_full_batch_vs_gpu_oracle():
    """BASELINE.json configs[1] exactly as benchmarked: vits14, config.yaml res4 (heads 16, self query) + res5 (heads 24,
    query from res4), stdc_2 channel counts, B = 64 @224."""
    _run_pipeline_vs_gpu_oracle("dinov2_vits14", 224, 64,
                                [("scalekd_res4", 512, 16, True), ("scalekd_res5", 1024, 24, False)])


def test_cfg3_shapes_vs_gpu_oracle():
    """configs[2] per-GPU shapes: vitb14 -> convnext_tiny channels (384 / 768), head dims 48 / 32, B = 32 @224."""
    _run_pipeline_vs_gpu_oracle("dinov2_vitb14", 224, 32,
                                [("scalekd_res4", 384, 16, True), ("scalekd_res5", 768, 24, False)])


def test_cfg4_shapes_vs_gpu_oracle():
    """configs[3] shapes: vitl14 (D = 1024, 24 layers, re-used blocks 18-22) @518 = 37 x 37 tokens (odd grid, 1369-token
    attention), swin_tiny channels, 16 heads (head_dim 64; the shipped 24 does not divide 1024), at B = 2."""
    _run_pipeline_vs_gpu_oracle("dinov2_vitl14", 518, 2,
                                [("scalekd_res4", 384, 16, True), ("scalekd_res5", 768, 16, False)])


@pytest.mark.parametrize("name,size,B,depth", [("dinov2_vitl14", 518, 2, 4), ("dinov2_vitg14", 224, 4, 3),
                                               ("dinov2_vitg14", 518, 1, 2)])
def test_teacher_true_width_vs_gpu_oracle(name, size, B, depth):
    """The widths cfg4 / cfg5 run (train.py:103-108: vitl14 D = 1024 / 16 heads, vitg14 D = 1536 / 24 heads / fused
    SwiGLU 4096) at a reduced depth: exercises the cta_group::2 GEMM path (K >= 1024) and the SwiGLU kernels end to end."""
    from oracle import dinov2_ref
    full = dinov2_ref.TEACHER_CFGS[name]
    cfg = dinov2_ref.VitCfg(full.dim, depth, full.heads, full.ffn_hidden, full.swiglu)
    t, cfg, sd = _teacher_pair(cfg, seed=11)
    x = torch.randn(B, 3, size, size, generator=torch.Generator().manual_seed(0)).cuda()
    sd_d = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        ref = dinov2_ref.teacher_feature_map(sd_d, cfg, x)
    got = t(x)["feature_map"]
    assert got.shape == ref.shape and got.stride() == ref.stride()
    cos = torch.nn.functional.cosine_similarity(got.float().flatten(1), ref.flatten(1), dim=1)
    tok = torch.nn.functional.cosine_similarity(got.float(), ref, dim=1)
    print(f"{name} L={depth} @{size}: cos min {cos.min().item():.6f}, token cos min {tok.min().item():.5f}, "
          f"max abs {(got.float() - ref).abs().max().item():.4f}")
    assert cos.min().item() >= 0.9999, cos
    assert tok.min().item() >= 0.999, tok.min()


def test_teacher_vs_huggingface_dinov2_at_518():
    """Second, independent pin of the CUDA teacher: transformers' Dinov2Model (same published architecture, separate code
    base) with the same weights, at 518 pixels where no position-embedding interpolation is involved (HF interpolates
    with size=, the hub with scale_factor=: they differ at other resolutions)."""
    pytest.importorskip("transformers")
    from test_oracle_teacher import _hf_model
    t, cfg, sd = _teacher_pair("dinov2_vits14", seed=5)
    hf = _hf_model(cfg, sd, 518).cuda()
    x = torch.randn(1, 3, 518, 518, generator=torch.Generator().manual_seed(0)).cuda()
    with torch.no_grad():
        ref = hf(pixel_values=x).last_hidden_state[:, 1:]          # normed patch tokens [B, HW, D]
    got = t(x)["feature_map"].float().flatten(2).transpose(1, 2)   # [B, D, H, W] -> [B, HW, D]
    cos = torch.nn.functional.cosine_similarity(got.flatten(1), ref.flatten(1), dim=1)
    print(f"CUDA teacher vs HF Dinov2Model @518: cos {cos.min().item():.6f}, max abs {(got - ref).abs().max().item():.4f}")
    assert cos.min().item() >= 0.9999, cos


def test_teacher_real_weights_if_available():
    """Optional: with real hub checkpoints on disk (DINOV2_WEIGHTS_DIR/<name>_pretrain.pth, the files
    torch.hub.load('facebookresearch/dinov2', ...) downloads -- models/backbones/dinov2.py:20), the CUDA teacher must
    match the fp32 oracle with those weights. Skips offline."""
    d = os.environ.get("DINOV2_WEIGHTS_DIR")
    name = "dinov2_vits14"
    path = os.path.join(d, f"{name}_pretrain.pth") if d else None
    if not path or not os.path.isfile(path):
        pytest.skip("no real DINOv2 checkpoint on disk (set DINOV2_WEIGHTS_DIR)")
    from dinov2_distillation_b200 import teacher
    from oracle import dinov2_ref
    t = teacher.DINOv2ViT(name).cuda().eval()
    sd = {k: v.cuda() for k, v in torch.load(path, map_location="cpu").items()}
    cfg = dinov2_ref.TEACHER_CFGS[name]
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(0)).cuda()
    with torch.no_grad():
        ref = dinov2_ref.teacher_feature_map(sd, cfg, x)
    got = t(x)["feature_map"].float()
    cos = torch.nn.functional.cosine_similarity(got.flatten(1), ref.flatten(1), dim=1)
    assert cos.min().item() >= 0.9999, cos


def test_reference_loop_over_b200_shells_matches_distillation_step():
    """The reference's own orchestration (restated object-for-object in oracle.scalekd_ref.compute_losses_modules and
    pinned against DistillationModule._compute_losses on the CPU) driven over the B200 drop-ins -- ScaleKD shells and
    teacher.model.blocks -- must give what DistillationStep._compute_losses gives: the drop-ins work behind the
    reference's loop, not only behind ours (train/distillation_module.py:180-246)."""
    warnings.simplefilter("ignore")
    from dinov2_distillation_b200 import distill
    from oracle import scalekd_ref
    t, cfg, _ = _teacher_pair("dinov2_vits14", seed=1)
    specs, _ = _specs(384, 16, [("scalekd_res4", 512, 16, True), ("scalekd_res5", 1024, 24, False)])
    specs[0]["weight"], specs[1]["weight"] = 2.0, 0.5
    torch.manual_seed(3)
    step = distill.DistillationStep(None, t, specs).cuda().train()
    gen = torch.Generator().manual_seed(4)
    B = 8
    img = torch.randn(B, 3, 224, 224, generator=gen).cuda()
    base = {"res4": torch.randn(B, 512, 16, 16, generator=gen).cuda(), "res5": torch.randn(B, 1024, 16, 16, generator=gen).cuda()}
    T = t(img)["feature_map"]

    def run(fn):
        for p in step.losses.parameters():
            p.grad = None
        f = {k: v.clone().requires_grad_(True) for k, v in base.items()}
        # (BatchNorm running statistics advance on every training forward: irrelevant to the batch-statistics outputs)
        out = fn(f)
        out["loss"].backward()
        torch.cuda.synchronize()
        return out, {k: v.grad.clone() for k, v in f.items()}, {k: p.grad.clone() for k, p in step.losses.named_parameters()}

    weights = {s["kwargs"]["name"]: s["weight"] for s in specs}
    ours = run(lambda f: step._compute_losses({"student": f, "teacher": T}))
    theirs = run(lambda f: scalekd_ref.compute_losses_modules(step.losses, weights, f, T, t))
    assert list(ours[0].keys()) == list(theirs[0].keys())
    for k in ours[0]:
        a, b = ours[0][k].item(), theirs[0][k].item()
        # (the two loops launch the two branches in a different stream order: atomics in the BatchNorm statistics and
        # the loss reductions re-associate (measured: 1e-4 relative), nothing else differs)
        assert abs(a - b) <= max(1e-4 * abs(b), 2e-5), (k, a, b)
    # gradients: a handful of ReLU masks flip with the re-associated BatchNorm statistics (measured 3-4e-3, as in
    # test_cfg2_full_size_properties): the north-star gradient gate applies
    for k in base:
        assert rel(ours[1][k], theirs[1][k]) <= GRAD_RTOL, (k, rel(ours[1][k], theirs[1][k]))
    check_param_grads({k: (ours[2][k], v) for k, v in theirs[2].items()})
