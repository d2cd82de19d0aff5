"""CPU: pin the oracle (oracle/scalekd_ref.py) against the reference -- the golden vectors that oracle/make_golden.py
produced by executing /root/reference, and the live reference when it is present in this container."""
import os

import pytest
import torch

from oracle import dinov2_ref, ref_shims, scalekd_ref

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def _grad_sd(sd):
    return {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v.clone()) for k, v in sd.items()}


def test_port_matches_golden_scalekd_tiny():
    g = torch.load(os.path.join(GOLDEN, "scalekd_tiny.pt"))
    kw = g["kwargs"]
    sd = _grad_sd(g["state_dict"])
    S = g["preds_S"].clone().requires_grad_(True)
    out = scalekd_ref.scalekd_forward(sd, S, g["preds_T"], alpha=kw["alpha"], hw=kw["query_hw"],
                                      num_heads=kw["num_heads"], softmax_scale=kw["softmax_scale"])
    for k, v in g["out"].items():
        assert abs(out[k].item() - v.item()) <= 1e-5 * max(1.0, abs(v.item())), k
    out["loss"].backward()
    assert rel(S.grad, g["grad_S"]) < 1e-4
    for k, ref in g["grads"].items():
        if ref.norm() < 1e-6:
            continue
        assert rel(sd[k].grad, ref) < 2e-4, (k, rel(sd[k].grad, ref))


def test_port_matches_golden_pipeline():
    """res4 -> teacher blocks -> res5 chaining as the reference's DistillationModule._compute_losses does it."""
    g = torch.load(os.path.join(GOLDEN, "pipeline_tiny.pt"))
    cfg = dinov2_ref.VitCfg(*g["teacher_cfg"])
    tsd = g["teacher_sd"]
    feats = {k: v.clone().requires_grad_(True) for k, v in g["feats"].items()}
    losses = {}
    sds = {}
    for spec in g["specs"]:
        kw = spec["kwargs"]
        name = kw["name"]
        sds[name] = _grad_sd({k[len(name) + 1:]: v for k, v in g["losses_sd"].items() if k.startswith(name + ".")})
        losses[name] = dict(sd=sds[name], weight=spec["weight"], alpha=kw["alpha"], hw=kw["query_hw"],
                            num_heads=kw["num_heads"], softmax_scale=kw["softmax_scale"])
    blocks = [lambda x, i=i: dinov2_ref.block(tsd, i, x, cfg) for i in range(cfg.depth)]
    out = scalekd_ref.compute_losses(losses, feats, g["teacher_map"], blocks)
    assert sorted(out.keys()) == sorted(g["out"].keys())
    for k, v in g["out"].items():
        assert abs(out[k].item() - v.item()) <= 2e-5 * max(1.0, abs(v.item())), (k, out[k].item(), v.item())
    out["loss"].backward()
    for k in feats:
        assert rel(feats[k].grad, g["grad_feats"][k]) < 2e-4
    for k, ref in g["grads"].items():
        name, pk = k.split(".", 1)
        if ref.norm() < 1e-6:
            continue
        assert rel(sds[name][pk].grad, ref) < 5e-4, (k, rel(sds[name][pk].grad, ref))


def test_teacher_oracle_reproduces_fixture_map():
    g = torch.load(os.path.join(GOLDEN, "pipeline_tiny.pt"))
    cfg = dinov2_ref.VitCfg(*g["teacher_cfg"])
    with torch.no_grad():
        T = dinov2_ref.teacher_feature_map(g["teacher_sd"], cfg, g["img"])
    assert rel(T, g["teacher_map"]) < 1e-5


def test_stage_block_ranges():
    """_forward_specific_stage (train/distillation_module.py:162-176): only res4 is non-empty."""
    assert list(scalekd_ref.stage_block_range(12, "res4")) == [9, 10]
    assert list(scalekd_ref.stage_block_range(24, "res4")) == [18, 19, 20, 21, 22]
    assert list(scalekd_ref.stage_block_range(40, "res4")) == list(range(30, 39))
    for L in (12, 24, 40):
        assert list(scalekd_ref.stage_block_range(L, "res2")) == []
        assert list(scalekd_ref.stage_block_range(L, "res3")) == []


@pytest.mark.parametrize("R", [4, 16, 37])
def test_dct_zero_dc_equals_mean_subtraction(R):
    """The identity the CUDA frequency term relies on: idct2(zero_dc(dct2(x))) == x - mean_{H,W}(x)."""
    x = torch.randn(2, 5, R, R, dtype=torch.float64)
    wf, wi = scalekd_ref.dct_matrices(R, torch.float64)
    assert (scalekd_ref.idct2(scalekd_ref.dct2(x, wf), wi) - x).abs().max() < 1e-10
    X = scalekd_ref.dct2(x, wf).clone()
    X[:, :, 0, 0] = 0
    y = scalekd_ref.idct2(X, wi)
    assert (y - (x - x.mean(dim=(2, 3), keepdim=True))).abs().max() < 1e-10


def test_no_query_raises():
    sd = scalekd_ref.make_scalekd_state(8, 16, (2, 2), self_query=False)
    with pytest.raises(NotImplementedError):
        scalekd_ref.projector_forward(sd, "projector_0.", torch.randn(1, 8, 2, 2), None, hw=(2, 2), num_heads=2,
                                      softmax_scale=1.0)


needs_reference = pytest.mark.skipif(not ref_shims.reference_available(), reason="/root/reference not present")


@needs_reference
def test_port_matches_live_reference_scalekd():
    sk, _ = ref_shims.import_reference()
    kw = dict(name="scalekd_res5", alpha=[0.08, 0.06], student_dims=24, teacher_dims=48, query_hw=[5, 5], pos_hw=[5, 5],
              pos_dims=48, window_shapes=[1, 1], self_query=False, softmax_scale=[5.0, 3.0], num_heads=6)
    torch.manual_seed(0)
    m = sk.ScaleKD(**kw).train()
    S = torch.randn(3, 24, 5, 5, requires_grad=True)
    T = torch.randn(3, 48, 5, 5)
    qs, qf = torch.randn(3, 25, 48, requires_grad=True), torch.randn(3, 25, 48, requires_grad=True)
    sd = _grad_sd(m.state_dict())
    ref = m(S, T, query_s=qs, query_f=qf)
    ref["loss"].backward()
    S2, qs2, qf2 = (t.detach().clone().requires_grad_(True) for t in (S, qs, qf))
    out = scalekd_ref.scalekd_forward(sd, S2, T, qs2, qf2, alpha=kw["alpha"], hw=(5, 5), num_heads=6,
                                      softmax_scale=kw["softmax_scale"])
    out["loss"].backward()
    for k in ref:
        assert abs(out[k].item() - ref[k].item()) < 1e-5 * max(1, abs(ref[k].item())), k
    assert rel(S2.grad, S.grad) < 1e-4 and rel(qs2.grad, qs.grad) < 1e-4 and rel(qf2.grad, qf.grad) < 1e-4
    for k, p in m.named_parameters():
        if p.grad is not None and p.grad.norm() > 1e-6:
            assert rel(sd[k].grad, p.grad) < 5e-4, k


@needs_reference
def test_module_level_orchestration_matches_live_reference_compute_losses():
    """oracle compute_losses_modules (the loop the GPU tests drive over the B200 shells) against the reference's own
    DistillationModule._compute_losses, both over the reference's ScaleKD modules and a reference-shaped teacher."""
    import torch.nn as nn
    from oracle import dinov2_ref
    sk, dm = ref_shims.import_reference()
    if dm is None:
        pytest.skip("train.distillation_module not importable")
    cfg = dinov2_ref.VitCfg(64, 8, 4, 256)
    teacher = dinov2_ref.RefTeacher(cfg, dinov2_ref.make_state_dict(cfg, seed=21, pos_grid=4))
    common = dict(alpha=[0.08, 0.06], teacher_dims=64, query_hw=[4, 4], pos_hw=[4, 4], pos_dims=64,
                  window_shapes=[1, 1], softmax_scale=[5.0, 5.0])
    torch.manual_seed(31)
    mods = nn.ModuleDict({
        "scalekd_res4": sk.ScaleKD(**dict(common, name="scalekd_res4", student_dims=32, self_query=True, num_heads=4)),
        "scalekd_res5": sk.ScaleKD(**dict(common, name="scalekd_res5", student_dims=48, self_query=False, num_heads=8))}).train()
    weights = {"scalekd_res4": 2.0, "scalekd_res5": 0.5}
    mod = dm.DistillationModule.__new__(dm.DistillationModule)
    nn.Module.__init__(mod)
    mod.teacher, mod.losses, mod.loss_weights = teacher, mods, weights
    g = torch.Generator().manual_seed(32)
    T = teacher(torch.randn(2, 3, 56, 56, generator=g))["feature_map"]
    base = {"res4": torch.randn(2, 32, 4, 4, generator=g), "res5": torch.randn(2, 48, 4, 4, generator=g)}
    f1 = {k: v.clone().requires_grad_(True) for k, v in base.items()}
    ref = mod._compute_losses({"student": f1, "teacher": T})
    ref["loss"].backward()
    g1 = {k: p.grad.clone() for k, p in mods.named_parameters() if p.grad is not None}
    mods.zero_grad(set_to_none=True)
    f2 = {k: v.clone().requires_grad_(True) for k, v in base.items()}
    out = scalekd_ref.compute_losses_modules(mods, weights, f2, T, teacher)
    out["loss"].backward()
    assert list(out.keys()) == list(ref.keys())
    for k in ref:
        assert torch.allclose(out[k].detach(), ref[k].detach(), rtol=1e-6, atol=1e-7), k
    for k in f1:
        assert torch.allclose(f2[k].grad, f1[k].grad, rtol=1e-5, atol=1e-8), k
    for k, p in mods.named_parameters():
        if k in g1:
            assert torch.allclose(p.grad, g1[k], rtol=1e-5, atol=1e-8), k


@needs_reference
def test_closed_form_dct_matches_reference_fft_construction():
    sk, _ = ref_shims.import_reference()
    for R in (4, 16, 37):
        d = sk.DCT(resolution=R, device="cpu")
        wf, wi = scalekd_ref.dct_matrices(R)
        assert (d.forward_transform.weight - wf).abs().max() < 2e-5
        assert (d.inverse_transform.weight - wi).abs().max() < 2e-6


@needs_reference
def test_b200_module_state_dict_matches_reference():
    """Same names, shapes and (under the same seed) the same initial values as the reference module."""
    import warnings
    warnings.simplefilter("ignore")
    from dinov2_distillation_b200.scalekd import ScaleKD
    sk, _ = ref_shims.import_reference()
    kw = dict(name="scalekd_res4", alpha=[0.08, 0.06], student_dims=16, teacher_dims=32, query_hw=[3, 3], pos_hw=[3, 3],
              pos_dims=32, window_shapes=[1, 1], self_query=True, softmax_scale=[5.0, 5.0], num_heads=4)
    torch.manual_seed(7)
    a = sk.ScaleKD(**kw).state_dict()
    torch.manual_seed(7)
    b = ScaleKD(**kw).state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k


@pytest.mark.parametrize("n_in,n_out", [(7, 16), (14, 16), (17, 37), (33, 37), (16, 37), (32, 37), (16, 16), (20, 9)])
def test_resize_ref_matches_interpolate(n_in, n_out):
    """oracle/resize_ref.py against F.interpolate(bilinear, align_corners=False) -- the call ModelWrapper.forward makes
    (models/model_zoo.py:121-126) -- on the reference's own size pairs, forward and adjoint (through autograd)."""
    from oracle import resize_ref
    torch.manual_seed(n_in * 100 + n_out)
    x = torch.randn(2, 5, n_in, n_in + 1, dtype=torch.float64, requires_grad=True)
    size = (n_out, n_out + 2)
    ref = torch.nn.functional.interpolate(x, size=size, mode="bilinear", align_corners=False)
    got = resize_ref.resize_bilinear(x, size)
    assert torch.allclose(got, ref, atol=1e-12, rtol=0)
    g = torch.randn_like(ref)
    (gx_ref,) = torch.autograd.grad(ref, x, g, retain_graph=True)
    (gx,) = torch.autograd.grad(got, x, g)
    assert torch.allclose(gx, gx_ref, atol=1e-12, rtol=0)
    # rows of the 1-D operator sum to one: a 1x1 conv's bias commutes with the resize
    assert torch.allclose(resize_ref.resize_matrix(n_in, n_out).sum(1), torch.ones(n_out, dtype=torch.float64))


def test_resize_commutes_with_conv1x1():
    """The identity the fused path rests on: conv1x1(resize(x)) == resize(conv1x1(x)), bias included."""
    from oracle import resize_ref
    torch.manual_seed(0)
    x = torch.randn(2, 12, 7, 7, dtype=torch.float64)
    conv = torch.nn.Conv2d(12, 20, 1).double()
    a = conv(torch.nn.functional.interpolate(x, size=(16, 16), mode="bilinear", align_corners=False))
    b = resize_ref.resize_bilinear(conv(x), (16, 16))
    assert torch.allclose(a, b, atol=1e-12, rtol=0)


@pytest.mark.parametrize("tag", ["self", "ext"])
def test_port_matches_golden_scalekd_windows(tag):
    """window_shapes = [2, 2]: the port's separate_tokens / window-major output (losses/scalekd.py:305-314, :326-335)
    against the unmodified reference's outputs and gradients (oracle/make_golden.py: scalekd_windows)."""
    g = torch.load(os.path.join(GOLDEN, f"scalekd_win_{tag}.pt"))
    kw = g["kwargs"]
    sd = _grad_sd(g["state_dict"])
    S = g["preds_S"].clone().requires_grad_(True)
    out = scalekd_ref.scalekd_forward(sd, S, g["preds_T"], g["query_s"], g["query_f"], alpha=kw["alpha"],
                                      hw=kw["query_hw"], num_heads=kw["num_heads"], softmax_scale=kw["softmax_scale"],
                                      window_shapes=kw["window_shapes"])
    for k, v in g["out"].items():
        assert abs(out[k].item() - v.item()) <= 1e-5 * max(1.0, abs(v.item())), k
    out["loss"].backward()
    assert rel(S.grad, g["grad_S"]) < 1e-4
    for k, ref in g["grads"].items():
        if ref.norm() < 1e-6:
            continue
        assert rel(sd[k].grad, ref) < 2e-4, (k, rel(sd[k].grad, ref))
    # and the windows matter: the same weights without windows give a different loss
    out1 = scalekd_ref.scalekd_forward(_grad_sd(g["state_dict"]), g["preds_S"], g["preds_T"], g["query_s"], g["query_f"],
                                       alpha=kw["alpha"], hw=kw["query_hw"], num_heads=kw["num_heads"],
                                       softmax_scale=kw["softmax_scale"])
    assert abs(out1["loss"].item() - g["out"]["loss"].item()) > 1e-4
