"""CPU: host-side logic of the drop-in shells -- schema handling, error behaviour (no silent CPU fallback), batch
sharding, the flat gradient arena and the 2-rank gloo data-parallel path."""
import os
import socket
import warnings

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

warnings.simplefilter("ignore")


def _kw(**over):
    kw = dict(name="scalekd_res5", alpha=[0.08, 0.06], student_dims=16, teacher_dims=32, query_hw=[3, 3], pos_hw=[3, 3],
              pos_dims=32, window_shapes=[1, 1], self_query=True, softmax_scale=[5.0, 5.0], num_heads=4)
    kw.update(over)
    return kw


def test_scalekd_schema_and_parameter_names():
    from dinov2_distillation_b200.scalekd import LOSS_REGISTRY, ScaleKD
    assert LOSS_REGISTRY["scalekd"] is ScaleKD
    m = ScaleKD(**_kw())
    keys = set(m.state_dict().keys())
    for pj in ("projector_0", "projector_1"):
        for k in ("pos_embed", "proj_student.0.weight", "proj_student.0.bias", "proj_student.1.weight",
                  "proj_student.1.running_mean", "proj_student.1.running_var", "proj_student.1.num_batches_tracked",
                  "pos_attention.q.weight", "pos_attention.k.bias", "pos_attention.v.weight", "pos_attention.proj.bias",
                  "ffn.layers.0.0.weight", "ffn.layers.1.bias", "norm.weight", "norm_2.bias", "query.weight"):
            assert f"{pj}.{k}" in keys, k
    assert m.projector_0.pos_embed.shape == (1, 32, 3, 3)
    m2 = ScaleKD(**_kw(self_query=False))
    assert not any("query.weight" in k for k in m2.state_dict())
    # vits14 / Cs=512 / 16x16 / self_query: SURVEY.md section 8 (a9) counts 4 337 664 parameters
    big = ScaleKD(**_kw(student_dims=512, teacher_dims=384, query_hw=[16, 16], pos_hw=[16, 16], pos_dims=384, num_heads=24))
    assert sum(p.numel() for p in big.parameters()) == 4337664


def test_scalekd_errors_like_the_reference_and_never_falls_back_to_cpu():
    from dinov2_distillation_b200 import _lib
    from dinov2_distillation_b200.scalekd import ScaleKD
    m = ScaleKD(**_kw(self_query=False))
    with pytest.raises(NotImplementedError, match="There is no query"):
        m.project_feat_spat(torch.randn(1, 16, 3, 3))
    m = ScaleKD(**_kw())
    with pytest.raises(_lib.B200Error, match="no CPU fallback"):
        m(torch.randn(1, 16, 3, 3), torch.randn(1, 32, 3, 3))
    with pytest.raises(_lib.B200Error, match="no CPU fallback"):
        m.get_spat_loss(torch.randn(1, 9, 32), torch.randn(1, 32, 3, 3))
    with pytest.raises(ValueError):
        ScaleKD(**_kw(teacher_dims=30, pos_dims=30, num_heads=4))   # reference: RuntimeError inside reshape
    with pytest.raises(ValueError, match="window_shapes"):   # 2x2 windows do not tile a 3x3 grid (reference: .view fails)
        ScaleKD(**_kw(window_shapes=[2, 2])).cpu().projector_0(torch.randn(1, 16, 3, 3))


def test_teacher_shell_surface():
    from dinov2_distillation_b200 import _lib
    from dinov2_distillation_b200.teacher import DINOv2ViT, TEACHER_CONFIGS
    assert sorted(TEACHER_CONFIGS) == ["dinov2_vitb14", "dinov2_vitg14", "dinov2_vitl14", "dinov2_vits14"]
    with pytest.raises(KeyError):
        DINOv2ViT("dinov2_vitx14")
    t = DINOv2ViT("dinov2_vits14", weights="synthetic")
    assert all(not p.requires_grad for p in t.parameters())
    assert len(t.model.blocks) == 12 and not t.model.training
    keys = set(t.model.state_dict().keys())
    for k in ("cls_token", "pos_embed", "mask_token", "patch_embed.proj.weight", "blocks.0.norm1.weight",
              "blocks.11.attn.qkv.bias", "blocks.3.ls1.gamma", "blocks.5.mlp.fc2.weight", "norm.bias"):
        assert k in keys, k
    assert t.model.pos_embed.shape == (1, 1370, 384)
    with pytest.raises(_lib.B200Error, match="no CPU fallback"):
        t(torch.randn(1, 3, 224, 224))
    with pytest.raises(_lib.B200Error, match="no CPU fallback"):
        t.model.blocks[9](torch.randn(1, 256, 384))
    g = DINOv2ViT("dinov2_vitg14", weights="synthetic")
    assert "blocks.0.mlp.w12.weight" in g.model.state_dict() and g.model.blocks[0].mlp.w12.weight.shape == (8192, 1536)
    assert abs(sum(p.numel() for p in g.parameters()) - 1136.5e6) / 1136.5e6 < 0.005


def test_teacher_requires_real_weights_unless_synthetic_is_requested(tmp_path, monkeypatch):
    """models/backbones/dinov2.py:20 always loads pretrained weights: a missing / mistyped path must raise, never fall
    back to a random teacher; a hub-format checkpoint on disk loads strictly (file or directory form)."""
    from dinov2_distillation_b200.teacher import DINOv2ViT
    monkeypatch.delenv("DINOV2_WEIGHTS_DIR", raising=False)
    with pytest.raises(FileNotFoundError, match="no pretrained weights"):
        DINOv2ViT("dinov2_vits14")
    with pytest.raises(FileNotFoundError, match="not found"):
        DINOv2ViT("dinov2_vits14", weights=str(tmp_path / "nope.pth"))
    monkeypatch.setenv("DINOV2_WEIGHTS_DIR", str(tmp_path))
    with pytest.raises(FileNotFoundError, match="not found"):
        DINOv2ViT("dinov2_vits14")
    src = DINOv2ViT("dinov2_vits14", weights="synthetic", seed=7)
    torch.save(src.model.state_dict(), tmp_path / "dinov2_vits14_pretrain.pth")
    for t in (DINOv2ViT("dinov2_vits14"), DINOv2ViT("dinov2_vits14", weights=str(tmp_path / "dinov2_vits14_pretrain.pth"))):
        for k, v in src.model.state_dict().items():
            assert torch.equal(t.model.state_dict()[k], v), k


def test_distillation_step_mirrors_reference_orchestration():
    from dinov2_distillation_b200.distill import DistillationStep
    from dinov2_distillation_b200.teacher import DINOv2ViT
    specs = [{"type": "scalekd", "weight": 1.0, "kwargs": _kw(name="scalekd_res4", teacher_dims=384, pos_dims=384, num_heads=16)},
             {"type": "scalekd", "weight": 0.5, "kwargs": _kw(name="scalekd_res5", teacher_dims=384, pos_dims=384, num_heads=24, self_query=False)}]
    step = DistillationStep(None, DINOv2ViT("dinov2_vits14", weights="synthetic"), specs)
    assert sorted(step.losses.keys()) == ["scalekd_res4", "scalekd_res5"]
    assert step.loss_weights == {"scalekd_res4": 1.0, "scalekd_res5": 0.5}
    step.train()
    assert not step.teacher.training and step.losses.training

    class Rec(torch.nn.Module):
        def __init__(self, i, log):
            super().__init__()
            self.i, self.log = i, log

        def forward(self, x):
            self.log.append(self.i)
            return x

    log = []
    step.teacher.model.blocks = torch.nn.ModuleList([Rec(i, log) for i in range(12)])
    step._forward_specific_stage(torch.zeros(1, 4, 8), "res4")
    assert log == [9, 10]
    log.clear()
    step._forward_specific_stage(torch.zeros(1, 4, 8), "res3")
    assert log == []


def test_shard_batch_covers_everything_once():
    from dinov2_distillation_b200.distributed import shard_batch
    for gb, w in ((256, 8), (64, 1), (10, 4), (7, 8)):
        seen = [i for r in range(w) for i in shard_batch(gb, r, w)]
        assert seen == list(range(gb))


def test_flat_grad_arena_accumulates_autograd_in_place():
    from dinov2_distillation_b200.distributed import FlatGradArena
    lin = torch.nn.Linear(4, 3)
    arena = FlatGradArena(lin.parameters(), extra_numel=5)
    assert arena.numel == 4 * 3 + 3 + 5 and arena.extra.numel() == 5
    x = torch.randn(2, 4)
    lin(x).sum().backward()
    assert lin.weight.grad.data_ptr() == arena.buffer.data_ptr()
    assert torch.allclose(arena.buffer[:12].view(3, 4), x.sum(0).expand(3, 4))
    arena.zero()
    assert arena.buffer.abs().sum() == 0 and lin.weight.grad.data_ptr() == arena.buffer.data_ptr()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _dp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from dinov2_distillation_b200 import distributed as D
    r, w, _ = D.init_from_env(backend="gloo")
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))
    arena = D.FlatGradArena(model.parameters())
    data = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10
    idx = list(D.shard_batch(8, r, w))
    arena.zero()
    loss = model(data[idx]).pow(2).sum() / len(idx)      # per-rank mean, like the reference's loss / N (scalekd.py:86)
    loss.backward()
    arena.allreduce_mean()
    m = D.reduce_metrics({"loss": loss.detach(), "b": torch.tensor(float(r))})
    q.put((r, arena.buffer.clone(), float(m["loss"]), float(m["b"])))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_allreduce_matches_single_process():
    """world_size=2 on CPU (gloo): sharded batch + one flat mean-allreduce == the full-batch gradient."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.ReLU(), torch.nn.Linear(5, 1))
    data = torch.arange(8 * 6, dtype=torch.float32).view(8, 6) / 10
    full = model(data).pow(2).sum() / 8
    full.backward()
    ref = torch.cat([p.grad.flatten() for p in model.parameters()])
    assert torch.allclose(res[0][1], ref, atol=1e-5) and torch.allclose(res[1][1], ref, atol=1e-5)
    assert abs(res[0][2] - full.item()) < 1e-5 and abs(res[0][3] - 0.5) < 1e-6


def test_plugin_install_rebinds_the_reference_plugin_points():
    from oracle import ref_shims
    if not ref_shims.reference_available():
        pytest.skip("/root/reference not present")
    sk, dm = ref_shims.import_reference()
    ref_scalekd = sk.ScaleKD
    os.environ["NCCL_P2P_DISABLE"] = "1"
    import dinov2_distillation_b200.plugin as plugin
    from dinov2_distillation_b200.scalekd import ScaleKD
    try:
        touched = plugin.install()
        assert dm.LOSS_REGISTRY["scalekd"] is ScaleKD
        assert "NCCL_P2P_DISABLE" not in os.environ
        assert any("LOSS_REGISTRY" in k for k in touched)
        # the reference's own _initialize_loss now builds B200 modules from the config schema
        mod = dm.DistillationModule.__new__(dm.DistillationModule)
        torch.nn.Module.__init__(mod)
        mod.cfg = type("Cfg", (), {"loss": {"losses": [{"type": "scalekd", "weight": 1, "kwargs": _kw(name="scalekd_res5")}]}})()
        mod._initialize_loss()
        assert isinstance(mod.losses["scalekd_res5"], ScaleKD)
    finally:
        dm.LOSS_REGISTRY["scalekd"] = ref_scalekd
        sk.ScaleKD = ref_scalekd
        import losses
        losses.ScaleKD = ref_scalekd
        dm.ScaleKD = ref_scalekd


def test_compute_losses_dict_weights_and_gradients_with_stub_losses():
    """`_compute_losses` (train/distillation_module.py:180-246) on CPU with stub loss modules: key set and order, the
    res4 'frequency' term scored with get_spat_loss, the break after res5, query chaining, and the fused combine
    (`_CombineLosses`): every weighted entry and d(loss)/d(term) equal the reference's scalar arithmetic."""
    from dinov2_distillation_b200.distill import DistillationStep
    from dinov2_distillation_b200.teacher import DINOv2ViT

    class StubKD(torch.nn.Module):
        def __init__(self, base):
            super().__init__()
            self.p = torch.nn.Parameter(torch.tensor(float(base)))
            self.calls = []

        def project_feat_spat(self, x, query=None):
            self.calls.append(("spat", query is not None))
            return x * self.p

        def project_feat_freq(self, x, query=None):
            self.calls.append(("freq", query is not None))
            return x * (self.p + 1.0)

        def get_spat_loss(self, s, t):
            self.calls.append(("get_spat_loss",))
            return (s.sum() - t.sum()) ** 2, s.mean()

        def forward(self, s, t, query_s=None, query_f=None):
            self.calls.append(("forward", query_s is not None, query_f is not None))
            a, b = (s * self.p).sum() ** 2, (s * self.p).mean() ** 2
            return {"spatial_loss": a, "frequency_loss": b, "spatial_similarity": s.mean(), "frequency_similarity": s.std(),
                    "loss": a + b}

    specs = [{"type": "scalekd", "weight": 2.0, "kwargs": _kw(name="scalekd_res4", teacher_dims=384, pos_dims=384, num_heads=16)},
             {"type": "scalekd", "weight": 0.5, "kwargs": _kw(name="scalekd_res5", teacher_dims=384, pos_dims=384, num_heads=24, self_query=False)},
             {"type": "scalekd", "weight": 9.0, "kwargs": _kw(name="scalekd_res6", teacher_dims=384, pos_dims=384, num_heads=24)}]
    step = DistillationStep(None, DINOv2ViT("dinov2_vits14", weights="synthetic"), specs)
    step.two_streams = False
    step.losses = torch.nn.ModuleDict({"scalekd_res4": StubKD(1.5), "scalekd_res5": StubKD(0.7), "scalekd_res6": StubKD(3.0)})
    step._forward_specific_stage = lambda feat, layer: feat + 1.0
    f4 = torch.randn(2, 3, requires_grad=True)
    f5 = torch.randn(2, 3, requires_grad=True)
    T = torch.randn(2, 3)
    out = step._compute_losses({"student": {"res4": f4, "res5": f5, "res6": f5}, "teacher": T})
    assert list(out.keys()) == [
        "scalekd_res4_total_loss", "scalekd_res4_frequency_loss", "scalekd_res4_spatial_loss",
        "scalekd_res4_spatial_similarity", "scalekd_res4_frequency_similarity",
        "scalekd_res5_total_loss", "scalekd_res5_frequency_loss", "scalekd_res5_spatial_loss",
        "scalekd_res5_spatial_similarity", "scalekd_res5_frequency_similarity", "loss"]      # res6 never runs (break)
    k4, k5 = step.losses["scalekd_res4"], step.losses["scalekd_res5"]
    assert k4.calls == [("spat", False), ("freq", False), ("get_spat_loss",), ("get_spat_loss",)]
    assert k5.calls == [("forward", True, True)] and step.losses["scalekd_res6"].calls == []
    # the reference's arithmetic, written out
    s4 = ((f4 * 1.5 + 1.0).sum() - T.sum()) ** 2
    q4 = ((f4 * 2.5 + 1.0).sum() - T.sum()) ** 2
    a5, b5 = (f5 * 0.7).sum() ** 2, (f5 * 0.7).mean() ** 2
    ref = {"scalekd_res4_total_loss": (s4 + q4) * 2.0, "scalekd_res4_frequency_loss": q4 * 2.0, "scalekd_res4_spatial_loss": s4 * 2.0,
           "scalekd_res5_total_loss": (a5 + b5) * 0.5, "scalekd_res5_frequency_loss": b5 * 0.5, "scalekd_res5_spatial_loss": a5 * 0.5}
    ref["loss"] = ref["scalekd_res4_total_loss"] + ref["scalekd_res5_total_loss"]
    for k, v in ref.items():
        assert torch.allclose(out[k], v.detach(), rtol=1e-5), k
    g_ref = torch.autograd.grad(ref["loss"], [f4, f5], retain_graph=True)
    g = torch.autograd.grad(out["loss"], [f4, f5], retain_graph=True)
    for a, b in zip(g, g_ref):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)
    # a weighted entry other than the total back-propagates too
    (g4,) = torch.autograd.grad(out["scalekd_res4_frequency_loss"], [f4])
    (g4_ref,) = torch.autograd.grad(ref["scalekd_res4_frequency_loss"], [f4])
    assert torch.allclose(g4, g4_ref, rtol=1e-5, atol=1e-6)


def test_teacher_feature_cache_host_logic():
    """slot hand-out, capacity, all-or-nothing lookups -- and no CPU pool (no CPU fallback)."""
    from dinov2_distillation_b200._lib import B200Error
    from dinov2_distillation_b200.feature_cache import TeacherFeatureCache
    c = TeacherFeatureCache(capacity=3, tokens=4, dim=8, device="cpu")
    assert c.slots_for([10, 11], allocate=False) == [-1, -1] and len(c) == 0
    assert c.slots_for([10, 11, 10], allocate=True) == [0, 1, 0]
    assert c.slots_for([12, 13], allocate=True) == [2, -1]          # full: 13 is not cached
    assert len(c) == 3 and 12 in c and 13 not in c
    assert c.all_cached([10, 12]) and not c.all_cached([10, 13])
    with pytest.raises(B200Error):
        c.load([10])
    with pytest.raises(ValueError):
        TeacherFeatureCache(capacity=1, tokens=4, dim=6, device="cpu")
    with pytest.raises(ValueError):
        TeacherFeatureCache(capacity=1, tokens=4, dim=8, device="cpu", dtype=torch.float16)
