"""Host-side mirror of the reference's `DistillationModule` hot methods (train/distillation_module.py):
`_initialize_loss` (:112-137), `_forward_specific_stage` (:139-178), `_compute_losses` (:180-246),
`_extract_features` (:311-337) and `training_step` (:247-276) -- same names, same dict keys, same quirks -- for users
who run the distillation step without Lightning (bench.py, smoke(), the parity tests). With the real reference, use
`plugin.install()` instead and keep its own DistillationModule.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .scalekd import LOSS_REGISTRY

_STAGE_FRACTION = {"res2": 0.25, "res3": 0.50, "res4": 0.75}


class _CombineLosses(torch.autograd.Function):
    """Every weighted entry of the loss dict and the total from the raw (spatial, frequency) terms of the processed
    stages: out = A @ stack(terms) -- one concat and one matrix-vector product instead of ~13 scalar add / mul kernels,
    and one kernel in backward (A[row] * g) instead of a mul / select chain per term. A is a small constant matrix
    ([3 * stages + 1, 2 * stages]: per stage (s + f) * w, f * w, s * w; last row the total, train/distillation_module.py
    :225-246); the outputs are views of one buffer."""

    @staticmethod
    def forward(ctx, A, *terms):
        out = torch.mv(A, torch.stack(terms))
        ctx.save_for_backward(A)
        ctx.set_materialize_grads(False)
        return out.unbind(0)

    @staticmethod
    def backward(ctx, *grads):
        (A,) = ctx.saved_tensors
        gx = None
        for i, g in enumerate(grads):
            if g is not None:
                gx = A[i] * g if gx is None else torch.addcmul(gx, A[i], g)
        if gx is None:
            return (None,) * (1 + A.shape[1])
        return (None, *gx.unbind(0))


class DistillationStep(nn.Module):
    """student: callable images -> {'res4': [B,C,H,W], ...} (already resized to the teacher grid, as ModelWrapper does,
    models/model_zoo.py:118-128) or None when features are supplied directly; teacher: DINOv2ViT-like;
    loss_specs: the `loss.losses` list of the config (type / weight / kwargs)."""

    def __init__(self, student: Optional[nn.Module], teacher: nn.Module, loss_specs: List[dict],
                 teacher_key: str = "feature_map"):
        super().__init__()
        self.student = student
        self.teacher = teacher
        self.teacher_key = teacher_key
        self.teacher.eval()
        for p in self.teacher.parameters():
            p.requires_grad = False
        self.losses = nn.ModuleDict()
        self.loss_weights: Dict[str, float] = {}
        for spec in loss_specs:
            kwargs = dict(spec["kwargs"])
            name = kwargs.get("name", spec["type"])
            self.losses[name] = LOSS_REGISTRY[spec["type"]](**kwargs)
            self.loss_weights[name] = spec["weight"]

    def train(self, mode: bool = True):
        super().train(mode)
        self.teacher.eval()  # frozen (distillation_module.py:104-108)
        return self

    def _forward_specific_stage(self, feat, layer):
        n_total = len(self.teacher.model.blocks)
        start = int(n_total * _STAGE_FRACTION[layer])
        end = n_total - 1 if layer == "res4" else int(n_total / 4) - 1
        for i in range(start, end):
            feat = self.teacher.model.blocks[i](feat)
        return feat

    # The spatial and the frequency branch of a stage are independent until their losses are added (each: projector ->
    # re-used teacher blocks -> loss term). With `two_streams` they are issued on two CUDA streams: the kernels of one
    # branch fill the SMs the partial last wave of the other leaves idle (a 16 384 x 384 GEMM is 1.7 waves of tiles), and
    # autograd replays each branch's backward on its forward stream, so the backward overlaps the same way. Stream
    # fork / join is captured by GraphedDistillStep like any other dependency.
    two_streams = True

    def _branch_streams(self, ref: torch.Tensor):
        if not (self.two_streams and ref.is_cuda):
            return None, None
        if getattr(self, "_side_stream", None) is None or self._side_stream.device != ref.device:
            self._side_stream = torch.cuda.Stream(device=ref.device)
        return torch.cuda.current_stream(ref.device), self._side_stream

    def teacher_async(self, batch):
        """Frozen-teacher forward on its own CUDA stream: the projectors and the re-used teacher blocks of the loss path
        do not read the teacher features until the loss terms, so the ~90-kernel teacher chain overlaps them. Returns
        (feature_map, event); pass the event as features['teacher_ready'] and `_compute_losses` waits for it right
        before the first use."""
        dev = batch.device
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_teacher_stream", None) is None or self._teacher_stream.device != dev:
            self._teacher_stream = torch.cuda.Stream(device=dev)
        ts = self._teacher_stream
        # the teacher's cached bf16 weight pack (and the position embedding of this grid) are built on first use: do that
        # on the MAIN stream, before the fork -- the re-used teacher blocks read the same buffers from the main / side
        # streams, which only join the teacher stream at the loss terms
        model = getattr(self.teacher, "model", None)
        if model is not None and hasattr(model, "_ensure_pack"):
            model._ensure_pack()
            if hasattr(model, "_pos_for_grid") and batch.dim() == 4:
                model._pos_for_grid(batch.shape[2] // 14, batch.shape[3] // 14)
        ts.wait_stream(main)
        with torch.cuda.stream(ts), torch.no_grad():
            T = self.teacher(batch)[self.teacher_key]
            ev = torch.cuda.Event()
            ev.record(ts)
        T.record_stream(main)
        return T, ev

    @staticmethod
    def _await(ready):
        if ready is not None:
            torch.cuda.current_stream().wait_event(ready)

    def _combine_matrix(self, names, device):
        key = (tuple(names), tuple(float(self.loss_weights[n]) for n in names), str(device))
        cache = getattr(self, "_combine_cache", None)
        if cache is None or cache[0] != key:
            n = len(names)
            A = torch.zeros(3 * n + 1, 2 * n)
            for i, nm in enumerate(names):
                w = float(self.loss_weights[nm])
                A[3 * i, 2 * i] = A[3 * i, 2 * i + 1] = w      # total = (spatial + frequency) * weight
                A[3 * i + 1, 2 * i + 1] = w                    # frequency * weight
                A[3 * i + 2, 2 * i] = w                        # spatial * weight
                A[3 * n, 2 * i] = A[3 * n, 2 * i + 1] = w      # 'loss'
            self._combine_cache = cache = (key, A.to(device))
        return cache[1]

    def _compute_losses(self, features):
        loss_dict = {}
        spatial_query = frequency_query = None
        ready = features.get("teacher_ready")
        done, terms = [], []
        for name in sorted(self.losses.keys()):
            layer = name.split("_")[1]
            loss_fn = self.losses[name]
            s_feat = features["student"][layer]
            main, side = self._branch_streams(s_feat)
            if side is not None and ready is not None:
                features["teacher"].record_stream(side)   # produced on the teacher stream, read on both branches
            if "res5" in name:
                if side is not None and hasattr(loss_fn, "forward_two_streams"):
                    loss = loss_fn.forward_two_streams(s_feat, features["teacher"], spatial_query, frequency_query, main, side,
                                                       teacher_ready=ready)
                else:
                    self._await(ready)
                    loss = loss_fn(s_feat, features["teacher"], query_s=spatial_query, query_f=frequency_query)
                done.append(name)
                terms += [loss["spatial_loss"], loss["frequency_loss"]]
                loss_dict[f"{name}_spatial_similarity"] = loss["spatial_similarity"]
                loss_dict[f"{name}_frequency_similarity"] = loss["frequency_similarity"]
                break
            if side is None or not hasattr(loss_fn, "tokenize_for_both"):
                feat_spat = loss_fn.project_feat_spat(s_feat, query=spatial_query)
                feat_freq = loss_fn.project_feat_freq(s_feat, query=frequency_query)
                feat_spat = self._forward_specific_stage(feat_spat, layer)
                feat_freq = self._forward_specific_stage(feat_freq, layer)
                self._await(ready)
                spatial_loss, spatial_similarity = loss_fn.get_spat_loss(feat_spat, features["teacher"])
                # sic: the reference scores the "frequency" branch of non-res5 stages with the SPATIAL loss (:236-237)
                frequency_loss, frequency_similarity = loss_fn.get_spat_loss(feat_freq, features["teacher"])
            else:
                tok = loss_fn.tokenize_for_both(s_feat)          # shared by both projectors; before the fork
                side.wait_stream(main)
                feat_spat = loss_fn.projector_0(s_feat, query=spatial_query, tokens=tok)
                feat_spat = self._forward_specific_stage(feat_spat, layer)
                with torch.cuda.stream(side):
                    feat_freq = loss_fn.projector_1(s_feat, query=frequency_query, tokens=tok)
                    feat_freq = self._forward_specific_stage(feat_freq, layer)
                    self._await(ready)
                    frequency_loss, frequency_similarity = loss_fn.get_spat_loss(feat_freq, features["teacher"])
                self._await(ready)
                spatial_loss, spatial_similarity = loss_fn.get_spat_loss(feat_spat, features["teacher"])
                main.wait_stream(side)
                for t in (feat_freq, frequency_loss, frequency_similarity):   # allocated on `side`, read on `main`
                    t.record_stream(main)
            spatial_query, frequency_query = feat_spat, feat_freq
            done.append(name)
            terms += [spatial_loss, frequency_loss]
            loss_dict[f"{name}_spatial_similarity"] = spatial_similarity
            loss_dict[f"{name}_frequency_similarity"] = frequency_similarity
        if not done:
            loss_dict["loss"] = 0
            return loss_dict
        # weighted entries and the total (train/distillation_module.py:225-246) in one fused combine
        outs = _CombineLosses.apply(self._combine_matrix(done, terms[0].device), *terms)
        ordered = {}                                   # key order of the reference's dict (:225-246)
        for i, name in enumerate(done):
            ordered[f"{name}_total_loss"] = outs[3 * i]
            ordered[f"{name}_frequency_loss"] = outs[3 * i + 1]
            ordered[f"{name}_spatial_loss"] = outs[3 * i + 2]
            ordered[f"{name}_spatial_similarity"] = loss_dict[f"{name}_spatial_similarity"]
            ordered[f"{name}_frequency_similarity"] = loss_dict[f"{name}_frequency_similarity"]
        ordered["loss"] = outs[3 * len(done)]
        return ordered

    def _extract_features(self, batch):
        if batch.is_cuda and self.two_streams:
            teacher_features, ready = self.teacher_async(batch)     # overlaps the student and the projectors
            return {"student": self.student(batch), "teacher": teacher_features, "teacher_ready": ready}
        with torch.no_grad():
            teacher_features = self.teacher(batch)[self.teacher_key]
        return {"student": self.student(batch), "teacher": teacher_features}

    def training_step(self, batch, batch_idx=0):
        losses = self._compute_losses(self._extract_features(batch))
        self.last_losses = losses
        return losses["loss"]

    forward = training_step


class GraphedDistillStep:
    """The hot path of one training step captured ONCE into a CUDA graph and replayed: teacher forward -> ScaleKD
    forward -> backward to the student features and the ScaleKD parameters (~450 kernel launches at config.yaml
    shapes), with no Python, ctypes or allocator work on the critical path. Static input buffers are refreshed by
    (async) copies; outputs are static tensors that the next replay overwrites.

        g = GraphedDistillStep(step, img_example, {'res4': f4, 'res5': f5}, arena)
        losses, feat_grads = g(img, {'res4': f4, 'res5': f5})     # feat_grads feed the student's own backward

    The student network stays outside the graph (stock PyTorch, as in the reference)."""

    def __init__(self, step: "DistillationStep", img: torch.Tensor, feats: Dict[str, torch.Tensor], arena=None,
                 warmup: int = 3, split_teacher: bool = False):
        """split_teacher: capture TWO graphs -- the frozen teacher forward, and everything that touches trainable state
        (projectors, loss terms, backward). A data-parallel loop can then keep the gradient all-reduce of step i in
        flight under the teacher forward of step i + 1 (which depends on nothing a step changes) and only join it
        before the second graph: `run_teacher()` ... `arena.wait()` ... `run_losses()`."""
        if not img.is_cuda:
            raise RuntimeError("GraphedDistillStep needs CUDA tensors: there is no CPU fallback")
        self.step, self.arena = step, arena
        self.split_teacher = bool(split_teacher) and bool(feats)
        if arena is not None:
            arena.enable_direct_accumulation(step.losses)
        self.img = img.detach().clone()
        self.feats = {k: v.detach().clone().requires_grad_(True) for k, v in feats.items()}
        # module state the warm-up steps would otherwise leave behind: BatchNorm running statistics / counters advance
        # on every training forward, and parameter gradients accumulate when no arena zeroes them
        state = {n: b.detach().clone() for n, b in step.losses.named_buffers()}
        grads = {n: (None if p.grad is None else p.grad.detach().clone()) for n, p in step.losses.named_parameters()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if self.split_teacher:
            self.graph_teacher = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_teacher):
                self._T = self.step.teacher(self.img)[self.step.teacher_key]
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, pool=self.graph_teacher.pool()):
                self.out = self._run_losses(self._T)
        else:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.out = self._run()
        self.feat_grads = {k: v.grad for k, v in self.feats.items()}
        # undo what warm-up and capture did to the module: the first replay is step 0, as in the reference
        with torch.no_grad():
            for n, b in step.losses.named_buffers():
                b.copy_(state[n])
            if arena is not None:
                arena.zero()
            else:
                for n, p in step.losses.named_parameters():
                    if grads[n] is None:
                        if p.grad is not None:
                            p.grad.zero_()   # (the captured graph accumulates into this tensor: it must stay allocated)
                    else:
                        p.grad.copy_(grads[n])

    def _run_losses(self, T, ready=None):
        if self.arena is not None:
            self.arena.zero()
        for f in self.feats.values():
            f.grad = None
        feats = {"student": self.feats, "teacher": T}
        if ready is not None:
            feats["teacher_ready"] = ready
        out = self.step._compute_losses(feats)
        out["loss"].backward()
        return {k: v.detach() for k, v in out.items()}

    def _run(self):
        if not self.feats:
            return {"teacher": self.step.teacher(self.img)[self.step.teacher_key]}
        if self.split_teacher or not self.step.two_streams:
            return self._run_losses(self.step.teacher(self.img)[self.step.teacher_key])
        T, ready = self.step.teacher_async(self.img)
        return self._run_losses(T, ready)

    # ---- split form (data parallel): teacher forward and the trainable part as separate replays
    def run_teacher(self, img: Optional[torch.Tensor] = None) -> None:
        if img is not None and img.data_ptr() != self.img.data_ptr():
            self.img.copy_(img, non_blocking=True)
        self.graph_teacher.replay()

    def run_losses(self, feats: Optional[Dict[str, torch.Tensor]] = None):
        if feats is not None:
            for k, v in feats.items():
                if v.data_ptr() != self.feats[k].data_ptr():
                    self.feats[k].data.copy_(v, non_blocking=True)
        self.graph.replay()
        return self.out, self.feat_grads

    def __call__(self, img: Optional[torch.Tensor] = None, feats: Optional[Dict[str, torch.Tensor]] = None):
        if img is not None and img.data_ptr() != self.img.data_ptr():
            self.img.copy_(img, non_blocking=True)
        if feats is not None:
            for k, v in feats.items():
                if v.data_ptr() != self.feats[k].data_ptr():
                    self.feats[k].data.copy_(v, non_blocking=True)
        if self.split_teacher:
            self.graph_teacher.replay()
        self.graph.replay()
        return self.out, self.feat_grads

    # ---- input pipeline: the NEXT step's host batch travels H2D on a side stream while the current step computes
    def stage_inputs(self, img_host: torch.Tensor, feats_host: Optional[Dict[str, torch.Tensor]] = None) -> None:
        """Start the asynchronous host->device copy of the next step's (pinned) inputs into device staging buffers.
        `run_staged()` then moves them into the graph's static inputs (device->device, a few tens of microseconds)
        and replays. One staging set is enough: the copy waits until the previous set has been consumed."""
        if not hasattr(self, "_copy_stream"):
            self._copy_stream = torch.cuda.Stream()
            self._stage_img = torch.empty_like(self.img)
            self._stage_feats = {k: torch.empty_like(v.data) for k, v in self.feats.items()}
            self._stage_ready = torch.cuda.Event()
            self._stage_free = torch.cuda.Event()
            self._stage_free.record(torch.cuda.current_stream())
        cs = self._copy_stream
        cs.wait_event(self._stage_free)
        with torch.cuda.stream(cs):
            self._stage_img.copy_(img_host, non_blocking=True)
            for k, v in (feats_host or {}).items():
                self._stage_feats[k].copy_(v, non_blocking=True)
            self._stage_ready.record(cs)

    def consume_staged(self) -> None:
        """Move the staged inputs (see `stage_inputs`) into the graphs' static input buffers."""
        cur = torch.cuda.current_stream()
        cur.wait_event(self._stage_ready)
        self.img.copy_(self._stage_img, non_blocking=True)
        for k, v in self._stage_feats.items():
            self.feats[k].data.copy_(v, non_blocking=True)
        self._stage_free.record(cur)

    def run_staged(self):
        """Consume the staged inputs and replay the captured step."""
        self.consume_staged()
        if self.split_teacher:
            self.graph_teacher.replay()
        self.graph.replay()
        return self.out, self.feat_grads
