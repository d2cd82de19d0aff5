"""Drop-in for the reference's `ScaleKD` loss (losses/scalekd.py:12-127) and its `AttentionProjector` (:177-245).

Extension: `preds_S` may also be the RAW backbone map [B, Cs, h, w]; the bilinear resize of
`ModelWrapper.forward` (models/model_zoo.py:121-126) is then applied inside, after the 1x1 conv it commutes with.

Same constructor kwargs (the `loss.losses[*].kwargs` schema of config/config.yaml:40-59 plus the keys train.py injects),
same parameter / buffer names and shapes (so checkpoints and scripts/convert_to_anyma.py keep working), same methods
(`forward`, `project_feat_spat/freq`, `get_spat_loss`, `get_freq_loss`) and the same 5-key output dict. Forward and
backward run in libb200distill.so; the nn.Modules below only own the fp32 master parameters.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn
from torch.nn.init import trunc_normal_

from . import _lib as L
from . import ops


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------------------------------------ parameter containers
class _PosAttention(nn.Module):
    """Parameters of WindowMultiheadPosAttention (losses/scalekd.py:277-282)."""

    def __init__(self, embed_dims, num_heads, pos_dims, softmax_scale, window_shapes):
        super().__init__()
        self.embed_dims, self.num_heads = embed_dims, num_heads
        self.head_dims = embed_dims // num_heads
        # The reference accepts any divisor (losses/scalekd.py:273); the attention kernels here cover head dims that are
        # multiples of 8 up to 96 (config.yaml's 16 / 24 heads on every teacher; NOT the reference default of 8 heads on
        # vitl14 / vitg14 = 128 / 192). Refuse at construction, not at the first forward.
        if embed_dims % num_heads == 0 and (self.head_dims % 8 != 0 or self.head_dims > 96):
            raise ValueError(f"head_dim = teacher_dims / num_heads = {self.head_dims}: this library supports multiples of 8 "
                             "up to 96 (use more heads, e.g. 16 or 24 as in the reference's config.yaml)")
        self.softmax_scale = softmax_scale
        self.window_shapes = window_shapes
        self.q = nn.Linear(pos_dims, embed_dims, bias=True)
        self.k = nn.Linear(embed_dims, embed_dims, bias=True)
        self.v = nn.Linear(embed_dims, embed_dims, bias=True)
        self.proj = nn.Linear(embed_dims, embed_dims, bias=True)


class _FFN(nn.Module):
    """Parameters of FFN (losses/scalekd.py:431-462): layers.0.0 = Linear(D,4D) (+ReLU), layers.1 = Linear(4D,D)."""

    def __init__(self, embed_dims, feedforward_channels):
        super().__init__()
        self.layers = nn.Sequential(
            nn.Sequential(nn.Linear(embed_dims, feedforward_channels), nn.ReLU(inplace=True), nn.Dropout(0.0)),
            nn.Linear(feedforward_channels, embed_dims),
        )


_PARAM_ORDER = [
    # (C-ABI field, attribute path)
    ("conv_w", "proj_student.0.weight"), ("conv_b", "proj_student.0.bias"),
    ("bn_w", "proj_student.1.weight"), ("bn_b", "proj_student.1.bias"),
    ("pos_embed", "pos_embed"),
    ("q_w", "pos_attention.q.weight"), ("q_b", "pos_attention.q.bias"),
    ("k_w", "pos_attention.k.weight"), ("k_b", "pos_attention.k.bias"),
    ("v_w", "pos_attention.v.weight"), ("v_b", "pos_attention.v.bias"),
    ("p_w", "pos_attention.proj.weight"), ("p_b", "pos_attention.proj.bias"),
    ("ffn1_w", "ffn.layers.0.0.weight"), ("ffn1_b", "ffn.layers.0.0.bias"),
    ("ffn2_w", "ffn.layers.1.weight"), ("ffn2_b", "ffn.layers.1.bias"),
    ("ln1_w", "norm.weight"), ("ln1_b", "norm.bias"),
    ("ln2_w", "norm_2.weight"), ("ln2_b", "norm_2.bias"),
    ("query_w", "query.weight"),
]


class _ProjectorFn(torch.autograd.Function):
    """AttentionProjector forward/backward as two C calls. Inputs: x, query (or None), then the parameters in
    _PARAM_ORDER (query_w last, possibly None)."""

    @staticmethod
    def forward(ctx, proj, tokens, x, query, *params):
        lib = L.load()
        cfg = proj._cfg(training=proj.training, in_hw=tuple(x.shape[2:]))
        B = x.shape[0]
        x = x.contiguous()
        if query is not None:
            query = query.contiguous()
        pstruct = L.ProjectorParams()
        for (field, _), t in zip(_PARAM_ORDER, params):
            setattr(pstruct, field, None if t is None else t.data_ptr())
        bn = proj.proj_student[1]
        pstruct.bn_running_mean = bn.running_mean.data_ptr()
        pstruct.bn_running_var = bn.running_var.data_ptr()
        # nn.BatchNorm2d's step counter is bumped by the BN kernel itself (one launch less than `+= 1`)
        nbt = bn.num_batches_tracked
        counted = proj.training and nbt is not None and nbt.is_cuda and nbt.dtype == torch.int64
        pstruct.bn_num_batches_tracked = nbt.data_ptr() if counted else None
        out = torch.empty(B, cfg.HW, cfg.D, device=x.device, dtype=torch.float32)
        save = torch.empty(lib.b200_projector_save_bytes(C.byref(cfg), B), dtype=torch.uint8, device=x.device)
        ws_bytes = lib.b200_projector_ws_bytes(C.byref(cfg), B)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        L.check(lib.b200_projector_fwd_tok(C.byref(cfg), C.byref(pstruct), x.data_ptr(),
                                           None if query is None else query.data_ptr(), B, out.data_ptr(),
                                           save.data_ptr(), ws.data_ptr(), ws_bytes,
                                           None if tokens is None else tokens.data_ptr(), _stream()), "projector_fwd")
        ctx.tokens = tokens   # shared student tokens (ScaleKD tokenises preds_S once for both projectors)
        if proj.training and not counted and nbt is not None:
            bn.num_batches_tracked += 1
        ctx.proj, ctx.cfg, ctx.pstruct = proj, cfg, pstruct
        ctx.has_query = query is not None
        ctx.save_for_backward(x, query, save, *[p for p in params if p is not None])
        ctx.param_present = [p is not None for p in params]
        return out

    @staticmethod
    def backward(ctx, dout):
        lib = L.load()
        cfg, pstruct = ctx.cfg, ctx.pstruct
        saved = ctx.saved_tensors
        x, query, save = saved[0], saved[1], saved[2]
        live = list(saved[3:])
        B = x.shape[0]
        dout = dout.contiguous().float()
        params = []
        for present in ctx.param_present:
            params.append(live.pop(0) if present else None)
        # The C side ACCUMULATES every parameter gradient into the buffer it is given. Default: one flat zeroed buffer,
        # returned to autograd. With `accumulate_into_grad` (set by the flat gradient arena / graphed step) the kernels
        # add straight into the existing fp32 `.grad` (a view of the arena) and autograd gets None: that removes one
        # AccumulateGrad add kernel per parameter (21 per projector) from the step.
        direct = [bool(ctx.proj.accumulate_into_grad) and p is not None and p.grad is not None
                  and p.grad.dtype == torch.float32 and p.grad.is_contiguous() and p.grad.is_cuda for p in params]
        sizes = [0 if (p is None or d) else p.numel() for p, d in zip(params, direct)]
        flat = torch.zeros(sum(sizes), device=x.device, dtype=torch.float32) if sum(sizes) else None
        grads, off = [], 0
        gstruct = L.ProjectorGrads()
        for (field, _), p, n, d in zip(_PARAM_ORDER, params, sizes, direct):
            if p is None:
                grads.append(None)
                setattr(gstruct, field, None)
                continue
            if d:
                grads.append(None)
                setattr(gstruct, field, p.grad.data_ptr())
                continue
            gview = flat[off:off + n].view_as(p)
            off += n
            grads.append(gview)
            setattr(gstruct, field, gview.data_ptr())
        need_dx = ctx.needs_input_grad[2]
        need_dq = ctx.has_query and ctx.needs_input_grad[3]
        dx = torch.empty_like(x) if need_dx else None
        dquery = torch.empty_like(query) if need_dq else None
        ws_bytes = lib.b200_projector_ws_bytes(C.byref(cfg), B)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        tokens = ctx.tokens
        L.check(lib.b200_projector_bwd_tok(C.byref(cfg), C.byref(pstruct), C.byref(gstruct), x.data_ptr(),
                                           None if query is None else query.data_ptr(), dout.data_ptr(), B,
                                           None if dx is None else dx.data_ptr(), 0,
                                           None if dquery is None else dquery.data_ptr(), save.data_ptr(),
                                           ws.data_ptr(), ws_bytes, None if tokens is None else tokens.data_ptr(),
                                           _stream()), "projector_bwd")
        return (None, None, dx, dquery, *grads)


class AttentionProjector(nn.Module):
    """losses/scalekd.py:177-245. Constructed in the reference's order so a shared RNG seed gives identical init."""

    def __init__(self, student_dims, teacher_dims, hw_dims, pos_dims, window_shapes=(1, 1), self_query=True,
                 softmax_scale=1., num_heads=8):
        super().__init__()
        # backward kernels add into existing .grad buffers instead of returning gradients (see _ProjectorFn.backward)
        self.accumulate_into_grad = False
        self.hw_dims = tuple(int(v) for v in hw_dims)
        self.student_dims = int(student_dims)
        self.teacher_dims = int(teacher_dims)
        window_shapes = tuple(int(v) for v in window_shapes)
        if int(pos_dims) != self.teacher_dims:
            raise ValueError("pos_dims must equal teacher_dims (train.py:115-116 sets both to teacher.out_dim)")
        self.proj_student = nn.Sequential(nn.Conv2d(self.student_dims, self.teacher_dims, 1, stride=1, padding=0),
                                          nn.BatchNorm2d(self.teacher_dims), nn.ReLU())
        self.pos_embed = nn.Parameter(torch.zeros(1, self.teacher_dims, self.hw_dims[0], self.hw_dims[1]),
                                      requires_grad=True)
        self.pos_attention = _PosAttention(self.teacher_dims, int(num_heads), int(pos_dims), float(softmax_scale),
                                           window_shapes)
        self.ffn = _FFN(self.teacher_dims, self.teacher_dims * 4)
        self.norm = nn.LayerNorm([self.teacher_dims])
        self.norm_2 = nn.LayerNorm([self.teacher_dims])
        self.query = nn.Embedding(self.hw_dims[0] * self.hw_dims[1], self.teacher_dims) if self_query else None
        trunc_normal_(self.pos_embed, std=0.02)
        if self.teacher_dims % int(num_heads) != 0:
            # the reference fails later, inside .reshape (losses/scalekd.py:299); fail at construction instead
            raise ValueError(f"teacher_dims={self.teacher_dims} is not divisible by num_heads={num_heads}")

    def _cfg(self, training: bool, in_hw=None) -> L.ProjectorConfig:
        """`in_hw`: spatial size of the student map handed in. When it differs from the teacher grid the map is taken to
        be the RAW backbone feature (before ModelWrapper's bilinear resize, models/model_zoo.py:121-126) and the resize
        is fused behind the 1x1 conv (include/b200_distill.h: b200_projector_config.raw_h)."""
        bn = self.proj_student[1]
        H, W = self.hw_dims
        raw = (0, 0) if in_hw is None or tuple(int(v) for v in in_hw) == (H, W) else tuple(int(v) for v in in_hw)
        return L.ProjectorConfig(self.student_dims, self.teacher_dims, H * W,
                                 self.pos_attention.num_heads, float(self.pos_attention.softmax_scale), float(bn.eps),
                                 float(bn.momentum if bn.momentum is not None else 0.1), float(self.norm.eps),
                                 int(training), raw[0], raw[1], H, W, *self.pos_attention.window_shapes)

    def _params(self):
        out = []
        for _, path in _PARAM_ORDER:
            obj = self
            ok = True
            for part in path.split("."):
                obj = obj[int(part)] if part.isdigit() else getattr(obj, part, None)
                if obj is None:
                    ok = False
                    break
            out.append(obj if ok else None)
        return out

    def tokenize(self, x: torch.Tensor) -> torch.Tensor:
        """Token-major working copies of the student features (bf16 tokens + 3-term fp16 split), to be shared by the
        two projectors of a ScaleKD via `forward(x, tokens=...)`."""
        lib = L.load()
        cfg = self._cfg(training=self.training, in_hw=tuple(x.shape[2:]))
        B = x.shape[0]
        x = x.contiguous().float()
        tok = torch.empty(lib.b200_projector_tokens_bytes(C.byref(cfg), B), dtype=torch.uint8, device=x.device)
        ws_bytes = lib.b200_projector_ws_bytes(C.byref(cfg), B)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        L.check(lib.b200_projector_tokenize(C.byref(cfg), x.data_ptr(), B, tok.data_ptr(), ws.data_ptr(), ws_bytes,
                                            _stream()), "projector_tokenize")
        return tok

    def forward(self, x, query=None, tokens=None):
        if query is None and self.query is None:
            raise NotImplementedError("There is no query!")
        wh, ww = self.pos_attention.window_shapes
        if wh * ww > 1 and (self.hw_dims[0] != self.hw_dims[1] or self.hw_dims[0] % wh or self.hw_dims[1] % ww):
            # the reference takes the grid as sqrt(N) x sqrt(N) and .view()s it into windows (losses/scalekd.py:326-333)
            raise ValueError(f"window_shapes {wh}x{ww} need a square token grid they divide, got {self.hw_dims}")
        if not x.is_cuda:
            raise L.B200Error("AttentionProjector needs CUDA tensors: there is no CPU fallback")
        H, W = self.hw_dims
        if x.dim() != 4 or x.shape[1] != self.student_dims or x.shape[2] < 1 or x.shape[3] < 1:
            raise ValueError(f"expected student features [B,{self.student_dims},{H},{W}] (or the raw backbone map "
                             f"[B,{self.student_dims},h,w], resized to {H}x{W} inside), got {tuple(x.shape)}")
        return _ProjectorFn.apply(self, tokens, x.float(), None if query is None else query.float(), *self._params())


# ------------------------------------------------------------------------------------------------ loss terms
def _teacher_tokens(preds_T: torch.Tensor) -> Tuple[torch.Tensor, int]:
    """preds_T [B,D,H,W] -> (token-major fp32 tensor [B,Nt,D] whose memory the kernel can read, Nt).

    The teacher shell hands out a strided VIEW of token-major memory (reference: dinov2.py:40); in that case no copy
    is made and the kernel walks the view with row pitch Nt*D. Any other layout is converted by a transpose kernel."""
    B, D, H, W = preds_T.shape
    HW = H * W
    st = preds_T.stride()
    if preds_T.dtype == torch.float32 and st[1] == 1 and st[3] == D and st[2] == W * D and st[0] % D == 0 \
            and st[0] >= HW * D:
        nt = st[0] // D
        tok = torch.as_strided(preds_T, (B, HW, D), (st[0], D, 1))
        return tok, nt
    _, t32 = ops.nchw_to_tokens(preds_T.float().contiguous(), want_f32=True)
    return t32.view(B, HW, D), HW


class _KdLossFn(torch.autograd.Function):
    """(loss, similarity) of one ScaleKD term as TWO 0-dim autograd outputs (views of one 2-float buffer), so neither
    the forward (`out[0]`, `out[1]`) nor the backward (SelectBackward: zeros + copy per output) launches glue kernels:
    the backward kernel reads the two upstream gradients through separate pointers."""

    @staticmethod
    def forward(ctx, S, T_tok, nt, freq, alpha):
        S = S.contiguous()
        B, HW, D = S.shape
        lib = L.load()
        ws = torch.empty(int(lib.b200_kd_loss_ws_floats(B, HW, D)), device=S.device, dtype=torch.float32)
        out = torch.empty(2, device=S.device, dtype=torch.float32)
        L.check(lib.b200_kd_loss_fwd(S.data_ptr(), T_tok.data_ptr(), B, HW, D, nt, 0, int(freq), float(alpha),
                                     out.data_ptr(), ws.data_ptr(), _stream()), "kd_loss_fwd")
        ctx.save_for_backward(S, T_tok, ws)
        ctx.meta = (nt, int(freq), float(alpha))
        ctx.set_materialize_grads(False)
        loss, sim = out.unbind(0)
        return loss, sim

    @staticmethod
    def backward(ctx, g_loss, g_sim):
        S, T_tok, ws = ctx.saved_tensors
        nt, freq, alpha = ctx.meta
        B, HW, D = S.shape
        if g_loss is None and g_sim is None:
            return None, None, None, None, None
        if g_loss is None:
            g_loss = torch.zeros((), device=S.device, dtype=torch.float32)
        g_loss = g_loss.contiguous().float()
        g_sim = None if g_sim is None else g_sim.contiguous().float()
        dS = torch.empty_like(S)
        L.check(L.load().b200_kd_loss_bwd_split(S.data_ptr(), T_tok.data_ptr(), B, HW, D, nt, 0, freq, alpha,
                                                g_loss.data_ptr(), None if g_sim is None else g_sim.data_ptr(),
                                                dS.data_ptr(), 0, ws.data_ptr(), _stream()), "kd_loss_bwd")
        return dS, None, None, None, None


def _kd_term(preds_S: torch.Tensor, preds_T: torch.Tensor, alpha: float, freq: bool):
    if not preds_S.is_cuda:
        raise L.B200Error("ScaleKD loss terms need CUDA tensors: there is no CPU fallback")
    N, Cc, H, W = preds_T.shape
    if freq and H != W:
        raise ValueError("the frequency term needs square maps (losses/scalekd.py:107 builds DCT(resolution=H))")
    if tuple(preds_S.shape) != (N, H * W, Cc):
        raise ValueError(f"expected projected tokens [{N},{H * W},{Cc}], got {tuple(preds_S.shape)}")
    T_tok, nt = _teacher_tokens(preds_T.detach())
    return _KdLossFn.apply(preds_S.float(), T_tok, nt, freq, alpha)


class ScaleKD(nn.Module):
    """losses/scalekd.py:12-127."""

    def __init__(self, name, alpha, student_dims, teacher_dims, query_hw, pos_hw, pos_dims, window_shapes=(1, 1),
                 self_query=True, softmax_scale=[1, 1], num_heads=8):
        super().__init__()
        self.name = name
        self.alpha = [float(a) for a in alpha]
        self.self_query = bool(self_query)
        softmax_scale = [float(s) for s in softmax_scale]
        mk = lambda s: AttentionProjector(student_dims, teacher_dims, query_hw, pos_dims, window_shapes=window_shapes,  # noqa: E731
                                          self_query=self_query, softmax_scale=s, num_heads=num_heads)
        self.projector_0 = mk(softmax_scale[0])
        self.projector_1 = mk(softmax_scale[1])

    def forward(self, preds_S: torch.Tensor, preds_T: torch.Tensor, query_s: Optional[torch.Tensor] = None,
                query_f: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        preds_S_spat = self.project_feat_spat(preds_S, query=query_s)
        preds_S_freq = self.project_feat_freq(preds_S, query=query_f)
        spat_loss, spatial_similarity = self.get_spat_loss(preds_S_spat, preds_T)
        freq_loss, frequency_similarity = self.get_freq_loss(preds_S_freq, preds_T)
        return {"spatial_loss": spat_loss, "frequency_loss": freq_loss, "spatial_similarity": spatial_similarity,
                "frequency_similarity": frequency_similarity, "loss": spat_loss + freq_loss}

    def tokenize_for_both(self, preds_S: torch.Tensor):
        """Shared token-major working copy of preds_S for `projector_0(x, tokens=...)` / `projector_1(x, tokens=...)`
        (None when preds_S needs the projector's own conversions)."""
        return self._tokenize(preds_S)

    def forward_two_streams(self, preds_S, preds_T, query_s, query_f, main, side, teacher_ready=None):
        """`forward` with the spatial branch on `main` and the frequency branch on `side` (see
        distill.DistillationStep.two_streams). Same result; the caller has made `main` current."""
        tok = self._tokenize(preds_S)
        side.wait_stream(main)
        preds_S_spat = self.projector_0(preds_S, query=query_s, tokens=tok)
        with torch.cuda.stream(side):
            preds_S_freq = self.projector_1(preds_S, query=query_f, tokens=tok)
            if teacher_ready is not None:      # preds_T may still be in flight on the teacher's own stream
                side.wait_event(teacher_ready)
            freq_loss, frequency_similarity = self.get_freq_loss(preds_S_freq, preds_T)
        if teacher_ready is not None:
            main.wait_event(teacher_ready)
        spat_loss, spatial_similarity = self.get_spat_loss(preds_S_spat, preds_T)
        main.wait_stream(side)
        for t in (freq_loss, frequency_similarity):   # allocated on `side`, read on `main`
            t.record_stream(main)
        return {"spatial_loss": spat_loss, "frequency_loss": freq_loss, "spatial_similarity": spatial_similarity,
                "frequency_similarity": frequency_similarity, "loss": spat_loss + freq_loss}

    def _tokenize(self, preds_S: torch.Tensor):
        if not (preds_S.is_cuda and preds_S.dtype == torch.float32 and preds_S.is_contiguous() and preds_S.dim() == 4
                and preds_S.shape[1] == self.projector_0.student_dims):
            return None   # the projector's own checks / conversions apply
        return self.projector_0.tokenize(preds_S.detach())

    def project_feat_spat(self, preds_S, query=None):
        """Both projectors read the same preds_S (scalekd.py:27-46, distillation_module.py:230-231): it is tokenised
        HERE, every call, and handed over ONCE to the project_feat_freq call that follows with the same tensor object
        (never across steps: a step's input may be refreshed in place, e.g. the static buffers of a CUDA graph)."""
        tok = self._tokenize(preds_S)
        self._tok_handover = None if tok is None else (weakref.ref(preds_S), preds_S._version, tok)
        return self.projector_0(preds_S, query=query, tokens=tok)

    def project_feat_freq(self, preds_S, query=None):
        h, self._tok_handover = getattr(self, "_tok_handover", None), None
        tok = h[2] if (h is not None and h[0]() is preds_S and h[1] == preds_S._version) else self._tokenize(preds_S)
        return self.projector_1(preds_S, query=query, tokens=tok)

    def get_spat_loss(self, preds_S: torch.Tensor, preds_T: torch.Tensor):
        """alpha[0]/B * sum (S^ - T^)^2 over channel-normalised features, and the mean cosine (scalekd.py:67-92)."""
        return _kd_term(preds_S, preds_T, self.alpha[0], freq=False)

    def get_freq_loss(self, preds_S: torch.Tensor, preds_T: torch.Tensor):
        """Same after DCT -> zero DC -> inverse DCT, weighted alpha[1] (scalekd.py:95-127)."""
        return _kd_term(preds_S, preds_T, self.alpha[1], freq=True)


LOSS_REGISTRY = {"scalekd": ScaleKD}
