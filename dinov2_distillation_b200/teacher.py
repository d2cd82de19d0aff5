"""Drop-in for the reference's teacher wrapper `DINOv2ViT` (models/backbones/dinov2.py:5-46).

Same constructor (`model_name` in the four names of train.py:103-108), same `forward(x) -> {'feature_map': [B,D,H/14,W/14]}`
(a strided view of token-major memory, exactly like the reference's reshape+permute at dinov2.py:40), and a `.model`
attribute whose `.blocks[i](feat)` is differentiable w.r.t. `feat` (train/distillation_module.py:169-177).
`.model`'s parameters carry the hub state-dict names, so `dinov2_vit*14_pretrain.pth` files load unchanged.

All arithmetic runs in libb200distill.so (sm_100a kernels); PyTorch only owns the memory. There is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import warnings
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib as L

PATCH = 14
PATCH_K = 3 * PATCH * PATCH  # 588
PATCH_KP = 592               # padded to a multiple of 8 (16-byte TMA rows)


def _swiglu_hidden(dim: int) -> int:
    return (int(dim * 4 * 2 / 3) + 7) // 8 * 8


# name -> (dim, depth, heads, ffn hidden, swiglu)      [train.py:103-108; hub vit_small/base/large/giant2]
TEACHER_CONFIGS: Dict[str, dict] = {
    "dinov2_vits14": dict(dim=384, depth=12, heads=6, ffn=1536, swiglu=False),
    "dinov2_vitb14": dict(dim=768, depth=12, heads=12, ffn=3072, swiglu=False),
    "dinov2_vitl14": dict(dim=1024, depth=24, heads=16, ffn=4096, swiglu=False),
    "dinov2_vitg14": dict(dim=1536, depth=40, heads=24, ffn=_swiglu_hidden(1536), swiglu=True),
}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _LayerScale(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.gamma = nn.Parameter(torch.ones(dim))


class _Attn(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, 3 * dim)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _SwiGLU(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.w12 = nn.Linear(dim, 2 * hidden)
        self.w3 = nn.Linear(hidden, dim)


class _PatchEmbed(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, PATCH, PATCH)


class _VitBlockFn(torch.autograd.Function):
    """One teacher block on tokens [B,N,D]; input-gradient only (the teacher is frozen)."""

    @staticmethod
    def forward(ctx, x, block):
        vit = block._vit()
        x = x.contiguous()
        B, N, D = x.shape
        lib = L.load()
        cfg = vit._cfg_struct
        need_grad = bool(ctx.needs_input_grad[0])
        blk = vit._block_struct(block.index, need_grad)
        y = torch.empty_like(x)
        ws_bytes = lib.b200_vit_block_ws_bytes(C.byref(cfg), B, N)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        save = None
        if need_grad:
            save = torch.empty(lib.b200_vit_block_save_bytes(C.byref(cfg), B, N), dtype=torch.uint8, device=x.device)
        L.check(lib.b200_vit_block_fwd(C.byref(cfg), C.byref(blk), x.data_ptr(), y.data_ptr(), B, N,
                                       save.data_ptr() if save is not None else None, ws.data_ptr(), ws_bytes,
                                       _stream()), "vit_block_fwd")
        if need_grad:
            ctx.save_for_backward(x, save)
            ctx.block = block
        return y

    @staticmethod
    def backward(ctx, dy):
        x, save = ctx.saved_tensors
        block = ctx.block
        vit = block._vit()
        lib = L.load()
        cfg = vit._cfg_struct
        blk = vit._block_struct(block.index, True)
        B, N, D = x.shape
        dy = dy.contiguous().float()
        dx = torch.empty_like(x)
        ws_bytes = lib.b200_vit_block_ws_bytes(C.byref(cfg), B, N)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        L.check(lib.b200_vit_block_bwd_input(C.byref(cfg), C.byref(blk), x.data_ptr(), dy.data_ptr(), dx.data_ptr(), B, N,
                                             save.data_ptr(), ws.data_ptr(), ws_bytes, _stream()), "vit_block_bwd_input")
        return dx, None


class TeacherBlock(nn.Module):
    """hub `Block`: x += ls1(attn(norm1 x)); x += ls2(mlp(norm2 x)), LayerNorm eps 1e-6."""

    def __init__(self, dim, hidden, swiglu, index):
        super().__init__()
        self.index = index
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attn(dim)
        self.ls1 = _LayerScale(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _SwiGLU(dim, hidden) if swiglu else _Mlp(dim, hidden)
        self.ls2 = _LayerScale(dim)
        self._owner = None  # set by DinoVisionTransformerB200 (not a submodule reference -> no cycle in state_dict)

    def _vit(self):
        return self._owner()

    def forward(self, x):
        if not x.is_cuda:
            raise L.B200Error("TeacherBlock needs CUDA tensors: there is no CPU fallback")
        return _VitBlockFn.apply(x.float(), self)


class DinoVisionTransformerB200(nn.Module):
    """Parameter container with the hub's names + the engine tables the C ABI consumes."""

    def __init__(self, dim, depth, heads, ffn, swiglu):
        super().__init__()
        import weakref
        self.embed_dim = dim
        self.depth, self.heads, self.ffn_hidden, self.swiglu = depth, heads, ffn, swiglu
        self.patch_size = PATCH
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, 1 + 37 * 37, dim))
        self.mask_token = nn.Parameter(torch.zeros(1, dim))
        self.patch_embed = _PatchEmbed(dim)
        self.blocks = nn.ModuleList([TeacherBlock(dim, ffn, swiglu, i) for i in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        ref = weakref.ref(self)
        for b in self.blocks:
            b._owner = ref
        self._pack = None          # engine tables (built lazily per device)
        self.batch_streams = 1     # forward_tokens: 2 = half-batches on two CUDA streams (measured slower at cfg2: 6.33 vs 6.21 ms)
        self.min_images_per_stream = 8
        self._pos_cache = {}
        self._cfg_struct = L.VitConfig(dim, depth, heads, ffn, int(swiglu), 1e-6)
        self.register_load_state_dict_post_hook(lambda m, k: m._invalidate())

    # -- cache management -------------------------------------------------------------------------
    def _invalidate(self):
        self._pack = None
        self._pos_cache = {}

    def _apply(self, fn, *a, **k):
        self._invalidate()
        return super()._apply(fn, *a, **k)

    def _ensure_pack(self):
        dev = self.cls_token.device
        if dev.type != "cuda":
            raise L.B200Error("the B200 teacher runs on CUDA only: move it with .cuda() (no CPU fallback)")
        if self._pack is not None and self._pack["device"] == dev:
            return self._pack
        from . import ops
        D = self.embed_dim
        with torch.no_grad():
            pw = self.patch_embed.proj.weight.detach().float().reshape(D, PATCH_K)
            pw = F.pad(pw, (0, PATCH_KP - PATCH_K)).contiguous()
            pack = {"device": dev, "keep": [], "patch_w": ops.cast_bf16(pw), "blocks": (L.VitBlock * self.depth)(),
                    "has_T": [False] * self.depth}
            for i, b in enumerate(self.blocks):
                s = pack["blocks"][i]
                fc1 = b.mlp.w12 if self.swiglu else b.mlp.fc1
                fc2 = b.mlp.w3 if self.swiglu else b.mlp.fc2
                f32 = dict(ln1_w=b.norm1.weight, ln1_b=b.norm1.bias, ln2_w=b.norm2.weight, ln2_b=b.norm2.bias,
                           qkv_b=b.attn.qkv.bias, proj_b=b.attn.proj.bias, ls1=b.ls1.gamma, ls2=b.ls2.gamma,
                           fc1_b=fc1.bias, fc2_b=fc2.bias)
                for name, p in f32.items():
                    t = p.detach().float().contiguous()
                    pack["keep"].append(t)
                    setattr(s, name, t.data_ptr())
                for name, p in dict(qkv_w=b.attn.qkv.weight, proj_w=b.attn.proj.weight, fc1_w=fc1.weight,
                                    fc2_w=fc2.weight).items():
                    t = ops.cast_bf16(p.detach().float())
                    pack["keep"].append(t)
                    setattr(s, name, t.data_ptr())
            for name in ("patch_b", "cls", "norm_w", "norm_b"):
                src = {"patch_b": self.patch_embed.proj.bias, "cls": self.cls_token, "norm_w": self.norm.weight,
                       "norm_b": self.norm.bias}[name]
                pack[name] = src.detach().float().contiguous().view(-1)
        self._pack = pack
        return pack

    def _block_struct(self, i: int, need_transposed: bool):
        pack = self._ensure_pack()
        if need_transposed and not pack["has_T"][i]:
            from . import ops
            b = self.blocks[i]
            fc1 = b.mlp.w12 if self.swiglu else b.mlp.fc1
            fc2 = b.mlp.w3 if self.swiglu else b.mlp.fc2
            s = pack["blocks"][i]
            with torch.no_grad():
                tens = dict(
                    qkv_wT=ops.transpose_bf16(b.attn.qkv.weight.detach().float()),
                    proj_wT=ops.transpose_bf16(b.attn.proj.weight.detach().float(), b.ls1.gamma.detach().float().contiguous()),
                    fc1_wT=ops.transpose_bf16(fc1.weight.detach().float()),
                    fc2_wT=ops.transpose_bf16(fc2.weight.detach().float(), b.ls2.gamma.detach().float().contiguous()),
                )
            for name, t in tens.items():
                pack["keep"].append(t)
                setattr(s, name, t.data_ptr())
            pack["has_T"][i] = True
        return pack["blocks"][i]

    def _pos_for_grid(self, gh: int, gw: int) -> torch.Tensor:
        """hub interpolate_pos_encoding (bicubic, scale_factor=(n+0.1)/37, antialias off), computed once per grid on
        the host side of the ABI (not on the hot path: the result is cached)."""
        key = (gh, gw)
        if key in self._pos_cache:
            return self._pos_cache[key]
        pe = self.pos_embed.detach().float()
        n_pos = pe.shape[1] - 1
        D = pe.shape[-1]
        if not (gh * gw == n_pos and gh == gw):
            m = int(math.sqrt(n_pos))
            grid = pe[:, 1:].reshape(1, m, m, D).permute(0, 3, 1, 2)
            sf = (float(gh + 0.1) / m, float(gw + 0.1) / m)
            grid = F.interpolate(grid, mode="bicubic", antialias=False, scale_factor=sf)
            if tuple(grid.shape[-2:]) != (gh, gw):
                raise L.B200Error(f"pos-embed interpolation produced {tuple(grid.shape[-2:])}, wanted {(gh, gw)}")
            pe = torch.cat([pe[:, :1], grid.permute(0, 2, 3, 1).reshape(1, gh * gw, D)], dim=1)
        pos = pe.reshape(-1, D).contiguous()
        self._pos_cache[key] = pos
        return pos

    # -- forward ----------------------------------------------------------------------------------
    @torch.no_grad()
    def forward_tokens(self, x: torch.Tensor) -> torch.Tensor:
        """images [B,3,H,W] -> normed tokens fp32 [B, 1+HW, D] (cls first)."""
        if not x.is_cuda:
            raise L.B200Error("the B200 teacher needs CUDA input: there is no CPU fallback")
        pack = self._ensure_pack()
        x = x.contiguous().float()
        B, Cc, H, W = x.shape
        if Cc != 3 or H % PATCH or W % PATCH:
            raise ValueError(f"expected [B,3,H,W] with H,W multiples of {PATCH}, got {tuple(x.shape)}")
        gh, gw = H // PATCH, W // PATCH
        N = gh * gw + 1
        pos = self._pos_for_grid(gh, gw)
        lib = L.load()
        cfg = self._cfg_struct
        out = torch.empty(B, N, self.embed_dim, device=x.device, dtype=torch.float32)

        def run(x_part, out_part, nb):
            ws_bytes = lib.b200_vit_forward_ws_bytes(C.byref(cfg), nb, H, W)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
            L.check(lib.b200_vit_forward(C.byref(cfg), pack["blocks"], pack["patch_w"].data_ptr(), PATCH_KP,
                                         pack["patch_b"].data_ptr(), pack["cls"].data_ptr(), pos.data_ptr(),
                                         pack["norm_w"].data_ptr(), pack["norm_b"].data_ptr(), x_part.data_ptr(), nb, H, W,
                                         out_part.data_ptr(), ws.data_ptr(), ws_bytes, _stream()), "vit_forward")
            return ws

        if self.batch_streams > 1 and B >= 2 * self.min_images_per_stream:
            # images are independent through the teacher: two half-batches on two streams. Each chain of ~87 kernels has
            # partial last waves and serial non-GEMM kernels; the other half's kernels fill those gaps.
            main = torch.cuda.current_stream(x.device)
            if getattr(self, "_side_stream", None) is None or self._side_stream.device != x.device:
                self._side_stream = torch.cuda.Stream(device=x.device)
            side = self._side_stream
            b0 = B // 2
            side.wait_stream(main)
            keep = [run(x[:b0], out[:b0], b0)]
            with torch.cuda.stream(side):
                keep.append(run(x[b0:], out[b0:], B - b0))
            main.wait_stream(side)
            del keep   # (workspaces were allocated on `main`, which now follows both halves)
        else:
            run(x, out, B)
        return out

    def get_intermediate_layers(self, x, n=1, reshape=False, return_class_token=False, norm=True):
        """The one call pattern the reference uses (dinov2.py:32): n=1, norm=True."""
        if n != 1 or not norm or reshape:
            raise NotImplementedError("only get_intermediate_layers(x, n=1, norm=True) is on the distillation path")
        t = self.forward_tokens(x)
        if return_class_token:
            return ((t[:, 1:], t[:, 0]),)
        return (t[:, 1:],)

    def forward(self, x):
        return self.forward_tokens(x)[:, 0]


def _seeded_init_(model: DinoVisionTransformerB200, seed: int) -> None:
    """Synthetic weights (no checkpoints offline): N(0, 1/fan_in) linears, gamma ~ U(0.1, 1)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith("gamma"):
                p.copy_(0.1 + 0.9 * torch.rand(p.shape, generator=g))
            elif name.endswith("norm1.weight") or name.endswith("norm2.weight") or name == "norm.weight":
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith(".bias") or name in ("cls_token", "pos_embed"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
            elif name == "mask_token":
                p.zero_()
            elif p.dim() >= 2:
                fan_in = p[0].numel()
                p.copy_(torch.randn(p.shape, generator=g) / math.sqrt(fan_in))


class DINOv2ViT(nn.Module):
    """Reference signature: DINOv2ViT(model_name='dinov2_vitg14'). Extra keyword-only knobs are additive."""

    def __init__(self, model_name: str = "dinov2_vitg14", *, weights: Optional[str] = None, seed: int = 1):
        super().__init__()
        if model_name not in TEACHER_CONFIGS:
            raise KeyError(f"unknown teacher {model_name!r}; expected one of {sorted(TEACHER_CONFIGS)}")
        c = TEACHER_CONFIGS[model_name]
        self.model_name = model_name
        self.model = DinoVisionTransformerB200(c["dim"], c["depth"], c["heads"], c["ffn"], c["swiglu"])
        # The reference always loads pretrained hub weights (models/backbones/dinov2.py:20). Here: `weights` (a .pth file
        # or a directory holding <model_name>_pretrain.pth), else $DINOV2_WEIGHTS_DIR. Seeded synthetic weights are an
        # explicit opt-in (weights="synthetic": tests, bench.py, smoke) -- a missing or mistyped path raises instead of
        # silently distilling from a random teacher.
        if weights == "synthetic":
            _seeded_init_(self.model, seed)
        else:
            path = weights or os.environ.get("DINOV2_WEIGHTS_DIR")
            if not path:
                raise FileNotFoundError(
                    f"{model_name}: no pretrained weights given. Pass weights=<file or directory> or set "
                    "DINOV2_WEIGHTS_DIR (hub format: <model_name>_pretrain.pth); pass weights='synthetic' for seeded "
                    "random weights (benchmarks / tests only).")
            if os.path.isdir(path):
                path = os.path.join(path, f"{model_name}_pretrain.pth")
            if not os.path.isfile(path):
                raise FileNotFoundError(f"{model_name}: pretrained weights not found at {path!r}")
            self.model.load_state_dict(torch.load(path, map_location="cpu"), strict=True)
        self.model.eval()
        for p in self.model.parameters():
            p.requires_grad = False
        self.H = None
        self.W = None

    def forward(self, x):
        patch_embeddings, _cls = self.model.get_intermediate_layers(x, n=1, return_class_token=True)[0]
        if self.H is None:  # cached on the first call, like the reference (dinov2.py:34-36)
            self.H = x.shape[2] // PATCH
            self.W = x.shape[3] // PATCH
        self.B, _, self.D = patch_embeddings.shape
        feature_map = patch_embeddings.reshape(self.B, self.H, self.W, self.D).permute(0, 3, 1, 2)
        return {"feature_map": feature_map}
