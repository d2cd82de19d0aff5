"""ORACLE (test infrastructure only -- never imported by the product path).

CPU/fp32 restatement of the teacher arithmetic that the reference obtains from
`torch.hub.load('facebookresearch/dinov2', name)` (reference: models/backbones/dinov2.py:20, un-vendored, branch
`main`, unpinned) and calls through `get_intermediate_layers(x, n=1, return_class_token=True)`
(models/backbones/dinov2.py:32) and `.model.blocks[i](feat)` (train/distillation_module.py:169-177).

PARITY UNPINNED: the hub source and weights are not on disk and the reference holds no tests / golden vectors for this
boundary. This file restates the published algorithm of upstream `dinov2/models/vision_transformer.py`,
`dinov2/layers/{patch_embed,attention,block,mlp,swiglu_ffn,layer_scale}.py` and `dinov2/hub/backbones.py`
(`_make_dinov2_model`: img_size 518, patch 14, init_values 1.0, ffn 'mlp' / 'swiglufused' for vitg14, block_chunks 0,
no registers, interpolate_antialias False, interpolate_offset 0.1). It is cross-checked against the installed
`transformers` Dinov2Model in tests/test_oracle_teacher.py (same arithmetic, independent code base).

State-dict keys are the hub's, so a real `dinov2_vit*14_pretrain.pth` loads unchanged:
  cls_token, pos_embed[1,1370,D], mask_token, patch_embed.proj.{weight,bias},
  blocks.{i}.{norm1,norm2}.{weight,bias}, blocks.{i}.attn.{qkv,proj}.{weight,bias}, blocks.{i}.{ls1,ls2}.gamma,
  blocks.{i}.mlp.{fc1,fc2}.{weight,bias}  (vitg14: blocks.{i}.mlp.{w12,w3}.{weight,bias}), norm.{weight,bias}
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

PATCH = 14
POS_GRID = 37  # 518 / 14
LN_EPS = 1e-6
INTERPOLATE_OFFSET = 0.1


@dataclass(frozen=True)
class VitCfg:
    dim: int
    depth: int
    heads: int
    ffn_hidden: int  # fc1 output (mlp) or gated hidden size (swiglu)
    swiglu: bool = False


def _swiglu_hidden(dim: int) -> int:
    return (int(dim * 4 * 2 / 3) + 7) // 8 * 8


# reference names: train.py:103-108
TEACHER_CFGS: Dict[str, VitCfg] = {
    "dinov2_vits14": VitCfg(384, 12, 6, 1536),
    "dinov2_vitb14": VitCfg(768, 12, 12, 3072),
    "dinov2_vitl14": VitCfg(1024, 24, 16, 4096),
    "dinov2_vitg14": VitCfg(1536, 40, 24, _swiglu_hidden(1536), True),
}


def make_state_dict(cfg: VitCfg, seed: int = 1, pos_grid: int = POS_GRID) -> Dict[str, torch.Tensor]:
    """Seeded synthetic weights (there are no real checkpoints offline). Linear weights ~ N(0, 1/fan_in) so attention
    logits and MLP pre-activations are O(1); LayerScale gamma ~ U(0.1, 1) so both residual branches matter."""
    g = torch.Generator().manual_seed(seed)
    D = cfg.dim

    def lin(o, i):
        return torch.randn(o, i, generator=g) / math.sqrt(i), torch.randn(o, generator=g) * 0.02

    sd: Dict[str, torch.Tensor] = {}
    sd["cls_token"] = torch.randn(1, 1, D, generator=g) * 0.02
    sd["pos_embed"] = torch.randn(1, 1 + pos_grid * pos_grid, D, generator=g) * 0.02
    sd["mask_token"] = torch.zeros(1, D)
    w, b = lin(D, 3 * PATCH * PATCH)
    sd["patch_embed.proj.weight"] = w.view(D, 3, PATCH, PATCH).contiguous()
    sd["patch_embed.proj.bias"] = b
    for i in range(cfg.depth):
        p = f"blocks.{i}."
        for n in ("norm1", "norm2"):
            sd[p + n + ".weight"] = 1.0 + 0.1 * torch.randn(D, generator=g)
            sd[p + n + ".bias"] = 0.02 * torch.randn(D, generator=g)
        sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"] = lin(3 * D, D)
        sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"] = lin(D, D)
        sd[p + "ls1.gamma"] = 0.1 + 0.9 * torch.rand(D, generator=g)
        sd[p + "ls2.gamma"] = 0.1 + 0.9 * torch.rand(D, generator=g)
        if cfg.swiglu:
            sd[p + "mlp.w12.weight"], sd[p + "mlp.w12.bias"] = lin(2 * cfg.ffn_hidden, D)
            sd[p + "mlp.w3.weight"], sd[p + "mlp.w3.bias"] = lin(D, cfg.ffn_hidden)
        else:
            sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"] = lin(cfg.ffn_hidden, D)
            sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"] = lin(D, cfg.ffn_hidden)
    sd["norm.weight"] = 1.0 + 0.1 * torch.randn(D, generator=g)
    sd["norm.bias"] = 0.02 * torch.randn(D, generator=g)
    return sd


def interpolate_pos_embed(pos_embed: torch.Tensor, gh: int, gw: int) -> torch.Tensor:
    """upstream DinoVisionTransformer.interpolate_pos_encoding: bicubic, antialias off, with the hub's
    scale_factor=(n + 0.1)/M convention (NOT size=; the two differ at 224)."""
    n_pos = pos_embed.shape[1] - 1
    if gh * gw == n_pos and gh == gw:
        return pos_embed
    m = int(math.sqrt(n_pos))
    assert m * m == n_pos
    dim = pos_embed.shape[-1]
    pe = pos_embed.float()
    grid = pe[:, 1:].reshape(1, m, m, dim).permute(0, 3, 1, 2)
    sf = (float(gh + INTERPOLATE_OFFSET) / m, float(gw + INTERPOLATE_OFFSET) / m)
    grid = F.interpolate(grid, mode="bicubic", antialias=False, scale_factor=sf)
    assert grid.shape[-2:] == (gh, gw), (grid.shape, gh, gw)
    grid = grid.permute(0, 2, 3, 1).reshape(1, gh * gw, dim)
    return torch.cat([pe[:, :1], grid], dim=1).to(pos_embed.dtype)


def prepare_tokens(sd, x: torch.Tensor) -> torch.Tensor:
    """PatchEmbed (conv 14/14) -> flatten -> prepend cls -> + interpolated pos_embed."""
    B, _, H, W = x.shape
    t = F.conv2d(x, sd["patch_embed.proj.weight"], sd["patch_embed.proj.bias"], stride=PATCH)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat([sd["cls_token"].expand(B, -1, -1), t], dim=1)
    return t + interpolate_pos_embed(sd["pos_embed"], H // PATCH, W // PATCH)


def attention(sd, p: str, x: torch.Tensor, heads: int) -> torch.Tensor:
    B, N, D = x.shape
    hd = D // heads
    qkv = F.linear(x, sd[p + "qkv.weight"], sd[p + "qkv.bias"]).reshape(B, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * hd ** -0.5, qkv[1], qkv[2]
    a = (q @ k.transpose(-2, -1)).softmax(dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, N, D)
    return F.linear(o, sd[p + "proj.weight"], sd[p + "proj.bias"])


def ffn(sd, p: str, x: torch.Tensor, cfg: VitCfg) -> torch.Tensor:
    if cfg.swiglu:
        x1, x2 = F.linear(x, sd[p + "w12.weight"], sd[p + "w12.bias"]).chunk(2, dim=-1)
        return F.linear(F.silu(x1) * x2, sd[p + "w3.weight"], sd[p + "w3.bias"])
    h = F.gelu(F.linear(x, sd[p + "fc1.weight"], sd[p + "fc1.bias"]))
    return F.linear(h, sd[p + "fc2.weight"], sd[p + "fc2.bias"])


def block(sd, i: int, x: torch.Tensor, cfg: VitCfg) -> torch.Tensor:
    """upstream Block.forward in eval mode: x += ls1(attn(norm1 x)); x += ls2(mlp(norm2 x))."""
    p = f"blocks.{i}."
    D = cfg.dim
    h = F.layer_norm(x, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], LN_EPS)
    x = x + sd[p + "ls1.gamma"] * attention(sd, p + "attn.", h, cfg.heads)
    h = F.layer_norm(x, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], LN_EPS)
    return x + sd[p + "ls2.gamma"] * ffn(sd, p + "mlp.", h, cfg)


def forward_tokens(sd, cfg: VitCfg, x: torch.Tensor) -> torch.Tensor:
    """Normed last-layer tokens [B, 1+HW, D] (what get_intermediate_layers(n=1, norm=True) slices)."""
    t = prepare_tokens(sd, x)
    for i in range(cfg.depth):
        t = block(sd, i, t, cfg)
    return F.layer_norm(t, (cfg.dim,), sd["norm.weight"], sd["norm.bias"], LN_EPS)


def get_intermediate_layers(sd, cfg: VitCfg, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    t = forward_tokens(sd, cfg, x)
    return t[:, 1:], t[:, 0]


def teacher_feature_map(sd, cfg: VitCfg, x: torch.Tensor) -> torch.Tensor:
    """models/backbones/dinov2.py:27-46: patch tokens [B,HW,D] -> (strided view) [B, D, H/14, W/14]."""
    patches, _ = get_intermediate_layers(sd, cfg, x)
    B, _, D = patches.shape
    return patches.reshape(B, x.shape[2] // PATCH, x.shape[3] // PATCH, D).permute(0, 3, 1, 2)


class RefBlock(torch.nn.Module):
    """Callable stand-in for hub `model.blocks[i]` (used by the reference's _forward_specific_stage)."""

    def __init__(self, sd, i, cfg):
        super().__init__()
        self.sd, self.i, self.cfg = sd, i, cfg

    def forward(self, x):
        return block(self.sd, self.i, x, self.cfg)


class RefTeacher(torch.nn.Module):
    """Duck-typed like the reference's DINOv2ViT: forward -> {'feature_map'}, .model.blocks."""

    def __init__(self, cfg: VitCfg, sd):
        super().__init__()
        self.cfg, self.sd = cfg, sd
        self.model = torch.nn.Module()
        self.model.blocks = torch.nn.ModuleList([RefBlock(sd, i, cfg) for i in range(cfg.depth)])

    def forward(self, x):
        with torch.no_grad():
            return {"feature_map": teacher_feature_map(self.sd, self.cfg, x)}
