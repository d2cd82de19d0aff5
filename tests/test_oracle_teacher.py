"""CPU: the teacher restatement (oracle/dinov2_ref.py) against an independent code base -- the installed HuggingFace
`transformers` Dinov2Model (same arithmetic; `facebookresearch/dinov2` itself is not on disk, so parity with the hub is
UNPINNED by the reference and this is the strongest cross-check available offline)."""
import math

import pytest
import torch

from oracle import dinov2_ref


def _hf_model(cfg, sd, image_size):
    tr = pytest.importorskip("transformers")
    hc = tr.Dinov2Config(hidden_size=cfg.dim, num_hidden_layers=cfg.depth, num_attention_heads=cfg.heads,
                         mlp_ratio=cfg.ffn_hidden // cfg.dim if not cfg.swiglu else 4, image_size=image_size,
                         patch_size=14, use_swiglu_ffn=cfg.swiglu, layer_norm_eps=1e-6, qkv_bias=True,
                         layerscale_value=1.0, hidden_act="gelu")
    m = tr.Dinov2Model(hc).eval()
    D = cfg.dim
    hsd = {}
    hsd["embeddings.cls_token"] = sd["cls_token"]
    hsd["embeddings.mask_token"] = sd["mask_token"]
    hsd["embeddings.position_embeddings"] = sd["pos_embed"]
    hsd["embeddings.patch_embeddings.projection.weight"] = sd["patch_embed.proj.weight"]
    hsd["embeddings.patch_embeddings.projection.bias"] = sd["patch_embed.proj.bias"]
    for i in range(cfg.depth):
        p, h = f"blocks.{i}.", f"encoder.layer.{i}."
        hsd[h + "norm1.weight"], hsd[h + "norm1.bias"] = sd[p + "norm1.weight"], sd[p + "norm1.bias"]
        hsd[h + "norm2.weight"], hsd[h + "norm2.bias"] = sd[p + "norm2.weight"], sd[p + "norm2.bias"]
        w, b = sd[p + "attn.qkv.weight"], sd[p + "attn.qkv.bias"]
        for j, nm in enumerate(("query", "key", "value")):
            hsd[h + f"attention.attention.{nm}.weight"] = w[j * D:(j + 1) * D]
            hsd[h + f"attention.attention.{nm}.bias"] = b[j * D:(j + 1) * D]
        hsd[h + "attention.output.dense.weight"], hsd[h + "attention.output.dense.bias"] = sd[p + "attn.proj.weight"], sd[p + "attn.proj.bias"]
        hsd[h + "layer_scale1.lambda1"], hsd[h + "layer_scale2.lambda1"] = sd[p + "ls1.gamma"], sd[p + "ls2.gamma"]
        if cfg.swiglu:
            hsd[h + "mlp.weights_in.weight"], hsd[h + "mlp.weights_in.bias"] = sd[p + "mlp.w12.weight"], sd[p + "mlp.w12.bias"]
            hsd[h + "mlp.weights_out.weight"], hsd[h + "mlp.weights_out.bias"] = sd[p + "mlp.w3.weight"], sd[p + "mlp.w3.bias"]
        else:
            hsd[h + "mlp.fc1.weight"], hsd[h + "mlp.fc1.bias"] = sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]
            hsd[h + "mlp.fc2.weight"], hsd[h + "mlp.fc2.bias"] = sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"]
    hsd["layernorm.weight"], hsd["layernorm.bias"] = sd["norm.weight"], sd["norm.bias"]
    missing, unexpected = m.load_state_dict(hsd, strict=False)
    assert not unexpected, unexpected
    assert all("mask_token" in k or "position_ids" in k for k in missing), missing
    return m


@pytest.mark.parametrize("swiglu", [False, True])
def test_restatement_matches_hf_dinov2_without_interpolation(swiglu):
    """At the native grid (no pos-embed interpolation) the two implementations must agree to fp32 round-off."""
    grid = 5
    hidden = dinov2_ref._swiglu_hidden(96) if swiglu else 384
    cfg = dinov2_ref.VitCfg(96, 3, 3, hidden, swiglu)
    sd = dinov2_ref.make_state_dict(cfg, seed=5, pos_grid=grid)
    x = torch.randn(2, 3, grid * 14, grid * 14, generator=torch.Generator().manual_seed(0))
    hf = _hf_model(cfg, sd, grid * 14)
    with torch.no_grad():
        ours = dinov2_ref.forward_tokens(sd, cfg, x)
        theirs = hf(pixel_values=x).last_hidden_state
    assert (ours - theirs).abs().max().item() < 2e-4, (ours - theirs).abs().max().item()


def test_pos_embed_interpolation_hub_convention():
    """scale_factor=(n+0.1)/M bicubic (hub) -- identity at the native grid, right shape otherwise, cls untouched."""
    pe = torch.randn(1, 1 + 37 * 37, 8)
    assert dinov2_ref.interpolate_pos_embed(pe, 37, 37) is pe
    out = dinov2_ref.interpolate_pos_embed(pe, 16, 16)
    assert out.shape == (1, 257, 8)
    assert torch.equal(out[:, 0], pe[:, 0])
    # differs from the size= convention (HF) -- the reason the hub convention is restated explicitly
    grid = pe[:, 1:].reshape(1, 37, 37, 8).permute(0, 3, 1, 2)
    hf_style = torch.nn.functional.interpolate(grid, size=(16, 16), mode="bicubic", align_corners=False)
    assert (out[:, 1:].reshape(1, 16, 16, 8).permute(0, 3, 1, 2) - hf_style).abs().max() > 1e-3


def test_teacher_table_and_feature_map_view():
    assert dinov2_ref.TEACHER_CFGS["dinov2_vitg14"].ffn_hidden == 4096
    cfg = dinov2_ref.VitCfg(32, 2, 2, 64)
    sd = dinov2_ref.make_state_dict(cfg, seed=2, pos_grid=3)
    x = torch.randn(2, 3, 42, 42)
    with torch.no_grad():
        fm = dinov2_ref.teacher_feature_map(sd, cfg, x)
    assert fm.shape == (2, 32, 3, 3)
    assert fm.stride() == (10 * 32, 1, 3 * 32, 32)   # strided view of token-major memory (dinov2.py:40)
    n_params = sum(v.numel() for k, v in dinov2_ref.make_state_dict(dinov2_ref.TEACHER_CFGS["dinov2_vits14"]).items())
    assert abs(n_params - 22.06e6) / 22.06e6 < 0.01  # run.ipynb:109 reports 22.1 M
