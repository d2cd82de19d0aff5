"""ORACLE (test infrastructure only -- never imported by the product path).

Restatement of the student-feature resize of `ModelWrapper.forward` (models/model_zoo.py:118-128):
`F.interpolate(feat, size=patch_size, mode='bilinear', align_corners=False)`. The arithmetic lives in PyTorch
(aten upsample_bilinear2d / area_pixel_compute_source_index, torch 2.11 in this image): for output index d of an axis
resized from `n_in` to `n_out`,

    src = max(0, (d + 0.5) * n_in / n_out - 0.5);  i0 = floor(src);  i1 = i0 + (i0 < n_in - 1);  l1 = src - i0

and the output is (1 - l1) * x[i0] + l1 * x[i1], separably over H and W. Written here as two explicit dense matrices so
the B200 kernels (csrc/elementwise.cu: bilinear_tokens_{fwd,bwd}_kernel) and the fused resize->conv1x1 path of the
projector have an independent check.

PINNING: tests/test_oracle_scalekd.py::test_resize_ref_matches_interpolate compares this file with
torch.nn.functional.interpolate -- the very call the reference makes -- on the reference's own size pairs
(7->16, 14->16, 17->37, 33->37, 16->37, 32->37; SURVEY.md section 8 student-tap table).
"""
from __future__ import annotations

import torch


def resize_matrix(n_in: int, n_out: int, dtype=torch.float64) -> torch.Tensor:
    """R [n_out, n_in] with out = R @ in along one axis (bilinear, align_corners=False)."""
    R = torch.zeros(n_out, n_in, dtype=dtype)
    scale = n_in / n_out
    for d in range(n_out):
        src = max(0.0, (d + 0.5) * scale - 0.5)
        i0 = min(int(src), n_in - 1)
        i1 = i0 + (1 if i0 < n_in - 1 else 0)
        l1 = src - i0
        R[d, i0] += 1.0 - l1
        R[d, i1] += l1
    return R


def resize_bilinear(x: torch.Tensor, size) -> torch.Tensor:
    """x [B, C, h, w] -> [B, C, H, W] (models/model_zoo.py:121-126)."""
    H, W = int(size[0]), int(size[1])
    Ry = resize_matrix(x.shape[2], H, dtype=x.dtype).to(x.device)
    Rx = resize_matrix(x.shape[3], W, dtype=x.dtype).to(x.device)
    return torch.einsum("yh,bchw,xw->bcyx", Ry, x, Rx)
