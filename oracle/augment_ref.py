"""ORACLE (test infrastructure only -- never imported by the product path).

CPU restatement of the per-image pixel work of the reference's input pipeline, datasets/augmentations.py:24-78, WITHOUT
RandAugment (:52-58), for given random parameters:
  RandomResizedCrop(size, scale, BICUBIC) (:36-40): torchvision crops the PIL image, then PIL resizes it. The arithmetic is
      Pillow's (third-party; libImaging/Resample.c, the 8-bits-per-channel path; algorithm unchanged since 3.x, pinned here
      against the installed Pillow): per axis precompute_coeffs (scale = in / out, support = 2 * max(scale, 1), Keys cubic
      a = -0.5, weights normalised by their sequential sum in double), normalize_coeffs_8bpc (22-bit fixed point), then the
      horizontal pass into an 8-bit image and the vertical pass on that (clip8((2^21 + sum) >> 22)).
  RandomHorizontalFlip (:41), ToTensor + Normalize (:60-66), RandomErasing(value=0) (:44-49).
Pinned in tests/test_oracle_augment.py against the reference's own torchvision transforms on PIL images, same seed: equal
bit for bit."""
from __future__ import annotations

import math

import numpy as np
import torch

MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)
PRECISION_BITS = 32 - 8 - 2


def _bicubic(x: float) -> float:
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pil_coeffs(in_size: int, out_size: int):
    """precompute_coeffs + normalize_coeffs_8bpc: dense int64 matrix [out_size, in_size]."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ss = 1.0 / filterscale
    K = np.zeros((out_size, in_size), dtype=np.int64)
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size)
        w = [_bicubic((x - center + 0.5) * ss) for x in range(xmin, xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x, v in zip(range(xmin, xmax), w):
            if ww != 0.0:
                v = v / ww
            f = v * (1 << PRECISION_BITS)
            K[xx, x] = int(-0.5 + f) if v < 0 else int(0.5 + f)
    return K


def pil_resize_bicubic(img_u8_hwc: np.ndarray, size: int) -> np.ndarray:
    h, w, _ = img_u8_hwc.shape
    half = 1 << (PRECISION_BITS - 1)
    x = img_u8_hwc.astype(np.int64)
    Kx, Ky = pil_coeffs(w, size), pil_coeffs(h, size)
    tmp = np.clip((np.einsum("hwc,ow->hoc", x, Kx) + half) >> PRECISION_BITS, 0, 255)      # horizontal pass, 8-bit image
    out = np.clip((np.einsum("hoc,ph->poc", tmp, Ky) + half) >> PRECISION_BITS, 0, 255)    # vertical pass
    return out.astype(np.uint8)


def augment_one(img_u8_hwc: torch.Tensor, crop, flip: int, erase, size: int) -> torch.Tensor:
    t, l, h, w = (int(v) for v in crop)
    r = pil_resize_bicubic(img_u8_hwc[t:t + h, l:l + w].numpy(), size)
    x = torch.from_numpy(r).permute(2, 0, 1).to(torch.float32).div(255)            # ToTensor
    if int(flip):
        x = x.flip(-1)
    x = x.sub(torch.tensor(MEAN).view(3, 1, 1)).div(torch.tensor(STD).view(3, 1, 1))   # Normalize
    i, j, eh, ew = (int(v) for v in erase)
    if eh > 0:
        x[:, i:i + eh, j:j + ew] = 0.0
    return x


def augment_batch(images, crop, flip, erase, size: int) -> torch.Tensor:
    return torch.stack([augment_one(im.cpu(), crop[b].tolist(), int(flip[b]), erase[b].tolist(), size)
                        for b, im in enumerate(images)])
