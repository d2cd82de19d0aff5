"""GPU input pipeline: `DataAugmentationDINO` (datasets/augmentations.py:24-78) without RandAugment, as one
fused kernel over the output batch.

    aug = GpuAugment(global_crops_scale=(0.32, 1.0), global_crops_size=224)      # the reference's constructor arguments
    batch = aug(list_of_uint8_hwc_images)                                        # -> fp32 [B, 3, 224, 224] on the GPU

The reference transforms one PIL image at a time on DataLoader workers (datasets/CustomDataset.py:156-182) and its
published run is input bound (run.ipynb: 1.24 it/s). Here the host only decodes and draws the random parameters --
with torchvision's own `get_params`, in the reference's order, so that the distributions (and, under the same seed, the
draws) are the reference's -- and the pixel work (antialiased bicubic resize of the crop, flip, ToTensor, Normalize,
RandomErasing) runs in `b200_augment_batch` (csrc/augment.cu). RandAugment (augmentations.py:52-58) is NOT covered."""
from __future__ import annotations

import ctypes as C
import math
from typing import List, Optional, Sequence

import torch

from . import _lib as L

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)   # datasets/augmentations.py:12-13
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


def sample_params(sizes: Sequence[Sequence[int]], out_size: int, scale: Sequence[float],
                  ratio: Sequence[float] = (3.0 / 4.0, 4.0 / 3.0), flip_p: float = 0.5, erase_p: float = 0.25,
                  erase_scale: Sequence[float] = (0.02, 1.0 / 3.0), erase_ratio: Sequence[float] = (0.3, 3.3)):
    """Per-image random parameters in the order the reference pipeline draws them from the global torch RNG:
    RandomResizedCrop.get_params, RandomHorizontalFlip's torch.rand(1), RandomErasing's torch.rand(1) and get_params.
    sizes: (H, W) per image. Returns int32 CPU tensors crop [B,4], flip [B], erase [B,4] (erase height 0 = none)."""
    from torchvision import transforms as T
    crop, flip, erase = [], [], []
    for (h, w) in sizes:
        img = torch.empty(3, int(h), int(w), device="meta")
        crop.append(list(T.RandomResizedCrop.get_params(img, list(scale), list(ratio))))
        flip.append(int(torch.rand(1).item() < flip_p))
        e = [0, 0, 0, 0]
        if torch.rand(1).item() < erase_p:
            out = torch.empty(3, out_size, out_size, device="meta")
            i, j, eh, ew, _ = T.RandomErasing.get_params(out, scale=tuple(erase_scale), ratio=tuple(erase_ratio), value=[0.0])
            if not (eh == out_size and ew == out_size):   # (get_params' "return the original image" fallback)
                e = [i, j, eh, ew]
        erase.append(e)
    return (torch.tensor(crop, dtype=torch.int32), torch.tensor(flip, dtype=torch.int32),
            torch.tensor(erase, dtype=torch.int32))


def augment_batch(images: List[torch.Tensor], crop: torch.Tensor, flip: torch.Tensor, erase: torch.Tensor, out_size: int,
                  mean: Sequence[float] = IMAGENET_DEFAULT_MEAN, std: Sequence[float] = IMAGENET_DEFAULT_STD,
                  device="cuda") -> torch.Tensor:
    """images: uint8 [H, W, 3] tensors (CPU or CUDA; any sizes). Applies the given parameters on the GPU."""
    dev = torch.device(device)
    if dev.type != "cuda":
        raise L.B200Error("augment_batch: CUDA only (there is no CPU fallback)")
    B = len(images)
    if B == 0 or crop.shape != (B, 4) or flip.shape != (B,) or erase.shape != (B, 4):
        raise ValueError("one crop / flip / erase row per image")
    sizes, offs, total = [], [], 0
    for im in images:
        if im.dtype != torch.uint8 or im.dim() != 3 or im.shape[2] != 3:
            raise ValueError("images must be uint8 [H, W, 3]")
        sizes.append((im.shape[0], im.shape[1]))
        offs.append(total)
        total += im.numel()
    for (h, w), c in zip(sizes, crop.tolist()):
        if not (0 <= c[0] and 0 <= c[1] and c[2] > 0 and c[3] > 0 and c[0] + c[2] <= h and c[1] + c[3] <= w):
            raise ValueError(f"crop box {c} outside a {h} x {w} image")
    if all(im.device.type == "cpu" for im in images):
        packed = torch.empty(total, dtype=torch.uint8, pin_memory=True)
        for im, o in zip(images, offs):
            packed[o:o + im.numel()] = im.reshape(-1)
        packed = packed.to(dev, non_blocking=True)
    else:
        packed = torch.cat([im.to(dev).reshape(-1) for im in images])
    meta = lambda t, dt=torch.int32: torch.as_tensor(t, dtype=dt).to(dev, non_blocking=True)
    d_off, d_hw = meta(offs, torch.int64), meta(sizes)
    d_crop, d_flip, d_erase = meta(crop), meta(flip), meta(erase)
    out = torch.empty(B, 3, out_size, out_size, device=dev, dtype=torch.float32)
    m3, s3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    # Pillow's ksize for the largest down-scale of the batch
    big = max(max(c[2], c[3]) for c in crop.tolist())
    max_taps = 2 * math.ceil(2.0 * max(big / out_size, 1.0)) + 1
    lib = L.load()
    ws_bytes = lib.b200_augment_ws_bytes(B, out_size, max_taps)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        L.check(lib.b200_augment_batch(packed.data_ptr(), d_off.data_ptr(), d_hw.data_ptr(), d_crop.data_ptr(),
                                       d_flip.data_ptr(), d_erase.data_ptr(), out.data_ptr(), B, out_size, max_taps, m3, s3,
                                       ws.data_ptr(), ws_bytes, torch.cuda.current_stream().cuda_stream), "augment_batch")
    return out


class GpuAugment:
    """Constructor arguments of the reference's DataAugmentationDINO (datasets/augmentations.py:25-31)."""

    def __init__(self, global_crops_scale, global_crops_size: int = 224, device="cuda"):
        self.global_crops_scale = tuple(global_crops_scale)
        self.global_crops_size = int(global_crops_size)
        self.device = device

    def __call__(self, images: List[torch.Tensor]) -> torch.Tensor:
        crop, flip, erase = sample_params([(im.shape[0], im.shape[1]) for im in images], self.global_crops_size,
                                          self.global_crops_scale)
        return augment_batch(images, crop, flip, erase, self.global_crops_size, device=self.device)
