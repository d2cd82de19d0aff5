"""Teacher-feature cache: keep the frozen teacher's output per sample in HBM and skip its forward when every
sample of a batch has been seen.

The reference recomputes `DINOv2ViT.forward` (models/backbones/dinov2.py:27-46) for every batch of every epoch although
the teacher is frozen (train/distillation_module.py:106-108) -- its output depends on the input image only. That is
forced by the reference's random augmentations (datasets/augmentations.py:24-78: a new crop / flip / RandAugment draw per
epoch); with a deterministic pipeline (validation, no augmentation, or one fixed draw per sample id) the features can be
cached: 256 x 384 bf16 values per vits14 image @224 = 197 KB, so the 180 GB of a B200 hold ~0.9 M images. Also serves
"two-pass" use: several students / loss configurations distilled from one teacher pass.

    cache = TeacherFeatureCache(capacity=100_000, tokens=256, dim=384, device="cuda")
    teacher = CachedTeacher(DINOv2ViT("dinov2_vits14", weights=...).cuda(), cache)
    out = teacher(images, ids=sample_ids)        # {'feature_map': [B, D, H/14, W/14]} like the reference

Host side: a dict sample id -> pool row. Device side: `b200_feature_cache_store/load` (csrc/elementwise.cu)."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional, Sequence

import torch
from torch import nn

from . import _lib as L


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class TeacherFeatureCache:
    """Fixed-capacity pool [capacity, tokens, dim] in bf16 (default; cosine to the fp32 features 1 - 2e-6) or fp32 (exact).
    Rows are handed out first come first served; when the pool is full new samples are simply not cached."""

    def __init__(self, capacity: int, tokens: int, dim: int, device="cuda", dtype: torch.dtype = torch.bfloat16):
        if dtype not in (torch.bfloat16, torch.float32):
            raise ValueError("cache dtype must be torch.bfloat16 or torch.float32")
        if capacity <= 0 or tokens <= 0 or dim <= 0 or dim % 4:
            raise ValueError("capacity, tokens, dim must be positive and dim a multiple of 4")
        self.capacity, self.tokens, self.dim, self.dtype = int(capacity), int(tokens), int(dim), dtype
        self.device = torch.device(device)
        self.pool = torch.empty(self.capacity, self.tokens, self.dim, device=self.device, dtype=dtype) \
            if self.device.type == "cuda" else None   # (CPU: host logic only -- the product path has no CPU fallback)
        self._slot: Dict[int, int] = {}
        self.hits = 0
        self.misses = 0

    def __len__(self) -> int:
        return len(self._slot)

    def __contains__(self, sample_id) -> bool:
        return int(sample_id) in self._slot

    # ---- host logic
    def slots_for(self, ids: Iterable[int], allocate: bool) -> list:
        """Pool row per id (-1: not cached). allocate=True hands out free rows to unseen ids while any are left."""
        out = []
        for i in ids:
            i = int(i)
            s = self._slot.get(i, -1)
            if s < 0 and allocate and len(self._slot) < self.capacity:
                s = len(self._slot)
                self._slot[i] = s
            out.append(s)
        return out

    def all_cached(self, ids: Sequence[int]) -> bool:
        return all(int(i) in self._slot for i in ids)

    # ---- device side
    def _need_pool(self):
        if self.pool is None:
            raise L.B200Error("TeacherFeatureCache: the pool lives on a CUDA device; there is no CPU path")

    def store(self, ids: Sequence[int], tokens: torch.Tensor) -> int:
        """tokens: fp32 [B, tokens, dim] (any batch / token strides, dim contiguous: the teacher's strided view goes in as
        is). Returns how many of the batch were newly cached."""
        self._need_pool()
        if tokens.dim() != 3 or tokens.shape[1:] != (self.tokens, self.dim) or tokens.dtype != torch.float32 or tokens.stride(2) != 1:
            raise ValueError(f"expected fp32 [B, {self.tokens}, {self.dim}] with contiguous last dim, got {tuple(tokens.shape)}")
        fresh = [int(i) not in self._slot for i in ids]
        slots = self.slots_for(ids, allocate=True)
        slots = [s if f else -1 for s, f in zip(slots, fresh)]   # rows already cached are not rewritten
        if all(s < 0 for s in slots):
            return 0
        sl = torch.tensor(slots, dtype=torch.int64, device=self.device)
        L.check(L.load().b200_feature_cache_store(tokens.data_ptr(), tokens.stride(0), tokens.stride(1), sl.data_ptr(),
                                                  self.pool.data_ptr(), int(self.dtype == torch.bfloat16), tokens.shape[0],
                                                  self.tokens, self.dim, _stream()), "feature_cache_store")
        return sum(1 for s in slots if s >= 0)

    def load(self, ids: Sequence[int]) -> Optional[torch.Tensor]:
        """fp32 [B, tokens, dim] when EVERY id is cached, else None (a partially cached batch runs the teacher: one
        forward over the whole batch costs the same as over the missing part at these sizes)."""
        self._need_pool()
        if not self.all_cached(ids):
            self.misses += 1
            return None
        self.hits += 1
        sl = torch.tensor(self.slots_for(ids, allocate=False), dtype=torch.int64, device=self.device)
        out = torch.empty(len(sl), self.tokens, self.dim, device=self.device, dtype=torch.float32)
        L.check(L.load().b200_feature_cache_load(self.pool.data_ptr(), int(self.dtype == torch.bfloat16), sl.data_ptr(),
                                                 out.data_ptr(), len(sl), self.tokens, self.dim, _stream()), "feature_cache_load")
        return out


class CachedTeacher(nn.Module):
    """Wraps a `DINOv2ViT` shell: `forward(x)` is the reference call (models/backbones/dinov2.py:27); `forward(x, ids=...)`
    serves the feature map from the cache when the whole batch is there, otherwise runs the teacher and caches the result.
    `.model` (and so `.model.blocks`, train/distillation_module.py:169-177) is the wrapped teacher's."""

    def __init__(self, teacher: nn.Module, cache: TeacherFeatureCache):
        super().__init__()
        self.teacher = teacher
        self.cache = cache

    @property
    def model(self):
        return self.teacher.model

    def forward(self, x: torch.Tensor, ids: Optional[Sequence[int]] = None):
        gh, gw = x.shape[2] // 14, x.shape[3] // 14
        if ids is not None:
            if len(ids) != x.shape[0]:
                raise ValueError("one sample id per image")
            if gh * gw != self.cache.tokens:
                raise ValueError(f"the cache holds {self.cache.tokens} tokens per image, this resolution has {gh * gw}")
            tok = self.cache.load(ids)
            if tok is not None:
                B, _, D = tok.shape
                return {"feature_map": tok.reshape(B, gh, gw, D).permute(0, 3, 1, 2)}
        out = self.teacher(x)
        if ids is not None:
            fm = out["feature_map"]                      # [B, D, gh, gw] view of token-major memory
            tok = fm.permute(0, 2, 3, 1).flatten(1, 2)   # -> [B, HW, D] (still a view: strides (N*D, D, 1))
            self.cache.store(ids, tok)
        return out
