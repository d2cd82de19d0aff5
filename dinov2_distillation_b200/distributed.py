"""Data-parallel plumbing for the distillation step (one process per GPU, torch.distributed).

The path shards by batch: every rank holds the full frozen teacher and its own images, the teacher forward needs no
communication, and the only exchange per step is the mean-allreduce of the trainable gradients (student + ScaleKD) --
what Lightning's DDP does for the reference (config/config.yaml:66, train.py:262), but as ONE collective over a flat
fp32 arena instead of 25 MB buckets, plus ONE packed collective for the logged scalars instead of one `sync_dist`
all-reduce per key (train/distillation_module.py:353-358). The reference's NCCL_P2P_DISABLE=1 (train.py:23) is NOT
inherited: NCCL runs over NVLink 5 / NVSwitch.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> tuple[int, int, int]:
    """(rank, world, local_rank) from torchrun's environment; initialises the default group when world > 1."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        os.environ.pop("NCCL_P2P_DISABLE", None)
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world, local


def shard_batch(global_batch: int, rank: int, world: int) -> range:
    """Contiguous batch split (remainder goes to the first ranks), like DistributedSampler without shuffling."""
    base, rem = divmod(global_batch, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


class FlatGradArena:
    """All trainable gradients in one contiguous fp32 buffer; `.grad` of every parameter is a view into it, so autograd
    (and the projector backward kernels behind it) accumulate straight into the arena and one all-reduce covers them."""

    def __init__(self, params: Iterable[torch.nn.Parameter], extra_numel: int = 0):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params and extra_numel == 0:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device if self.params else torch.device("cpu")
        self.numel = sum(p.numel() for p in self.params) + int(extra_numel)
        self.buffer = torch.zeros(self.numel, dtype=torch.float32, device=dev)
        off = 0
        self.views = []
        for p in self.params:
            v = self.buffer[off:off + p.numel()].view_as(p)
            off += p.numel()
            p.grad = v
            self.views.append(v)
        self.extra = self.buffer[off:]

    @staticmethod
    def enable_direct_accumulation(module: torch.nn.Module, on: bool = True) -> int:
        """Let every sub-module that supports it (ScaleKD's AttentionProjector) add its parameter gradients straight
        into the arena views from inside its backward kernels. Returns the number of modules switched."""
        n = 0
        for m in module.modules():
            if hasattr(m, "accumulate_into_grad"):
                m.accumulate_into_grad = bool(on)
                n += 1
        return n

    def zero(self) -> None:
        self.wait()
        self.buffer.zero_()
        for p, v in zip(self.params, self.views):  # re-attach if an optimizer set grads to None
            if p.grad is None or p.grad.data_ptr() != v.data_ptr():
                p.grad = v

    def allreduce_mean(self, group=None, async_op: bool = False):
        """One collective per step over the whole arena (mean over ranks, like DDP).

        Default: stream-ordered -- when this returns, every later kernel on the CURRENT stream (optimizer step, norm,
        zero) sees the reduced gradients; nothing on the host blocks. async_op=True (NCCL only) leaves the collective
        running on NCCL's stream and returns its work handle; the arena remembers it and `wait()` / `zero()` /
        `ArenaAdamW.step()` make the current stream wait for it, so a discarded handle cannot race."""
        self.wait()
        if not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        if dist.get_backend(group) == "nccl":
            work = dist.all_reduce(self.buffer, op=dist.ReduceOp.AVG, group=group, async_op=True)
            if async_op:
                self._pending = work
                return work
            work.wait()   # NCCL: the current stream waits for the collective (no host block)
            return None
        work = dist.all_reduce(self.buffer, op=dist.ReduceOp.SUM, group=group, async_op=True)
        work.wait()
        self.buffer.div_(dist.get_world_size(group))
        return None

    def wait(self) -> None:
        """Order the current stream after a pending asynchronous all-reduce (no-op otherwise)."""
        work = getattr(self, "_pending", None)
        if work is not None:
            work.wait()
            self._pending = None


def reduce_metrics(metrics: Dict[str, torch.Tensor], group=None) -> Dict[str, torch.Tensor]:
    """Mean of every logged scalar across ranks with ONE collective (keys sorted for a rank-independent order)."""
    keys = sorted(metrics.keys())
    packed = torch.stack([metrics[k].detach().float().reshape(()) for k in keys])
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        packed = packed / dist.get_world_size(group)
    return {k: packed[i] for i, k in enumerate(keys)}
