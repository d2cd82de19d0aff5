"""Optimizer step for the ScaleKD parameters on flat arenas.

The reference hands student + loss parameters to `torch.optim.AdamW` (train/distillation_module.py:440-502,
config/config.yaml:25-30) and lets Lightning clip the global gradient norm (train.py:267-268). `ArenaAdamW` does the
same arithmetic for the loss-module parameters in two kernel launches: parameters, gradients (`FlatGradArena`) and both
moments live in contiguous fp32 buffers, the squared gradient norm is reduced on the device, and the clip coefficient is
applied inside the AdamW kernel (the student's share of the global norm can be passed in as a device scalar).
"""
from __future__ import annotations

from typing import Iterable, Optional, Tuple

import torch

from . import _lib as L
from .distributed import FlatGradArena


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class ArenaAdamW:
    """AdamW (decoupled weight decay, no amsgrad) + optional global-norm clipping over a FlatGradArena.

        arena = FlatGradArena(step.losses.parameters())
        opt = ArenaAdamW(arena, lr=1e-3, betas=(0.9, 0.999), weight_decay=0.01, max_grad_norm=1.0)
        ... backward ...; arena.allreduce_mean(); opt.step(extra_sq_norm=student_sq_norm)   # then arena.zero()

    The parameters are re-pointed to views of ONE flat fp32 buffer (values preserved), like the arena does for `.grad`.
    `lr` may be changed between steps (schedulers: `opt.lr = ...`)."""

    def __init__(self, arena: FlatGradArena, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 1e-2, max_grad_norm: Optional[float] = None):
        if not arena.params:
            raise ValueError("empty arena")
        if not arena.buffer.is_cuda:
            raise L.B200Error("ArenaAdamW needs CUDA parameters: there is no CPU fallback")
        self.arena = arena
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.max_grad_norm = None if max_grad_norm is None else float(max_grad_norm)
        self.n = sum(p.numel() for p in arena.params)
        dev = arena.buffer.device
        self.flat_params = torch.empty(self.n, dtype=torch.float32, device=dev)
        off = 0
        for p in arena.params:
            if p.dtype != torch.float32:
                raise ValueError("ArenaAdamW handles fp32 master parameters")
            view = self.flat_params[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            off += p.numel()
        self.exp_avg = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.sq_norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.steps = 0

    def grad_norm(self) -> torch.Tensor:
        """Global L2 norm of the arena's gradients as of the last `step()` (device scalar)."""
        return self.sq_norm.sqrt()

    @torch.no_grad()
    def step(self, extra_sq_norm: Optional[torch.Tensor] = None) -> None:
        lib = L.load()
        self.steps += 1
        self.arena.wait()   # a still-running asynchronous all-reduce of the gradients
        g = self.arena.buffer
        clip = self.max_grad_norm is not None and self.max_grad_norm > 0
        if clip:
            self.sq_norm.zero_()
            L.check(lib.b200_sqnorm_f32(g.data_ptr(), self.n, self.sq_norm.data_ptr(), _stream()), "sqnorm")
        L.check(lib.b200_adamw_step(self.flat_params.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(),
                                    self.exp_avg_sq.data_ptr(), self.n, self.lr, self.betas[0], self.betas[1], self.eps,
                                    self.weight_decay, self.steps, self.sq_norm.data_ptr() if clip else None,
                                    None if extra_sq_norm is None else extra_sq_norm.data_ptr(),
                                    self.max_grad_norm if clip else 0.0, _stream()), "adamw_step")

    def state_dict(self) -> dict:
        return {"steps": self.steps, "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
                "max_grad_norm": self.max_grad_norm}

    def load_state_dict(self, sd: dict) -> None:
        self.steps = int(sd["steps"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        self.lr, self.betas, self.eps = float(sd["lr"]), tuple(sd["betas"]), float(sd["eps"])
        self.weight_decay, self.max_grad_norm = float(sd["weight_decay"]), sd["max_grad_norm"]
