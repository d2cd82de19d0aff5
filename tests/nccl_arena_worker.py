"""Worker of tests/test_gpu_multi.py (launched by torch.distributed.run, one process per GPU, backend nccl).

Every rank holds the same ScaleKD replica and its own shard of the batch (data parallel, config/config.yaml:64-66,
train.py:262): after the backward, the flat gradient arena goes through ONE mean all-reduce. Checked here:
  * the reduced arena equals the host-side average of the per-rank gradients (all-gathered before the reduction);
  * the asynchronous form (handle discarded by the caller) is still ordered before ArenaAdamW.step(): the parameter
    update equals torch.optim.AdamW applied to the host-averaged gradient;
  * reduce_metrics gives the mean of the per-rank loss values with one collective.
"""
import os
import sys
import warnings

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
warnings.simplefilter("ignore")

from dinov2_distillation_b200 import distributed as D  # noqa: E402
from dinov2_distillation_b200.optim import ArenaAdamW  # noqa: E402
from dinov2_distillation_b200.scalekd import ScaleKD  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def main():
    rank, world, local = D.init_from_env()
    dev = torch.device("cuda", local)
    torch.manual_seed(3)   # same replica on every rank
    m = ScaleKD(name="scalekd_res5", alpha=[0.08, 0.06], student_dims=256, teacher_dims=384, query_hw=[16, 16],
                pos_hw=[16, 16], pos_dims=384, window_shapes=[1, 1], self_query=True, softmax_scale=[5.0, 5.0],
                num_heads=24).to(dev).train()
    gen = torch.Generator().manual_seed(100 + rank)   # a different shard of the batch on every rank
    B = 4
    S = torch.randn(B, 256, 16, 16, generator=gen).to(dev).requires_grad_(True)
    T = torch.randn(B, 384, 16, 16, generator=gen).to(dev)
    arena = D.FlatGradArena(m.parameters())
    arena.enable_direct_accumulation(m)
    arena.zero()
    out = m(S, T)
    out["loss"].backward()
    local_grad = arena.buffer.clone()
    gathered = [torch.zeros_like(local_grad) for _ in range(world)]
    dist.all_gather(gathered, local_grad)
    host_mean = torch.stack([g.cpu().double() for g in gathered]).mean(0)
    assert rel(gathered[0], gathered[world - 1]) > 1e-3, "ranks must see different data"

    # (1) stream-ordered form
    arena.allreduce_mean()
    e1 = rel(arena.buffer, host_mean)

    # (2) asynchronous form, handle discarded; optimizer step right behind it
    arena.buffer.copy_(local_grad)
    p0 = [p.detach().clone() for p in arena.params]
    opt = ArenaAdamW(arena, lr=1e-2, betas=(0.9, 0.999), weight_decay=0.01, max_grad_norm=1.0)
    arena.allreduce_mean(async_op=True)
    opt.step()
    torch.cuda.synchronize()
    ref_params = [p.clone().cpu().double().requires_grad_(True) for p in p0]
    off = 0
    for p in ref_params:
        p.grad = host_mean[off:off + p.numel()].view_as(p).clone()
        off += p.numel()
    torch.nn.utils.clip_grad_norm_(ref_params, 1.0)
    ref_opt = torch.optim.AdamW(ref_params, lr=1e-2, betas=(0.9, 0.999), weight_decay=0.01)
    ref_opt.step()
    # the update of ALL parameters as one vector (per tensor it is ill conditioned where the gradient is exactly zero --
    # the conv bias under BatchNorm: the update is lr * wd * p = 1e-4 |p|, a few fp32 ulps of p), and every new parameter
    # value against the float64 reference to fp32 rounding of (|p| + lr)
    upd = torch.cat([(p.detach() - q).flatten().cpu().double() for p, q in zip(arena.params, p0)])
    upd_ref = torch.cat([(r.detach() - q.cpu().double()).flatten() for q, r in zip(p0, ref_params)])
    e2 = ((upd - upd_ref).norm() / upd_ref.norm()).item()
    e2b = max(((p.detach().cpu().double() - r.detach()).abs() / (r.detach().abs() + 1e-2)).max().item()
              for p, r in zip(arena.params, ref_params))

    # (3) packed metrics
    met = D.reduce_metrics({"loss": out["loss"].detach(), "rank": torch.tensor(float(rank), device=dev)})
    losses = [torch.zeros(1, device=dev) for _ in range(world)]
    dist.all_gather(losses, out["loss"].detach().reshape(1))
    e3 = abs(met["loss"].item() - torch.stack(losses).mean().item())
    e4 = abs(met["rank"].item() - (world - 1) / 2)

    print(f"rank {rank}/{world}: arena vs host mean {e1:.2e}; AdamW update vs reference {e2:.2e} (worst element {e2b:.2e}); "
          f"metrics {e3:.2e} {e4:.2e}", flush=True)
    ok = e1 < 1e-6 and e2 < 1e-4 and e2b < 1e-5 and e3 < 1e-6 and e4 < 1e-6
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
