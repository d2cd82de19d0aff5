#!/usr/bin/env python
"""bench.py -- distillation hot-path throughput (teacher fwd + ScaleKD fwd/bwd) on N B200s of one node.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's CPU path (oracle port) on the host cores

Prints ONE JSON line (rank 0). A "step" is one pass of the hot path over one synthetic batch:
frozen DINOv2 teacher forward -> ScaleKD projectors / re-used teacher blocks / loss terms forward -> backward to the
student features and the projector parameters (-> one NCCL mean-allreduce of the flat gradient arena when N > 1).
The student network itself is out of scope (stock PyTorch in the reference): its tapped features are synthetic tensors
of the shapes ModelWrapper hands to the losses.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "distill images/sec (teacher fwd+ScaleKD fwd/bwd)"
UNIT = "images/s"

# BASELINE.json configs. per-GPU batch is fixed as N grows (weak scaling, like `batch_size ... #per gpu`, config.yaml:75)
WORKLOADS = {
    "cfg1": dict(desc="dinov2_vits14 + resnet_18, ScaleKD res5 only, B=2 @224", teacher="dinov2_vits14", size=224, batch=2,
                 losses=[("scalekd_res5", 512, 24, True)], student_params=11_180_000),
    "cfg2": dict(desc="dinov2_vits14 -> stdc_2, config.yaml ScaleKD res4+res5, B=64/GPU @224", teacher="dinov2_vits14",
                 size=224, batch=64, losses=[("scalekd_res4", 512, 16, True), ("scalekd_res5", 1024, 24, False)],
                 raw={"res4": 14, "res5": 7}, student_params=9_350_000),
    "cfg3": dict(desc="dinov2_vitb14 -> convnext_tiny, ScaleKD res4+res5, B=32/GPU @224", teacher="dinov2_vitb14",
                 size=224, batch=32, losses=[("scalekd_res4", 384, 16, True), ("scalekd_res5", 768, 24, False)],
                 raw={"res4": 14, "res5": 7}, student_params=27_860_000),
    "cfg4": dict(desc="dinov2_vitl14 -> swin_tiny, ScaleKD res4+res5 (heads 16), B=32/GPU @518", teacher="dinov2_vitl14",
                 size=518, batch=32, losses=[("scalekd_res4", 384, 16, True), ("scalekd_res5", 768, 16, False)],
                 raw={"res4": 33, "res5": 17}, student_params=27_550_000),
    "cfg5": dict(desc="dinov2_vitg14 teacher forward only, B=64/GPU @224", teacher="dinov2_vitg14", size=224, batch=64,
                 losses=[]),
}
TEACHER_DIMS = {"dinov2_vits14": (384, 12, 6, 1536, False), "dinov2_vitb14": (768, 12, 12, 3072, False),
                "dinov2_vitl14": (1024, 24, 16, 4096, False), "dinov2_vitg14": (1536, 40, 24, 4096, True)}


def loss_specs(wl):
    D = TEACHER_DIMS[wl["teacher"]][0]
    g = wl["size"] // 14
    specs = []
    for name, cs, heads, self_query in wl["losses"]:
        specs.append({"type": "scalekd", "weight": 1.0, "kwargs": dict(
            name=name, alpha=[0.08, 0.06], student_dims=cs, teacher_dims=D, query_hw=[g, g], pos_hw=[g, g], pos_dims=D,
            window_shapes=[1, 1], self_query=self_query, softmax_scale=[5.0, 5.0], num_heads=heads)})
    return specs


def algorithmic_gflop_per_image(wl):
    """SURVEY.md section 8(d) formulas (GEMM 2MNK, attention 4 N^2 D fwd / 8 N^2 D bwd; elementwise not counted)."""
    D, L, h, F, swiglu = TEACHER_DIMS[wl["teacher"]]
    hw = (wl["size"] // 14) ** 2
    n = hw + 1

    def block(ntok):
        ffn = 6 * ntok * D * F if swiglu else 4 * ntok * D * F
        return 8 * ntok * D * D + ffn + 4 * ntok * ntok * D

    teacher = 2 * hw * 588 * D + L * block(n)
    proj_f = sum(2 * hw * cs * D + 24 * hw * D * D + 4 * hw * hw * D for _, cs, _, _ in wl["losses"]) * 2
    stage_f = 0
    if any("res4" in nm for nm, *_ in wl["losses"]):
        stage_f = 2 * sum(block(hw) for _ in range(int(L * 0.75), L - 1))
    stage_b = 0
    if stage_f:
        per = (8 * hw * D * D + (6 if swiglu else 4) * hw * D * F) + 8 * hw * hw * D
        stage_b = 2 * per * len(range(int(L * 0.75), L - 1))
    total = teacher + proj_f + stage_f + 2 * proj_f + stage_b
    return dict(teacher=teacher / 1e9, projector_fwd=proj_f / 1e9, stage_fwd=stage_f / 1e9,
                backward=(2 * proj_f + stage_b) / 1e9, total=total / 1e9)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """`nvidia-smi -lms 50` on this rank's GPU, started BEFORE the warm-up (the tool needs up to a second to produce its
    first line -- longer with 8 GPUs in the box) and kept running; `with sampler:` marks a timed window, and the summary
    covers only the lines that arrived inside the marked windows."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index, self.windows, self._t0 = [], None, index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def __enter__(self):
        self._t0 = time.perf_counter()
        return self

    def __exit__(self, *a):
        self.windows.append((self._t0, time.perf_counter()))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, r in self.rows:
            if not any(a - 0.05 <= ts <= b + 0.05 for a, b in self.windows):   # a line reports the 50 ms before it
                continue
            try:
                sm.append(float(r[0])); mx.append(float(r[1])); pw.append(float(r[2]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw)}


# ------------------------------------------------------------------------------------------------ reference arms
def reference_step_factory(wl, batch, device="cpu", seed=0):
    """The reference's own (stock PyTorch) path for this workload: oracle teacher restatement + oracle
    ScaleKD/_compute_losses port (oracle/*.py; /root/reference itself cannot travel to the GPU box). Returns a callable
    running one fwd+bwd step on `device` -- the host CPU for the reference arm / cpu_baseline leg, cuda for the
    gpu_eager_baseline leg (the same modules as stock eager PyTorch on the B200)."""
    from oracle import dinov2_ref, scalekd_ref
    cfg = dinov2_ref.TEACHER_CFGS[wl["teacher"]]
    tsd = {k: v.to(device) for k, v in dinov2_ref.make_state_dict(cfg, seed=1).items()}
    g = wl["size"] // 14
    gen = torch.Generator().manual_seed(seed)
    img = torch.randn(batch, 3, wl["size"], wl["size"], generator=gen).to(device)
    feats, losses = {}, {}
    for i, (name, cs, heads, self_query) in enumerate(wl["losses"]):
        layer = name.split("_")[1]
        feats[layer] = torch.randn(batch, cs, g, g, generator=gen).to(device).requires_grad_(True)
        sd = {k: v.to(device) for k, v in scalekd_ref.make_scalekd_state(cs, cfg.dim, (g, g), self_query, seed=3 + i).items()}
        for v in sd.values():
            if v.is_floating_point():
                v.requires_grad_(True)
        losses[name] = dict(sd=sd, weight=1.0, alpha=[0.08, 0.06], hw=(g, g), num_heads=heads, softmax_scale=[5.0, 5.0])
    blocks = [lambda x, i=i: dinov2_ref.block(tsd, i, x, cfg) for i in range(cfg.depth)]

    def step():
        for f in feats.values():
            f.grad = None
        for l in losses.values():
            for v in l["sd"].values():
                v.grad = None
        with torch.no_grad():
            T = dinov2_ref.teacher_feature_map(tsd, cfg, img)
        if not losses:
            return T.float().mean()
        out = scalekd_ref.compute_losses(losses, feats, T, blocks)
        out["loss"].backward()
        return out["loss"].detach()

    return step


def time_cpu_reference(wl, batch, steps=5, warmup=2, budget_s=30.0):
    """all host threads, `warmup` untimed steps, then best of up to `steps` timed steps
    (perf_counter around teacher fwd -> losses -> backward). Bounded: stops early once `budget_s` of timed work is spent
    (the 518-pixel workloads take tens of seconds per CPU step)."""
    torch.set_num_threads(os.cpu_count() or 1)
    step = reference_step_factory(wl, batch)
    t_first = time.perf_counter()
    float(step())
    t_first = time.perf_counter() - t_first
    done_w = 1
    while done_w < warmup and t_first * (done_w + 1) < budget_s / 2:
        float(step())
        done_w += 1
    times = []
    while len(times) < max(steps, 1) and (not times or sum(times) + times[-1] < budget_s):
        t0 = time.perf_counter()
        float(step())
        times.append(time.perf_counter() - t0)
    best, mean = min(times), sum(times) / len(times)
    return {"ips": batch / best, "s_best": best, "s_mean": mean, "timed_steps": len(times), "warmups": done_w}


def cpu_sample_text(r, b, wl_name):
    return (f"best of {r['timed_steps']} timed steps after {r['warmups']} warm-ups, B={b} images of {wl_name} through the "
            f"oracle port of the reference (teacher restatement + ScaleKD/_compute_losses, torch CPU fp32): "
            f"{r['s_best']:.3f} s/step best, {r['s_mean']:.3f} s mean")


def run_reference_arm(args, wl, rank, world):
    if rank != 0:
        return
    b = min(wl["batch"], args.cpu_batch)
    r = time_cpu_reference(wl, b, steps=min(max(args.steps, 1), 5), warmup=min(max(args.warmup, 1), 2))
    cores = torch.get_num_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": r["ips"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": r["timed_steps"], "warmup": r["warmups"], "ms_per_step": r["s_best"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "bounded_sample": f"B={b} images per step on the host CPU",
                   "timing": "best-of-N step time (BASELINE.md section 4)"},
        "cpu_baseline": {"value": r["ips"], "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": cpu_sample_text(r, b, args.workload) + f", {cores} threads"},
        "e2e": {"value": r["ips"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def time_gpu_eager_baseline(wl, batch, device, steps=10, warmup=3):
    """The bar SURVEY.md section 8(d) / BASELINE.md section 4 set: the SAME reference modules as stock eager PyTorch on
    this B200 (cuBLAS / ATen kernels, autograd), same batch, CUDA-event timed. Three arithmetic settings: fp32 with TF32
    off (the oracle's own setting), fp32 with TF32 on (the reference's, train.py:304) and bf16 autocast."""
    out = {}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        step = reference_step_factory(wl, batch, device=device)
        for name, tf32, ac in (("fp32_tf32_off", False, None), ("fp32_tf32_on", True, None), ("bf16_autocast", True, torch.bfloat16)):
            torch.backends.cuda.matmul.allow_tf32 = tf32
            torch.backends.cudnn.allow_tf32 = tf32

            def run():
                if ac is None:
                    return step()
                with torch.autocast("cuda", dtype=ac):
                    return step()
            try:
                for _ in range(warmup):
                    run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    run()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / steps
                out[name] = {"value": batch / (ms / 1e3), "unit": UNIT, "ms_per_step": ms}
            except torch.cuda.OutOfMemoryError:
                out[name] = {"value": None, "note": "out of memory at this batch"}
                torch.cuda.empty_cache()
        del step
        torch.cuda.empty_cache()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    out["what"] = (f"oracle port of the reference modules (teacher restatement + ScaleKD/_compute_losses) as stock eager "
                   f"PyTorch on cuda, B={batch}, {steps} steps after {warmup} warm-ups, CUDA events; device-resident inputs")
    return out


# ------------------------------------------------------------------------------------------------ B200 arm
def build_gpu_step(wl, device):
    warnings.simplefilter("ignore")
    from dinov2_distillation_b200 import distributed as D
    from dinov2_distillation_b200.distill import DistillationStep
    from dinov2_distillation_b200.teacher import DINOv2ViT
    torch.manual_seed(3)
    teacher = DINOv2ViT(wl["teacher"], weights="synthetic")
    step = DistillationStep(None, teacher, loss_specs(wl)).to(device).train()
    g = wl["size"] // 14
    gen = torch.Generator().manual_seed(0)
    B = wl["batch"]
    host = {"img": torch.randn(B, 3, wl["size"], wl["size"], generator=gen).pin_memory()}
    for name, cs, _, _ in wl["losses"]:
        layer = name.split("_")[1]
        r = wl.get("raw", {}).get(layer, g) if wl.get("raw_student") else g   # --raw-student: the backbone's own tap size
        host[layer] = torch.randn(B, cs, r, r, generator=gen).pin_memory()
    params = [p for p in step.losses.parameters()]
    # data parallel: the arena also carries a student-sized tail, so the ONE all-reduce per step has the size SURVEY.md
    # section 8(e) gives it (student + ScaleKD gradients; the student's own backward is stock PyTorch and out of scope)
    extra = int(wl.get("student_params", 0)) if wl.get("dp_student_tail") else 0
    arena = D.FlatGradArena(params, extra_numel=extra) if params else None
    if arena is not None:
        arena.enable_direct_accumulation(step.losses)   # backward kernels add straight into the arena views
    return step, host, arena


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--cpu-batch", type=int, default=8, help="images per step of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the stock-eager-PyTorch-on-this-GPU baseline leg")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--dense-table", default="", help="write the per-shape table of the dense launches (roofline leg) here")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of one CUDA graph per step")
    ap.add_argument("--no-student-tail", action="store_true",
                    help="N > 1: all-reduce only the ScaleKD gradients (default: plus a student-sized tail, SURVEY 8e)")
    ap.add_argument("--serial-allreduce", action="store_true",
                    help="N > 1: wait for the all-reduce right after the step instead of overlapping it with the next "
                         "step's teacher forward")
    ap.add_argument("--raw-student", action="store_true",
                    help="feed the student's RAW tap maps (e.g. 14x14 / 7x7 at 224) and fuse ModelWrapper's bilinear resize "
                         "into the projector (SURVEY 8 f1) instead of the already-resized maps the metric is quoted on")
    args = ap.parse_args()
    wl = dict(WORKLOADS[args.workload])
    if args.batch:
        wl["batch"] = args.batch
    wl["raw_student"] = bool(args.raw_student)
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference_arm(args, wl, rank, world)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback); use --impl reference")
    from dinov2_distillation_b200 import _lib as L
    from dinov2_distillation_b200 import distributed as D
    import torch.distributed as dist
    rank, world, local = D.init_from_env()
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    lib = L.load()

    wl["dp_student_tail"] = world > 1 and not args.no_student_tail
    overlap = world > 1 and not args.serial_allreduce and not args.no_graph and bool(wl["losses"])
    step, host, arena = build_gpu_step(wl, device)
    dev = {k: v.to(device) for k, v in host.items()}
    layers = [n.split("_")[1] for n, *_ in wl["losses"]]
    feats = {k: dev[k].clone().requires_grad_(True) for k in layers}

    graphed = None
    if not args.no_graph:
        from dinov2_distillation_b200.distill import GraphedDistillStep
        graphed = GraphedDistillStep(step, dev["img"], {k: dev[k] for k in layers}, arena, split_teacher=overlap)

    def hot_path(img, feats):
        """One step with inputs resident in HBM (img / feats=None: reuse the graph's static input buffers)."""
        if overlap:
            # data parallel: the all-reduce of the PREVIOUS step stays in flight under this step's frozen-teacher forward
            # (nothing a step changes feeds it) and is joined before anything reads trainable state or the arena
            graphed.run_teacher(img)
            arena.wait()
            out, _ = graphed.run_losses(feats)
            arena.allreduce_mean(async_op=True)
            return out
        if graphed is not None:
            out, _ = graphed(img, feats)
            if not layers:
                out = out["teacher"]
        else:
            if arena is not None:
                arena.zero()
            for f in feats.values():
                f.grad = None
            T = step.teacher(img)["feature_map"]
            if not layers:
                return T
            out = step._compute_losses({"student": feats, "teacher": T})
            out["loss"].backward()
        if arena is not None and world > 1:
            arena.allreduce_mean()
        return out

    def sync_all():
        if arena is not None:
            arena.wait()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    res_img, res_feats = (None, None) if graphed is not None else (dev["img"], feats)
    clocks = ClockSampler(local).start()       # streaming before the warm-up; windows are marked below
    for _ in range(args.warmup):
        hot_path(res_img, res_feats)
    sync_all()

    # ---- timed region: device-resident inputs, CUDA events, max over ranks
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib.b200_reset_launch_count()
    with clocks:
        sync_all()
        e0.record()
        for _ in range(args.steps):
            hot_path(res_img, res_feats)
        if arena is not None:
            arena.wait()           # the last step's all-reduce ends inside the timed region
        e1.record()
        sync_all()
    launches = int(lib.b200_launch_count())
    if graphed is not None:
        # graph replays do not pass through the library's launch counter: count one eager step instead
        lib.b200_reset_launch_count()
        graphed._run()
        torch.cuda.synchronize()
        launches = int(lib.b200_launch_count()) * args.steps
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = wl["batch"] * world * args.steps / (ms / 1e3)

    # ---- end to end: pinned host inputs -> H2D every step, loss metrics -> D2H every step
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    # Two pinned result buffers: the loss of step i is copied D2H asynchronously and READ on the host after step i+1 has
    # been enqueued (one-step-deferred logging, the way a training loop reads its metrics without draining the GPU);
    # every step's result is still read inside the timed region -- the last one before the closing event.
    metrics_host = [torch.empty(16, dtype=torch.float32).pin_memory() for _ in range(2)]
    d2h = 0
    pending = []          # [(event, host buffer)] of the step whose result has not been read yet
    step_no = [0]

    def read_pending():
        while pending:
            ev, buf = pending.pop(0)
            ev.synchronize()
            float(buf[0])

    def e2e_step(last=False):
        nonlocal d2h
        if graphed is not None:
            # this step's inputs were staged H2D (side stream) while the previous step computed; stage the next
            # step's inputs now so that its copy overlaps this step -- every step still copies its own batch
            if overlap:
                graphed.consume_staged()
                graphed.run_teacher()
                arena.wait()
                out, _ = graphed.run_losses()
                arena.allreduce_mean(async_op=True)
            else:
                out, _ = graphed.run_staged()
            if not layers:
                out = out["teacher"]
            if arena is not None and world > 1 and not overlap:
                arena.allreduce_mean()
            if not last:
                graphed.stage_inputs(host["img"], {k: host[k] for k in layers})
        else:
            img = host["img"].to(device, non_blocking=True)
            f = {k: host[k].to(device, non_blocking=True).requires_grad_(True) for k in layers}
            out = hot_path(img, f)
        if layers:
            vals = torch.stack([v.detach().float().reshape(()) for _, v in sorted(out.items())])
        else:
            vals = out.float().mean().reshape(1)
        buf = metrics_host[step_no[0] & 1]
        step_no[0] += 1
        buf[:vals.numel()].copy_(vals, non_blocking=True)
        d2h = vals.numel() * 4
        ev = torch.cuda.Event()
        ev.record()
        read_pending()                       # the previous step's loss: its copy finished while this step was enqueued
        pending.append((ev, buf))
        if last:
            read_pending()                   # nothing is left unread when the region closes

    e2e_value = None
    if not args.no_e2e:
        if graphed is not None:
            graphed.stage_inputs(host["img"], {k: host[k] for k in layers})
        for _ in range(2):
            e2e_step()
        e2e_step(last=True)
        sync_all()
        with clocks:
            e0.record()
            if graphed is not None:
                graphed.stage_inputs(host["img"], {k: host[k] for k in layers})   # step 0's copy is inside the region
            for i in range(args.steps):
                e2e_step(last=(i == args.steps - 1))
            if arena is not None:
                arena.wait()
            e1.record()
            sync_all()
        t = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_value = wl["batch"] * world * args.steps / (float(t.item()) / 1e3)

    clocks.stop()
    clocks_summary = clocks.summary()

    # ---- roofline leg: per-launch CUDA-event timing of the dense kernels over the same steps (rank 0 reports)
    roofline, extra = None, {}
    if not args.no_roofline:
        import ctypes as C
        # kernels are timed ALONE on one stream: the two-branch overlap of the timed step (DistillationStep.two_streams)
        # would let a neighbour's kernels run between the events that bracket each GEMM
        step.two_streams = False
        ms_one_stream = None
        if graphed is not None:
            from dinov2_distillation_b200.distill import GraphedDistillStep
            g1 = GraphedDistillStep(step, dev["img"], {k: dev[k] for k in layers}, arena)
            for _ in range(3):
                g1()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                g1()
            e1.record()
            torch.cuda.synchronize()
            ms_one_stream = e0.elapsed_time(e1) / 10
        lib.b200_profile_enable(1)
        stat_names = ("stat_attn_tc_fwd", "stat_attn_pp_fwd", "stat_attn_mma_fwd", "stat_attn_tc_bwd", "stat_attn_mma_bwd")
        for nm in stat_names:
            lib.b200_set_option(nm.encode(), 0)
        nprof = max(2, min(args.steps, 5))
        for _ in range(nprof):   # eager launches (events bracket each dense kernel); not part of `value`
            if graphed is not None:
                g1._run()
            else:
                hot_path(dev["img"], feats)
        torch.cuda.synchronize()
        # per-launch records (launch order): per-category sums, and a table by (category, algorithmic flops) on request
        cap = 4096 * nprof
        r_ms, r_fl, r_cat, r_by = (C.c_float * cap)(), (C.c_double * cap)(), (C.c_int * cap)(), (C.c_double * cap)()
        n_rec = lib.b200_profile_read_records(cap, r_ms, r_fl, r_cat, r_by)
        if n_rec < 0:
            L.check(n_rec, "profile_read_records")
        lib.b200_profile_enable(0)
        ms_c, fl_c, n_c = [0.0] * 3, [0.0] * 3, [0] * 3
        table = {}
        for i in range(n_rec):
            c = r_cat[i]
            if 0 <= c < 3:
                ms_c[c] += r_ms[i]
                fl_c[c] += r_fl[i]
                n_c[c] += 1
                e = table.setdefault((c, r_fl[i], r_by[i]), [0, 0.0])
                e[0] += 1
                e[1] += r_ms[i]
        # what the event pair itself adds to each bracketed launch (measured live with a null kernel)
        br, bb = C.c_float(), C.c_float()
        L.check(lib.b200_profile_event_overhead(256, C.byref(br), C.byref(bb), torch.cuda.current_stream().cuda_stream),
                "profile_event_overhead")
        ev_over_us = max(0.0, br.value - bb.value)
        # the roofline that binds each shape: tensor time (algorithmic FLOPs / sustained bf16 peak) against HBM time
        # (algorithmic operand + result bytes / copy bandwidth); what fraction of that bound the launch achieves, net of the
        # event overhead. Aggregate: sum of bounds / sum of net times over the GEMM launches.
        try:
            _pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            _pk = {}
        pk_tf, pk_gb = float(_pk.get("bf16_tflops_sustained", 1400.0)), float(_pk.get("hbm_gbs", 6545.0))
        bound_us_sum, net_us_sum = 0.0, 0.0
        rows_tbl = []
        for (c, fl, by), (cnt, ms) in sorted(table.items(), key=lambda kv: -kv[1][1]):
            us = ms * 1e3 / cnt
            t_tensor, t_hbm = fl / pk_tf * 1e-6, by / pk_gb * 1e-3
            bound = max(t_tensor, t_hbm)
            net = max(us - ev_over_us, 1e-3)
            if c == 0:
                bound_us_sum += bound * cnt
                net_us_sum += net * cnt
            rows_tbl.append((c, fl, by, cnt, us, ms, t_tensor, t_hbm, bound, net))
        frac_of_bound = bound_us_sum / net_us_sum if net_us_sum > 0 else None
        if args.dense_table and rank == 0:
            names = {0: "gemm", 1: "attention fwd", 2: "attention bwd"}
            with open(args.dense_table, "w") as f:
                f.write(f"# dense launches of one {args.workload} step, by (kind, algorithmic GFLOP, algorithmic MB): CUDA events around each "
                        f"launch, eager, one stream, {nprof} steps; event-pair overhead {ev_over_us:.2f} us per launch\n\n"
                        f"Bound = max(FLOP / {pk_tf:.0f} TFLOP/s, bytes / {pk_gb:.0f} GB/s): the roofline that binds the shape "
                        f"(`hbm` where the operand + result traffic takes longer than the math); `of bound` = bound / (measured - event "
                        f"overhead). GEMM launches together: {frac_of_bound * 100 if frac_of_bound else 0:.0f} % of their bounds.\n\n")
                f.write("| kind | GFLOP | MB | launches / step | avg us | TFLOP/s | us / step | tensor us | hbm us | binds | of bound |\n"
                        "|---|---:|---:|---:|---:|---:|---:|---:|---:|---|---:|\n")
                for (c, fl, by, cnt, us, ms, t_tensor, t_hbm, bound, net) in rows_tbl:
                    f.write(f"| {names[c]} | {fl / 1e9:.2f} | {by / 1e6:.1f} | {cnt / nprof:.1f} | {us:.1f} | {fl / us / 1e6:.0f} | "
                            f"{ms * 1e3 / nprof:.0f} | {t_tensor:.1f} | {t_hbm:.1f} | {'hbm' if t_hbm > t_tensor else 'tensor'} | "
                            f"{bound / net * 100:.0f} % |\n")
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
        traffic = None
        try:
            # measured once per change under `ncu --set full` (never inside a timed run); see profiles/
            tp = [os.path.join(ROOT, "profiles", f) for f in ("r02_traffic.json", "r01_traffic.json")]
            traffic = json.load(open([f for f in tp if os.path.exists(f)][0]))["gemm_v2_kernel"]["traffic_bytes_per_launch"]
        except Exception:
            pass
        if n_c[0] > 0 and ms_c[0] > 0:
            ach = fl_c[0] / (ms_c[0] * 1e-3) / 1e12
            net_ms = ms_c[0] - n_c[0] * ev_over_us * 1e-3
            roofline = {"bound": "tensor", "kernel": "gemm_v2_kernel", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                        "frac": ach / peak, "traffic": traffic,
                        "event_overhead_us_per_launch": ev_over_us,
                        "event_overhead_note": "a null kernel bracketed launch by launch like every record here, minus the "
                                               "same kernel back to back (b200_profile_event_overhead, measured in this run); "
                                               "`achieved` / `frac` keep it in (conservative), the *_net fields take it out",
                        "frac_of_binding_roofline_net": frac_of_bound,
                        "frac_of_binding_roofline_note": "per GEMM shape the bound is max(FLOP / sustained bf16 peak, algorithmic operand + "
                                                         "result bytes / copy bandwidth); sum of bounds / sum of (event time - event overhead): "
                                                         "the K = 384 GEMMs with fp32 residual or saved activations are HBM-bound shapes",
                        "achieved_net": fl_c[0] / (net_ms * 1e-3) / 1e12 if net_ms > 0 else None,
                        "frac_net": fl_c[0] / (net_ms * 1e-3) / 1e12 / peak if net_ms > 0 else None, "traffic_unit": "bytes/launch (ncu dram read+write, profiles/r0*_gemm_v2_ncu_full.md)",
                        "peak_source": peak_src,
                        "peak_kind": "sustained (the kernel is timed inside a long step; the burst figure is for a kernel timed alone)",
                        "launches_per_step": n_c[0] / nprof, "avg_launch_us": ms_c[0] * 1e3 / n_c[0],
                        "algorithmic_gflop_per_launch": fl_c[0] / n_c[0] / 1e9,
                        "share_of_step": (ms_c[0] / nprof) / (ms_one_stream or ms_per_step),
                        "share_note": "GEMM device time per step / step time, both with the step on ONE stream (kernels "
                                      "timed alone); the timed `value` overlaps the spatial and frequency branches on two",
                        "ms_per_step_one_stream": ms_one_stream}
        step.two_streams = True
        # which attention kernels the step launched (per step): tcgen05 (tc / pp) vs the mma.sync fallback
        extra["attention_paths_per_step"] = {nm[5:]: lib.b200_get_option(nm.encode()) / nprof for nm in stat_names}
        for i, nm in ((1, "attention_fwd"), (2, "attention_bwd")):
            if n_c[i] > 0 and ms_c[i] > 0:
                extra[nm] = {"tflops": fl_c[i] / (ms_c[i] * 1e-3) / 1e12, "ms_per_step": ms_c[i] / nprof,
                             "launches_per_step": n_c[i] / nprof}

    # ---- stock eager PyTorch on this GPU (rank 0, N=1 only): the same reference modules, same batch -- the real bar
    gpu_eager = None
    if rank == 0 and world == 1 and not args.no_gpu_eager:
        torch.cuda.empty_cache()
        gpu_eager = time_gpu_eager_baseline(wl, wl["batch"], device)

    # ---- CPU baseline (rank 0, N=1 only): bounded sample of the same workload on the host cores
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        b = min(wl["batch"], args.cpu_batch)
        r = time_cpu_reference(wl, b)
        cpu = {"value": r["ips"], "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": cpu_sample_text(r, b, args.workload)}

    if rank == 0:
        gf = algorithmic_gflop_per_image(wl)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "per_gpu_batch": wl["batch"],
                       "global_batch": wl["batch"] * world, "image_size": wl["size"],
                       "student_features": ("synthetic (student network out of scope)" if not wl.get("raw_student") else
                                            f"synthetic RAW backbone maps {wl.get('raw')}, bilinear resize fused into the projector"),
                       "parallelism": f"dp{world}",
                       "allreduce": (None if world == 1 else
                                     {"bytes": int(arena.numel * 4) if arena is not None else 0,
                                      "student_tail_params": int(wl.get("student_params", 0)) if wl.get("dp_student_tail") else 0,
                                      "overlap": ("in flight under the next step's teacher forward (two graphs: teacher / "
                                                  "projectors+losses+backward), joined before the second graph" if overlap
                                                  else "stream-ordered right after the step")}),
                       "precision": "teacher bf16 operands / fp32 accum+residual; projector fwd fp16 operands, bwd bf16",
                       "l2": "per-step working set (activations >> 126 MB L2) ; no explicit flush",
                       "launch": "eager" if graphed is None else "one CUDA graph replay per step",
                       "e2e_read": "loss dict copied D2H every step; the host reads step i's values after step i+1 is enqueued",
                       "algorithmic_gflop_per_image": gf},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks_summary,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "gpu_eager_baseline": gpu_eager,
            "model_tflops": value * gf["total"] / 1e3 / world,
            "kernels": extra,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
