#!/usr/bin/env python
"""Kernel timeline of a few graph replays of the cfg2 step (torch.profiler / CUPTI): per-stream busy time, overlap and
idle gaps. usage: python tools/timeline.py [steps]  -> gpurun_out/timeline.txt"""
import json, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from dinov2_distillation_b200.distill import GraphedDistillStep

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
wl = dict(bench.WORKLOADS["cfg2"])
dev = torch.device("cuda", 0)
step, host, arena = bench.build_gpu_step(wl, dev)
d = {k: v.to(dev) for k, v in host.items()}
layers = [n.split("_")[1] for n, *_ in wl["losses"]]
g = GraphedDistillStep(step, d["img"], {k: d[k] for k in layers}, arena)
for _ in range(5):
    g()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(steps):
        g()
    torch.cuda.synchronize()
os.makedirs("gpurun_out", exist_ok=True)
prof.export_chrome_trace("gpurun_out/trace.json")
ev = json.load(open("gpurun_out/trace.json"))["traceEvents"]
ks = [e for e in ev if e.get("cat") == "kernel"]
ks.sort(key=lambda e: e["ts"])
t0 = ks[0]["ts"]
# last replay only
span = (ks[-1]["ts"] + ks[-1]["dur"] - t0)
per = span / steps
last = [e for e in ks if e["ts"] - t0 >= per * (steps - 1) - 50]
s0 = last[0]["ts"]
with open("gpurun_out/timeline.txt", "w") as f:
    busy = collections.defaultdict(float)
    for e in last:
        busy[e["args"].get("stream")] += e["dur"]
    end = max(e["ts"] + e["dur"] for e in last)
    f.write(f"# last replay: {len(last)} kernels, span {end - s0:.1f} us, busy per stream {dict(busy)}\n")
    # union coverage
    iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in last)
    cov, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
    for a, b in iv[1:]:
        if a > cur_e:
            cov += cur_e - cur_s; cur_s, cur_e = a, b
        else:
            cur_e = max(cur_e, b)
    cov += cur_e - cur_s
    f.write(f"# time with at least one kernel running: {cov:.1f} us; idle {end - s0 - cov:.1f} us\n")
    for e in last:
        f.write(f"{e['ts'] - s0:9.1f} {e['dur']:8.1f} s{e['args'].get('stream')} {e['name'][:90]}\n")
print(open("gpurun_out/timeline.txt").read()[:600])
os.remove("gpurun_out/trace.json")
