#!/usr/bin/env python
"""HBM-bound kernels from an `ncu --set full --page raw --csv` export: achieved DRAM GB/s per launch (dram read + write
bytes / gpu__time_duration) against the measured copy bandwidth in MEASURED_PEAKS.json.
usage: python tools/ncu_hbm_summary.py raw.csv [peak_GBps] > profiles/rNN_hbm_kernels_ncu.md"""
import csv
import json
import os
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else None
if peak is None:
    try:
        mp = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        peak = float(mp.get("hbm_gbps_burst") or mp.get("hbm_gbps") or 6545.0)
    except Exception:
        peak = 6545.0


def col(name):
    return hdr.index(name)


def to_bytes(v, u):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def to_us(v, u):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}[u]


ci = {k: col(k) for k in ("Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum",
                          "dram__bytes_write.sum", "launch__registers_per_thread",
                          "sm__warps_active.avg.pct_of_peak_sustained_active")}
agg = {}
for r in rows[2:]:
    name = re.sub(r"\(.*", "", r[ci["Kernel Name"]].replace("void ", "")).replace("b200::", "")
    t = to_us(r[ci["gpu__time_duration.sum"]], units[ci["gpu__time_duration.sum"]])
    rd = to_bytes(r[ci["dram__bytes_read.sum"]], units[ci["dram__bytes_read.sum"]])
    wr = to_bytes(r[ci["dram__bytes_write.sum"]], units[ci["dram__bytes_write.sum"]])
    key = (name, r[ci["Grid Size"]], r[ci["Block Size"]])
    agg.setdefault(key, []).append((t, rd, wr, r[ci["launch__registers_per_thread"]],
                                    r[ci["sm__warps_active.avg.pct_of_peak_sustained_active"]]))
print(f"Peak for the fraction: {peak:.0f} GB/s (MEASURED_PEAKS.json copy bandwidth). One row per (kernel, grid): median launch.\n")
print("| kernel | grid | block | launches | time [us] | dram read [MB] | dram write [MB] | GB/s | % of peak | regs | occupancy % |")
print("|---|---|---|---:|---:|---:|---:|---:|---:|---:|---:|")
for (name, grid, block), v in sorted(agg.items(), key=lambda kv: -sorted(x[1] + x[2] for x in kv[1])[len(kv[1]) // 2]):
    v.sort(key=lambda x: x[0])
    t, rd, wr, regs, occ = v[len(v) // 2]
    gbps = (rd + wr) / t * 1e-3
    print(f"| `{name}` | {grid} | {block} | {len(v)} | {t:.1f} | {rd / 1e6:.1f} | {wr / 1e6:.1f} | {gbps:.0f} | {gbps / peak * 100:.0f} | {regs} | {float(occ):.0f} |")
