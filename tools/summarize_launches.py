#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: share of device time per kernel.
usage: python tools/summarize_launches.py gpurun_out/launches_X.csv > profiles/rNN_launches_X.md"""
import collections
import csv
import re
import sys

path = sys.argv[1]
lines = [l for l in open(path) if not l.startswith("==")]
agg = collections.defaultdict(lambda: [0, 0.0])
tot, n = 0.0, 0
for row in csv.DictReader(lines):
    try:
        t = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    unit = row.get("Metric Unit", "")
    t = t / 1e3 if unit == "ns" else (t * 1e3 if unit == "ms" else t)
    name = re.sub(r"<.*", "", row["Kernel Name"]).replace("void ", "")
    name = re.sub(r"\(.*", "", name)
    agg[name][0] += 1
    agg[name][1] += t
    tot += t
    n += 1
print(f"# ncu launch list summary: {path}\n")
print(f"{n} launches, {tot / 1e3:.3f} ms of device time (cold-cache, serialised under ncu: compare SHARES, not absolutes)\n")
print("| share | total us | launches | avg us | kernel |\n|---:|---:|---:|---:|---|")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| {t / tot * 100:.2f}% | {t:.1f} | {c} | {t / c:.1f} | `{k}` |")
