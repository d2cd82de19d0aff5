#!/usr/bin/env python
"""One step of an `ncu --metrics gpu__time_duration.sum --csv` launch list, in launch order, with grids and durations.
usage: python tools/step_trace.py gpurun_out/launches_X.csv [marker-kernel-substring]
A step starts at each launch of the marker kernel (default: patch_im2col); the LAST complete step is printed."""
import csv
import re
import sys

path = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else "patch_im2col"
lines = [l for l in open(path) if not l.startswith("==")]
rows = []
for row in csv.DictReader(lines):
    try:
        t = float(row["Metric Value"].replace(",", ""))
    except Exception:
        continue
    unit = row.get("Metric Unit", "")
    t = t / 1e3 if unit == "ns" else (t * 1e3 if unit == "ms" else t)
    full = row["Kernel Name"]
    name = re.sub(r"\(.*", "", full.replace("void ", ""))
    name = name.replace("b200::", "").replace("(int)", "").replace("(bool)", "")
    rows.append((name, row["Grid Size"], row["Block Size"], t))
starts = [i for i, r in enumerate(rows) if marker in r[0]]
if len(starts) >= 2:
    a, b = starts[-2], starts[-1]
else:
    a, b = 0, len(rows)
step = rows[a:b]
tot = sum(r[3] for r in step)
print(f"# launches {len(step)}  device time {tot/1e3:.3f} ms (serialised under ncu)")
for i, (n, g, bl, t) in enumerate(step):
    print(f"{i:4d} {t:8.1f} us  {g:>16s} {bl:>14s}  {n[:110]}")
