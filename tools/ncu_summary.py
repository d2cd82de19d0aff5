#!/usr/bin/env python
"""Markdown summary of an `ncu --page raw --csv` export: one row per captured launch with the metrics the roofline
needs (duration, DRAM bytes, tensor-pipe %, DRAM %, registers, grid). usage: ncu_summary.py raw.csv > profiles/x.md"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = [
    ("Kernel Name", "kernel"), ("launch__grid_size", "grid"), ("launch__registers_per_thread", "regs"),
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active", "hmma %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1 %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"),
]
idx = [(hdr.index(k), n) for k, n in want if k in hdr]
print("| " + " | ".join(f"{n} [{units[i]}]" if units[i] else n for i, n in idx) + " |")
print("|" + "---|" * len(idx))
tot_b, n = 0.0, 0
for r in rows[2:]:
    cells = []
    for i, nme in idx:
        v = r[i]
        if nme == "kernel":
            v = v.replace("void ", "").replace("b200::", "").split("(CUtensorMap")[0].replace("(int)", "").replace("(bool)", "")
            v = "`" + v[:60] + "`"
        cells.append(v)
    print("| " + " | ".join(cells) + " |")
