#!/usr/bin/env python
"""Hottest SASS instructions (warp-stall samples) of an `ncu --page source --csv --print-source sass` export.
usage: ncu_hot.py source.csv[.gz] [top-n]"""
import csv, gzip, sys
path = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = list(csv.reader(f))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hdr_i]
ci = {k: hdr.index(k) for k in hdr}
samp = ci["# Samples"]
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
tot = sum(int(r[samp] or 0) for r in body)
print(f"total samples {tot}")
idx = sorted(range(len(body)), key=lambda i: -int(body[i][samp] or 0))[:n]
for i in sorted(idx):
    r = body[i]
    st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{i:6d} {int(r[samp]):7d} {100*int(r[samp])/tot:5.1f}%  {r[ci['Source']].strip()[:70]:70s} " + " ".join(f"{nm}:{v}" for v, nm in st if v))
