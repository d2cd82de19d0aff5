#!/usr/bin/env python
"""Hottest SASS instructions (warp-stall samples) of an `ncu --page source --csv --print-source sass` export.
usage: ncu_hot.py source.csv[.gz] [top-n] [section-index]   (an export holds one section per captured launch)"""
import csv, gzip, sys
path = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
only = int(sys.argv[3]) if len(sys.argv) > 3 else None
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
rows = list(csv.reader(f))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for k, hdr_i in enumerate(starts):
    if only is not None and k != only:
        continue
    end = starts[k + 1] - 1 if k + 1 < len(starts) else len(rows)
    name = rows[hdr_i - 1][1] if hdr_i > 0 and rows[hdr_i - 1] and rows[hdr_i - 1][0] == "Kernel Name" else "?"
    hdr = rows[hdr_i]
    ci = {h: i for i, h in enumerate(hdr)}
    samp = ci["# Samples"]
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = [r for r in rows[hdr_i + 1:end] if len(r) == len(hdr)]
    tot = sum(int(r[samp] or 0) for r in body)
    print(f"== section {k}: {name[:110]}\n   {len(body)} SASS lines, total samples {tot}")
    agg = {}
    for r in body:
        for c in stall_cols:
            agg[hdr[c][6:]] = agg.get(hdr[c][6:], 0) + int(r[c] or 0)
    print("   stalls: " + " ".join(f"{nm}:{100 * v / max(tot, 1):.1f}%" for nm, v in sorted(agg.items(), key=lambda t: -t[1])[:8]))
    idx = sorted(range(len(body)), key=lambda i: -int(body[i][samp] or 0))[:n]
    for i in sorted(idx):
        r = body[i]
        st = sorted(((int(r[c] or 0), hdr[c][6:]) for c in stall_cols), reverse=True)[:3]
        print(f"{i:6d} {int(r[samp]):7d} {100 * int(r[samp]) / max(tot, 1):5.1f}%  {r[ci['Source']].strip()[:70]:70s} "
              + " ".join(f"{nm}:{v}" for v, nm in st if v))
