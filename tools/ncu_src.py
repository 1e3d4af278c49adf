#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` dump: per kernel launch, the SASS lines with the
most warp-stall samples and their dominant stall reasons.  usage: ncu_src.py dump.csv [top] [kernel_index]"""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
only = int(sys.argv[3]) if len(sys.argv) > 3 else None
kernels, cur, hdr = [], None, None
for row in csv.reader(open(path)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "rows": []}
        kernels.append(cur)
        hdr = None
        continue
    if row[0] == "Address":
        hdr = row
        continue
    if cur is not None and hdr is not None:
        cur["rows"].append(row)
for ki, k in enumerate(kernels):
    if only is not None and ki != only:
        continue
    idx = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    s_i = idx["# Samples"]
    total = sum(int(r[s_i] or 0) for r in k["rows"])
    print(f"=== kernel {ki}: total samples {total}")
    agg = {c: sum(int(r[idx[c]] or 0) for r in k["rows"]) for c in stall_cols}
    print("   stall totals:", ", ".join(f"{c[6:]}={v}" for c, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v)[:400])
    ranked = sorted(enumerate(k["rows"]), key=lambda ir: -int(ir[1][s_i] or 0))[:top]
    for li, r in sorted(ranked):
        st = sorted(((int(r[idx[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
        print(f"   {li:5d} {int(r[s_i]):7d} {100.0*int(r[s_i])/max(total,1):5.1f}%  {r[1].strip()[:70]:70s} "
              + " ".join(f"{n}={v}" for v, n in st if v))
