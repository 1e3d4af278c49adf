#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page raw --csv` output: one line per launch with the figures the roofline
discussion needs (duration, DRAM bytes and % of peak, L2 / L1 / SM throughput %, occupancy, registers, issue
utilisation, tensor-pipe utilisation, top warp-stall reasons).  usage: ncu_raw.py raw.csv [--md]"""
import csv
import sys

path = sys.argv[1]
md = "--md" in sys.argv
rows = list(csv.reader(open(path)))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
SCALE = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "second": 1e6, "s": 1e6, "nsecond": 1e-3,
         "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def gs(r, name):
    """value scaled to microseconds / bytes according to the unit row"""
    i = ix.get(name)
    if i is None or r[i] == "":
        return 0.0
    return float(r[i].replace(",", "")) * SCALE.get(units[i], 1.0)



def g(r, name, default=0.0):
    i = ix.get(name)
    if i is None or r[i] == "":
        return default
    try:
        return float(r[i].replace(",", ""))
    except ValueError:
        return default


stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")
              and "not_issued" not in h]
tensor_cols = [h for h in hdr if "pipe_tensor" in h and "pct_of_peak_sustained_active" in h]
out = []
for r in data:
    name = r[ix["Kernel Name"]]
    short = name.split("(")[0].replace("void ", "").replace("fce::", "").replace("unnamed>::", "").replace("<unnamed>::", "")
    dur = gs(r, "gpu__time_duration.sum")
    rd, wr = gs(r, "dram__bytes_read.sum"), gs(r, "dram__bytes_write.sum")
    stalls = sorted(((g(r, c), c[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for c in stall_cols),
                    reverse=True)[:3]
    tens = max([g(r, c) for c in tensor_cols] + [0.0])
    out.append(dict(id=r[ix["ID"]], k=short[:46], grid=r[ix["launch__grid_size"]], blk=r[ix["launch__block_size"]], us=dur,
                    mb=(rd + wr) / 1e6, gbs=(rd + wr) / 1e3 / max(dur, 1e-9),
                    dram=g(r, "dram__throughput.avg.pct_of_peak_sustained_elapsed"),
                    l2=g(r, "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
                    l1=g(r, "l1tex__throughput.avg.pct_of_peak_sustained_elapsed"),
                    sm=g(r, "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                    occ=g(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                    regs=int(g(r, "launch__registers_per_thread")),
                    issue=g(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"), tensor=tens,
                    stalls=" ".join(f"{n}={v:.1f}" for v, n in stalls if v > 0)))
fmt = "{id:>3} {k:46s} {grid:>7} {blk:>4} {us:8.1f}us {mb:8.1f}MB {gbs:7.0f}GB/s dram{dram:5.1f}% L2{l2:5.1f}% L1{l1:5.1f}% sm{sm:5.1f}% occ{occ:5.1f}% r{regs:<3d} issue{issue:5.1f}% tc{tensor:5.1f}% | {stalls}"
if md:
    print("| id | kernel | grid | block | µs | DRAM MB | GB/s | dram % | L2 % | L1 % | SM % | occ % | regs | issue % | tensor % | top stalls |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for o in out:
        print("| {id} | `{k}` | {grid} | {blk} | {us:.1f} | {mb:.1f} | {gbs:.0f} | {dram:.1f} | {l2:.1f} | {l1:.1f} | {sm:.1f} | {occ:.1f} | {regs} | {issue:.1f} | {tensor:.1f} | {stalls} |".format(**o))
else:
    for o in out:
        print(fmt.format(**o))
