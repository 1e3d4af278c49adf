#!/usr/bin/env python
"""Roofline floor of the compiled plan (CPU only): per node, max(algorithmic bytes / HBM peak, flops / TC peak)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fce_yolo_b200.plan import compile_model, DT_SIZE, View
from fce_yolo_b200.tasks import DetectionModel, variant_cfg, yaml_model_load

yaml = sys.argv[1] if len(sys.argv) > 1 else "yolo11s-fce.yaml"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
size = int(sys.argv[3]) if len(sys.argv) > 3 else 640
variant = {5: ("CoordAtt", []), 8: ("CoordAtt", [])} if yaml == "yolo11s-fce.yaml" else None
cfg = variant_cfg(yaml_model_load(yaml), variant)
model = DetectionModel(cfg).fuse().eval()
plan = compile_model(model, B, size, size, "bf16", torch.device("cpu"), input_u8=True, nms=dict(conf=0.25, iou=0.7))
pk = json.load(open("MEASURED_PEAKS.json"))
hbm, tc = pk["hbm_gbs"] * 1e9, pk["bf16_tflops_sustained"] * 1e12
tot_t = tot_b = tot_f = 0.0
rows = []
for n in plan.nodes:
    def vb(v):
        return v.B * v.H * v.W * v.C * DT_SIZE[v.dtype]
    byts = sum(vb(v) for v in n.reads) + sum(vb(v) for v in n.writes if n.fn != "fce_nms")
    t = max(byts / hbm, n.flops / tc)
    rows.append((t, n.tag, n.fn, byts, n.flops))
    tot_t += t; tot_b += byts; tot_f += n.flops
rows.sort(reverse=True)
for t, tag, fn, byts, fl in rows[:25]:
    print(f"{tag:28s} {fn:18s} {t*1e6:7.1f} us  {byts/1e6:8.1f} MB {fl/1e9:8.1f} GF  {'TC' if fl/tc > byts/hbm else 'HBM'}")
print(f"nodes {len(plan.nodes)}  total bytes {tot_b/1e9:.2f} GB  flops {tot_f/1e12:.3f} TF  roofline floor {tot_t*1e3:.3f} ms "
      f"-> {B/tot_t:.0f} img/s   (bytes-only {tot_b/hbm*1e3:.3f} ms, flops-only {tot_f/tc*1e3:.3f} ms)")
