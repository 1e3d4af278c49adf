#!/usr/bin/env python
"""One eager pass of the bench workload's plan bracketed by cudaProfilerStart/Stop, for
    ncu --profile-from-start off [--set full --import-source on | --metrics gpu__time_duration.sum] python tools/profile_step.py
Prints the node order (tag, C-ABI entry point) so that launch ids in the report map back to plan nodes."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
from fce_yolo_b200 import _lib  # noqa: E402
from fce_yolo_b200.predict import Predictor  # noqa: E402
from fce_yolo_b200.weights import synth_images  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=None)
ap.add_argument("--yaml", default=None)
ap.add_argument("--size", type=int, default=None)
ap.add_argument("--nodes", default=None, help="write the node order to this file")
ap.add_argument("--config", type=int, default=None, help="BASELINE.json configs[i] (default: bench.py's workload)")
a = ap.parse_args()
w, _ = bench.workload(a.config)
if a.yaml:
    w["yaml"], w["variant"] = a.yaml, None
if a.batch:
    w["batch"] = a.batch
if a.size:
    w["size"] = a.size
_lib.load(check_device=True)
dev = torch.device("cuda:0")
cfg, model, sd = bench.build_model(w)
B, S = w["batch"], w["size"]
pred = Predictor(model, B, S, precision=w["precision"], device=dev, conf=w["conf"], iou=w["iou"], max_det=w["max_det"],
                 input_u8=True, use_graph=False)
img = (synth_images(1234, B, S, S) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
pred.inp.copy_(img)
for _ in range(3):
    pred.run_device()
torch.cuda.synchronize()
torch.cuda.profiler.start()
pred.run_device()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
if a.nodes:
    with open(a.nodes, "w") as f:
        for i, n in enumerate(pred.ex.plan.nodes):
            f.write(f"{i},{n.tag},{n.fn},{n.flops / 1e9:.3f},{n.bytes / 1e6:.3f}\n")
print("profiled", len(pred.ex.plan.nodes), "plan nodes")
