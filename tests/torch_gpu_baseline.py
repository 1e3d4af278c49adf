#!/usr/bin/env python
"""The "kernel to beat" of SURVEY 8d: the same graph through stock PyTorch on the SAME GPU - the oracle's functional
torch forward (test infrastructure, oracle/fce_oracle.py: F.conv2d / softmax / max_pool2d / interpolate ...) moved to
the device in bf16 with channels_last tensors, i.e. cuDNN / cuBLAS kernels launched op by op like the reference's eager
predict path, + torchvision's batched NMS when it is installed.  Prints one JSON line; never imported by the package.
Lives under tests/ because it executes the oracle (test infrastructure) - it is a measurement aid, not a test.

    python tests/torch_gpu_baseline.py [--config N] [--batch B] [--steps K] [--dtype bf16|fp16|fp32] [--device cuda:0]

Not part of bench.py's contract: a reference point for profiles/README.md (run by tools/gpu_job.sh)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
from fce_yolo_b200.weights import synth_images  # noqa: E402
from oracle import fce_oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--device", default="cuda:0")
    a = ap.parse_args()
    w, _ = bench.workload(a.config)
    if a.batch:
        w["batch"] = a.batch
    if a.size:
        w["size"] = a.size
    dev = torch.device(a.device)
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[a.dtype]
    cfg, model, sd = bench.build_model(w)
    sd_dev = {k: (v.to(dev, dt) if v.is_floating_point() else v.to(dev)) for k, v in sd.items()}
    sd_dev = {k: (v.contiguous(memory_format=torch.channels_last) if v.ndim == 4 else v) for k, v in sd_dev.items()}
    B, S = w["batch"], w["size"]
    x = synth_images(1234, B, S, S).to(dev, dt).contiguous(memory_format=torch.channels_last)
    try:
        from torchvision.ops import batched_nms
    except Exception:  # torchvision is optional
        batched_nms = None

    def step():
        y, _ = O.forward(cfg, cfg["scale"], sd_dev, x)
        if batched_nms is None:
            return y
        y = y.float()
        out = []
        for yi in y:  # the reference loops over images too (nms.py:91)
            conf, cls = yi[4:].max(0)
            keep = conf > w["conf"]
            b = yi[:4, keep].t()
            xyxy = torch.cat((b[:, :2] - b[:, 2:] / 2, b[:, :2] + b[:, 2:] / 2), 1)
            out.append(batched_nms(xyxy, conf[keep], cls[keep], w["iou"])[: w["max_det"]])
        return out

    sync = torch.cuda.synchronize if dev.type == "cuda" else (lambda: None)
    with torch.inference_mode():
        for _ in range(a.warmup):
            step()
        sync()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            step()
        sync()
    ms = (time.perf_counter() - t0) * 1e3 / a.steps
    print(json.dumps({"impl": "torch-eager", "device": str(dev), "dtype": a.dtype, "metric": "images/sec",
                      "value": round(B / (ms * 1e-3), 2), "ms_per_step": round(ms, 3), "batch": B, "size": S,
                      "nms": "torchvision.batched_nms" if batched_nms else None, "config": a.config,
                      "note": "oracle's functional torch forward on the device (cuDNN / cuBLAS, channels_last), op by op"}))


if __name__ == "__main__":
    main()
