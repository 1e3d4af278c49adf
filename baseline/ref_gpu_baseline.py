#!/usr/bin/env python
"""SURVEY 2.3 / 8d "kernel to beat on the same box": the UNMODIFIED reference's own DetectionModel on the GPU through
stock PyTorch - eager (ATen -> cuDNN / cuBLAS, ``channels_last``, bf16 or fp16 = the reference's ``half=True``) and
``torch.compile`` (what ``attempt_compile``, utils/torch_utils.py:908, turns on) - followed by the reference's own
``non_max_suppression`` (utils/nms.py:13-166, torchvision nms on the device), same model / weights / batch as bench.py's
arm.  Measurement infrastructure: prints ONE JSON line; spawned by bench.py (``gpu_baseline`` key), never imported by the
package.

    python baseline/ref_gpu_baseline.py --config 2 [--batch 256] [--modes eager,compile] [--dtype bf16] [--steps 10]
"""
import argparse
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--modes", default="eager")
    ap.add_argument("--device", default="cuda:0")
    a = ap.parse_args()

    import torch

    import bench
    import ref_env

    w, name = bench.workload(a.config)
    if a.batch:
        w["batch"] = a.batch
    out = {"impl": "reference-gpu", "workload": name, "batch": w["batch"], "dtype": a.dtype, "modes": {}}
    try:
        ref_env.import_reference()
    except RuntimeError as e:
        out["unavailable"] = str(e)
        print(json.dumps(out))
        return
    from ultralytics.utils import nms as ref_nms

    from fce_yolo_b200.weights import synth_images

    dev = torch.device(a.device)
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[a.dtype]
    model = ref_env.reference_model(w["yaml"], w.get("variant"), w["seed"]).to(dev).to(dt)
    model = model.to(memory_format=torch.channels_last)
    B, S = w["batch"], w["size"]
    x = synth_images(1234, B, S, S).to(dev, dt).contiguous(memory_format=torch.channels_last)
    torch.backends.cudnn.benchmark = True

    def timed(fn, steps, warmup):
        with torch.inference_mode():
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    for mode in a.modes.split(","):
        t0 = time.perf_counter()
        try:
            net = model
            if mode == "compile":
                net = torch.compile(model)  # attempt_compile's default backend (inductor)

            def fwd():
                return net(x)

            def fwd_nms():
                y = net(x)
                y = y[0] if isinstance(y, (list, tuple)) else y
                return ref_nms.non_max_suppression(y, w["conf"], w["iou"], max_det=w["max_det"])

            ms_f = timed(fwd, a.steps, a.warmup)
            out["modes"][mode] = {"forward_ms": round(ms_f, 3), "forward_images_per_s": round(B / ms_f * 1e3, 1),
                                  "setup_s": round(time.perf_counter() - t0 - ms_f * (a.steps + a.warmup) / 1e3, 1)}
            if mode == "eager":
                # The reference's NMS loops over images in Python with several device syncs each (nms.py:91-161) and gives
                # up after 2.0 + 0.05 * batch seconds (nms.py:70, 162-164: "NMS time limit exceeded", remaining images
                # get no detections): one step is enough to see which regime it is in.
                ms_p = timed(fwd_nms, 1, 0)
                out["modes"][mode].update({"predict_ms": round(ms_p, 3), "predict_images_per_s": round(B / ms_p * 1e3, 1),
                                           "nms_time_limit_s": 2.0 + 0.05 * B,
                                           "nms_hit_time_limit": bool(ms_p - ms_f > (2.0 + 0.05 * B) * 1e3 * 0.95)})
        except Exception as e:  # noqa: BLE001 - a failing compile backend must not lose the eager number
            out["modes"][mode] = {"error": f"{type(e).__name__}: {str(e)[:200]}"}
    out["note"] = ("reference DetectionModel (baseline/_ref), fused, synthetic weights, channels_last; forward = model(x); "
                   "predict = forward + the reference's non_max_suppression on the device; CUDA events, "
                   f"{a.steps} steps after {a.warmup} warm-ups")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
