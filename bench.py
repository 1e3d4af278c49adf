#!/usr/bin/env python
"""bench.py - images/s of the FCE-YOLOv11 predict step (forward + DFL decode + NMS) on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config i]

Workload at every N (weak scaling, one replica per GPU): the north-star target configuration, BASELINE.json
configs[2] - yolo11m-bifpn.yaml, bf16, batch 256 per GPU, 640x640 synthetic images, synthetic (seeded) weights.  A step
is one pass of the hot path over one batch.  At N=1 the other four BASELINE configs are measured in the same run and
reported under `configs` (their own batch / image size; same code path).

 value        : whole-job images/s with the batch resident in HBM (CUDA-event timing, max over ranks).
 e2e          : same metric through the public Predictor API with HOST buffers: pinned uint8 NHWC images copied H2D,
                detections copied D2H, every step, inside the timed region.
 roofline     : dominant kernel class (the tcgen05 implicit-GEMM conv), from per-launch CUDA-event timing of one more
                eager pass; per-class figures for every other kernel under roofline.classes.
 cpu_baseline : the UNMODIFIED reference (baseline/_ref) - YOLO(cfg).predict(tensor, device="cpu") - on the host cores,
                bounded sample (N=1, rank 0).
 gpu_baseline : the reference's own DetectionModel through stock PyTorch (cuDNN / cuBLAS bf16 channels_last, eager and
                torch.compile) + its own NMS on the same GPU, same batch (N=1, rank 0; baseline/ref_gpu_baseline.py).
 --impl reference : the reference arm - the same CPU measurement as a line of its own.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

COMMON = dict(precision="bf16", conf=0.25, iou=0.7, max_det=300, seed=1)
CONFIGS = {
    0: dict(name="yolo11n-fce predict: forward+DFL decode+NMS, bf16, batch 1, 640x640 (BASELINE configs[0] on the GPU path)",
            yaml="yolo11n-fce.yaml", variant=None, batch=1, size=640),
    1: dict(name="yolo11s-fce (CoordAtt@L5,L8) predict: forward+DFL decode+NMS, bf16, batch 64/GPU, 640x640 "
                 "(BASELINE configs[1])",
            yaml="yolo11s-fce.yaml", variant={5: ("CoordAtt", []), 8: ("CoordAtt", [])}, batch=64, size=640),
    2: dict(name="yolo11m-bifpn predict: forward+DFL decode+NMS, bf16, batch 256/GPU, 640x640 (BASELINE configs[2], the "
                 "north-star target configuration)",
            yaml="yolo11m-bifpn.yaml", variant=None, batch=256, size=640),
    3: dict(name="yolo11s-fce (CoordCrossAtt@L5, BiCoordCrossAtt[8 heads]@L8) predict: forward+DFL decode+NMS, bf16, "
                 "batch 128, 640x640 (BASELINE configs[3])",
            yaml="yolo11s-fce.yaml", variant={5: ("CoordCrossAtt", [512, 16, 2]), 8: ("BiCoordCrossAtt", [512, 8, 8])},
            batch=128, size=640),
    4: dict(name="yolo11x-fce predict: forward+DFL decode+NMS, bf16, batch 32/GPU, 1280x1280 (BASELINE configs[4])",
            yaml="yolo11x-fce.yaml", variant=None, batch=32, size=1280),
}
DEFAULT_CONFIG = 2


def workload(idx=None):
    """(parameter dict, name) of BASELINE.json configs[idx]; idx None = the judged default."""
    w = dict(COMMON)
    w.update(CONFIGS[DEFAULT_CONFIG if idx is None else idx])
    return w, w.pop("name")


WORKLOAD, WORKLOAD_NAME = workload()


def config_obj(name, w, world):
    """The `config` object - identical in both arms (the reference arm reports its bounded sample elsewhere)."""
    return {"workload": name, "global_batch": w["batch"] * world, "image_size": w["size"],
            "parallelism": f"dp{world} (replicas, image-sharded)",
            "l2": "inputs larger than L2 (uint8 batch + >1 GB of activations per step)"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], src="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, src="fallback")


def write_only_peak(device):
    """HBM WRITE-only bandwidth, measured live (GB/s): a fill of 2 GiB.  On these B200s a pure write stream reaches about
    3.9 TB/s where reads and read+write copies reach 6.5 - 6.6 TB/s, so a kernel that mostly writes (the stem: 0.3 GB
    in, 3.4 GB out) is bounded by this number, not by MEASURED_PEAKS.json's copy bandwidth.  Reported next to the contract's
    roofline, never instead of it."""
    x = torch.empty(2 << 30, dtype=torch.uint8, device=device)
    x.zero_()
    torch.cuda.synchronize(device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        x.zero_()
    e1.record()
    torch.cuda.synchronize(device)
    return 5 * (2 << 30) / (e0.elapsed_time(e1) * 1e-3) / 1e9


def build_model(w):
    from fce_yolo_b200.tasks import DetectionModel, variant_cfg, yaml_model_load
    from fce_yolo_b200.weights import load_synthetic

    cfg = variant_cfg(yaml_model_load(w["yaml"]), w.get("variant"))
    model = DetectionModel(cfg).fuse().eval()
    sd = load_synthetic(model, w["seed"])
    return cfg, model, sd


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = max(mx, float(r[2]))
                for n, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------- CPU arms
def cpu_port_run(w, cfg, sd, n_images, steps, warmup, threads):
    """Fallback when baseline/_ref is absent: the oracle port of the reference predict path on the host: forward (torch
    fp32 CPU) + NMS.  Returns (images/s, ms per step)."""
    from fce_yolo_b200.weights import synth_images
    from oracle import fce_oracle as O
    from oracle import nms_oracle

    torch.set_num_threads(threads)
    x = synth_images(1234, n_images, w["size"], w["size"])
    scale = cfg["scale"]

    def step():
        y, _ = O.forward(cfg, scale, sd, x)
        nms_oracle.non_max_suppression(y.numpy(), w["conf"], w["iou"], max_det=w["max_det"])

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return n_images * steps / dt, dt / steps * 1e3


def cpu_reference_run(w, n_images, steps, warmup, shipped_steps=2):
    """The reference exactly as shipped: ``YOLO(cfg).predict(tensor, device="cpu", half=False)`` from
    baseline/_ref - LoadTensor path (data/loaders.py:562-632), AutoBackend fuse (nn/autobackend.py:203-207), forward,
    the reference's own NMS + Results.  The reference caps its CPU threads at min(8, ncpu-1) (utils/__init__.py:44,
    utils/torch_utils.py:224): that default is timed first (`as_shipped`), then the same predictor with every host core.
    Same synthetic weights and images as the GPU arm (they decide how much work the NMS has)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    import ref_env

    from fce_yolo_b200.weights import synth_images

    yolo = ref_env.reference_yolo(w["yaml"], w.get("variant"), w["seed"])
    x = synth_images(1234, n_images, w["size"], w["size"])
    kw = dict(device="cpu", half=False, imgsz=w["size"], conf=w["conf"], iou=w["iou"], max_det=w["max_det"],
              verbose=False, save=False)

    def run(k):
        t0 = time.perf_counter()
        speed = None
        n_det = 0
        for _ in range(k):
            res = yolo.predict(x, **kw)
            speed = res[0].speed
            n_det = sum(len(r.boxes) for r in res)
        return (time.perf_counter() - t0) / max(k, 1), speed, n_det

    run(1)  # sets the predictor up: select_device("cpu") applies the thread cap here
    shipped_threads = torch.get_num_threads()
    sec_shipped, speed_shipped, _ = run(shipped_steps)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    for _ in range(warmup):
        run(1)
    sec, speed, n_det = run(steps)
    return dict(ips=n_images / sec, ms=sec * 1e3, cores=cores, speed_ms_per_image=speed,
                detections_per_image=n_det / n_images,
                as_shipped=dict(threads=shipped_threads, images_per_s=round(n_images / sec_shipped, 3),
                                ms_per_step=round(sec_shipped * 1e3, 2), speed_ms_per_image=speed_shipped))


def reference_line(a, w, name):
    """--impl reference: one JSON line for the reference's CPU implementation on this box's host cores."""
    n = 8  # images per step: the bounded sample of the workload (B in {1, 8})
    sample = f"{n} images per step x {a.steps} steps of the same model / image size / thresholds (bounded sample of the batch)"
    try:
        r = cpu_reference_run(w, n, a.steps, a.warmup)
        kind, cores, ips, ms = "reference", r["cores"], r["ips"], r["ms"]
        extra = {"as_shipped": r["as_shipped"], "speed_ms_per_image": r["speed_ms_per_image"],
                 "detections_per_image": r["detections_per_image"],
                 "api": "ultralytics.YOLO(cfg).predict(torch tensor BCHW, device='cpu', half=False) from baseline/_ref"}
    except Exception as e:  # noqa: BLE001 - reference not installed here: time the oracle port instead and say so
        cfg, model, sd = build_model(w)
        cores = os.cpu_count() or 1
        ips, ms = cpu_port_run(w, cfg, sd, n, a.steps, a.warmup, cores)
        kind, extra = "port", {"why_port": f"{type(e).__name__}: {str(e)[:160]}"}
    line = {"impl": "reference", "metric": "images/sec", "value": round(ips, 3), "unit": "images/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": round(ms, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_obj(name, w, a.gpus),
            "cpu_baseline": {"value": round(ips, 3), "unit": "images/s", "cores": cores, "kind": kind, "sample": sample,
                             **extra},
            "e2e": {"value": round(ips, 3), "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------- GPU arm
def kernel_table(pred, name):
    """Per-launch CUDA-event times of one eager pass (3 repetitions after a warm-up) -> (roofline object, rows)."""
    import ctypes as C

    ex = pred.ex
    stream = torch.cuda.current_stream()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(ex._calls) + 1)]
    reps = 3
    acc = [0.0] * len(ex._calls)
    for r in range(reps + 1):
        torch.cuda.synchronize()
        evs[0].record()
        for i, (fn, args, n) in enumerate(ex._calls):
            fn(*args, C.c_void_p(stream.cuda_stream))
            evs[i + 1].record()
        torch.cuda.synchronize()
        if r:  # first rep is warm-up
            for i in range(len(ex._calls)):
                acc[i] += evs[i].elapsed_time(evs[i + 1]) / reps
    classes, table = {}, []
    pk = peaks()
    from fce_yolo_b200.plan import DT_SIZE, View
    wpk = write_only_peak(pred.device)
    for (fn, args, n), ms in zip(ex._calls, acc):
        # the fused Detect epilogues are launches of the same tcgen05 conv kernel: one class
        cls = "fce_conv2d" if n.fn in ("fce_conv2d_detect", "fce_conv1x1_chain") else n.fn  # dense convs on the tensor cores
        c = classes.setdefault(cls, dict(ms=0.0, flops=0.0, bytes=0.0, launches=0, ideal=0.0, wbytes=0.0, ideal_w=0.0))
        c["ms"] += ms
        # this launch's own roofline: the slower of its tensor time and its HBM time
        ideal = max(n.flops / (pk["tf_sustained"] * 1e12), n.bytes / (pk["hbm"] * 1e9)) * 1e3
        c["ideal"] += ideal
        # ... and with the write stream held to the measured write-only bandwidth
        wb = sum(v.B * v.H * v.W * v.C * DT_SIZE[v.dtype] for v in n.writes if isinstance(v, View)) if n.bytes else 0.0
        c["wbytes"] += wb
        c["ideal_w"] += max(ideal, wb / (wpk * 1e9) * 1e3)
        c["flops"] += n.flops
        c["bytes"] += n.bytes
        c["launches"] += 1
        table.append((n.tag, n.fn, ms, n.flops, n.bytes))
    total = sum(c["ms"] for c in classes.values())
    fn, c = max(classes.items(), key=lambda kv: kv[1]["ms"])
    if c["flops"] > 0:
        ach = c["flops"] / (c["ms"] * 1e-3) / 1e12
        roof = {"kernel": fn, "bound": "tensor", "achieved": round(ach, 2), "peak": pk["tf_sustained"],
                "unit": "TFLOP/s", "frac": round(ach / pk["tf_sustained"], 4), "traffic": None,
                "peak_source": pk["src"] + " (sustained bf16 GEMM)", "share_of_step": round(c["ms"] / total, 3),
                "launches_per_step": c["launches"], "avg_launch_ms": round(c["ms"] / c["launches"], 4),
                # many conv launches of this model sit LEFT of the ridge (1x1s with Cin, Cout <= 512): fraction
                # of the per-launch roofline min(tensor peak, intensity x HBM peak), summed over the launches
                "instance_roofline_frac": round(c["ideal"] / c["ms"], 4),
                "algorithmic_gflop_per_launch": round(c["flops"] / c["launches"] / 1e9, 2)}
    else:
        ach = c["bytes"] / (c["ms"] * 1e-3) / 1e9
        roof = {"kernel": fn, "bound": "hbm", "achieved": round(ach, 1), "peak": pk["hbm"], "unit": "GB/s",
                "frac": round(ach / pk["hbm"], 4), "traffic": None, "peak_source": pk["src"],
                "share_of_step": round(c["ms"] / total, 3), "launches_per_step": c["launches"],
                "avg_launch_ms": round(c["ms"] / c["launches"], 4)}
    # measured DRAM traffic of the dominant class (ncu capture of the same step, committed under profiles/):
    # bytes per launch next to the algorithmic bytes per launch - a ratio well above 1 means wasted re-reads
    tj = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            tc = tj.get("classes", {}).get(fn)
            if tc and tj.get("workload") == name:
                roof["traffic"] = round(tc["dram_bytes_per_step"] / c["launches"], 1)
                roof["traffic_unit"] = "bytes per launch (ncu dram__bytes_read+write, profiles/traffic.json)"
                roof["algorithmic_bytes_per_launch"] = round(c["bytes"] / c["launches"], 1)
                if tj.get("note"):
                    roof["traffic_note"] = tj["note"]
        except (OSError, ValueError, KeyError):
            tj = {}
    roof["write_only_peak_GB/s"] = round(wpk, 1)
    roof["classes"] = {
        k: {"ms": round(v["ms"], 3), "share": round(v["ms"] / total, 3), "launches": v["launches"],
            **({"TFLOP/s": round(v["flops"] / (v["ms"] * 1e-3) / 1e12, 2)} if v["flops"] else {}),
            "roofline_frac": round(v["ideal"] / v["ms"], 3),
            **({"write_GB/s": round(v["wbytes"] / (v["ms"] * 1e-3) / 1e9, 1),
                "roofline_frac_write_aware": round(v["ideal_w"] / v["ms"], 3)} if v["wbytes"] else {}),
            **({"GB/s": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1),
                "hbm_frac": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9 / pk["hbm"], 3)} if v["bytes"] else {})}
        for k, v in sorted(classes.items(), key=lambda kv: -kv[1]["ms"])}
    if tj.get("workload") == name:
        for k, v in roof["classes"].items():
            tc = tj.get("classes", {}).get(k)
            if tc:
                v["dram_mb_per_step"] = round(tc["dram_bytes_per_step"] / 1e6, 1)
                v["algorithmic_mb_per_step"] = round(classes[k]["bytes"] / 1e6, 1)
    return roof, table


def measure(w, steps, warm, rank, world, dev, opts, sampler=None):
    """One configuration on this rank's GPU (all ranks run the same code: every barrier / collective below is reached
    by every rank).  Returns the per-rank result dict; times are already the max over ranks."""
    import torch.distributed as dist

    from fce_yolo_b200.predict import Predictor
    from fce_yolo_b200.runner import DetectionGather
    from fce_yolo_b200.weights import synth_images

    cfg, model, sd = build_model(w)
    B, S = w["batch"], w["size"]
    pred = Predictor(model, B, S, precision=w["precision"], device=dev, conf=w["conf"], iou=w["iou"],
                     max_det=w["max_det"], input_u8=True, use_graph=not opts.no_graph,
                     overlap_nms=not (opts.no_overlap or opts.no_graph))
    # synthetic uint8 NHWC batch (same seeded images, quantised), resident in HBM for `value`
    img = (synth_images(1234 + rank, B, S, S) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    h_img = img.pin_memory()
    pred.inp.copy_(h_img)
    torch.cuda.synchronize()

    gatherer = DetectionGather(pred.ex) if world > 1 else None  # one collective per step, out of the NMS's own buffer

    def step_device():
        pred.run_device()
        if world > 1:
            with torch.cuda.stream(pred._out_stream()):  # behind this step's NMS (its side stream in overlap mode)
                gatherer.gather()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warm):
        step_device()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        step_device()
    pred.join()  # the last step's NMS (and gather) belong to the timed region
    if world > 1:
        torch.cuda.current_stream().wait_stream(pred._out_stream())
    e1.record()
    barrier()
    ms_dev = e0.elapsed_time(e1)
    n_det = float(pred.count.float().mean())  # detections per image the NMS emitted for this batch

    # e2e: host buffers through the public API (Predictor.pipeline): every step's batch is copied H2D from pinned
    # memory and its detections D2H inside the timed region; the copy of batch i+1 overlaps the graph of batch i.
    # Two distinct pinned batches alternate so that no step can reuse the previous step's upload.
    h_img2 = (255 - img).pin_memory()
    for _ in pred.pipeline([h_img, h_img2] * 2):
        pass
    barrier()
    t0 = time.perf_counter()
    n_out = 0
    for det_h, cnt_h in pred.pipeline((h_img if s % 2 == 0 else h_img2) for s in range(steps)):
        n_out += int(cnt_h[0])  # touch the result on the host
    barrier()
    wall_e2e = (time.perf_counter() - t0) * 1e3

    if sampler is not None and rank == 0 and len(sampler.rows) < 3:  # keep the GPU busy until the sampler has data
        # rank-0-only code: LOCAL work only (run_device, no gather) - a collective here would wait for ranks that are
        # already in the all_reduce below (this hung a 4-GPU run, where nvidia-smi starts more slowly)
        t_end = time.perf_counter() + 1.5
        while time.perf_counter() < t_end and len(sampler.rows) < 3:
            pred.run_device()
            pred.join()
            torch.cuda.synchronize()
    t = torch.tensor([ms_dev, wall_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = t.tolist()
    n_img = B * world * steps
    pred.inp.copy_(h_img)
    pred.run_device()
    pred.join()
    torch.cuda.synchronize()
    y = pred.ex.outputs()[0]
    cand = float((y[:, 4:].amax(1) > w["conf"]).sum(1).float().mean())
    return dict(pred=pred, model=model, cfg=cfg, sd=sd, h_img=h_img, ms_dev=ms_dev, ms_e2e=ms_e2e, n_img=n_img,
                value=n_img / (ms_dev * 1e-3), e2e=n_img / (ms_e2e * 1e-3), detections_per_image=n_det,
                candidates_per_image=cand)


def latency_b1(w, model, h_img, dev, opts):
    """Batch-1 latency, BASELINE's second metric: back-to-back graph replays and Predictor.infer from pinned host."""
    from fce_yolo_b200.predict import Predictor

    p1 = Predictor(model, 1, w["size"], precision=w["precision"], device=dev, conf=w["conf"], iou=w["iou"],
                   max_det=w["max_det"], input_u8=True, use_graph=not opts.no_graph)
    p1.inp.copy_(h_img[:1])
    for _ in range(10):
        p1.run_device()
    torch.cuda.synchronize()
    n1 = 200
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    for _ in range(n1):
        p1.run_device()
    l1.record()
    torch.cuda.synchronize()
    dev_ms = l0.elapsed_time(l1) / n1
    t0 = time.perf_counter()
    for _ in range(n1):
        p1.infer(h_img[:1])  # H2D + graph + D2H + host sync, every image
    host_ms = (time.perf_counter() - t0) * 1e3 / n1
    return {"device_ms_per_img": round(dev_ms, 4), "e2e_ms_per_img": round(host_ms, 4), "launches": p1.launches_per_call,
            "note": "batch 1, same model/size; device = back-to-back graph replays, e2e = Predictor.infer from pinned host"}


def gpu_baseline(config_idx, batch, timeout_s=420):
    """Spawns baseline/ref_gpu_baseline.py (the reference DetectionModel through stock PyTorch on this GPU)."""
    cmd = [sys.executable, os.path.join(ROOT, "baseline", "ref_gpu_baseline.py"), "--batch", str(batch), "--modes",
           "eager,compile", "--steps", "5", "--warmup", "2"]
    if config_idx is not None:
        cmd += ["--config", str(config_idx)]
    try:
        r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"unavailable": f"rc={r.returncode}: {r.stderr.strip()[-200:]}"}
    except subprocess.TimeoutExpired:
        return {"unavailable": f"timed out after {timeout_s} s"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": f"{type(e).__name__}: {e}"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="override batch per GPU (exploration only)")
    ap.add_argument("--yaml", default=None)
    ap.add_argument("--size", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other four BASELINE configs")
    ap.add_argument("--config", type=int, default=None, choices=sorted(CONFIGS),
                    help="BASELINE.json configs[i] (exploration; the judged workload is configs[2], the default)")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="run the NMS in line instead of underneath the next forward")
    ap.add_argument("--kernel-times", default=None, help="write the per-node timing table to this file")
    a = ap.parse_args()

    w, name = workload(a.config)
    if a.batch or a.yaml or a.size:
        w.update({k: v for k, v in (("batch", a.batch), ("yaml", a.yaml), ("size", a.size)) if v})
        if a.yaml:
            w["variant"] = None
        name = f"{w['yaml']} predict bf16 batch {w['batch']}/GPU {w['size']}x{w['size']} (exploration)"
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    warm = max(a.warmup, 3)

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 alone)
    if a.impl == "reference":
        if rank != 0:
            return
        reference_line(a, w, name)
        return

    # ------------------------------------------------------------------ our arm (GPU)
    import torch.distributed as dist

    from fce_yolo_b200 import _lib

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.load(check_device=True)
    if world > 1:
        # NCCL prints its version banner to stdout when the first communicator is created; stdout is reserved for the
        # ONE JSON line, so file descriptor 1 points at stderr while the communicator comes up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            import datetime

            # a short collective timeout: a rank-asymmetric bug must fail loudly instead of holding the box
            dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)

    # clocks / throttle reasons are sampled from before the warm-up to the end of the e2e loop: the GPU is under the
    # same load throughout, and a 20-step timed region alone can be shorter than nvidia-smi's start-up
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    m = measure(w, a.steps, warm, rank, world, dev, a, sampler)
    clocks = sampler.stop() if rank == 0 else None
    pred = m["pred"]

    roof, lat, cpu, gpub, others = None, None, None, None, None
    if rank == 0:
        roof, table = kernel_table(pred, name)
        if a.kernel_times:
            with open(a.kernel_times, "w") as f:
                f.write("tag,fn,ms,gflop,mbytes\n")
                for tag, fnn, ms, fl, by in table:
                    f.write(f"{tag},{fnn},{ms:.4f},{fl / 1e9:.3f},{by / 1e6:.3f}\n")
    if rank == 0 and not a.no_latency:
        lat = latency_b1(w, m["model"], m["h_img"], dev, a)
    line = None
    if rank == 0:
        line = {
            "metric": "images/sec", "value": round(m["value"], 2), "unit": "images/s", "n_gpus": world,
            "steps": a.steps, "warmup": warm, "ms_per_step": round(m["ms_dev"] / a.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": config_obj(name, w, world),
            "impl_detail": {"cuda_graph": not a.no_graph, "arena_mb": round(pred.ex.nbytes / 2 ** 20, 1),
                            "nms_overlap": "NMS of step i runs on a side stream under the forward of step i+1; the last "
                                           "step's NMS is inside the timed region" if pred.overlap else False,
                            "bf16_convs_on_cuda_cores": len(pred.ex.simt_bf16_convs),
                            "detections_per_image": round(m["detections_per_image"], 1),
                            "nms_candidates_per_image": round(m["candidates_per_image"], 1)},
            "e2e": {"value": round(m["e2e"], 2), "unit": "images/s",
                    "h2d_bytes_per_step": pred.h2d_bytes(), "d2h_bytes_per_step": pred.d2h_bytes(),
                    "ms_per_step": round(m["ms_e2e"] / a.steps, 4),
                    "api": "Predictor.pipeline(pinned uint8 NHWC batches) -> pinned (det, count); H2D of batch i+1 "
                           "overlaps the graph of batch i"},
            "gpu_launches": pred.launches_per_call * a.steps,
            "latency_b1": lat, "clocks": clocks, "roofline": roof, "cpu_baseline": None, "gpu_baseline": None,
        }
    # free the big plan before the extra legs
    B_main = w["batch"]
    del pred
    m.clear()
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------ the other BASELINE configs (N=1, default run)
    if rank == 0 and world == 1 and a.config is None and not (a.no_configs or a.batch or a.yaml or a.size):
        others = {}
        for idx in sorted(CONFIGS):
            if idx == DEFAULT_CONFIG:
                continue
            wi, ni = workload(idx)
            try:
                mi = measure(wi, max(5, a.steps // 2), 3, 0, 1, dev, a)
                ri, _ = kernel_table(mi["pred"], ni)
                others[f"configs[{idx}]"] = {
                    "workload": ni, "value": round(mi["value"], 2), "unit": "images/s",
                    "ms_per_step": round(mi["ms_dev"] * wi["batch"] / mi["n_img"], 4), "e2e": round(mi["e2e"], 2),
                    "conv_TFLOP/s": ri["classes"].get("fce_conv2d", {}).get("TFLOP/s"),
                    "conv_frac_of_sustained_peak": round(ri["classes"].get("fce_conv2d", {}).get("TFLOP/s", 0.0) /
                                                         peaks()["tf_sustained"], 4),
                    "conv_instance_roofline_frac": ri["classes"].get("fce_conv2d", {}).get("roofline_frac"),
                    "detections_per_image": round(mi["detections_per_image"], 1),
                    "gpu_launches_per_step": mi["pred"].launches_per_call}
                if idx == 0:
                    others[f"configs[{idx}]"]["latency_b1"] = latency_b1(wi, mi["model"], mi["h_img"], dev, a)
                mi.clear()
            except Exception as e:  # noqa: BLE001 - an extra config must never lose the headline line
                others[f"configs[{idx}]"] = {"workload": ni, "error": f"{type(e).__name__}: {str(e)[:200]}"}
            torch.cuda.empty_cache()
        line["configs"] = others

    # ------------------------------------------------------------------ baselines on this box (rank 0, N=1 only)
    if rank == 0 and world == 1 and not a.no_gpu_baseline:
        line["gpu_baseline"] = gpu_baseline(a.config, B_main)
        modes = line["gpu_baseline"].get("modes") or {}
        # ours = forward + decode + NMS; the reference numbers below are its FORWARD ALONE (its NMS on this many candidates
        # runs into its own time limit, see modes.eager.nms_hit_time_limit) - so these ratios understate the gap
        for k in ("eager", "compile"):
            if modes.get(k, {}).get("forward_images_per_s"):
                line["gpu_baseline"][f"ours_predict_over_reference_{k}_forward"] = round(
                    line["value"] / modes[k]["forward_images_per_s"], 2)
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        n = 8
        try:
            r = cpu_reference_run(w, n, 3, 1)
            cpu = {"value": round(r["ips"], 3), "unit": "images/s", "cores": r["cores"], "kind": "reference",
                   "sample": f"YOLO(cfg).predict(tensor, device='cpu') from baseline/_ref, {n} images/step x 3 steps, "
                             f"{r['cores']} threads, same model / image size / thresholds",
                   "as_shipped": r["as_shipped"], "speed_ms_per_image": r["speed_ms_per_image"]}
        except Exception as e:  # noqa: BLE001
            cfg, model, sd = build_model(w)
            cores = os.cpu_count() or 1
            ips, ms = cpu_port_run(w, cfg, sd, 4, 3, 1, cores)
            cpu = {"value": round(ips, 3), "unit": "images/s", "cores": cores, "kind": "port",
                   "sample": f"oracle port (torch fp32 CPU forward + C/numpy NMS), 4 images/step x 3 steps, {cores} threads",
                   "why_port": f"{type(e).__name__}: {str(e)[:160]}"}
        line["cpu_baseline"] = cpu

    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
