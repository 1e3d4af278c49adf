"""Public predict API: the call a user makes (mirror of ``YOLO(...).predict`` for tensors / uint8 batches,
reference engine/model.py:477-535 -> engine/predictor.py:276-381 -> detect/predict.py:33-80).

One Predictor owns one compiled plan (forward + decode + NMS captured in a single CUDA graph), a static
device input buffer and pinned host staging buffers, so a call is: H2D copy, graph launch, D2H of the padded
detections.  ``__call__`` returns the reference's list of ``[n_i, 6]`` (x1,y1,x2,y2,conf,cls) tensors.
"""
from __future__ import annotations

import torch

from .engine import Executor
from .plan import compile_model


class Predictor:
    def __init__(self, model, batch: int, imgsz=(640, 640), precision: str = "bf16", device=None, conf: float = 0.25,
                 iou: float = 0.7, max_det: int = 300, agnostic_nms: bool = False, multi_label: bool = False,
                 input_u8: bool = True, use_graph: bool = True):
        if isinstance(imgsz, int):
            imgsz = (imgsz, imgsz)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise RuntimeError("fce_yolo_b200 runs on a B200 GPU only (no CPU fallback)")
        self.batch, self.imgsz, self.input_u8 = batch, imgsz, input_u8
        if model.training:
            model.eval()
        with torch.cuda.device(self.device):
            plan = compile_model(model, batch, imgsz[0], imgsz[1], precision, self.device, input_u8=input_u8,
                                 nms=dict(conf=conf, iou=iou, max_det=max_det, agnostic=agnostic_nms,
                                          multi_label=multi_label))
            self.ex = Executor(plan, use_graph=use_graph)
            self.stream = torch.cuda.Stream(self.device)
            self.inp = self.ex.input_tensor()
            self.det, self.keep, self.count = self.ex.detections()
            self.h_det = torch.empty(self.det.shape, dtype=torch.float32).pin_memory()
            self.h_count = torch.empty(self.count.shape, dtype=torch.int32).pin_memory()
        self.launches_per_call = self.ex.launches_per_run

    @property
    def input_shape(self):
        return tuple(self.inp.shape)

    def run_device(self):
        """Forward + decode + NMS on whatever is in the static input buffer (no host traffic)."""
        self.ex.run()
        return self.det, self.keep, self.count

    def infer(self, images: torch.Tensor):
        """images: host (ideally pinned) or device tensor matching ``input_shape`` - uint8 NHWC when
        ``input_u8`` else fp32 NCHW in [0,1].  Returns pinned host (det [B,max_det,6], count [B])."""
        if tuple(images.shape) != tuple(self.inp.shape):
            raise ValueError(f"expected input of shape {tuple(self.inp.shape)}, got {tuple(images.shape)}")
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self.inp.copy_(images, non_blocking=True)
            self.ex.run()
            self.h_det.copy_(self.det, non_blocking=True)
            self.h_count.copy_(self.count, non_blocking=True)
        self.stream.synchronize()
        return self.h_det, self.h_count

    def __call__(self, images: torch.Tensor):
        det, count = self.infer(images)
        return [det[b, :n].clone() for b, n in enumerate(count.tolist())]

    def h2d_bytes(self):
        return self.inp.numel() * self.inp.element_size()

    def d2h_bytes(self):
        return self.h_det.numel() * 4 + self.h_count.numel() * 4
