"""Public predict API: the call a user makes (mirror of ``YOLO(...).predict`` for tensors / uint8 batches,
reference engine/model.py:477-535 -> engine/predictor.py:276-381 -> detect/predict.py:33-80).

One Predictor owns one compiled plan (forward + decode + NMS captured in a single CUDA graph), a static
device input buffer and pinned host staging buffers, so a call is: H2D copy, graph launch, D2H of the padded
detections.  ``__call__`` returns the reference's list of ``[n_i, 6]`` (x1,y1,x2,y2,conf,cls) tensors.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .engine import Executor
from .plan import compile_model


def scale_meta(img1_shape, img0_shapes) -> np.ndarray:
    """[B, 5] fp32 {gain, pad_x, pad_y, w0, h0} for fce_scale_boxes: the gain / padding arithmetic of ops.scale_boxes
    (ultralytics/utils/ops.py:118-121) for letterboxed shape ``img1_shape`` = (h, w) and original shapes (h0, w0)."""
    m = np.empty((len(img0_shapes), 5), dtype=np.float32)
    for b, s0 in enumerate(img0_shapes):
        gain = min(img1_shape[0] / s0[0], img1_shape[1] / s0[1])
        m[b] = (gain, round((img1_shape[1] - s0[1] * gain) / 2 - 0.1), round((img1_shape[0] - s0[0] * gain) / 2 - 0.1),
                s0[1], s0[0])
    return m


class Predictor:
    def __init__(self, model, batch: int, imgsz=(640, 640), precision: str = "bf16", device=None, conf: float = 0.25,
                 iou: float = 0.7, max_det: int = 300, agnostic_nms: bool = False, multi_label: bool = False,
                 input_u8: bool = True, use_graph: bool = True, overlap_nms: bool = False,
                 fuse_decode: bool = True):
        if isinstance(imgsz, int):
            imgsz = (imgsz, imgsz)
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise RuntimeError("fce_yolo_b200 runs on a B200 GPU only (no CPU fallback)")
        self.batch, self.imgsz, self.input_u8 = batch, imgsz, input_u8
        if model.training:
            model.eval()
        nc = int(getattr(list(model.model)[-1], "nc", 0) or 0)
        self.names = getattr(model, "names", None) or {i: f"{i}" for i in range(nc)}  # tasks.py:371 default names
        with torch.cuda.device(self.device):
            plan = compile_model(model, batch, imgsz[0], imgsz[1], precision, self.device, input_u8=input_u8,
                                 nms=dict(conf=conf, iou=iou, max_det=max_det, agnostic=agnostic_nms,
                                          multi_label=multi_label),
                                 fuse_decode=fuse_decode)  # predict never returns Detect's raw logit maps
            self.ex = Executor(plan, use_graph=use_graph)
            self.stream = torch.cuda.Stream(self.device)
            self.stream_ = self.stream
            self.inp = self.ex.input_tensor()
            self.det, self.keep, self.count = self.ex.detections()
            # det and count sit back to back in one arena buffer (plan.Plan.nms): ONE device->host copy per call
            self.detcount = self.ex.bytes(plan.outputs["detcount"])
            self.h_det, self.h_count, self._h_dc = self._host_slot()
            self.overlap = bool(overlap_nms and use_graph)
            if self.overlap:  # NMS of call i runs on a side stream underneath the forward of call i+1 (engine.py)
                self.ex.enable_overlap()
        self.launches_per_call = self.ex.launches_per_run

    def _host_slot(self):
        """Pinned host mirror of the packed (det, count) buffer: (det view [B,max_det,6], count view [B], raw bytes)."""
        raw = torch.empty(self.detcount.numel(), dtype=torch.uint8).pin_memory()
        nd = self.det.numel() * 4
        return (raw[:nd].view(torch.float32).view(self.det.shape), raw[nd:nd + 4 * self.batch].view(torch.int32), raw)

    def _run(self):
        if self.overlap:
            self.ex.run_overlapped()
        else:
            self.ex.run()

    def _out_stream(self):
        """Stream on which det / keep / count of the last call are valid in stream order."""
        return self.ex.tail_stream if self.overlap else torch.cuda.current_stream(self.device)

    def join(self):
        """Makes the caller's current stream wait for the detections of the last run_device()."""
        self.ex.join()

    @property
    def input_shape(self):
        return tuple(self.inp.shape)

    def run_device(self):
        """Forward + decode + NMS on whatever is in the static input buffer (no host traffic).  In overlap mode the
        NMS is still in flight on a side stream when this returns: call join() before reading the outputs on the
        current stream (a device-wide synchronize also does)."""
        self._run()
        return self.det, self.keep, self.count

    def infer(self, images: torch.Tensor):
        """images: host (ideally pinned) or device tensor matching ``input_shape`` - uint8 NHWC when
        ``input_u8`` else fp32 NCHW in [0,1].  Returns pinned host (det [B,max_det,6], count [B])."""
        if tuple(images.shape) != tuple(self.inp.shape):
            raise ValueError(f"expected input of shape {tuple(self.inp.shape)}, got {tuple(images.shape)}")
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self.inp.copy_(images, non_blocking=True)
            self._run()
            self.join()
            self._h_dc.copy_(self.detcount, non_blocking=True)
        self.stream.synchronize()
        return self.h_det, self.h_count

    def pipeline(self, batches):
        """Pipelined inference over an iterable of host batches (each shaped like ``input_shape``, ideally pinned):
        the host->device copy of batch i+1 runs on a copy stream while batch i is in the graph, and the padded
        detections of batch i come back device->host on the compute stream.  Yields, per batch and in order, pinned
        host tensors (det [B,max_det,6], count [B]) - valid until the next-but-one ``next()`` (three result slots: the
        device->host copy of batch i+2 is enqueued while the caller still holds batch i, into a different slot).
        Every batch still crosses PCIe inside the call: this is the end-to-end path, only overlapped."""
        dev = self.device
        with torch.cuda.device(dev):
            if not hasattr(self, "_pipe"):
                self._pipe = dict(
                    copy_stream=torch.cuda.Stream(dev),
                    stage=[torch.empty_like(self.inp) for _ in range(2)],
                    slot=[self._host_slot() for _ in range(3)],
                    copied=[torch.cuda.Event() for _ in range(2)],
                    consumed=[torch.cuda.Event() for _ in range(2)],
                    done=[torch.cuda.Event() for _ in range(3)])
            P = self._pipe
            cs, ms = P["copy_stream"], self.stream_
            it = iter(batches)

            def upload(i, images):
                if tuple(images.shape) != tuple(self.inp.shape):
                    raise ValueError(f"expected input of shape {tuple(self.inp.shape)}, got {tuple(images.shape)}")
                with torch.cuda.stream(cs):
                    cs.wait_event(P["consumed"][i % 2])  # the staging slot was drained by compute two batches ago
                    P["stage"][i % 2].copy_(images, non_blocking=True)
                    P["copied"][i % 2].record(cs)

            for ev in P["consumed"]:
                ev.record(ms)
            nxt = next(it, None)
            i = 0
            if nxt is not None:
                upload(0, nxt)
            pending = None
            while nxt is not None:
                cur_i = i
                nxt = next(it, None)
                with torch.cuda.stream(ms):
                    ms.wait_event(P["copied"][cur_i % 2])
                    self.inp.copy_(P["stage"][cur_i % 2], non_blocking=True)  # device-to-device, ~30 us
                    P["consumed"][cur_i % 2].record(ms)
                    self._run()
                    with torch.cuda.stream(self._out_stream()):  # in overlap mode: behind the NMS, on its side stream
                        P["slot"][cur_i % 3][2].copy_(self.detcount, non_blocking=True)  # det + count: one copy
                        P["done"][cur_i % 3].record(torch.cuda.current_stream(dev))
                if nxt is not None:
                    upload(cur_i + 1, nxt)  # overlaps the graph that was just launched
                if pending is not None:
                    P["done"][pending % 3].synchronize()
                    yield P["slot"][pending % 3][0], P["slot"][pending % 3][1]
                pending = cur_i
                i += 1
            if pending is not None:
                P["done"][pending % 3].synchronize()
                yield P["slot"][pending % 3][0], P["slot"][pending % 3][1]

    def predict(self, images, as_results: bool = False, names=None, paths=None):
        """The reference's ``YOLO(...).predict(list_of_bgr_frames)`` for raw uint8 HWC BGR images of any size (at most
        ``batch`` of them): LetterBox + BGR->RGB on the GPU (fce_letterbox) -> forward + decode + NMS (one CUDA graph) ->
        ops.scale_boxes + clip on the GPU (fce_scale_boxes) -> one D2H of the padded detections.  Returns the
        reference's list of ``[n_i, 6]`` (x1, y1, x2, y2, conf, cls) tensors in ORIGINAL image coordinates
        (engine/predictor.py:151-201, models/yolo/detect/predict.py:33-122).  ``as_results=True`` wraps them like
        ``DetectionPredictor.construct_result`` does (predict.py:109-122): a list of ``results.Results`` whose ``.boxes``
        mirror the reference's ``Boxes`` accessors (xyxy / conf / cls / xywh / xyxyn / xywhn)."""
        from .preprocess import LetterBoxGPU

        n = len(images)
        if not 0 < n <= self.batch:
            raise ValueError(f"predict() takes 1..{self.batch} images per call, got {n}")
        if not self.input_u8:
            raise RuntimeError("predict() on raw frames needs a Predictor built with input_u8=True")
        if not hasattr(self, "_lb"):
            self._lb = LetterBoxGPU(self.imgsz, auto=False, device=self.device)
            self._meta_pin = torch.empty(self.batch, 5, dtype=torch.float32).pin_memory()
            self._meta_dev = torch.empty(self.batch, 5, dtype=torch.float32, device=self.device)
        meta = scale_meta(self.imgsz, [im.shape[:2] for im in images])
        self._meta_pin[:n].copy_(torch.from_numpy(meta))
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self._lb(images, out=self.inp[:n])
            self._meta_dev[:n].copy_(self._meta_pin[:n], non_blocking=True)
            self._run()
            self.join()
            st = L.load().fce_scale_boxes(C.c_void_p(self.det.data_ptr()), C.c_void_p(self.count.data_ptr()),
                                          C.c_void_p(self._meta_dev.data_ptr()), n, self.det.shape[1],
                                          C.c_void_p(self.stream.cuda_stream))
            L.check(st, "fce_scale_boxes")
            self._h_dc.copy_(self.detcount, non_blocking=True)
        self.stream.synchronize()
        dets = [self.h_det[b, :c].clone() for b, c in enumerate(self.h_count[:n].tolist())]
        if not as_results:
            return dets
        from .results import Results

        names = names if names is not None else getattr(self, "names", None)
        return [Results(im.shape[:2], d, names=names, path=(paths[i] if paths else f"image{i}.jpg"), orig_img=im)
                for i, (im, d) in enumerate(zip(images, dets))]

    def __call__(self, images: torch.Tensor):
        det, count = self.infer(images)
        return [det[b, :n].clone() for b, n in enumerate(count.tolist())]

    def h2d_bytes(self):
        return self.inp.numel() * self.inp.element_size()

    def d2h_bytes(self):
        return self._h_dc.numel()
