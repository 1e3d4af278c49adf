"""``YOLO``: the reference's user-facing entry point for this path (ultralytics/engine/model.py ``Model.predict`` /
``Model.val`` as used with a detect ``*.yaml``), as a thin composition of the pieces of this package:

    from fce_yolo_b200 import YOLO
    model = YOLO("yolo11s-fce.yaml")            # or a packed model written by packed.save_packed (".fcepack")
    results = model.predict(list_of_bgr_frames, imgsz=640, conf=0.25, iou=0.7)   # -> list[results.Results]
    print(results[0].boxes.xyxy, results[0].summary())
    stats = model.val(batches_of_(frames, labels))                                 # -> mp / mr / mAP50 / mAP50-95

``predict`` = LetterBox + BGR->RGB on the GPU -> the compiled plan (one CUDA graph: forward + decode + NMS) ->
``scale_boxes`` on the GPU -> ``Results`` (``predict.Predictor.predict``).  ``val`` runs the same plan with the
validator's settings (conf 0.001, multi-label, val.py:105-126), matches on the GPU (``val.ValStats``) and reduces
on the host (``metrics.detection_metrics``).  Predictors are cached per (batch, imgsz, thresholds): a new shape
compiles a new plan, like the reference re-runs ``setup_model`` / warm-up (engine/predictor.py:293-330).

There is no CPU fallback: construction works anywhere (it only builds the module graph), the first ``predict`` /
``val`` needs a B200."""
from __future__ import annotations

import os

import numpy as np
import torch

from .tasks import DetectionModel


class YOLO:
    task = "detect"

    def __init__(self, model="yolo11n-fce.yaml", precision: str = "bf16", device=None, verbose: bool = False):
        if isinstance(model, torch.nn.Module):
            self.model = model
        elif isinstance(model, dict) or str(model).endswith((".yaml", ".yml")):
            self.model = DetectionModel(model)
        elif str(model).endswith(".fcepack"):
            from .packed import load_packed

            self.model = load_packed(str(model))
        else:
            raise ValueError(f"'{model}': expected a detect *.yaml, a cfg dict, a .fcepack file or an nn.Module "
                             "(pickled .pt checkpoints are converted once with packed.save_packed)")
        self.model.eval()
        if hasattr(self.model, "fuse"):
            self.model.fuse()
        self.precision, self.device, self.verbose = precision, device, verbose
        nc = int(getattr(list(self.model.model)[-1], "nc", 0) or 0)
        self.names = getattr(self.model, "names", None) or {i: f"{i}" for i in range(nc)}
        self._predictors = {}
        self._predictor_cls = None  # test seam: defaults to predict.Predictor

    # ------------------------------------------------------------------------------------------------
    def _predictor(self, batch, imgsz, conf, iou, max_det, agnostic_nms, multi_label):
        key = (batch, tuple(imgsz), float(conf), float(iou), int(max_det), bool(agnostic_nms), bool(multi_label))
        if key not in self._predictors:
            if self._predictor_cls is None:
                from .predict import Predictor as cls
            else:
                cls = self._predictor_cls
            self._predictors[key] = cls(self.model, batch, tuple(imgsz), precision=self.precision, device=self.device,
                                        conf=conf, iou=iou, max_det=max_det, agnostic_nms=agnostic_nms,
                                        multi_label=multi_label, input_u8=True)
        return self._predictors[key]

    @staticmethod
    def _imgsz(imgsz):
        hw = (imgsz, imgsz) if isinstance(imgsz, int) else tuple(imgsz)
        if len(hw) != 2 or any(v <= 0 or v % 32 for v in hw):
            raise ValueError(f"imgsz={imgsz} must be one or two positive multiples of 32 (max stride, loaders.py:603-609)")
        return hw

    @staticmethod
    def _frames(source):
        if isinstance(source, np.ndarray):
            source = [source] if source.ndim == 3 else list(source)
        frames = list(source)
        for im in frames:
            if not (isinstance(im, np.ndarray) and im.ndim == 3 and im.shape[2] == 3 and im.dtype == np.uint8):
                raise TypeError("predict() takes uint8 HWC BGR images (numpy), one array or a list of them "
                                "(files / streams / URLs are the reference's loaders: outside this path)")
        return frames

    def predict(self, source, imgsz=640, conf: float = 0.25, iou: float = 0.7, max_det: int = 300,
                agnostic_nms: bool = False, batch: int | None = None, **unused):
        """Returns one ``Results`` per image, in order (Model.predict, engine/model.py; DetectionPredictor defaults:
        cfg/default.yaml conf 0.25, iou 0.7, max_det 300)."""
        frames = self._frames(source)
        if not frames:
            return []
        hw = self._imgsz(imgsz)
        bs = int(batch) if batch else min(len(frames), 64)
        p = self._predictor(bs, hw, conf, iou, max_det, agnostic_nms, False)
        out = []
        for i in range(0, len(frames), bs):
            out += p.predict(frames[i:i + bs], as_results=True, names=self.names,
                             paths=[f"image{i + j}.jpg" for j in range(len(frames[i:i + bs]))])
        return out

    __call__ = predict

    def val(self, batches, imgsz=640, conf: float = 0.001, iou: float = 0.7, max_det: int = 300, batch: int | None = None):
        """``batches``: the path of a YOLO-format ``data.yaml`` (read by ``data.iter_val_batches``), or any
        iterable of (frames, labels) with ``frames`` a list of uint8 HWC BGR images and ``labels`` a
        list of float arrays [n_i, 5] = (cls, x1, y1, x2, y2) in ORIGINAL image pixels.  Validator settings of the
        reference (val.py:105-126: conf 0.001, multi-label NMS).  Returns ``metrics.detection_metrics``'s dict (on
        rank 0 when torch.distributed is initialised, None elsewhere)."""
        from .val import ValStats

        hw = self._imgsz(imgsz)
        if isinstance(batches, (str, os.PathLike, dict)):  # a data.yaml (YOLO-format dataset): Model.val(data=...)
            from .data import iter_val_batches

            batches = iter_val_batches(batches, batch=int(batch) if batch else 16)
        stats, p = ValStats(), None
        for frames, labels in batches:
            frames = self._frames(frames)
            if len(frames) != len(labels):
                raise ValueError("one label array per image")
            if p is None:
                p = self._predictor(int(batch) if batch else len(frames), hw, conf, iou, max_det, False, True)
            p.predict(frames)  # detections of this batch stay on the device, already in original-image coordinates
            n = len(frames)
            lab = [np.asarray(l, dtype=np.float32).reshape(-1, 5) for l in labels]
            offs = np.concatenate(([0], np.cumsum([len(l) for l in lab]))).astype(np.int64)
            cat = torch.from_numpy(np.concatenate(lab) if offs[-1] else np.zeros((0, 5), np.float32))
            stats.update(p.det[:n], p.count[:n], cat[:, 1:5], cat[:, 0], offs)
        return stats.metrics()

    def save(self, path: str, dtype=torch.bfloat16):
        """Pickle-free packed model (packed.save_packed): reload with ``YOLO(path)``."""
        from .packed import save_packed

        if not str(path).endswith(".fcepack"):
            raise ValueError("save() writes the packed format: use a .fcepack file name")
        return save_packed(self.model, str(path), dtype=dtype)

    def info(self):
        n_p = sum(p.numel() for p in self.model.parameters())
        return {"layers": len(list(self.model.model)), "parameters": n_p, "names": len(self.names), "task": self.task}
