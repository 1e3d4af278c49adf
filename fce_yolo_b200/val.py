"""Validation statistics on the GPU: the per-image ``_process_batch`` / ``update_metrics`` loop of
``DetectionValidator`` (ultralytics/models/yolo/detect/val.py:168-211, 274-288) without the per-image device->host
round trips.  ``match_predictions`` turns the padded NMS output of a whole batch into the ``tp [n, 10]`` matrix in
one launch; ``ValStats`` accumulates the reference's ``stats`` dict (tp, conf, pred_cls, target_cls, target_img -
ultralytics/utils/metrics.py:1118) and gathers it to rank 0 over NCCL (runner.gather_stats_to_rank0) in place of
the pickled ``dist.gather_object`` of val.py:222-242.  ``ap_per_class`` itself (metrics.py:817) stays on the host:
``ValStats.metrics()`` hands the gathered statistics to ``metrics.detection_metrics`` (numpy restatement, checked
against the live reference to 1e-12)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from .runner import gather_stats_to_rank0

IOUV = torch.linspace(0.5, 0.95, 10)  # val.py:52


def match_predictions(det: torch.Tensor, count: torch.Tensor, gt_boxes: torch.Tensor, gt_cls: torch.Tensor,
                      gt_offsets, iouv: torch.Tensor | None = None) -> torch.Tensor:
    """det [B, max_det, 6], count [B] (NMS outputs, CUDA); labels of image b = rows gt_offsets[b]:gt_offsets[b+1] of
    gt_boxes [G, 4] (xyxy, same coordinate space as det) / gt_cls [G].  Returns bool tp [B, max_det, n_iou]."""
    if not det.is_cuda:
        raise RuntimeError("fce_yolo_b200 validation matching runs on the GPU only (no CPU fallback)")
    dev = det.device
    B, max_det = det.shape[0], det.shape[1]
    offs = np.asarray(gt_offsets, dtype=np.int32)
    if offs.shape != (B + 1,) or offs[0] != 0 or (np.diff(offs) < 0).any():
        raise ValueError("gt_offsets must be B+1 non-decreasing offsets starting at 0")
    G = int(offs[-1])
    if gt_boxes.shape != (G, 4) or gt_cls.shape != (G,):
        raise ValueError(f"labels must be [{G}, 4] boxes and [{G}] classes")
    iouv = (IOUV if iouv is None else iouv).to(dev, torch.float32).contiguous()
    gb = gt_boxes.to(dev, torch.float32).contiguous() if G else torch.zeros(1, 4, device=dev)
    gc = gt_cls.to(dev, torch.float32).contiguous() if G else torch.zeros(1, device=dev)
    go = torch.from_numpy(offs).to(dev)
    tp = torch.empty(B, max_det, iouv.numel(), dtype=torch.uint8, device=dev)
    d32 = det.contiguous()
    cnt = count.to(torch.int32).contiguous()
    with torch.cuda.device(dev):
        st = L.load(check_device=True).fce_match_predictions(
            C.c_void_p(d32.data_ptr()), C.c_void_p(cnt.data_ptr()), C.c_void_p(gb.data_ptr()), C.c_void_p(gc.data_ptr()),
            C.c_void_p(go.data_ptr()), C.c_void_p(iouv.data_ptr()), B, max_det, iouv.numel(),
            int(np.diff(offs).max()) if B else 0, C.c_void_p(tp.data_ptr()),
            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    L.check(st, "fce_match_predictions")
    return tp.bool()


class ValStats:
    """Accumulates the reference's validation statistics over batches on the device."""

    def __init__(self, device=None):
        self.parts = {k: [] for k in ("tp", "conf", "pred_cls", "target_cls", "target_img")}
        # where EMPTY statistics live: a rank whose shard holds no batch (more ranks than batches) still takes part in the
        # collectives of result(), and NCCL needs its (empty) tensors on this rank's GPU
        self.device = torch.device(device) if device is not None else (
            torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu"))

    def update(self, det, count, gt_boxes, gt_cls, gt_offsets):
        tp = match_predictions(det, count, gt_boxes, gt_cls, gt_offsets)
        B, max_det = det.shape[:2]
        valid = torch.arange(max_det, device=det.device)[None, :] < count[:, None]
        self.parts["tp"].append(tp[valid])
        self.parts["conf"].append(det[..., 4][valid])
        self.parts["pred_cls"].append(det[..., 5][valid])
        self.parts["target_cls"].append(gt_cls.to(det.device, torch.float32))
        offs = np.asarray(gt_offsets)
        for b in range(B):  # classes present per image (np.unique(cls), val.py:185)
            self.parts["target_img"].append(torch.unique(gt_cls[offs[b]:offs[b + 1]]).to(det.device, torch.float32))

    def result(self, group=None):
        """Concatenated stats, gathered to rank 0 when torch.distributed is initialised (None on other ranks)."""
        dev = next((p[0].device for p in self.parts.values() if p), self.device)
        stats = {k: (torch.cat(v) if v else torch.zeros((0, 10) if k == "tp" else (0,), device=dev)) for k, v in self.parts.items()}
        stats["tp"] = stats["tp"].to(torch.uint8)
        out = gather_stats_to_rank0(stats, group)
        if out is None:
            return None
        out["tp"] = out["tp"].bool()
        return out

    def metrics(self, group=None):
        """mp / mr / mAP50 / mAP50-95 and the per-class table (``metrics.detection_metrics``) from everything
        accumulated so far; on rank 0 only when torch.distributed is initialised (None elsewhere) - the reference's
        ``DetectionValidator.get_stats`` (val.py:244-262) after ``gather_stats``."""
        from .metrics import detection_metrics

        st = self.result(group)
        if st is None:
            return None
        return detection_metrics(st["tp"], st["conf"], st["pred_cls"], st["target_cls"])
