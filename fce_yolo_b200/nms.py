"""GPU non-max suppression behind the reference's ``non_max_suppression`` signature
(ultralytics/utils/nms.py:13-29), backed by the batched sm_100a kernel (csrc/nms.cu).

Two entry points:
 * ``nms_batched`` - fixed-shape, sync-free: det [B, max_det, 6], keep [B, max_det] (int64 anchor
   indices, -1 padded), count [B] (int32).  This is what the multi-GPU runner all-gathers.
 * ``non_max_suppression`` - the reference's list-of-tensors API (one device->host read of the counts).

Intentional divergences (documented in DESIGN.md): no wall-clock bail-out (nms.py:81,162-164 is
nondeterministic), the input is not overwritten in place (nms.py:86 does, through a transposed view), and
exact score ties among more than ``max_nms`` candidates are ordered by ascending index (the reference's
pre-sort is an unstable argsort, nms.py:138).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L


def nms_batched(prediction: torch.Tensor, conf_thres: float = 0.25, iou_thres: float = 0.45, classes=None,
                agnostic: bool = False, multi_label: bool = False, max_det: int = 300, nc: int = 0,
                max_nms: int = 30000, max_wh: int = 7680, workspace: torch.Tensor | None = None, out=None):
    if not (0 <= conf_thres <= 1):
        raise AssertionError(f"Invalid Confidence threshold {conf_thres}, valid values are between 0.0 and 1.0")
    if not (0 <= iou_thres <= 1):
        raise AssertionError(f"Invalid IoU {iou_thres}, valid values are between 0.0 and 1.0")
    if not prediction.is_cuda:
        raise RuntimeError("fce_yolo_b200 NMS runs on the GPU only (no CPU fallback)")
    lib = L.load(check_device=True)
    p = prediction
    if p.dtype != torch.float32 or not p.is_contiguous():
        p = p.float().contiguous()
    B, ch, A = p.shape
    nc = nc or ch - 4
    if ch != 4 + nc:
        raise ValueError("extra (mask) channels are outside the detection path")
    multi_label = bool(multi_label and nc > 1)
    cls_t = None
    if classes is not None:
        cls_t = torch.as_tensor(list(classes), dtype=torch.int32, device=p.device)
    d = L.NmsDesc(B=B, A=A, nc=nc, conf_thres=float(conf_thres), iou_thres=float(iou_thres), max_det=int(max_det),
                  max_nms=int(max_nms), multi_label=int(multi_label), agnostic=int(bool(agnostic)),
                  max_wh=float(max_wh), n_classes=0 if cls_t is None else cls_t.numel())
    need = lib.fce_nms_workspace(C.byref(d))
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        workspace = torch.empty(need, dtype=torch.uint8, device=p.device)
    if out is None:
        det = torch.empty(B, max_det, 6, dtype=torch.float32, device=p.device)
        keep = torch.empty(B, max_det, dtype=torch.int64, device=p.device)
        count = torch.empty(B, dtype=torch.int32, device=p.device)
    else:
        det, keep, count = out
    with torch.cuda.device(p.device):
        st = lib.fce_nms(C.byref(d), C.c_void_p(p.data_ptr()), C.c_void_p(cls_t.data_ptr() if cls_t is not None else 0),
                         C.c_void_p(det.data_ptr()), C.c_void_p(keep.data_ptr()), C.c_void_p(count.data_ptr()),
                         C.c_void_p(workspace.data_ptr()), C.c_size_t(workspace.numel() * workspace.element_size()),
                         C.c_void_p(torch.cuda.current_stream(p.device).cuda_stream))
    L.check(st, "fce_nms")
    return det, keep, count


def non_max_suppression(prediction, conf_thres: float = 0.25, iou_thres: float = 0.45, classes=None,
                        agnostic: bool = False, multi_label: bool = False, labels=(), max_det: int = 300,
                        nc: int = 0, max_time_img: float = 0.05, max_nms: int = 30000, max_wh: int = 7680,
                        rotated: bool = False, end2end: bool = False, return_idxs: bool = False):
    if isinstance(prediction, (list, tuple)):
        prediction = prediction[0]  # (inference_out, raw) - nms.py:61-62
    if rotated or end2end or prediction.shape[-1] == 6 or (labels and any(len(l) for l in labels)):
        raise NotImplementedError("rotated / end2end / a-priori-label NMS is outside the FCE detection path")
    det, keep, count = nms_batched(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, max_det, nc,
                                   max_nms, max_wh)
    counts = count.tolist()  # the one device->host sync the list-of-tensors API needs
    output = [det[b, :n] for b, n in enumerate(counts)]
    if return_idxs:
        return output, [keep[b, :n] for b, n in enumerate(counts)]
    return output
