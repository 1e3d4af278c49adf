"""CPU oracle for validation matching - TEST INFRASTRUCTURE ONLY (see fce_oracle.py).

Restates ``box_iou`` (ultralytics/utils/metrics.py:57-77) and the non-scipy branch of
``BaseValidator.match_predictions`` (ultralytics/engine/validator.py:266-306) as used by
``DetectionValidator._process_batch`` (ultralytics/models/yolo/detect/val.py:274-288).

What the reference's sort / unique / unique sequence computes, per IoU threshold t:
  * a detection d is a candidate if its best same-class label (highest IoU) has IoU >= t - ``np.unique`` over the
    detection column keeps the first, i.e. highest-IoU, match of every detection;
  * of the candidates sharing a best label, the one with the LOWEST detection index is correct - the second
    ``np.unique`` runs over a list that is by then ordered by detection index.
Exact IoU ties between two labels of one detection are ordered by numpy's unstable argsort in the reference, i.e.
unspecified; here (and in the CUDA kernel) the label with the higher index wins.  All arithmetic is fp32 in the
reference's operation order.

Pinned against the live reference (tests/golden/make_golden.py, ``match_*`` fixtures).
"""
from __future__ import annotations

import numpy as np

IOUV = np.linspace(0.5, 0.95, 10, dtype=np.float32)  # val.py:52 torch.linspace(0.5, 0.95, 10)


def box_iou(box1: np.ndarray, box2: np.ndarray, eps: float = 1e-7) -> np.ndarray:
    """[N,4] x [M,4] xyxy fp32 -> [N,M] fp32, op order of metrics.py:71-77."""
    b1, b2 = box1.astype(np.float32), box2.astype(np.float32)
    a1, a2 = b1[:, None, :2], b1[:, None, 2:]
    c1, c2 = b2[None, :, :2], b2[None, :, 2:]
    wh = np.clip(np.minimum(a2, c2) - np.maximum(a1, c1), 0, None).astype(np.float32)
    inter = wh[..., 0] * wh[..., 1]
    area1 = (a2 - a1)[..., 0] * (a2 - a1)[..., 1]
    area2 = (c2 - c1)[..., 0] * (c2 - c1)[..., 1]
    return (inter / (((area1 + area2) - inter) + np.float32(eps))).astype(np.float32)


def match_predictions(pred_boxes, pred_cls, gt_boxes, gt_cls, iouv=IOUV) -> np.ndarray:
    """-> bool [n_pred, len(iouv)] (val.py:274-288).  Empty predictions or labels give all-False."""
    n, g = len(pred_cls), len(gt_cls)
    tp = np.zeros((n, len(iouv)), dtype=bool)
    if n == 0 or g == 0:
        return tp
    iou = box_iou(np.asarray(gt_boxes), np.asarray(pred_boxes))  # labels x detections
    iou = iou * (np.asarray(gt_cls, dtype=np.float32)[:, None] == np.asarray(pred_cls, dtype=np.float32)[None, :])
    best_l = np.zeros(n, dtype=np.int64)
    best_i = np.zeros(n, dtype=np.float32)
    for d in range(n):
        col = iou[:, d]
        m = col.max()
        best_i[d] = m
        best_l[d] = np.nonzero(col == m)[0].max()  # ties: higher label index
    for ti, t in enumerate(iouv):
        taken = set()
        for d in range(n):  # ascending detection index: the first candidate of a label wins
            if best_i[d] >= t and best_l[d] not in taken:
                taken.add(int(best_l[d]))
                tp[d, ti] = True
    return tp
