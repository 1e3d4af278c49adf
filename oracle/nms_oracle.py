"""CPU oracle for ``non_max_suppression`` - TEST INFRASTRUCTURE ONLY (see fce_oracle.py header).

numpy restatement of ultralytics/utils/nms.py:13-166 for the detection case
(rotated=False, end2end=False, labels=()), with the greedy suppression loop in C
(oracle/nms_oracle.c).  Pinned by tests/golden/nms_*.npz which hold the keep indices / class ids
the real reference returned for the same seeded inputs (tests/golden/make_golden.py).
"""
from __future__ import annotations

import ctypes

import numpy as np

from .build_oracle import build

_lib = None


def _nms_core(boxes: np.ndarray, scores: np.ndarray, thr: float) -> np.ndarray:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.fce_oracle_nms.restype = ctypes.c_int
        _lib.fce_oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_double,
                                        ctypes.c_void_p]
    boxes = np.ascontiguousarray(boxes, dtype=np.float32)
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    n = boxes.shape[0]
    keep = np.zeros(max(n, 1), dtype=np.int64)
    k = _lib.fce_oracle_nms(boxes.ctypes.data, scores.ctypes.data, n, float(thr), keep.ctypes.data)
    return keep[:k]


def non_max_suppression(pred: np.ndarray, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                        multi_label=False, max_det=300, nc=0, max_nms=30000, max_wh=7680):
    """pred: [B, 4+nc, A] fp32 (cx,cy,w,h,scores...).  Returns (dets, idxs): per image
    ``[n,6]`` fp32 rows (x1,y1,x2,y2,conf,cls) in descending confidence and int64 anchor indices."""
    pred = np.asarray(pred, dtype=np.float32)
    B = pred.shape[0]
    nc = nc or pred.shape[1] - 4
    conf_f = np.float32(conf_thres)  # torch compares an fp32 tensor with the python scalar cast to fp32
    multi_label = multi_label and nc > 1  # nms.py:82
    dets, idxs = [], []
    for b in range(B):
        x = pred[b].T.copy()  # [A, 4+nc]  (nms.py:84)
        half = x[:, 2:4] / np.float32(2)  # ops.py:236-239 xywh2xyxy
        xy = x[:, 0:2].copy()
        x[:, 0:2] = xy - half
        x[:, 2:4] = xy + half
        ids = np.arange(x.shape[0], dtype=np.int64)
        cand = x[:, 4:4 + nc].max(1) > conf_f  # nms.py:76
        x, ids = x[cand], ids[cand]
        if x.shape[0] == 0:
            dets.append(np.zeros((0, 6), np.float32)); idxs.append(np.zeros((0,), np.int64)); continue
        box, cls = x[:, :4], x[:, 4:4 + nc]
        if multi_label:  # nms.py:114-118: torch.where is row-major (anchor, then class)
            i, j = np.nonzero(cls > conf_f)
            x = np.concatenate([box[i], cls[i, j][:, None], j[:, None].astype(np.float32)], 1)
            ids = ids[i]
        else:  # nms.py:119-124: first maximal class wins (torch.max returns the first index on ties)
            j = cls.argmax(1)
            conf = cls[np.arange(cls.shape[0]), j]
            f = conf > conf_f
            x = np.concatenate([box, conf[:, None], j[:, None].astype(np.float32)], 1)[f]
            ids = ids[f]
        if classes is not None:  # nms.py:127-131
            f = np.isin(x[:, 5], np.asarray(classes, dtype=np.float32))
            x, ids = x[f], ids[f]
        n = x.shape[0]
        if n == 0:
            dets.append(np.zeros((0, 6), np.float32)); idxs.append(np.zeros((0,), np.int64)); continue
        if n > max_nms:  # nms.py:136-140 (stable here; torch's argsort is unspecified on ties)
            o = np.argsort(-x[:, 4], kind="stable")[:max_nms]
            x, ids = x[o], ids[o]
        c = x[:, 5:6] * np.float32(0 if agnostic else max_wh)  # nms.py:143
        boxes = x[:, :4] + c  # fp32 add, nms.py:149
        keep = _nms_core(boxes, x[:, 4], iou_thres)[:max_det]  # nms.py:154-157
        dets.append(x[keep].astype(np.float32))
        idxs.append(ids[keep])
    return dets, idxs
