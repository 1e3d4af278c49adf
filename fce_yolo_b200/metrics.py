"""Detection metrics from the validation statistics (host, numpy): the step after ``val.ValStats.result()``.

Restates what the reference computes in ``ap_per_class`` / ``compute_ap`` / ``Metric`` (ultralytics/utils/metrics.py:
785-915, 969-1015) for the detect task: per-class precision / recall at the max-F1 confidence, AP at each IoU threshold
(101-point interpolated area under the precision envelope) and their means (mp, mr, mAP50, mAP50-95).  It runs once per
validation on a few thousand rows, so it stays on the host; the per-image matching that feeds it is the GPU
kernel (``fce_match_predictions``).  Plotting and the confusion matrix are outside the path."""
from __future__ import annotations

import numpy as np

_GRID = np.linspace(0.0, 1.0, 1000)     # confidence grid of the P / R / F1 curves
_RECALL_PTS = np.linspace(0.0, 1.0, 101)  # COCO-style interpolation points


def _area_under_envelope(recall: np.ndarray, precision: np.ndarray):
    """AP of one (class, IoU threshold): sentinel-extended curve, monotone precision envelope, 101-point integral."""
    r = np.concatenate(([0.0], recall, [1.0]))
    p = np.concatenate(([1.0], precision, [0.0]))
    p = np.maximum.accumulate(p[::-1])[::-1]
    trapz = getattr(np, "trapezoid", None) or np.trapz
    return float(trapz(np.interp(_RECALL_PTS, r, p), _RECALL_PTS)), r, p


def _box_smooth(y: np.ndarray, frac: float) -> np.ndarray:
    """Moving average over ~2*frac of the curve with edge replication (odd window)."""
    n = round(len(y) * frac * 2) // 2 + 1
    pad = np.ones(n // 2)
    return np.convolve(np.concatenate((pad * y[0], y, pad * y[-1])), np.ones(n) / n, mode="valid")


def detection_metrics(tp, conf, pred_cls, target_cls, eps: float = 1e-16) -> dict:
    """tp [n, n_iou] bool, conf [n], pred_cls [n], target_cls [m] (tensors or arrays; the ``ValStats.result()`` dict
    unpacked).  Returns ``classes`` (those with labels), per-class ``p``, ``r``, ``f1``, ``ap`` [nc, n_iou], the counts
    ``tp`` / ``fp`` at the max-F1 confidence, and the means ``mp``, ``mr``, ``map50``, ``map75``, ``map``."""
    as_np = lambda t: t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)  # noqa: E731
    tp, conf, pred_cls, target_cls = (as_np(t) for t in (tp, conf, pred_cls, target_cls))
    tp = (tp if tp.ndim == 2 else tp.reshape(len(conf), 1)).astype(np.float64)
    order = np.argsort(-conf)
    tp, conf, pred_cls = tp[order], conf[order], pred_cls[order]
    classes, n_labels = np.unique(target_cls, return_counts=True)
    nc, n_iou = len(classes), tp.shape[1]
    ap = np.zeros((nc, n_iou))
    p_curve, r_curve = np.zeros((nc, _GRID.size)), np.zeros((nc, _GRID.size))
    for ci, c in enumerate(classes):
        sel = pred_cls == c
        if not sel.any() or n_labels[ci] == 0:
            continue
        hits = tp[sel].cumsum(0)
        misses = (1.0 - tp[sel]).cumsum(0)
        recall = hits / (n_labels[ci] + eps)
        precision = hits / (hits + misses)
        # curves over confidence at the first IoU threshold (conf decreases along the rows -> negate for np.interp)
        r_curve[ci] = np.interp(-_GRID, -conf[sel], recall[:, 0], left=0)
        p_curve[ci] = np.interp(-_GRID, -conf[sel], precision[:, 0], left=1)
        for j in range(n_iou):
            ap[ci, j] = _area_under_envelope(recall[:, j], precision[:, j])[0]
    f1_curve = 2 * p_curve * r_curve / (p_curve + r_curve + eps)
    best = int(_box_smooth(f1_curve.mean(0), 0.1).argmax()) if nc else 0
    p, r, f1 = p_curve[:, best], r_curve[:, best], f1_curve[:, best]
    n_tp = (r * n_labels).round()
    n_fp = (n_tp / (p + eps) - n_tp).round()
    mean = lambda a: float(a.mean()) if len(a) else 0.0  # noqa: E731
    return {"classes": classes.astype(int), "p": p, "r": r, "f1": f1, "ap": ap, "tp": n_tp, "fp": n_fp,
            "mp": mean(p), "mr": mean(r), "map50": mean(ap[:, 0]) if nc else 0.0,
            "map75": mean(ap[:, 5]) if nc and n_iou > 5 else 0.0, "map": mean(ap) if nc else 0.0,
            "conf_at_best_f1": float(_GRID[best])}
