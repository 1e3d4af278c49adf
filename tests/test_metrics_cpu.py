"""metrics.detection_metrics (host numpy) against the LIVE reference's ap_per_class / Metric
(ultralytics/utils/metrics.py:785-915, 969-1015) on synthetic validation statistics, plus closed-form cases that run
everywhere (/root/reference is absent on the GPU box)."""
import os
import sys

import numpy as np
import pytest

from fce_yolo_b200.metrics import detection_metrics

REF = "/root/reference"


def _stats(seed, n=400, m=150, nc=7, n_iou=10):
    g = np.random.default_rng(seed)
    conf = g.random(n).astype(np.float32)
    pred_cls = g.integers(0, nc, n).astype(np.float32)
    target_cls = g.integers(0, nc - 1, m).astype(np.float32)   # the last class has predictions but no labels
    base = g.random(n) < 0.2 + 0.6 * conf                      # confident predictions are right more often
    tp = np.stack([base & (g.random(n) < 1.0 - 0.08 * j) for j in range(n_iou)], 1)
    tp = np.logical_and.accumulate(tp, 1)                       # a match at a strict IoU is a match at a looser one
    return tp, conf, pred_cls, target_cls


def test_perfect_and_empty_predictions():
    cls = np.array([0.0, 0.0, 1.0, 2.0])
    out = detection_metrics(np.ones((4, 10), bool), np.array([0.9, 0.8, 0.7, 0.6]), cls, cls)
    assert list(out["classes"]) == [0, 1, 2]
    assert np.allclose(out["ap"], 0.995, atol=1e-3) and abs(out["map"] - 0.995) < 1e-3  # 101-point interpolation
    assert np.allclose(out["r"], 1.0) and np.allclose(out["p"], 1.0) and list(out["tp"]) == [2, 1, 1] and not out["fp"].any()
    none = detection_metrics(np.zeros((0, 10), bool), np.zeros(0), np.zeros(0), cls)
    assert none["map"] == 0.0 and none["mp"] == 0.0 and none["ap"].shape == (3, 10)
    wrong = detection_metrics(np.zeros((3, 10), bool), np.array([0.9, 0.5, 0.4]), np.array([0.0, 1.0, 1.0]), cls)
    assert wrong["map50"] == 0.0 and not wrong["tp"].any()


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ultralytics")), reason="reference tree not present")
@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_against_live_reference(seed):
    os.environ.setdefault("YOLO_CONFIG_DIR", "/tmp/ulcfg")
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.append(REF)
    from ultralytics.utils.metrics import ap_per_class

    tp, conf, pred_cls, target_cls = _stats(seed)
    r_tp, r_fp, r_p, r_r, r_f1, r_ap, r_classes, *_ = ap_per_class(tp, conf, pred_cls, target_cls, plot=False)
    mine = detection_metrics(tp, conf, pred_cls, target_cls)
    assert np.array_equal(mine["classes"], r_classes)
    for k, ref in (("ap", r_ap), ("p", r_p), ("r", r_r), ("f1", r_f1), ("tp", r_tp), ("fp", r_fp)):
        assert np.allclose(mine[k], ref, rtol=0, atol=1e-12), k
    assert abs(mine["mp"] - r_p.mean()) < 1e-12 and abs(mine["mr"] - r_r.mean()) < 1e-12
    assert abs(mine["map50"] - r_ap[:, 0].mean()) < 1e-12 and abs(mine["map"] - r_ap.mean()) < 1e-12
