"""YOLO facade on the GPU (runs last): predict() must return exactly what Predictor.predict returns, wrapped; val()
must turn the model's own detections, used as labels, into non-trivial metrics through the GPU matcher."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_facade_predict_and_val():
    from fce_yolo_b200 import YOLO
    from fce_yolo_b200.predict import Predictor
    from fce_yolo_b200.results import Results
    from fce_yolo_b200.weights import load_synthetic

    model = YOLO("yolo11n-fce.yaml")
    load_synthetic(model.model, 0)
    g = np.random.default_rng(3)
    frames = [g.integers(0, 256, size=s + (3,), dtype=np.uint8) for s in ((480, 640), (300, 500), (640, 640))]
    res = model.predict(frames, imgsz=640, conf=0.05)
    assert len(res) == 3 and all(isinstance(r, Results) for r in res)
    direct = Predictor(model.model, 3, 640, precision="bf16", conf=0.05).predict(frames)
    for r, d, im in zip(res, direct, frames):
        assert r.orig_shape == im.shape[:2] and torch.equal(r.boxes.data, d)
        if len(r):
            b = r.boxes
            assert float(b.xyxyn.max()) <= 1.0 + 1e-6 and float(b.xyxy.min()) >= 0.0
            assert r.summary()[0]["class"] == int(b.cls[0])
    assert sum(len(r) for r in res) > 0
    # the model's own detections as ground truth: every label has an identical prediction unless the validator's
    # multi-label NMS suppresses it, so AP@0.5 cannot be zero
    labels = [np.concatenate([r.boxes.cls.numpy()[:, None], r.boxes.xyxy.numpy()], 1) for r in res]
    stats = model.val([(frames, labels)], imgsz=640)
    assert stats is not None and stats["ap"].shape[1] == 10
    assert 0.0 < stats["map50"] <= 1.0 and 0.0 <= stats["map"] <= stats["map50"] + 1e-9
    assert set(stats["classes"].tolist()) == set(int(c) for l in labels for c in l[:, 0])
