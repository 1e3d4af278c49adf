"""YOLO facade (fce_yolo_b200.model): construction, argument handling, batching and predictor caching on CPU with a
fake predictor (the real one needs a B200 and must say so), packed save / reload.  The GPU behaviour of every piece it
composes is covered elsewhere (Predictor.predict, ValStats, fce_match_predictions)."""
import numpy as np
import pytest
import torch

from fce_yolo_b200 import YOLO
from fce_yolo_b200.results import Results


class FakePredictor:
    made = []

    def __init__(self, model, batch, imgsz, **kw):
        self.batch, self.imgsz, self.kw, self.calls = batch, imgsz, kw, []
        FakePredictor.made.append(self)

    def predict(self, images, as_results=False, names=None, paths=None):
        assert 0 < len(images) <= self.batch
        self.calls.append(len(images))
        dets = [torch.tensor([[1.0, 2.0, 3.0 + i, 4.0, 0.5, 7.0]]) for i in range(len(images))]
        if not as_results:
            return dets
        return [Results(im.shape[:2], d, names=names, path=paths[i], orig_img=im) for i, (im, d) in enumerate(zip(images, dets))]


def _frames(n, h=48, w=64):
    return [np.full((h, w, 3), i, dtype=np.uint8) for i in range(n)]


def test_construction_and_rejected_inputs(tmp_path):
    m = YOLO("yolo11n-fce.yaml")
    assert m.task == "detect" and len(m.names) == 80 and m.info()["layers"] == 26
    assert not any(isinstance(x, torch.nn.BatchNorm2d) for x in m.model.modules())  # fused like AutoBackend does
    assert YOLO(m.model).model is m.model
    with pytest.raises(ValueError):
        YOLO("weights.pt")
    with pytest.raises(FileNotFoundError):
        YOLO("no-such-model.yaml")
    with pytest.raises(ValueError):
        m.predict(_frames(1), imgsz=100)          # not a multiple of the max stride
    with pytest.raises(TypeError):
        m.predict([np.zeros((8, 8, 3), np.float32)])  # not uint8 frames
    assert m.predict([]) == []
    p = tmp_path / "n.fcepack"
    m.save(str(p))
    again = YOLO(str(p))
    sd0, sd1 = m.model.state_dict(), again.model.state_dict()
    assert sd0.keys() == sd1.keys()
    for k in sd0:  # bf16 storage: one rounding of the fused fp32 weights
        assert torch.allclose(sd0[k].float(), sd1[k].float(), rtol=2 ** -7, atol=1e-6), k
    with pytest.raises(ValueError):
        m.save(str(tmp_path / "n.pt"))


def test_predict_batches_caches_and_wraps():
    FakePredictor.made.clear()
    m = YOLO("yolo11n-fce.yaml")
    m._predictor_cls = FakePredictor
    res = m.predict(_frames(5), imgsz=(64, 96), conf=0.3, batch=2)
    assert len(res) == 5 and all(isinstance(r, Results) for r in res)
    (p,) = FakePredictor.made
    assert (p.batch, p.imgsz, p.calls) == (2, (64, 96), [2, 2, 1])
    assert p.kw["conf"] == 0.3 and p.kw["iou"] == 0.7 and p.kw["max_det"] == 300 and p.kw["input_u8"] is True
    assert p.kw["multi_label"] is False and p.kw["precision"] == "bf16"
    assert [r.path for r in res] == [f"image{i}.jpg" for i in range(5)] and res[3].orig_shape == (48, 64)
    assert res[0].names is m.names and res[4].summary()[0]["class"] == 7
    m.predict(_frames(2), imgsz=(64, 96), conf=0.3, batch=2)      # same key: the compiled plan is reused
    assert len(FakePredictor.made) == 1 and p.calls == [2, 2, 1, 2]
    m(_frames(1), imgsz=64)                                        # __call__ = predict; new shape -> new plan
    assert len(FakePredictor.made) == 2 and FakePredictor.made[1].batch == 1
    single = m.predict(_frames(1)[0], imgsz=64)                    # one HWC array is one image
    assert len(single) == 1


def test_first_predict_needs_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        YOLO("yolo11n-fce.yaml").predict(_frames(1), imgsz=64)
