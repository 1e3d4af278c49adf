"""results.Boxes / results.Results: the host wrappers returned by Predictor.predict(..., as_results=True) must give the
same numbers as the reference's Boxes / Results accessors (ultralytics/engine/results.py:815-1008, 749-800).  The
formulas are checked on fixed values everywhere; the comparison with the LIVE reference runs in the build container only
(/root/reference is absent on the GPU box)."""
import os
import sys

import numpy as np
import pytest
import torch

from fce_yolo_b200.results import Boxes, Results

REF = "/root/reference"


def _dets(n=7, seed=0, h=480, w=640):
    g = torch.Generator().manual_seed(seed)
    x1 = torch.rand(n, generator=g) * (w - 50)
    y1 = torch.rand(n, generator=g) * (h - 50)
    bw = torch.rand(n, generator=g) * 45 + 2
    bh = torch.rand(n, generator=g) * 45 + 2
    conf = torch.rand(n, generator=g)
    cls = torch.randint(0, 80, (n,), generator=g).float()
    return torch.stack([x1, y1, x1 + bw, y1 + bh, conf, cls], 1)


def test_boxes_accessors_on_fixed_values():
    d = torch.tensor([[100.0, 50.0, 150.0, 100.0, 0.9, 0.0], [200.0, 150.0, 300.0, 250.0, 0.8, 1.0]])
    b = Boxes(d, (480, 640))
    assert len(b) == 2 and b.shape == (2, 6) and b.id is None and not b.is_track
    assert torch.equal(b.xyxy, d[:, :4]) and torch.equal(b.conf, d[:, 4]) and torch.equal(b.cls, d[:, 5])
    assert torch.equal(b.xywh, torch.tensor([[125.0, 75.0, 50.0, 50.0], [250.0, 200.0, 100.0, 100.0]]))
    assert torch.allclose(b.xyxyn, d[:, :4] / torch.tensor([640.0, 480.0, 640.0, 480.0]))
    assert torch.allclose(b.xywhn, b.xywh / torch.tensor([640.0, 480.0, 640.0, 480.0]))
    assert torch.equal(b.xyxy, d[:, :4])  # the derived formats never write through to the data
    one = b[1]
    assert len(one) == 1 and torch.equal(one.xyxy[0], d[1, :4])
    n = b.numpy()
    assert isinstance(n.data, np.ndarray) and np.allclose(n.xywhn, b.xywhn.numpy())
    assert Boxes(d[0], (480, 640)).shape == (1, 6)  # a single row is promoted like the reference does
    with pytest.raises(AssertionError):
        Boxes(torch.zeros(3, 5), (10, 10))


def test_results_summary_and_empty():
    d = torch.tensor([[10.0, 20.0, 110.0, 220.0, 0.87654321, 3.0]])
    r = Results((400, 200), d, names={i: f"c{i}" for i in range(80)}, path="a.jpg")
    assert len(r) == 1 and r.orig_shape == (400, 200)
    assert r.summary() == [{"name": "c3", "class": 3, "confidence": 0.87654,
                            "box": {"x1": 10.0, "y1": 20.0, "x2": 110.0, "y2": 220.0}}]
    assert r.summary(normalize=True, decimals=3)[0]["box"] == {"x1": 0.05, "y1": 0.05, "x2": 0.55, "y2": 0.55}
    e = Results((400, 200), torch.zeros(0, 6), names={})
    assert len(e) == 0 and e.summary() == [] and e.boxes.xywhn.shape == (0, 4)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ultralytics")), reason="reference tree not present")
@pytest.mark.parametrize("seed,shape", [(0, (480, 640)), (1, (1080, 1920)), (2, (33, 77))])
def test_against_live_reference(seed, shape):
    os.environ.setdefault("YOLO_CONFIG_DIR", "/tmp/ulcfg")
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.append(REF)
    from ultralytics.engine.results import Boxes as RefBoxes
    from ultralytics.engine.results import Results as RefResults

    d = _dets(9, seed, *shape)
    mine, ref = Boxes(d.clone(), shape), RefBoxes(d.clone(), shape)
    for attr in ("xyxy", "conf", "cls", "xywh", "xyxyn", "xywhn"):
        assert torch.equal(getattr(mine, attr), getattr(ref, attr)), attr
    names = {i: f"n{i}" for i in range(80)}
    img = np.zeros(shape + (3,), dtype=np.uint8)
    rr = RefResults(img, path="p.jpg", names=names, boxes=d.clone())
    mr = Results(shape, d.clone(), names=names, path="p.jpg", orig_img=img)
    assert len(mr) == len(rr) and mr.orig_shape == tuple(rr.orig_shape)
    for norm in (False, True):
        assert mr.summary(normalize=norm) == rr.summary(normalize=norm)
