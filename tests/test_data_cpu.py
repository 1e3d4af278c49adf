"""data.iter_val_batches: YOLO-format dataset -> (frames, pixel labels) batches, on a synthetic dataset written with
cv2 (lossless PNG).  Label conversion is checked against the reference's own xywhn2xyxy when the reference tree is
present."""
import os
import sys

import numpy as np
import pytest

from fce_yolo_b200.data import iter_val_batches, label_path, load_data_yaml, read_labels

REF = "/root/reference"


def _make(tmp_path, n=5):
    import cv2

    root = tmp_path / "ds"
    (root / "images" / "val").mkdir(parents=True)
    (root / "labels" / "val").mkdir(parents=True)
    g = np.random.default_rng(0)
    truth = []
    for i in range(n):
        h, w = int(g.integers(40, 90)), int(g.integers(50, 120))
        img = g.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        cv2.imwrite(str(root / "images" / "val" / f"im{i}.png"), img)
        k = i % 3  # image 0 and 3: background (no label file / empty file)
        rows = np.column_stack([g.integers(0, 80, k), g.uniform(0.3, 0.7, k), g.uniform(0.3, 0.7, k),
                                g.uniform(0.05, 0.5, k), g.uniform(0.05, 0.5, k)]) if k else np.zeros((0, 5))
        if i != 0:
            np.savetxt(root / "labels" / "val" / f"im{i}.txt", rows, fmt="%.6f")
        truth.append((img, rows, h, w))
    (root / "data.yaml").write_text("path: .\nval: images/val\nnames:\n" + "".join(f"  {i}: c{i}\n" for i in range(80)))
    return root, truth


def test_batches_frames_and_labels(tmp_path):
    root, truth = _make(tmp_path)
    d = load_data_yaml(str(root / "data.yaml"))
    assert len(d["names"]) == 80 and os.path.isabs(d["path"])
    batches = list(iter_val_batches(str(root / "data.yaml"), batch=2))
    assert [len(f) for f, _ in batches] == [2, 2, 1]
    flat = [(f, l) for fs, ls in batches for f, l in zip(fs, ls)]
    for (img, lab), (t_img, t_rows, h, w) in zip(flat, truth):
        assert img.dtype == np.uint8 and np.array_equal(img, t_img)   # BGR exactly as written
        assert lab.dtype == np.float32 and lab.shape == (len(t_rows), 5)
        if len(t_rows):
            r = t_rows
            exp = np.column_stack([r[:, 0], (r[:, 1] - r[:, 3] / 2) * w, (r[:, 2] - r[:, 4] / 2) * h,
                                   (r[:, 1] + r[:, 3] / 2) * w, (r[:, 2] + r[:, 4] / 2) * h])
            assert np.allclose(lab, exp, atol=1e-3)
    assert label_path("/a/images/val/x.jpg") == "/a/labels/val/x.txt"
    with pytest.raises(ValueError):
        bad = root / "labels" / "val" / "bad.txt"
        bad.write_text("1 0.5 0.5 2.0 0.1\n")
        read_labels(str(bad), 10, 10)
    with pytest.raises(KeyError):
        (root / "broken.yaml").write_text("path: .\nnames: [a]\n")
        load_data_yaml(str(root / "broken.yaml"))


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ultralytics")), reason="reference tree not present")
def test_label_conversion_matches_reference(tmp_path):
    os.environ.setdefault("YOLO_CONFIG_DIR", "/tmp/ulcfg")
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.append(REF)
    from ultralytics.utils.ops import xywhn2xyxy

    g = np.random.default_rng(1)
    rows = np.column_stack([g.integers(0, 80, 6), g.uniform(0.3, 0.7, 6), g.uniform(0.3, 0.7, 6),
                            g.uniform(0.05, 0.5, 6), g.uniform(0.05, 0.5, 6)]).astype(np.float32)
    p = tmp_path / "l.txt"
    np.savetxt(p, rows, fmt="%.8f")
    mine = read_labels(str(p), 375, 500)
    ref = xywhn2xyxy(np.loadtxt(p, dtype=np.float32)[:, 1:], w=500, h=375)
    assert np.allclose(mine[:, 1:], ref, atol=1e-3) and np.array_equal(mine[:, 0], rows[:, 0])
