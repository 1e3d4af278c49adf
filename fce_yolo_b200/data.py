"""Minimal validation-set reader for the detect path: a YOLO-format dataset (``data.yaml`` with ``path`` / ``val`` /
``names``, images under ``images/<split>``, one ``labels/<split>/<stem>.txt`` per image with rows
``cls cx cy w h`` normalised to the image - the layout ``check_det_dataset`` expects, ultralytics/data/utils.py:385-470)
-> batches of (BGR frames, pixel-space labels) for ``YOLO.val``.

Deliberately small: no caching, augmentation, rect batching or multiprocessing (the reference's dataset / dataloader
stack stays with the reference).  Labels come out as ``[n, 5]`` float32 ``(cls, x1, y1, x2, y2)`` in ORIGINAL image
pixels - the space in which ``Predictor.predict`` returns its boxes (``ops.scale_boxes``), so no letterbox bookkeeping
is needed on the label side (the reference scales labels into the letterboxed space and the predictions back,
models/yolo/detect/val.py:147-165; same comparison, one conversion less)."""
from __future__ import annotations

import glob
import os

import numpy as np
import yaml

IMG_EXT = (".bmp", ".jpeg", ".jpg", ".png", ".tif", ".tiff", ".webp")


def load_data_yaml(path: str) -> dict:
    with open(path) as f:
        d = yaml.safe_load(f)
    for key in ("val", "names"):
        if key not in d:
            raise KeyError(f"{path}: '{key}' is required (data/utils.py:413-431)")
    if isinstance(d["names"], (list, tuple)):
        d["names"] = dict(enumerate(d["names"]))
    root = d.get("path") or os.path.dirname(os.path.abspath(path))
    d["path"] = root if os.path.isabs(root) else os.path.join(os.path.dirname(os.path.abspath(path)), root)
    return d


def image_files(split_dir: str):
    files = sorted(p for p in glob.glob(os.path.join(split_dir, "**", "*"), recursive=True)
                   if p.lower().endswith(IMG_EXT))
    if not files:
        raise FileNotFoundError(f"no images under {split_dir}")
    return files


def label_path(img_path: str) -> str:
    """.../images/<split>/x.jpg -> .../labels/<split>/x.txt (data/utils.py:46-49 img2label_paths)."""
    sa, sb = f"{os.sep}images{os.sep}", f"{os.sep}labels{os.sep}"
    head, _, tail = img_path.rpartition(sa)
    return (head + sb + tail if head else img_path).rsplit(".", 1)[0] + ".txt"


def read_labels(txt: str, height: int, width: int) -> np.ndarray:
    """Rows ``cls cx cy w h`` (normalised) -> [n, 5] float32 (cls, x1, y1, x2, y2) in pixels; a missing or empty file
    is an image without objects (background), like the reference."""
    if not os.path.isfile(txt):
        return np.zeros((0, 5), np.float32)
    rows = [ln.split() for ln in open(txt).read().strip().splitlines() if ln.strip()]
    if not rows:
        return np.zeros((0, 5), np.float32)
    a = np.asarray(rows, dtype=np.float32)
    if a.shape[1] != 5:
        raise ValueError(f"{txt}: expected 5 columns (cls cx cy w h), got {a.shape[1]} (segments / keypoints are outside the path)")
    if (a[:, 1:] < 0).any() or (a[:, 1:] > 1 + 1e-4).any():
        raise ValueError(f"{txt}: non-normalised or out-of-bounds coordinates")
    out = np.empty_like(a)
    out[:, 0] = a[:, 0]
    out[:, 1] = (a[:, 1] - a[:, 3] / 2) * width
    out[:, 2] = (a[:, 2] - a[:, 4] / 2) * height
    out[:, 3] = (a[:, 1] + a[:, 3] / 2) * width
    out[:, 4] = (a[:, 2] + a[:, 4] / 2) * height
    return out


def iter_val_batches(data, batch: int = 16, split: str = "val"):
    """Yields (frames, labels) per batch: ``frames`` uint8 HWC BGR arrays as cv2 decodes them, ``labels`` as above.
    ``data``: path of a data.yaml, or a dict already loaded by ``load_data_yaml``."""
    import cv2

    d = load_data_yaml(data) if isinstance(data, (str, os.PathLike)) else data
    files = image_files(os.path.join(d["path"], d[split]))
    for i in range(0, len(files), batch):
        frames, labels = [], []
        for f in files[i:i + batch]:
            im = cv2.imread(f, cv2.IMREAD_COLOR)
            if im is None:
                raise OSError(f"cannot decode {f}")
            frames.append(im)
            labels.append(read_labels(label_path(f), im.shape[0], im.shape[1]))
        yield frames, labels
