"""Host-side wrappers of the detections, mirroring the part of the reference's result objects that the detect
predict path fills in (ultralytics/engine/results.py:815-1008 ``Boxes``; :188-330, 749-800 ``Results``), so that code
written against ``YOLO(...).predict(...)`` keeps reading ``r.boxes.xyxy / .conf / .cls / .xywh / .xyxyn / .xywhn``,
``len(r)``, ``r.summary()`` unchanged.  Pure tensor views: no copies until a derived format is asked for.

Out of scope (not on the detect path): masks, keypoints, OBB, probs, tracking ids, plotting and saving."""
from __future__ import annotations

import numpy as np
import torch


class Boxes:
    """Rows are (x1, y1, x2, y2, conf, cls) in ORIGINAL image pixels; ``orig_shape`` = (height, width)."""

    def __init__(self, data, orig_shape):
        if data.ndim == 1:
            data = data[None, :]
        if data.shape[-1] != 6:  # the reference also takes 7 columns (track id): tracking is outside the path
            raise AssertionError(f"expected 6 values per box but got {data.shape[-1]}")
        self.data = data
        self.orig_shape = tuple(int(v) for v in orig_shape)
        self.is_track = False

    # -- raw columns -----------------------------------------------------------------------------------
    @property
    def xyxy(self):
        return self.data[:, :4]

    @property
    def conf(self):
        return self.data[:, -2]

    @property
    def cls(self):
        return self.data[:, -1]

    @property
    def id(self):
        return None

    # -- derived formats (ops.xyxy2xywh, results.py:945-1007) --------------------------------------------
    def _like(self):
        return torch.empty_like(self.xyxy) if isinstance(self.data, torch.Tensor) else np.empty_like(self.xyxy)

    @property
    def xywh(self):
        b, y = self.xyxy, self._like()
        y[..., 0] = (b[..., 0] + b[..., 2]) / 2
        y[..., 1] = (b[..., 1] + b[..., 3]) / 2
        y[..., 2] = b[..., 2] - b[..., 0]
        y[..., 3] = b[..., 3] - b[..., 1]
        return y

    def _normalised(self, t):
        t[..., [0, 2]] /= self.orig_shape[1]
        t[..., [1, 3]] /= self.orig_shape[0]
        return t

    @property
    def xyxyn(self):
        b = self.xyxy
        return self._normalised(b.clone() if isinstance(b, torch.Tensor) else np.copy(b))

    @property
    def xywhn(self):
        return self._normalised(self.xywh)

    # -- container protocol / device moves (results.py BaseTensor :30-185) --------------------------------
    @property
    def shape(self):
        return self.data.shape

    def __len__(self):
        return len(self.data)

    def __getitem__(self, idx):
        return Boxes(self.data[idx], self.orig_shape)

    def cpu(self):
        return self if isinstance(self.data, np.ndarray) else Boxes(self.data.cpu(), self.orig_shape)

    def numpy(self):
        return self if isinstance(self.data, np.ndarray) else Boxes(self.data.cpu().numpy(), self.orig_shape)

    def cuda(self):
        return Boxes(torch.as_tensor(self.data).cuda(), self.orig_shape)

    def to(self, *args, **kwargs):
        return Boxes(torch.as_tensor(self.data).to(*args, **kwargs), self.orig_shape)

    def __repr__(self):
        return f"Boxes(n={len(self)}, orig_shape={self.orig_shape})"


class Results:
    """One image's detections (results.py:188-330): ``boxes``, ``names``, ``orig_shape``, ``path``, ``speed``."""

    def __init__(self, orig_shape, boxes, names=None, path="", orig_img=None, speed=None):
        self.orig_img = orig_img
        self.orig_shape = tuple(int(v) for v in orig_shape)
        self.boxes = Boxes(boxes, self.orig_shape) if boxes is not None else None
        self.masks = self.probs = self.keypoints = self.obb = None
        self.names = names if names is not None else {}
        self.path = path
        self.speed = speed or {"preprocess": None, "inference": None, "postprocess": None}

    def __len__(self):
        return len(self.boxes) if self.boxes is not None else 0

    def __getitem__(self, idx):
        r = Results(self.orig_shape, None, self.names, self.path, self.orig_img, self.speed)
        r.boxes = self.boxes[idx]
        return r

    def cpu(self):
        r = Results(self.orig_shape, None, self.names, self.path, self.orig_img, self.speed)
        r.boxes = self.boxes.cpu()
        return r

    def numpy(self):
        r = Results(self.orig_shape, None, self.names, self.path, self.orig_img, self.speed)
        r.boxes = self.boxes.numpy()
        return r

    def summary(self, normalize: bool = False, decimals: int = 5):
        """List of ``{"name", "class", "confidence", "box": {x1, y1, x2, y2}}`` dicts (results.py:749-800)."""
        h, w = self.orig_shape if normalize else (1, 1)
        out = []
        rows = self.boxes.data.tolist() if self.boxes is not None else []
        for x1, y1, x2, y2, conf, cls in rows:
            cid = int(cls)
            out.append({"name": self.names[cid] if cid in self.names or isinstance(self.names, (list, tuple)) else str(cid),
                        "class": cid, "confidence": round(conf, decimals),
                        "box": {"x1": round(x1 / w, decimals), "y1": round(y1 / h, decimals),
                                "x2": round(x2 / w, decimals), "y2": round(y2 / h, decimals)}})
        return out

    def __repr__(self):
        return f"Results(n={len(self)}, orig_shape={self.orig_shape}, path={self.path!r})"
