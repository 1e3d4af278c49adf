"""GPU preprocessing in front of the predict path.

Mirror of ``BasePredictor.pre_transform`` + ``preprocess`` (ultralytics/engine/predictor.py:151-201) for lists of
uint8 HWC BGR images: LetterBox (ultralytics/data/augment.py:1509-1647) -> BGR->RGB -> the network's uint8 NHWC
input (the /255 of predictor.py:171 is folded into the stem weights).  On real inputs this step is CPU-bound in the
reference (cv2.resize + copyMakeBorder per image); here one H2D copy ships the raw images and ``fce_letterbox``
resizes / pads / flips the whole batch in one launch, bit-exactly.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L

PAD_VALUE = 114


def letterbox_geometry(shape, new_shape=(640, 640), auto=False, scaleup=True, center=True, stride=32):
    """Sizes and paddings of LetterBox for a source of ``shape`` = (h, w) (augment.py:1592-1621)."""
    if isinstance(new_shape, int):
        new_shape = (new_shape, new_shape)
    h, w = int(shape[0]), int(shape[1])
    r = min(new_shape[0] / h, new_shape[1] / w)
    if not scaleup:
        r = min(r, 1.0)
    new_w, new_h = round(w * r), round(h * r)
    dw, dh = new_shape[1] - new_w, new_shape[0] - new_h
    if auto:
        dw, dh = dw % stride, dh % stride
    if center:
        dw, dh = dw / 2, dh / 2
    top, bottom = (round(dh - 0.1) if center else 0), round(dh + 0.1)
    left, right = (round(dw - 0.1) if center else 0), round(dw + 0.1)
    return new_w, new_h, top, left, new_h + top + bottom, new_w + left + right


def _taps(src: int, dst: int, clamp_fraction: bool) -> np.ndarray:
    """[dst, 4] int32 {i0, i1, w0, w1}: OpenCV's linear-resize taps along one axis (11-bit weights).  The x axis clamps
    the fraction at the border, the y axis only the row index."""
    scale = 1.0 / (dst / src)
    f = ((np.arange(dst, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_fraction:
        lo, hi = s < 0, s >= src - 1
        f[lo | hi] = 0
        s[lo] = 0
        s[hi] = src - 1
    t = np.empty((dst, 4), dtype=np.int32)
    t[:, 0] = np.clip(s, 0, src - 1)
    t[:, 1] = np.clip(s + 1, 0, src - 1)
    t[:, 2] = np.rint((np.float32(1) - f) * np.float32(2048))
    t[:, 3] = np.rint(f * np.float32(2048))
    return t


class LetterBoxGPU:
    """``LetterBoxGPU(640)(list_of_bgr_uint8_images) -> uint8 CUDA tensor [B, out_h, out_w, 3] (RGB)``.
    All images of a call must letterbox to the same output size (always true with ``auto=False``; with ``auto=True``
    the reference also requires same-shaped sources, predictor.py:193-199)."""

    def __init__(self, new_shape=640, auto: bool = False, scaleup: bool = True, center: bool = True, stride: int = 32,
                 device=None):
        self.new_shape = (new_shape, new_shape) if isinstance(new_shape, int) else tuple(new_shape)
        self.auto, self.scaleup, self.center, self.stride = auto, scaleup, center, stride
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise RuntimeError("fce_yolo_b200 preprocessing runs on the GPU only (no CPU fallback)")
        self.lib = L.load(check_device=True)
        self._pin = self._dev = None  # staging: [meta int32 | images uint8], grown on demand
        self._h2d_done = None         # event behind the last host->device copy out of the pinned staging buffer

    def _staging(self, nbytes: int):
        if self._pin is None or self._pin.numel() < nbytes:
            n = int(nbytes * 1.25) + 4096
            self._pin = torch.empty(n, dtype=torch.uint8).pin_memory()
            self._dev = torch.empty(n, dtype=torch.uint8, device=self.device)
        return self._pin, self._dev

    def __call__(self, images, out: torch.Tensor | None = None) -> torch.Tensor:
        if not len(images):
            raise ValueError("empty batch")
        B = len(images)
        geo = []
        for im in images:
            if not (isinstance(im, np.ndarray) and im.dtype == np.uint8 and im.ndim == 3 and im.shape[2] == 3):
                raise ValueError("images must be uint8 HWC arrays with 3 channels (BGR)")
            geo.append(letterbox_geometry(im.shape[:2], self.new_shape, self.auto, self.scaleup, self.center, self.stride))
        out_h, out_w = geo[0][4], geo[0][5]
        if any((g[4], g[5]) != (out_h, out_w) for g in geo):
            raise ValueError("images of one batch letterbox to different sizes (use auto=False or same-shaped sources)")
        # staging layout (bytes): items [B x 40] | pad to 16 | xtab [B, out_w, 4] int32 | ytab [B, out_h, 4] int32 | images
        item_sz = C.sizeof(L.LetterboxItem)
        off_x = (B * item_sz + 15) & ~15
        off_y = off_x + B * out_w * 16
        off_img = off_y + B * out_h * 16
        img_offs, cur = [], off_img
        for im in images:
            img_offs.append(cur)
            cur = (cur + im.shape[0] * im.shape[1] * 3 + 15) & ~15
        if self._h2d_done is not None:
            # the previous call's non_blocking copy may still be reading the pinned buffer: rewriting it now would corrupt
            # that batch (two back-to-back standalone calls) - wait for the DMA, not for the whole stream
            self._h2d_done.synchronize()
        pin, dev = self._staging(cur)
        pin_np = pin.numpy()
        base = dev.data_ptr()
        items = (L.LetterboxItem * B)()
        xt = pin_np[off_x:off_y].view(np.int32).reshape(B, out_w, 4)
        yt = pin_np[off_y:off_img].view(np.int32).reshape(B, out_h, 4)
        for b, (im, g, o) in enumerate(zip(images, geo, img_offs)):
            new_w, new_h, top, left = g[:4]
            h, w = im.shape[:2]
            items[b] = L.LetterboxItem(src=base + o, src_pitch=w * 3, H=h, W=w, new_w=new_w, new_h=new_h, top=top, left=left,
                                       reserved=0)
            xt[b, :new_w] = _taps(w, new_w, True)
            yt[b, :new_h] = _taps(h, new_h, False)
            pin_np[o:o + h * w * 3] = np.ascontiguousarray(im).reshape(-1)
        pin_np[:B * item_sz] = np.frombuffer(bytes(items), dtype=np.uint8)
        if out is None:
            out = torch.empty(B, out_h, out_w, 3, dtype=torch.uint8, device=self.device)
        elif tuple(out.shape) != (B, out_h, out_w, 3) or out.dtype != torch.uint8 or not out.is_contiguous():
            raise ValueError(f"out must be a contiguous uint8 tensor of shape {(B, out_h, out_w, 3)}")
        with torch.cuda.device(self.device):
            dev[:cur].copy_(pin[:cur], non_blocking=True)
            if self._h2d_done is None:
                self._h2d_done = torch.cuda.Event()
            self._h2d_done.record(torch.cuda.current_stream(self.device))
            st = self.lib.fce_letterbox(C.c_void_p(base), C.c_void_p(base + off_x), C.c_void_p(base + off_y), B, out_h,
                                        out_w, PAD_VALUE, C.c_void_p(out.data_ptr()),
                                        C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        L.check(st, "fce_letterbox")
        return out
