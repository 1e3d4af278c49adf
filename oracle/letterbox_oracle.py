"""CPU oracle for the predict-side preprocessing - TEST INFRASTRUCTURE ONLY (see fce_oracle.py).

Restates, in numpy integer arithmetic,
  * ``LetterBox.__call__`` (ultralytics/data/augment.py:1589-1631): scale ratio, rounded unpadded size, centred /
    stride-modulo padding, resize if needed, constant border 114;
  * the BGR->RGB flip of ``BasePredictor.preprocess`` (ultralytics/engine/predictor.py:163-165);
  * ``cv2.resize(..., INTER_LINEAR)`` for 8-bit images.  OpenCV is a third-party dependency of the reference
    (``opencv-python>=4.6.0``, pyproject.toml; 4.13.0 installed in the build container), its source is not under
    /root/reference, so the published algorithm (modules/imgproc/src/resize.cpp: ``resizeGeneric_`` with
    ``HResizeLinear`` / ``VResizeLinear<uchar,int,short,FixedPtCast<int,uchar,22>,VResizeLinearVec_32s8u>``) is
    restated: 11-bit fixed-point coefficients ``cvRound(w * 2048)``; horizontal taps clamp the FRACTION at the image
    border (fx = 0), vertical taps clamp only the ROW INDEX (both taps read the border row with their un-clamped
    weights); the vertical pass is ``(((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2``.

Pinned: bit-exact against cv2 4.13 itself on random images (tests/golden/make_golden.py stores the reference's own
LetterBox output; tests/test_oracle_golden.py replays it).
"""
from __future__ import annotations

import numpy as np

PAD_VALUE = 114  # augment.py:1537 padding_value


def letterbox_geometry(shape, new_shape=(640, 640), auto=False, scaleup=True, center=True, stride=32):
    """augment.py:1592-1621.  shape = (h, w) of the source.  Returns dict(new_w, new_h, top, bottom, left, right,
    out_h, out_w)."""
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
        dw /= 2
        dh /= 2
    top, bottom = (round(dh - 0.1) if center else 0), round(dh + 0.1)
    left, right = (round(dw - 0.1) if center else 0), round(dw + 0.1)
    return dict(new_w=new_w, new_h=new_h, top=top, bottom=bottom, left=left, right=right,
                out_h=new_h + top + bottom, out_w=new_w + left + right)


def linear_coeffs(src: int, dst: int, clamp_fraction: bool):
    """Tap indices and 11-bit weights of cv2's linear resize along one axis (resize.cpp, ``resize`` coefficient loop).
    ``scale`` is 1 / (dst / src) in double, the coordinate is rounded to float32 before the floor."""
    scale = 1.0 / (dst / src)
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_fraction:  # x axis: "if (sx < 0) fx = 0, sx = 0;  if (sx >= ssize.width - 1) fx = 0, sx = ssize.width - 1"
        lo = s < 0
        f[lo] = 0
        s[lo] = 0
        hi = s >= src - 1
        f[hi] = 0
        s[hi] = src - 1
    w1 = np.rint(f * np.float32(2048)).astype(np.int32)  # saturate_cast<short>(float) = round half to even
    w0 = np.rint((np.float32(1) - f) * np.float32(2048)).astype(np.int32)
    i0 = np.clip(s, 0, src - 1).astype(np.int32)  # y axis: rows clipped in resizeGeneric_Invoker
    i1 = np.clip(s + 1, 0, src - 1).astype(np.int32)
    return i0, i1, w0, w1


def resize_linear_u8(img: np.ndarray, new_w: int, new_h: int) -> np.ndarray:
    H, W = img.shape[:2]
    x0, x1, a0, a1 = linear_coeffs(W, new_w, True)
    y0, y1, b0, b1 = linear_coeffs(H, new_h, False)
    im = img.astype(np.int64)
    rows = im[:, x0, :] * a0[None, :, None] + im[:, x1, :] * a1[None, :, None]  # HResizeLinear, scale 2^11
    s0, s1 = rows[y0], rows[y1]
    out = (((b0[:, None, None] * (s0 >> 4)) >> 16) + ((b1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox(img_bgr: np.ndarray, new_shape=(640, 640), auto=False, scaleup=True, center=True, stride=32) -> np.ndarray:
    """uint8 HWC BGR image -> uint8 HWC **RGB** letterboxed image (what the network's uint8 NHWC input holds)."""
    g = letterbox_geometry(img_bgr.shape[:2], new_shape, auto, scaleup, center, stride)
    img = img_bgr
    if (img.shape[1], img.shape[0]) != (g["new_w"], g["new_h"]):
        img = resize_linear_u8(img, g["new_w"], g["new_h"])
    out = np.full((g["out_h"], g["out_w"], 3), PAD_VALUE, dtype=np.uint8)
    out[g["top"]:g["top"] + g["new_h"], g["left"]:g["left"] + g["new_w"]] = img
    return out[..., ::-1].copy()


def scale_boxes(img1_shape, boxes: np.ndarray, img0_shape) -> np.ndarray:
    """ops.scale_boxes (ultralytics/utils/ops.py:102-134, ratio_pad=None, padding=True, xyxy) + clip_boxes (:152-177) on
    fp32 boxes [n, 4] of the letterboxed image (h1, w1) -> original image (h0, w0).  Same fp32 op order as torch on CPU:
    subtract the (integer) pad, divide by the gain rounded to fp32, clamp."""
    gain = min(img1_shape[0] / img0_shape[0], img1_shape[1] / img0_shape[1])
    pad_x = round((img1_shape[1] - img0_shape[1] * gain) / 2 - 0.1)
    pad_y = round((img1_shape[0] - img0_shape[0] * gain) / 2 - 0.1)
    b = boxes.astype(np.float32).copy()
    b[:, [0, 2]] -= np.float32(pad_x)
    b[:, [1, 3]] -= np.float32(pad_y)
    b /= np.float32(gain)
    b[:, [0, 2]] = b[:, [0, 2]].clip(0, img0_shape[1])
    b[:, [1, 3]] = b[:, [1, 3]].clip(0, img0_shape[0])
    return b
