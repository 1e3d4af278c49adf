"""Deterministic synthetic weights for fused FCE-YOLO models.

There is no network (no checkpoints, no datasets), and the reference's raw random init makes
eval-mode activations collapse to ~1e-10 by layer 12 (BN running stats are 0/1), which would
make every parity check vacuous.  This module plays the role of ``initialize_weights`` +
BN calibration (reference: ultralytics/utils/torch_utils.py:463-473) for the
*fused* graph: every tensor of the state dict is a pure function of (seed, key, shape), drawn
with numpy's PCG64 so that the build container (which generates the golden fixtures from the
real reference) and the GPU box produce bit-identical weights.

Variance is chosen so activations stay O(1) through the network: conv weights
N(0, g^2/fan_in) with g compensating SiLU's second-moment loss, small biases, BiFPN fusion
weights in (0.5, 1.5) plus an occasional negative one (exercises the relu in
fce_block.py:55), classification bias shifted so that a few hundred anchors clear conf=0.25.
"""
from __future__ import annotations

import re
import zlib

import numpy as np
import torch

SILU_GAIN = 1.65  # self-stabilising: SiLU needs gain 2 for small inputs, sqrt(2) for large


def _rng(seed: int, key: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(key.encode())]))


def synth_tensor(seed: int, key: str, shape, cls_bias: float = -4.5) -> np.ndarray:
    shape = tuple(int(s) for s in shape)
    r = _rng(seed, key)
    leaf = key.rsplit(".", 1)[-1]
    if key.endswith("dfl.conv.weight"):
        return np.arange(shape[1], dtype=np.float32).reshape(shape)  # block.py:70-72
    if leaf == "w" and len(shape) == 1:  # BiFPN_Concat.w
        w = r.uniform(0.5, 1.5, size=shape).astype(np.float32)
        if len(shape) and shape[0] == 3:
            w[1] = -0.25  # relu() must zero this branch
        return w
    if leaf == "weight" and len(shape) == 4:
        fan_in = shape[1] * shape[2] * shape[3]
        gain = SILU_GAIN
        # bare projections feeding a softmax / sigmoid: keep logits moderate
        if any(t in key for t in ("proj_q", "proj_k", "q_conv", "k_conv", "qkv", "attn.pe")):
            gain = 1.0
        elif any(t in key for t in (".out_h.", ".out_w.", ".cv_h.", ".cv_w.")) or key.endswith(".proj.weight"):
            gain = 2.0
        # residual branches (Bottleneck.cv2 inside C3k2/C3k, PSABlock attn.proj / ffn.1): damp so
        # x + f(x) does not grow geometrically at the deeper m/l/x scales
        elif re.search(r"\.m\.\d+\.cv2\.conv\.weight$", key):
            gain = 0.9
        elif "attn.proj" in key or ".ffn.1." in key:
            gain = 0.5
        # Detect's last 1x1s (head.py:94,103): spread the DFL / class logits so decode and NMS
        # see a non-degenerate distribution
        elif re.search(r"\.cv3\.\d+\.2\.weight$", key):
            gain = 18.0
        elif re.search(r"\.cv2\.\d+\.2\.weight$", key):
            gain = 10.0
        return (r.standard_normal(size=shape) * (gain / np.sqrt(fan_in))).astype(np.float32)
    if leaf == "bias":
        b = (r.standard_normal(size=shape) * 0.1).astype(np.float32)
        parts = key.split(".")
        # Detect.cv3[i][2].bias : final class logits (head.py:103, bias_init head.py:176)
        if "cv3" in parts and parts[-2] == "2":
            b += np.float32(cls_bias)
        if "cv2" in parts and parts[-2] == "2" and len(parts) >= 5:
            b += np.float32(1.0)  # head.py:175 box bias
        return b
    raise KeyError(f"no synthesis rule for {key} {shape}")


@torch.no_grad()
def load_synthetic(model: torch.nn.Module, seed: int = 0, cls_bias: float = -4.5) -> dict:
    """Fill a *fused* model in place; returns the fp32 state dict (CPU tensors)."""
    sd = model.state_dict()
    out = {}
    for k, v in sd.items():
        if not torch.is_floating_point(v):
            continue
        if ".bn." in k:
            raise ValueError(f"load_synthetic expects a fused model, found {k}")
        t = torch.from_numpy(synth_tensor(seed, k, v.shape, cls_bias))
        v.copy_(t.to(v.dtype))
        out[k] = t
    return out


def synth_images(seed: int, batch: int, h: int, w: int, ch: int = 3) -> torch.Tensor:
    """Synthetic BCHW fp32 images in [0,1) (the reference's LoadTensor contract,
    ultralytics/data/loaders.py:597-616), reproducible across machines."""
    r = np.random.Generator(np.random.PCG64([seed, 0xF0CE]))
    return torch.from_numpy(r.random(size=(batch, ch, h, w), dtype=np.float32))


def synth_predictions(seed: int, batch: int, anchors: int, nc: int = 80, imgsz: int = 640,
                      sharp: float = 8.0) -> torch.Tensor:
    """Synthetic pre-NMS tensor [B, 4+nc, A]: random-init weights never produce
    confident classes, so NMS parity is checked on these instead."""
    r = np.random.Generator(np.random.PCG64([seed, 0x9A5]))
    p = np.empty((batch, 4 + nc, anchors), dtype=np.float32)
    p[:, 0:2] = r.random(size=(batch, 2, anchors), dtype=np.float32) * np.float32(imgsz)
    p[:, 2:4] = r.random(size=(batch, 2, anchors), dtype=np.float32) * np.float32(100) + np.float32(5)
    p[:, 4:] = r.random(size=(batch, nc, anchors), dtype=np.float32) ** np.float32(sharp)
    return torch.from_numpy(p)
