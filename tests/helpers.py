"""Shared test helpers: golden loading, oracle-side model setup, error metrics."""
import os

import numpy as np
import torch
import yaml

from cases import variant_cfg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
CFG_DIR = os.path.join(ROOT, "fce_yolo_b200", "cfg")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def split_yaml_name(name: str):
    """'yolo11s-fce.yaml' -> ('yolo11-fce.yaml', 's')  (reference tasks.py:1759-1764)."""
    import re

    m = re.match(r"(yolo11)([nslmx])(.*\.yaml)$", name)
    return m.group(1) + m.group(3), m.group(2)


def load_cfg(case):
    base, scale = split_yaml_name(case["yaml"])
    d = yaml.safe_load(open(os.path.join(CFG_DIR, base)))
    return variant_cfg(d, case.get("variant")), scale


def rel_l2(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def rel_max(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def assert_nms_equal(idx, det, gidx, gdet, tie_perm=False):
    """Bit-exact comparison of NMS output.  ``tie_perm``: rows sharing an identical score may be
    permuted - only used where the reference itself is unordered: with more than max_nms candidates it
    pre-sorts with an *unstable* argsort (ultralytics/utils/nms.py:138), so the order of exact score
    ties is unspecified there; this implementation (and the oracle) define it as index-ascending."""
    idx, gidx = np.asarray(idx), np.asarray(gidx)
    det, gdet = np.asarray(det), np.asarray(gdet)
    assert idx.shape == gidx.shape and det.shape == gdet.shape
    if not tie_perm:
        assert np.array_equal(idx, gidx)
        assert np.array_equal(det, gdet)
        return
    assert np.array_equal(det[:, 4], gdet[:, 4])
    s = det[:, 4]
    start = 0
    for end in range(1, len(s) + 1):
        if end == len(s) or s[end] != s[start]:
            a = sorted(zip(idx[start:end].tolist(), map(tuple, det[start:end].tolist())))
            b = sorted(zip(gidx[start:end].tolist(), map(tuple, gdet[start:end].tolist())))
            assert a == b
            start = end
