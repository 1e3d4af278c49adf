"""Graph builder: YAML -> module list (mirror of the reference's L2 layer).

``yaml_model_load`` / ``parse_model`` / ``DetectionModel`` keep the reference's names, argument
meaning and resolution rules (ultralytics/nn/tasks.py:1489-1781, :339-420) for the module set on
the FCE hot path; anything else raises.  Nothing here runs a PyTorch forward: strides are derived
from the graph, and ``DetectionModel.forward`` executes the compiled plan on the GPU.
"""
from __future__ import annotations

import copy
import math
import os
import re

import torch
import torch.nn as nn
import yaml

from . import modules as M

CFG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cfg")

_WIDTH_SCALED = {"Conv", "C3k2", "SPPF", "C2PSA"}  # the "base_modules" subset (tasks.py:1525-1563)
_REPEATED = {"C3k2", "C2PSA"}  # "repeat_modules" subset (tasks.py:1564-1581)


def make_divisible(x, divisor: int) -> int:
    return int(math.ceil(x / divisor) * divisor)


def yaml_model_load(path) -> dict:
    """'yolo11s-fce.yaml' -> dict of yolo11-fce.yaml with d['scale']='s' (tasks.py:1746-1781)."""
    path = str(path)
    name = os.path.basename(path)
    unified = re.sub(r"(\d+)([nslmx])(.+)?$", r"\1\3", name)
    scale = ""
    m = re.search(r"yolo(?:v|e-)?\d+([nslmx])", os.path.splitext(name)[0])
    if m:
        scale = m.group(1)
    for cand in (path, os.path.join(os.path.dirname(path), unified), os.path.join(CFG_DIR, name),
                 os.path.join(CFG_DIR, unified)):
        if os.path.isfile(cand):
            with open(cand) as f:
                d = yaml.safe_load(f)
            d["scale"] = scale
            d["yaml_file"] = path
            return d
    raise FileNotFoundError(f"model config '{path}' not found (looked in {CFG_DIR} too)")


def variant_cfg(d: dict, variant) -> dict:
    """YAML variants as cfg dicts (the reference accepts a dict in DetectionModel, tasks.py:377): replace
    backbone rows, e.g. {5: ("CoordAtt", []), 8: ("CoordAtt", [])} for BASELINE config 2."""
    d = copy.deepcopy(d)
    for row, spec in (variant or {}).items():
        d["backbone"][int(row)] = [-1, 1, spec[0], list(spec[1])]
    return d


def _fce_defaults(inp: int, args: list, width: float, max_ch: float, with_heads: bool):
    """Argument resolution for CoordAtt / CoordCrossAtt / BiCoordCrossAtt rows (tasks.py:1636-1708):
    oup defaults to inp and is width-scaled only when given; reduction defaults to
    clamp(floor(sqrt(inp)), 8, 32); heads default to min(8, inp//32) lowered until dim/head >= 8."""
    oup = make_divisible(min(args[0], max_ch) * width, 8) if args else inp
    red = args[1] if len(args) > 1 else max(8, min(32, int(inp ** 0.5)))
    if not with_heads:
        return [inp, oup, red]
    if len(args) > 2:
        heads = args[2]
    else:
        base = max(8, inp // red)
        heads = max(1, min(8, inp // 32))
        while heads > 1 and base // heads < 8:
            heads -= 1
    return [inp, oup, red, heads]


def parse_model(d: dict, ch: int, verbose: bool = False):
    """Returns (nn.Sequential, sorted save list); each layer carries .i .f .type .np like the
    reference's (tasks.py:1732-1741)."""
    nc, scales = d.get("nc"), d.get("scales")
    depth, width, max_ch = d.get("depth_multiple", 1.0), d.get("width_multiple", 1.0), float("inf")
    scale = d.get("scale")
    if scales:
        if not scale:
            scale = next(iter(scales))
        depth, width, max_ch = scales[scale]
    if d.get("activation"):
        raise NotImplementedError("custom activation: only the default SiLU has a fused epilogue")

    chans = [ch]
    layers, save = [], []
    for i, (f, n, name, args) in enumerate(d["backbone"] + d["head"]):
        args = [nc if a == "nc" else a for a in args]
        n_rep = max(round(n * depth), 1) if n > 1 else n
        if name in _WIDTH_SCALED:
            c1 = chans[f]
            c2 = make_divisible(min(args[0], max_ch) * width, 8)
            cargs = [c1, c2, *args[1:]]
            if name in _REPEATED:
                cargs.insert(2, n_rep)
            if name == "C3k2" and scale in "mlx":
                while len(cargs) < 4:
                    cargs.append(False)
                cargs[3] = True
            mod = getattr(M, name)(*cargs)
        elif name == "nn.Upsample":
            c2, cargs = chans[f], list(args)
            mod = nn.Upsample(*cargs)
        elif name == "Concat":
            c2, cargs = sum(chans[x] for x in f), list(args)
            mod = M.Concat(*cargs)
        elif name == "BiFPN_Concat":
            c1 = [chans[x] for x in f] if isinstance(f, list) else [chans[f]]
            c2 = args[0] if args else max(c1)
            c2 = make_divisible(min(c2, max_ch) * width, 8)  # yes, scaled again (tasks.py:1634)
            cargs = [c1, c2]
            mod = M.BiFPN_Concat(*cargs)
        elif name in ("CoordAtt", "CoordCrossAtt", "BiCoordCrossAtt"):
            cargs = _fce_defaults(chans[f], args, width, max_ch, with_heads=name != "CoordAtt")
            c2 = cargs[1]
            mod = getattr(M, name)(*cargs)
        elif name == "Detect":
            cargs = [args[0], [chans[x] for x in f]]
            c2 = None
            mod = M.Detect(*cargs)
        else:
            raise NotImplementedError(f"module '{name}' is outside the FCE-YOLO hot path")
        mod.i, mod.f, mod.type = i, f, name
        mod.np = sum(p.numel() for p in mod.parameters())
        if verbose:
            print(f"{i:>3}{str(f):>14}{n_rep:>3}{mod.np:>10}  {name:<18}{cargs}")
        save.extend(x % i for x in ([f] if isinstance(f, int) else f) if x != -1)
        layers.append(mod)
        if i == 0:
            chans = []
        chans.append(c2)
    return nn.Sequential(*layers), sorted(save)


def graph_strides(model: nn.Sequential) -> list[int]:
    """Cumulative stride at each Detect input, derived from the graph (the reference probes this with
    a 256x256 forward, tasks.py:396-411)."""
    s = []
    for m in model:
        f = m.f if isinstance(m.f, int) else m.f[0]
        cur = 1 if m.i == 0 else s[f if f >= 0 else m.i + f]
        if isinstance(m, M.Conv):
            cur = cur * m.conv.stride[0]
        elif isinstance(m, nn.Upsample):
            cur = cur // int(m.scale_factor)
        s.append(cur)
    det = model[-1]
    return [s[j] for j in det.f]


class DetectionModel(nn.Module):
    """Mirror of ultralytics.nn.tasks.DetectionModel (tasks.py:339-420) for the detection graphs."""

    def __init__(self, cfg="yolo11n-fce.yaml", ch: int = 3, nc: int | None = None, verbose: bool = False,
                 scale: str | None = None):
        super().__init__()
        self.yaml = cfg if isinstance(cfg, dict) else yaml_model_load(cfg)
        self.yaml = copy.deepcopy(self.yaml)
        if scale:
            self.yaml["scale"] = scale
        self.yaml["channels"] = ch
        if nc and nc != self.yaml["nc"]:
            self.yaml["nc"] = nc
        self.model, self.save = parse_model(copy.deepcopy(self.yaml), ch=ch, verbose=verbose)
        self.names = {i: f"{i}" for i in range(self.yaml["nc"])}
        self.inplace = True
        det = self.model[-1]
        if isinstance(det, M.Detect):
            det.stride = torch.tensor([float(s) for s in graph_strides(self.model)])
            self.stride = det.stride
            det.bias_init()
        else:
            self.stride = torch.tensor([32.0])
        self._fused = False

    def is_fused(self) -> bool:
        return not any(isinstance(m, nn.BatchNorm2d) for m in self.modules())

    def fuse(self, verbose: bool = False):
        M.fuse_module(self)
        self._fused = True
        return self

    def forward(self, x, *args, **kwargs):
        return self.predict(x, *args, **kwargs)

    def predict(self, x, profile=False, visualize=False, augment=False, embed=None):
        if profile or visualize or augment or embed:
            raise NotImplementedError("profile/visualize/augment/embed are not on the B200 hot path")
        from .engine import run_model

        return run_model(self, x)
