"""Packed model format: the on-disk side of the path without pickle.

The reference ships weights as ``.pt`` files - a pickled ``nn.Module`` (fp16 EMA copy) inside a dict
(ultralytics/engine/trainer.py:584-623, loaded by ultralytics/nn/tasks.py:1371-1486), which needs the whole Python
package importable at the same class paths just to deserialise.  A ``.fcepack`` file is

    8 bytes   magic "FCEPACK1"
    8 bytes   little-endian length N of the JSON header
    N bytes   JSON: {"cfg": <model yaml as dict>, "meta": {...}, "tensors": {key: {"dtype", "shape", "offset", "nbytes"}}}
    padding   to a 64-byte boundary
    data      raw little-endian tensors, each 64-byte aligned

holding the FUSED state dict (BatchNorm folded: the only form the inference path needs) under the reference's own
``state_dict`` keys (``model.<i>.<attr>...conv.weight / conv.bias``), so it can be produced from a reference model or
from this package's mirror and read by either.  Loading is a header parse plus one ``numpy.memmap``: O(read), no code
execution.  ``load_packed`` rebuilds the mirror ``DetectionModel`` (parameter container) from the embedded cfg; the
plan compiler then packs kernel-layout weights as usual.
"""
from __future__ import annotations

import json
import struct

import numpy as np
import torch

MAGIC = b"FCEPACK1"
_ALIGN = 64
_DT = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16, "int64": torch.int64}


def _cfg_jsonable(cfg: dict) -> dict:
    return json.loads(json.dumps(cfg, default=lambda o: o.tolist() if hasattr(o, "tolist") else str(o)))


def save_packed(model, path: str, dtype: torch.dtype | None = None, meta: dict | None = None) -> dict:
    """Writes ``model`` (reference or mirror DetectionModel with a ``.yaml`` cfg dict) to ``path``.  An un-fused model is
    folded on a deep copy first (reference formula, torch_utils.py:237-267).  ``dtype`` optionally stores floating
    tensors narrower (torch.bfloat16 halves the file; BiFPN fusion weights and DFL stay fp32)."""
    import copy

    from .modules import fuse_module

    if any(isinstance(m, torch.nn.BatchNorm2d) for m in model.modules()):
        model = fuse_module(copy.deepcopy(model))
    cfg = getattr(model, "yaml", None)
    if not isinstance(cfg, dict):
        raise ValueError("model has no .yaml cfg dict to embed")
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    index, blobs, off = {}, [], 0
    for k, t in sd.items():
        if not t.is_floating_point() and t.dtype != torch.int64:
            continue  # e.g. num_batches_tracked never survives fusing
        keep32 = k.endswith(".w") or "dfl" in k
        if dtype is not None and t.is_floating_point() and not keep32:
            t = t.to(dtype)
        name = str(t.dtype).replace("torch.", "")
        if name not in _DT:
            raise ValueError(f"unsupported dtype {t.dtype} for {k}")
        raw = t.contiguous().view(torch.uint8).numpy().tobytes() if t.numel() else b""
        index[k] = {"dtype": name, "shape": list(t.shape), "offset": off, "nbytes": len(raw)}
        blobs.append(raw)
        off += (len(raw) + _ALIGN - 1) // _ALIGN * _ALIGN
    header = {"cfg": _cfg_jsonable(cfg), "meta": dict(meta or {}, format=1, fused=True), "tensors": index}
    hb = json.dumps(header).encode()
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", len(hb)))
        f.write(hb)
        f.write(b"\0" * (-(16 + len(hb)) % _ALIGN))
        for raw in blobs:
            f.write(raw)
            f.write(b"\0" * (-len(raw) % _ALIGN))
    return header


def read_header(path: str):
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError(f"{path} is not a FCEPACK1 file")
        (n,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(n).decode())
    data_off = (16 + n + _ALIGN - 1) // _ALIGN * _ALIGN
    return header, data_off


def load_state_dict(path: str) -> tuple[dict, dict]:
    """(header, {key: tensor}) - tensors are copies out of one read-only memory map."""
    header, data_off = read_header(path)
    mm = np.memmap(path, dtype=np.uint8, mode="r", offset=data_off) if any(
        t["nbytes"] for t in header["tensors"].values()) else np.zeros(0, np.uint8)
    sd = {}
    for k, t in header["tensors"].items():
        raw = torch.from_numpy(np.array(mm[t["offset"]:t["offset"] + t["nbytes"]]))
        sd[k] = raw.view(_DT[t["dtype"]]).reshape(t["shape"]).clone() if t["nbytes"] else torch.zeros(t["shape"], dtype=_DT[t["dtype"]])
    return header, sd


def load_packed(path: str):
    """Rebuilds the (fused, eval) mirror DetectionModel from a ``.fcepack`` file; weights come back in fp32 whatever
    the storage dtype (the plan compiler chooses the kernel dtype)."""
    from .tasks import DetectionModel

    header, sd = load_state_dict(path)
    model = DetectionModel(header["cfg"]).fuse().eval()
    own = model.state_dict()
    missing = [k for k in own if k not in sd]
    extra = [k for k in sd if k not in own]
    if missing or extra:
        raise ValueError(f"{path}: state dict mismatch (missing {missing[:3]}..., unexpected {extra[:3]}...)")
    model.load_state_dict({k: v.to(own[k].dtype) for k, v in sd.items()}, strict=True)
    return model


def convert_checkpoint(pt_path: str, out_path: str, dtype: torch.dtype | None = None) -> dict:
    """Reference checkpoint ingestion: a ``.pt`` written by the reference's trainer (engine/trainer.py:584-623) is a
    pickled dict whose ``ema`` / ``model`` entries are whole ``nn.Module`` objects (fp16), which only deserialise where
    the reference package is importable at its original class paths (nn/tasks.py:1371-1486 ``torch_safe_load`` /
    ``load_checkpoint``).  This reads one - preferring the EMA weights like ``load_checkpoint`` does (tasks.py:1467) -,
    casts to fp32, folds BatchNorm and writes the pickle-free ``.fcepack`` twin.  Run it ONCE wherever ``ultralytics`` (the
    reference) can be imported; serving then needs neither pickle nor the reference."""
    try:
        import ultralytics  # noqa: F401  (the unpickler resolves ultralytics.nn.tasks.DetectionModel etc.)
    except ImportError as e:
        raise RuntimeError("convert_checkpoint() unpickles reference nn.Modules: the reference package (ultralytics) must be "
                           "importable here; the resulting .fcepack file no longer needs it") from e
    ckpt = torch.load(pt_path, map_location="cpu", weights_only=False)
    model = ckpt.get("ema") or ckpt["model"] if isinstance(ckpt, dict) else ckpt
    model = model.float().eval()
    meta = {"source": "reference .pt", "epoch": ckpt.get("epoch") if isinstance(ckpt, dict) else None}
    return save_packed(model, out_path, dtype=dtype, meta=meta)
