"""Packed model format: pickle-free round trip of the fused state dict + cfg, from the mirror model and -
where /root/reference exists - from the reference's own DetectionModel; the reloaded model compiles to the same plan
(same packed weights) and, interpreted on CPU, gives the same outputs."""
import os
import sys

import pytest
import torch

from cases import FORWARD_CASES
from helpers import load_cfg, rel_max
from plan_interp import Interp

from fce_yolo_b200 import packed
from fce_yolo_b200.plan import compile_model
from fce_yolo_b200.tasks import DetectionModel
from fce_yolo_b200.weights import load_synthetic, synth_images


def _run(model, x):
    plan = compile_model(model, x.shape[0], x.shape[2], x.shape[3], "fp32", torch.device("cpu"))
    it = Interp(plan, reuse_memory=True)
    it.input_tensor().copy_(x)
    it.run()
    return it.outputs()[0].clone(), plan


@pytest.mark.parametrize("name", ["n_fce_64", "s_cca_bicca8_64"])
def test_round_trip_mirror(tmp_path, name):
    case = FORWARD_CASES[name]
    cfg, scale = load_cfg(case)
    m = DetectionModel(cfg, scale=scale).fuse().eval()
    load_synthetic(m, case["seed"])
    path = str(tmp_path / "m.fcepack")
    hdr = packed.save_packed(m, path, meta={"note": "test"})
    assert hdr["meta"]["fused"] and set(hdr["tensors"]) == set(m.state_dict())
    assert open(path, "rb").read(8) == packed.MAGIC
    m2 = packed.load_packed(path)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    x = synth_images(case["img_seed"], 1, 64, 64)
    y1, p1 = _run(m, x)
    y2, p2 = _run(m2, x)
    assert torch.equal(y1, y2) and len(p1.nodes) == len(p2.nodes)


def test_unfused_model_is_folded_and_bf16_storage(tmp_path):
    cfg, scale = load_cfg(FORWARD_CASES["n_fce_64"])
    m = DetectionModel(cfg, scale=scale).eval()  # still carries BatchNorm
    g = torch.Generator().manual_seed(1)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_var.copy_(torch.rand(mod.num_features, generator=g) + 0.5)
            mod.running_mean.copy_(torch.randn(mod.num_features, generator=g) * 0.1)
    path = str(tmp_path / "m.fcepack")
    packed.save_packed(m, path)
    assert any(isinstance(x, torch.nn.BatchNorm2d) for x in m.modules())  # the caller's model is untouched
    m2 = packed.load_packed(path)
    x = synth_images(3, 1, 64, 64)
    y1, _ = _run(m, x)  # the plan compiler folds BN itself
    y2, _ = _run(m2, x)
    assert rel_max(y2, y1) < 1e-5
    small = str(tmp_path / "m16.fcepack")
    packed.save_packed(m, small, dtype=torch.bfloat16)
    assert os.path.getsize(small) < 0.6 * os.path.getsize(path)
    y3, _ = _run(packed.load_packed(small), x)
    assert rel_max(y3, y1) < 5e-2  # bf16-rounded weights


def test_rejects_foreign_files(tmp_path):
    p = tmp_path / "x.bin"
    p.write_bytes(b"not a pack file at all")
    with pytest.raises(ValueError):
        packed.load_packed(str(p))


@pytest.mark.skipif(not os.path.isdir("/root/reference/ultralytics"), reason="reference tree not present")
def test_pack_from_reference_model(tmp_path):
    """A model built by the reference's own parse_model packs under the same keys and reloads into the mirror."""
    os.environ.setdefault("YOLO_CONFIG_DIR", "/tmp/ulcfg")
    sys.dont_write_bytecode = True
    if "/root/reference" not in sys.path:
        sys.path.append("/root/reference")
    import ultralytics.nn.tasks as T

    ref = T.DetectionModel(T.yaml_model_load("yolo11n-fce.yaml"), verbose=False).eval()
    ref.fuse()
    load_synthetic(ref, 7)
    path = str(tmp_path / "ref.fcepack")
    packed.save_packed(ref, path)
    m2 = packed.load_packed(path)
    x = synth_images(9, 1, 64, 64)
    with torch.no_grad():
        y_ref = ref(x)[0]
    y2, _ = _run(m2, x)
    assert rel_max(y2, y_ref) < 1e-4
