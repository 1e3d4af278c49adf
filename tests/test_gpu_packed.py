"""SURVEY 8f-3 on the GPU: a model loaded from a ``.fcepack`` file predicts exactly what the in-memory model predicts, and
a reference-style ``.pt`` checkpoint (pickled fp16 nn.Module, engine/trainer.py:584-623) converted with
``packed.convert_checkpoint`` does too."""
import copy
import os
import sys

import pytest
import torch

import detection_parity as DP

pytestmark = pytest.mark.gpu


def _detections(model, x):
    from fce_yolo_b200.predict import Predictor

    p = Predictor(model, x.shape[0], x.shape[1], precision="bf16", conf=0.25, iou=0.7, input_u8=True)
    det, cnt = p.infer(x.pin_memory())
    return det.clone(), cnt.clone(), p.keep.cpu().clone()


@pytest.mark.parametrize("name", ["cfg0_n_fce", "cfg3_s_cca_bicca8"])
def test_fcepack_model_predicts_like_the_in_memory_model(tmp_path, name):
    from fce_yolo_b200 import packed

    case = dict(DP.CONFIGS[name], size=320, batch=2)
    _, model, _ = DP.build(case)
    x = DP.u8_batch(5, 2, 320)
    ref = _detections(model, x)
    path = str(tmp_path / "m.fcepack")
    packed.save_packed(model, path)
    got = _detections(packed.load_packed(path), x)
    assert int(ref[1].sum()) > 50
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
    # bf16 storage halves the file; conv weights round the same way the plan does, only the fp32 biases move
    path16 = str(tmp_path / "m16.fcepack")
    packed.save_packed(model, path16, dtype=torch.bfloat16)
    assert os.path.getsize(path16) < 0.6 * os.path.getsize(path)
    d16, c16, _ = _detections(packed.load_packed(path16), x)
    assert (c16 - ref[1]).abs().max() <= 30
    assert int(c16.sum()) > 50


def test_reference_pt_checkpoint_converts_and_predicts(tmp_path):
    sys.path.insert(0, os.path.join(DP.ROOT, "baseline"))
    import ref_env

    if not ref_env.installed():
        pytest.skip("baseline/_ref not built")
    from fce_yolo_b200 import packed

    case = dict(DP.CONFIGS["cfg1_s_coordatt"], size=320, batch=2)
    ref_model = ref_env.reference_model(case["yaml"], case.get("variant"), case["seed"])
    # what the reference's trainer writes (engine/trainer.py:595-611): a dict with a pickled fp16 copy of the module
    pt = str(tmp_path / "last.pt")
    torch.save({"epoch": 3, "model": None, "ema": copy.deepcopy(ref_model).half(), "updates": 1, "optimizer": None,
                "train_args": {}, "date": "synthetic", "version": "8.3.242"}, pt)
    out = str(tmp_path / "last.fcepack")
    hdr = packed.convert_checkpoint(pt, out)
    assert hdr["meta"]["fused"] and hdr["meta"]["epoch"] == 3
    loaded = packed.load_packed(out)
    # the same weights in the mirror, rounded through fp16 like the checkpoint
    _, mirror, _ = DP.build(case)
    with torch.no_grad():
        for p_ in mirror.parameters():
            p_.copy_(p_.half().float())
    x = DP.u8_batch(6, 2, 320)
    got, ref = _detections(loaded, x), _detections(mirror, x)
    assert int(ref[1].sum()) > 50
    for a, b in zip(got, ref):
        assert torch.equal(a, b)
