"""Drop-in boundary against the LIVE reference (build container only: /root/reference is absent on the GPU box,
where these tests skip).  The plan compiler must walk the reference's own nn.Module instances, install() must
rebind only `_predict_once` / `non_max_suppression`, and nothing about the reference's state may change."""
import inspect
import os
import sys

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ultralytics")), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref():
    os.environ.setdefault("YOLO_CONFIG_DIR", "/tmp/ulcfg")
    sys.dont_write_bytecode = True
    if REF not in sys.path:
        sys.path.append(REF)
    import ultralytics.nn.tasks as T
    from ultralytics.utils import nms

    return T, nms


def _build(T, name, seed, variant=None):
    from cases import variant_cfg

    from fce_yolo_b200.weights import load_synthetic

    d = variant_cfg(T.yaml_model_load(name), variant)
    m = T.DetectionModel(d, verbose=False).eval()
    m.fuse()
    load_synthetic(m, seed)
    return m


@pytest.mark.parametrize("name,variant", [
    ("yolo11n-fce.yaml", None),
    ("yolo11s-fce.yaml", {5: ("CoordCrossAtt", [512, 16, 2]), 8: ("CoordAtt", [])}),
    ("yolo11m-bifpn.yaml", None),
])
def test_plan_compiles_from_reference_instances(ref, name, variant):
    """Same node list / packed weights the CUDA executor would launch, interpreted on CPU, equals the reference's
    own forward on its own module instances (fp32, 1e-4)."""
    from helpers import rel_max
    from plan_interp import Interp

    from fce_yolo_b200.plan import compile_model
    from fce_yolo_b200.weights import synth_images

    T, _ = ref
    m = _build(T, name, 3, variant)
    x = synth_images(5, 1, 64, 64)
    with torch.no_grad():
        y_ref, raw_ref = m(x)
    plan = compile_model(m, 1, 64, 64, "fp32", torch.device("cpu"))
    it = Interp(plan, reuse_memory=True)
    it.input_tensor().copy_(x)
    it.run()
    y, raw = it.outputs()
    assert rel_max(y, y_ref) < 1e-4
    for a, b in zip(raw, raw_ref):
        assert rel_max(a, b) < 1e-4


def test_install_rebinds_only_the_layer_loop(ref):
    import fce_yolo_b200

    T, nms = ref
    m = _build(T, "yolo11n-fce.yaml", 0)
    keys = list(m.state_dict())
    orig_nms = nms.non_max_suppression
    x = torch.rand(1, 3, 64, 64)
    with torch.no_grad():
        y0 = m(x)[0]
    fce_yolo_b200.install(m, "bf16")
    try:
        assert list(m.state_dict()) == keys and type(m).__module__ == "ultralytics.nn.tasks"
        # eval + CPU tensor: loud failure, never a silent CPU path
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m(x)
        # training mode / profile are not the hot path: the reference's own method runs
        m.train()
        out = m(x)
        assert isinstance(out, list) and len(out) == 3
        m.eval()
        # NMS: same signature as the reference, CPU tensors keep going to the original
        assert nms.non_max_suppression is not orig_nms
        assert list(inspect.signature(nms.non_max_suppression).parameters) == \
            list(inspect.signature(orig_nms).parameters)
        from fce_yolo_b200.weights import synth_predictions

        p = synth_predictions(7, 1, 2100)
        a = nms.non_max_suppression(p.clone(), 0.25, 0.7)
        b = orig_nms(p.clone(), 0.25, 0.7)
        assert all(torch.equal(u, v) for u, v in zip(a, b))
    finally:
        fce_yolo_b200.uninstall(m)
        fce_yolo_b200.uninstall_nms()
    assert nms.non_max_suppression is orig_nms
    with torch.no_grad():
        assert torch.equal(m(x)[0], y0)


def test_install_rejects_what_the_reference_rejects(ref):
    """CoordCrossAtt defaults at m-scale give mip=23, heads=2: the reference dies in view() at the first forward
    (fce_block.py:166); install() reports it up front."""
    import fce_yolo_b200
    from fce_yolo_b200.plan import PlanError

    T, _ = ref
    from cases import variant_cfg

    d = variant_cfg(T.yaml_model_load("yolo11m-fce.yaml"), {5: ("CoordCrossAtt", [])})
    try:
        m = T.DetectionModel(d, verbose=False).eval()
    except Exception:
        pytest.skip("the reference already fails in its construction-time stride probe")
    with pytest.raises(PlanError):
        fce_yolo_b200.install(m)
