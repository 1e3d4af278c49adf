"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle and the committed
reference goldens.  Tolerances (north_star): fp32 mode 1e-4 relative per layer; bf16 mode 2e-2 relative L2
per layer, teacher-forced (each layer is fed the fp32 oracle's input); NMS indices / class ids bit-exact."""
import numpy as np
import pytest
import torch

from cases import FORWARD_CASES, MODULE_CASES, NMS_CASES
from helpers import assert_nms_equal, golden, load_cfg, rel_l2, rel_max
from test_oracle_golden import module_case_io

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_TOL = 2e-2
E2E_BF16_MIN_FRAC = {}  # per-case exceptions to the 99.5 % box criterion of test_end_to_end_bf16 (none)


def dev():
    return torch.device("cuda:0")


def build(case):
    from fce_yolo_b200.tasks import DetectionModel
    from fce_yolo_b200.weights import load_synthetic, synth_images

    cfg, scale = load_cfg(case)
    model = DetectionModel(cfg, scale=scale).fuse().eval()
    sd = load_synthetic(model, case["seed"])
    x = synth_images(case["img_seed"], case["batch"], case["size"], case["size"])
    return cfg, scale, model, sd, x


@pytest.mark.parametrize("name", ["n_fce_64", "s_coordatt_64", "s_cca_bicca8_64", "m_bifpn_64", "x_fce_64",
                                  "n_stock_64"])
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_layers_teacher_forced(name, prec):
    """Every top-level layer, fed the oracle's own input, against the oracle's output."""
    from fce_yolo_b200.engine import run_module
    from oracle import fce_oracle as O

    case = FORWARD_CASES[name]
    cfg, scale, model, sd, x = build(case)
    (yo, rawo), ys = O.forward(cfg, scale, sd, x, keep_layers=True)
    graph = O.resolve_graph(cfg, scale)
    worst = {}
    for i, m in enumerate(model.model):
        inp = O.layer_inputs(graph, ys, x, i)
        if type(m).__name__ == "Upsample":
            continue  # folded into its consumer; covered by the BiFPN / Concat layers and the e2e test
        ref = ys[i]
        if isinstance(inp, list):
            out = run_module(m, [t.to(dev()) for t in inp], precision=prec)
        else:
            out = run_module(m, inp.to(dev()), precision=prec)
        if isinstance(ref, tuple):  # Detect: (y, raw)
            pairs = [(out[0], ref[0])] + list(zip(out[1], ref[1]))
        else:
            pairs = [(out, ref)]
        for a, b in pairs:
            a = a.float().cpu()
            assert a.shape == b.shape
            if prec == "fp32":
                e = max(rel_l2(a, b), rel_max(a, b))
                assert e < FP32_TOL, (name, i, type(m).__name__, e)
            else:
                e = rel_l2(a, b)
                assert e < BF16_TOL, (name, i, type(m).__name__, e)
            worst[type(m).__name__] = max(worst.get(type(m).__name__, 0), e)
    print(name, prec, {k: f"{v:.2e}" for k, v in worst.items()})


@pytest.mark.parametrize("name", list(FORWARD_CASES))
def test_end_to_end_fp32_vs_golden(name):
    """Whole graph in fp32 mode against the reference's own output (golden) and the oracle."""
    from fce_yolo_b200.engine import run_model

    case = FORWARD_CASES[name]
    cfg, scale, model, sd, x = build(case)
    y, raw = run_model(model, x.to(dev()), precision="fp32")
    g = golden("fwd_" + name)
    sub = case.get("y_stride", 1)
    assert rel_max(y.cpu()[:, :, ::sub], g["y"]) < 2e-4
    for i, r in enumerate(raw):
        s = g[f"raw{i}_sum"]
        assert abs(r.double().abs().sum().item() - s[1]) / s[1] < 1e-4
    # graph replay gives the same answer as the first (eager) run
    y2, _ = run_model(model, x.to(dev()), precision="fp32")
    assert torch.equal(y, y2)


@pytest.mark.parametrize("name", ["n_fce_64", "s_coordatt_64", "s_cca_bicca8_64", "m_bifpn_64", "x_fce_64", "n_fce_640"])
def test_end_to_end_bf16(name):
    """bf16 mode end to end.  The reference's own bf16 forward drifts up to 0.16 relL2 from its fp32 forward
    (SURVEY E.2), so the end-to-end bound is loose; boxes are checked by IoU against the fp32 oracle."""
    from fce_yolo_b200.engine import run_model
    from oracle import fce_oracle as O

    case = FORWARD_CASES[name]
    cfg, scale, model, sd, x = build(case)
    yo, rawo = O.forward(cfg, scale, sd, x)
    y, raw = run_model(model, x.to(dev()), precision="bf16")
    y = y.cpu()
    assert rel_l2(y[:, :4], yo[:, :4]) < 0.05
    # IoU of every decoded box against the fp32 oracle's box for the same anchor
    def xyxy(t):
        cx, cy, w, h = t[:, 0], t[:, 1], t[:, 2], t[:, 3]
        return cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2

    ax1, ay1, ax2, ay2 = xyxy(y)
    bx1, by1, bx2, by2 = xyxy(yo)
    iw = (torch.minimum(ax2, bx2) - torch.maximum(ax1, bx1)).clamp(min=0)
    ih = (torch.minimum(ay2, by2) - torch.maximum(ay1, by1)).clamp(min=0)
    inter = iw * ih
    iou = inter / ((ax2 - ax1) * (ay2 - ay1) + (bx2 - bx1) * (by2 - by1) - inter)
    frac = (iou >= 0.99).float().mean().item()
    print(name, "bf16 boxes with IoU>=0.99:", frac, "min IoU", iou.min().item())
    import json
    import os

    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_e2e_bf16_pytest.jsonl"), "a") as f:
            f.write(json.dumps({"case": name, "frac_iou99_per_anchor": frac, "min_iou": iou.min().item()}) + "\n")
    # north_star: IoU >= 0.99 for >= 99.5 % - per ANCHOR here (every decoded box, no NMS selection), against the fp32
    # oracle; the matched-detection form of the criterion at BASELINE image sizes is tests/test_gpu_detections.py
    assert frac >= E2E_BF16_MIN_FRAC.get(name, 0.995)


@pytest.mark.parametrize("name", list(MODULE_CASES))
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_fce_modules_vs_reference_golden(name, prec):
    from fce_yolo_b200.engine import run_module

    case = MODULE_CASES[name]
    mod, sd, xs = module_case_io(case)
    xs = [t.to(dev()) for t in xs]
    y = run_module(mod.eval(), xs if case["cls"] == "BiFPN_Concat" else xs[0], precision=prec)
    g = golden("mod_" + name)["y"]
    if prec == "fp32":
        assert max(rel_l2(y.cpu(), g), rel_max(y.cpu(), g)) < FP32_TOL
    else:
        assert rel_l2(y.float().cpu(), g) < BF16_TOL


CONV_SHAPES = [
    # cin, cout, k, s, H, W, B, act
    (3, 16, 3, 2, 64, 64, 2, True), (16, 32, 3, 2, 32, 32, 2, True), (32, 32, 1, 1, 16, 16, 2, True),
    (48, 64, 1, 1, 16, 16, 1, True), (16, 8, 3, 1, 16, 16, 2, True), (8, 16, 3, 1, 16, 16, 2, True),
    (64, 64, 3, 1, 20, 20, 3, True), (128, 128, 3, 2, 40, 40, 2, True), (256, 80, 1, 1, 20, 20, 2, False),
    (64, 64, 3, 1, 80, 80, 1, True), (192, 128, 1, 1, 40, 24, 2, True), (96, 48, 1, 1, 12, 20, 2, True),
    (384, 256, 1, 1, 8, 8, 4, True), (512, 512, 3, 2, 8, 8, 2, True), (256, 256, 3, 1, 6, 10, 2, False),
]


@pytest.mark.parametrize("shape", CONV_SHAPES)
@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_conv_shapes(shape, prec):
    """Conv(+BN folded)+SiLU over the channel/stride/size combinations of SURVEY B.2, incl. ragged tiles."""
    from fce_yolo_b200 import modules as M
    from fce_yolo_b200.engine import run_module

    cin, cout, k, s, H, W, B, act = shape
    g = torch.Generator().manual_seed(cin * 1000 + cout + k + s)
    m = M.Conv(cin, cout, k, s, act=act)
    M.fuse_module(m)
    m.conv.weight.data.copy_(torch.randn(m.conv.weight.shape, generator=g) / (cin * k * k) ** 0.5)
    m.conv.bias.data.copy_(torch.randn(cout, generator=g) * 0.1)
    x = torch.randn(B, cin, H, W, generator=g)
    ref = torch.nn.functional.conv2d(x, m.conv.weight, m.conv.bias, stride=s, padding=k // 2)
    if act:
        ref = torch.nn.functional.silu(ref)
    y = run_module(m.eval(), x.to(dev()), precision=prec).float().cpu()
    if prec == "fp32":
        assert max(rel_l2(y, ref), rel_max(y, ref)) < FP32_TOL
    else:
        assert rel_l2(y, ref) < BF16_TOL


@pytest.mark.parametrize("name", list(NMS_CASES))
def test_nms_bit_exact(name):
    """Keep indices, class ids and copied-through boxes/conf: bit-exact vs the reference golden and the oracle."""
    from fce_yolo_b200.nms import non_max_suppression
    from oracle import nms_oracle

    case = NMS_CASES[name]
    p = case["make"]()
    out, idx = non_max_suppression(p.to(dev()), return_idxs=True, **case["kw"])
    g = golden("nms_" + name)
    dets_o, idxs_o = nms_oracle.non_max_suppression(p.numpy(), **case["kw"])
    for b in range(p.shape[0]):
        o, i = out[b].cpu().numpy(), idx[b].cpu().numpy()
        assert_nms_equal(i, o, idxs_o[b], dets_o[b])  # vs oracle: exact incl. tie order
        assert_nms_equal(i, o, g[f"idx{b}"], g[f"det{b}"], tie_perm=case.get("tie_perm", False))


def test_nms_on_model_output():
    """NMS on a real forward output (fp32 mode) agrees with the oracle's NMS on the oracle's forward output:
    same kept anchors and classes wherever the scores are not within fp32 noise of each other."""
    from fce_yolo_b200.engine import run_model
    from fce_yolo_b200.nms import non_max_suppression
    from oracle import nms_oracle

    case = dict(FORWARD_CASES["n_fce_640"])
    cfg, scale, model, sd, x = build(case)
    y, _ = run_model(model, x.to(dev()), precision="fp32")
    conf = float(y[:, 4:].amax(1).flatten().kthvalue(int(0.9 * y.shape[2])).values)  # ~10% of anchors pass
    conf = min(max(conf, 0.01), 0.9)
    out, idx = non_max_suppression(y.clone(), conf_thres=conf, iou_thres=0.7, return_idxs=True)
    dets_o, idxs_o = nms_oracle.non_max_suppression(y.cpu().numpy(), conf_thres=conf, iou_thres=0.7)
    assert_nms_equal(idx[0].cpu().numpy(), out[0].cpu().numpy(), idxs_o[0], dets_o[0])
    assert len(idxs_o[0]) > 10


def test_large_batch_properties():
    """BASELINE-size run (s-scale CoordAtt variant, B=64, 640x640, bf16): size-independent properties -
    every image of a batch of identical images gives identical output, and the batched result equals the
    batch-1 result for the same image (images are independent: SURVEY 8e)."""
    from fce_yolo_b200.engine import run_model

    case = dict(FORWARD_CASES["s_coordatt_64"], size=640, batch=1)
    cfg, scale, model, sd, x = build(case)
    y1, _ = run_model(model, x.to(dev()), precision="bf16")
    y1 = y1.clone()
    xb = x.repeat(64, 1, 1, 1)
    yb, _ = run_model(model, xb.to(dev()), precision="bf16")
    assert torch.equal(yb[0], yb[63])
    # batch 1 and batch 64 may run a layer on DIFFERENT kernels (strip / im2col, single CTA / CTA pair are chosen by
    # problem size) whose fp32 accumulation order over K differs: equal up to bf16 rounding of a few elements, far below
    # the 2e-2 bf16 tolerance - not bit-equal
    assert rel_l2(yb[0:1].cpu(), y1.cpu()) < 2e-3
    assert torch.isfinite(yb).all()


def test_cpu_tensor_is_rejected():
    from fce_yolo_b200 import modules as M

    m = M.fuse_module(M.Conv(8, 8, 1)).eval()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 8, 4, 4))
