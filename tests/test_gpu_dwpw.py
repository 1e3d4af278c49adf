"""fce_dwpw_conv (depthwise 3x3 -> 1x1 in one pass; head.py:101-102 = DWConv conv.py:185-199 + Conv conv.py:80-89) against
(a) the two-launch route fce_dwconv3x3 + fce_conv2d, BIT FOR BIT (same fp32 operation order in the depthwise part, same
bf16 rounding of the intermediate, same MMA sequence), and (b) a torch fp32 reference of the same two ops, with the
tolerance of one bf16 rounding per layer."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from fce_yolo_b200 import _lib as L
    return L.load(check_device=True), L


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr() if t is not None else 0)


# (B, H, W, C, Cout): Detect class-branch shapes of the BASELINE configs (m: 256 -> 256 at 80 / 40 / 20, s: 128 / 256 / 512
# -> 128, n first block: 64 -> 80), ragged maps (tiles cut by both borders), one unit, many units per CTA
SHAPES = [(2, 80, 80, 256, 256), (3, 40, 40, 256, 256), (2, 20, 20, 256, 256), (2, 80, 80, 128, 128), (2, 40, 40, 256, 128),
          (3, 20, 20, 512, 128), (1, 80, 80, 64, 80), (1, 33, 37, 64, 48), (2, 7, 13, 128, 16), (1, 5, 3, 64, 256),
          (40, 40, 40, 128, 128), (1, 160, 160, 64, 64),
          # 64 Ki weight elements (the largest the kernel parks: 7-row units, three input stages), an odd number of units,
          # one column tile
          (1, 8, 48, 256, 256), (3, 9, 16, 512, 128),
          # weights too large to park (m-scale P4 / P5 first block, 256 KB): streamed chunk by chunk with the input
          (2, 40, 40, 512, 256), (3, 20, 20, 512, 256), (1, 24, 40, 768, 192)]


@pytest.mark.parametrize("B,H,W,Cc,Cout", SHAPES)
@pytest.mark.parametrize("acts", [(1, 1), (0, 0), (1, 2)])
@pytest.mark.parametrize("sliced", [False, True])
def test_dwpw_equals_two_launches(lib, B, H, W, Cc, Cout, acts, sliced):
    l, L = lib
    dw_act, pw_act = acts
    g = torch.Generator().manual_seed(B * 7 + H * 100 + W + Cc + Cout)
    ip, io = (Cc + 32, 16) if sliced else (Cc, 0)
    op, oo = (Cout + 48, 16) if sliced else (Cout, 0)
    xb = torch.randn(B, H, W, ip, generator=g).to(torch.bfloat16).cuda()
    wd = (torch.randn(Cc, 3, 3, generator=g) * 0.3)
    bd = torch.randn(Cc, generator=g) * 0.1
    wp = (torch.randn(Cout, Cc, generator=g) / Cc ** 0.5).to(torch.bfloat16)
    bp = torch.randn(Cout, generator=g) * 0.1
    wd_d, bd_d, wp_d, bp_d = wd.view(Cc, 9).t().contiguous().cuda(), bd.cuda(), wp.cuda(), bp.cuda()
    fill = torch.randn(B, H, W, op, generator=g).to(torch.bfloat16).cuda()
    y1, y2 = fill.clone(), fill.clone()
    d = L.DwpwDesc(B=B, H=H, W=W, C=Cc, Cout=Cout, in_pitch=ip, in_off=io, out_pitch=op, out_off=oo, dw_act=dw_act,
                   pw_act=pw_act)
    assert l.fce_dwpw_route(C.byref(d)) == 1
    L.check(l.fce_dwpw_conv(C.byref(d), _p(xb), _p(wd_d), _p(bd_d), _p(wp_d), _p(bp_d), _p(y1), _stream()), "fce_dwpw_conv")
    # the two-launch route through a dense bf16 intermediate
    mid = torch.empty(B, H, W, Cc, dtype=torch.bfloat16, device="cuda")
    dd = L.DwconvDesc(B=B, H=H, W=W, C=Cc, in_pitch=ip, in_off=0, out_pitch=Cc, out_off=0, add_pitch=0, add_off=0,
                      act=dw_act, dtype=L.BF16)
    L.check(l.fce_dwconv3x3(C.byref(dd), C.c_void_p(xb.data_ptr() + io * 2), _p(wd_d), _p(bd_d), _p(None), _p(mid),
                            _stream()), "fce_dwconv3x3")
    dc = L.ConvDesc(B=B, H=H, W=W, Cin=Cc, Cout=Cout, in_pitch=Cc, in_off=0, out_pitch=op, out_off=oo, res_pitch=0,
                    res_off=0, k=1, stride=1, act=pw_act, in_dtype=L.BF16, w_dtype=L.BF16, out_dtype=L.BF16,
                    in_layout=L.NHWC, in_scale=1.0, impl=2)
    L.check(l.fce_conv2d(C.byref(dc), _p(mid), _p(wp_d), _p(bp_d), _p(None), _p(y2), _stream()), "fce_conv2d")
    torch.cuda.synchronize()
    assert torch.equal(y1[..., :oo], fill[..., :oo]) and torch.equal(y1[..., oo + Cout:], fill[..., oo + Cout:])
    a, b = y1[..., oo:oo + Cout], y2[..., oo:oo + Cout]
    assert torch.isfinite(a.float()).all()
    assert torch.equal(a, b), f"max |diff| {(a.float() - b.float()).abs().max().item()}"
    # torch fp32 reference of the same two layers (bf16 intermediate as stored by the reference's bf16 forward)
    x = xb[..., io:io + Cc].float().cpu().permute(0, 3, 1, 2)
    t = F.conv2d(x, wd.view(Cc, 1, 3, 3), bd, padding=1, groups=Cc)
    t = (F.silu(t) if dw_act == 1 else t).to(torch.bfloat16).float()
    ref = F.conv2d(t, wp.float().view(Cout, Cc, 1, 1), bp)
    ref = F.silu(ref) if pw_act == 1 else (torch.sigmoid(ref) if pw_act == 2 else ref)
    out = a.float().cpu().permute(0, 3, 1, 2)
    l2 = ((out - ref).norm() / ref.norm()).item()
    assert l2 < 6e-3, l2


def test_dwpw_route_rejects(lib):
    """Shapes outside the kernel (the plan compiler then issues the two launches): channel counts off the 64 / 16 grid,
    more than 256 outputs."""
    l, L = lib
    base = dict(B=1, H=40, W=40, in_off=0, out_off=0, dw_act=1, pw_act=1)
    for Cc, Cout in [(80, 80), (64, 40), (384, 384), (768, 384), (512, 264)]:
        d = L.DwpwDesc(C=Cc, Cout=Cout, in_pitch=Cc, out_pitch=Cout, **base)
        assert l.fce_dwpw_route(C.byref(d)) == 0
    d = L.DwpwDesc(C=256, Cout=256, in_pitch=256, out_pitch=256, **base)
    assert l.fce_dwpw_route(C.byref(d)) == 1
    x = torch.zeros(1, 40, 40, 80, dtype=torch.bfloat16, device="cuda")
    d = L.DwpwDesc(C=80, Cout=80, in_pitch=80, out_pitch=80, **base)
    z = torch.zeros(9 * 80, device="cuda")
    assert l.fce_dwpw_conv(C.byref(d), _p(x), _p(z), _p(z), _p(x), _p(z), _p(x), _stream()) == -2


@pytest.mark.parametrize("yaml,size,batch", [("yolo11m-bifpn.yaml", 320, 4), ("yolo11s-fce.yaml", 256, 4)])
def test_plan_with_fused_producers_matches_two_launch_plan(yaml, size, batch):
    """The whole predict plan with fce_stem2_conv / fce_dwpw_conv in it against the same plan with the two-launch routes
    (Plan.FUSED_STEM2 / FUSED_DWPW off): fce_dwpw_conv alone leaves the prediction tensor BIT-identical; the fused stem pair
    changes a stem value by one bf16 ulp now and then, which this deep synthetic network amplifies like any other bf16
    rounding - so the yardstick is the fp32 oracle: the fused plan is no further from it than the two-launch plan."""
    from fce_yolo_b200.plan import Plan
    from fce_yolo_b200.predict import Predictor
    from fce_yolo_b200.tasks import DetectionModel
    from fce_yolo_b200.weights import load_synthetic, synth_images
    from oracle import fce_oracle as O

    model = DetectionModel(yaml).fuse().eval()
    sd = load_synthetic(model, 2)
    img = (synth_images(11, batch, size, size) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous().pin_memory()
    saved = (Plan.FUSED_STEM2, Plan.FUSED_DWPW)
    out = {}
    try:
        for key, (s2, dw) in {"two": (False, False), "dwpw": (False, True), "both": (True, True)}.items():
            Plan.FUSED_STEM2, Plan.FUSED_DWPW = s2, dw
            pred = Predictor(model, batch, size, precision="bf16", conf=0.25, iou=0.7, input_u8=True)
            det, cnt = [t.clone() for t in pred.infer(img)]
            fns = [n.fn for n in pred.ex.plan.nodes]
            assert ("fce_stem2_conv" in fns) == s2 and ("fce_dwpw_conv" in fns) == dw
            out[key] = (pred.ex.outputs()[0].float().cpu().clone(), det, cnt, len(fns))
            del pred
    finally:
        Plan.FUSED_STEM2, Plan.FUSED_DWPW = saved
    assert out["both"][3] < out["dwpw"][3] < out["two"][3]  # fewer launches
    assert torch.equal(out["dwpw"][0], out["two"][0]) and torch.equal(out["dwpw"][1], out["two"][1])
    yo = O.forward(model.yaml, model.yaml["scale"], sd, img.permute(0, 3, 1, 2).float() / 255.0)[0].double()
    err = {k: ((out[k][0].double() - yo).norm() / yo.norm()).item() for k in ("two", "both")}
    assert err["both"] <= 1.25 * err["two"] + 1e-3, err
    n0, n1 = int(out["two"][2].sum()), int(out["both"][2].sum())
    assert n0 > 0 and abs(n0 - n1) <= max(2, n0 // 20)
