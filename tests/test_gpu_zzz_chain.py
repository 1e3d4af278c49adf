"""(Runs after the other GPU test files: added last, and the files before it keep the order they were last verified in.)
fce_conv1x1_chain (C3k.cv3 chained into C3k2.cv2 in one pass; block.py:338-340 + :303-307, both Conv.forward_fuse
conv.py:80-89) against (a) the two-launch route fce_conv2d + fce_conv2d through a concat buffer, BIT FOR BIT (same MMA
shapes and K order, same bf16 rounding of the intermediate), and (b) a torch fp32 reference of the same two layers with the
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


# (B, H, W, c, n, Cout): c = C3k2's hidden width (x1 and t have c channels, x2 (1 + n) * c).  The m-scale blocks of the
# BASELINE configs (64 -> 256 at 160, 128 -> 512 / 256 at 80), n-scale ones (64 -> 128, 128 -> 256), two inner blocks
# (l / x topology), ragged M (last tile cut), fewer tiles than SMs, one tile, several tiles per CTA
SHAPES = [(2, 160, 160, 64, 1, 256), (2, 80, 80, 128, 1, 512), (2, 80, 80, 128, 1, 256), (3, 40, 40, 64, 1, 128),
          (1, 20, 20, 128, 1, 256), (1, 33, 37, 64, 2, 128), (1, 9, 13, 128, 2, 384), (1, 8, 16, 64, 1, 128),
          (5, 80, 80, 64, 1, 256), (3, 100, 100, 128, 1, 512)]


@pytest.mark.parametrize("B,H,W,c,n,Cout", SHAPES)
@pytest.mark.parametrize("acts", [(1, 1), (0, 0), (1, 2)])
@pytest.mark.parametrize("sliced", [False, True])
def test_chain_equals_two_launches(lib, B, H, W, c, n, Cout, acts, sliced):
    l, L = lib
    act1, act2 = acts
    c2 = (1 + n) * c
    g = torch.Generator().manual_seed(B * 7 + H * 100 + W + c + Cout + n)
    # x1: a channel slice of a wider buffer (C3k's [chain out | cv2(x) | cv1(x)] buffer: pitch 3c/2); x2: the concat buffer
    p1, o1 = (c * 3 // 2, 0) if not sliced else (c * 2, 16)
    p2, o2 = (c2, 0) if not sliced else (c2 + c, 0)   # sliced: the two-launch layout [x2 | t] itself
    op, oo = (Cout, 0) if not sliced else (Cout + 48, 16)
    x1b = torch.randn(B, H, W, p1, generator=g).to(torch.bfloat16).cuda()
    x2b = torch.randn(B, H, W, p2, generator=g).to(torch.bfloat16).cuda()
    w1 = (torch.randn(c, c, generator=g) / c ** 0.5).to(torch.bfloat16).cuda()
    b1 = (torch.randn(c, generator=g) * 0.1).cuda()
    w2 = (torch.randn(Cout, c2 + c, generator=g) / (c2 + c) ** 0.5).to(torch.bfloat16).cuda()
    b2 = (torch.randn(Cout, generator=g) * 0.1).cuda()
    fill = torch.randn(B, H, W, op, generator=g).to(torch.bfloat16).cuda()
    y1, y2 = fill.clone(), fill.clone()
    d = L.ChainDesc(B=B, H=H, W=W, c1=c, cm=c, c2=c2, Cout=Cout, x1_pitch=p1, x1_off=o1, x2_pitch=p2, x2_off=o2,
                    out_pitch=op, out_off=oo, act1=act1, act2=act2)
    assert l.fce_conv1x1_chain_route(C.byref(d)) == 1
    L.check(l.fce_conv1x1_chain(C.byref(d), _p(x1b), _p(w1), _p(b1), _p(x2b), _p(w2), _p(b2), _p(y1), _stream()),
            "fce_conv1x1_chain")
    # the two-launch route: cv3 writes t into the last slice of a dense concat buffer, cv2 reads the whole buffer
    cat = torch.empty(B, H, W, c2 + c, dtype=torch.bfloat16, device="cuda")
    cat[..., :c2] = x2b[..., o2:o2 + c2]

    def conv(x, xoff, cin, pitch, w, b, y, yoff, cout, ypitch, act):
        dc = L.ConvDesc(B=B, H=H, W=W, Cin=cin, Cout=cout, in_pitch=pitch, in_off=xoff, out_pitch=ypitch, out_off=yoff,
                        res_pitch=0, res_off=0, k=1, stride=1, act=act, in_dtype=L.BF16, w_dtype=L.BF16, out_dtype=L.BF16,
                        in_layout=L.NHWC, in_scale=1.0, impl=4)
        L.check(l.fce_conv2d(C.byref(dc), _p(x), _p(w), _p(b), _p(None), _p(y), _stream()), "fce_conv2d")

    conv(x1b, o1, c, p1, w1, b1, cat, c2, c, c2 + c, act1)
    conv(cat, 0, c2 + c, c2 + c, w2, b2, y2, oo, Cout, op, act2)
    torch.cuda.synchronize()
    assert torch.equal(y1[..., :oo], fill[..., :oo]) and torch.equal(y1[..., oo + Cout:], fill[..., oo + Cout:])
    a, b = y1[..., oo:oo + Cout], y2[..., oo:oo + Cout]
    assert torch.isfinite(a.float()).all()
    assert torch.equal(a, b), f"max |diff| {(a.float() - b.float()).abs().max().item()}"
    # torch fp32 reference of the same two layers (bf16 intermediate, as the reference's bf16 forward stores it)
    f = lambda t, act: F.silu(t) if act == 1 else (torch.sigmoid(t) if act == 2 else t)
    x1 = x1b[..., o1:o1 + c].float().reshape(-1, c)
    x2 = x2b[..., o2:o2 + c2].float().reshape(-1, c2)
    t = f(x1 @ w1.float().t() + b1, act1).to(torch.bfloat16).float()
    ref = f(torch.cat([x2, t], 1) @ w2.float().t() + b2, act2)
    l2 = ((a.float().reshape(-1, Cout) - ref).norm() / ref.norm()).item()
    assert l2 < 6e-3, l2


def test_chain_route_rejects(lib):
    """Shapes outside the kernel (the plan compiler then issues the two launches)."""
    l, L = lib
    base = dict(B=1, H=40, W=40, x1_off=0, x2_off=0, out_off=0, act1=1, act2=1)
    for c, n, Cout in [(32, 1, 128), (256, 1, 512), (96, 2, 384), (64, 1, 192), (64, 1, 64)]:
        d = L.ChainDesc(c1=c, cm=c, c2=(1 + n) * c, Cout=Cout, x1_pitch=c, x2_pitch=(1 + n) * c, out_pitch=Cout, **base)
        assert l.fce_conv1x1_chain_route(C.byref(d)) == 0
    d = L.ChainDesc(c1=32, cm=32, c2=64, Cout=128, x1_pitch=32, x2_pitch=64, out_pitch=128, **base)
    x = torch.zeros(1, 40, 40, 128, dtype=torch.bfloat16, device="cuda")
    z = torch.zeros(512, device="cuda")
    assert l.fce_conv1x1_chain(C.byref(d), _p(x), _p(x), _p(z), _p(x), _p(x), _p(z), _p(x), _stream()) == -2


@pytest.mark.parametrize("yaml,size,batch", [("yolo11m-bifpn.yaml", 320, 4), ("yolo11n-fce.yaml", 256, 3)])
def test_plan_with_chain_matches_two_launch_plan(yaml, size, batch):
    """The whole predict plan with fce_conv1x1_chain in it against the same plan with Plan.FUSED_C3K_TAIL off: the
    prediction tensor and the detections are BIT-identical (the fused kernel reproduces both launches exactly)."""
    from fce_yolo_b200.plan import Plan
    from fce_yolo_b200.predict import Predictor
    from fce_yolo_b200.tasks import DetectionModel
    from fce_yolo_b200.weights import load_synthetic, synth_images

    model = DetectionModel(yaml).fuse().eval()
    load_synthetic(model, 2)
    img = (synth_images(11, batch, size, size) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous().pin_memory()
    saved = Plan.FUSED_C3K_TAIL
    out = {}
    try:
        for flag in (True, False):
            Plan.FUSED_C3K_TAIL = flag
            pred = Predictor(model, batch, size, precision="bf16", conf=0.25, iou=0.7, input_u8=True, use_graph=False)
            det, cnt = [t.clone() for t in pred.infer(img)]
            fns = [n.fn for n in pred.ex.plan.nodes]
            assert ("fce_conv1x1_chain" in fns) == flag
            out[flag] = (pred.ex.outputs()[0].float().cpu().clone(), det, cnt, len(fns))
            del pred
    finally:
        Plan.FUSED_C3K_TAIL = saved
    assert out[True][3] < out[False][3]  # fewer launches
    assert torch.isfinite(out[True][0]).all()
    assert torch.equal(out[True][0], out[False][0]) and torch.equal(out[True][1], out[False][1])
    assert torch.equal(out[True][2], out[False][2]) and int(out[True][2].sum()) > 0
