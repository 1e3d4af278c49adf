"""fce_stem2_conv (the first two convs of the graph in one pass: Conv(3, 64, 3, 2) on the uint8 image + Conv(64, C1, 3, 2),
yolo11-fce.yaml:20-21, conv.py:80-89) against (a) the two-launch route fce_stem_conv + fce_conv2d: same operands and the
same bf16 rounding of the stem map, but the stem GEMM runs on tcgen05 instead of mma.sync (its bias enters as an fp16 K
column) and the second conv accumulates its taps plane by plane - a stem value now and then rounds to the neighbouring
bf16, so the outputs agree to rounding noise, not bit for bit: relative L2 <= 2e-3, no output off by more than 2 % of the
largest magnitude; and (b) a torch fp32 reference of the two layers with the tolerance of two bf16 roundings (the same bound
the two-launch route meets)."""
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


# (B, H, W, C0, C1): the m-scale pair (64 -> 128) and the s-scale pair (32 -> 64) at full size (one image: 230 units), small maps
# (one unit, ragged units on both borders), H / W multiples of 4 only, several units per CTA (batch 40 at 160^2: 600 units),
# narrow and wide outputs
SHAPES = [(1, 640, 640, 64, 128), (2, 64, 64, 64, 128), (1, 96, 160, 64, 128), (1, 68, 36, 64, 64), (3, 32, 32, 64, 16),
          (1, 4, 4, 64, 32), (40, 160, 160, 64, 128), (2, 128, 256, 64, 96),
          (1, 640, 640, 32, 64), (2, 64, 64, 32, 64), (1, 68, 36, 32, 48), (40, 160, 160, 32, 64), (2, 96, 128, 32, 192)]


@pytest.mark.parametrize("B,H,W,C0,C1", SHAPES)
@pytest.mark.parametrize("acts", [(1, 1), (0, 0)])
@pytest.mark.parametrize("sliced", [False, True])
def test_stem2_equals_two_launches(lib, B, H, W, C0, C1, acts, sliced):
    l, L = lib
    act0, act1 = acts
    g = torch.Generator().manual_seed(B * 7 + H * 100 + W + C1)
    x = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)
    w0 = torch.randn(C0, 3, 3, 3, generator=g) * 0.3
    b0 = torch.randn(C0, generator=g) * 0.2
    w1 = (torch.randn(C1, C0, 3, 3, generator=g) / (9 * C0) ** 0.5)
    b1 = torch.randn(C1, generator=g) * 0.1
    wk = torch.zeros(C0, 32)
    wk[:, :27] = w0.permute(0, 2, 3, 1).reshape(C0, 27) / 255.0
    wk_d, b0_d = wk.bfloat16().cuda(), b0.cuda()
    w1_d, b1_d = w1.permute(0, 2, 3, 1).contiguous().bfloat16().cuda(), b1.cuda()  # OHWI
    xd = x.cuda()
    H0, W0, H1, W1 = H // 2, W // 2, H // 4, W // 4
    op, oo = (C1 + 48, 16) if sliced else (C1, 0)
    fill = torch.randn(B, H1, W1, op, generator=g).to(torch.bfloat16).cuda()
    y1, y2 = fill.clone(), fill.clone()
    d = L.Stem2Desc(B=B, H=H, W=W, C0=C0, C1=C1, out_pitch=op, out_off=oo, act0=act0, act1=act1)
    assert l.fce_stem2_route(C.byref(d)) == 1
    L.check(l.fce_stem2_conv(C.byref(d), _p(xd), _p(wk_d), _p(b0_d), _p(w1_d), _p(b1_d), _p(y1), _stream()), "fce_stem2_conv")
    # two launches through a dense stem map
    mid = torch.empty(B, H0, W0, C0, dtype=torch.bfloat16, device="cuda")
    ds = L.StemDesc(B=B, H=H, W=W, Cout=C0, out_pitch=C0, out_off=0, act=act0, in_dtype=L.U8, in_layout=L.NHWC)
    L.check(l.fce_stem_conv(C.byref(ds), _p(xd), _p(wk_d), _p(b0_d), _p(mid), _stream()), "fce_stem_conv")
    dc = L.ConvDesc(B=B, H=H0, W=W0, Cin=C0, Cout=C1, in_pitch=C0, in_off=0, out_pitch=op, out_off=oo, res_pitch=0,
                    res_off=0, k=3, stride=2, act=act1, in_dtype=L.BF16, w_dtype=L.BF16, out_dtype=L.BF16,
                    in_layout=L.NHWC, in_scale=1.0, impl=2)
    L.check(l.fce_conv2d(C.byref(dc), _p(mid), _p(w1_d), _p(b1_d), _p(None), _p(y2), _stream()), "fce_conv2d")
    torch.cuda.synchronize()
    assert torch.equal(y1[..., :oo], fill[..., :oo]) and torch.equal(y1[..., oo + C1:], fill[..., oo + C1:])
    a, b = y1[..., oo:oo + C1], y2[..., oo:oo + C1]
    assert torch.isfinite(a.float()).all()
    af, bfl = a.float(), b.float()
    assert ((af - bfl).norm() / bfl.norm()).item() < 2e-3
    assert ((af - bfl).abs().max() / bfl.abs().max()).item() < 2e-2
    # torch fp32 reference (bf16 operands, bf16 stem map)
    xr = x.float().permute(0, 3, 1, 2)
    w0r = wk_d.float().cpu()[:, :27].reshape(C0, 3, 3, 3).permute(0, 3, 1, 2)
    t = F.conv2d(xr, w0r, b0, stride=2, padding=1)
    t = (F.silu(t) if act0 == 1 else t).to(torch.bfloat16).float()
    ref = F.conv2d(t, w1_d.float().cpu().permute(0, 3, 1, 2), b1, stride=2, padding=1)
    ref = F.silu(ref) if act1 == 1 else ref
    out = a.float().cpu().permute(0, 3, 1, 2)
    l2 = ((out - ref).norm() / ref.norm()).item()
    l2_two = ((b.float().cpu().permute(0, 3, 1, 2) - ref).norm() / ref.norm()).item()
    assert l2 < 6e-3, (l2, l2_two)
    assert l2 < 1.25 * l2_two + 1e-4, (l2, l2_two)  # no less accurate than the two launches


def test_stem2_route_rejects(lib):
    l, L = lib
    for kw in [dict(C0=16, C1=32), dict(C0=96, C1=192), dict(C0=64, C1=192), dict(C0=32, C1=256), dict(C0=64, C1=40),
               dict(C0=64, C1=128, H=66)]:
        args = dict(B=1, H=64, W=64, C0=64, C1=128, out_pitch=256, out_off=0, act0=1, act1=1)
        args.update(kw)
        assert l.fce_stem2_route(C.byref(L.Stem2Desc(**args))) == 0
    d = L.Stem2Desc(B=1, H=64, W=64, C0=64, C1=128, out_pitch=128, out_off=0, act0=1, act1=1)
    assert l.fce_stem2_route(C.byref(d)) == 1
