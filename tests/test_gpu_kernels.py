"""Stand-alone parity of individual sm_100a kernels through the C ABI against a torch fp32 reference of the same
op (floating-point kernels keep a torch reference next to the oracle).  Tolerances are written per test."""
import ctypes as C
import math

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


# ------------------------------------------------------------------------------------------------ fused stem
@pytest.mark.parametrize("cout", [16, 32, 48, 64, 96])
@pytest.mark.parametrize("kind", ["u8_nhwc", "f32_nchw", "f32_nhwc"])
@pytest.mark.parametrize("shape", [(2, 64, 64), (1, 96, 160), (3, 640, 640), (1, 34, 70)])
def test_stem_conv_vs_torch(lib, cout, kind, shape):
    """fce_stem_conv = SiLU(conv3x3/s2/p1(x * scale) + b) with bf16 operands, fp32 accumulate, bf16 output:
    max |err| <= 1e-2 * max|ref|, relL2 <= 4e-3 (one bf16 rounding of the output + __expf SiLU)."""
    l, L = lib
    B, H, W = shape
    g = torch.Generator().manual_seed(cout * 1000 + H + W)
    w = (torch.randn(cout, 3, 3, 3, generator=g) * 0.3)
    bias = torch.randn(cout, generator=g) * 0.2
    if kind == "u8_nhwc":
        x = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8)
        scale, x_ref = 1.0 / 255.0, x.float().permute(0, 3, 1, 2)
        dt, lay = L.U8, L.NHWC
    else:
        xr = torch.rand(B, 3, H, W, generator=g)
        x = xr if kind == "f32_nchw" else xr.permute(0, 2, 3, 1).contiguous()
        scale, x_ref = 1.0, xr.bfloat16().float()
        dt, lay = L.F32, (L.NCHW if kind == "f32_nchw" else L.NHWC)
    wk = torch.zeros(cout, 32)
    wk[:, :27] = w.permute(0, 2, 3, 1).reshape(cout, 27) * scale
    wk = wk.bfloat16().cuda()
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    pitch, off = cout + 16, 8  # destination is a channel slice of a wider buffer
    y = torch.full((B, Ho, Wo, pitch), 7.0, dtype=torch.bfloat16, device="cuda")
    d = L.StemDesc(B=B, H=H, W=W, Cout=cout, out_pitch=pitch, out_off=0, act=L.ACT_SILU, in_dtype=dt, in_layout=lay)
    xd, bd = x.cuda(), bias.cuda()
    st = l.fce_stem_conv(C.byref(d), C.c_void_p(xd.data_ptr()), C.c_void_p(wk.data_ptr()), C.c_void_p(bd.data_ptr()),
                         C.c_void_p(y.data_ptr() + off * 2), _stream())
    L.check(st, "fce_stem_conv")
    torch.cuda.synchronize()
    w_ref = wk.float().cpu()[:, :27].reshape(cout, 3, 3, 3).permute(0, 3, 1, 2)
    ref = F.silu(F.conv2d(x_ref, w_ref, bias, stride=2, padding=1))
    out = y[..., off:off + cout].float().cpu().permute(0, 3, 1, 2)
    assert ((out - ref).abs().max() / ref.abs().max()).item() < 1e-2
    assert ((out - ref).norm() / ref.norm()).item() < 4e-3
    assert bool((y[..., :off] == 7.0).all() and (y[..., off + cout:] == 7.0).all())


def test_stem_conv_rejects_bad_args(lib):
    l, L = lib
    d = L.StemDesc(B=1, H=64, W=64, Cout=20, out_pitch=20, out_off=0, act=L.ACT_SILU, in_dtype=L.U8, in_layout=L.NHWC)
    t = torch.zeros(64 * 64 * 32, device="cuda")
    p = C.c_void_p(t.data_ptr())
    assert l.fce_stem_conv(C.byref(d), p, p, p, p, _stream()) == -2
    d.Cout, d.out_pitch = 32, 36
    assert l.fce_stem_conv(C.byref(d), p, p, p, p, _stream()) == -3
    assert l.fce_stem_conv(C.byref(d), None, p, p, p, _stream()) == -1


# ------------------------------------------------------------------------------------------------ C2PSA attention
@pytest.mark.parametrize("B,N,heads", [(2, 400, 4), (1, 1600, 2), (3, 4, 2), (2, 100, 1), (1, 81, 6), (64, 400, 4)])
@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_psa_attention_vs_torch(lib, B, N, heads, dtype):
    """out[:, n] = sum_m softmax_m(scale <q_n, k_m>) v_m  (block.py:1293-1302) on [Q|K|V]-ordered NHWC channels.
    fp32 kernel: relL2 <= 1e-5.  bf16 kernel (mma.sync, probabilities rounded to bf16 before P V): relL2 <= 1e-2
    against fp32 math on the same bf16 inputs."""
    l, L = lib
    kd, hd = 32, 64
    pitch = heads * (2 * kd + hd) + 8
    g = torch.Generator().manual_seed(N * 10 + heads)
    qkv = torch.randn(B, N, pitch, generator=g)
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    qkv_d = qkv.to(tdt).cuda()
    out_pitch = heads * hd + 8
    out = torch.full((B, N, out_pitch), 3.0, dtype=tdt, device="cuda")
    scale = kd ** -0.5
    d = L.PsaDesc(B=B, N=N, heads=heads, kd=kd, hd=hd, qkv_pitch=pitch, q_off=0, k_off=heads * kd, v_off=2 * heads * kd,
                  out_pitch=out_pitch, out_off=8, dtype=L.BF16 if dtype == "bf16" else L.F32, scale=scale)
    st = l.fce_psa_attention(C.byref(d), C.c_void_p(qkv_d.data_ptr()), C.c_void_p(out.data_ptr()), _stream())
    L.check(st, "fce_psa_attention")
    torch.cuda.synchronize()
    x = qkv_d.float()
    q = x[..., :heads * kd].view(B, N, heads, kd).transpose(1, 2)
    k = x[..., heads * kd:2 * heads * kd].view(B, N, heads, kd).transpose(1, 2)
    v = x[..., 2 * heads * kd:2 * heads * kd + heads * hd].view(B, N, heads, hd).transpose(1, 2)
    att = torch.softmax((q @ k.transpose(-1, -2)) * scale, dim=-1)
    ref = (att @ v).transpose(1, 2).reshape(B, N, heads * hd)
    got = out[..., 8:].float()
    err = ((got - ref).norm() / ref.norm()).item()
    assert err < (1e-2 if dtype == "bf16" else 1e-5), err
    assert bool((out[..., :8] == 3.0).all())
    assert math.isfinite(err)


# ------------------------------------------------------------------------------------------------ depthwise 3x3
@pytest.mark.parametrize("B,H,W,Cc", [(2, 80, 80, 128), (1, 20, 20, 512), (3, 7, 13, 64), (1, 40, 40, 72), (2, 33, 50, 256)])
@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
@pytest.mark.parametrize("with_add,act", [(False, 1), (True, 0)])
def test_dwconv3x3_vs_torch(lib, B, H, W, Cc, dtype, with_add, act):
    """fce_dwconv3x3 (DWConv conv.py:185-199; Attention.pe + add block.py:1302) reading and writing channel
    slices of wider buffers; the add operand aliases the destination (as in the C2PSA plan).
    fp32: relL2 <= 1e-5;  bf16 storage (fp32 math, one output rounding): relL2 <= 4e-3, max <= 1e-2."""
    l, L = lib
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    g = torch.Generator().manual_seed(H * 100 + W + Cc)
    ip, io, op, oo = Cc + 16, 8, Cc + 24, 16
    xb = torch.randn(B, H, W, ip, generator=g).to(tdt).cuda()
    yb = torch.randn(B, H, W, op, generator=g).to(tdt).cuda()
    y0 = yb.clone()
    w = (torch.randn(Cc, 3, 3, generator=g) * 0.3)
    bias = torch.randn(Cc, generator=g) * 0.1
    wp = w.view(Cc, 9).t().contiguous().cuda()
    bp = bias.cuda()
    esz = xb.element_size()
    d = L.DwconvDesc(B=B, H=H, W=W, C=Cc, in_pitch=ip, in_off=0, out_pitch=op, out_off=0, add_pitch=op if with_add else 0,
                     add_off=0, act=act, dtype=L.BF16 if dtype == "bf16" else L.F32)
    yptr = yb.data_ptr() + oo * esz
    st = l.fce_dwconv3x3(C.byref(d), C.c_void_p(xb.data_ptr() + io * esz), C.c_void_p(wp.data_ptr()),
                         C.c_void_p(bp.data_ptr()), C.c_void_p(yptr if with_add else 0), C.c_void_p(yptr), _stream())
    L.check(st, "fce_dwconv3x3")
    torch.cuda.synchronize()
    x = xb[..., io:io + Cc].float().cpu().permute(0, 3, 1, 2)
    ref = F.conv2d(x, w.view(Cc, 1, 3, 3), bias, padding=1, groups=Cc)
    if act == 1:
        ref = F.silu(ref)
    if with_add:
        ref = ref + y0[..., oo:oo + Cc].float().cpu().permute(0, 3, 1, 2)
    out = yb[..., oo:oo + Cc].float().cpu().permute(0, 3, 1, 2)
    l2 = ((out - ref).norm() / ref.norm()).item()
    mx = ((out - ref).abs().max() / ref.abs().max()).item()
    assert l2 < (4e-3 if dtype == "bf16" else 1e-5), l2
    assert mx < (1e-2 if dtype == "bf16" else 1e-5), mx
    assert torch.equal(yb[..., :oo], y0[..., :oo]) and torch.equal(yb[..., oo + Cc:], y0[..., oo + Cc:])


# ------------------------------------------------------------------------------------------------ SPPF pyramid
@pytest.mark.parametrize("B,H,W,Cc", [(2, 20, 20, 256), (1, 40, 40, 384), (3, 2, 2, 128), (1, 5, 9, 40), (1, 80, 80, 32)])
@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_sppf_pool_bit_exact(lib, B, H, W, Cc, dtype):
    """fce_sppf_pool: slices 1..3 = three chained 5x5/s1/p2 max-pools of slice 0 (block.py:228-232).  max() is
    exact in every precision: the result must equal torch's max_pool2d BIT FOR BIT, and slice 0 must be untouched.
    40 x 40 is the P5 map of a 1280^2 image (BASELINE config 5)."""
    l, L = lib
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    g = torch.Generator().manual_seed(H * 100 + W + Cc)
    pitch, off = 4 * Cc + 16, 8
    buf = torch.randn(B, H, W, pitch, generator=g).to(tdt).cuda()
    b0 = buf.clone()
    d = L.SppfDesc(B=B, H=H, W=W, C=Cc, pitch=pitch, off=0, dtype=L.BF16 if dtype == "bf16" else L.F32)
    st = l.fce_sppf_pool(C.byref(d), C.c_void_p(buf.data_ptr() + off * buf.element_size()), _stream())
    L.check(st, "fce_sppf_pool")
    torch.cuda.synchronize()
    y = b0[..., off:off + Cc].float().permute(0, 3, 1, 2)
    for level in (1, 2, 3):
        y = F.max_pool2d(y, 5, 1, 2)
        got = buf[..., off + level * Cc:off + (level + 1) * Cc].float().permute(0, 3, 1, 2)
        assert torch.equal(got, y), level
    assert torch.equal(buf[..., :off + Cc], b0[..., :off + Cc])
    assert torch.equal(buf[..., off + 4 * Cc:], b0[..., off + 4 * Cc:])


# ------------------------------------------------------------------------------------------------ CoordAtt gate MLP
@pytest.mark.parametrize("rows_h,rows_w,Cc,mip,oup", [(160, 160, 256, 16, 256), (80, 40, 128, 8, 128), (5, 9, 64, 11, 96), (3, 300, 768, 32, 64),
                                                      (33, 1, 512, 32, 512), (64 * 80, 64 * 80, 256, 16, 256),
                                                      (7, 20, 768, 23, 384)])
def test_coordatt_mlp_vs_torch(lib, rows_h, rows_w, Cc, mip, oup):
    """fce_coordatt_mlp: a = sigmoid(W2 SiLU(W1 s + b1) + b2) per strip row, (W2, b2) switching from cv_h to cv_w at
    row rows_h (fce_block.py:104-113).  fp32 throughout: 1e-5 relative to the largest output."""
    l, L = lib
    g = torch.Generator().manual_seed(rows_h + 7 * rows_w + Cc + mip)
    rows = rows_h + rows_w
    s_pitch, o_pitch = Cc + 8, oup + 4
    s = torch.randn(rows, s_pitch, generator=g).cuda()
    w1, b1 = torch.randn(mip, Cc, generator=g).cuda() / math.sqrt(Cc), torch.randn(mip, generator=g).cuda()
    wh, bh = torch.randn(oup, mip, generator=g).cuda() / math.sqrt(mip), torch.randn(oup, generator=g).cuda()
    ww, bw = torch.randn(oup, mip, generator=g).cuda() / math.sqrt(mip), torch.randn(oup, generator=g).cuda()
    out = torch.full((rows, o_pitch), -7.0).cuda()
    d = L.CoordAttMlpDesc(rows_h=rows_h, rows_w=rows_w, C=Cc, mip=mip, oup=oup, s_pitch=s_pitch, out_pitch=o_pitch,
                          act1=L.ACT_SILU, act2=L.ACT_SIGMOID)
    w1t = w1.view(mip, Cc // 4, 4).permute(1, 0, 2).contiguous()  # [C/4][mip][4]
    wht, wwt = wh.t().contiguous(), ww.t().contiguous()
    st = l.fce_coordatt_mlp(C.byref(d), *[C.c_void_p(t.data_ptr()) for t in (s, w1t, b1, wht, bh, wwt, bw, out)],
                            _stream())
    L.check(st, "fce_coordatt_mlp")
    torch.cuda.synchronize()
    y = F.silu(s[:, :Cc].double() @ w1.double().t() + b1.double())
    ref = torch.cat([torch.sigmoid(y[:rows_h] @ wh.double().t() + bh.double()),
                     torch.sigmoid(y[rows_h:] @ ww.double().t() + bw.double())]).float()
    assert (out[:, :oup] - ref).abs().max().item() <= 1e-5
    assert torch.all(out[:, oup:] == -7.0)  # the pitch padding is not written


def test_coordatt_mlp_rejects_bad_args(lib):
    l, L = lib
    t = torch.zeros(64, 64).cuda()
    p = C.c_void_p(t.data_ptr())
    d = L.CoordAttMlpDesc(rows_h=4, rows_w=4, C=30, mip=8, oup=32, s_pitch=32, out_pitch=32, act1=1, act2=2)
    assert l.fce_coordatt_mlp(C.byref(d), p, p, p, p, p, p, p, p, _stream()) == -3  # C not a multiple of 4
    d.C, d.mip = 32, 65
    assert l.fce_coordatt_mlp(C.byref(d), p, p, p, p, p, p, p, p, _stream()) == -2  # hidden width above the kernel's limit
    d.mip = 8
    d.C = 32
    assert l.fce_coordatt_mlp(C.byref(d), p, p, None, p, p, p, p, p, _stream()) == -1


# ------------------------------------------------------------------------------------------------ gate application
@pytest.mark.parametrize("B,H,W,Cc", [(2, 80, 80, 256), (3, 40, 40, 512), (1, 13, 27, 128), (2, 20, 20, 64),
                                      (1, 7, 5, 1024), (2, 10, 12, 40)])
@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_gate_apply_vs_torch(lib, B, H, W, Cc, mode, dtype):
    """fce_gate_apply (fce_block.py:116, 180, 283-284) on both launch shapes (column walk for wide channel counts,
    row-per-CTA otherwise), pitched views.  fp32: 1e-6 relative; bf16: one output rounding (2^-8) + fast sigmoid."""
    l, L = lib
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    g = torch.Generator().manual_seed(B + 3 * H + 5 * W + Cc + mode)
    pitch, off = Cc + 16, 8
    xb = torch.randn(B, H, W, pitch, generator=g).to(tdt).cuda()
    yb = torch.full((B, H, W, pitch), 3.0).to(tdt).cuda()
    gh = torch.rand(B, H, Cc, generator=g).cuda() if mode != 2 else torch.randn(B, H, Cc, generator=g).cuda()
    gw = torch.rand(B, W, Cc, generator=g).cuda() if mode != 2 else torch.randn(B, W, Cc, generator=g).cuda()
    d = L.GateDesc(B=B, H=H, W=W, C=Cc, mode=mode, in_pitch=pitch, in_off=off, out_pitch=pitch, out_off=off,
                   dtype=L.BF16 if dtype == "bf16" else L.F32, gh_bstride=H * Cc, gh_rstride=Cc, gw_bstride=W * Cc,
                   gw_rstride=Cc)
    st = l.fce_gate_apply(C.byref(d), C.c_void_p(xb.data_ptr()), C.c_void_p(gh.data_ptr()),
                          C.c_void_p(gw.data_ptr() if mode != 1 else 0), C.c_void_p(yb.data_ptr()), _stream())
    L.check(st, "fce_gate_apply")
    torch.cuda.synchronize()
    x = xb[..., off:off + Cc].float()
    a, b = gh[:, :, None, :], gw[:, None, :, :]
    ref = x * a * b if mode == 0 else (x * a if mode == 1 else x * torch.sigmoid(a + b))
    got = yb[..., off:off + Cc].float()
    tol = 1e-6 if dtype == "fp32" and mode != 2 else (2e-6 if dtype == "fp32" else 2 ** -7)
    assert ((got - ref).abs() <= tol * ref.abs() + 1e-6).all()
    assert torch.all(yb[..., :off].float() == 3.0) and torch.all(yb[..., off + Cc:].float() == 3.0)


# ------------------------------------------------------------------------------------------------ coordinate pooling
@pytest.mark.parametrize("B,H,W,Cc", [(2, 80, 80, 256), (64, 40, 40, 512), (3, 20, 20, 512), (1, 160, 160, 64),
                                      (2, 13, 27, 128), (1, 7, 100, 40), (5, 33, 50, 256),
                                      # large (image, chunk) counts: the TMA-fed kernel (bf16) - every (RB, NCOL) variant,
                                      # ragged last band, partial last channel chunk, W not a multiple of 16
                                      (64, 80, 80, 128), (32, 160, 160, 64), (48, 20, 20, 256), (40, 33, 50, 200),
                                      (100, 7, 100, 40), (128, 40, 40, 64)])
@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_coord_pool_vs_torch(lib, B, H, W, Cc, dtype):
    """fce_coord_pool: strip[b, 0:H] = mean over W, strip[B*H + b*W ...] = mean over H (fce_block.py:101-102), every
    channel-group width (8 / 16 / 32 vector lanes) and both band modes (single band, bands + finish kernel).
    fp32 accumulation in a different order than torch: 2e-6 relative to the largest mean."""
    l, L = lib
    tdt = torch.bfloat16 if dtype == "bf16" else torch.float32
    g = torch.Generator().manual_seed(B + 3 * H + 5 * W + Cc)
    pitch, off = Cc + 16, 8
    xb = (torch.randn(B, H, W, pitch, generator=g) + 0.5).to(tdt).cuda()
    strip = torch.full((B * (H + W), Cc), -5.0).cuda()
    d = L.PoolDesc(B=B, H=H, W=W, C=Cc, pitch=pitch, off=0, dtype=L.BF16 if dtype == "bf16" else L.F32)
    nws = l.fce_coord_pool_workspace(C.byref(d))
    ws = torch.empty(max(nws, 16), dtype=torch.uint8).cuda()
    st = l.fce_coord_pool(C.byref(d), C.c_void_p(xb.data_ptr() + off * xb.element_size()), C.c_void_p(strip.data_ptr()),
                          C.c_void_p(ws.data_ptr()), C.c_size_t(nws), _stream())
    L.check(st, "fce_coord_pool")
    torch.cuda.synchronize()
    x = xb[..., off:off + Cc].double()
    ref = torch.cat([x.mean(2).reshape(B * H, Cc), x.mean(1).reshape(B * W, Cc)]).float()
    assert (strip - ref).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item())


# ------------------------------------------------------------------------------------------------ Detect decode
@pytest.mark.parametrize("B,shapes,nc,pad", [(3, [(80, 80), (40, 40), (20, 20)], 80, 0), (2, [(8, 8), (4, 4), (2, 2)], 80, 16),
                                             (1, [(13, 7), (5, 9)], 4, 4), (2, [(20, 20)], 12, 0),
                                             (1, [(160, 160), (80, 80), (40, 40), (3, 3)], 80, 0)])
def test_detect_decode_vs_oracle(lib, B, shapes, nc, pad):
    """fce_detect_decode (head.py:149-167, DFL block.py:76-79, tal.py:352-376) against the CPU oracle on random
    logits: level boundaries inside a warp's 32-anchor slab, ragged level sizes, padded raw pitch, 1-4 levels.
    Boxes within 2e-3 pixel (approximate ex2 / rcp forms), class probabilities within 2e-6."""
    from oracle.fce_oracle import detect_decode
    l, L = lib
    g = torch.Generator().manual_seed(B + nc + len(shapes))
    R = 16
    no = 4 * R + nc
    raws = [torch.randn(B, h, w, no + pad, generator=g) * 3.0 for (h, w) in shapes]
    strides = [8.0, 16.0, 32.0, 64.0][:len(shapes)]
    A = sum(h * w for h, w in shapes)
    d = L.DecodeDesc(B=B, nl=len(shapes), nc=nc, reg_max=R)
    for i, (h, w) in enumerate(shapes):
        d.H[i], d.W[i], d.stride[i], d.raw_pitch[i] = h, w, strides[i], no + pad
    dev = [r.cuda() for r in raws]
    y = torch.full((B, 4 + nc, A), -9.0).cuda()
    ptrs = [C.c_void_p(t.data_ptr()) for t in dev] + [C.c_void_p(0)] * (4 - len(dev))
    st = l.fce_detect_decode(C.byref(d), *ptrs, C.c_void_p(y.data_ptr()), _stream())
    L.check(st, "fce_detect_decode")
    torch.cuda.synchronize()
    ref = detect_decode([r[..., :no].permute(0, 3, 1, 2).contiguous() for r in raws], strides, R)
    got = y.cpu()
    assert (got[:, :4] - ref[:, :4]).abs().max().item() <= 2e-3
    assert (got[:, 4:] - ref[:, 4:]).abs().max().item() <= 2e-6


# ------------------------------------------------------------------------------------------------ fused Detect epilogues
@pytest.mark.parametrize("B,H,W,Cin,nc", [(2, 80, 80, 128, 80), (3, 20, 20, 256, 80), (1, 13, 9, 64, 16), (5, 40, 40, 64, 80)])
def test_conv2d_detect_vs_torch(lib, B, H, W, Cin, nc):
    """fce_conv2d_detect: the last 1x1 conv of a Detect branch with the decode in its tcgen05 epilogue (head.py:94,103,
    149-167).  mode 1 writes sigmoid class scores, mode 2 the DFL boxes, both straight into y[B, 4+nc, A] at the
    level's anchor offset; the rest of y must stay untouched.  bf16 operands, fp32 accumulation: scores within 2e-3 of
    a torch fp32 conv on the same bf16 operands (accumulation order), boxes within 2e-2 pixel."""
    from oracle.fce_oracle import detect_decode
    l, L = lib
    g = torch.Generator().manual_seed(B + H + W + Cin + nc)
    A, a_base, stride = H * W + 37, 21, 16.0
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16).cuda()
    y = torch.full((B, 4 + nc, A), -3.0).cuda()
    refs = {}
    for mode, Cout in ((1, nc), (2, 64)):
        w = (torch.randn(Cout, 1, 1, Cin, generator=g) / math.sqrt(Cin) * 2).to(torch.bfloat16).cuda()
        b = torch.randn(Cout, generator=g).cuda()
        d = L.ConvDesc(B=B, H=H, W=W, Cin=Cin, Cout=Cout, in_pitch=Cin, in_off=0, out_pitch=0, out_off=0, res_pitch=0,
                       res_off=0, k=1, stride=1, act=L.ACT_NONE, in_dtype=L.BF16, w_dtype=L.BF16, out_dtype=L.F32,
                       in_layout=L.NHWC, in_scale=1.0, impl=0)
        e = L.DetectEpiDesc(mode=mode, A=A, a_base=a_base, rows=4 + nc, reg_max=16, stride=stride)
        st = l.fce_conv2d_detect(C.byref(d), C.byref(e), C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()),
                                 C.c_void_p(b.data_ptr()), C.c_void_p(y.data_ptr()), _stream())
        L.check(st, "fce_conv2d_detect")
        refs[mode] = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b)  # [B, Cout, H, W]
    torch.cuda.synchronize()
    sl = slice(a_base, a_base + H * W)
    cls_ref = torch.sigmoid(refs[1].reshape(B, nc, -1))
    assert (y[:, 4:, sl] - cls_ref).abs().max().item() <= 2e-3
    full = torch.cat([refs[2], torch.zeros(B, 1, H, W, device=y.device)], 1).cpu()
    box_ref = detect_decode([full], [stride], 16)[:, :4]
    assert (y[:, :4, sl].cpu() - box_ref).abs().max().item() <= 2e-2
    assert torch.all(y[:, :, :a_base] == -3.0) and torch.all(y[:, :, a_base + H * W:] == -3.0)


def test_conv2d_detect_rejects_what_it_cannot_fuse(lib):
    l, L = lib
    t = torch.zeros(1 << 16, dtype=torch.bfloat16).cuda()
    p = C.c_void_p(t.data_ptr())
    d = L.ConvDesc(B=1, H=8, W=8, Cin=64, Cout=64, in_pitch=64, in_off=0, out_pitch=0, out_off=0, res_pitch=0, res_off=0,
                   k=1, stride=1, act=L.ACT_NONE, in_dtype=L.BF16, w_dtype=L.BF16, out_dtype=L.F32, in_layout=L.NHWC,
                   in_scale=1.0, impl=0)
    e = L.DetectEpiDesc(mode=2, A=64, a_base=0, rows=84, reg_max=8, stride=8.0)
    assert l.fce_conv2d_detect(C.byref(d), C.byref(e), p, p, p, p, _stream()) == -2   # reg_max other than 16
    e.reg_max, e.a_base = 16, 1
    assert l.fce_conv2d_detect(C.byref(d), C.byref(e), p, p, p, p, _stream()) == -1   # level does not fit into A
    e.a_base, d.k = 0, 3
    assert l.fce_conv2d_detect(C.byref(d), C.byref(e), p, p, p, p, _stream()) == -2   # only the 1x1 tail convs


# ------------------------------------------------------------------------------------------------ BiFPN-fused conv epilogue
@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 80, 80, 256, 128), (3, 20, 20, 512, 256), (1, 6, 10, 64, 32)])
@pytest.mark.parametrize("res_up", [0, 1])
@pytest.mark.parametrize("scales", [(0.705, 0.295), (0.0, 0.43), (0.57, 1.0)])
def test_conv_weighted_sum_epilogue(lib, B, H, W, Cin, Cout, res_up, scales):
    """fce_conv2d with weighted / res_up: y = out_scale * SiLU(conv1x1(x) + b) + res_scale * res, res optionally a
    half-resolution map read through a nearest 2x upsample - BiFPN_Concat's weighted sum (fce_block.py:55-63) inside
    its realign conv.  A zero weight (relu(w) = 0) is a real case.  bf16 in / out: 2^-7 relative + accumulation noise."""
    l, L = lib
    g = torch.Generator().manual_seed(B + H + Cin + Cout + res_up)
    osc, rsc = scales
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(Cout, 1, 1, Cin, generator=g) / math.sqrt(Cin)).to(torch.bfloat16).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    rh, rw = (H // 2, W // 2) if res_up else (H, W)
    res = torch.randn(B, rh, rw, Cout, generator=g).to(torch.bfloat16).cuda()
    y = torch.zeros(B, H, W, Cout, dtype=torch.bfloat16).cuda()
    d = L.ConvDesc(B=B, H=H, W=W, Cin=Cin, Cout=Cout, in_pitch=Cin, in_off=0, out_pitch=Cout, out_off=0, res_pitch=Cout,
                   res_off=0, k=1, stride=1, act=L.ACT_SILU, in_dtype=L.BF16, w_dtype=L.BF16, out_dtype=L.BF16,
                   in_layout=L.NHWC, in_scale=1.0, impl=0, weighted=1, out_scale=osc, res_scale=rsc, res_up=res_up)
    st = l.fce_conv2d(C.byref(d), C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()),
                      C.c_void_p(res.data_ptr()), C.c_void_p(y.data_ptr()), _stream())
    L.check(st, "fce_conv2d weighted")
    torch.cuda.synchronize()
    conv = F.silu(F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b))
    r = res.float().permute(0, 3, 1, 2)
    if res_up:
        r = F.interpolate(r, scale_factor=2, mode="nearest")
    ref = (osc * conv + rsc * r).permute(0, 2, 3, 1)
    err = (y.float() - ref).abs()
    assert (err <= 2 ** -7 * ref.abs() + 2e-2).all(), err.max().item()


@pytest.mark.parametrize("B,H,W,Cin,Cout", [(2, 80, 80, 512, 256), (3, 20, 20, 512, 512), (1, 6, 10, 64, 32),
                                            (5, 40, 40, 384, 128)])
@pytest.mark.parametrize("res_up", [0, 1])
@pytest.mark.parametrize("impl", [0, 3, 4])
def test_conv_pre_activation_residual(lib, B, H, W, Cin, Cout, res_up, impl):
    """fce_conv2d with weighted = 2: y = SiLU(conv1x1(x) + res + b) - the second half of a 1x1 conv over a two-input
    BiFPN sum (conv(w0 a + w1 b) = (w0 W) a + (w1 W) b), res optionally at half resolution (the upsampled operand: a
    1x1 conv commutes with nearest upsampling).  Single-CTA and CTA-pair kernels."""
    l, L = lib
    g = torch.Generator().manual_seed(B + H + Cin + Cout + res_up)
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16).cuda()
    w = (torch.randn(Cout, 1, 1, Cin, generator=g) / math.sqrt(Cin)).to(torch.bfloat16).cuda()
    b = torch.randn(Cout, generator=g).cuda()
    rh, rw = (H // 2, W // 2) if res_up else (H, W)
    res = torch.randn(B, rh, rw, Cout, generator=g).to(torch.bfloat16).cuda()
    y = torch.zeros(B, H, W, Cout, dtype=torch.bfloat16).cuda()
    d = L.ConvDesc(B=B, H=H, W=W, Cin=Cin, Cout=Cout, in_pitch=Cin, in_off=0, out_pitch=Cout, out_off=0, res_pitch=Cout,
                   res_off=0, k=1, stride=1, act=L.ACT_SILU, in_dtype=L.BF16, w_dtype=L.BF16, out_dtype=L.BF16,
                   in_layout=L.NHWC, in_scale=1.0, impl=impl, weighted=2, out_scale=1.0, res_scale=1.0, res_up=res_up)
    st = l.fce_conv2d(C.byref(d), C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()),
                      C.c_void_p(res.data_ptr()), C.c_void_p(y.data_ptr()), _stream())
    if impl == 3 and st == -2:
        pytest.skip("no CTA-pair tiling for this shape")
    L.check(st, "fce_conv2d pre-activation residual")
    torch.cuda.synchronize()
    r = res.float().permute(0, 3, 1, 2)
    if res_up:
        r = F.interpolate(r, scale_factor=2, mode="nearest")
    ref = F.silu(F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), b) + r).permute(0, 2, 3, 1)
    err = (y.float() - ref).abs()
    assert (err <= 2 ** -7 * ref.abs() + 2e-2).all(), err.max().item()
    d.weighted = 2
    assert l.fce_conv2d(C.byref(d), C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(b.data_ptr()), None,
                        C.c_void_p(y.data_ptr()), _stream()) == -1  # weighted = 2 without a residual


def test_conv_weighted_sum_needs_the_tensor_core_1x1(lib):
    l, L = lib
    t = torch.zeros(1 << 16, dtype=torch.bfloat16).cuda()
    p = C.c_void_p(t.data_ptr())
    d = L.ConvDesc(B=1, H=8, W=8, Cin=32, Cout=32, in_pitch=32, in_off=0, out_pitch=32, out_off=0, res_pitch=32, res_off=0,
                   k=3, stride=1, act=L.ACT_SILU, in_dtype=L.BF16, w_dtype=L.BF16, out_dtype=L.BF16, in_layout=L.NHWC,
                   in_scale=1.0, impl=0, weighted=1, out_scale=0.5, res_scale=0.5, res_up=0)
    assert l.fce_conv2d(C.byref(d), p, p, p, p, p, _stream()) == -2      # 3x3: no weighted epilogue
    d.k, d.res_up, d.H = 1, 1, 7
    assert l.fce_conv2d(C.byref(d), p, p, p, p, p, _stream()) == -1      # odd map cannot be a 2x upsample
