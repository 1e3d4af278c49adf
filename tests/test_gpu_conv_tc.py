"""tcgen05 implicit-GEMM convolution (fce_conv2d, impl=2) through the C ABI against a torch fp32 reference
of the same op (F.conv2d on the bf16-rounded operands, fp32 math) - the floating-point kernel keeps a torch
reference next to the oracle.  Tolerance: the output is rounded to bf16 once (2^-9 relative) on top of fp32
accumulation-order noise, so max |err| <= 1e-2 * max|ref| and relL2 <= 4e-3 (bf16 out) / 1e-5 (fp32 out)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (B, H, W, Cin, Cout, k, stride, act, res, out_f32, in_pitch_extra, out_pitch_extra)
CASES = [
    (1, 16, 16, 64, 64, 1, 1, 1, False, False, 0, 0),
    (2, 20, 20, 64, 128, 1, 1, 1, False, False, 0, 0),       # M = 800: ragged last tile
    (2, 12, 12, 32, 64, 1, 1, 1, False, False, 0, 0),        # kc = 32 (64B swizzle)
    (2, 12, 12, 16, 32, 1, 1, 0, False, False, 0, 0),        # kc = 16 (32B swizzle)
    (2, 12, 12, 96, 80, 1, 1, 1, False, False, 32, 64),      # kc = 32 x3, Cout = 80, channel-slice views
    (1, 40, 40, 256, 512, 1, 1, 1, False, False, 0, 0),      # two N tiles
    (1, 40, 40, 128, 64, 1, 1, 0, False, True, 0, 80),       # fp32 output into a wider buffer (Detect)
    (2, 16, 16, 128, 128, 1, 1, 1, True, False, 0, 128),     # residual
    (1, 16, 16, 64, 64, 3, 1, 1, False, False, 0, 0),
    (2, 20, 20, 64, 64, 3, 1, 1, True, False, 64, 0),        # tiles cross rows and images, residual
    (2, 24, 24, 32, 16, 3, 1, 1, False, False, 0, 0),        # small channels
    (2, 24, 24, 16, 32, 3, 1, 1, True, False, 16, 32),
    (2, 32, 32, 64, 128, 3, 2, 1, False, False, 0, 0),       # stride 2
    (3, 40, 24, 128, 256, 3, 2, 1, False, False, 0, 0),      # stride 2, H != W
    (1, 80, 80, 128, 128, 3, 1, 1, False, False, 0, 0),
    (4, 20, 20, 512, 512, 3, 1, 1, False, False, 0, 0),      # long K, two N tiles
    (1, 8, 8, 192, 384, 1, 1, 2, False, False, 0, 0),        # sigmoid, bn = 192
    # 3x3 stride-1 strip ("halo") kernel shapes: weights resident, odd widths, bands that do not divide H
    (2, 40, 40, 64, 64, 3, 1, 1, True, False, 0, 64),
    (2, 80, 80, 64, 32, 3, 1, 1, False, False, 0, 0),
    (1, 160, 160, 32, 16, 3, 1, 1, False, False, 0, 0),
    (1, 160, 160, 16, 32, 3, 1, 1, True, False, 16, 0),
    (3, 20, 20, 128, 32, 3, 1, 1, False, False, 0, 0),       # two K chunks
    (2, 13, 27, 64, 48, 3, 1, 0, False, True, 0, 16),        # odd sizes, fp32 out
    # strip kernel with STREAMED weights (Detect box branch: 3x3, Cin = 128..512 -> 64; the weights do not fit)
    (2, 80, 80, 128, 64, 3, 1, 1, False, False, 0, 0),
    (3, 40, 40, 256, 64, 3, 1, 1, False, False, 0, 0),
    (5, 20, 20, 512, 64, 3, 1, 1, False, False, 0, 0),       # whole image per band, 8-16 K chunks
    (2, 33, 47, 384, 64, 3, 1, 1, True, False, 128, 64),     # odd sizes, residual, channel-slice views
    (1, 24, 40, 320, 32, 3, 1, 0, False, False, 0, 0),       # kc = 64 x5, 32 output channels
    (2, 40, 40, 160, 64, 3, 1, 1, False, False, 0, 0),       # Cin % 64 != 0: kc = 32 x5
    (3, 20, 20, 128, 128, 3, 1, 1, True, False, 0, 0),       # C3k bottleneck at P5: 128 outputs, residual
    (2, 40, 40, 192, 96, 3, 1, 1, False, False, 64, 32),
    # 1x1 layers with several N tiles on a full persistent grid
    (8, 80, 80, 192, 256, 1, 1, 1, False, False, 0, 0),      # 2 N tiles of 128
    (4, 80, 80, 256, 512, 1, 1, 1, True, False, 0, 0),       # 4 N tiles, residual
    (4, 80, 80, 256, 384, 1, 1, 1, False, False, 64, 0),     # 3 N tiles
    (5, 80, 80, 128, 320, 1, 1, 0, False, False, 0, 64),     # ragged last N tile (320 = 128 + 128 + 64)
    # strip kernel on maps wider than one TMA box: column tiles
    (1, 24, 320, 48, 48, 3, 1, 1, True, False, 0, 0),
    (1, 9, 515, 32, 64, 3, 1, 1, False, False, 0, 0),
    (2, 24, 160, 96, 96, 3, 1, 1, True, False, 96, 0),       # extra column tiles (96-output accumulator)
]


# CTA-pair (cta_group::2) kernel, forced with impl = 3: odd tile counts (the pair's last half is past M), every N-tile
# shape (Cout = 32 ... 512: bn = 32, 48, 80, 128, 192, 256, two N tiles), resident and streamed weights, 1x1 / 3x3 s1 / s2,
# residual, fp32 output, channel-slice views, kc = 16 / 32 / 64
PAIR_CASES = [
    (2, 16, 16, 64, 64, 1, 1, 1, False, False, 0, 0),        # 4 tiles = 2 pairs, weights resident
    (3, 16, 16, 64, 128, 1, 1, 1, False, False, 0, 0),       # 6 tiles
    (2, 20, 20, 64, 128, 1, 1, 1, False, False, 0, 0),       # M = 800: 7 tiles (odd), ragged last tile
    (1, 24, 16, 32, 48, 1, 1, 1, False, False, 0, 0),        # 3 tiles (odd), kc = 32, bn = 48
    (2, 12, 12, 16, 32, 1, 1, 0, False, False, 0, 0),        # kc = 16, bn = 32 (16 rows of B per CTA)
    (2, 12, 12, 96, 80, 1, 1, 1, False, False, 32, 64),      # S = 3, Cout = 80 (40 rows per CTA), channel-slice views
    (1, 40, 40, 256, 512, 1, 1, 1, False, False, 0, 0),      # two N tiles of 256, streamed weights, 13 tiles (odd)
    (1, 40, 40, 128, 64, 1, 1, 0, False, True, 0, 80),       # fp32 output into a wider buffer
    (2, 16, 16, 128, 128, 1, 1, 1, True, False, 0, 128),     # residual
    (2, 24, 24, 768, 512, 1, 1, 1, False, False, 0, 0),      # long-K 1x1, two N tiles
    (1, 16, 16, 192, 384, 1, 1, 2, False, False, 0, 0),      # sigmoid, bn = 192
    (2, 20, 20, 64, 64, 3, 1, 1, True, False, 64, 0),        # 3x3 s1 through TMA-im2col, tiles cross rows and images
    (1, 80, 80, 128, 128, 3, 1, 1, False, False, 0, 0),      # 50 tiles, streamed weights
    (2, 40, 40, 256, 256, 3, 1, 1, True, False, 0, 0),       # C3k bottleneck at m scale: K = 2304, bn = 256
    (2, 32, 32, 64, 128, 3, 2, 1, False, False, 0, 0),       # stride 2
    (3, 40, 24, 128, 256, 3, 2, 1, False, False, 0, 0),      # stride 2, H != W, 6 tiles
    (5, 20, 20, 256, 512, 3, 2, 1, False, False, 0, 0),      # stride 2, two N tiles, M = 500 (4 tiles, ragged)
    (4, 20, 20, 512, 512, 3, 1, 1, False, False, 0, 0),      # long K, two N tiles, 13 tiles (odd)
    (8, 80, 80, 256, 512, 1, 1, 1, True, False, 0, 0),       # full grid of pairs: 2 x 256 columns, residual
    (8, 80, 80, 192, 384, 1, 1, 1, False, False, 0, 0),      # full grid of pairs: 3 x 128 columns
    (2, 24, 24, 32, 16, 3, 1, 1, False, False, 0, 0),        # bn = 16: no legal pair tiling -> FCE_ERR_UNSUPPORTED
]


def run_case(lib, L, case, impl):
    B, H, W, Cin, Cout, k, s, act, has_res, out_f32, ipx, opx = case
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(hash(case) & 0xFFFF)
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
    in_pitch, in_off = Cin + ipx, ipx // 2
    out_pitch, out_off = Cout + opx, opx // 2
    xb = torch.randn(B, H, W, in_pitch, generator=g).to(dev, torch.bfloat16)
    w = (torch.randn(Cout, k, k, Cin, generator=g) / (k * k * Cin) ** 0.5).to(dev, torch.bfloat16)  # OHWI
    bias = torch.randn(Cout, generator=g).to(dev)
    odt = torch.float32 if out_f32 else torch.bfloat16
    yb = torch.full((B, Ho, Wo, out_pitch), 7.0, dtype=odt, device=dev)
    rb = torch.randn(B, Ho, Wo, out_pitch, generator=g).to(dev, torch.bfloat16) if has_res else None
    d = L.ConvDesc(B=B, H=H, W=W, Cin=Cin, Cout=Cout, in_pitch=in_pitch, in_off=in_off, out_pitch=out_pitch,
                   out_off=out_off, res_pitch=out_pitch if has_res else 0, res_off=out_off if has_res else 0, k=k,
                   stride=s, act=act, in_dtype=L.BF16, w_dtype=L.BF16, out_dtype=L.F32 if out_f32 else L.BF16,
                   in_layout=L.NHWC, in_scale=1.0, impl=impl)
    st = lib.fce_conv2d(C.byref(d), C.c_void_p(xb.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(bias.data_ptr()),
                        C.c_void_p(rb.data_ptr() if has_res else 0), C.c_void_p(yb.data_ptr()),
                        C.c_void_p(torch.cuda.current_stream().cuda_stream))
    L.check(st, f"fce_conv2d {case}")
    torch.cuda.synchronize()
    x = xb[..., in_off:in_off + Cin].float().permute(0, 3, 1, 2)
    ref = F.conv2d(x, w.float().permute(0, 3, 1, 2), bias, stride=s, padding=pad)
    if act == 1:
        ref = F.silu(ref)
    elif act == 2:
        ref = torch.sigmoid(ref)
    if has_res:
        ref = ref + rb[..., out_off:out_off + Cout].float().permute(0, 3, 1, 2)
    out = yb[..., out_off:out_off + Cout].float().permute(0, 3, 1, 2)
    err_max = ((out - ref).abs().max() / ref.abs().max()).item()
    err_l2 = ((out - ref).norm() / ref.norm()).item()
    # channels outside the destination view must be untouched
    untouched = True
    if opx:
        untouched = bool((yb[..., :out_off] == 7.0).all() and (yb[..., out_off + Cout:] == 7.0).all())
    return err_max, err_l2, untouched


@pytest.fixture(scope="module")
def lib():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from fce_yolo_b200 import _lib as L
    return L.load(check_device=True), L


@pytest.mark.parametrize("case", CASES, ids=[f"c{i}" for i in range(len(CASES))])
def test_conv_tc_vs_torch(lib, case):
    l, L = lib
    err_max, err_l2, untouched = run_case(l, L, case, impl=2)
    print(case, f"max {err_max:.2e} l2 {err_l2:.2e}")
    assert untouched
    assert err_max < 1e-2, case
    assert err_l2 < (2e-5 if case[9] else 4e-3), case


@pytest.mark.parametrize("case", PAIR_CASES, ids=[f"p{i}" for i in range(len(PAIR_CASES))])
def test_conv_tc_pair_vs_torch(lib, case):
    """The cta_group::2 kernel against torch, and against the single-CTA kernel (same MMA arithmetic: the accumulation
    order over K is identical, so the two must agree to the last bit)."""
    l, L = lib
    if case[4] < 32:
        with pytest.raises(ValueError):
            run_case(l, L, case, impl=3)
        return
    buf = (C.c_longlong * 4)()
    l.fce_conv_stats(buf, 1)
    err_max, err_l2, untouched = run_case(l, L, case, impl=3)
    l.fce_conv_stats(buf, 0)
    assert buf[1] == 1 and buf[0] == 0 and buf[2] == 0 and buf[3] == 0, list(buf)  # it really was the pair kernel
    print(case, f"max {err_max:.2e} l2 {err_l2:.2e}")
    assert untouched
    assert err_max < 1e-2, case
    assert err_l2 < (2e-5 if case[9] else 4e-3), case
    e1 = run_case(l, L, case, impl=4)
    assert (err_max, err_l2) == (e1[0], e1[1]), (case, err_max, e1)


# 3x3 stride-1 strip kernel as CTA pairs (impl = 6): resident and streamed weights, residual (staged by TMA), odd unit
# counts (the last pair's second CTA is a dummy), bands that do not divide H, two K chunks, Cout = 32 ... 128
HALO_PAIR_CASES = [
    (2, 40, 40, 64, 64, 3, 1, 1, True, False, 0, 64),        # resident weights, residual
    (2, 80, 80, 64, 32, 3, 1, 1, False, False, 0, 0),
    (3, 20, 20, 128, 32, 3, 1, 1, False, False, 0, 0),       # 3 units (odd): dummy half
    (1, 160, 160, 32, 32, 3, 1, 1, True, False, 32, 0),      # m-scale L2 bottleneck: thin, wide map
    (2, 80, 80, 128, 64, 3, 1, 1, False, False, 0, 0),       # streamed weights (Detect box branch)
    (3, 40, 40, 256, 64, 3, 1, 1, False, False, 0, 0),
    (5, 20, 20, 512, 64, 3, 1, 1, False, False, 0, 0),       # 5 units (odd), 8 K chunks
    (2, 33, 47, 384, 64, 3, 1, 1, True, False, 128, 64),     # odd sizes, residual, channel-slice views
    (3, 20, 20, 128, 128, 3, 1, 1, True, False, 0, 0),       # C3k bottleneck: 128 outputs streamed, residual
    (2, 40, 40, 128, 128, 3, 1, 1, False, False, 0, 0),      # the m-scale 128 -> 128 layers
    (2, 40, 40, 192, 96, 3, 1, 1, False, False, 64, 32),
    # maps wider than one TMA box (254 output columns): column tiles, each with its own halo columns
    (1, 24, 320, 48, 48, 3, 1, 1, True, False, 0, 0),        # x-scale L2 bottleneck at 1280^2: two tiles of 160
    (2, 11, 300, 64, 64, 3, 1, 1, False, False, 0, 32),      # 2 x 150
    (1, 9, 515, 32, 64, 3, 1, 1, True, False, 32, 0),        # 3 x 172, ragged last tile (171 columns)
    (1, 7, 255, 64, 32, 3, 1, 1, False, False, 0, 0),        # just past one box: 128 + 127
    # more column tiles than the box needs, where one row would waste the 96-output accumulator (x scale)
    (2, 24, 160, 96, 96, 3, 1, 1, False, False, 0, 0),
    (2, 24, 160, 96, 96, 3, 1, 1, True, False, 96, 0),
    (1, 17, 100, 96, 96, 3, 1, 1, True, False, 0, 32),       # ragged tiles
]


@pytest.mark.parametrize("case", HALO_PAIR_CASES, ids=[f"h{i}" for i in range(len(HALO_PAIR_CASES))])
def test_conv_halo_pair_vs_torch(lib, case):
    l, L = lib
    buf = (C.c_longlong * 4)()
    l.fce_conv_stats(buf, 1)
    err_max, err_l2, untouched = run_case(l, L, case, impl=6)
    l.fce_conv_stats(buf, 0)
    assert buf[1] == 1 and buf[0] == 0 and buf[2] == 0 and buf[3] == 0, list(buf)
    print(case, f"max {err_max:.2e} l2 {err_l2:.2e}")
    assert untouched
    assert err_max < 1e-2, case
    assert err_l2 < 4e-3, case
    try:
        e1 = run_case(l, L, case, impl=5)
    except ValueError:
        return  # the single-CTA strip kernel does not take this shape (e.g. 128 outputs on a large problem)
    assert (err_max, err_l2) == (e1[0], e1[1]), (case, err_max, e1)


def test_conv_stats_show_the_route(lib):
    """fce_conv2d_route / fce_conv_stats make the kernel choice visible: a bf16 conv with 8 output channels leaves the
    tensor cores (SIMT), a 64-channel one does not."""
    l, L = lib
    buf = (C.c_longlong * 4)()
    l.fce_conv_stats(buf, 1)
    run_case(l, L, (1, 16, 16, 64, 64, 1, 1, 1, False, False, 0, 0), impl=0)
    l.fce_conv_stats(buf, 1)
    assert buf[0] + buf[1] + buf[2] == 1 and buf[3] == 0
    run_case(l, L, (1, 16, 16, 64, 8, 1, 1, 1, False, False, 0, 0), impl=0)
    l.fce_conv_stats(buf, 1)
    assert buf[3] == 1 and buf[0] + buf[1] + buf[2] == 0


if __name__ == "__main__":  # quick report: python tests/test_gpu_conv_tc.py
    import sys
    sys.path.insert(0, ".")
    from fce_yolo_b200 import _lib as L
    torch.backends.cudnn.allow_tf32 = False
    l = L.load(check_device=True)
    tc_impl = int(sys.argv[1]) if len(sys.argv) > 1 else 2  # 3 = CTA-pair kernel, 4 = single-CTA kernel (no strip kernel)
    for i, case in enumerate(CASES + PAIR_CASES):
        try:
            em, el, ut = run_case(l, L, case, impl=tc_impl)
            es, esl, _ = run_case(l, L, case, impl=1)
            print(f"c{i} {case}: tc max {em:.3e} l2 {el:.3e} untouched {ut} | simt max {es:.3e} l2 {esl:.3e}", flush=True)
        except Exception as e:  # noqa
            print(f"c{i} {case}: EXC {e}", flush=True)
            break
