#!/usr/bin/env python
"""Micro-benchmark of fce_conv2d on the conv shapes of the bench workload (yolo11s-fce, batch 64, 640x640).

    python tools/conv_bench.py [--reps 10] [--once] [--only NAME,...] [--impl 2]

Times each shape with CUDA events (L2 flushed between repetitions by writing a 256 MB buffer) and prints
TFLOP/s plus the HBM-roofline time (input read once + output written once [+ residual]).  --once launches
every shape a single time (for `ncu -k regex:conv_tc`)."""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fce_yolo_b200 import _lib as L

# name, B, H, W, Cin, Cout, k, stride, residual, out_f32
SHAPES = [
    ("L1   3x3s2  32->64  @320", 64, 320, 320, 32, 64, 3, 2, 0, 0),
    ("L2.cv1 1x1  64->64  @160", 64, 160, 160, 64, 64, 1, 1, 0, 0),
    ("L2.m.cv1 3x3 32->16 @160", 64, 160, 160, 32, 16, 3, 1, 0, 0),
    ("L2.m.cv2 3x3 16->32 @160", 64, 160, 160, 16, 32, 3, 1, 1, 0),
    ("L2.cv2 1x1  96->128 @160", 64, 160, 160, 96, 128, 1, 1, 0, 0),
    ("L3   3x3s2 128->128 @160", 64, 160, 160, 128, 128, 3, 2, 0, 0),
    ("L4.cv1 1x1 128->128 @80 ", 64, 80, 80, 128, 128, 1, 1, 0, 0),
    ("L4.m.cv1 3x3 64->32 @80 ", 64, 80, 80, 64, 32, 3, 1, 0, 0),
    ("L4.m.cv2 3x3 32->64 @80 ", 64, 80, 80, 32, 64, 3, 1, 1, 0),
    ("L4.cv2 1x1 192->256 @80 ", 64, 80, 80, 192, 256, 1, 1, 0, 0),
    ("L6   3x3s2 256->256 @80 ", 64, 80, 80, 256, 256, 3, 2, 0, 0),
    ("L7.c3k 3x3  64->64  @40 ", 64, 40, 40, 64, 64, 3, 1, 1, 0),
    ("L7.cv2 1x1 384->256 @40 ", 64, 40, 40, 384, 256, 1, 1, 0, 0),
    ("L9   3x3s2 256->512 @40 ", 64, 40, 40, 256, 512, 3, 2, 0, 0),
    ("L10.c3k 3x3 128->128 @20", 64, 20, 20, 128, 128, 3, 1, 1, 0),
    ("L10.cv2 1x1 768->512 @20", 64, 20, 20, 768, 512, 1, 1, 0, 0),
    ("L11.cv2 1x1 1024->512@20", 64, 20, 20, 1024, 512, 1, 1, 0, 0),
    ("Det.cv2.0.0 3x3 128->64@80", 64, 80, 80, 128, 64, 3, 1, 0, 0),
    ("Det.cv2.0.1 3x3 64->64 @80", 64, 80, 80, 64, 64, 3, 1, 0, 0),
    ("Det.cv2.1.0 3x3 256->64@40", 64, 40, 40, 256, 64, 3, 1, 0, 0),
    ("Det.cv2.2.0 3x3 512->64@20", 64, 20, 20, 512, 64, 3, 1, 0, 0),
    ("Det.cv3.0.2 1x1 128->80@80", 64, 80, 80, 128, 80, 1, 1, 0, 1),
    ("Det.cv2.0.2 1x1 64->64 @80 f32", 64, 80, 80, 64, 64, 1, 1, 0, 1),
    # m scale, batch 256 (the north-star configuration): the layers furthest from their own roofline
    ("M1   3x3s2  64->128 @320 b256", 256, 320, 320, 64, 128, 3, 2, 0, 0),
    ("M2.m.cv1 3x3 32->32 @160 b256", 256, 160, 160, 32, 32, 3, 1, 0, 0),
    ("M2.m.cv2 3x3 32->32 @160 b256 +res", 256, 160, 160, 32, 32, 3, 1, 1, 0),
    ("M2.cv3 1x1 64->64 @160 b256", 256, 160, 160, 64, 64, 1, 1, 0, 0),
    ("M4.m.cv1 3x3 64->64 @80 b256", 256, 80, 80, 64, 64, 3, 1, 0, 0),
    ("M2.cv2 1x1 192->256 @160 b256", 256, 160, 160, 192, 256, 1, 1, 0, 0),
    ("M4.cv2 1x1 384->512 @80 b256", 256, 80, 80, 384, 512, 1, 1, 0, 0),
    ("M16.cv2 1x1 384->256 @80 b256", 256, 80, 80, 384, 256, 1, 1, 0, 0),
    ("M6.cv1 1x1 512->512 @40 b256", 256, 40, 40, 512, 512, 1, 1, 0, 0),
    # L2-resident probes of the TMA feed rate (run with --noflush): A traffic only, tiny N
    ("probe tiled 1x1 1024->16 @80 b4", 4, 80, 80, 1024, 16, 1, 1, 0, 0),
    ("probe im2col 3x3 128->16 @80 b4", 4, 80, 80, 128, 16, 3, 1, 0, 0),
    ("probe im2col 3x3 64->16 @80 b8", 8, 80, 80, 64, 16, 3, 1, 0, 0),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--only", default=None)
    ap.add_argument("--impl", type=int, default=2)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--noflush", action="store_true")
    ap.add_argument("--act", type=int, default=1)
    ap.add_argument("--in-pitch-mult", type=int, default=1, help="read the input as a channel slice of a buffer this many times wider")
    ap.add_argument("--out-pitch-mult", type=int, default=1, help="write the output into a channel slice of a wider buffer")
    ap.add_argument("--prof", action="store_true", help="per-role cycle accounting of one launch per shape")
    ap.add_argument("--skip-tma", action="store_true", help="debug: producers skip the loads (MMA+epilogue rate)")
    ap.add_argument("--dbg", type=int, default=0, help="debug bits for timed runs: 1 skip TMA loads, 2 skip TMA stores, "
                                                      "4 skip epilogue math (results are garbage)")
    ap.add_argument("--nohalo", action="store_true", help="strip kernel off (sets FCE_HALO_MODE=0 before the library loads)")
    ap.add_argument("--nostream", action="store_true", help="strip kernel only with resident weights (FCE_HALO_MODE=2)")
    a = ap.parse_args()
    if a.nohalo or a.nostream:  # kernel-selection knob of conv_halo.cu, read once at load time
        os.environ["FCE_HALO_MODE"] = "0" if a.nohalo else "2"
    lib = L.load(check_device=True)
    debug = hasattr(lib, "fce_conv_tc_set_profile") and lib.fce_conv_tc_set_profile.argtypes is not None
    if (a.prof or a.dbg or a.skip_tma) and not debug:
        sys.exit("--prof / --dbg / --skip-tma need a debug build of the library: FCE_DEBUG=1 python fce_yolo_b200/build.py")
    dev = torch.device("cuda:0")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    tot_ms = tot_roof = 0.0
    for (name, B, H, W, Cin, Cout, k, s, has_res, f32) in SHAPES:
        if a.only and not any(t in name for t in a.only.split(",")):
            continue
        if a.batch and not name.startswith("probe"):
            B = a.batch
        pad = k // 2
        Ho, Wo = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
        x = torch.randn(B, H, W, Cin * a.in_pitch_mult, device=dev).to(torch.bfloat16)
        w = (torch.randn(Cout, k, k, Cin, device=dev) / (k * k * Cin) ** 0.5).to(torch.bfloat16)
        bias = torch.randn(Cout, device=dev)
        y = torch.empty(B, Ho, Wo, Cout * a.out_pitch_mult, device=dev, dtype=torch.float32 if f32 else torch.bfloat16)
        r = torch.randn(B, Ho, Wo, Cout, device=dev).to(torch.bfloat16) if has_res else None
        d = L.ConvDesc(B=B, H=H, W=W, Cin=Cin, Cout=Cout, in_pitch=Cin * a.in_pitch_mult, in_off=0,
                       out_pitch=Cout * a.out_pitch_mult, out_off=0, res_pitch=Cout if has_res else 0, res_off=0, k=k, stride=s, act=a.act, in_dtype=L.BF16,
                       w_dtype=L.BF16, out_dtype=L.F32 if f32 else L.BF16, in_layout=L.NHWC, in_scale=1.0, impl=a.impl)
        args = (C.byref(d), C.c_void_p(x.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(bias.data_ptr()),
                C.c_void_p(r.data_ptr() if has_res else 0), C.c_void_p(y.data_ptr()), C.c_void_p(st))
        if a.once:
            L.check(lib.fce_conv2d(*args), name)
            torch.cuda.synchronize()
            continue
        if a.prof:
            L.check(lib.fce_conv2d(*args), name)
            torch.cuda.synchronize()
            flush.zero_()
            lib.fce_conv_tc_set_profile(3 if a.skip_tma else 1)
            L.check(lib.fce_conv2d(*args), name)
            torch.cuda.synchronize()
            lib.fce_conv_tc_set_profile(0)
            buf = (C.c_longlong * (148 * 16))()
            lib.fce_conv_tc_profile(buf, 148 * 16)
            import numpy as np
            pr = np.array(buf[:], dtype=np.float64).reshape(148, 16)
            tiles = -(-(B * Ho * Wo) // 128)
            m = pr.mean(0)
            print(f"{name:34s} tiles/CTA {tiles/148:6.1f} | A-prod wait {m[0]:9.0f} / {m[1]:9.0f} | MMA wait-full {m[4]:9.0f} "
                  f"wait-tmem {m[5]:9.0f} / {m[6]:9.0f} | epi wait {m[7]:9.0f} / {m[8]:9.0f} stg-wait {m[9]:9.0f} | dma wait {m[10]:9.0f} / {m[11]:9.0f}", flush=True)
            continue
        if debug:
            lib.fce_conv_tc_set_profile(a.dbg << 1)
        for _ in range(2):
            L.check(lib.fce_conv2d(*args), name)
        ms = []
        for _ in range(a.reps):
            if not a.noflush:
                flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.fce_conv2d(*args)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        if debug:
            lib.fce_conv_tc_set_profile(0)
        ms.sort()
        t = ms[len(ms) // 2]
        flops = 2.0 * B * Ho * Wo * Cout * Cin * k * k
        byts = (x.numel() * 2 // a.in_pitch_mult + y.numel() * y.element_size() // a.out_pitch_mult +
                (r.numel() * 2 if has_res else 0) + w.numel() * 2)
        roof_ms = max(byts / 6553.9e9, flops / 1407e12) * 1e3
        tot_ms += t
        tot_roof += roof_ms
        print(f"{name:34s} {t*1e3:8.1f} us  {flops/t/1e9:7.1f} TF/s  {byts/t/1e6:7.0f} GB/s  roofline {roof_ms*1e3:7.1f} us "
              f"({roof_ms/t*100:5.1f}%)", flush=True)
    if not a.once and tot_ms:
        print(f"TOTAL {tot_ms*1e3:.1f} us, roofline {tot_roof*1e3:.1f} us ({tot_roof/tot_ms*100:.1f}%)")


if __name__ == "__main__":
    main()
