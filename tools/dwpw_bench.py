"""Times fce_dwpw_conv (depthwise 3x3 -> 1x1 in one pass) against the two-launch route fce_dwconv3x3 + fce_conv2d on the
Detect class-branch shapes, L2 flushed between launches.  Usage: python tools/dwpw_bench.py [--batch 256] [--reps 5]
(FCE_DWPW_PAIR=0/1 forces single CTAs / CTA pairs.)"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fce_yolo_b200 import _lib as L  # noqa: E402

SHAPES = [("m P3 256->256 @80", 80, 256, 256), ("m P4 512->256 @40", 40, 512, 256), ("m P4 256->256 @40", 40, 256, 256),
          ("m P5 512->256 @20", 20, 512, 256), ("s P3 128->128 @80", 80, 128, 128), ("s P4 256->128 @40", 40, 256, 128),
          ("s P5 512->128 @20", 20, 512, 128)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default=None)
    ap.add_argument("--no-two", action="store_true")
    a = ap.parse_args()
    lib = L.load(check_device=True)
    dev = torch.device("cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)  # noqa: E731

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        ms = []
        for _ in range(a.reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        return ms[len(ms) // 2]

    for name, hw, Cc, Cout in SHAPES:
        if a.only and a.only not in name:
            continue
        B = a.batch
        x = torch.randn(B, hw, hw, Cc, device=dev).to(torch.bfloat16)
        wd = (torch.randn(9, Cc, device=dev) * 0.3).contiguous()
        bd = torch.randn(Cc, device=dev) * 0.1
        wp = (torch.randn(Cout, Cc, device=dev) / Cc ** 0.5).to(torch.bfloat16)
        bp = torch.randn(Cout, device=dev) * 0.1
        y = torch.empty(B, hw, hw, Cout, device=dev, dtype=torch.bfloat16)
        mid = torch.empty(B, hw, hw, Cc, device=dev, dtype=torch.bfloat16)
        d = L.DwpwDesc(B=B, H=hw, W=hw, C=Cc, Cout=Cout, in_pitch=Cc, in_off=0, out_pitch=Cout, out_off=0, dw_act=1, pw_act=1)
        dd = L.DwconvDesc(B=B, H=hw, W=hw, C=Cc, in_pitch=Cc, in_off=0, out_pitch=Cc, out_off=0, add_pitch=0, add_off=0,
                          act=1, dtype=L.BF16)
        dc = L.ConvDesc(B=B, H=hw, W=hw, Cin=Cc, Cout=Cout, in_pitch=Cc, in_off=0, out_pitch=Cout, out_off=0, res_pitch=0,
                        res_off=0, k=1, stride=1, act=1, in_dtype=L.BF16, w_dtype=L.BF16, out_dtype=L.BF16,
                        in_layout=L.NHWC, in_scale=1.0, impl=0)
        t_f = None
        if lib.fce_dwpw_route(C.byref(d)) == 1:
            t_f = timed(lambda: L.check(lib.fce_dwpw_conv(C.byref(d), p(x), p(wd), p(bd), p(wp), p(bp), p(y), st), "dwpw"))
        if t_f is not None and hasattr(lib, "fce_dwpw_profile"):  # debug build: per-role cycle accounting of the last launch
            buf = (C.c_longlong * (148 * 16))()
            lib.fce_dwpw_profile(buf, 148 * 16)
            import statistics
            col = lambda k: statistics.mean(buf[i * 16 + k] for i in range(148))  # noqa: E731
            print(f"    cycles/CTA: DW warp wait-input {col(0):9.0f} wait-A-empty {col(1):9.0f} total {col(2):9.0f} | MMA wait-A-full "
                  f"{col(4):9.0f} wait-acc-empty {col(5):9.0f} total {col(6):9.0f} | epilogue wait-acc-full {col(8):9.0f} total {col(9):9.0f}")
        t_a = t_b = float("nan")
        if not a.no_two:
            t_a = timed(lambda: L.check(lib.fce_dwconv3x3(C.byref(dd), p(x), p(wd), p(bd), p(None), p(mid), st), "dw"))
            t_b = timed(lambda: L.check(lib.fce_conv2d(C.byref(dc), p(mid), p(wp), p(bp), p(None), p(y), st), "pw"))
        px = B * hw * hw
        hbm_us = px * (Cc + Cout) * 2 / 6553.9e9 * 1e6
        print(f"{name:22s} b{B}: fused {t_f * 1e3 if t_f else float('nan'):8.1f} us   two launches {t_a * 1e3:7.1f} + {t_b * 1e3:7.1f} "
              f"= {(t_a + t_b) * 1e3:7.1f} us   HBM floor {hbm_us:6.1f} us")


if __name__ == "__main__":
    main()
