"""Times fce_conv1x1_chain (C3k.cv3 chained into C3k2.cv2 in one pass) against the two-launch route fce_conv2d + fce_conv2d on
the C3k2 tails of the BASELINE configs, L2 flushed between launches, and checks that both give the same bits.
Usage: python tools/chain_bench.py [--batch 256] [--reps 5] [--only m.L2]"""
import argparse
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fce_yolo_b200 import _lib as L  # noqa: E402

# name, map size, c (hidden width), inner blocks, Cout
SHAPES = [("m.L2 64 -> 256 @160", 160, 64, 1, 256), ("m.L4 128 -> 512 @80", 80, 128, 1, 512),
          ("m.L16 128 -> 256 @80", 80, 128, 1, 256), ("s.L6 64 -> 256 @40 (n scale widths x2)", 40, 64, 1, 256),
          ("n.L8 128 -> 256 @20", 20, 128, 1, 256)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default=None)
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

    for name, hw, c, n, Cout in SHAPES:
        if a.only and a.only not in name:
            continue
        B, c2 = a.batch, (1 + n) * c
        p1 = c * 3 // 2  # C3k's [chain out | cv2(x) | cv1(x)] buffer
        x1 = torch.randn(B, hw, hw, p1, device=dev).to(torch.bfloat16)
        cat = torch.randn(B, hw, hw, c2 + c, device=dev).to(torch.bfloat16)   # two-launch layout; the chain reads [:c2]
        w1 = (torch.randn(c, c, device=dev) / c ** 0.5).to(torch.bfloat16)
        b1 = torch.randn(c, device=dev) * 0.1
        w2 = (torch.randn(Cout, c2 + c, device=dev) / (c2 + c) ** 0.5).to(torch.bfloat16)
        b2 = torch.randn(Cout, device=dev) * 0.1
        y1 = torch.empty(B, hw, hw, Cout, device=dev, dtype=torch.bfloat16)
        y2 = torch.empty_like(y1)
        d = L.ChainDesc(B=B, H=hw, W=hw, c1=c, cm=c, c2=c2, Cout=Cout, x1_pitch=p1, x1_off=0, x2_pitch=c2 + c, x2_off=0,
                        out_pitch=Cout, out_off=0, act1=1, act2=1)

        def conv(x, cin, pitch, w, b, y, yoff, cout, ypitch):
            dc = L.ConvDesc(B=B, H=hw, W=hw, Cin=cin, Cout=cout, in_pitch=pitch, in_off=0, out_pitch=ypitch, out_off=yoff,
                            res_pitch=0, res_off=0, k=1, stride=1, act=1, in_dtype=L.BF16, w_dtype=L.BF16, out_dtype=L.BF16,
                            in_layout=L.NHWC, in_scale=1.0, impl=0)
            L.check(lib.fce_conv2d(C.byref(dc), p(x), p(w), p(b), p(None), p(y), st), "fce_conv2d")

        t_a = timed(lambda: conv(x1, c, p1, w1, b1, cat, c2, c, c2 + c))
        t_b = timed(lambda: conv(cat, c2 + c, c2 + c, w2, b2, y2, 0, Cout, Cout))
        if lib.fce_conv1x1_chain_route(C.byref(d)) != 1:
            print(f"{name:42s} cv3 {t_a * 1e3:7.1f} + cv2 {t_b * 1e3:7.1f} us; chain: shape not taken")
            continue
        t_f = timed(lambda: L.check(lib.fce_conv1x1_chain(C.byref(d), p(x1), p(w1), p(b1), p(cat), p(w2), p(b2), p(y1), st),
                                    "fce_conv1x1_chain"))
        same = torch.equal(y1, y2)
        px = B * hw * hw
        gb = px * (c + c2 + Cout) * 2 / 1e9
        print(f"{name:42s} cv3 {t_a * 1e3:7.1f} + cv2 {t_b * 1e3:7.1f} = {(t_a + t_b) * 1e3:7.1f} us -> chain {t_f * 1e3:7.1f} us "
              f"({gb / t_f * 1e3:6.0f} GB/s algorithmic)  bit-identical: {same}")


if __name__ == "__main__":
    main()
