"""Times fce_stem2_conv (stem + second conv in one pass) against fce_stem_conv + fce_conv2d, L2 flushed between launches.
Usage: python tools/stem2_bench.py [--batch 256] [--size 640] [--c1 128]; debug builds print the per-role cycle accounting."""
import argparse
import ctypes as C
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fce_yolo_b200 import _lib as L  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--size", type=int, default=640)
    ap.add_argument("--c1", type=int, default=128)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--no-two", action="store_true")
    a = ap.parse_args()
    lib = L.load(check_device=True)
    dev = torch.device("cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)  # noqa: E731
    B, S, C0, C1 = a.batch, a.size, 64, a.c1
    x = torch.randint(0, 256, (B, S, S, 3), dtype=torch.uint8, device=dev)
    wk = torch.zeros(C0, 32, device=dev)
    wk[:, :27] = torch.randn(C0, 27, device=dev) * 0.3 / 255
    wk = wk.bfloat16()
    b0 = torch.randn(C0, device=dev) * 0.2
    w1 = (torch.randn(C1, 3, 3, C0, device=dev) / (9 * C0) ** 0.5).bfloat16()
    b1 = torch.randn(C1, device=dev) * 0.1
    y = torch.empty(B, S // 4, S // 4, C1, dtype=torch.bfloat16, device=dev)
    mid = torch.empty(B, S // 2, S // 2, C0, dtype=torch.bfloat16, device=dev)
    d = L.Stem2Desc(B=B, H=S, W=S, C0=C0, C1=C1, out_pitch=C1, out_off=0, act0=1, act1=1)
    ds = L.StemDesc(B=B, H=S, W=S, Cout=C0, out_pitch=C0, out_off=0, act=1, in_dtype=L.U8, in_layout=L.NHWC)
    dc = L.ConvDesc(B=B, H=S // 2, W=S // 2, Cin=C0, Cout=C1, in_pitch=C0, in_off=0, out_pitch=C1, out_off=0, res_pitch=0,
                    res_off=0, k=3, stride=2, act=1, in_dtype=L.BF16, w_dtype=L.BF16, out_dtype=L.BF16, in_layout=L.NHWC,
                    in_scale=1.0, impl=0)

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

    t_f = timed(lambda: L.check(lib.fce_stem2_conv(C.byref(d), p(x), p(wk), p(b0), p(w1), p(b1), p(y), st), "stem2"))
    if hasattr(lib, "fce_stem2_profile"):
        buf = (C.c_longlong * (148 * 16))()
        lib.fce_stem2_profile(buf, 148 * 16)
        col = lambda k: statistics.mean(buf[i * 16 + k] for i in range(148))  # noqa: E731
        print(f"cycles/CTA: builder wait-A-empty {col(0):.0f} total {col(1):.0f} | stem epilogue wait-acc {col(2):.0f} wait-plane-free "
              f"{col(3):.0f} total {col(4):.0f} | MMA wait-A-full {col(5):.0f} wait-stem-acc-empty {col(6):.0f} wait-plane-full "
              f"{col(7):.0f} wait-acc-empty {col(8):.0f} total {col(9):.0f} | epilogue wait {col(10):.0f} total {col(11):.0f}")
    t_a = t_b = float("nan")
    if not a.no_two:
        t_a = timed(lambda: L.check(lib.fce_stem_conv(C.byref(ds), p(x), p(wk), p(b0), p(mid), st), "stem"))
        t_b = timed(lambda: L.check(lib.fce_conv2d(C.byref(dc), p(mid), p(w1), p(b1), p(None), p(y), st), "conv"))
    print(f"b{B} {S}x{S} 3->64->{C1}: fused {t_f * 1e3:8.1f} us   two launches {t_a * 1e3:7.1f} + {t_b * 1e3:7.1f} = "
          f"{(t_a + t_b) * 1e3:7.1f} us")


if __name__ == "__main__":
    main()
