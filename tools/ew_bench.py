#!/usr/bin/env python
"""Micro-benchmark of the bandwidth-class kernels on the bench workload's shapes (batch 64, bf16).
Prints achieved algorithmic GB/s against the measured HBM copy peak (MEASURED_PEAKS.json)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from fce_yolo_b200 import _lib as L

PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, flush, reps=10):
    for _ in range(2):
        fn()
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


def main():
    lib = L.load(check_device=True)
    dev = torch.device("cuda:0")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    P = lambda t: C.c_void_p(t.data_ptr() if t is not None else 0)

    def report(name, ms, byts):
        print(f"{name:40s} {ms*1e3:8.1f} us  {byts/ms/1e6:7.0f} GB/s  ({byts/ms/1e6/PEAK*100:5.1f}% of measured HBM peak)", flush=True)

    for (H, Cc) in [(80, 128), (40, 256), (20, 512)]:
        x = torch.randn(B, H, H, Cc, device=dev).to(torch.bfloat16)
        y = torch.empty_like(x)
        w = torch.randn(9, Cc, device=dev)
        b = torch.randn(Cc, device=dev)
        d = L.DwconvDesc(B=B, H=H, W=H, C=Cc, in_pitch=Cc, in_off=0, out_pitch=Cc, out_off=0, add_pitch=0, add_off=0,
                         act=1, dtype=L.BF16)
        ms = timeit(lambda: lib.fce_dwconv3x3(C.byref(d), P(x), P(w), P(b), P(None), P(y), st), flush)
        report(f"dwconv3x3 {H}x{H}x{Cc}", ms, 2 * x.numel() * 2)
    for (H, Cc) in [(80, 256), (40, 512)]:
        x = torch.randn(B, H, H, Cc, device=dev).to(torch.bfloat16)
        strip = torch.empty(B * 2 * H, Cc, device=dev)
        d = L.PoolDesc(B=B, H=H, W=H, C=Cc, pitch=Cc, off=0, dtype=L.BF16)
        nws = lib.fce_coord_pool_workspace(C.byref(d))
        ws = torch.empty(max(nws, 16), dtype=torch.uint8, device=dev)
        ms = timeit(lambda: lib.fce_coord_pool(C.byref(d), P(x), P(strip), P(ws), C.c_size_t(nws), st), flush)
        report(f"coord_pool {H}x{H}x{Cc}", ms, x.numel() * 2)
        y = torch.empty_like(x)
        gh = torch.rand(B * H, Cc, device=dev)
        gw = torch.rand(B * H, Cc, device=dev)
        g = L.GateDesc(B=B, H=H, W=H, C=Cc, mode=0, in_pitch=Cc, in_off=0, out_pitch=Cc, out_off=0, dtype=L.BF16,
                       gh_bstride=H * Cc, gh_rstride=Cc, gw_bstride=H * Cc, gw_rstride=Cc)
        ms = timeit(lambda: lib.fce_gate_apply(C.byref(g), P(x), P(gh), P(gw), P(y), st), flush)
        report(f"gate_apply {H}x{H}x{Cc}", ms, 2 * x.numel() * 2)
    for (H, Cc, mip) in [(80, 256, 16), (40, 512, 16)]:
        rows = B * H
        strip = torch.randn(2 * rows, Cc, device=dev)
        out = torch.empty(2 * rows, Cc, device=dev)
        w1t, b1 = torch.randn(Cc // 4, mip, 4, device=dev), torch.randn(mip, device=dev)
        wht, bh = torch.randn(mip, Cc, device=dev), torch.randn(Cc, device=dev)
        d = L.CoordAttMlpDesc(rows_h=rows, rows_w=rows, C=Cc, mip=mip, oup=Cc, s_pitch=Cc, out_pitch=Cc, act1=1, act2=2)
        ms = timeit(lambda: lib.fce_coordatt_mlp(C.byref(d), P(strip), P(w1t), P(b1), P(wht), P(bh), P(wht), P(bh), P(out),
                                                 st), flush)
        report(f"coordatt_mlp {2 * rows}x{Cc}->{mip}->{Cc}", ms, 2 * strip.numel() * 4)
    for (H, Cc) in [(20, 256)]:
        buf = torch.randn(B, H, H, 4 * Cc, device=dev).to(torch.bfloat16)
        d = L.SppfDesc(B=B, H=H, W=H, C=Cc, pitch=4 * Cc, off=0, dtype=L.BF16)
        ms = timeit(lambda: lib.fce_sppf_pool(C.byref(d), P(buf), st), flush)
        report(f"sppf_pool {H}x{H}x{Cc}", ms, 4 * B * H * H * Cc * 2)
    # reference point: a plain device copy of 210 MB
    a = torch.empty(105 << 20, dtype=torch.uint8, device=dev)
    c = torch.empty_like(a)
    ms = timeit(lambda: c.copy_(a), flush)
    report("torch copy 105 MB (reference)", ms, 2 * a.numel())


if __name__ == "__main__":
    main()
