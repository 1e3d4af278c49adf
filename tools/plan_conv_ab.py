#!/usr/bin/env python
"""A/B timing of every DISTINCT dense-conv node of a bench workload's plan under the kernel choices of fce_conv2d:
automatic (impl 0/2: strip kernel / single CTA / CTA pair as the library decides), single-CTA implicit GEMM (impl 4) and
CTA-pair implicit GEMM (impl 3).  The nodes run on the plan's own buffers with their own descriptors, one at a time,
L2 flushed between repetitions, CUDA events.  Drives the pair / strip heuristics in csrc/conv_tc.cu.

    python tools/plan_conv_ab.py --config 2 [--batch 64] [--reps 5] [--csv out.csv]
"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import bench  # noqa: E402
from fce_yolo_b200 import _lib as L  # noqa: E402
from fce_yolo_b200.predict import Predictor  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=None)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--csv", default=None)
    ap.add_argument("--dense", action="store_true", help="also time every conv on dense scratch tensors")
    ap.add_argument("--only", default=None, help="comma-separated substrings of node tags to time")
    a = ap.parse_args()
    w, name = bench.workload(a.config)
    if a.batch:
        w["batch"] = a.batch
    lib = L.load(check_device=True)
    dev = torch.device("cuda:0")
    cfg, model, sd = bench.build_model(w)
    pred = Predictor(model, w["batch"], w["size"], precision="bf16", device=dev, use_graph=False)
    pred.run_device()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    pk = bench.peaks()
    seen = {}
    rows = []
    tot = {"auto": 0.0, "single": 0.0, "pair": 0.0, "strip": 0.0, "strip2": 0.0, "best": 0.0}
    for fn, args, n in pred.ex._calls:
        if n.fn not in ("fce_conv2d", "fce_conv2d_detect"):
            continue
        d = n.desc
        if d.in_dtype != L.BF16 or d.w_dtype != L.BF16:
            continue
        if a.only and not any(t in n.tag for t in a.only.split(",")):
            continue
        key = (n.fn, d.B, d.H, d.W, d.Cin, d.Cout, d.k, d.stride, d.res_pitch != 0, d.out_dtype, d.weighted, d.res_up) + (
            (d.in_pitch, d.out_pitch, d.res_pitch) if a.dense else ())
        if key in seen:
            seen[key][0] += 1
            continue
        res = {}
        for _ in range(3):  # untimed: the first variant timed after the previous node's runs came out 5-10 % slow
            flush.zero_()
            fn(*args, st)
        torch.cuda.synchronize()
        for label, impl in (("auto", 0), ("single", 4), ("pair", 3), ("strip", 5), ("strip2", 6)):
            saved = d.impl
            d.impl = impl
            rc = fn(*args, st)
            torch.cuda.synchronize()
            if rc != 0:
                res[label] = None
                d.impl = saved
                continue
            ms = []
            for _ in range(a.reps):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn(*args, st)
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            ms.sort()
            res[label] = ms[len(ms) // 2]
            d.impl = saved
        if a.dense and n.fn == "fce_conv2d":
            # the same conv on DENSE scratch tensors (pitch = channel count): what the channel-slice views cost
            import copy
            dd = copy.copy(d)
            Ho = (d.H + 2 * (d.k // 2) - d.k) // d.stride + 1
            Wo = (d.W + 2 * (d.k // 2) - d.k) // d.stride + 1
            xs = torch.randn(d.B, d.H, d.W, d.Cin, device=dev).to(torch.bfloat16)  # real data: zeros run cooler and faster
            ys = torch.empty(d.B, Ho, Wo, d.Cout, dtype=torch.float32 if d.out_dtype == L.F32 else torch.bfloat16, device=dev)
            rs = None
            if d.res_pitch:
                q = 2 if d.res_up else 1
                rs = torch.randn(d.B, Ho // q, Wo // q, d.Cout, device=dev).to(torch.bfloat16)
            dd.in_pitch, dd.in_off, dd.out_pitch, dd.out_off = d.Cin, 0, d.Cout, 0
            dd.res_pitch, dd.res_off = (d.Cout if rs is not None else 0), 0
            dargs = [C.byref(dd), C.c_void_p(xs.data_ptr()), args[2], args[3],
                     C.c_void_p(rs.data_ptr() if rs is not None else 0), C.c_void_p(ys.data_ptr())]
            if fn(*dargs, st) == 0:
                torch.cuda.synchronize()
                ms = []
                for _ in range(a.reps):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    fn(*dargs, st)
                    e1.record()
                    torch.cuda.synchronize()
                    ms.append(e0.elapsed_time(e1))
                ms.sort()
                res["dense"] = ms[len(ms) // 2]
            del xs, ys, rs
        seen[key] = [1, n.tag, res, n.flops, n.bytes, (d.in_pitch, d.Cin, d.out_pitch, d.Cout)]
    print(f"{'node':30s} {'shape':34s} cnt {'auto us':>9s} {'single':>9s} {'pair':>9s} {'strip':>9s} {'strip2':>9s}  auto TF  best TF  roofline us")
    if a.dense:
        gain = 0.0
        print("dense-tensor twin of every conv (what the channel-slice views cost):")
        for key, (cnt, tag, res, fl, by, pit) in seen.items():
            if res.get("dense") is None:
                continue
            g = cnt * (res["auto"] - res["dense"])
            gain += g
            if abs(g) > 0.004:
                print(f"  {tag:30s} x{cnt} in {pit[1]}/{pit[0]} out {pit[3]}/{pit[2]}  views {res['auto']*1e3:8.1f} us  dense "
                      f"{res['dense']*1e3:8.1f} us  gain {g*1e3:7.1f} us")
        print(f"  TOTAL gain if every view were dense: {gain:.3f} ms per step")
    for key, (cnt, tag, res, fl, by, *_r) in seen.items():
        _, B, H, W, Cin, Cout, k, s, has_res, odt, wt, ru = key[:12]
        shape = f"{k}x{k}s{s} {Cin}->{Cout} @{H}x{W}" + (" +res" if has_res else "") + (" f32" if odt == L.F32 else "")
        roof = max(fl / (pk["tf_sustained"] * 1e12), by / (pk["hbm"] * 1e9)) * 1e6
        vals = {k_: v for k_, v in res.items() if v is not None and k_ != "dense"}
        best = min(vals.values())
        f = lambda v: f"{v * 1e3:9.1f}" if v is not None else "        -"  # noqa: E731
        print(f"{tag:30s} {shape:34s} {cnt:3d} {f(res['auto'])} {f(res['single'])} {f(res['pair'])} {f(res['strip'])} {f(res['strip2'])}  "
              f"{fl / res['auto'] / 1e9:7.0f} {fl / best / 1e9:8.0f}  {roof:9.1f}", flush=True)
        for k_ in ("auto", "single", "pair", "strip", "strip2"):
            tot[k_] += cnt * (res[k_] if res[k_] is not None else res["auto"])
        tot["best"] += cnt * best
        rows.append((tag, shape, cnt, res["auto"], res["single"], res["pair"], res["strip"], res["strip2"], fl, by))
    print("TOTAL ms per step: " + "  ".join(f"{k_} {v:.3f}" for k_, v in tot.items()))
    if a.csv:
        with open(a.csv, "w") as fcsv:
            fcsv.write("tag,shape,count,auto_ms,single_ms,pair_ms,strip_ms,strip_pair_ms,gflop,mbytes\n")
            for r in rows:
                fcsv.write(",".join(str(x) for x in r[:3]) + "," + ",".join("" if x is None else f"{x:.5f}" for x in r[3:8]) +
                           f",{r[8] / 1e9:.3f},{r[9] / 1e6:.3f}\n")


if __name__ == "__main__":
    main()
