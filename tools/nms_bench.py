#!/usr/bin/env python
"""Times fce_nms (batch 64, 8400 anchors, 80 classes) on (a) low scores - no candidate passes conf, the
random-weight bench case, (b) SURVEY 8d synthetic scores (rand**8: ~300 kept per image), predict and val settings."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from fce_yolo_b200.nms import nms_batched


def synth(B, A=8400, nc=80, seed=7, low=False):
    g = torch.Generator().manual_seed(seed)
    p = torch.empty(B, 4 + nc, A)
    p[:, 0:2] = torch.rand(B, 2, A, generator=g) * 640
    p[:, 2:4] = torch.rand(B, 2, A, generator=g) * 100 + 5
    s = torch.rand(B, nc, A, generator=g)
    p[:, 4:] = s * 0.01 if low else s ** 8
    return p.cuda()


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
for name, low, kw in [("no candidates, predict", True, dict(conf_thres=0.25, iou_thres=0.7)),
                      ("rand**8, predict", False, dict(conf_thres=0.25, iou_thres=0.7)),
                      ("rand**8, val (multi-label, conf .001)", False, dict(conf_thres=0.001, iou_thres=0.7, multi_label=True))]:
    p = synth(B, low=low)
    ws = torch.empty(B * 8400 * 80 * 16, dtype=torch.uint8, device="cuda")
    det, keep, count = nms_batched(p, workspace=ws, **kw)
    # time the C-ABI call alone (pre-built arguments, outputs reused)
    import ctypes as C
    from fce_yolo_b200 import _lib as L
    lib = L.load()
    d = L.NmsDesc(B=B, A=8400, nc=80, conf_thres=kw["conf_thres"], iou_thres=kw["iou_thres"], max_det=300,
                  max_nms=30000, multi_label=int(kw.get("multi_label", False)), agnostic=0, max_wh=7680.0, n_classes=0)
    args = (C.byref(d), C.c_void_p(p.data_ptr()), C.c_void_p(0), C.c_void_p(det.data_ptr()), C.c_void_p(keep.data_ptr()),
            C.c_void_p(count.data_ptr()), C.c_void_p(ws.data_ptr()), C.c_size_t(ws.numel()),
            C.c_void_p(torch.cuda.current_stream().cuda_stream))
    ms = timeit(lambda: lib.fce_nms(*args))
    print(f"{name:42s} B={B}: {ms*1e3:9.1f} us   kept/img mean {count.float().mean().item():.1f}  "
          f"({p.numel()*4/ms/1e6:.0f} GB/s of pred)", flush=True)
