"""Spec-level detection parity of the PRODUCT path (test infrastructure; imported by tests/test_gpu_detections.py and
runnable as a script that writes the measured fractions to a JSON file for profiles/).

north_star: "final boxes within IoU >= 0.99 for at least 99.5 % of matched detections", "integer outputs must be
bit-exact: the NMS keep-indices and class ids given identical pre-NMS scores".

What runs: the SAME uint8 batch goes through ``Predictor.infer`` exactly as ``bench.py`` configures it (uint8 stem,
Detect decode fused into the head's tcgen05 epilogues in bf16 mode, NMS overlapped on a side stream, CUDA graphs)
and through the oracle (``oracle.fce_oracle.forward`` = restated reference forward, head.py:149-167 decode;
``oracle.nms_oracle.non_max_suppression`` = restated utils/nms.py:72-161 + torchvision nms).  Detections are matched
by KEPT ANCHOR INDEX (predict mode keeps at most one class per anchor, nms.py:120-122), then:

  * class ids of matched detections must be equal;
  * IoU(ours, oracle) >= 0.99 for >= 99.5 % of the matched detections;
  * fp32 mode is held against the fp32 oracle; bf16 mode against the oracle run in bf16 storage (the
    reference's own bf16 forward drifts from its fp32 forward, so bf16 is judged against a bf16 run of the same
    math), and its fractions against the fp32 oracle are recorded next to it;
  * "given identical pre-NMS scores": the oracle NMS applied to OUR prediction tensor must reproduce our keep indices
    and class ids bit for bit (this is the integer half of the criterion, at BASELINE image size, through the
    overlapped graph path).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)

# The five BASELINE.json configs at their own image size; batch cut to what the CPU oracle finishes in seconds.
CONFIGS = {
    "cfg0_n_fce": dict(yaml="yolo11n-fce.yaml", variant=None, size=640, batch=4, seed=0),
    "cfg1_s_coordatt": dict(yaml="yolo11s-fce.yaml", variant={5: ("CoordAtt", []), 8: ("CoordAtt", [])}, size=640,
                            batch=4, seed=1),
    "cfg2_m_bifpn": dict(yaml="yolo11m-bifpn.yaml", variant=None, size=640, batch=4, seed=1),
    "cfg3_s_cca_bicca8": dict(yaml="yolo11s-fce.yaml", size=640, batch=4, seed=1,
                              variant={5: ("CoordCrossAtt", [512, 16, 2]), 8: ("BiCoordCrossAtt", [512, 8, 8])}),
    "cfg4_x_fce_1280": dict(yaml="yolo11x-fce.yaml", variant=None, size=1280, batch=2, seed=1),
    # beyond BASELINE.json: the remaining scale of the FCE graph (l: two repeats, C3k everywhere) and the stock Concat neck
    "extra_l_fce_320": dict(yaml="yolo11l-fce.yaml", variant=None, size=320, batch=2, seed=2),
    "extra_s_stock_320": dict(yaml="yolo11s.yaml", variant=None, size=320, batch=2, seed=3),
}
CONF, IOU, MAX_DET = 0.25, 0.7, 300


def build(case):
    from fce_yolo_b200.tasks import DetectionModel, variant_cfg, yaml_model_load
    from fce_yolo_b200.weights import load_synthetic

    cfg = variant_cfg(yaml_model_load(case["yaml"]), case.get("variant"))
    model = DetectionModel(cfg).fuse().eval()
    sd = load_synthetic(model, case["seed"])
    return cfg, model, sd


def u8_batch(seed, batch, size):
    from fce_yolo_b200.weights import synth_images

    return (synth_images(seed, batch, size, size) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()


def oracle_predictions(cfg, sd, img_u8, dtype=torch.float32):
    """[B, 4+nc, A] fp32 prediction tensor of the oracle for a uint8 NHWC batch.  dtype=bfloat16: weights and
    activations stored in bf16 like `model.bfloat16()` would (SURVEY finding 6), decode from the head's logits in fp32
    (SURVEY 7 "bf16 head precision": the product decodes in fp32 from fp32 accumulators)."""
    from oracle import fce_oracle as O

    x = img_u8.permute(0, 3, 1, 2).to(torch.float32) / 255.0  # predictor.py:168-172: uint8 -> float, /255
    if dtype == torch.float32:
        return O.forward(cfg, cfg["scale"], sd, x)[0]
    sdl = {k: (v.to(dtype) if v.is_floating_point() else v) for k, v in sd.items()}
    _, raw = O.forward(cfg, cfg["scale"], sdl, x.to(dtype))
    return O.detect_decode([r.float() for r in raw], (8, 16, 32))


def _iou(a, b):
    iw = np.clip(np.minimum(a[:, 2], b[:, 2]) - np.maximum(a[:, 0], b[:, 0]), 0, None)
    ih = np.clip(np.minimum(a[:, 3], b[:, 3]) - np.maximum(a[:, 1], b[:, 1]), 0, None)
    inter = iw * ih
    ua = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1]) + (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1]) - inter
    return inter / np.maximum(ua, 1e-12)


def match(det, keep, count, dets_o, idxs_o):
    """Matches by kept anchor index.  Returns a dict of counts and the IoU / class agreement of the matched pairs."""
    n_ours = n_or = n_match = n_cls_eq = 0
    ious = []
    for b in range(len(dets_o)):
        n = int(count[b])
        ours = {int(a): i for i, a in enumerate(keep[b, :n].tolist())}
        n_ours += n
        n_or += len(idxs_o[b])
        ia, ib = [], []
        for j, a in enumerate(np.asarray(idxs_o[b]).tolist()):
            if a in ours:
                ia.append(ours[a])
                ib.append(j)
        if not ia:
            continue
        A, Bm = det[b, ia], np.asarray(dets_o[b])[ib]
        n_match += len(ia)
        n_cls_eq += int((A[:, 5] == Bm[:, 5]).sum())
        ious.append(_iou(A[:, :4].astype(np.float64), Bm[:, :4].astype(np.float64)))
    ious = np.concatenate(ious) if ious else np.zeros(0)
    return dict(n_ours=n_ours, n_oracle=n_or, n_matched=n_match, n_class_equal=n_cls_eq,
                frac_iou99=float((ious >= 0.99).mean()) if len(ious) else 0.0,
                min_iou=float(ious.min()) if len(ious) else 0.0,
                matched_of_oracle=n_match / max(n_or, 1))


def evaluate(name, precision, device="cuda:0"):
    """Runs config `name` through the product path and both oracles; returns the statistics dict."""
    from fce_yolo_b200.predict import Predictor
    from oracle import nms_oracle

    case = CONFIGS[name]
    cfg, model, sd = build(case)
    B, S = case["batch"], case["size"]
    dev = torch.device(device)
    pred = Predictor(model, B, S, precision=precision, device=dev, conf=CONF, iou=IOU, max_det=MAX_DET, input_u8=True,
                     use_graph=True, overlap_nms=True, fuse_decode=True)  # bench.py's configuration
    img = u8_batch(1234, B, S)
    other = (255 - img).pin_memory()
    h = img.pin_memory()
    # a stream of different batches: the overlapped NMS of call i runs under the forward of call i+1
    pred.infer(other)
    pred.infer(h)
    pred.infer(other)
    det, count = [t.clone().numpy() for t in pred.infer(h)]
    keep = pred.keep.cpu().numpy()
    y_ours = pred.ex.outputs()[0].float().cpu()
    fns = [n.fn for n in pred.ex.plan.nodes]
    out = dict(config=name, precision=precision, batch=B, size=S, overlap=bool(pred.overlap),
               fused_decode="fce_conv2d_detect" in fns, launches=pred.launches_per_call)

    # (1) integer half: oracle NMS on OUR pre-NMS tensor -> identical keep indices, class ids, copied-through boxes
    dets_s, idxs_s = nms_oracle.non_max_suppression(y_ours.numpy(), CONF, IOU, max_det=MAX_DET)
    same = True
    for b in range(B):
        n = int(count[b])
        same &= n == len(idxs_s[b]) and bool((keep[b, :n] == np.asarray(idxs_s[b])).all())
        same &= bool((det[b, :n, 5] == np.asarray(dets_s[b])[:, 5]).all()) if n else True
        same &= bool(np.array_equal(det[b, :n, :5], np.asarray(dets_s[b])[:, :5])) if n else True
    out["nms_bit_exact_on_own_scores"] = bool(same)
    out["detections_per_image"] = float(count.mean())
    out["candidates_per_image"] = float((y_ours[:, 4:].amax(1) > CONF).sum(1).float().mean())

    # (2) float half: matched detections vs the oracle(s)
    y32 = oracle_predictions(cfg, sd, img, torch.float32)
    d32, i32 = nms_oracle.non_max_suppression(y32.numpy(), CONF, IOU, max_det=MAX_DET)
    out["vs_fp32_oracle"] = match(det, keep, count, d32, i32)
    # per-anchor boxes (no NMS selection effects): IoU of every decoded box against the fp32 oracle's box
    def xyxy(t):
        return np.stack([t[:, 0] - t[:, 2] / 2, t[:, 1] - t[:, 3] / 2, t[:, 0] + t[:, 2] / 2, t[:, 1] + t[:, 3] / 2], -1)
    a = xyxy(y_ours[:, :4].permute(0, 2, 1).reshape(-1, 4).double().numpy())
    b = xyxy(y32[:, :4].permute(0, 2, 1).reshape(-1, 4).double().numpy())
    iou_all = _iou(a, b)
    out["per_anchor_vs_fp32_oracle"] = dict(frac_iou99=float((iou_all >= 0.99).mean()), min_iou=float(iou_all.min()),
                                            score_max_abs_err=float((y_ours[:, 4:] - y32[:, 4:]).abs().max()))
    if precision == "bf16":
        y16 = oracle_predictions(cfg, sd, img, torch.bfloat16)
        d16, i16 = nms_oracle.non_max_suppression(y16.numpy(), CONF, IOU, max_det=MAX_DET)
        out["vs_bf16_oracle"] = match(det, keep, count, d16, i16)
        # how far the reference-style bf16 run itself is from the fp32 oracle (context for the two rows above)
        k16 = np.full((B, MAX_DET), -1, dtype=np.int64)
        dd16 = np.zeros((B, MAX_DET, 6), dtype=np.float32)
        c16 = np.zeros(B, dtype=np.int32)
        for bb in range(B):
            n = len(i16[bb])
            k16[bb, :n], dd16[bb, :n], c16[bb] = i16[bb], d16[bb], n
        out["bf16_oracle_vs_fp32_oracle"] = match(dd16, k16, c16, d32, i32)
        b16 = xyxy(y16[:, :4].permute(0, 2, 1).reshape(-1, 4).double().numpy())
        out["per_anchor_vs_bf16_oracle"] = dict(frac_iou99=float((_iou(a, b16) >= 0.99).mean()))
        out["per_anchor_bf16_oracle_vs_fp32_oracle"] = dict(frac_iou99=float((_iou(b16, b) >= 0.99).mean()))
    del pred
    torch.cuda.empty_cache()
    return out


def main():
    import argparse

    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "parity_detections.json"))
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    res = []
    for name in CONFIGS:
        if a.only and a.only not in name:
            continue
        for prec in ("fp32", "bf16"):
            r = evaluate(name, prec)
            print(json.dumps(r), flush=True)
            res.append(r)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
