"""Case tables shared by make_golden.py (reference side) and the tests (oracle / CUDA side)."""
import torch

from fce_yolo_b200.tasks import variant_cfg  # noqa: F401  (re-exported for make_golden.py / tests)
from fce_yolo_b200.weights import synth_predictions


FORWARD_CASES = {
    # every layer of the shipped FCE graph, small image
    "n_fce_64": dict(yaml="yolo11n-fce.yaml", seed=0, img_seed=11, batch=2, size=64, layers="all"),
    # BASELINE config 2: CoordAtt at L5/L8
    "s_coordatt_64": dict(yaml="yolo11s-fce.yaml", seed=1, img_seed=12, batch=1, size=64, layers=[4, 5, 7, 8],
                          variant={5: ("CoordAtt", []), 8: ("CoordAtt", [])}),
    # BASELINE config 4: CoordCrossAtt + 8-head BiCoordCrossAtt
    "s_cca_bicca8_64": dict(yaml="yolo11s-fce.yaml", seed=2, img_seed=13, batch=1, size=64, layers=[4, 5, 7, 8],
                            variant={5: ("CoordCrossAtt", [512, 16, 2]), 8: ("BiCoordCrossAtt", [512, 8, 8])}),
    # BASELINE config 3: m-scale BiFPN graph (C3k everywhere, one realign conv)
    "m_bifpn_64": dict(yaml="yolo11m-bifpn.yaml", seed=3, img_seed=14, batch=1, size=64, layers=[10, 12, 15, 18, 21]),
    # BASELINE config 5 graph (depth 1.0: n=2 repeats), tiny image
    "x_fce_64": dict(yaml="yolo11x-fce.yaml", seed=4, img_seed=15, batch=1, size=64, layers=[5, 8, 12]),
    # full-size image, output sub-sampled along anchors
    "n_fce_640": dict(yaml="yolo11n-fce.yaml", seed=0, img_seed=16, batch=1, size=640, layers=[], y_stride=8),
    # stock Concat neck
    "n_stock_64": dict(yaml="yolo11n.yaml", seed=5, img_seed=17, batch=1, size=64, layers=[12, 15]),
}

MODULE_CASES = {
    # inp != oup exercises the identity 1x1 (fce_block.py:95, 233); non-square maps
    "coordatt_64_96": dict(cls="CoordAtt", args=[64, 96, 8], seed=21, img_seed=31, batch=2, inputs=[(64, 12, 20)]),
    "coordatt_mip11": dict(cls="CoordAtt", args=[128, 128, 11], seed=22, img_seed=32, batch=1, inputs=[(128, 16, 16)]),
    "cca_64_h2": dict(cls="CoordCrossAtt", args=[64, 64, 8, 2], seed=23, img_seed=33, batch=2, inputs=[(64, 10, 24)]),
    "cca_mip11_h1": dict(cls="CoordCrossAtt", args=[128, 128, 11, 1], seed=24, img_seed=34, batch=1, inputs=[(128, 8, 8)]),
    "bicca_64_96": dict(cls="BiCoordCrossAtt", args=[64, 96, 8, 4], seed=25, img_seed=35, batch=2, inputs=[(64, 12, 20)]),
    "bicca_d1": dict(cls="BiCoordCrossAtt", args=[64, 64, 16, 8], seed=26, img_seed=36, batch=1, inputs=[(64, 16, 16)]),
    "bifpn3": dict(cls="BiFPN_Concat", args=[[64, 128, 128], 32], seed=27, img_seed=37, batch=2,
                   inputs=[(64, 10, 10), (128, 10, 10), (128, 10, 10)]),
    "bifpn2_id": dict(cls="BiFPN_Concat", args=[[64, 64], 64], seed=28, img_seed=40, batch=1,
                      inputs=[(64, 8, 12), (64, 8, 12)]),
}


def _ties():
    # exact score ties + IoU exactly at threshold + disjoint boxes (SURVEY D)
    p = torch.zeros(1, 84, 64)
    for a in range(64):
        p[0, 0, a] = 50 + (a % 8) * 30
        p[0, 1, a] = 50 + (a // 8) * 5  # heavy vertical overlap inside a column
        p[0, 2, a] = 20
        p[0, 3, a] = 20
        p[0, 4 + (a % 3), a] = 0.5 if a % 2 else 0.75
    return p


def _class79():
    # fp32 class offset 79*7680 quantises coordinates to 1/16 px (nms.py:143,149)
    p = synth_predictions(9, 1, 2048, sharp=2.0)
    p[:, 4:] *= 0.1
    p[:, 4 + 79] = synth_predictions(10, 1, 2048, sharp=1.0)[:, 4]
    p[:, 0:2] = p[:, 0:2] * 0.2 + 300  # crowd them so suppression decisions sit near the threshold
    return p


def _empty():
    p = synth_predictions(11, 2, 512)
    p[:, 4:] *= 0.01
    return p


NMS_CASES = {
    "predict": dict(make=lambda: synth_predictions(7, 2, 8400), kw=dict(conf_thres=0.25, iou_thres=0.7, max_det=300)),
    "predict_iou45": dict(make=lambda: synth_predictions(8, 1, 8400, sharp=4.0),
                          kw=dict(conf_thres=0.25, iou_thres=0.45, max_det=300)),
    "val_multilabel": dict(make=lambda: synth_predictions(7, 1, 8400, sharp=3.0),
                           kw=dict(conf_thres=0.001, iou_thres=0.7, max_det=300, multi_label=True), tie_perm=True),
    "agnostic": dict(make=lambda: synth_predictions(12, 1, 4200), kw=dict(conf_thres=0.25, iou_thres=0.5, agnostic=True)),
    "classes": dict(make=lambda: synth_predictions(13, 1, 4200), kw=dict(conf_thres=0.25, iou_thres=0.7, classes=[0, 3, 79])),
    "ties": dict(make=_ties, kw=dict(conf_thres=0.25, iou_thres=0.6)),
    "class79": dict(make=_class79, kw=dict(conf_thres=0.25, iou_thres=0.7, max_det=300)),
    "empty": dict(make=_empty, kw=dict(conf_thres=0.25, iou_thres=0.7)),
    "maxdet20": dict(make=lambda: synth_predictions(14, 1, 8400), kw=dict(conf_thres=0.25, iou_thres=0.7, max_det=20)),
    # max_nms below the candidate count, both <= 1024: the kept set must be the BEST max_nms by score (nms.py:136-140)
    "maxnms_cut": dict(make=lambda: synth_predictions(16, 2, 700, sharp=4.0),
                       kw=dict(conf_thres=0.25, iou_thres=0.7, max_det=300, max_nms=150), tie_perm=True),
    "a33600": dict(make=lambda: synth_predictions(15, 1, 33600, imgsz=1280, sharp=1.5),
                   kw=dict(conf_thres=0.25, iou_thres=0.7, max_det=300), tie_perm=True),
}


# LetterBox + BGR->RGB: (source h, w), new_shape, LetterBox kwargs.  Small sizes keep the fixtures small.
LETTERBOX_CASES = {
    "lb_up_landscape": dict(shape=(45, 70), new_shape=(96, 96), kw={}, seed=41),
    "lb_down_portrait": dict(shape=(201, 133), new_shape=(96, 96), kw={}, seed=42),
    "lb_auto_rect": dict(shape=(120, 213), new_shape=(128, 128), kw=dict(auto=True), seed=43),
    "lb_noscaleup": dict(shape=(40, 52), new_shape=(96, 96), kw=dict(scaleup=False), seed=44),
    "lb_same_size": dict(shape=(96, 96), new_shape=(96, 96), kw={}, seed=45),
    "lb_half": dict(shape=(192, 192), new_shape=(96, 96), kw={}, seed=46),
}


def letterbox_image(case):
    import numpy as np

    rng = np.random.default_rng(case["seed"])
    return rng.integers(0, 256, (case["shape"][0], case["shape"][1], 3), dtype=np.uint8)


# ops.scale_boxes + clip: letterboxed shape, original shape, seed
SCALE_BOXES_CASES = {
    "sb_landscape": dict(img1=(640, 640), img0=(480, 640), seed=51),
    "sb_hd": dict(img1=(384, 640), img0=(1080, 1920), seed=52),
    "sb_portrait_up": dict(img1=(640, 640), img0=(333, 217), seed=53),
    "sb_1280": dict(img1=(1280, 1280), img0=(3000, 4000), seed=54),
}


def scale_boxes_input(case, n=64):
    import numpy as np

    rng = np.random.default_rng(case["seed"])
    h, w = case["img1"]
    xy = rng.uniform(-20, 1, (n, 2)).astype(np.float32) * 0 + rng.uniform(-30, max(h, w) + 30, (n, 2)).astype(np.float32)
    wh = rng.uniform(1, 300, (n, 2)).astype(np.float32)
    return np.concatenate([xy, xy + wh], 1).astype(np.float32)


# validation matching: n_pred, n_gt, nc, seed.  Ground truth = jittered copies of some predictions so
# that IoUs spread over [0.5, 1]; several predictions per label exercise the "lowest index wins" rule.
MATCH_CASES = {
    "match_dense": dict(n_pred=120, n_gt=40, nc=3, seed=61),
    "match_sparse": dict(n_pred=300, n_gt=7, nc=80, seed=62),
    "match_crowd": dict(n_pred=64, n_gt=90, nc=2, seed=63),
    "match_one": dict(n_pred=1, n_gt=1, nc=1, seed=64),
    "match_nopred": dict(n_pred=0, n_gt=5, nc=4, seed=65),
    "match_nogt": dict(n_pred=9, n_gt=0, nc=4, seed=66),
}


def match_inputs(case):
    import numpy as np

    rng = np.random.default_rng(case["seed"])
    n, g, nc = case["n_pred"], case["n_gt"], case["nc"]
    xy = rng.uniform(0, 560, (max(g, 1), 2))
    wh = rng.uniform(20, 120, (max(g, 1), 2))
    gt = np.concatenate([xy, xy + wh], 1).astype(np.float32)[:g]
    gt_cls = rng.integers(0, nc, g).astype(np.float32)
    pred = np.zeros((n, 4), dtype=np.float32)
    pred_cls = rng.integers(0, nc, n).astype(np.float32)
    for d in range(n):
        if g and rng.random() < 0.8:
            l = rng.integers(0, g)
            jit = rng.normal(0, rng.choice([1.0, 4.0, 10.0]), 4)
            pred[d] = gt[l] + jit
            if rng.random() < 0.85:
                pred_cls[d] = gt_cls[l]
        else:
            p = rng.uniform(0, 560, 2)
            pred[d] = np.concatenate([p, p + rng.uniform(20, 120, 2)])
    return pred.astype(np.float32), pred_cls, gt, gt_cls
