"""Generates tests/golden/*.npz from the LIVE reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 YOLO_CONFIG_DIR=/tmp/ulcfg python tests/golden/make_golden.py

Imports /root/reference/ultralytics, builds each model through the reference's own
DetectionModel/parse_model, fuses it, loads the synthetic state dict from
fce_yolo_b200.weights (same keys - load is strict), runs the reference forward and stores what
the parity tests compare against.  /root/reference does not exist on the GPU box, hence the
committed fixtures.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

from ultralytics.nn.tasks import DetectionModel, yaml_model_load  # noqa: E402
from ultralytics.nn.modules import fce_block  # noqa: E402
from ultralytics.utils import nms as ref_nms  # noqa: E402

from fce_yolo_b200.weights import load_synthetic, synth_images, synth_predictions, synth_tensor  # noqa: E402

from cases import (FORWARD_CASES, LETTERBOX_CASES, MATCH_CASES, MODULE_CASES, NMS_CASES, SCALE_BOXES_CASES,  # noqa: E402
                   letterbox_image, match_inputs, scale_boxes_input, variant_cfg)


def build_ref(case):
    d = yaml_model_load(case["yaml"])
    d = variant_cfg(d, case.get("variant"))
    m = DetectionModel(d, verbose=False).eval()
    m.fuse()
    load_synthetic(m, case["seed"])
    return m


def forward_case(name, case):
    m = build_ref(case)
    x = synth_images(case["img_seed"], case["batch"], case["size"], case["size"])
    outs = []
    hooks = [l.register_forward_hook(lambda mod, i, o: outs.append(o)) for l in m.model]
    with torch.no_grad():
        y, raw = m(x)
    for h in hooks:
        h.remove()
    blob = {}
    keep = case["layers"]
    for i, o in enumerate(outs[:-1]):
        if keep == "all" or i in keep:
            blob[f"layer{i}"] = o.numpy()
    sub = case.get("y_stride", 1)
    blob["y"] = y.numpy()[:, :, ::sub]
    for i, r in enumerate(raw):
        blob[f"raw{i}_sum"] = np.array([r.double().sum().item(), r.double().abs().sum().item()])
    np.savez_compressed(os.path.join(HERE, f"fwd_{name}.npz"), **blob)
    print(name, {k: v.shape for k, v in blob.items() if k in ("y",)}, "layers:", len(blob))


def module_case(name, case):
    cls = getattr(fce_block, case["cls"])
    mod = cls(*case["args"]).eval()
    # realign Conv inside BiFPN_Concat / cv1 inside CoordAtt carry BN: fold it the reference way
    from ultralytics.nn.modules.conv import Conv
    from ultralytics.utils.torch_utils import fuse_conv_and_bn
    for sub in mod.modules():
        if isinstance(sub, Conv) and hasattr(sub, "bn"):
            sub.conv = fuse_conv_and_bn(sub.conv, sub.bn)
            delattr(sub, "bn")
            sub.forward = sub.forward_fuse
    sd = mod.state_dict()
    for k, v in sd.items():
        v.copy_(torch.from_numpy(synth_tensor(case["seed"], "mod." + k, v.shape)))
    xs = [synth_images(case["img_seed"] + j, case["batch"], h, w, c) * 4 - 2 for j, (c, h, w) in enumerate(case["inputs"])]
    with torch.no_grad():
        y = mod(xs if case["cls"] == "BiFPN_Concat" else xs[0])
    np.savez_compressed(os.path.join(HERE, f"mod_{name}.npz"), y=y.numpy())
    print(name, tuple(y.shape))


def nms_case(name, case):
    p = case["make"]()
    kw = dict(case["kw"])
    out, idx = ref_nms.non_max_suppression(p.clone(), return_idxs=True, **kw)
    blob = {}
    for b, (o, i) in enumerate(zip(out, idx)):
        blob[f"det{b}"] = o.numpy().astype(np.float32).reshape(-1, 6)
        blob[f"idx{b}"] = i.numpy().astype(np.int64).reshape(-1)
    np.savez_compressed(os.path.join(HERE, f"nms_{name}.npz"), **blob)
    print(name, [v.shape[0] for k, v in blob.items() if k.startswith("idx")])


def letterbox_case(name, case):
    """The reference's own preprocessing: LetterBox (data/augment.py:1589-1631, cv2.resize + copyMakeBorder) followed
    by the BGR->RGB flip of BasePredictor.preprocess (engine/predictor.py:163-165)."""
    from ultralytics.data.augment import LetterBox

    img = letterbox_image(case)
    out = LetterBox(case["new_shape"], stride=32, **case["kw"])(image=img)[..., ::-1]
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), out=np.ascontiguousarray(out))
    print(name, img.shape, "->", out.shape)


def scale_boxes_case(name, case):
    from ultralytics.utils import ops

    b = torch.from_numpy(scale_boxes_input(case))
    out = ops.scale_boxes(case["img1"], b.clone(), case["img0"] + (3,))
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), out=out.numpy())
    print(name, tuple(out.shape))


def match_case(name, case):
    """DetectionValidator._process_batch (models/yolo/detect/val.py:274-288) of the live reference."""
    from ultralytics.models.yolo.detect.val import DetectionValidator

    v = DetectionValidator.__new__(DetectionValidator)  # only iouv / niou are used by the two methods
    v.iouv = torch.linspace(0.5, 0.95, 10)
    v.niou = 10
    pred, pred_cls, gt, gt_cls = match_inputs(case)
    out = v._process_batch({"bboxes": torch.from_numpy(pred), "cls": torch.from_numpy(pred_cls)},
                           {"bboxes": torch.from_numpy(gt), "cls": torch.from_numpy(gt_cls)})["tp"]
    np.savez_compressed(os.path.join(HERE, f"{name}.npz"), tp=np.asarray(out, dtype=bool))
    print(name, out.shape, int(np.asarray(out).sum()))


if __name__ == "__main__":
    import torchvision  # noqa: F401  (the branch ultralytics takes in practice: nms.py:151-154)
    which = sys.argv[1:] or ["fwd", "mod", "nms", "lb"]
    if "match" in which or "lb" in which:
        for n, c in MATCH_CASES.items():
            match_case(n, c)
    if "sb" in which or "lb" in which:
        for n, c in SCALE_BOXES_CASES.items():
            scale_boxes_case(n, c)
    if "lb" in which:
        for n, c in LETTERBOX_CASES.items():
            letterbox_case(n, c)
    if "fwd" in which:
        for n, c in FORWARD_CASES.items():
            forward_case(n, c)
    if "mod" in which:
        for n, c in MODULE_CASES.items():
            module_case(n, c)
    if "nms" in which:
        for n, c in NMS_CASES.items():
            nms_case(n, c)
    for w in which:  # a single NMS case by name: python make_golden.py nms:maxnms_cut
        if w.startswith("nms:"):
            nms_case(w[4:], NMS_CASES[w[4:]])
