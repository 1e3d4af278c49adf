"""``YOLO(...).val()`` of the UNMODIFIED reference (baseline/_ref) with ``fce_yolo_b200.install()`` on the GPU: the
reference's own validator / dataloader / metrics code (engine/validator.py:130-264, models/yolo/detect/val.py:105-211,
utils/metrics.py) runs unchanged, only ``_predict_once`` and ``non_max_suppression`` are rebound.  On a small synthetic
dataset whose labels are the reference's own top predictions, the metrics of the installed run must equal the metrics of
the reference's plain CPU run (fp32 mode: 1e-3; bf16 mode: close)."""
import os
import sys

import numpy as np
import pytest

import detection_parity as DP

sys.path.insert(0, os.path.join(DP.ROOT, "baseline"))
import ref_env  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_env.installed(), reason="baseline/_ref not built (no reference tree at build time)")]
KEYS = ("metrics/precision(B)", "metrics/recall(B)", "metrics/mAP50(B)", "metrics/mAP50-95(B)")


def _dataset(root, yolo, n=8, size=320):
    import cv2
    import yaml

    for sub in ("images/train", "images/val", "labels/train", "labels/val"):
        os.makedirs(os.path.join(root, sub), exist_ok=True)
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 256, (size, size, 3), dtype=np.uint8) for _ in range(n)]
    res = yolo.predict(imgs, device="cpu", imgsz=size, conf=0.25, iou=0.7, verbose=False, save=False)
    for i, (im, r) in enumerate(zip(imgs, res)):
        lines = [f"{int(r.boxes.cls[k])} " + " ".join(f"{v:.6f}" for v in r.boxes.xywhn[k].tolist())
                 for k in range(min(6, len(r.boxes)))]
        for split in ("train", "val"):
            cv2.imwrite(os.path.join(root, f"images/{split}/{i:03d}.png"), im)  # lossless: both runs see the same pixels
            with open(os.path.join(root, f"labels/{split}/{i:03d}.txt"), "w") as f:
                f.write("\n".join(lines) + "\n")
    with open(os.path.join(root, "data.yaml"), "w") as f:
        yaml.safe_dump({"path": root, "train": "images/train", "val": "images/val", "names": {i: f"c{i}" for i in range(80)}}, f)
    return os.path.join(root, "data.yaml")


def test_reference_val_with_install_matches_reference_cpu_val(tmp_path):
    import fce_yolo_b200

    ref_env.offline_val_env()
    yolo = ref_env.reference_yolo("yolo11n-fce.yaml", None, 0)
    data = _dataset(str(tmp_path / "ds"), yolo)
    # rect=False: square letterbox without the validator's half-stride padding (labels come from un-padded predictions)
    kw = dict(data=data, imgsz=320, batch=4, plots=False, workers=0, verbose=False, half=False, rect=False)
    ref = yolo.val(device="cpu", **kw).results_dict
    assert ref["metrics/mAP50(B)"] > 0.1  # not vacuous: the labels are found
    for prec, tol in (("fp32", 1e-3), ("bf16", 0.08)):
        fce_yolo_b200.install(yolo.model, precision=prec)
        try:
            got = yolo.val(device=0, **kw).results_dict
        finally:
            fce_yolo_b200.uninstall(yolo.model)
            fce_yolo_b200.uninstall_nms()
        for k in KEYS:
            assert abs(got[k] - ref[k]) <= tol, (prec, k, got[k], ref[k])
