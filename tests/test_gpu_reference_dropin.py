"""install() on the UNMODIFIED reference, on the GPU, through the reference's own public API.

The reference package is the pip-installed copy under baseline/_ref (built in the build container by
__graft_entry__.build(), shipped to the GPU box with the snapshot).  ``YOLO(yaml).predict(tensor, device=0)`` runs the
reference's predictor / AutoBackend / Results code unchanged (engine/model.py:477-535, engine/predictor.py:276-381,
models/yolo/detect/predict.py:33-122); fce_yolo_b200.install() has rebound only ``DetectionModel._predict_once`` and
``ultralytics.utils.nms.non_max_suppression``.  The detections must equal the oracle's (forward + NMS restated on
the CPU): same count, same class ids, boxes within fp32 noise in fp32 mode and IoU >= 0.99 in bf16 mode.
"""
import os
import sys

import numpy as np
import pytest
import torch

import detection_parity as DP

sys.path.insert(0, os.path.join(DP.ROOT, "baseline"))
import ref_env  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_env.installed(), reason="baseline/_ref not built (no reference tree at build time)")]


def _oracle_dets(cfg, sd, x, conf, iou, max_det, H, W):
    from oracle import fce_oracle as O
    from oracle import nms_oracle

    y = O.forward(cfg, cfg["scale"], sd, x)[0]
    dets, idxs = nms_oracle.non_max_suppression(y.numpy(), conf, iou, max_det=max_det)
    out = []
    for d in dets:  # ops.scale_boxes with identical shapes: gain 1, pad 0, then clip_boxes (ops.py:102-134, 152)
        d = np.array(d, dtype=np.float32, copy=True)
        d[:, [0, 2]] = d[:, [0, 2]].clip(0, W)
        d[:, [1, 3]] = d[:, [1, 3]].clip(0, H)
        out.append(d)
    return out, idxs


@pytest.mark.parametrize("name,prec", [("cfg0_n_fce", "fp32"), ("cfg0_n_fce", "bf16"), ("cfg1_s_coordatt", "bf16"),
                                       ("cfg3_s_cca_bicca8", "fp32")])
def test_reference_yolo_predict_runs_on_the_b200_plan(name, prec):
    import fce_yolo_b200

    case = dict(DP.CONFIGS[name], size=320, batch=2)
    cfg, _, sd = DP.build(case)
    yolo = ref_env.reference_yolo(case["yaml"], case.get("variant"), case["seed"])
    model = yolo.model
    assert type(model).__module__ == "ultralytics.nn.tasks"  # the reference's own DetectionModel
    from ultralytics.utils import nms as ref_nms

    orig_nms = ref_nms.non_max_suppression
    from fce_yolo_b200.weights import synth_images

    S = case["size"]
    x = synth_images(77, case["batch"], S, S)
    fce_yolo_b200.install(model, precision=prec)
    try:
        assert ref_nms.non_max_suppression is not orig_nms
        lib_calls_before = fce_yolo_b200._lib.load().fce_abi_version()
        res = yolo.predict(x, device=0, half=False, imgsz=S, conf=0.25, iou=0.7, max_det=300, verbose=False, save=False)
        assert lib_calls_before == 1
        # the model the predictor ran IS the installed instance (AutoBackend keeps nn.Modules, autobackend.py:200-208)
        assert "_predict_once" in vars(yolo.predictor.model.model)
        ref_d, _ = _oracle_dets(cfg, sd, x, 0.25, 0.7, 300, S, S)
        assert len(res) == case["batch"]
        total = 0
        for r, d in zip(res, ref_d):
            got = r.boxes.data.float().cpu().numpy()
            assert r.boxes.data.is_cuda and got.shape[1] == 6
            if prec == "fp32":
                assert got.shape == d.shape
                assert np.array_equal(got[:, 5], d[:, 5])
                assert np.allclose(got[:, :4], d[:, :4], atol=2e-3, rtol=0)
                assert np.allclose(got[:, 4], d[:, 4], atol=1e-5, rtol=0)
                total += len(d)
            else:
                # bf16: match on (class, near-identical box): every oracle detection that survives in both runs
                n = 0
                for row in d:
                    same = got[got[:, 5] == row[5]]
                    if len(same) == 0:
                        continue
                    iou = DP._iou(same[:, :4].astype(np.float64), np.repeat(row[None, :4], len(same), 0).astype(np.float64))
                    n += int(iou.max() >= 0.99)
                assert n >= 0.85 * len(d), (n, len(d))
                total += n
        assert total >= 100  # hundreds of detections: not vacuous
    finally:
        fce_yolo_b200.uninstall(model)
        fce_yolo_b200.uninstall_nms()
    assert ref_nms.non_max_suppression is orig_nms


def test_reference_forward_equals_installed_forward_fp32():
    """Same reference DetectionModel instance, same CUDA tensor: its own eager forward (cuDNN) vs the installed plan in
    fp32 mode - prediction tensor and raw head maps within 1e-4 (TF32 off)."""
    import fce_yolo_b200
    from fce_yolo_b200.weights import synth_images

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    case = DP.CONFIGS["cfg2_m_bifpn"]
    model = ref_env.reference_model(case["yaml"], case.get("variant"), case["seed"]).cuda()
    x = synth_images(78, 2, 256, 256).cuda()
    with torch.no_grad():
        y0, raw0 = model(x)
    fce_yolo_b200.install(model, precision="fp32", patch_nms=False)
    try:
        with torch.no_grad():
            y1, raw1 = model(x)
            y2, _ = model(x.flip(0))
        # a second forward must not overwrite what the first returned (the reference returns fresh tensors)
        assert not torch.equal(y1, y2)
        rel = ((y1 - y0).norm() / y0.norm()).item()
        assert rel < 1e-4, rel
        for a, b in zip(raw1, raw0):
            assert ((a - b).norm() / b.norm()).item() < 1e-4
        with torch.no_grad():
            y3, _ = model(x)
        assert torch.equal(y1, y3)
    finally:
        fce_yolo_b200.uninstall(model)
