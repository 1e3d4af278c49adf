"""Drop-in installation behind the UNCHANGED reference classes.

    from ultralytics import YOLO
    import fce_yolo_b200
    yolo = YOLO("yolo11s-fce.yaml")
    fce_yolo_b200.install(yolo.model, precision="bf16")     # DetectionModel (ultralytics/nn/tasks.py:339)
    yolo.predict(tensor_or_images, device=0)                # reference API, B200 kernels underneath

``install`` keeps every reference class, parameter and state-dict key; it rebinds ONE method on the instance -
``BaseModel._predict_once`` (tasks.py:160-188), the layer loop - to the compiled-plan executor, and
``install_nms`` rebinds the module attribute ``ultralytics.utils.nms.non_max_suppression`` that
``detect/predict.py:54``, ``detect/val.py:115`` and ``autobackend.py:925`` look up at call time.

The plan compiler dispatches on class *names* and attribute names (``cv1``, ``m``, ``conv``, ``realign_convs`` ...)
so it walks the reference's instances exactly like this package's mirror classes.  Packed weights are cached per
(input shape, dtype, device) and rebuilt when any parameter's (data_ptr, version) changes - ``.to()/.half()/
load_state_dict()/fuse()`` on the reference model therefore just work.

Eval-mode CUDA inference is the hot path.  What is NOT the hot path is handed back to the reference's own method
untouched: training mode (loss needs the autograd graph) and ``profile/visualize/embed`` requests.  A CPU tensor in
eval mode raises - there is no CPU fallback of the kernels.
"""
from __future__ import annotations

import types

import torch

_ORIG = "_fce_orig_predict_once"


def _fast_predict_once(self, x, profile=False, visualize=False, embed=None):
    if self.training or profile or visualize or embed:
        return getattr(self, _ORIG)(x, profile, visualize, embed)
    from .engine import run_model

    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        # the reference probes strides / warms up on CPU tensors only at construction time, before install()
        raise RuntimeError("fce_yolo_b200.install: eval-mode forward needs a CUDA tensor (no CPU fallback); "
                           "call uninstall(model) to restore the reference path")
    prec = getattr(self, "fce_precision", None)
    if prec is None:  # follow the model's dtype, as AutoBackend's .half()/.float() would (autobackend.py:208,219)
        prec = "fp32" if next(self.parameters()).dtype == torch.float32 else "bf16"
    y, raw = run_model(self, x, precision=prec)
    # run_model's outputs alias the executor's arena and are overwritten by the next forward; the reference returns fresh
    # tensors (callers keep predictions across batches, compare two forwards, do TTA) - so this boundary copies
    y, raw = y.clone(), [r.clone() for r in raw]
    det = self.model[-1]
    if getattr(det, "export", False):
        return y.to(x.dtype) if x.dtype != torch.uint8 else y
    # eval-mode Detect returns (y, x_list) (head.py:121-124); y in the caller's dtype like the reference
    if x.dtype in (torch.float16, torch.bfloat16):
        return y.to(x.dtype), [r.to(x.dtype) for r in raw]
    return y, raw


def install(model, precision: str | None = "bf16", patch_nms: bool = True):
    """Routes ``model`` (a reference or mirror DetectionModel) through the B200 plan.  precision: 'bf16' (bf16
    storage, fp32 accumulate - BASELINE's mode), 'fp32' (1e-4 parity mode) or None (follow the model's dtype)."""
    from .plan import PlanError, compile_model

    if not hasattr(model, "model") or type(model.model[-1]).__name__ != "Detect":
        raise PlanError("install() expects a detection model whose last layer is Detect")
    if precision not in (None, "bf16", "fp32"):
        raise PlanError(f"precision must be 'bf16', 'fp32' or None, got {precision}")
    # validate now: unsupported modules / argument combinations fail at install time, not mid-predict
    compile_model(model, 1, 64, 64, precision or "bf16", torch.device("cpu"))
    if not hasattr(model, _ORIG):
        object.__setattr__(model, _ORIG, model._predict_once)
    model.fce_precision = precision
    object.__setattr__(model, "_predict_once", types.MethodType(_fast_predict_once, model))
    if patch_nms and type(model).__module__.split(".")[0] == "ultralytics":
        install_nms()  # the mirror DetectionModel has no reference NMS module to patch
    return model


def uninstall(model):
    if hasattr(model, _ORIG):
        object.__setattr__(model, "_predict_once", getattr(model, _ORIG))
        object.__delattr__(model, _ORIG)
    return model


_NMS_ORIG = None


def install_nms():
    """Rebinds ultralytics.utils.nms.non_max_suppression (nms.py:13-29, same signature) to the GPU batched NMS
    for CUDA inputs of the plain detection kind; everything else (rotated, end2end, a-priori labels, CPU
    tensors) goes to the original function."""
    global _NMS_ORIG
    try:
        from ultralytics.utils import nms as ref
    except ImportError as e:  # the mirror-only deployment has nothing to patch
        raise RuntimeError("install_nms() needs the reference package importable") from e
    if _NMS_ORIG is not None:
        return
    from .nms import non_max_suppression as fast

    orig = ref.non_max_suppression

    def non_max_suppression(prediction, conf_thres=0.25, iou_thres=0.45, classes=None, agnostic=False,
                            multi_label=False, labels=(), max_det=300, nc=0, max_time_img=0.05, max_nms=30000,
                            max_wh=7680, rotated=False, end2end=False, return_idxs=False):
        p = prediction[0] if isinstance(prediction, (list, tuple)) else prediction
        plain = not (rotated or end2end or p.shape[-1] == 6 or (labels and any(len(l) for l in labels)))
        if plain and p.is_cuda:
            return fast(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, labels, max_det, nc,
                        max_time_img, max_nms, max_wh, rotated, end2end, return_idxs)
        return orig(prediction, conf_thres, iou_thres, classes, agnostic, multi_label, labels, max_det, nc,
                    max_time_img, max_nms, max_wh, rotated, end2end, return_idxs)

    _NMS_ORIG = orig
    ref.non_max_suppression = non_max_suppression


def uninstall_nms():
    global _NMS_ORIG
    if _NMS_ORIG is not None:
        from ultralytics.utils import nms as ref

        ref.non_max_suppression = _NMS_ORIG
        _NMS_ORIG = None
