"""ctypes binding of libfce_yolo_b200.so (include/fce_yolo_b200.h).

Loading is strict: a missing library, a missing symbol or a non-sm_100 device raises - the product
path has no fallback of any kind.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libfce_yolo_b200.so")

BF16, F32, U8 = 0, 1, 2
ACT_NONE, ACT_SILU, ACT_SIGMOID = 0, 1, 2
NHWC, NCHW = 0, 1

STATUS = {0: "OK", -1: "bad argument", -2: "unsupported shape/dtype", -3: "misaligned view",
          -4: "workspace too small", -5: "CUDA error"}

i32, f32, f64, i64 = C.c_int32, C.c_float, C.c_double, C.c_int64


class ConvDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("B", "H", "W", "Cin", "Cout", "in_pitch", "in_off", "out_pitch", "out_off",
                                   "res_pitch", "res_off", "k", "stride", "act", "in_dtype", "w_dtype", "out_dtype",
                                   "in_layout")] + [("in_scale", f32), ("impl", i32), ("weighted", i32), ("out_scale", f32),
                                                    ("res_scale", f32), ("res_up", i32)]


class DetectEpiDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("mode", "A", "a_base", "rows", "reg_max")] + [("stride", f32)]


class PackDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("B", "H", "W", "Cin", "k", "stride", "Kpad", "in_dtype", "in_layout")]


class StemDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("B", "H", "W", "Cout", "out_pitch", "out_off", "act", "in_dtype", "in_layout")]


class Stem2Desc(C.Structure):
    _fields_ = [(n, i32) for n in ("B", "H", "W", "C0", "C1", "out_pitch", "out_off", "act0", "act1")]


class LetterboxItem(C.Structure):
    _fields_ = [("src", C.c_uint64)] + [(n, i32) for n in ("src_pitch", "H", "W", "new_w", "new_h", "top", "left",
                                                            "reserved")]


class DwconvDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("B", "H", "W", "C", "in_pitch", "in_off", "out_pitch", "out_off", "add_pitch",
                                   "add_off", "act", "dtype")]


class DwpwDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("B", "H", "W", "C", "Cout", "in_pitch", "in_off", "out_pitch", "out_off", "dw_act",
                                   "pw_act")]


class ChainDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("B", "H", "W", "c1", "cm", "c2", "Cout", "x1_pitch", "x1_off", "x2_pitch", "x2_off",
                                   "out_pitch", "out_off", "act1", "act2")]


class SppfDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("B", "H", "W", "C", "pitch", "off", "dtype")]


class UpsampleDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("B", "H", "W", "C", "in_pitch", "in_off", "out_pitch", "out_off", "dtype")]


class BifpnDesc(C.Structure):
    _fields_ = [("B", i32), ("H", i32), ("W", i32), ("C", i32), ("n", i32), ("pitch", i32 * 3), ("off", i32 * 3),
                ("up", i32 * 3), ("wn", f32 * 3), ("out_pitch", i32), ("out_off", i32), ("dtype", i32)]


class CopyDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("B", "H", "W", "C", "in_pitch", "in_off", "out_pitch", "out_off", "dtype")]


class PoolDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("B", "H", "W", "C", "pitch", "off", "dtype")]


class CoordAttMlpDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("rows_h", "rows_w", "C", "mip", "oup", "s_pitch", "out_pitch", "act1", "act2")]


class StripAttnDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("B", "heads", "dh", "Lq", "Lk")] + [("scale", f32)] + \
               [(n, i64) for n in ("q_bstride", "q_rstride", "k_bstride", "k_rstride", "v_bstride", "v_rstride",
                                   "o_bstride", "o_rstride")]


class GateDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("B", "H", "W", "C", "mode", "in_pitch", "in_off", "out_pitch", "out_off",
                                   "dtype")] + [(n, i64) for n in ("gh_bstride", "gh_rstride", "gw_bstride",
                                                                   "gw_rstride")]


class PsaDesc(C.Structure):
    _fields_ = [(n, i32) for n in ("B", "N", "heads", "kd", "hd", "qkv_pitch", "q_off", "k_off", "v_off",
                                   "out_pitch", "out_off", "dtype")] + [("scale", f32)]


class DecodeDesc(C.Structure):
    _fields_ = [("B", i32), ("nl", i32), ("nc", i32), ("reg_max", i32), ("H", i32 * 4), ("W", i32 * 4),
                ("stride", f32 * 4), ("raw_pitch", i32 * 4)]


class NmsDesc(C.Structure):
    _fields_ = [("B", i32), ("A", i32), ("nc", i32), ("conf_thres", f32), ("iou_thres", f64), ("max_det", i32),
                ("max_nms", i32), ("multi_label", i32), ("agnostic", i32), ("max_wh", f32), ("n_classes", i32)]


_P = C.c_void_p
_SIGS = {
    "fce_abi_version": (C.c_int, []),
    "fce_last_cuda_error": (C.c_char_p, []),
    "fce_device_ok": (C.c_int, []),
    "fce_conv2d": (C.c_int, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P]),
    "fce_conv2d_detect": (C.c_int, [C.POINTER(ConvDesc), C.POINTER(DetectEpiDesc), _P, _P, _P, _P, _P]),
    "fce_stem_pack": (C.c_int, [C.POINTER(PackDesc), _P, _P, _P]),
    "fce_stem_conv": (C.c_int, [C.POINTER(StemDesc), _P, _P, _P, _P, _P]),
    "fce_stem2_conv": (C.c_int, [C.POINTER(Stem2Desc), _P, _P, _P, _P, _P, _P, _P]),
    "fce_stem2_route": (C.c_int, [C.POINTER(Stem2Desc)]),
    "fce_match_predictions": (C.c_int, [_P, _P, _P, _P, _P, _P, i32, i32, i32, i32, _P, _P]),
    "fce_scale_boxes": (C.c_int, [_P, _P, _P, i32, i32, _P]),
    "fce_letterbox": (C.c_int, [_P, _P, _P, i32, i32, i32, i32, _P, _P]),
    "fce_conv2d_route": (C.c_int, [C.POINTER(ConvDesc), _P, _P, _P, _P]),
    "fce_conv_stats": (C.c_int, [C.POINTER(C.c_longlong), C.c_int]),
    "fce_dwconv3x3": (C.c_int, [C.POINTER(DwconvDesc), _P, _P, _P, _P, _P, _P]),
    "fce_dwpw_conv": (C.c_int, [C.POINTER(DwpwDesc), _P, _P, _P, _P, _P, _P, _P]),
    "fce_dwpw_route": (C.c_int, [C.POINTER(DwpwDesc)]),
    "fce_conv1x1_chain": (C.c_int, [C.POINTER(ChainDesc), _P, _P, _P, _P, _P, _P, _P, _P]),
    "fce_conv1x1_chain_route": (C.c_int, [C.POINTER(ChainDesc)]),
    "fce_sppf_pool": (C.c_int, [C.POINTER(SppfDesc), _P, _P]),
    "fce_upsample2x": (C.c_int, [C.POINTER(UpsampleDesc), _P, _P, _P]),
    "fce_bifpn_fuse": (C.c_int, [C.POINTER(BifpnDesc), _P, _P, _P, _P, _P]),
    "fce_copy_view": (C.c_int, [C.POINTER(CopyDesc), _P, _P, _P]),
    "fce_coord_pool": (C.c_int, [C.POINTER(PoolDesc), _P, _P, _P, C.c_size_t, _P]),
    "fce_coord_pool_workspace": (C.c_size_t, [C.POINTER(PoolDesc)]),
    "fce_coordatt_mlp": (C.c_int, [C.POINTER(CoordAttMlpDesc), _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "fce_strip_attn": (C.c_int, [C.POINTER(StripAttnDesc), _P, _P, _P, _P, _P]),
    "fce_gate_apply": (C.c_int, [C.POINTER(GateDesc), _P, _P, _P, _P, _P]),
    "fce_psa_attention": (C.c_int, [C.POINTER(PsaDesc), _P, _P, _P]),
    "fce_detect_decode": (C.c_int, [C.POINTER(DecodeDesc), _P, _P, _P, _P, _P, _P]),
    "fce_nms_workspace": (C.c_size_t, [C.POINTER(NmsDesc)]),
    "fce_nms": (C.c_int, [C.POINTER(NmsDesc), _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
}

_lib = None


class FceLibraryError(RuntimeError):
    pass


def load(check_device: bool = False):
    """Returns the loaded CDLL with typed signatures; raises FceLibraryError if unavailable."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FceLibraryError(
                f"{LIB_PATH} is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(fce_yolo_b200 has no CPU or PyTorch fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise FceLibraryError(f"{LIB_PATH} does not export {name}") from e
            fn.restype, fn.argtypes = res, args
        # debug builds only (FCE_DEBUG=1 python fce_yolo_b200/build.py): not part of the product ABI, not in the header
        for name, (res, args) in {"fce_conv_tc_set_profile": (None, [C.c_int]),
                                  "fce_conv_tc_profile": (C.c_int, [C.POINTER(C.c_longlong), C.c_int])}.items():
            fn = getattr(lib, name, None)
            if fn is not None:
                fn.restype, fn.argtypes = res, args
        _lib = lib
    if check_device and not _lib.fce_device_ok():
        raise FceLibraryError("fce_yolo_b200 kernels are built for sm_100a (B200) only; no such device is current")
    return _lib


def exported_symbols():
    return list(_SIGS)


def check(status: int, what: str):
    if status == 0:
        return
    msg = STATUS.get(status, f"status {status}")
    if status == -5:
        msg += ": " + load().fce_last_cuda_error().decode()
    if status in (-1, -2, -3):
        raise ValueError(f"{what}: {msg}")
    raise RuntimeError(f"{what}: {msg}")
