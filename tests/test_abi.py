"""The C-ABI library loads and exports every symbol include/fce_yolo_b200.h declares (no GPU needed)."""
import os
import re

from helpers import ROOT


def test_header_symbols_are_exported():
    from fce_yolo_b200 import _lib

    hdr = open(os.path.join(ROOT, "include", "fce_yolo_b200.h")).read()
    declared = set(re.findall(r"\b(fce_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 15
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())
    assert lib.fce_abi_version() == 1


def test_desc_struct_sizes_match_header():
    """ctypes mirrors must have the C struct sizes (compile a tiny C program against the header)."""
    import ctypes
    import subprocess
    import tempfile

    from fce_yolo_b200 import _lib as L

    names = {"fce_conv_desc": L.ConvDesc, "fce_stem_desc": L.StemDesc, "fce_letterbox_item": L.LetterboxItem, "fce_pack_desc": L.PackDesc, "fce_dwconv_desc": L.DwconvDesc, "fce_sppf_desc": L.SppfDesc,
             "fce_upsample_desc": L.UpsampleDesc, "fce_bifpn_desc": L.BifpnDesc, "fce_copy_desc": L.CopyDesc,
             "fce_pool_desc": L.PoolDesc, "fce_strip_attn_desc": L.StripAttnDesc, "fce_gate_desc": L.GateDesc,
             "fce_psa_desc": L.PsaDesc, "fce_decode_desc": L.DecodeDesc, "fce_nms_desc": L.NmsDesc,
             "fce_coordatt_mlp_desc": L.CoordAttMlpDesc, "fce_detect_epi_desc": L.DetectEpiDesc, "fce_dwpw_desc": L.DwpwDesc,
             "fce_chain_desc": L.ChainDesc}
    src = '#include <stdio.h>\n#include "fce_yolo_b200.h"\nint main(){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe], text=True)
    for line in out.strip().splitlines():
        n, sz = line.split()
        assert ctypes.sizeof(names[n]) == int(sz), n


def test_missing_gpu_fails_loudly():
    import pytest
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fce_yolo_b200.tasks import DetectionModel

    m = DetectionModel("yolo11n-fce.yaml").fuse().eval()
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 3, 64, 64))
