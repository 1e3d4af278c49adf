"""Pins oracle/ against fixtures generated from the live reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from cases import FORWARD_CASES, MODULE_CASES, NMS_CASES
from helpers import assert_nms_equal, golden, load_cfg, rel_l2, rel_max

from oracle import fce_oracle as O
from oracle import nms_oracle


def _oracle_sd(cfg, scale, seed):
    """State dict for the oracle: keys + shapes come from the mirror model (fused)."""
    from fce_yolo_b200.tasks import DetectionModel
    from fce_yolo_b200.weights import load_synthetic

    m = DetectionModel(cfg, scale=scale).fuse()
    return load_synthetic(m, seed)


@pytest.mark.parametrize("name", list(FORWARD_CASES))
def test_forward_matches_reference(name):
    from fce_yolo_b200.weights import synth_images

    case = FORWARD_CASES[name]
    if case["size"] > 64 and case["yaml"].startswith("yolo11x"):
        pytest.skip("kept small")
    g = golden("fwd_" + name)
    cfg, scale = load_cfg(case)
    sd = _oracle_sd(cfg, scale, case["seed"])
    x = synth_images(case["img_seed"], case["batch"], case["size"], case["size"])
    (y, raw), ys = O.forward(cfg, scale, sd, x, keep_layers=True)
    sub = case.get("y_stride", 1)
    assert rel_max(y[:, :, ::sub], g["y"]) < 2e-5
    for k in g.files:
        if k.startswith("layer"):
            i = int(k[5:])
            assert ys[i].shape == g[k].shape
            assert rel_max(ys[i], g[k]) < 2e-5, k
    for i, r in enumerate(raw):
        s = g[f"raw{i}_sum"]
        assert abs(r.double().abs().sum().item() - s[1]) / s[1] < 1e-5


_MOD_FN = {
    "CoordAtt": lambda sd, xs, a: O.coord_att(sd, "mod", xs[0]),
    "CoordCrossAtt": lambda sd, xs, a: O.coord_cross_att(sd, "mod", xs[0], a[3]),
    "BiCoordCrossAtt": lambda sd, xs, a: O.bi_coord_cross_att(sd, "mod", xs[0], a[3]),
    "BiFPN_Concat": lambda sd, xs, a: O.bifpn_concat(sd, "mod", xs),
}


def module_case_io(case):
    """Synthetic weights + inputs of a module-level case (shared with the GPU tests)."""
    from fce_yolo_b200 import modules as M
    from fce_yolo_b200.weights import synth_images, synth_tensor

    mod = getattr(M, case["cls"])(*case["args"])
    M.fuse_module(mod)
    sd = {}
    for k, v in mod.state_dict().items():
        t = torch.from_numpy(synth_tensor(case["seed"], "mod." + k, v.shape))
        v.copy_(t)
        sd["mod." + k] = t
    xs = [synth_images(case["img_seed"] + j, case["batch"], h, w, c) * 4 - 2
          for j, (c, h, w) in enumerate(case["inputs"])]
    return mod, sd, xs


@pytest.mark.parametrize("name", list(MODULE_CASES))
def test_module_matches_reference(name):
    case = MODULE_CASES[name]
    _, sd, xs = module_case_io(case)
    with torch.no_grad():
        y = _MOD_FN[case["cls"]](sd, xs, case["args"])
    g = golden("mod_" + name)["y"]
    assert y.shape == g.shape
    assert rel_max(y, g) < 1e-5


@pytest.mark.parametrize("name", list(NMS_CASES))
def test_nms_matches_reference(name):
    case = NMS_CASES[name]
    p = case["make"]().numpy()
    dets, idxs = nms_oracle.non_max_suppression(p, **case["kw"])
    g = golden("nms_" + name)
    for b in range(p.shape[0]):
        # keep indices, class ids and the copied-through boxes/conf: all bit-exact
        assert_nms_equal(idxs[b], dets[b], g[f"idx{b}"], g[f"det{b}"], tie_perm=case.get("tie_perm", False))


# ---------------------------------------------------------------------------------------------------
# preprocessing: LetterBox + BGR->RGB, bit-exact against the reference (cv2) fixtures
# ---------------------------------------------------------------------------------------------------
from cases import LETTERBOX_CASES, letterbox_image  # noqa: E402


@pytest.mark.parametrize("name", list(LETTERBOX_CASES))
def test_letterbox_oracle_matches_reference(name):
    import numpy as np

    from oracle import letterbox_oracle as LB

    case = LETTERBOX_CASES[name]
    got = LB.letterbox(letterbox_image(case), case["new_shape"], stride=32, **case["kw"])
    ref = golden(name)["out"]
    assert got.shape == ref.shape and np.array_equal(got, ref)


def test_letterbox_host_tables_match_oracle():
    """The product's host-side tap tables / geometry (fce_yolo_b200/preprocess.py) against the oracle's restatement."""
    import numpy as np

    from fce_yolo_b200 import preprocess as P
    from oracle import letterbox_oracle as LB

    for src, dst in [(45, 96), (201, 64), (96, 96), (1080, 360), (375, 480), (7, 640), (640, 7)]:
        for clamp in (True, False):
            i0, i1, w0, w1 = LB.linear_coeffs(src, dst, clamp)
            t = P._taps(src, dst, clamp)
            assert np.array_equal(t[:, 0], i0) and np.array_equal(t[:, 1], i1)
            assert np.array_equal(t[:, 2], w0) and np.array_equal(t[:, 3], w1)
    for shape, kw in [((480, 640), {}), ((333, 517), dict(auto=True)), ((50, 70), dict(scaleup=False)),
                      ((1280, 960), dict(auto=True)), ((100, 80), dict(center=False))]:
        g = LB.letterbox_geometry(shape, (640, 640), stride=32, **kw)
        assert P.letterbox_geometry(shape, (640, 640), stride=32, **kw) == \
            (g["new_w"], g["new_h"], g["top"], g["left"], g["out_h"], g["out_w"])


from cases import SCALE_BOXES_CASES, scale_boxes_input  # noqa: E402


@pytest.mark.parametrize("name", list(SCALE_BOXES_CASES))
def test_scale_boxes_oracle_matches_reference(name):
    """ops.scale_boxes + clip_boxes of the live reference (fixtures) vs the numpy restatement: bit-exact fp32."""
    import numpy as np

    from fce_yolo_b200.predict import scale_meta
    from oracle import letterbox_oracle as LB

    case = SCALE_BOXES_CASES[name]
    got = LB.scale_boxes(case["img1"], scale_boxes_input(case), case["img0"])
    assert np.array_equal(got, golden(name)["out"])
    m = scale_meta(case["img1"], [case["img0"]])[0]
    gain = min(case["img1"][0] / case["img0"][0], case["img1"][1] / case["img0"][1])
    assert m[0] == np.float32(gain) and m[3] == case["img0"][1] and m[4] == case["img0"][0]


from cases import MATCH_CASES, match_inputs  # noqa: E402


@pytest.mark.parametrize("name", list(MATCH_CASES))
def test_match_oracle_matches_reference(name):
    """DetectionValidator._process_batch of the live reference (fixtures) vs the restatement: identical tp matrices."""
    import numpy as np

    from oracle import match_oracle as MO

    pred, pred_cls, gt, gt_cls = match_inputs(MATCH_CASES[name])
    got = MO.match_predictions(pred, pred_cls, gt, gt_cls)
    ref = golden(name)["tp"].reshape(got.shape)
    assert np.array_equal(got, ref)
