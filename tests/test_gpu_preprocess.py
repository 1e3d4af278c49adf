"""GPU preprocessing (fce_letterbox through the C ABI via LetterBoxGPU): bit-exact against the reference's own
LetterBox + BGR->RGB output (fixtures generated with the live reference / cv2) and against the oracle restatement on
full-size, mixed-size batches.  Byte work: the bar is equality."""
import numpy as np
import pytest
import torch

from cases import LETTERBOX_CASES, letterbox_image
from helpers import golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(LETTERBOX_CASES))
def test_letterbox_matches_reference_golden(name):
    from fce_yolo_b200.preprocess import LetterBoxGPU

    case = LETTERBOX_CASES[name]
    lb = LetterBoxGPU(case["new_shape"], stride=32, **case["kw"])
    out = lb([letterbox_image(case)])
    torch.cuda.synchronize()
    ref = golden(name)["out"]
    assert tuple(out.shape) == (1,) + ref.shape
    assert np.array_equal(out[0].cpu().numpy(), ref)


def test_mixed_size_batch_matches_oracle_640():
    """One launch, eight differently sized sources (camera, HD, portrait, tiny, already 640) -> [8, 640, 640, 3]."""
    from fce_yolo_b200.preprocess import LetterBoxGPU
    from oracle import letterbox_oracle as LB

    rng = np.random.default_rng(5)
    shapes = [(480, 640), (1080, 1920), (375, 500), (720, 1280), (1333, 800), (31, 57), (640, 640), (641, 639)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    lb = LetterBoxGPU(640)
    out = lb(imgs)
    again = lb(imgs[::-1])  # staging buffers are reused across calls
    torch.cuda.synchronize()
    for b, im in enumerate(imgs):
        ref = LB.letterbox(im, (640, 640))
        assert np.array_equal(out[b].cpu().numpy(), ref), shapes[b]
        assert np.array_equal(again[len(imgs) - 1 - b].cpu().numpy(), ref)


def test_auto_rect_and_output_buffer():
    from fce_yolo_b200.preprocess import LetterBoxGPU
    from oracle import letterbox_oracle as LB

    rng = np.random.default_rng(6)
    imgs = [rng.integers(0, 256, (720, 1280, 3), dtype=np.uint8) for _ in range(3)]
    lb = LetterBoxGPU(640, auto=True)
    buf = torch.zeros(3, 384, 640, 3, dtype=torch.uint8, device="cuda")
    out = lb(imgs, out=buf)
    torch.cuda.synchronize()
    assert out.data_ptr() == buf.data_ptr()
    for b, im in enumerate(imgs):
        assert np.array_equal(out[b].cpu().numpy(), LB.letterbox(im, (640, 640), auto=True))
    with pytest.raises(ValueError):
        lb([imgs[0], rng.integers(0, 256, (640, 640, 3), dtype=np.uint8)])  # different letterboxed sizes
    with pytest.raises(ValueError):
        lb([imgs[0].astype(np.float32)])


def test_preprocess_feeds_predictor():
    """Raw BGR frames -> LetterBoxGPU -> Predictor.run_device on the predictor's own input buffer == feeding the
    oracle-letterboxed batch through Predictor.infer (identical detections)."""
    from fce_yolo_b200.predict import Predictor
    from fce_yolo_b200.preprocess import LetterBoxGPU
    from fce_yolo_b200.tasks import DetectionModel
    from fce_yolo_b200.weights import load_synthetic
    from oracle import letterbox_oracle as LB

    model = DetectionModel("yolo11n-fce.yaml").fuse().eval()
    load_synthetic(model, 0)
    pred = Predictor(model, 2, 320, precision="bf16", conf=0.05)
    rng = np.random.default_rng(7)
    imgs = [rng.integers(0, 256, (240, 320, 3), dtype=np.uint8), rng.integers(0, 256, (400, 300, 3), dtype=np.uint8)]
    LetterBoxGPU(320)(imgs, out=pred.inp)
    det, keep, count = pred.run_device()
    det, count = det.clone(), count.clone()
    ref_in = torch.from_numpy(np.stack([LB.letterbox(im, (320, 320)) for im in imgs]))
    h_det, h_count = pred.infer(ref_in.pin_memory())
    assert torch.equal(count.cpu(), h_count) and torch.equal(det.cpu(), h_det)


def test_scale_boxes_bit_exact():
    """fce_scale_boxes vs the reference's ops.scale_boxes output (fixtures): same fp32 bits; rows >= count untouched."""
    import ctypes as C

    from cases import SCALE_BOXES_CASES, scale_boxes_input
    from fce_yolo_b200 import _lib as L
    from fce_yolo_b200.predict import scale_meta

    lib = L.load(check_device=True)
    names = list(SCALE_BOXES_CASES)
    B, max_det = len(names), 80
    det = torch.full((B, max_det, 6), 7.0)
    for b, n in enumerate(names):
        det[b, :64, :4] = torch.from_numpy(scale_boxes_input(SCALE_BOXES_CASES[n]))
    before = det.clone()
    d = det.cuda()
    count = torch.tensor([64, 64, 10, 0], dtype=torch.int32, device="cuda")
    meta = np.concatenate([scale_meta(SCALE_BOXES_CASES[n]["img1"], [SCALE_BOXES_CASES[n]["img0"]]) for n in names])
    m = torch.from_numpy(meta).cuda()
    st = lib.fce_scale_boxes(C.c_void_p(d.data_ptr()), C.c_void_p(count.data_ptr()), C.c_void_p(m.data_ptr()), B, max_det,
                             C.c_void_p(torch.cuda.current_stream().cuda_stream))
    L.check(st, "fce_scale_boxes")
    out = d.cpu()
    for b, n in enumerate(names):
        c = int(count[b])
        assert np.array_equal(out[b, :c, :4].numpy(), golden(n)["out"][:c])
        assert torch.equal(out[b, c:], before[b, c:]) and torch.equal(out[b, :, 4:], before[b, :, 4:])


def test_predict_raw_frames_end_to_end():
    """Predictor.predict(list of raw BGR frames) == oracle letterbox -> Predictor.infer -> oracle scale_boxes."""
    from fce_yolo_b200.predict import Predictor
    from fce_yolo_b200.tasks import DetectionModel
    from fce_yolo_b200.weights import load_synthetic
    from oracle import letterbox_oracle as LB

    model = DetectionModel("yolo11n-fce.yaml").fuse().eval()
    load_synthetic(model, 0)
    pred = Predictor(model, 3, 320, precision="bf16", conf=0.05)
    rng = np.random.default_rng(8)
    imgs = [rng.integers(0, 256, s + (3,), dtype=np.uint8) for s in [(240, 320), (400, 300)]]  # 2 frames, batch 3
    got = pred.predict(imgs)
    ref_in = torch.zeros(3, 320, 320, 3, dtype=torch.uint8)
    for b, im in enumerate(imgs):
        ref_in[b] = torch.from_numpy(LB.letterbox(im, (320, 320)))
    h_det, h_count = pred.infer(ref_in.pin_memory())
    assert len(got) == 2
    for b, im in enumerate(imgs):
        c = int(h_count[b])
        assert got[b].shape == (c, 6)
        ref = h_det[b, :c].clone().numpy()
        ref[:, :4] = LB.scale_boxes((320, 320), ref[:, :4], im.shape[:2])
        assert np.array_equal(got[b].numpy(), ref)
