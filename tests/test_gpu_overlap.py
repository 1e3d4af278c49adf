"""NMS-overlap mode of the Predictor (NMS of call i on a side stream underneath the forward of call i+1): detections
must be bit-identical to the in-line mode for a stream of DIFFERENT batches - i.e. the forward of batch i+1 must
not disturb the NMS of batch i (the two share the decoded prediction buffer) - through run_device / pipeline /
infer."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model():
    from fce_yolo_b200.tasks import DetectionModel
    from fce_yolo_b200.weights import load_synthetic

    m = DetectionModel("yolo11n-fce.yaml").fuse().eval()
    load_synthetic(m, 0)
    return m


def test_overlap_matches_inline_over_a_stream_of_batches():
    from fce_yolo_b200.predict import Predictor

    model = _model()
    B, S = 8, 320
    inline = Predictor(model, B, S, precision="bf16", conf=0.05)
    over = Predictor(model, B, S, precision="bf16", conf=0.05, overlap_nms=True)
    assert over.overlap and not inline.overlap
    g = torch.Generator().manual_seed(3)
    batches = [torch.randint(0, 256, (B, S, S, 3), generator=g, dtype=torch.uint8).pin_memory() for _ in range(6)]
    ref = []
    for x in batches:
        d, c = inline.infer(x)
        ref.append((d.clone(), c.clone()))
    # pipeline(): uploads, graphs and downloads of consecutive batches overlap
    got = [(d.clone(), c.clone()) for d, c in over.pipeline(batches)]
    assert len(got) == len(ref)
    for (d, c), (rd, rc) in zip(got, ref):
        assert torch.equal(c, rc) and torch.equal(d, rd)
    assert sum(int(c.sum()) for _, c in ref) > 0  # the comparison is not vacuous
    # run_device() back to back, outputs read after join()
    for x, (rd, rc) in zip(batches, ref):
        over.inp.copy_(x)
        det, keep, count = over.run_device()
        over.join()
        assert torch.equal(count.cpu(), rc) and torch.equal(det.cpu(), rd)
    # infer()
    d, c = over.infer(batches[2])
    assert torch.equal(c, ref[2][1]) and torch.equal(d, ref[2][0])


def test_overlap_needs_graphs_and_nms_tail():
    from fce_yolo_b200.predict import Predictor

    p = Predictor(_model(), 1, 64, overlap_nms=True, use_graph=False)
    assert not p.overlap  # silently in-line without graphs: same results, no side stream
