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


def test_fused_decode_matches_decode_kernel():
    """Predictor(fuse_decode=True) - Detect's last convs decode in their tcgen05 epilogue, no logit maps, no decode
    launch - against the unfused plan (fp32 logits + fce_detect_decode) on the same batch: class scores within 2e-6
    (approximate ex2 / rcp), boxes within 2e-2 pixel, and the same detections after NMS."""
    from fce_yolo_b200.predict import Predictor

    model = _model()
    B, S = 4, 640
    plain = Predictor(model, B, S, precision="bf16", conf=0.05, fuse_decode=False)
    fused = Predictor(model, B, S, precision="bf16", conf=0.05, fuse_decode=True)
    fns = [n.fn for n in fused.ex.plan.nodes]
    assert "fce_detect_decode" not in fns and fns.count("fce_conv2d_detect") == 6
    assert fused.launches_per_call == plain.launches_per_call - 1
    g = torch.Generator().manual_seed(11)
    x = torch.randint(0, 256, (B, S, S, 3), generator=g, dtype=torch.uint8).pin_memory()
    d0, c0 = [t.clone() for t in plain.infer(x)]
    d1, c1 = [t.clone() for t in fused.infer(x)]
    y0, y1 = plain.ex.outputs()[0], fused.ex.outputs()[0]
    assert (y0[:, 4:] - y1[:, 4:]).abs().max().item() <= 2e-6
    assert (y0[:, :4] - y1[:, :4]).abs().max().item() <= 2e-2
    assert torch.equal(c0, c1) and int(c0.sum()) > 0
    assert torch.allclose(d0, d1, atol=2e-2, rtol=0)


@pytest.mark.parametrize("overlap", [False, True])
def test_head_branches_match_single_stream(overlap):
    """Plan.HEAD_STREAMS: the six Detect conv chains captured as concurrent branches of the CUDA graph (event edges from
    Plan.dependencies()).  Same kernels on the same data in a different schedule: predictions and detections must be
    BIT-IDENTICAL to the single-stream graph, over a stream of different batches, with and without the NMS overlap."""
    from fce_yolo_b200.plan import Plan
    from fce_yolo_b200.predict import Predictor

    model = _model()
    B, S = 4, 640
    saved = Plan.HEAD_STREAMS
    preds = {}
    try:
        for hs in (False, True):
            Plan.HEAD_STREAMS = hs
            preds[hs] = Predictor(model, B, S, precision="bf16", conf=0.05, overlap_nms=overlap)
    finally:
        Plan.HEAD_STREAMS = saved
    assert max(n.stream for n in preds[True].ex.plan.nodes) >= 6 and max(n.stream for n in preds[False].ex.plan.nodes) == 0
    g = torch.Generator().manual_seed(5)
    total = 0
    for it in range(5):
        x = torch.randint(0, 256, (B, S, S, 3), generator=g, dtype=torch.uint8).pin_memory()
        d0, c0 = [t.clone() for t in preds[False].infer(x)]
        d1, c1 = [t.clone() for t in preds[True].infer(x)]
        assert torch.equal(c0, c1) and torch.equal(d0, d1), it
        assert torch.equal(preds[False].ex.outputs()[0], preds[True].ex.outputs()[0]), it
        total += int(c0.sum())
    assert total > 0


def test_bf16_convs_that_leave_the_tensor_cores_are_reported():
    """Executor.simt_bf16_convs lists the bf16 convs fce_conv2d would route to the CUDA-core kernel (n scale: the 8-channel
    bottleneck convs, Cout % 16 != 0); the m-scale graph has none, and strict_tc=True turns a non-empty list into an error."""
    from fce_yolo_b200.engine import Executor
    from fce_yolo_b200.plan import PlanError, compile_model
    from fce_yolo_b200.tasks import DetectionModel

    dev = torch.device("cuda:0")
    n_model = _model()
    ex = Executor(compile_model(n_model, 1, 64, 64, "bf16", dev), use_graph=False)
    assert ex.simt_bf16_convs and all(".m." in t for t in ex.simt_bf16_convs)
    with pytest.raises(PlanError):
        Executor(compile_model(n_model, 1, 64, 64, "bf16", dev), use_graph=False, strict_tc=True)
    m_model = DetectionModel("yolo11m-bifpn.yaml").fuse().eval()
    assert Executor(compile_model(m_model, 1, 64, 64, "bf16", dev), use_graph=False, strict_tc=True).simt_bf16_convs == []
