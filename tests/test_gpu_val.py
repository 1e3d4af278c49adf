"""Validation matching on the GPU (fce_match_predictions through the C ABI): identical tp matrices to the live
reference's DetectionValidator._process_batch (fixtures) and to the oracle on a full batch; ValStats bookkeeping."""
import numpy as np
import pytest
import torch

from cases import MATCH_CASES, match_inputs
from helpers import golden

pytestmark = pytest.mark.gpu


def _pack(names, max_det=300):
    B = len(names)
    det = torch.zeros(B, max_det, 6)
    count = torch.zeros(B, dtype=torch.int32)
    gts, gcs, offs = [], [], [0]
    for b, n in enumerate(names):
        pred, pred_cls, gt, gt_cls = match_inputs(MATCH_CASES[n])
        k = len(pred_cls)
        det[b, :k, :4] = torch.from_numpy(pred)
        det[b, :k, 4] = torch.linspace(0.9, 0.1, k) if k else 0
        det[b, :k, 5] = torch.from_numpy(pred_cls)
        det[b, k:, :4] = 777.0  # garbage beyond count must be ignored
        count[b] = k
        gts.append(torch.from_numpy(gt).reshape(-1, 4))
        gcs.append(torch.from_numpy(gt_cls))
        offs.append(offs[-1] + len(gt_cls))
    return det, count, torch.cat(gts), torch.cat(gcs), offs


def test_match_batch_equals_reference_fixtures():
    from fce_yolo_b200.val import match_predictions

    names = list(MATCH_CASES)
    det, count, gb, gc, offs = _pack(names)
    tp = match_predictions(det.cuda(), count.cuda(), gb, gc, offs).cpu().numpy()
    for b, n in enumerate(names):
        k = int(count[b])
        ref = golden(n)["tp"].reshape(k, 10)
        assert np.array_equal(tp[b, :k], ref), n
        assert not tp[b, k:].any()


def test_match_large_batch_vs_oracle_and_stats():
    from fce_yolo_b200.val import ValStats, match_predictions
    from oracle import match_oracle as MO

    rng = np.random.default_rng(3)
    names = [list(MATCH_CASES)[i] for i in rng.integers(0, 3, 48)]
    det, count, gb, gc, offs = _pack(names)
    tp = match_predictions(det.cuda(), count.cuda(), gb, gc, offs).cpu().numpy()
    total = 0
    for b in range(len(names)):
        k = int(count[b])
        ref = MO.match_predictions(det[b, :k, :4].numpy(), det[b, :k, 5].numpy(), gb[offs[b]:offs[b + 1]].numpy(),
                                   gc[offs[b]:offs[b + 1]].numpy())
        assert np.array_equal(tp[b, :k], ref)
        total += k
    vs = ValStats()
    vs.update(det.cuda(), count.cuda(), gb, gc, offs)
    out = vs.result()
    assert out["tp"].shape == (total, 10) and out["conf"].shape == (total,) and out["target_cls"].shape == (offs[-1],)
    assert int(out["tp"].sum()) == int(tp.sum())


def test_match_rejects_cpu_and_bad_offsets():
    from fce_yolo_b200.val import match_predictions

    det, count, gb, gc, offs = _pack(["match_one"])
    with pytest.raises(RuntimeError):
        match_predictions(det, count, gb, gc, offs)
    with pytest.raises(ValueError):
        match_predictions(det.cuda(), count.cuda(), gb, gc, [0, 5])
