"""N > 1 on real GPUs (skipped on a single-GPU box; the CPU twin is tests/test_runner_gloo.py): validation statistics
and the per-step detection gather over NCCL, bit-identical to a single-process run (tests/dist_val_check.py)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
@pytest.mark.parametrize("images", [56, 8])  # 8 images = one batch: every rank but the first has an EMPTY shard
def test_val_stats_and_detection_gather_over_nccl(images):
    n = min(torch.cuda.device_count(), 8)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(29541 + images), os.path.join(ROOT, "tests", "dist_val_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT,
                       env=dict(os.environ, FCE_DIST_VAL_IMAGES=str(images)))
    line = next((ln for ln in reversed(r.stdout.splitlines()) if ln.startswith("{")), None)
    assert r.returncode == 0 and line, (r.stdout[-1500:], r.stderr[-1500:])
    out = json.loads(line)
    assert out["dist_val"] == "ok" and out["world"] == n and out["n_pred"] >= 300 * images * 0.9, out


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least two GPUs")
def test_two_devices_in_one_process():
    """One process driving two GPUs: the > 48 KB dynamic shared-memory opt-in (cudaFuncSetAttribute) is a per-device
    attribute - every kernel family must repeat it on the second device (the tcgen05 convs ask for up to 227 KB).  Same
    model, same batch on cuda:0 and cuda:1: identical detections."""
    import detection_parity as DP
    from fce_yolo_b200.predict import Predictor

    case = dict(DP.CONFIGS["cfg1_s_coordatt"], size=320, batch=4)
    _, model, _ = DP.build(case)
    x = DP.u8_batch(9, 4, 320).pin_memory()
    outs = []
    for dev in ("cuda:0", "cuda:1"):
        p = Predictor(model, 4, 320, precision="bf16", device=torch.device(dev), conf=0.25, iou=0.7, overlap_nms=True)
        det, cnt = p.infer(x)
        outs.append((det.clone(), cnt.clone(), p.keep.cpu().clone()))
    assert int(outs[0][1].sum()) > 100
    for a, b in zip(*outs):
        assert torch.equal(a, b)
