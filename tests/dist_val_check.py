#!/usr/bin/env python
"""Multi-GPU validation statistics over NCCL (BASELINE configs[4]: "NCCL gather of detections for val
statistics").  Launch with one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/dist_val_check.py

Every rank validates its contiguous shard of a synthetic dataset (runner.shard_range = the reference's
ContiguousDistributedSampler, data/build.py:172-189) with the val settings of the reference (conf 0.001, iou 0.7,
multi-label, val.py:105-126), matches predictions to labels on its GPU (fce_match_predictions) and the statistics are
gathered to rank 0 with runner.gather_stats_to_rank0 (two collectives in total) - what replaces the reference's pickled
dist.gather_object (val.py:222-242).  Rank 0 then validates the WHOLE dataset alone and requires the gathered statistics to
be bit-identical (same rows in the same order), and checks runner.DetectionGather (the per-step collective of predict)
against its own recomputation of every rank's batch.  Prints one JSON line; exit code 0 = parity."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import datetime

    dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    from fce_yolo_b200.metrics import detection_metrics
    from fce_yolo_b200.predict import Predictor
    from fce_yolo_b200.runner import DetectionGather, shard_range
    from fce_yolo_b200.tasks import DetectionModel
    from fce_yolo_b200.val import ValStats
    from fce_yolo_b200.weights import load_synthetic, synth_images

    # 7 batches by default: ranks get unequal shares; FCE_DIST_VAL_IMAGES=8 leaves every rank but the first WITHOUT a batch
    N_IMG, B, S = int(os.environ.get("FCE_DIST_VAL_IMAGES", 56)), 8, 320
    model = DetectionModel("yolo11n-fce.yaml").fuse().eval()
    load_synthetic(model, 0)
    pred = Predictor(model, B, S, precision="bf16", device=dev, conf=0.001, iou=0.7, max_det=300, multi_label=True,
                     input_u8=True, overlap_nms=True)
    images = (synth_images(99, N_IMG, S, S) * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()

    def run_batch(i0):
        """Detections of images [i0, i0 + B) (short last batch padded with zeros) + labels derived from them."""
        n = min(B, N_IMG - i0)
        x = torch.zeros(B, S, S, 3, dtype=torch.uint8)
        x[:n] = images[i0:i0 + n]
        pred.inp.copy_(x.to(dev))
        det, keep, count = pred.run_device()
        pred.join()
        torch.cuda.synchronize()
        det, count = det[:n].clone(), count[:n].clone()
        gts, gcs, offs = [], [], [0]
        for b in range(n):  # labels: every 7th detection, shifted by 2 px (IoU ~0.9 with itself, misses the higher thresholds)
            k = int(count[b])
            sel = det[b, :k:7]
            gts.append(sel[:, :4] + 2.0)
            gcs.append(sel[:, 5])
            offs.append(offs[-1] + len(sel))
        return det, count, torch.cat(gts), torch.cat(gcs), offs

    def validate(lo, hi):
        vs = ValStats()
        for i0 in range(lo, hi, B):
            det, count, gb, gc, offs = run_batch(i0)
            vs.update(det, count, gb, gc, offs)
        return vs

    lo, hi = shard_range(N_IMG, B, rank, world)
    merged = validate(lo, hi).result()  # collective: every rank calls it
    # the per-step collective of predict: every rank's first batch, gathered in one all_gather out of the NMS buffer
    run_batch(lo if lo < hi else 0)
    g = DetectionGather(pred.ex)
    with torch.cuda.stream(pred._out_stream()):
        det_all, cnt_all = g.gather()
    torch.cuda.synchronize()
    det_all, cnt_all = det_all.clone(), cnt_all.clone()
    ok, info = True, {}
    if rank == 0:
        full = validate(0, N_IMG)
        parts = {k: (torch.cat(v) if v else None) for k, v in full.parts.items()}
        for k in ("tp", "conf", "pred_cls", "target_cls", "target_img"):
            a, b = merged[k], parts[k]
            same = a.shape == b.shape and torch.equal(a.to(b.dtype), b)
            ok &= bool(same)
            info[k] = [list(a.shape), bool(same)]
        m = detection_metrics(merged["tp"], merged["conf"], merged["pred_cls"], merged["target_cls"])
        info["metrics"] = {k: round(float(m[k]), 6) for k in ("mp", "mr", "map50", "map") if k in m}
        info["n_pred"] = int(merged["conf"].shape[0])
        for r in range(world):
            rlo, rhi = shard_range(N_IMG, B, r, world)
            d, c, *_ = run_batch(rlo if rlo < rhi else 0)
            n = d.shape[0]
            same = torch.equal(cnt_all[r, :n], c) and all(
                torch.equal(det_all[r, b, :int(c[b])], d[b, :int(c[b])]) for b in range(n))  # rows beyond count are padding
            ok &= bool(same)
            info[f"gather_rank{r}"] = bool(same)
        print(json.dumps({"dist_val": "ok" if ok else "MISMATCH", "world": world, "backend": "nccl", **info}))
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
