"""Multi-GPU inference: one process per GPU, full replica each, batch sharded by image.

Images are independent (BN folded, attention per image, NMS per image - SURVEY 8e), so the only
collective is the gather of results: fixed-shape padded detections ``[B_local, max_det, 6]`` + counts, which
replaces the reference's two pickled ``dist.gather_object`` calls (ultralytics/models/yolo/detect/val.py:222-242).
Works over NCCL (NVLink/NVSwitch) on GPUs and over gloo on CPU tensors (used by the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_items: int, batch: int, rank: int, world: int):
    """Contiguous, batch-aligned slice of a dataset for `rank` - mirrors the reference's
    ContiguousDistributedSampler._get_rank_indices (ultralytics/data/build.py:172-189): whole batches are
    dealt to ranks in order, the first ranks take the remainder batches, order after gather = dataset order."""
    n_batches = (n_items + batch - 1) // batch
    per, extra = divmod(n_batches, world)
    start_b = rank * per + min(rank, extra)
    end_b = start_b + per + (1 if rank < extra else 0)
    return min(start_b * batch, n_items), min(end_b * batch, n_items)


def gather_detections(det: torch.Tensor, count: torch.Tensor, group=None):
    """All-gather padded detections.  det [B,max_det,6] fp32, count [B] int32 -> ([W*B,max_det,6], [W*B]) in
    rank order on every rank.  One fixed-shape collective per tensor, issued on the caller's stream."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return det, count
    world = dist.get_world_size(group)
    out_det = det.new_empty((world * det.shape[0],) + tuple(det.shape[1:]))
    out_cnt = count.new_empty((world * count.shape[0],))
    dist.all_gather_into_tensor(out_det, det.contiguous(), group=group)
    dist.all_gather_into_tensor(out_cnt, count.contiguous(), group=group)
    return out_det, out_cnt


def gather_stats_to_rank0(stats: dict, group=None):
    """Validation statistics (tp [n,10] bool, conf [n], pred_cls [n], ...; reference metrics.py:1118) gathered to
    rank 0 as padded tensors instead of pickles.  Returns the merged dict on rank 0, None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return stats
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    merged = {}
    for k in sorted(stats):
        t = stats[k]
        n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        ns = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(ns, n, group=group)
        m = int(max(int(x) for x in ns))
        pad = t.new_zeros((m,) + tuple(t.shape[1:]))
        pad[: t.shape[0]] = t
        outs = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(outs, pad, group=group)
        if rank == 0:
            merged[k] = torch.cat([o[: int(c)] for o, c in zip(outs, ns)])
    return merged if rank == 0 else None
