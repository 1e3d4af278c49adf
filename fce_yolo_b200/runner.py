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
    rank order on every rank.  ONE fixed-shape collective (detections and counts travel in one packed byte buffer),
    issued on the caller's stream.  Functional form (allocates the packed buffers per call): the per-step path of a
    Predictor is :class:`DetectionGather`, which sends straight out of the buffer the NMS kernel wrote."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return det, count
    world = dist.get_world_size(group)
    B = det.shape[0]
    nd = det.numel() * 4
    send = torch.cat([det.contiguous().view(-1).view(torch.uint8), count.to(torch.int32).contiguous().view(torch.uint8)])
    recv = send.new_empty((world, send.numel()))
    dist.all_gather_into_tensor(recv.view(-1), send, group=group)
    out_det = recv[:, :nd].contiguous().view(torch.float32).view((world * B,) + tuple(det.shape[1:]))
    out_cnt = recv[:, nd:].contiguous().view(torch.int32).view(world * B)
    return out_det, out_cnt


class DetectionGather:
    """The one collective of the predict path, without a copy on either side: the plan places the NMS outputs
    ``det [B, max_det, 6] fp32`` and ``count [B] int32`` back to back in ONE arena buffer (plan.Plan.nms), so the NMS
    kernel writes straight into the send buffer; the receive buffer is allocated once.  ``gather()`` issues a single
    ``all_gather_into_tensor`` on the current stream and returns views ``det [W, B, max_det, 6]``, ``count [W, B]`` into
    the receive buffer (rank-major = dataset order with runner.shard_range).  Replaces the reference's two pickled
    ``dist.gather_object`` calls (ultralytics/models/yolo/detect/val.py:222-242)."""

    def __init__(self, executor, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        o = executor.plan.outputs
        if "detcount" not in o:
            raise RuntimeError("DetectionGather needs a plan that ends in NMS")
        self.send = executor.bytes(o["detcount"])
        self.B, self.max_det = executor.plan.B, o["max_det"]
        self.nd = self.B * self.max_det * 24
        self.recv = torch.empty((self.world, self.send.numel()), dtype=torch.uint8, device=self.send.device)

    def views(self, buf2d):
        W = buf2d.shape[0]
        det = buf2d[:, :self.nd].view(torch.float32).view(W, self.B, self.max_det, 6)
        cnt = buf2d[:, self.nd:self.nd + 4 * self.B].view(torch.int32)
        return det, cnt

    def gather(self):
        if self.world == 1:
            return self.views(self.send.view(1, -1))
        dist.all_gather_into_tensor(self.recv.view(-1), self.send, group=self.group)
        return self.views(self.recv)


def gather_stats_to_rank0(stats: dict, group=None):
    """Validation statistics (tp [n,10] bool, conf [n], pred_cls [n], ...; reference metrics.py:1118) gathered to
    rank 0 as padded tensors instead of pickles: ONE collective for the row counts of every key, one host
    synchronisation, ONE collective for all payloads (packed into a single byte buffer).  Returns the merged dict on
    rank 0, None elsewhere."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return stats
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    keys = sorted(stats)
    dev = stats[keys[0]].device
    n_local = torch.tensor([stats[k].shape[0] for k in keys], dtype=torch.int64, device=dev)
    n_all = torch.empty((world, len(keys)), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(n_all.view(-1), n_local, group=group)
    n_all = n_all.tolist()  # the one host sync
    import math

    row_bytes = [math.prod(stats[k].shape[1:]) * stats[k].element_size() for k in keys]  # (an empty tensor has rows too)
    n_max = [max(n_all[r][i] for r in range(world)) for i in range(len(keys))]
    seg = [-(-(n_max[i] * row_bytes[i]) // 16) * 16 for i in range(len(keys))]  # 16-byte aligned segments
    send = torch.zeros(max(sum(seg), 16), dtype=torch.uint8, device=dev)
    off = 0
    for i, k in enumerate(keys):
        t = stats[k].contiguous()
        nb = t.numel() * t.element_size()
        if nb:
            send[off:off + nb] = t.view(-1).view(torch.uint8)
        off += seg[i]
    recv = torch.empty((world, send.numel()), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv.view(-1), send, group=group)
    if rank != 0:
        return None
    merged = {}
    off = 0
    for i, k in enumerate(keys):
        t = stats[k]
        parts = []
        for r in range(world):
            nb = n_all[r][i] * row_bytes[i]
            parts.append(recv[r, off:off + nb].contiguous().view(t.dtype).view((n_all[r][i],) + tuple(t.shape[1:])))
        merged[k] = torch.cat(parts)
        off += seg[i]
    return merged
