"""N>1 host logic on CPU: world_size-2 gloo processes exercise the sharding and the (only) collectives of the
multi-GPU path - the fixed-shape all-gather of padded detections and the gather of validation statistics that
replace the reference's pickled dist.gather_object calls (ultralytics/models/yolo/detect/val.py:222-242)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fce_yolo_b200.runner import DetectionGather, gather_detections, gather_stats_to_rank0, shard_range


def test_shard_range_matches_contiguous_sampler():
    # reference data/build.py:172-189: whole batches dealt in order, first ranks take the remainder
    for n, b, w in [(100, 8, 2), (64, 64, 8), (1000, 32, 4), (7, 4, 3), (0, 4, 2)]:
        spans = [shard_range(n, b, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (s0, e0), (s1, e1) in zip(spans, spans[1:]):
            assert e0 == s1 and s0 <= e0
        sizes = [-(-(e - s) // b) for s, e in spans]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
        assert all(s % b == 0 for s, _ in spans if s < n)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B, max_det = 3, 5
        g = torch.Generator().manual_seed(100 + rank)
        det = torch.rand(B, max_det, 6, generator=g)
        count = torch.tensor([rank + 1, 0, max_det], dtype=torch.int32)
        all_det, all_cnt = gather_detections(det, count)
        ok = all_det.shape == (world * B, max_det, 6) and all_cnt.shape == (world * B,)
        for r in range(world):
            gr = torch.Generator().manual_seed(100 + r)
            ok &= torch.equal(all_det[r * B:(r + 1) * B], torch.rand(B, max_det, 6, generator=gr))
            ok &= all_cnt[r * B:(r + 1) * B].tolist() == [r + 1, 0, max_det]
        # the zero-copy per-step gatherer on a stand-in executor: det and count adjacent in one byte buffer
        nd = B * max_det * 24
        packed = torch.cat([det.view(-1).view(torch.uint8), count.view(torch.uint8)])

        class _Plan:
            outputs = {"detcount": "dc", "max_det": max_det}

        _Plan.B = B

        class _Ex:
            plan = _Plan

            @staticmethod
            def bytes(v):
                return packed

        dg = DetectionGather(_Ex)
        d4, c2 = dg.gather()
        ok &= d4.shape == (world, B, max_det, 6) and c2.shape == (world, B) and dg.send.numel() == nd + 4 * B
        ok &= torch.equal(d4.reshape(world * B, max_det, 6), all_det) and torch.equal(c2.reshape(-1), all_cnt)
        # ragged validation statistics (different n per rank, one rank empty-ish)
        n = 4 if rank == 0 else 1
        stats = {"tp": torch.full((n, 10), bool(rank)), "conf": torch.arange(n, dtype=torch.float32) + 10 * rank,
                 "pred_cls": torch.full((n,), rank, dtype=torch.float32)}
        merged = gather_stats_to_rank0(stats)
        if rank == 0:
            ok &= merged["tp"].shape == (5, 10) and merged["conf"].tolist() == [0, 1, 2, 3, 10]
            ok &= merged["pred_cls"].tolist() == [0, 0, 0, 0, 1] and merged["tp"][4].all().item() \
                and not merged["tp"][:4].any().item()
        else:
            ok &= merged is None
        # a rank with NO statistics at all (more ranks than batches) still joins the collectives
        empty = {"tp": torch.zeros((0, 10), dtype=torch.uint8), "conf": torch.zeros(0)} if rank == 1 else \
            {"tp": torch.ones((2, 10), dtype=torch.uint8), "conf": torch.tensor([0.5, 0.25])}
        m2 = gather_stats_to_rank0(empty)
        if rank == 0:
            ok &= m2["tp"].shape == (2, 10) and m2["conf"].tolist() == [0.5, 0.25]
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == {0: True, 1: True}


def test_single_process_is_identity():
    det, cnt = torch.rand(2, 3, 6), torch.tensor([1, 2], dtype=torch.int32)
    a, b = gather_detections(det, cnt)
    assert a is det and b is cnt
