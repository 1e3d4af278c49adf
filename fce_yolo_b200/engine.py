"""Plan executor: arena allocation, stream-ordered launches, CUDA-graph replay, module/model entry points.

A compiled Plan (plan.py) is bound to one arena (a single torch.uint8 allocation, so PyTorch's caching
allocator still owns the memory), laid out by a liveness-driven first-fit allocator: a buffer's bytes are
recycled as soon as its last consumer has been issued.  At 180 GB of HBM the point is not fitting - it is
keeping the working set small enough that consecutive layers hit the 126 MB L2 more often.

Launches go to ``torch.cuda.current_stream()``; after one eager pass the whole plan is captured into a
CUDA graph (350-590 eager ATen launches in the reference become one graph launch).
"""
from __future__ import annotations

import ctypes as C
import gc
import os
import weakref

import torch

from . import _lib as L
from .plan import Buf, Plan, PlanError, View, compile_model, compile_module, DT_SIZE, TORCH_DT

ALIGN = 1024
NVTX = os.environ.get("FCE_NVTX") == "1"
_DEFAULT_PRECISION = "bf16"


def set_default_precision(p: str):
    global _DEFAULT_PRECISION
    if p not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    _DEFAULT_PRECISION = p


def _align(n):
    return (n + ALIGN - 1) // ALIGN * ALIGN


def assign_offsets(plan: Plan, reuse: bool = True) -> int:
    """First-fit arena layout by liveness.  Returns the arena size in bytes."""
    bufs = [b for b in plan.bufs if b.first >= 0 or b.persistent]
    top = 0
    if not reuse:
        for b in bufs:
            b.offset = top
            top += _align(b.nbytes)
        return top
    for b in bufs:  # persistent buffers first, never recycled
        if b.persistent:
            b.offset = top
            top += _align(b.nbytes)
    free: list[list[int]] = []  # [offset, size], sorted by offset
    by_first: dict[int, list[Buf]] = {}
    by_last: dict[int, list[Buf]] = {}
    for b in bufs:
        if b.persistent:
            continue
        by_first.setdefault(b.first, []).append(b)
        by_last.setdefault(b.last, []).append(b)
    for idx in range(len(plan.nodes)):
        for b in by_first.get(idx, []):
            need = _align(b.nbytes)
            best = None
            for k, (off, sz) in enumerate(free):
                if sz >= need and (best is None or sz < free[best][1]):
                    best = k
            if best is None:
                if free and free[-1][0] + free[-1][1] == top:  # grow the trailing hole
                    b.offset = free[-1][0]
                    top = b.offset + need
                    free.pop()
                else:
                    b.offset = top
                    top += need
            else:
                off, sz = free[best]
                b.offset = off
                if sz == need:
                    free.pop(best)
                else:
                    free[best] = [off + need, sz - need]
        for b in by_last.get(idx, []):
            free.append([b.offset, _align(b.nbytes)])
            free.sort()
            merged = []
            for off, sz in free:
                if merged and merged[-1][0] + merged[-1][1] == off:
                    merged[-1][1] += sz
                else:
                    merged.append([off, sz])
            free = merged
    return top


class Arena:
    """One allocation holding every plan buffer; hands out torch views over it (device-agnostic)."""

    def __init__(self, plan: Plan, reuse_memory: bool = True):
        self.plan = plan
        self.device = plan.device
        size = assign_offsets(plan, reuse=reuse_memory)
        self.mem = torch.empty(max(size, ALIGN), dtype=torch.uint8, device=self.device)
        self.nbytes = size
        self.base = self.mem.data_ptr()

    def _flat(self, b: Buf) -> torch.Tensor:
        return self.mem[b.offset:b.offset + b.nbytes].view(TORCH_DT[b.dtype])

    def tensor(self, v: View) -> torch.Tensor:
        """Torch view [B, H, W, C_view] (strided, zero-copy) over a plan view."""
        b = v.buf
        t = self._flat(b).view(b.B * b.H * b.W, b.C)[v.row0:v.row0 + v.B * v.H * v.W, v.c0:v.c0 + v.C]
        return t.unflatten(0, (v.B, v.H, v.W))

    def nchw(self, v: View) -> torch.Tensor:
        """Logical NCHW tensor (channels_last strides) - what the reference API hands around."""
        return self.tensor(v).permute(0, 3, 1, 2)

    def input_tensor(self, i: int = 0) -> torch.Tensor:
        b = self.plan.inputs[i].buf
        return self._flat(b).view(b.B, b.H, b.W, b.C)  # for the NCHW image buffer this *is* [B, C, H, W]

    def bytes(self, v: View) -> torch.Tensor:
        """Flat uint8 view over a byte-buffer view (plan.raw_buf and channel slices of it)."""
        b = v.buf
        if b.dtype != L.U8 or v.row0:
            raise ValueError("bytes() is for untyped byte buffers")
        return self.mem[b.offset + v.c0:b.offset + v.c0 + v.C]

    def detections(self):
        """(det [B,max_det,6] fp32, keep [B,max_det] int64, count [B] int32) views, if the plan ends in NMS."""
        o = self.plan.outputs
        B, md = self.plan.B, o["max_det"]
        det = self.bytes(o["det"]).view(torch.float32).view(B, md, 6)
        keep = self.bytes(o["keep"]).view(torch.int64).view(B, md)
        count = self.bytes(o["count"]).view(torch.int32).view(B)
        return det, keep, count

    def outputs(self):
        o = self.plan.outputs
        yv = o["y"]
        if "raw" in o:  # Detect: y is stored as [B, 4+nc, A]
            b = yv.buf
            return self._flat(b).view(b.B, b.W, b.C), [self.nchw(r) for r in o["raw"]]
        return self.nchw(yv)


class Executor(Arena):
    """Binds a Plan to device memory and runs it through the C ABI."""

    def __init__(self, plan: Plan, reuse_memory: bool = True, use_graph: bool = True, strict_tc: bool = False):
        self.lib = L.load(check_device=True)
        super().__init__(plan, reuse_memory)
        base = self.base
        if base % 256:
            raise RuntimeError("arena base is not 256-byte aligned")
        self._calls = []
        for n in plan.nodes:
            fn = getattr(self.lib, n.fn)
            args = [C.byref(n.desc)]
            for p in n.ptrs:
                if isinstance(p, View):
                    args.append(C.c_void_p(base + p.buf.offset + p.byte_offset()))
                elif isinstance(p, torch.Tensor):
                    args.append(C.c_void_p(p.data_ptr()))
                elif isinstance(p, C.Structure):  # a second descriptor (kept alive by the node)
                    args.append(C.byref(p))
                elif p is None:
                    args.append(C.c_void_p(0))
                else:
                    args.append(C.c_size_t(int(p)))
            self._calls.append((fn, args, n))
        # Which bf16 convs leave the tensor cores?  fce_conv2d falls through to the CUDA-core kernel when the tcgen05 path
        # rejects a shape (channel counts that are not multiples of 16, unaligned views ...): that must be VISIBLE, not
        # found in a profile.  (fp32-mode plans and the fp32 pooled strips are CUDA-core work by design.)
        self.simt_bf16_convs = []
        for fn, args, n in self._calls:
            d = n.desc
            if n.fn == "fce_conv2d" and d.in_dtype == L.BF16 and d.w_dtype == L.BF16 and d.impl == 0:
                if self.lib.fce_conv2d_route(*args[:3], args[4], args[5]) == 0:
                    self.simt_bf16_convs.append(n.tag)
        if strict_tc and self.simt_bf16_convs:
            raise PlanError(f"bf16 convolutions off the tensor cores: {self.simt_bf16_convs}")
        self.use_graph = use_graph
        self._side_streams = {}  # branch id -> torch stream (plans with Node.stream > 0, graph mode only)
        self.graph = None
        self._warm = False
        self.launches_per_run = len(self._calls)

    # -- execution ----------------------------------------------------------------------------------
    def _launch_all(self, stream_ptr):
        s = C.c_void_p(stream_ptr)
        if NVTX:  # FCE_NVTX=1: one NVTX range per plan node (entry point + layer tag) around eager launches, so that
            # a timeline / ncu --nvtx capture maps kernels back to reference layers
            for fn, args, n in self._calls:
                torch.cuda.nvtx.range_push(f"{n.fn}:{n.tag}")
                st = fn(*args, s)
                torch.cuda.nvtx.range_pop()
                if st != 0:
                    L.check(st, f"{n.fn} [{n.tag}]")
            return
        for fn, args, n in self._calls:
            st = fn(*args, s)
            if st != 0:
                L.check(st, f"{n.fn} [{n.tag}]")

    def _launch_branched(self, lo: int, hi: int):
        """Launches nodes [lo, hi) with the plan's side branches (Node.stream > 0) on side streams, ordered by events
        derived from Plan.dependencies().  Called under CUDA-graph capture, where the events become graph edges: e.g.
        the Detect chains of the fine levels run concurrently with the small-map layers that finish the neck.  Every
        side stream is joined back into the current stream before returning."""
        main = torch.cuda.current_stream(self.device)
        deps = self.plan.dependencies(lo, hi)
        stream_of = {i: self.plan.nodes[i].stream for i in range(lo, hi)}
        if not any(stream_of.values()):
            sp = C.c_void_p(main.cuda_stream)
            for fn, args, n in self._calls[lo:hi]:
                st = fn(*args, sp)
                if st != 0:
                    L.check(st, f"{n.fn} [{n.tag}]")
            return
        # cross-stream edges only (same-stream order is implicit); per foreign stream the LATEST node implies the rest
        cross_of = {}
        for i in range(lo, hi):
            latest = {}
            for j in deps[i]:
                if stream_of[j] != stream_of[i]:
                    latest[stream_of[j]] = max(latest.get(stream_of[j], -1), j)
            cross_of[i] = sorted(latest.values())
        needed = {j for i in range(lo, hi) for j in cross_of[i]}
        side, events, used = self._side_streams, {}, set()
        for i in range(lo, hi):
            fn, args, n = self._calls[i]
            k = stream_of[i]
            if k and k not in side:
                side[k] = torch.cuda.Stream(self.device)
            st_obj = main if k == 0 else side[k]
            cross = cross_of[i]
            for j in cross:
                st_obj.wait_event(events[j])
            if k and k not in used:
                used.add(k)
                if not cross:  # a branch must descend from the capturing stream
                    st_obj.wait_stream(main)
            rc = fn(*args, C.c_void_p(st_obj.cuda_stream))
            if rc != 0:
                L.check(rc, f"{n.fn} [{n.tag}]")
            if i in needed:
                ev = torch.cuda.Event()
                ev.record(st_obj)
                events[i] = ev
        for k in used:
            main.wait_stream(side[k])

    def _capture_stream(self):
        """Capture stream on THIS executor's device.  torch.cuda.graph() otherwise re-uses one class-level stream created
        on whatever device was current the first time any graph was captured in the process: on a second GPU the plan's
        launches would then miss the capture (invalid resource handle on the first side-branch launch)."""
        if getattr(self, "_cap_stream", None) is None:
            self._cap_stream = torch.cuda.Stream(self.device)
        return self._cap_stream

    def _capture(self, lo: int, hi: int):
        """One CUDA graph of the nodes [lo, hi).  Python's cycle collector must not run inside the capture window: it may
        finalize DEAD executors of the same process (they sit in reference cycles), and destroying their CUDA graphs is a call
        CUDA forbids while a capture is open - the destructor only warns ("operation not permitted when stream is capturing
        (function reset)"), but the open capture is invalidated and its next launch fails with "operation failed due to a
        previous error during capture" (seen in a 650-test process; torch.cuda.graph no longer collects before capturing).
        So: collect now, outside the window, and keep the collector off until the capture has ended."""
        g = torch.cuda.CUDAGraph()
        gc.collect()
        gc_was_on = gc.isenabled()
        gc.disable()
        try:
            with torch.cuda.graph(g, stream=self._capture_stream()):
                self._launch_branched(lo, hi)
        finally:
            if gc_was_on:
                gc.enable()
        return g

    def run(self):
        stream = torch.cuda.current_stream(self.device)
        if not self._warm:
            self._launch_all(stream.cuda_stream)  # eager pass: sets func attributes, validates every launch
            stream.synchronize()
            self._warm = True
            if self.use_graph:
                self.graph = self._capture(0, len(self._calls))
                self.graph.replay()  # outputs of this call come from the replayed graph
            return
        if self.graph is not None:
            self.graph.replay()
        else:
            self._launch_all(stream.cuda_stream)

    def run_eager(self):
        self._launch_all(torch.cuda.current_stream(self.device).cuda_stream)

    # -- NMS overlapped with the next call's forward -------------------------------------------------
    # The batched NMS is a latency-bound tail (one CTA per image: 64 of 148 SMs at batch 64, ~5 % of the step).  In
    # overlap mode the plan is replayed as three graphs: [everything before the first writer of y] and [the writers of y:
    # the decode, or the six fused Detect epilogues] on the caller's
    # stream, [NMS] on a side stream, so the NMS of call i runs underneath the forward of call i+1.  The only buffer
    # both touch is the decoded prediction tensor y: the decode of call i+1 waits for the NMS of call i (an event that
    # fired ~3 ms earlier in steady state).  NMS reads / writes persistent buffers only (y, its workspace, det, keep,
    # count), which the arena's lifetime packing never recycles.  Consumers of det / keep / count must order
    # themselves after `nms_done` (join(), or work on `tail_stream`).
    def enable_overlap(self):
        fns = [n.fn for n in self.plan.nodes]
        ybuf = self.plan.outputs["y"].buf if "y" in self.plan.outputs else None
        # first node that writes the prediction tensor: fce_detect_decode, or the first fused Detect epilogue
        k = next((i for i, nd in enumerate(self.plan.nodes) if any(v.buf is ybuf for v in nd.writes)), None)
        if len(fns) < 3 or fns[-1] != "fce_nms" or k is None or k >= len(fns) - 1:
            raise RuntimeError("NMS overlap needs a plan that ends with the prediction writers and fce_nms")
        if any(any(v.buf is ybuf for v in nd.reads) for nd in self.plan.nodes[:k]):
            raise RuntimeError("NMS overlap: a node before the prediction writers reads the prediction tensor")
        if not self.use_graph:
            raise RuntimeError("NMS overlap needs CUDA graphs")
        with torch.cuda.device(self.device):
            self.tail_stream = torch.cuda.Stream(self.device)
            self.dec_done, self.nms_done = torch.cuda.Event(), torch.cuda.Event()
            stream = torch.cuda.current_stream(self.device)
            self._launch_all(stream.cuda_stream)  # eager pass: func attributes, launch validation
            stream.synchronize()
            self._warm = True
            n = len(self._calls)
            self._segs = []
            for lo, hi in ((0, k), (k, n - 1), (n - 1, n)):
                self._segs.append(self._capture(lo, hi))
            self.nms_done.record(stream)
        self.overlap = True

    def run_overlapped(self):
        main = torch.cuda.current_stream(self.device)
        head, dec, nms = self._segs
        head.replay()
        main.wait_event(self.nms_done)  # the previous call's NMS is done reading y
        dec.replay()
        self.dec_done.record(main)
        with torch.cuda.stream(self.tail_stream):
            self.tail_stream.wait_event(self.dec_done)
            nms.replay()
            self.nms_done.record(self.tail_stream)

    def join(self):
        """Orders the caller's current stream after the last NMS (no-op unless overlap mode is on)."""
        if getattr(self, "overlap", False):
            torch.cuda.current_stream(self.device).wait_event(self.nms_done)


# ---------------------------------------------------------------------------------------------------
# caches + public entry points used by modules.py / tasks.py
# ---------------------------------------------------------------------------------------------------
_model_cache: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


def _params_version(m: torch.nn.Module):
    return tuple((p.data_ptr(), p._version) for p in m.parameters()) + \
        tuple((b.data_ptr(), b._version) for b in m.buffers())


def _require_cuda(x):
    t = x[0] if isinstance(x, (list, tuple)) else x
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("fce_yolo_b200 runs on a B200 GPU only: pass CUDA tensors (there is no CPU fallback)")
    return t.device


def precision_of(model) -> str:
    p = getattr(model, "fce_precision", None)
    return p or _DEFAULT_PRECISION


def get_executor(model, x: torch.Tensor, precision: str | None = None, **kw) -> Executor:
    device = _require_cuda(x)
    precision = precision or precision_of(model)
    B, Cc, H, W = x.shape
    u8 = x.dtype == torch.uint8
    key = (B, H, W, precision, u8, device.index, tuple(sorted(kw.items())))
    per_model = _model_cache.setdefault(model, {})
    ver = _params_version(model)
    hit = per_model.get(key)
    if hit is not None and hit[0] == ver:
        return hit[1]
    with torch.cuda.device(device):
        plan = compile_model(model, B, H, W, precision, device, input_u8=u8, impl=kw.get("impl", 0))
        ex = Executor(plan, reuse_memory=kw.get("reuse_memory", True), use_graph=kw.get("use_graph", True))
    per_model[key] = (ver, ex)
    return ex


@torch.no_grad()
def run_model(model, x: torch.Tensor, precision: str | None = None, **kw):
    """DetectionModel forward: returns (y[B,4+nc,A] fp32, [raw NCHW fp32 per level]) like the reference's
    eval-mode Detect (head.py:121-124).  Outputs alias the executor's arena: clone to keep them across calls."""
    if model.training:
        raise RuntimeError("fce_yolo_b200 implements the inference forward only (model.eval())")
    ex = get_executor(model, x, precision, **kw)
    with torch.cuda.device(ex.device):
        inp = ex.input_tensor()
        if x.dtype == torch.uint8:
            if x.shape != inp.shape:
                raise ValueError(f"uint8 input must be NHWC {tuple(inp.shape)}")
            inp.copy_(x, non_blocking=True)
        else:
            inp.copy_(x.to(torch.float32), non_blocking=True)
        ex.run()
        return ex.outputs()


_module_cache: "weakref.WeakKeyDictionary" = weakref.WeakKeyDictionary()


@torch.no_grad()
def run_module(m, x, precision: str | None = None, impl: int = 0):
    """Per-module drop-in: logical NCHW tensor(s) in, logical NCHW tensor out (channels_last memory)."""
    device = _require_cuda(x)
    precision = precision or precision_of(m)
    xs = list(x) if isinstance(x, (list, tuple)) else [x]
    shapes = tuple(tuple(t.shape) for t in xs)
    key = (shapes, precision, device.index, impl)
    per = _module_cache.setdefault(m, {})
    ver = _params_version(m)
    hit = per.get(key)
    if hit is None or hit[0] != ver:
        with torch.cuda.device(device):
            plan = compile_module(m, [tuple(s) for s in shapes], precision, device, impl=impl)
            ex = Executor(plan, reuse_memory=False, use_graph=False)
        per[key] = (ver, ex)
    else:
        ex = hit[1]
    with torch.cuda.device(device):
        for i, t in enumerate(xs):
            ex.nchw(ex.plan.inputs[i]).copy_(t)
        ex.run_eager()
        out = ex.outputs()
    if isinstance(out, tuple):
        return out[0].clone(), [r.clone() for r in out[1]]
    return out.to(xs[0].dtype)
