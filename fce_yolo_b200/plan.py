"""Static plan compiler: nn.Module graph -> list of kernel launches over NHWC buffers.

The reference executes ~350-590 ATen ops per forward from a Python loop
(ultralytics/nn/tasks.py:160-188).  Here the same graph is compiled ONCE, for a fixed
(batch, height, width, precision), into a flat list of C-ABI calls (include/fce_yolo_b200.h):

 * every Conv/BN/SiLU (+ Bottleneck / PSABlock residual) is one fused conv launch;
 * torch.cat / chunk / split never move data: producers write, consumers read, channel slices of
   one NHWC buffer;
 * nn.Upsample feeding BiFPN_Concat is folded into the fusion kernel's read pattern, and a 1x1
   realign conv behind an Upsample runs at the low resolution (pointwise ops commute with nearest
   upsampling);
 * q/k/v strip projections that share an input are merged into one conv by stacking weights.

The compiler walks modules by *class name* and attribute names only, so it accepts both the mirror
classes in fce_yolo_b200.modules and the reference's own ultralytics instances (drop-in).
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass, field

import torch
import torch.nn as nn

from . import _lib as L

DT_SIZE = {L.BF16: 2, L.F32: 4, L.U8: 1}
TORCH_DT = {L.BF16: torch.bfloat16, L.F32: torch.float32, L.U8: torch.uint8}


@dataclass
class Buf:
    """An NHWC activation buffer [B, H, W, C] (strips: [1, rows, 1, C])."""
    id: int
    B: int
    H: int
    W: int
    C: int
    dtype: int
    persistent: bool = False
    first: int = -1
    last: int = -1
    offset: int = -1  # byte offset in the arena

    @property
    def nbytes(self):
        return self.B * self.H * self.W * self.C * DT_SIZE[self.dtype]


@dataclass
class View:
    """Channel slice [c0, c0+C) and row range of a Buf."""
    buf: Buf
    c0: int
    C: int
    B: int
    H: int
    W: int
    row0: int = 0  # first pixel-row (strips only)

    @property
    def pitch(self):
        return self.buf.C

    @property
    def dtype(self):
        return self.buf.dtype

    def ch(self, c0, c1):
        assert 0 <= c0 < c1 <= self.C
        return View(self.buf, self.c0 + c0, c1 - c0, self.B, self.H, self.W, self.row0)

    def rows(self, r0, r1):
        """Row sub-range of a strip view (B == 1, W == 1)."""
        assert self.B == 1 and self.W == 1 and 0 <= r0 < r1 <= self.H
        return View(self.buf, self.c0, self.C, 1, r1 - r0, 1, self.row0 + r0)

    def byte_offset(self):
        return (self.row0 * self.buf.C + self.c0) * DT_SIZE[self.buf.dtype]


@dataclass
class LazyUp:
    """A nearest-2x upsample that has not been materialised (consumer may fold it)."""
    src: View
    mat: View | None = None


@dataclass
class LazySum:
    """A two-input BiFPN_Concat node with identity realigns, sum_i wn[i] * x_i (fce_block.py:55-63), that has not been
    materialised: a 1x1-conv consumer folds it into two convs (Plan.conv_of_sum) and the fused map is never stored."""
    terms: list             # [(View, weight, upsampled)]
    H: int
    W: int
    C: int
    tag: str = ""
    mat: View | None = None


@dataclass
class Node:
    fn: str                 # C-ABI symbol
    desc: object            # ctypes struct
    ptrs: list              # entries: View | torch.Tensor | None | ("ws", nbytes) | int
    reads: list = field(default_factory=list)
    writes: list = field(default_factory=list)
    tag: str = ""
    flops: float = 0.0      # algorithmic FLOPs (2*MAC) for convs
    bytes: float = 0.0      # algorithmic bytes for bandwidth kernels
    stream: int = 0         # 0 = the caller's stream; k > 0 = side branch k of the captured graph (engine.Executor)
    disjoint: str = ""      # nodes sharing a non-empty key write DISJOINT parts of a common buffer (no WAW edge)


class PlanError(ValueError):
    pass


class Plan:
    fuse_decode = False  # set by compile_model: Detect's last convs decode in their epilogue (fce_conv2d_detect)
    # Detect's conv chains as concurrent branches of the captured graph (fused-decode plans).  None = automatic: on for
    # small workloads, where the step is a chain of latency-bound launches (measured, yolo11s-fce 640^2: batch 1
    # 0.714 -> 0.614 ms), off for large batches, where every kernel already fills the GPU and concurrent persistent
    # kernels only time-slice (batch 64: 3.157 -> 3.195 ms).  FCE_HEAD_STREAMS=0/1 or Plan.HEAD_STREAMS = bool overrides.
    HEAD_STREAMS = {"0": False, "1": True}.get(os.environ.get("FCE_HEAD_STREAMS", ""), None)
    HEAD_STREAMS_MAX_PIXELS = 8 * 640 * 640  # automatic mode: batch * H * W of the network input
    FUSED_BIFPN = True  # False: realign convs + fce_bifpn_fuse as separate launches (A/B timing, cross-check)
    # False (or FCE_FUSED_SUM=0): identity-realign BiFPN nodes always run fce_bifpn_fuse (A/B timing, cross-check)
    FUSED_SUM = os.environ.get("FCE_FUSED_SUM", "1") != "0"
    FUSED_COORDATT_MLP = True  # False: cv1 / cv_h / cv_w as three strip convs (A/B timing, cross-check)
    FUSED_C3K_IN = True  # False: C3k.cv1 and C3k.cv2 as two launches (A/B timing, cross-check)
    # False (or FCE_FUSED_C3K_TAIL=0): the last C3k's cv3 and C3k2's cv2 as two launches (A/B timing, cross-check)
    FUSED_C3K_TAIL = os.environ.get("FCE_FUSED_C3K_TAIL", "1") != "0"
    # False (or FCE_FUSED_DWPW=0): Detect's DWConv + 1x1 blocks as two launches (A/B timing, cross-check)
    FUSED_DWPW = os.environ.get("FCE_FUSED_DWPW", "1") != "0"
    # False (or FCE_FUSED_STEM2=0): the first two convs as two launches (A/B timing, cross-check)
    FUSED_STEM2 = os.environ.get("FCE_FUSED_STEM2", "1") != "0"
    FUSED_STEM = True  # False: stem = fce_stem_pack + K=32 tcgen05 conv (kept for A/B timing and as a cross-check)

    def __init__(self, batch: int, precision: str, device, impl: int = 0):
        if precision not in ("bf16", "fp32"):
            raise PlanError(f"precision must be 'bf16' or 'fp32', got {precision}")
        self.B = batch
        self.precision = precision
        self.act_dt = L.BF16 if precision == "bf16" else L.F32
        self.device = device
        self.impl = impl
        self.bufs: list[Buf] = []
        self.cur_stream = 0     # stream id stamped on the nodes being added (0 = main)
        self.branching = False  # set by compile_model: independent sub-chains become branches of the captured graph
        self.n_branches = 0
        self.nodes: list[Node] = []
        self.weights: list[torch.Tensor] = []  # keeps packed device tensors alive
        self.layer_out: dict[int, object] = {}
        self.inputs: list[View] = []
        self.outputs: dict[str, object] = {}

    # ------------------------------------------------------------------ buffers
    def new_buf(self, H, W, C, dtype=None, B=None, persistent=False) -> View:
        b = Buf(len(self.bufs), self.B if B is None else B, H, W, C, self.act_dt if dtype is None else dtype,
                persistent)
        self.bufs.append(b)
        return View(b, 0, C, b.B, H, W)

    def strip_buf(self, rows, C) -> View:
        return self.new_buf(rows, 1, C, dtype=L.F32, B=1)

    def _w(self, t: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
        t = t.detach().to(device=self.device, dtype=dtype).contiguous()
        self.weights.append(t)
        return t

    def add(self, node: Node):
        idx = len(self.nodes)
        node.stream = self.cur_stream
        for v in node.reads + node.writes:
            b = v.buf
            if b.first < 0:
                b.first = idx
            b.last = idx
            if node.stream:  # touched by a concurrent branch: never recycled by the sequential lifetime packing
                b.persistent = True
        self.nodes.append(node)

    @staticmethod
    def _overlap(a: View, b: View) -> bool:
        """Do two views touch common elements?  Channel slices of one buffer (concat-by-offset) and row ranges of a
        strip are told apart; anything else on the same buffer counts as overlapping."""
        if a.buf is not b.buf:
            return False
        if a.c0 >= b.c0 + b.C or b.c0 >= a.c0 + a.C:
            return False
        if a.buf.B == 1 and a.buf.W == 1 and a.W == 1 and b.W == 1:  # strip rows
            if a.row0 >= b.row0 + b.H or b.row0 >= a.row0 + a.H:
                return False
        return True

    def dependencies(self, lo: int = 0, hi: int | None = None):
        """deps[i] = indices j < i (both inside [lo, hi)) that node i must be ordered after: RAW, WAR and WAW between
        overlapping views, except WAW between nodes that declare the same `disjoint` key.  Used by the executor to turn
        the side branches into graph edges.  Conservative: views are compared by buffer, channel range and (strips) row
        range only, so a missing edge is impossible; redundant (transitively implied) edges are kept."""
        hi = len(self.nodes) if hi is None else hi
        deps = {}
        for i in range(lo, hi):
            n = self.nodes[i]
            d = set()
            for j in range(lo, i):
                m = self.nodes[j]
                hit = any(self._overlap(r, w) for r in n.reads for w in m.writes) or \
                    any(self._overlap(w, r) for w in n.writes for r in m.reads)
                if not hit and not (n.disjoint and m.disjoint == n.disjoint):
                    hit = any(self._overlap(w, v) for w in n.writes for v in m.writes)
                if hit:
                    d.add(j)
            deps[i] = d
        return deps

    def branch(self):
        """Context manager: nodes added inside go to a fresh side branch when branching is on (small workloads)."""
        plan = self

        class _B:
            def __enter__(self_b):
                self_b.saved = plan.cur_stream
                if plan.branching and plan.cur_stream == 0:
                    plan.n_branches += 1
                    plan.cur_stream = plan.n_branches
                return self_b

            def __exit__(self_b, *exc):
                plan.cur_stream = self_b.saved
                return False

        return _B()

    # ------------------------------------------------------------------ conv family
    @staticmethod
    def conv_params(m):
        """(weight[O,I,kh,kw] fp32, bias[O] fp32, k, stride, groups, act) of a Conv-like module or bare Conv2d,
        with BatchNorm folded if still present (reference torch_utils.py:237-267)."""
        if isinstance(m, nn.Conv2d):
            conv, bn, act = m, None, L.ACT_NONE
        else:
            conv, bn = m.conv, getattr(m, "bn", None)
            a = getattr(m, "act", None)
            if isinstance(a, nn.SiLU):
                act = L.ACT_SILU
            elif a is None or isinstance(a, nn.Identity):
                act = L.ACT_NONE
            else:
                raise PlanError(f"activation {type(a).__name__} has no fused epilogue (only SiLU / identity)")
        w = conv.weight.detach().float()
        b = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=w.device)
        if isinstance(bn, nn.BatchNorm2d):
            s = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
            w = w * s.view(-1, 1, 1, 1)
            b = bn.bias.detach().float() + (b - bn.running_mean.detach().float()) * s
        k = conv.kernel_size[0]
        if conv.kernel_size[0] != conv.kernel_size[1] or k not in (1, 3) or conv.dilation[0] != 1:
            raise PlanError(f"conv kernel {conv.kernel_size} / dilation {conv.dilation} outside the path")
        if conv.padding[0] != k // 2:
            raise PlanError("only 'same' padding (k//2) is on the path")
        return w, b, k, conv.stride[0], conv.groups, act

    def conv(self, m, x: View, dst: View | None = None, res: View | None = None, out_dtype=None, act=None,
             w_override=None, b_override=None, in_layout=L.NHWC, in_scale=1.0, tag="", out_scale=1.0, res_scale=1.0,
             res_up=False, res_pre=False) -> View:
        w, b, k, s, g, a = self.conv_params(m) if m is not None else (w_override, b_override, 1, 1, 1, L.ACT_NONE)
        if w_override is not None:
            w, b = w_override, b_override
        if act is not None:
            a = act
        if g != 1:
            raise PlanError("grouped conv goes through dwconv()")
        Cout, Cin = w.shape[0], w.shape[1]
        if Cin != x.C:
            raise PlanError(f"{tag}: conv expects {Cin} input channels, view has {x.C}")
        Ho = (x.H + 2 * (k // 2) - k) // s + 1
        Wo = (x.W + 2 * (k // 2) - k) // s + 1
        strip = x.dtype == L.F32 and self.act_dt != L.F32 and in_layout == L.NHWC  # fp32 strips in bf16 mode
        if out_dtype is None:
            out_dtype = L.F32 if strip else self.act_dt
        if dst is None:
            dst = self.new_buf(Ho, Wo, Cout, dtype=out_dtype, B=x.B)
        if (dst.H, dst.W, dst.C, dst.B) != (Ho, Wo, Cout, x.B):
            raise PlanError(f"{tag}: conv output {(x.B, Ho, Wo, Cout)} does not fit destination "
                            f"{(dst.B, dst.H, dst.W, dst.C)}")
        w_dt = L.F32 if (x.dtype == L.F32 and (strip or self.act_dt == L.F32)) else self.act_dt
        if in_layout == L.NCHW:
            w_dt = self.act_dt
        wp = self._w(w.permute(0, 2, 3, 1), TORCH_DT[w_dt])  # OHWI
        bp = self._w(b)
        d = L.ConvDesc(B=x.B, H=x.H, W=x.W, Cin=Cin, Cout=Cout, in_pitch=x.pitch, in_off=0, out_pitch=dst.pitch,
                       out_off=0, res_pitch=res.pitch if res is not None else 0, res_off=0, k=k, stride=s, act=a,
                       in_dtype=x.dtype, w_dtype=w_dt, out_dtype=dst.dtype, in_layout=in_layout, in_scale=in_scale,
                       impl=self.impl, weighted=2 if res_pre else (1 if (out_scale != 1.0 or res_scale != 1.0) else 0),
                       out_scale=float(out_scale), res_scale=float(res_scale), res_up=1 if res_up else 0)
        rs = 2 if res_up else 1
        if res is not None and (res.C != Cout or res.dtype != dst.dtype or (res.H * rs, res.W * rs) != (Ho, Wo)):
            raise PlanError(f"{tag}: residual view does not match the conv output")
        self.add(Node("fce_conv2d", d, [x, wp, bp, res, dst], reads=[x] + ([res] if res is not None else []),
                      writes=[dst], tag=tag, flops=2.0 * x.B * Ho * Wo * Cout * Cin * k * k,
                      bytes=(x.B * x.H * x.W * Cin * DT_SIZE[x.dtype] + x.B * Ho * Wo * Cout * DT_SIZE[dst.dtype]
                             * (2 if res is not None else 1) + Cout * Cin * k * k * DT_SIZE[w_dt])))
        return dst

    def stem(self, m, img: View, in_layout, in_scale, tag="") -> View:
        """First Conv of the graph.  bf16 mode, 3x3/s2 on 3 channels: the fused stem kernel (fce_stem_conv); for
        channel counts it does not cover (or Plan.FUSED_STEM = False): patch packing (fce_stem_pack) into a [M, 32] bf16 matrix
        followed by a K = 32 1x1 conv on the tensor cores; anything else: the generic conv."""
        w, b, k, s, g, a = self.conv_params(m)
        Cout, Cin = w.shape[0], w.shape[1]
        if not (self.act_dt == L.BF16 and k == 3 and s == 2 and g == 1 and Cin == 3 and Cout % 16 == 0
                and self.impl != 1):
            return self.conv(m, img, in_layout=in_layout, in_scale=in_scale, tag=tag)
        H, W = (img.H, img.W)
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        # OHWI weights flattened to [Cout, 27] (+5 zero columns), input scale folded in
        wk = torch.zeros(Cout, 32, 1, 1, device=w.device)
        wk[:, :27, 0, 0] = (w.permute(0, 2, 3, 1).reshape(Cout, 27) * in_scale)
        if Cout in (16, 32, 48, 64, 96) and a in (L.ACT_SILU, L.ACT_NONE) and self.FUSED_STEM:
            # one fused pass: image tile -> smem -> mma.sync -> bias + SiLU -> NHWC bf16
            out = self.new_buf(Ho, Wo, Cout, dtype=L.BF16, B=img.B)
            wp = self._w(wk.view(Cout, 32), torch.bfloat16)
            bp = self._w(b)
            d = L.StemDesc(B=img.B, H=H, W=W, Cout=Cout, out_pitch=out.pitch, out_off=0, act=a, in_dtype=img.dtype,
                           in_layout=in_layout)
            self.add(Node("fce_stem_conv", d, [img, wp, bp, out], reads=[img], writes=[out], tag=tag,
                          flops=2.0 * img.B * Ho * Wo * Cout * 27,
                          bytes=img.B * (H * W * 3.0 * DT_SIZE[img.dtype] + Ho * Wo * Cout * 2.0)))
            return out
        packed = self.new_buf(Ho, Wo, 32, dtype=L.BF16, B=img.B)
        d = L.PackDesc(B=img.B, H=H, W=W, Cin=3, k=3, stride=2, Kpad=32, in_dtype=img.dtype, in_layout=in_layout)
        self.add(Node("fce_stem_pack", d, [img, packed], reads=[img], writes=[packed], tag=tag + ".pack",
                      bytes=img.B * (H * W * 3 * DT_SIZE[img.dtype] + Ho * Wo * 64.0)))
        return self.conv(None, packed, w_override=wk, b_override=b, act=a, tag=tag)

    def stem2(self, m0, m1, img: View, in_scale, tag0="", tag1=""):
        """The first two Convs of the graph (3x3/s2 on the uint8 NHWC image, then 3x3/s2) as ONE launch (fce_stem2_conv:
        the stem map stays in shared memory).  Returns the second conv's output view, or None when the kernel does not
        take the shape (the caller then emits the two layers as usual)."""
        if not (self.FUSED_STEM2 and self.FUSED_STEM and self.act_dt == L.BF16 and self.impl in (0, 2) and img.dtype == L.U8):
            return None
        if type(m1).__name__ != "Conv" or getattr(m1, "f", -1) != -1:
            return None
        w0, b0, k0, s0, g0, a0 = self.conv_params(m0)
        w1, b1, k1, s1, g1, a1 = self.conv_params(m1)
        if (k0, s0, g0, k1, s1, g1) != (3, 2, 1, 3, 2, 1) or w0.shape[1] != 3 or w1.shape[1] != w0.shape[0]:
            return None
        C0, C1 = w0.shape[0], w1.shape[0]
        H, W = img.H, img.W
        d = L.Stem2Desc(B=img.B, H=H, W=W, C0=C0, C1=C1, out_pitch=C1, out_off=0, act0=a0, act1=a1)
        if L.load().fce_stem2_route(ctypes.byref(d)) != 1:
            return None
        out = self.new_buf(H // 4, W // 4, C1, dtype=L.BF16, B=img.B)
        wk = torch.zeros(C0, 32, device=w0.device)
        wk[:, :27] = w0.permute(0, 2, 3, 1).reshape(C0, 27) * in_scale  # as Plan.stem: OHWI rows, input scale folded in
        ptrs = [img, self._w(wk, torch.bfloat16), self._w(b0), self._w(w1.permute(0, 2, 3, 1), torch.bfloat16), self._w(b1), out]
        px0, px1 = img.B * (H // 2) * (W // 2), img.B * (H // 4) * (W // 4)
        self.add(Node("fce_stem2_conv", d, ptrs, reads=[img], writes=[out], tag=f"{tag0}+{tag1}",
                      flops=2.0 * px0 * C0 * 27 + 2.0 * px1 * C1 * C0 * 9,
                      bytes=img.B * H * W * 3.0 + px1 * C1 * 2.0 + C1 * C0 * 9 * 2.0))
        return out

    def dwconv(self, m, x: View, dst: View | None = None, add: View | None = None, tag="") -> View:
        w, b, k, s, g, a = self.conv_params(m)
        if k != 3 or s != 1 or g != x.C or w.shape[0] != x.C or w.shape[1] != 1:
            raise PlanError(f"{tag}: only depthwise 3x3 stride-1 convs are on the path")
        if dst is None:
            dst = self.new_buf(x.H, x.W, x.C)
        wp = self._w(w.view(x.C, 9).t())  # [9][C]
        bp = self._w(b)
        d = L.DwconvDesc(B=x.B, H=x.H, W=x.W, C=x.C, in_pitch=x.pitch, in_off=0, out_pitch=dst.pitch, out_off=0,
                         add_pitch=add.pitch if add is not None else 0, add_off=0, act=a, dtype=x.dtype)
        esz = DT_SIZE[x.dtype]
        self.add(Node("fce_dwconv3x3", d, [x, wp, bp, add, dst], reads=[x] + ([add] if add is not None else []),
                      writes=[dst], tag=tag, bytes=(2 + (add is not None)) * x.B * x.H * x.W * x.C * esz))
        return dst

    def dwpw(self, m_dw, m_pw, x: View, dst: View | None = None, tag="") -> View:
        """DWConv(c, c, 3) followed by Conv(c, c2, 1) (one block of Detect.cv3, head.py:101-102) as ONE launch
        (fce_dwpw_conv: the depthwise result stays in shared memory) where the kernel takes the shape; two launches
        otherwise.  Same values either way (tests/test_gpu_dwpw.py)."""
        wd, bd, kd, sd, gd, ad = self.conv_params(m_dw)
        wp_, bp_, kp, sp, gp, ap = self.conv_params(m_pw)
        Cc, Cout = x.C, wp_.shape[0]
        d = L.DwpwDesc(B=x.B, H=x.H, W=x.W, C=Cc, Cout=Cout, in_pitch=x.pitch, in_off=0,
                       out_pitch=dst.pitch if dst is not None else Cout, out_off=0, dw_act=ad, pw_act=ap)
        ok = (self.FUSED_DWPW and self.act_dt == L.BF16 and self.impl in (0, 2) and x.dtype == L.BF16
              and (kd, sd, gd) == (3, 1, Cc) and wd.shape[0] == Cc and wd.shape[1] == 1
              and (kp, sp, gp) == (1, 1, 1) and wp_.shape[1] == Cc and (dst is None or dst.dtype == L.BF16)
              and L.load().fce_dwpw_route(ctypes.byref(d)) == 1)
        if not ok:
            t = self.dwconv(m_dw, x, tag=tag + ".0")
            return self.conv(m_pw, t, dst=dst, tag=tag + ".1")
        if dst is None:
            dst = self.new_buf(x.H, x.W, Cout)
        if (dst.H, dst.W, dst.C, dst.B) != (x.H, x.W, Cout, x.B):
            raise PlanError(f"{tag}: dwpw output does not fit its destination")
        ptrs = [x, self._w(wd.view(Cc, 9).t()), self._w(bd), self._w(wp_.view(Cout, Cc), torch.bfloat16), self._w(bp_), dst]
        px = x.B * x.H * x.W
        self.add(Node("fce_dwpw_conv", d, ptrs, reads=[x], writes=[dst], tag=tag + ".0+1",
                      flops=2.0 * px * Cout * Cc, bytes=px * (Cc + Cout) * 2.0 + Cout * Cc * 2.0))
        return dst

    # ------------------------------------------------------------------ blocks
    def bottleneck(self, m, x: View, dst=None, tag="") -> View:
        t = self.conv(m.cv1, x, tag=tag + ".cv1")
        return self.conv(m.cv2, t, dst=dst, res=x if m.add else None, tag=tag + ".cv2")

    def _c3k_in_fused(self, m) -> bool:
        p1, p2 = self.conv_params(m.cv1), self.conv_params(m.cv2)
        return bool(self.FUSED_C3K_IN and p1[2:] == p2[2:] and p1[2:5] == (1, 1, 1) and p1[0].shape == p2[0].shape
                    and m.cv1.conv.out_channels % 8 == 0)

    def c3k(self, m, x: View, dst=None, tag="", defer_cv3=False) -> View:
        """defer_cv3: everything but cv3; returns cv3's INPUT view (the caller chains cv3 into its own 1x1, c3k2())."""
        c_ = m.cv1.conv.out_channels
        blocks = list(m.m)
        if not blocks:
            raise PlanError("C3k without bottlenecks")
        p1, p2 = self.conv_params(m.cv1), self.conv_params(m.cv2)
        if self._c3k_in_fused(m):
            # cv1 and cv2 are 1x1 convs of the SAME input (block.py:338-340): ONE launch with the weights stacked reads x
            # once instead of twice (both layers sit left of the ridge: HBM-bound).  Buffer layout [chain out | cv2(x) |
            # cv1(x)]: the stacked conv writes channels [c_, 3c_), the bottleneck chain reads the last slice and ends in the
            # first, cv3 reads the first two - still no concat copy anywhere.
            buf = self.new_buf(x.H, x.W, 3 * c_)
            self.conv(None, x, dst=buf.ch(c_, 3 * c_), w_override=torch.cat([p2[0], p1[0]]),
                      b_override=torch.cat([p2[1], p1[1]]), act=p1[5], tag=tag + ".cv2+cv1")
            a = buf.ch(2 * c_, 3 * c_)
            for j, blk in enumerate(blocks):
                a = self.bottleneck(blk, a, dst=buf.ch(0, c_) if j == len(blocks) - 1 else None, tag=f"{tag}.m.{j}")
            if defer_cv3:
                return buf.ch(0, 2 * c_)
            return self.conv(m.cv3, buf.ch(0, 2 * c_), dst=dst, tag=tag + ".cv3")
        cat = self.new_buf(x.H, x.W, 2 * c_)
        a = self.conv(m.cv1, x, tag=tag + ".cv1")
        for j, blk in enumerate(blocks):
            a = self.bottleneck(blk, a, dst=cat.ch(0, c_) if j == len(blocks) - 1 else None, tag=f"{tag}.m.{j}")
        with self.branch():  # independent of the bottleneck chain: same input, its own half of the concat buffer
            self.conv(m.cv2, x, dst=cat.ch(c_, 2 * c_), tag=tag + ".cv2")
        if defer_cv3:
            return cat
        return self.conv(m.cv3, cat, dst=dst, tag=tag + ".cv3")

    def _c3k_tail_desc(self, m, H, W, B, dst):
        """Descriptor of fce_conv1x1_chain for C3k2 `m` (its last block's cv3 chained into cv2) if that kernel takes the block,
        else None."""
        c, n = m.c, len(m.m)
        if not (self.FUSED_C3K_TAIL and self.act_dt == L.BF16 and self.impl in (0, 2) and n >= 1
                and type(m.m[-1]).__name__ == "C3k" and list(m.m[-1].m)):
            return None
        p3, p2 = self.conv_params(m.m[-1].cv3), self.conv_params(m.cv2)
        Cout = p2[0].shape[0]
        if p3[2:5] != (1, 1, 1) or p2[2:5] != (1, 1, 1) or tuple(p3[0].shape[:2]) != (c, c) or p2[0].shape[1] != (2 + n) * c:
            return None
        if dst is not None and dst.dtype != L.BF16:
            return None
        c_ = m.m[-1].cv1.conv.out_channels
        x1_pitch = 3 * c_ if self._c3k_in_fused(m.m[-1]) else 2 * c_
        d = L.ChainDesc(B=B, H=H, W=W, c1=c, cm=c, c2=(1 + n) * c, Cout=Cout, x1_pitch=x1_pitch, x1_off=0,
                        x2_pitch=(1 + n) * c, x2_off=0, out_pitch=dst.pitch if dst is not None else Cout, out_off=0,
                        act1=p3[5], act2=p2[5])
        return d if L.load().fce_conv1x1_chain_route(ctypes.byref(d)) == 1 else None

    def c3k2(self, m, x, dst=None, tag="") -> View:
        c, n = m.c, len(m.m)
        chain = self._c3k_tail_desc(m, x.H, x.W, getattr(x, "B", self.B), dst)
        cat = self.new_buf(x.H, x.W, ((1 if chain is not None else 2) + n) * c)
        if isinstance(x, LazySum):  # BiFPN node in front: folded into cv1
            self.conv_of_sum(m.cv1, x, dst=cat.ch(0, 2 * c), tag=tag + ".cv1")
        else:
            self.conv(m.cv1, x, dst=cat.ch(0, 2 * c), tag=tag + ".cv1")
        prev = cat.ch(c, 2 * c)
        for j, blk in enumerate(m.m):
            if chain is not None and j == n - 1:
                # the last C3k's cv3 output t is only ever read by cv2: ONE launch computes t = cv3(..) tile by tile into shared
                # memory and uses it as the last K slice of cv2 (fce_conv1x1_chain) - the concat buffer has no slot for it
                x1 = self.c3k(blk, prev, tag=f"{tag}.m.{j}", defer_cv3=True)
                return self._chain(blk.cv3, m.cv2, chain, x1, cat, dst, tag=f"{tag}.m.{j}.cv3+cv2")
            out = cat.ch((2 + j) * c, (3 + j) * c)
            if type(blk).__name__ == "C3k":
                self.c3k(blk, prev, dst=out, tag=f"{tag}.m.{j}")
            else:
                self.bottleneck(blk, prev, dst=out, tag=f"{tag}.m.{j}")
            prev = out
        return self.conv(m.cv2, cat, dst=dst, tag=tag + ".cv2")

    def _chain(self, m1, m2, d, x1: View, x2: View, dst, tag="") -> View:
        w1, b1 = self.conv_params(m1)[:2]
        w2, b2 = self.conv_params(m2)[:2]
        if (x1.C, x2.C, x1.pitch, x2.pitch) != (d.c1, d.c2, d.x1_pitch, d.x2_pitch) or x1.dtype != L.BF16 or x2.dtype != L.BF16:
            raise PlanError(f"{tag}: views do not match the chain descriptor")
        if dst is None:
            dst = self.new_buf(x2.H, x2.W, d.Cout)
        if (dst.H, dst.W, dst.C, dst.B, dst.pitch) != (x2.H, x2.W, d.Cout, x2.B, d.out_pitch):
            raise PlanError(f"{tag}: chain output does not fit its destination")
        ptrs = [x1, self._w(w1.view(d.cm, d.c1), torch.bfloat16), self._w(b1), x2,
                self._w(w2.view(d.Cout, d.c2 + d.cm), torch.bfloat16), self._w(b2), dst]
        px = x2.B * x2.H * x2.W
        self.add(Node("fce_conv1x1_chain", d, ptrs, reads=[x1, x2], writes=[dst], tag=tag,
                      flops=2.0 * px * (d.cm * d.c1 + d.Cout * (d.c2 + d.cm)),
                      bytes=px * (d.c1 + d.c2 + d.Cout) * 2.0 + (d.cm * d.c1 + d.Cout * (d.c2 + d.cm)) * 2.0))
        return dst

    def sppf(self, m, x: View, dst=None, tag="") -> View:
        k = m.m.kernel_size if isinstance(m.m.kernel_size, int) else m.m.kernel_size[0]
        if k != 5:
            raise PlanError("SPPF kernel must be 5")
        c_ = m.cv1.conv.out_channels
        cat = self.new_buf(x.H, x.W, 4 * c_)
        self.conv(m.cv1, x, dst=cat.ch(0, c_), tag=tag + ".cv1")
        d = L.SppfDesc(B=x.B, H=x.H, W=x.W, C=c_, pitch=cat.pitch, off=0, dtype=cat.dtype)
        self.add(Node("fce_sppf_pool", d, [cat], reads=[cat], writes=[cat], tag=tag + ".pool",
                      bytes=4.0 * x.B * x.H * x.W * c_ * DT_SIZE[cat.dtype]))
        return self.conv(m.cv2, cat, dst=dst, tag=tag + ".cv2")

    def psa_block(self, blk, b: View, tag=""):
        at = blk.attn
        nh, kd, hd = at.num_heads, at.key_dim, at.head_dim
        c = b.C
        # qkv output channels re-ordered from per-head [q|k|v] (block.py:1296-1298) to [Q | K | V]
        w, bias, *_ = self.conv_params(at.qkv)
        per = 2 * kd + hd
        idx_q = [h * per + j for h in range(nh) for j in range(kd)]
        idx_k = [h * per + kd + j for h in range(nh) for j in range(kd)]
        idx_v = [h * per + 2 * kd + j for h in range(nh) for j in range(hd)]
        perm = torch.tensor(idx_q + idx_k + idx_v, device=w.device)
        qkv = self.conv(at.qkv, b, w_override=w[perm], b_override=bias[perm], tag=tag + ".attn.qkv")
        att = self.new_buf(b.H, b.W, c)
        d = L.PsaDesc(B=b.B, N=b.H * b.W, heads=nh, kd=kd, hd=hd, qkv_pitch=qkv.pitch, q_off=0, k_off=nh * kd,
                      v_off=2 * nh * kd, out_pitch=att.pitch, out_off=0, dtype=qkv.dtype, scale=float(at.scale))
        self.add(Node("fce_psa_attention", d, [qkv, att], reads=[qkv], writes=[att], tag=tag + ".attn.core",
                      flops=2.0 * b.B * nh * (b.H * b.W) ** 2 * (kd + hd)))
        self.dwconv(at.pe, qkv.ch(2 * nh * kd, 2 * nh * kd + c), dst=att, add=att, tag=tag + ".attn.pe")
        self.conv(at.proj, att, dst=b, res=b if blk.add else None, tag=tag + ".attn.proj")
        f = self.conv(blk.ffn[0], b, tag=tag + ".ffn.0")
        self.conv(blk.ffn[1], f, dst=b, res=b if blk.add else None, tag=tag + ".ffn.1")

    def c2psa(self, m, x: View, dst=None, tag="") -> View:
        c = m.c
        ab = self.new_buf(x.H, x.W, 2 * c)
        self.conv(m.cv1, x, dst=ab, tag=tag + ".cv1")
        for j, blk in enumerate(m.m):
            self.psa_block(blk, ab.ch(c, 2 * c), tag=f"{tag}.m.{j}")
        return self.conv(m.cv2, ab, dst=dst, tag=tag + ".cv2")

    # ------------------------------------------------------------------ resampling / fusion
    def materialize(self, x, dst: View | None = None) -> View:
        if isinstance(x, View):
            if dst is None:
                return x
            d = L.CopyDesc(B=x.B, H=x.H, W=x.W, C=x.C, in_pitch=x.pitch, in_off=0, out_pitch=dst.pitch, out_off=0,
                           dtype=x.dtype)
            self.add(Node("fce_copy_view", d, [x, dst], reads=[x], writes=[dst], tag="copy",
                          bytes=2.0 * x.B * x.H * x.W * x.C * DT_SIZE[x.dtype]))
            return dst
        if x.mat is not None and dst is None:
            return x.mat
        if isinstance(x, LazySum):
            views, ups, wn = [t[0] for t in x.terms], [1 if t[2] else 0 for t in x.terms], [t[1] for t in x.terms]
            out = self._fuse_node(views, ups, wn, x.H, x.W, x.C, dst, x.tag)
            if dst is None:
                x.mat = out
            return out
        s = x.src
        out = dst if dst is not None else self.new_buf(2 * s.H, 2 * s.W, s.C, dtype=s.dtype)
        d = L.UpsampleDesc(B=s.B, H=s.H, W=s.W, C=s.C, in_pitch=s.pitch, in_off=0, out_pitch=out.pitch, out_off=0,
                           dtype=s.dtype)
        self.add(Node("fce_upsample2x", d, [s, out], reads=[s], writes=[out], tag="upsample",
                      bytes=5.0 * s.B * s.H * s.W * s.C * DT_SIZE[s.dtype]))
        if dst is None:
            x.mat = out
        return out

    def concat(self, m, xs, tag="") -> View:
        if getattr(m, "d", 1) != 1:
            raise PlanError("Concat is only supported along channels")
        shapes = [(2 * x.src.H, 2 * x.src.W, x.src.C) if isinstance(x, LazyUp) else (x.H, x.W, x.C) for x in xs]
        H, W = shapes[0][:2]
        out = self.new_buf(H, W, sum(s[2] for s in shapes))
        c0 = 0
        for x, s in zip(xs, shapes):
            if s[:2] != (H, W):
                raise PlanError(f"{tag}: Concat inputs differ in size")
            self.materialize(x, dst=out.ch(c0, c0 + s[2]))
            c0 += s[2]
        return out

    def bifpn(self, m, xs, dst=None, tag="") -> View:
        if not 2 <= len(xs) <= 3:
            raise PlanError("BiFPN_Concat fuses 2 or 3 inputs")
        w = torch.relu(m.w.detach().float().cpu())
        wn = (w / (w.sum() + m.epsilon)).tolist()  # fce_block.py:55-56
        views, ups = [], []
        # Fusable: full-resolution inputs whose realign is a 1x1 Conv on the tensor-core path; at most ONE other operand
        # (each fused epilogue adds a single residual).  pending = [(index, conv, input view)], others = [(view, w, up)].
        def fusable(i, x):
            r = m.realign_convs[i]
            return (self.FUSED_BIFPN and self.act_dt == L.BF16 and self.impl != 1 and not isinstance(x, LazyUp)
                    and not isinstance(r, nn.Identity) and self.conv_params(r)[2:5] == (1, 1, 1)
                    and x.C % 16 == 0 and self.conv_params(r)[0].shape[0] % 16 == 0)
        n_fus = sum(1 for i, x in enumerate(xs) if fusable(i, x))
        # Two-input nodes only: measured on the three-input node (two chained fused convs, s scale, batch 64) the epilogue
        # residual reads cost more than the separate fusion launch saves (67 -> 72 us), while the two-input nodes gain
        # (108 -> 80 us at 80x80, 31 -> 21 us at 20x20)
        fuse = len(xs) == 2 and n_fus >= 1
        pending, others = [], []
        for i, x in enumerate(xs):
            r = m.realign_convs[i]
            up = isinstance(x, LazyUp)
            v = x.src if up else x
            if fuse and fusable(i, x):
                pending.append((i, r, v))
                # geometry checks below see the conv's OUTPUT shape
                views.append(View(v.buf, 0, self.conv_params(r)[0].shape[0], v.B, v.H, v.W))
                ups.append(0)
                continue
            if not isinstance(r, nn.Identity):
                with self.branch():  # realign convs of one node are independent of each other
                    v = self.conv(r, v, tag=f"{tag}.realign.{i}")  # at low resolution when upsampled
            if fuse:
                others.append((v, wn[i], up))
            views.append(v)
            ups.append(1 if up else 0)
        H, W = (views[0].H * 2, views[0].W * 2) if ups[0] else (views[0].H, views[0].W)
        C = views[0].C
        for v, u in zip(views, ups):
            if (v.H * (2 if u else 1), v.W * (2 if u else 1), v.C) != (H, W, C):
                raise PlanError(f"{tag}: BiFPN inputs disagree after realignment")
        if pending:
            # Weighted sum inside the realign convs' epilogues (bf16 plans): the first pending conv folds in the one
            # other operand (identity or upsampled view), every later one the running sum; the last writes dst.  The
            # realigned maps are never stored and re-read, and there is no separate fusion launch.
            acc, acc_w, acc_up = (others[0][0], others[0][1], others[0][2]) if others else (None, 1.0, False)
            for j, (i, r, v) in enumerate(pending):
                last = j == len(pending) - 1
                out = dst if (last and dst is not None) else self.new_buf(H, W, C)
                if acc is None:  # first of several convs and nothing to fold yet: plain conv, scaled later
                    acc, acc_w, acc_up = self.conv(r, v, tag=f"{tag}.realign.{i}"), wn[i], False
                    continue
                acc = self.conv(r, v, dst=out, res=acc, out_scale=wn[i], res_scale=acc_w, res_up=acc_up,
                                tag=f"{tag}.realign.{i}+fuse")
                acc_w, acc_up = 1.0, False
            return acc
        identity = all(isinstance(r, nn.Identity) for r in m.realign_convs)
        if (dst is None and self.FUSED_SUM and self.act_dt == L.BF16 and self.impl != 1 and len(xs) == 2 and identity
                and C % 16 == 0 and sum(ups) == 1):
            # nothing to compute yet: a 1x1-conv consumer (C3k2.cv1) takes the two maps directly.  Only with ONE upsampled
            # operand - its half of the conv then runs on a quarter of the pixels.  Two full-resolution operands: the
            # second conv costs what the fusion launch costs (m scale, batch 256, 20x20: 138 us separate, 167 us folded)
            return LazySum([(v, wn[i], bool(ups[i])) for i, v in enumerate(views)], H, W, C, tag)
        return self._fuse_node(views, ups, wn, H, W, C, dst, tag)

    def _fuse_node(self, views, ups, wn, H, W, C, dst, tag) -> View:
        if dst is None:
            dst = self.new_buf(H, W, C)
        d = L.BifpnDesc(B=self.B, H=H, W=W, C=C, n=len(views), out_pitch=dst.pitch, out_off=0, dtype=dst.dtype)
        for i, (v, u) in enumerate(zip(views, ups)):
            d.pitch[i], d.off[i], d.up[i], d.wn[i] = v.pitch, 0, u, wn[i]
        ptrs = views + [None] * (3 - len(views)) + [dst]
        esz = DT_SIZE[dst.dtype]
        alg = sum(v.B * v.H * v.W * v.C for v in views) * esz + self.B * H * W * C * esz
        self.add(Node("fce_bifpn_fuse", d, ptrs, reads=views, writes=[dst], tag=tag + ".fuse", bytes=alg))
        return dst

    def conv_of_sum(self, m, x: LazySum, dst: View | None = None, tag="") -> View:
        """1x1 Conv over a weighted sum of two maps without forming the sum:
        act(W (w0 a + w1 b) + bias) = act((w1 W) b + [(w0 W) a] + bias).  The bracket runs first as a plain conv with no
        bias / activation - at LOW resolution when a is the upsampled operand (a 1x1 conv commutes with nearest
        upsampling) - and joins the main conv's accumulator in its epilogue (fce_conv_desc.weighted == 2)."""
        w, b, k, s, g, a = self.conv_params(m)
        if (k, s, g) != (1, 1, 1) or w.shape[0] % 16:
            return self.conv(m, self.materialize(x), dst=dst, tag=tag)
        (v0, w0, u0), (v1, w1, u1) = x.terms
        if u1:  # the main conv runs on the full-resolution operand
            (v0, w0, u0), (v1, w1, u1) = (v1, w1, u1), (v0, w0, u0)
        side = self.conv(None, v0, w_override=w * w0, b_override=torch.zeros_like(b), act=L.ACT_NONE, tag=tag + ".side")
        return self.conv(None, v1, dst=dst, res=side, res_up=u0, res_pre=True, w_override=w * w1, b_override=b, act=a,
                         tag=tag + "+sum")

    # ------------------------------------------------------------------ coordinate attention family
    def _pool(self, x: View, tag) -> View:
        strip = self.strip_buf(x.B * (x.H + x.W), x.C)
        d = L.PoolDesc(B=x.B, H=x.H, W=x.W, C=x.C, pitch=x.pitch, off=0, dtype=x.dtype)
        bands = -(-x.H // 8)  # fce_coord_pool_workspace: fp32 column partials per 8-row band
        ws_bytes = x.B * bands * x.W * x.C * 4 if bands > 1 else 0
        ws = self.new_buf(1, 1, ws_bytes, dtype=L.U8, B=1) if ws_bytes else None
        self.add(Node("fce_coord_pool", d, [x, strip, ws, ws_bytes], reads=[x], writes=[strip] + ([ws] if ws else []),
                      tag=tag + ".pool", bytes=x.B * x.H * x.W * x.C * DT_SIZE[x.dtype]))
        return strip

    def _strip_attn(self, q: View, k: View, v: View, B, Lq, Lk, heads, scale, tag) -> View:
        out = self.strip_buf(B * Lq, q.C)
        dh = q.C // heads
        d = L.StripAttnDesc(B=B, heads=heads, dh=dh, Lq=Lq, Lk=Lk, scale=scale,
                            q_bstride=Lq * q.pitch, q_rstride=q.pitch, k_bstride=Lk * k.pitch, k_rstride=k.pitch,
                            v_bstride=Lk * v.pitch, v_rstride=v.pitch, o_bstride=Lq * out.pitch, o_rstride=out.pitch)
        self.add(Node("fce_strip_attn", d, [q, k, v, out], reads=[q, k, v], writes=[out], tag=tag))
        return out

    def _gate(self, x: View, gh: View, gw: View | None, mode, dst, tag) -> View:
        if dst is None:
            dst = self.new_buf(x.H, x.W, x.C)
        d = L.GateDesc(B=x.B, H=x.H, W=x.W, C=x.C, mode=mode, in_pitch=x.pitch, in_off=0, out_pitch=dst.pitch,
                       out_off=0, dtype=x.dtype, gh_bstride=x.H * gh.pitch, gh_rstride=gh.pitch,
                       gw_bstride=x.W * gw.pitch if gw is not None else 0, gw_rstride=gw.pitch if gw is not None else 0)
        self.add(Node("fce_gate_apply", d, [x, gh, gw, dst], reads=[x, gh] + ([gw] if gw is not None else []),
                      writes=[dst], tag=tag + ".gate", bytes=2.0 * x.B * x.H * x.W * x.C * DT_SIZE[x.dtype]))
        return dst

    def _identity(self, m, x: View, tag) -> View:
        idt = getattr(m, "identity", None)
        if idt is None or isinstance(idt, nn.Identity):
            return x
        return self.conv(idt, x, tag=tag + ".identity")

    def coord_att(self, m, x: View, dst=None, tag="") -> View:
        BH = x.B * x.H
        s = self._pool(x, tag)
        w1, b1, k1, _, _, act1 = self.conv_params(m.cv1)
        wh, bh, kh, *_ = self.conv_params(m.cv_h)
        ww, bw, kw, *_ = self.conv_params(m.cv_w)
        mip, oup = w1.shape[0], wh.shape[0]
        if (self.FUSED_COORDATT_MLP and (k1, kh, kw) == (1, 1, 1) and x.C % 4 == 0 and oup % 4 == 0 and x.C <= 1024
                and mip <= 64 and 4 * (mip * (x.C + oup) + oup + 16 * x.C + 1100) <= 200 * 1024):
            # cv1 (+BN, SiLU) -> cv_h / cv_w -> sigmoid on the strips in ONE launch (second layer transposed for the kernel)
            a = self.strip_buf(s.H, oup)
            d = L.CoordAttMlpDesc(rows_h=BH, rows_w=s.H - BH, C=x.C, mip=mip, oup=oup, s_pitch=s.pitch,
                                  out_pitch=a.pitch, act1=act1, act2=L.ACT_SIGMOID)
            ptrs = [s, self._w(w1.view(mip, x.C // 4, 4).permute(1, 0, 2)), self._w(b1), self._w(wh.view(oup, mip).t()), self._w(bh),
                    self._w(ww.view(oup, mip).t()), self._w(bw), a]
            self.add(Node("fce_coordatt_mlp", d, ptrs, reads=[s], writes=[a], tag=tag + ".mlp",
                          flops=2.0 * s.H * mip * (x.C + oup), bytes=4.0 * s.H * (x.C + oup)))
            a_h, a_w = a.rows(0, BH), a.rows(BH, a.H)
        else:
            y = self.conv(m.cv1, s, tag=tag + ".cv1")
            a_h = self.conv(m.cv_h, y.rows(0, BH), act=L.ACT_SIGMOID, tag=tag + ".cv_h")
            a_w = self.conv(m.cv_w, y.rows(BH, y.H), act=L.ACT_SIGMOID, tag=tag + ".cv_w")
        return self._gate(self._identity(m, x, tag), a_h, a_w, 0, dst, tag)

    def coord_cross_att(self, m, x: View, dst=None, tag="") -> View:
        BH = x.B * x.H
        mip, heads = m.mip, m.num_heads
        if mip % heads:
            raise PlanError(f"CoordCrossAtt: mip={mip} is not divisible by num_heads={heads} "
                            "(the reference fails in view(), fce_block.py:166)")
        if m.proj.out_channels != x.C:
            raise PlanError("CoordCrossAtt multiplies the raw input: needs oup == inp (fce_block.py:180)")
        s = self._pool(x, tag)
        y = self.conv(m.cv1, s, tag=tag + ".cv1")
        q = self.conv(m.q_conv, y.rows(0, BH), tag=tag + ".q")
        wk, bk, *_ = self.conv_params(m.k_conv)
        wv, bv, *_ = self.conv_params(m.v_conv)
        kv = self.conv(None, y.rows(BH, y.H), w_override=torch.cat([wk, wv]), b_override=torch.cat([bk, bv]),
                       tag=tag + ".kv")
        z = self._strip_attn(q, kv.ch(0, mip), kv.ch(mip, 2 * mip), x.B, x.H, x.W, heads, float(m.scale), tag + ".attn")
        g = self.conv(m.proj, z, act=L.ACT_SIGMOID, tag=tag + ".proj")
        return self._gate(x, g, None, 1, dst, tag)

    def bi_coord_cross_att(self, m, x: View, dst=None, tag="") -> View:
        BH = x.B * x.H
        mid, heads = m.mid_dim, m.num_heads
        s = self._pool(x, tag)

        def stacked(names):
            ws, bs = zip(*[self.conv_params(getattr(m, n))[:2] for n in names])
            return torch.cat(ws), torch.cat(bs)

        # rows from x_h feed q_h, k_w, v_w ; rows from x_w feed q_w, k_h, v_h (fce_block.py:246-268)
        w_h, b_h = stacked(["proj_q_h", "proj_k_w", "proj_v_w"])
        w_w, b_w = stacked(["proj_q_w", "proj_k_h", "proj_v_h"])
        ph = self.conv(None, s.rows(0, BH), w_override=w_h, b_override=b_h, tag=tag + ".proj_xh")
        pw = self.conv(None, s.rows(BH, s.H), w_override=w_w, b_override=b_w, tag=tag + ".proj_xw")
        zh = self._strip_attn(ph.ch(0, mid), pw.ch(mid, 2 * mid), pw.ch(2 * mid, 3 * mid), x.B, x.H, x.W, heads,
                              float(m.scale), tag + ".attn_h")
        zw = self._strip_attn(pw.ch(0, mid), ph.ch(mid, 2 * mid), ph.ch(2 * mid, 3 * mid), x.B, x.W, x.H, heads,
                              float(m.scale), tag + ".attn_w")
        gh = self.conv(m.out_h, zh, tag=tag + ".out_h")
        gw = self.conv(m.out_w, zw, tag=tag + ".out_w")
        return self._gate(self._identity(m, x, tag), gh, gw, 2, dst, tag)

    # ------------------------------------------------------------------ head
    def detect(self, m, xs, tag=""):
        if getattr(m, "end2end", False):
            raise PlanError("end2end heads are outside the path")
        nl, nc, R = m.nl, m.nc, m.reg_max
        if nl > 4:
            raise PlanError("at most 4 detection levels")
        xs = [self.materialize(x) for x in xs]
        strides = [float(s) for s in m.stride.tolist()]
        A = sum(x.H * x.W for x in xs)
        # Fused decode: the last 1x1 conv of each branch decodes in its tcgen05 epilogue and writes the prediction
        # tensor directly - no fp32 logit maps, no decode launch.  bf16 plans whose head fits fce_conv2d_detect.
        fused = (self.fuse_decode and self.act_dt == L.BF16 and self.impl != 1 and R == 16 and nc % 16 == 0
                 and all(m.cv2[i][2].in_channels % 16 == 0 and m.cv3[i][2].in_channels % 16 == 0 for i in range(nl)))
        raws, tails = [], []
        # Concurrent branches: the 2 * nl conv chains of the head are independent of each other and, for the finer
        # levels, of the rest of the neck - each gets its own branch of the captured graph (HEAD_STREAMS), so the big P3
        # chains run underneath the latency-bound small-map layers that finish the neck.
        ids = {}

        def branch(i, b):
            if not (self.branching and fused):
                return 0
            if (i, b) not in ids:
                self.n_branches += 1
                ids[(i, b)] = self.n_branches
            return ids[(i, b)]
        for i, x in enumerate(xs):
            raw = None if fused else self.new_buf(x.H, x.W, 4 * R + nc, dtype=L.F32, persistent=True)
            self.cur_stream = branch(i, 0)
            t = self.conv(m.cv2[i][0], x, tag=f"{tag}.cv2.{i}.0")
            t = self.conv(m.cv2[i][1], t, tag=f"{tag}.cv2.{i}.1")
            if fused:
                tails.append((m.cv2[i][2], t, 2, i, f"{tag}.cv2.{i}.2"))
            else:
                self.conv(m.cv2[i][2], t, dst=raw.ch(0, 4 * R), tag=f"{tag}.cv2.{i}.2")
            c = x
            self.cur_stream = branch(i, 1)
            for j in (0, 1):
                blk = m.cv3[i][j]
                if len(blk) != 2:
                    raise PlanError("legacy Detect heads (dense 3x3 class branch) are outside the path")
                c = self.dwpw(blk[0], blk[1], c, tag=f"{tag}.cv3.{i}.{j}")
            if fused:
                tails.append((m.cv3[i][2], c, 1, i, f"{tag}.cv3.{i}.2"))
            else:
                self.conv(m.cv3[i][2], c, dst=raw.ch(4 * R, 4 * R + nc), tag=f"{tag}.cv3.{i}.2")
                raws.append(raw)
            self.cur_stream = 0
        y = self.new_buf(1, 4 + nc, A, dtype=L.F32, persistent=True)  # logical [B, 4+nc, A]
        if fused:
            # the six y-writing convs go LAST: in overlap mode they form the short segment that waits for the previous
            # call's NMS to be done with y (engine.enable_overlap)
            a_base = [sum(v.H * v.W for v in xs[:i]) for i in range(nl)]
            for conv_m, t, mode, i, tg in tails:
                w, b, k, s_, g, a = self.conv_params(conv_m)
                if (k, s_, g, a) != (1, 1, 1, L.ACT_NONE):
                    raise PlanError(f"{tg}: the Detect branch must end in a bare 1x1 conv")
                Cout, Cin = w.shape[0], w.shape[1]
                d = L.ConvDesc(B=t.B, H=t.H, W=t.W, Cin=Cin, Cout=Cout, in_pitch=t.pitch, in_off=0, out_pitch=0, out_off=0,
                               res_pitch=0, res_off=0, k=1, stride=1, act=L.ACT_NONE, in_dtype=t.dtype, w_dtype=L.BF16,
                               out_dtype=L.F32, in_layout=L.NHWC, in_scale=1.0, impl=self.impl)
                e = L.DetectEpiDesc(mode=mode, A=A, a_base=a_base[i], rows=4 + nc, reg_max=R, stride=strides[i])
                wp = self._w(w.permute(0, 2, 3, 1), torch.bfloat16)
                out_rows = 4 if mode == 2 else nc
                self.cur_stream = branch(i, 0 if mode == 2 else 1)
                self.add(Node("fce_conv2d_detect", d, [e, t, wp, self._w(b), y], reads=[t], writes=[y], tag=tg,
                              flops=2.0 * t.B * t.H * t.W * Cout * Cin,
                              bytes=t.B * t.H * t.W * (Cin * 2.0 + out_rows * 4.0) + Cout * Cin * 2.0,
                              disjoint="detect-y"))  # each tail owns its rows / anchor range of y
                self.cur_stream = 0
            return y, []
        d = L.DecodeDesc(B=self.B, nl=nl, nc=nc, reg_max=R)
        for i, r in enumerate(raws):
            d.H[i], d.W[i], d.stride[i], d.raw_pitch[i] = r.H, r.W, strides[i], r.pitch
        ptrs = raws + [None] * (4 - nl) + [y]
        self.add(Node("fce_detect_decode", d, ptrs, reads=raws, writes=[y], tag=tag + ".decode",
                      bytes=self.B * A * ((4 * R + nc) + (4 + nc)) * 4.0))
        return y, raws

    def raw_buf(self, nbytes: int) -> View:
        """Untyped persistent byte buffer (NMS outputs / workspace)."""
        return self.new_buf(1, 1, int(nbytes), dtype=L.U8, B=1, persistent=True)

    def nms(self, y: View, nc: int, conf=0.25, iou=0.45, max_det=300, max_nms=30000, multi_label=False,
            agnostic=False, max_wh=7680.0, tag="nms"):
        """Appends the batched NMS (reference nms.py:13-166) so predict = forward + decode + NMS is ONE plan."""
        A = y.buf.C
        B = y.buf.B
        d = L.NmsDesc(B=B, A=A, nc=nc, conf_thres=float(conf), iou_thres=float(iou), max_det=int(max_det),
                      max_nms=int(max_nms), multi_label=int(bool(multi_label and nc > 1)), agnostic=int(bool(agnostic)),
                      max_wh=float(max_wh), n_classes=0)
        cap = A * nc if d.multi_label else A
        ws_bytes = B * cap * 16
        # det and count back to back in ONE buffer: the NMS kernel writes straight into the send buffer of the multi-GPU
        # gather (runner.DetectionGather: one collective, no packing copy)
        nd = B * max_det * 6 * 4
        detcount = self.raw_buf(nd + B * 4)
        det, count = detcount.ch(0, nd), detcount.ch(nd, nd + B * 4)
        keep = self.raw_buf(B * max_det * 8)
        ws = self.raw_buf(ws_bytes)
        self.add(Node("fce_nms", d, [y, None, det, keep, count, ws, ws_bytes], reads=[y],
                      writes=[det, keep, count, ws], tag=tag, bytes=B * ((4.0 + nc) * A * 4 + max_det * 6 * 4)))
        self.outputs.update({"det": det, "keep": keep, "count": count, "max_det": max_det, "detcount": detcount})
        return det, keep, count

    # ------------------------------------------------------------------ dispatch
    def emit(self, m, x, tag=""):
        """x: View | LazyUp | list of those.  Returns View | LazyUp | (y, raws) for Detect."""
        name = type(m).__name__
        if isinstance(x, list):  # only a single 1x1-conv consumer folds a pending BiFPN sum
            x = [self.materialize(e) if isinstance(e, LazySum) else e for e in x]
        if name == "Upsample":
            sf = m.scale_factor
            if m.mode != "nearest" or float(sf if not isinstance(sf, tuple) else sf[0]) != 2.0:
                raise PlanError("only nearest 2x upsampling is on the path")
            return LazyUp(self.materialize(x))
        if name == "BiFPN_Concat":
            return self.bifpn(m, list(x), tag=tag)
        if name == "Concat":
            return self.concat(m, list(x), tag=tag)
        if name == "Detect":
            return self.detect(m, list(x), tag=tag)
        if name == "C3k2" and isinstance(x, LazySum):
            return self.c3k2(m, x, tag=tag)
        x = self.materialize(x)
        if name in ("Conv", "Conv2d"):
            if name == "Conv" and m.conv.groups != 1:
                return self.dwconv(m, x, tag=tag)
            return self.conv(m, x, tag=tag)
        if name == "DWConv":
            return self.dwconv(m, x, tag=tag)
        fn = {"Bottleneck": self.bottleneck, "C3k": self.c3k, "C3k2": self.c3k2, "SPPF": self.sppf,
              "C2PSA": self.c2psa, "CoordAtt": self.coord_att, "CoordCrossAtt": self.coord_cross_att,
              "BiCoordCrossAtt": self.bi_coord_cross_att}.get(name)
        if fn is None:
            raise PlanError(f"module {name} is outside the FCE-YOLO hot path")
        return fn(m, x, tag=tag)


def compile_model(model, batch: int, height: int, width: int, precision: str, device, impl: int = 0,
                  input_u8: bool = False, nms: dict | None = None, fuse_decode: bool = False) -> Plan:
    """Whole-graph plan of a DetectionModel (mirror or reference): restates the routing of
    BaseModel._predict_once (tasks.py:172-188) at compile time."""
    layers = list(model.model)
    if height % 32 or width % 32:
        raise PlanError("image height/width must be multiples of 32 (reference loaders.py:603-609)")
    p = Plan(batch, precision, device, impl)
    p.fuse_decode = bool(fuse_decode)  # predict-only plans: Detect's raw logit maps (x_list, head.py:124) are not kept
    hs = Plan.HEAD_STREAMS
    p.branching = bool(batch * height * width <= Plan.HEAD_STREAMS_MAX_PIXELS if hs is None else hs) and p.fuse_decode
    first = layers[0]
    if type(first).__name__ != "Conv":
        raise PlanError("the graph must start with a Conv stem")
    cin = first.conv.in_channels
    if input_u8:
        img = p.new_buf(height, width, cin, dtype=L.U8, persistent=True)  # NHWC uint8
    else:
        img = p.new_buf(cin, height, width, dtype=L.F32, persistent=True)  # raw NCHW fp32 storage
    p.inputs = [img]
    ys = []
    x = None
    # layer 0's map is consumed by layer 1 alone (no later layer routes from it): the pair can run as one launch
    later_from0 = any(0 in ([f_] if isinstance(f_, int) else list(f_)) for f_ in (getattr(m_, "f", -1) for m_ in layers[2:]))
    fused01 = None
    for i, m in enumerate(layers):
        f = getattr(m, "f", -1)
        if i == 0:
            xin = View(img.buf, 0, cin, batch, height, width)  # for NCHW the geometry is carried by the desc
            if input_u8 and len(layers) > 1 and not later_from0:
                fused01 = p.stem2(m, layers[1], xin, 1.0 / 255.0, tag0="model.0", tag1="model.1")
            if fused01 is not None:
                x = None  # the stem map does not exist
            elif input_u8:
                x = p.stem(m, xin, L.NHWC, 1.0 / 255.0, tag="model.0")
            else:
                x = p.stem(m, xin, L.NCHW, 1.0, tag="model.0")
        elif i == 1 and fused01 is not None:
            x = fused01
        else:
            if f != -1:
                x = ys[f] if isinstance(f, int) else [x if j == -1 else ys[j] for j in f]
            x = p.emit(m, x, tag=f"model.{i}")
        ys.append(x)
        p.layer_out[i] = x
    last = ys[-1]
    if isinstance(last, tuple):
        p.outputs = {"y": last[0], "raw": last[1]}
    else:
        p.outputs = {"y": p.materialize(last)}
    for v in ([p.outputs["y"]] + p.outputs.get("raw", [])):
        v.buf.persistent = True
    if nms is not None:
        if "raw" not in p.outputs:
            raise PlanError("NMS needs a Detect head")
        p.nms(p.outputs["y"], layers[-1].nc, **nms)
    return p


def compile_module(m, in_shapes, precision: str, device, impl: int = 0) -> Plan:
    """Plan for ONE module (per-module drop-in and teacher-forced parity tests).
    in_shapes: list of (B, C, H, W)."""
    B = in_shapes[0][0]
    p = Plan(B, precision, device, impl)
    ins = [p.new_buf(h, w, c, persistent=True) for (_, c, h, w) in in_shapes]
    p.inputs = ins
    name = type(m).__name__
    multi = name in ("BiFPN_Concat", "Concat", "Detect")
    out = p.emit(m, list(ins) if multi else ins[0], tag=name)
    if isinstance(out, tuple):
        p.outputs = {"y": out[0], "raw": out[1]}
    else:
        p.outputs = {"y": p.materialize(out)}
    for v in ([p.outputs["y"]] + p.outputs.get("raw", [])):
        v.buf.persistent = True
    return p
