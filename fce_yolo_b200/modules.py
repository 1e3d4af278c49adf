"""Host-side mirror of the reference operator surface.

Same class names, positional constructor signatures, attribute names and therefore the same
``state_dict`` keys as the reference's ``ultralytics.nn.modules`` (conv.py, block.py, head.py,
fce_block.py) for the modules yolo11{,-fce,-bifpn}.yaml use - so a reference checkpoint's state
dict loads with ``strict=True`` and the plan compiler (plan.py) can walk either these classes
or the reference's own instances.

These modules are *parameter containers*: there is no PyTorch math here.  ``forward`` hands the
module to the engine, which compiles it to a static plan of sm_100a kernels and runs it on the
tensor's CUDA stream.  On a CPU tensor, or if the CUDA library is missing, it raises - there is
deliberately no eager fallback.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

__all__ = (
    "Conv", "DWConv", "Concat", "Bottleneck", "C3k", "C3k2", "SPPF", "Attention", "PSABlock", "C2PSA",
    "DFL", "Detect", "BiFPN_Concat", "CoordAtt", "CoordCrossAtt", "BiCoordCrossAtt", "fuse_module",
)


def _same_pad(k: int, p=None, d: int = 1) -> int:
    """'same' padding for odd kernels (reference conv.py:30-36)."""
    k = d * (k - 1) + 1
    return k // 2 if p is None else p


class _PlanModule(nn.Module):
    """Base: forward = compile-to-plan + run on CUDA (engine.run_module)."""

    def forward(self, x):
        from .engine import run_module

        return run_module(self, x)


class Conv(_PlanModule):
    """conv2d(bias=False) + BatchNorm2d + SiLU; BN is folded by ``fuse_module`` (conv.py:39-89)."""

    def __init__(self, c1, c2, k=1, s=1, p=None, g=1, d=1, act=True):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, _same_pad(k, p, d), groups=g, dilation=d, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)  # torch_utils.py:468-470
        self.act = nn.SiLU() if act is True else act if isinstance(act, nn.Module) else nn.Identity()


class DWConv(Conv):
    """Depth-wise Conv, groups = gcd(c1, c2) (conv.py:185-199)."""

    def __init__(self, c1, c2, k=1, s=1, d=1, act=True):
        super().__init__(c1, c2, k, s, g=math.gcd(c1, c2), d=d, act=act)


class Concat(_PlanModule):
    def __init__(self, dimension=1):
        super().__init__()
        self.d = dimension


class Bottleneck(_PlanModule):
    def __init__(self, c1, c2, shortcut=True, g=1, k=(3, 3), e=0.5):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, k[0], 1)
        self.cv2 = Conv(c_, c2, k[1], 1, g=g)
        self.add = shortcut and c1 == c2


class C3k(_PlanModule):
    """C3 with two kxk Bottlenecks (block.py:317-340, 1087-1105)."""

    def __init__(self, c1, c2, n=1, shortcut=True, g=1, e=0.5, k=3):
        super().__init__()
        c_ = int(c2 * e)
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c1, c_, 1, 1)
        self.cv3 = Conv(2 * c_, c2, 1)
        self.m = nn.Sequential(*(Bottleneck(c_, c_, shortcut, g, k=(k, k), e=1.0) for _ in range(n)))


class C3k2(_PlanModule):
    """C2f layout with Bottleneck or C3k inner blocks (block.py:283-307, 1064-1084)."""

    def __init__(self, c1, c2, n=1, c3k=False, e=0.5, g=1, shortcut=True):
        super().__init__()
        self.c = int(c2 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1)
        self.m = nn.ModuleList(
            C3k(self.c, self.c, 2, shortcut, g) if c3k else Bottleneck(self.c, self.c, shortcut, g) for _ in range(n)
        )


class SPPF(_PlanModule):
    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * 4, c2, 1, 1)
        self.m = nn.MaxPool2d(kernel_size=k, stride=1, padding=k // 2)


class Attention(_PlanModule):
    """block.py:1247-1304."""

    def __init__(self, dim, num_heads=8, attn_ratio=0.5):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.key_dim = int(self.head_dim * attn_ratio)
        self.scale = self.key_dim ** -0.5
        self.qkv = Conv(dim, dim + self.key_dim * num_heads * 2, 1, act=False)
        self.proj = Conv(dim, dim, 1, act=False)
        self.pe = Conv(dim, dim, 3, 1, g=dim, act=False)


class PSABlock(_PlanModule):
    def __init__(self, c, attn_ratio=0.5, num_heads=4, shortcut=True):
        super().__init__()
        self.attn = Attention(c, attn_ratio=attn_ratio, num_heads=num_heads)
        self.ffn = nn.Sequential(Conv(c, c * 2, 1), Conv(c * 2, c, 1, act=False))
        self.add = shortcut


class C2PSA(_PlanModule):
    def __init__(self, c1, c2, n=1, e=0.5):
        super().__init__()
        assert c1 == c2
        self.c = int(c1 * e)
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv(2 * self.c, c1, 1)
        self.m = nn.Sequential(*(PSABlock(self.c, attn_ratio=0.5, num_heads=self.c // 64) for _ in range(n)))


class DFL(nn.Module):
    """Holds the arange(16) projection so state dicts match (block.py:58-79); the decode kernel
    computes the softmax-integral directly and never reads this weight."""

    def __init__(self, c1=16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(c1, dtype=torch.float).view(1, c1, 1, 1)
        self.c1 = c1


class Detect(_PlanModule):
    """head.py:26-212, legacy=False layout (depth-wise class branch)."""

    dynamic = False
    export = False
    end2end = False
    max_det = 300
    shape = None
    legacy = False
    xyxy = False

    def __init__(self, nc=80, ch=()):
        super().__init__()
        self.nc = nc
        self.nl = len(ch)
        self.reg_max = 16
        self.no = nc + self.reg_max * 4
        self.stride = torch.zeros(self.nl)
        c2, c3 = max((16, ch[0] // 4, self.reg_max * 4)), max(ch[0], min(self.nc, 100))
        self.cv2 = nn.ModuleList(
            nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * self.reg_max, 1)) for x in ch
        )
        self.cv3 = nn.ModuleList(
            nn.Sequential(
                nn.Sequential(DWConv(x, x, 3), Conv(x, c3, 1)),
                nn.Sequential(DWConv(c3, c3, 3), Conv(c3, c3, 1)),
                nn.Conv2d(c3, self.nc, 1),
            )
            for x in ch
        )
        self.dfl = DFL(self.reg_max)

    def bias_init(self):
        """head.py:169-180."""
        for a, b, s in zip(self.cv2, self.cv3, self.stride):
            a[-1].bias.data[:] = 1.0
            b[-1].bias.data[: self.nc] = math.log(5 / self.nc / (640 / s) ** 2)


class BiFPN_Concat(_PlanModule):
    """Fast-normalised weighted fusion with 1x1 realign convs (fce_block.py:13-63)."""

    def __init__(self, c1, c2=None):
        super().__init__()
        self.output_ch = c2 if c2 else max(c1)
        self.realign_convs = nn.ModuleList(
            Conv(ch, self.output_ch, 1, 1) if ch != self.output_ch else nn.Identity() for ch in c1
        )
        self.w = nn.Parameter(torch.ones(len(c1), dtype=torch.float32), requires_grad=True)
        self.epsilon = 1e-4


class CoordAtt(_PlanModule):
    """fce_block.py:65-116."""

    def __init__(self, inp, oup, reduction=32):
        super().__init__()
        mip = max(8, inp // reduction)
        self.cv1 = Conv(inp, mip, k=1, s=1, p=0)
        self.cv_h = nn.Conv2d(mip, oup, kernel_size=1, stride=1, padding=0)
        self.cv_w = nn.Conv2d(mip, oup, kernel_size=1, stride=1, padding=0)
        self.identity = nn.Conv2d(inp, oup, 1) if inp != oup else nn.Identity()


class CoordCrossAtt(_PlanModule):
    """fce_block.py:119-180."""

    def __init__(self, inp, oup, reduction=32, num_heads=1):
        super().__init__()
        self.mip = max(8, inp // reduction)
        self.num_heads = num_heads
        self.scale = (self.mip // num_heads) ** -0.5
        if self.mip % num_heads:
            # the reference dies in forward with a view() RuntimeError (fce_block.py:166); fail at build
            raise RuntimeError(f"CoordCrossAtt: mip={self.mip} is not divisible by num_heads={num_heads}")
        self.cv1 = nn.Conv2d(inp, self.mip, kernel_size=1)
        self.q_conv = nn.Conv2d(self.mip, self.mip, 1)
        self.k_conv = nn.Conv2d(self.mip, self.mip, 1)
        self.v_conv = nn.Conv2d(self.mip, self.mip, 1)
        self.proj = nn.Conv2d(self.mip, oup, 1)


class BiCoordCrossAtt(_PlanModule):
    """fce_block.py:183-284."""

    def __init__(self, inp, oup, reduction=32, num_heads=4):
        super().__init__()
        self.num_heads = num_heads
        self.dim_head = max(8, inp // reduction) // num_heads
        self.mid_dim = self.dim_head * num_heads
        self.scale = self.dim_head ** -0.5
        for br in ("h", "w"):
            for nm in ("q", "k", "v"):
                setattr(self, f"proj_{nm}_{br}", nn.Conv2d(inp, self.mid_dim, 1))
            setattr(self, f"out_{br}", nn.Conv2d(self.mid_dim, oup, 1))
        self.identity = nn.Conv2d(inp, oup, 1) if inp != oup else nn.Identity()


@torch.no_grad()
def fold_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d):
    """W' = diag(g/sqrt(var+eps)) W ; b' = beta - g*mean/sqrt(var+eps) (+ scaled conv bias).
    Same algebra as the reference's fuse_conv_and_bn (torch_utils.py:237-267), returned as tensors."""
    s = bn.weight / torch.sqrt(bn.running_var + bn.eps)
    w = conv.weight * s.view(-1, 1, 1, 1)
    b0 = conv.bias if conv.bias is not None else torch.zeros_like(bn.running_mean)
    return w, bn.bias + (b0 - bn.running_mean) * s


@torch.no_grad()
def fuse_module(root: nn.Module) -> nn.Module:
    """In-place BN folding for every Conv-like child (mirror of BaseModel.fuse, tasks.py:223-252)."""
    for m in root.modules():
        bn = getattr(m, "bn", None)
        if isinstance(bn, nn.BatchNorm2d) and isinstance(getattr(m, "conv", None), nn.Conv2d):
            w, b = fold_bn(m.conv, bn)
            m.conv.weight.data = w
            m.conv.register_parameter("bias", nn.Parameter(b, requires_grad=False))
            m.conv.requires_grad_(False)
            del m.bn
    return root
