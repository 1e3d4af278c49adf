"""CPU oracle for the FCE-YOLOv11 detection forward path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``fce_yolo_b200/`` may import this file; only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs use it, and only as the checker / the CPU baseline - never as the product path.

What it is: a functional, state-dict driven restatement (plain torch fp32 ops on CPU) of the
reference forward.  It owns no nn.Module: every block is a function of ``(sd, prefix, x)``
where ``sd`` is the *fused* state dict (``...conv.weight`` / ``...conv.bias``) whose keys are
exactly the reference's (``model.<i>.<attr>...``).  Each function cites the reference lines
it restates (paths relative to /root/reference).

Pinning: the reference's own test-suite holds no numerical fixture for this path (its
tests/test_python.py:237-244 only constructs the YAMLs), so the oracle is pinned against
outputs of the reference itself, generated in the build container by
``tests/golden/make_golden.py`` (imports /root/reference/ultralytics, loads the same synthetic
state dict, stores per-layer outputs) and committed as ``tests/golden/*.npz``.
``tests/test_oracle_golden.py`` replays them.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# graph resolution: restates ultralytics/nn/tasks.py:1489-1743 (parse_model) for the module
# set used by yolo11{,-fce,-bifpn}.yaml
# --------------------------------------------------------------------------------------


def _ceil_to(x: float, d: int) -> int:
    """ultralytics/utils/ops.py:137-149 make_divisible."""
    return int(math.ceil(x / d) * d)


@dataclass
class Layer:
    i: int
    f: object  # int or list[int]
    kind: str
    args: list
    c_out: int
    extra: dict = field(default_factory=dict)


def resolve_graph(cfg: dict, scale: str | None = None, ch: int = 3) -> list[Layer]:
    """Resolve (from, repeats, module, args) rows into concrete constructor arguments.

    tasks.py:1501-1507 scale lookup; :1593 depth gain; :1595-1597 width scaling for "base"
    modules; :1599-1601 repeat insertion; :1606-1609 c3k forced for m/l/x; :1630-1708 FCE rows.
    """
    nc = cfg["nc"]
    scales = cfg.get("scales")
    scale = scale or cfg.get("scale") or next(iter(scales))
    depth, width, max_ch = scales[scale]
    rows = cfg["backbone"] + cfg["head"]
    chans: list[int] = []
    out: list[Layer] = []
    c_prev = ch
    for i, (f, n, kind, args) in enumerate(rows):
        args = [nc if a == "nc" else a for a in args]
        n = max(round(n * depth), 1) if n > 1 else n

        def cin(idx):
            return c_prev if (idx == -1 and not chans) else chans[idx]

        extra = {}
        if kind in ("Conv", "C3k2", "SPPF", "C2PSA"):
            c1 = cin(f)
            c2 = _ceil_to(min(args[0], max_ch) * width, 8)
            rest = list(args[1:])
            if kind in ("C3k2", "C2PSA"):
                rest.insert(0, n)  # repeats become a ctor argument (tasks.py:1599-1601)
            if kind == "C3k2":
                # args = [c1, c2, n, c3k, e]; c3k forced True for m/l/x (tasks.py:1608-1609)
                while len(rest) < 2:
                    rest.append(False)
                if scale in "mlx":
                    rest[1] = True
            a = [c1, c2, *rest]
        elif kind == "nn.Upsample":
            c2 = cin(f)
            a = list(args)
        elif kind == "Concat":
            c2 = sum(chans[j] for j in f)
            a = list(args)
        elif kind == "BiFPN_Concat":
            c1 = [chans[j] for j in f] if isinstance(f, list) else [chans[f]]
            c2 = args[0] if args else max(c1)
            c2 = _ceil_to(min(c2, max_ch) * width, 8)  # width applied a second time (tasks.py:1634)
            a = [c1, c2]
        elif kind in ("CoordAtt", "CoordCrossAtt", "BiCoordCrossAtt"):
            inp = cin(f)
            oup = _ceil_to(min(args[0], max_ch) * width, 8) if args else inp
            red = args[1] if len(args) > 1 else max(8, min(32, int(inp ** 0.5)))
            a = [inp, oup, red]
            if kind != "CoordAtt":
                if len(args) > 2:
                    heads = args[2]
                else:  # tasks.py:1670-1675 / :1700-1705
                    base = max(8, inp // red)
                    heads = max(1, min(8, inp // 32))
                    while heads > 1 and base // heads < 8:
                        heads -= 1
                a.append(heads)
            c2 = oup
        elif kind == "Detect":
            a = [args[0], [chans[j] for j in f]]
            c2 = None
        else:
            raise NotImplementedError(f"oracle: module {kind} is outside the hot path")
        out.append(Layer(i, f, kind, a, c2, extra))
        chans.append(c2)
    return out


# --------------------------------------------------------------------------------------
# blocks
# --------------------------------------------------------------------------------------


def conv_act(sd, p, x, act=True, groups=1, stride=None):
    """Fused Conv: act(conv2d(x) + b).  conv.py:80-89 (forward_fuse), autopad conv.py:30-36.

    ``p`` is the prefix of the ``Conv`` module, the fused nn.Conv2d lives at ``p + '.conv'``.
    Stride is not in the state dict - caller passes it (defaults to 1).
    """
    w = sd[p + ".conv.weight"]
    b = sd[p + ".conv.bias"]
    k = w.shape[-1]
    y = F.conv2d(x, w, b, stride=stride or 1, padding=k // 2, groups=groups)
    return F.silu(y) if act else y


def plain_conv(sd, p, x):
    """Bare biased nn.Conv2d 1x1 (fce_block.py:91-95,144-152,217-233; head.py:94,103)."""
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"])


def bottleneck(sd, p, x, shortcut=True):
    """block.py:452-476: x + cv2(cv1(x)) when shortcut and c1 == c2."""
    y = conv_act(sd, p + ".cv2", conv_act(sd, p + ".cv1", x))
    return x + y if (shortcut and y.shape[1] == x.shape[1]) else y


def c3k(sd, p, x, n=2, shortcut=True):
    """block.py:317-340 (C3.forward) with block.py:1087-1105 (two 3x3 Bottlenecks, e=1.0)."""
    a = conv_act(sd, p + ".cv1", x)
    for j in range(n):
        a = bottleneck(sd, f"{p}.m.{j}", a, shortcut)
    return conv_act(sd, p + ".cv3", torch.cat((a, conv_act(sd, p + ".cv2", x)), 1))


def c3k2(sd, p, x, n, use_c3k, shortcut=True):
    """block.py:303-307 (C2f.forward) with block.py:1064-1084 inner blocks."""
    y = list(conv_act(sd, p + ".cv1", x).chunk(2, 1))
    for j in range(n):
        q = f"{p}.m.{j}"
        y.append(c3k(sd, q, y[-1], 2, shortcut) if use_c3k else bottleneck(sd, q, y[-1], shortcut))
    return conv_act(sd, p + ".cv2", torch.cat(y, 1))


def sppf(sd, p, x, k=5):
    """block.py:228-232: cv2(cat[y, m(y), m(m(y)), m(m(m(y)))]), m = MaxPool2d(k, 1, k//2)."""
    y = [conv_act(sd, p + ".cv1", x)]
    for _ in range(3):
        y.append(F.max_pool2d(y[-1], k, 1, k // 2))
    return conv_act(sd, p + ".cv2", torch.cat(y, 1))


def psa_attention(sd, p, x, num_heads):
    """block.py:1284-1304 Attention.forward (attn_ratio 0.5)."""
    B, C, H, W = x.shape
    N = H * W
    hd = C // num_heads
    kd = int(hd * 0.5)
    qkv = conv_act(sd, p + ".qkv", x, act=False)
    q, k, v = qkv.view(B, num_heads, 2 * kd + hd, N).split([kd, kd, hd], dim=2)
    attn = ((q.transpose(-2, -1) @ k) * (kd ** -0.5)).softmax(dim=-1)
    o = (v @ attn.transpose(-2, -1)).view(B, C, H, W)
    o = o + conv_act(sd, p + ".pe", v.reshape(B, C, H, W), act=False, groups=C)
    return conv_act(sd, p + ".proj", o, act=False)


def c2psa(sd, p, x, n):
    """block.py:1453-1464 C2PSA.forward; PSABlock block.py:1343-1354; heads = c // 64 (:1451)."""
    c = x.shape[1] // 2
    a, b = conv_act(sd, p + ".cv1", x).split((c, c), 1)
    for j in range(n):
        q = f"{p}.m.{j}"
        b = b + psa_attention(sd, q + ".attn", b, c // 64)
        b = b + conv_act(sd, q + ".ffn.1", conv_act(sd, q + ".ffn.0", b), act=False)
    return conv_act(sd, p + ".cv2", torch.cat((a, b), 1))


def bifpn_concat(sd, p, xs, eps=1e-4):
    """fce_block.py:40-63: sum_i relu(w_i)/(sum relu(w)+eps) * realign_i(x_i)."""
    ys = []
    for j, t in enumerate(xs):
        key = f"{p}.realign_convs.{j}"
        ys.append(conv_act(sd, key, t) if (key + ".conv.weight") in sd else t)
    w = torch.relu(sd[p + ".w"])
    w = w / (w.sum() + eps)
    out = w[0] * ys[0]
    for j in range(1, len(ys)):
        out = out + w[j] * ys[j]
    return out


def _strip_means(x):
    """AdaptiveAvgPool2d((None,1)) and ((1,None)): fce_block.py:81-82,140-141,212-213."""
    return x.mean(dim=3, keepdim=True), x.mean(dim=2, keepdim=True)  # [B,C,H,1], [B,C,1,W]


def coord_att(sd, p, x):
    """fce_block.py:97-116."""
    B, C, H, W = x.shape
    xh, xw = _strip_means(x)
    y = conv_act(sd, p + ".cv1", torch.cat([xh, xw.permute(0, 1, 3, 2)], dim=2))
    yh, yw = torch.split(y, [H, W], dim=2)
    a_h = plain_conv(sd, p + ".cv_h", yh).sigmoid()
    a_w = plain_conv(sd, p + ".cv_w", yw.permute(0, 1, 3, 2)).sigmoid()
    idt = plain_conv(sd, p + ".identity", x) if (p + ".identity.weight") in sd else x
    return idt * a_h * a_w


def coord_cross_att(sd, p, x, heads):
    """fce_block.py:155-180.  Gate is H-only and multiplies raw x."""
    B, C, H, W = x.shape
    xh, xw = _strip_means(x)
    y = plain_conv(sd, p + ".cv1", torch.cat([xh, xw.permute(0, 1, 3, 2)], dim=2))
    mip = y.shape[1]
    yh, yw = torch.split(y, [H, W], dim=2)
    q = plain_conv(sd, p + ".q_conv", yh).view(B, heads, -1, H).permute(0, 1, 3, 2)
    k = plain_conv(sd, p + ".k_conv", yw).view(B, heads, -1, W)
    v = plain_conv(sd, p + ".v_conv", yw).view(B, heads, -1, W).permute(0, 1, 3, 2)
    attn = ((q @ k) * ((mip // heads) ** -0.5)).softmax(dim=-1)
    z = (attn @ v).permute(0, 1, 3, 2).contiguous().view(B, mip, H, 1)
    return x * plain_conv(sd, p + ".proj", z).sigmoid()


def bi_coord_cross_att(sd, p, x, heads):
    """fce_block.py:235-284.  Additive gate: sigmoid(g_h[H,1] + g_w[1,W])."""
    B, C, H, W = x.shape
    xh, xw = _strip_means(x)
    mid = sd[p + ".proj_q_h.weight"].shape[0]
    d = mid // heads
    sc = d ** -0.5

    def branch(sfx, src_q, src_kv, Lq, Lk):
        q = plain_conv(sd, f"{p}.proj_q_{sfx}", src_q).view(B, heads, d, Lq).permute(0, 1, 3, 2)
        k = plain_conv(sd, f"{p}.proj_k_{sfx}", src_kv).view(B, heads, d, Lk)
        v = plain_conv(sd, f"{p}.proj_v_{sfx}", src_kv).view(B, heads, d, Lk).permute(0, 1, 3, 2)
        a = ((q @ k) * sc).softmax(dim=-1)
        return (a @ v).permute(0, 1, 3, 2)  # [B, heads, d, Lq]

    g_h = plain_conv(sd, p + ".out_h", branch("h", xh, xw, H, W).reshape(B, mid, H, 1))
    g_w = plain_conv(sd, p + ".out_w", branch("w", xw, xh, W, H).reshape(B, mid, 1, W))
    idt = plain_conv(sd, p + ".identity", x) if (p + ".identity.weight") in sd else x
    return idt * torch.sigmoid(g_h + g_w)


def detect_raw(sd, p, feats):
    """head.py:114-120: per level cat(cv2[i](x), cv3[i](x)) -> [B, 64+nc, Hi, Wi]."""
    outs = []
    for i, x in enumerate(feats):
        b = conv_act(sd, f"{p}.cv2.{i}.1", conv_act(sd, f"{p}.cv2.{i}.0", x))
        b = plain_conv(sd, f"{p}.cv2.{i}.2", b)
        c = x
        for j in (0, 1):  # legacy=False head: [DWConv3x3, Conv1x1] twice (head.py:97-107)
            c = conv_act(sd, f"{p}.cv3.{i}.{j}.0", c, groups=c.shape[1])
            c = conv_act(sd, f"{p}.cv3.{i}.{j}.1", c)
        c = plain_conv(sd, f"{p}.cv3.{i}.2", c)
        outs.append(torch.cat((b, c), 1))
    return outs


def detect_decode(raw, strides, reg_max=16):
    """head.py:149-167 (_inference) + block.py:76-79 (DFL) + tal.py:352-376."""
    B = raw[0].shape[0]
    no = raw[0].shape[1]
    x_cat = torch.cat([r.reshape(B, no, -1) for r in raw], 2)
    box, cls = x_cat.split((reg_max * 4, no - reg_max * 4), 1)
    A = box.shape[-1]
    prob = box.view(B, 4, reg_max, A).softmax(2)
    dev = prob.device  # device-agnostic so that tests/torch_gpu_baseline.py can time these very ops through cuDNN
    dist = (prob * torch.arange(reg_max, dtype=prob.dtype, device=dev).view(1, 1, reg_max, 1)).sum(2)  # [B,4,A]
    pts, st = [], []
    for r, s in zip(raw, strides):
        h, w = r.shape[2:]
        sy, sx = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=dev) + 0.5,
                                torch.arange(w, dtype=torch.float32, device=dev) + 0.5, indexing="ij")
        pts.append(torch.stack((sx, sy), -1).view(-1, 2))
        st.append(torch.full((h * w,), float(s), device=dev))
    anchors = torch.cat(pts).t().unsqueeze(0)  # [1,2,A]
    st = torch.cat(st).view(1, 1, A)
    lt, rb = dist.chunk(2, 1)
    x1y1 = anchors - lt
    x2y2 = anchors + rb
    dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * st
    return torch.cat((dbox, cls.sigmoid()), 1)


# --------------------------------------------------------------------------------------
# whole graph: restates tasks.py:160-188 (_predict_once)
# --------------------------------------------------------------------------------------


def run_layer(L: Layer, sd, x, strides=(8, 16, 32)):
    p = f"model.{L.i}"
    a = L.args
    if L.kind == "Conv":
        return conv_act(sd, p, x, stride=a[3] if len(a) > 3 else 1)
    if L.kind == "C3k2":
        return c3k2(sd, p, x, a[2], a[3])
    if L.kind == "SPPF":
        return sppf(sd, p, x, a[2])
    if L.kind == "C2PSA":
        return c2psa(sd, p, x, a[2])
    if L.kind == "nn.Upsample":
        return F.interpolate(x, scale_factor=a[1], mode=a[2])
    if L.kind == "Concat":
        return torch.cat(x, a[0])
    if L.kind == "BiFPN_Concat":
        return bifpn_concat(sd, p, x)
    if L.kind == "CoordAtt":
        return coord_att(sd, p, x)
    if L.kind == "CoordCrossAtt":
        return coord_cross_att(sd, p, x, a[3])
    if L.kind == "BiCoordCrossAtt":
        return bi_coord_cross_att(sd, p, x, a[3])
    if L.kind == "Detect":
        raw = detect_raw(sd, p, x)
        return detect_decode(raw, strides), raw
    raise NotImplementedError(L.kind)


@torch.no_grad()
def forward(cfg: dict, scale: str, sd: dict, img: torch.Tensor, keep_layers: bool = False):
    """Returns (y[B,4+nc,A], raw list) and, if asked, the list of every layer's output."""
    graph = resolve_graph(cfg, scale)
    ys = []
    x = img
    for L in graph:
        if L.f != -1:
            x = ys[L.f] if isinstance(L.f, int) else [x if j == -1 else ys[j] for j in L.f]
        x = run_layer(L, sd, x)
        ys.append(x)
    return (ys[-1], ys) if keep_layers else ys[-1]


def layer_inputs(graph, ys, img, i):
    """Teacher-forcing helper: the input(s) layer i saw in a recorded run."""
    L = graph[i]
    prev = img if i == 0 else ys[i - 1]
    if L.f == -1:
        return prev
    if isinstance(L.f, int):
        return ys[L.f]
    return [prev if j == -1 else ys[j] for j in L.f]
