"""TEST-ONLY interpreter of a compiled Plan with torch ops on the CPU.

It executes the *same node list, descriptors, view offsets and packed weights* the CUDA executor would
launch, so the plan compiler (fusions, concat-by-offset, weight re-ordering, strides) can be validated
in the GPU-less build container.  It says nothing about the CUDA kernels - those are checked on the
B200 (tests marked gpu).  Never imported by the product.
"""
import torch
import torch.nn.functional as F

from fce_yolo_b200 import _lib as L
from fce_yolo_b200.engine import Arena
from fce_yolo_b200.plan import View


def _act(t, a):
    return F.silu(t) if a == L.ACT_SILU else torch.sigmoid(t) if a == L.ACT_SIGMOID else t


class Interp(Arena):
    def t(self, v):
        return self.tensor(v)

    def run(self):
        for n in self.plan.nodes:
            getattr(self, "_" + n.fn)(n.desc, n.ptrs)

    def _fce_conv2d(self, d, p):
        x, w, b, res, y = p
        if d.in_layout == L.NCHW:
            buf = x.buf
            xin = self._flat(buf).view(buf.B, buf.H, buf.W, buf.C).float()  # is [B,C,H,W]
        else:
            xin = View(x.buf, x.c0, x.C, d.B, d.H, d.W, x.row0)
            xin = self.t(xin).permute(0, 3, 1, 2).float()
        xin = xin * d.in_scale
        wt = w.float().permute(0, 3, 1, 2)  # OHWI -> OIHW
        o = F.conv2d(xin, wt, b.float(), stride=d.stride, padding=d.k // 2)
        r = None
        if res is not None:
            r = self.t(res).permute(0, 3, 1, 2).float()
            if d.res_up:
                r = F.interpolate(r, scale_factor=2, mode="nearest")
        if d.weighted == 2:  # the residual joins before the activation (BiFPN node folded into its consumer)
            o = _act(o + r, d.act)
        else:
            o = _act(o, d.act) * (d.out_scale if d.weighted else 1.0)
            if r is not None:
                o = o + r * (d.res_scale if d.weighted else 1.0)
        self.t(y).copy_(o.permute(0, 2, 3, 1))

    def _fce_stem_pack(self, d, p):
        x, a = p
        buf = x.buf
        flat = self._flat(buf)
        if d.in_layout == L.NCHW:
            xin = flat.view(d.B, 3, d.H, d.W).float()
        else:
            xin = flat.view(d.B, d.H, d.W, 3).permute(0, 3, 1, 2).float()
        cols = F.unfold(xin, kernel_size=3, padding=1, stride=2)  # [B, ci*9 + kh*3 + kw, L]
        Bn, _, Ln = cols.shape
        cols = cols.view(Bn, 3, 9, Ln).permute(0, 3, 2, 1).reshape(Bn, Ln, 27)  # K index (kh*3+kw)*3+ci
        out = torch.zeros(Bn, Ln, 32)
        out[..., :27] = cols
        self.t(a).copy_(out.view(self.t(a).shape))

    def _fce_stem_conv(self, d, p):
        x, w, b, y = p
        flat = self._flat(x.buf)
        if d.in_layout == L.NCHW:
            xin = flat.view(d.B, 3, d.H, d.W).float()
        else:
            xin = flat.view(d.B, d.H, d.W, 3).permute(0, 3, 1, 2).float()
        wt = w.float()[:, :27].reshape(d.Cout, 3, 3, 3).permute(0, 3, 1, 2)  # [co][kh][kw][ci] -> OIHW
        o = _act(F.conv2d(xin, wt, b.float(), stride=2, padding=1), d.act)
        self.t(y).copy_(o.permute(0, 2, 3, 1))

    def _fce_stem2_conv(self, d, p):
        x, w0, b0, w1, b1, y = p
        xin = self._flat(x.buf).view(d.B, d.H, d.W, 3).permute(0, 3, 1, 2).float()
        wt0 = w0.float()[:, :27].reshape(d.C0, 3, 3, 3).permute(0, 3, 1, 2)
        t = _act(F.conv2d(xin, wt0, b0.float(), stride=2, padding=1), d.act0).to(torch.bfloat16).float()
        o = _act(F.conv2d(t, w1.float().permute(0, 3, 1, 2), b1.float(), stride=2, padding=1), d.act1)
        self.t(y).copy_(o.permute(0, 2, 3, 1))

    def _fce_dwconv3x3(self, d, p):
        x, w, b, add, y = p
        xin = self.t(x).permute(0, 3, 1, 2).float()
        wt = w.float().t().reshape(d.C, 1, 3, 3)
        o = _act(F.conv2d(xin, wt, b.float(), padding=1, groups=d.C), d.act)
        if add is not None:
            o = o + self.t(add).permute(0, 3, 1, 2).float()
        self.t(y).copy_(o.permute(0, 2, 3, 1))

    def _fce_dwpw_conv(self, d, p):
        x, wd, bd, wp, bp, y = p
        xin = self.t(x).permute(0, 3, 1, 2).float()
        t = _act(F.conv2d(xin, wd.float().t().reshape(d.C, 1, 3, 3), bd.float(), padding=1, groups=d.C), d.dw_act)
        t = t.to(self.t(y).dtype).float()  # the intermediate is rounded to the activation dtype, as the two launches store it
        o = _act(F.conv2d(t, wp.float().view(d.Cout, d.C, 1, 1), bp.float()), d.pw_act)
        self.t(y).copy_(o.permute(0, 2, 3, 1))

    def _fce_conv1x1_chain(self, d, p):
        x1, w1, b1, x2, w2, b2, y = p
        a1 = self.t(x1).permute(0, 3, 1, 2).float()
        t = _act(F.conv2d(a1, w1.float().view(d.cm, d.c1, 1, 1), b1.float()), d.act1)
        t = t.to(self.t(y).dtype).float()  # the intermediate is rounded to the activation dtype, as the two launches store it
        a2 = torch.cat([self.t(x2).permute(0, 3, 1, 2).float(), t], 1)
        o = _act(F.conv2d(a2, w2.float().view(d.Cout, d.c2 + d.cm, 1, 1), b2.float()), d.act2)
        self.t(y).copy_(o.permute(0, 2, 3, 1))

    def _fce_sppf_pool(self, d, p):
        cat = p[0]
        cur = self.t(cat.ch(0, d.C)).permute(0, 3, 1, 2).float()
        for j in range(1, 4):
            cur = F.max_pool2d(cur, 5, 1, 2)
            self.t(cat.ch(j * d.C, (j + 1) * d.C)).copy_(cur.permute(0, 2, 3, 1))

    def _fce_upsample2x(self, d, p):
        x, y = p
        o = F.interpolate(self.t(x).permute(0, 3, 1, 2).float(), scale_factor=2, mode="nearest")
        self.t(y).copy_(o.permute(0, 2, 3, 1))

    def _fce_bifpn_fuse(self, d, p):
        acc = None
        for i in range(d.n):
            t = self.t(p[i]).permute(0, 3, 1, 2).float()
            if d.up[i]:
                t = F.interpolate(t, scale_factor=2, mode="nearest")
            acc = d.wn[i] * t if acc is None else acc + d.wn[i] * t
        self.t(p[3]).copy_(acc.permute(0, 2, 3, 1))

    def _fce_copy_view(self, d, p):
        self.t(p[1]).copy_(self.t(p[0]))

    def _fce_coord_pool(self, d, p):
        x, strip = p[0], p[1]
        t = self.t(x).float()  # [B,H,W,C]
        s = self.t(strip).view(-1, d.C)
        s[: d.B * d.H] = t.mean(2).reshape(-1, d.C)
        s[d.B * d.H:] = t.mean(1).reshape(-1, d.C)

    def _fce_coordatt_mlp(self, d, p):
        s = self.t(p[0]).reshape(-1, d.C)
        w1, b1, wht, bh, wwt, bw = [z.float().cpu() for z in p[1:7]]
        act = {0: lambda v: v, 1: torch.nn.functional.silu, 2: torch.sigmoid}
        w1 = w1.reshape(d.C // 4, d.mip, 4).permute(1, 0, 2).reshape(d.mip, d.C)  # undo the kernel's regrouping
        y = act[d.act1](s @ w1.t() + b1)
        o = self.t(p[7]).reshape(-1, d.oup)
        o[: d.rows_h] = act[d.act2](y[: d.rows_h] @ wht + bh)
        o[d.rows_h:] = act[d.act2](y[d.rows_h:] @ wwt + bw)

    def _fce_strip_attn(self, d, p):
        q, k, v, o = [self.t(z).reshape(-1, z.C) for z in p]
        B, h, dh = d.B, d.heads, d.dh
        q = q.view(B, d.Lq, h, dh).permute(0, 2, 1, 3)
        k = k.view(B, d.Lk, h, dh).permute(0, 2, 3, 1)
        v = v.view(B, d.Lk, h, dh).permute(0, 2, 1, 3)
        a = ((q @ k) * d.scale).softmax(-1)
        z = (a @ v).permute(0, 2, 1, 3).reshape(B * d.Lq, h * dh)
        self.t(p[3]).reshape(-1, h * dh).copy_(z) if self.t(p[3]).is_contiguous() else self.t(p[3]).copy_(
            z.view(1, B * d.Lq, 1, h * dh))

    def _fce_gate_apply(self, d, p):
        x, gh, gw, y = p
        t = self.t(x).float()
        a = self.t(gh).reshape(d.B, d.H, 1, d.C)
        if d.mode == 1:
            o = t * a
        else:
            b = self.t(gw).reshape(d.B, 1, d.W, d.C)
            o = t * a * b if d.mode == 0 else t * torch.sigmoid(a + b)
        self.t(y).copy_(o)

    def _fce_psa_attention(self, d, p):
        qkv, out = p
        t = self.t(qkv).float().reshape(d.B, d.N, -1)
        nh, kd, hd = d.heads, d.kd, d.hd
        q = t[..., d.q_off:d.q_off + nh * kd].view(d.B, d.N, nh, kd).permute(0, 2, 1, 3)
        k = t[..., d.k_off:d.k_off + nh * kd].view(d.B, d.N, nh, kd).permute(0, 2, 3, 1)
        v = t[..., d.v_off:d.v_off + nh * hd].view(d.B, d.N, nh, hd).permute(0, 2, 1, 3)
        a = ((q @ k) * d.scale).softmax(-1)
        o = (a @ v).permute(0, 2, 1, 3).reshape(d.B, d.N, nh * hd)
        self.t(out).copy_(o.view(self.t(out).shape))

    def _fce_conv2d_detect(self, d, p):
        """1x1 conv + the slice of Detect._inference this branch owns, written into y [B, 4+nc, A]."""
        from oracle.fce_oracle import detect_decode

        e, x, w, b, yv = p
        xin = self.t(View(x.buf, x.c0, x.C, d.B, d.H, d.W, x.row0)).permute(0, 3, 1, 2).float()
        o = F.conv2d(xin, w.float().permute(0, 3, 1, 2), b.float())  # [B, Cout, H, W]
        yb = yv.buf
        y = self._flat(yb).view(yb.B, yb.W, yb.C)
        sl = slice(e.a_base, e.a_base + d.H * d.W)
        if e.mode == 1:
            y[:, 4:4 + d.Cout, sl] = torch.sigmoid(o.reshape(d.B, d.Cout, -1))
        else:
            full = torch.cat([o, torch.zeros(d.B, 1, d.H, d.W)], 1)  # dummy class row for the oracle's splitter
            y[:, 0:4, sl] = detect_decode([full], [e.stride], e.reg_max)[:, :4]

    def _fce_detect_decode(self, d, p):
        from oracle.fce_oracle import detect_decode

        raws = [self.t(p[i]).permute(0, 3, 1, 2).float() for i in range(d.nl)]
        y = detect_decode(raws, [d.stride[i] for i in range(d.nl)], d.reg_max)
        yv = p[4]
        b = yv.buf
        self._flat(b).view(b.B, b.W, b.C).copy_(y)
