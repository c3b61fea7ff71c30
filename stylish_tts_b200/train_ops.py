"""Differentiable primitives of the training path: ``torch.autograd.Function``s whose forward
AND backward are the sm_100a kernels of ``csrc/`` (C ABI ``sty_*``).  PyTorch autograd only keeps the
tape between them and does the (B,C)-sized bookkeeping of normalisation statistics; every tensor-sized
pass is one of our kernels.  No CPU / ATen fallback: CPU tensors raise in ``_lib``.

Each primitive cites the reference lines whose autograd formula it replaces.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import torch
from torch.autograd import Function

from . import _lib as L
from ._lib import ACT_NONE, ACT_SNAKE, WgradArgs  # noqa: F401
from .engine import ConvW, conv1d, chan_layernorm, dwconv1d

F32 = torch.float32


def _new(shape, like):
    return torch.empty(shape, device=like.device, dtype=F32)


def _zeros(shape, like):
    return torch.zeros(shape, device=like.device, dtype=F32)


def transposed_weight(w_oik: torch.Tensor) -> ConvW:
    """weights of the adjoint conv: w'[ci, co, k] = w[co, ci, K-1-k] (stride 1, 'same' padding)."""
    return ConvW(w_oik.detach().flip(2).transpose(0, 1).contiguous(), None)


def wgrad(x, dy, K, dil, *, in_scale=None, in_shift=None, in_alpha=None, in_act=ACT_NONE, in_mask=None,
          out_mask=None, out_scale=1.0, umma=True):
    """-> dw (CO, CI, K) of a stride-1 'same' Conv1d (reference layout)."""
    B, CI, T = x.shape
    CO = dy.shape[1]
    dw = _zeros((CO, CI, K), x)
    a = WgradArgs()
    x_bs, x_cs = L._bct(x, "x")
    d_bs, d_cs = L._bct(dy, "dy")
    a.x, a.x_bs, a.x_cs = x.data_ptr(), x_bs, x_cs
    a.dy, a.dy_bs, a.dy_cs = dy.data_ptr(), d_bs, d_cs
    a.in_scale, a.in_shift, a.in_alpha = L.ptr(in_scale), L.ptr(in_shift), L.ptr(in_alpha)
    a.in_mask, a.out_mask, a.dw = L.ptr(in_mask), L.ptr(out_mask), dw.data_ptr()
    a.B, a.CI, a.CO, a.T, a.K, a.dil, a.pad, a.in_act = B, CI, CO, T, K, dil, (K - 1) * dil // 2, in_act
    a.out_scale = out_scale
    from . import engine as E
    a.tensor_cores = int(bool(umma and E.USE_UMMA))
    L.call("sty_conv1d_wgrad", C.byref(a), L.stream_ptr())
    return dw


def channel_sum(x, mask=None, scale=1.0):
    B, Cc, T = x.shape
    bs, cs = L._bct(x, "x")
    out = _zeros((Cc,), x)
    L.call("sty_channel_sum", x.data_ptr(), bs, cs, L.ptr(mask), out.data_ptr(), B, Cc, T, scale, L.stream_ptr())
    return out


def row_moments(x):
    B, Cc, T = x.shape
    bs, cs = L._bct(x, "x")
    mean, var = _new((B, Cc), x), _new((B, Cc), x)
    L.call("sty_row_moments", x.data_ptr(), bs, cs, mean.data_ptr(), var.data_ptr(), B, Cc, T, L.stream_ptr())
    return mean, var


def unshuffle(dy, CO, T, s):
    B = dy.shape[0]
    out = _new((B, CO, T), dy)
    L.call("sty_unshuffle", dy.data_ptr(), out.data_ptr(), B, CO, T, s, L.stream_ptr())
    return out


def prologue_bwd(dxp, x, *, scale, shift, alpha, mask, act, center=None, want_sums, c0=None, c1=None, add=None,
                 sums_only=False):
    """the two passes of the conv-prologue backward; returns (sums (B,C,3) or None, dx or None)."""
    B, Cc, T = x.shape
    x_bs, x_cs = L._bct(x, "x")
    sums = None
    if want_sums:
        sums = _new((B, Cc, 3), x)
        L.call("sty_prologue_bwd_reduce", dxp.data_ptr(), x.data_ptr(), x_bs, x_cs, L.ptr(scale), L.ptr(shift),
               L.ptr(alpha), L.ptr(mask), L.ptr(center), sums.data_ptr(), B, Cc, T, act, L.stream_ptr())
    if sums_only:
        return sums, None
    dx = _new((B, Cc, T), x)
    a_bs, a_cs = (0, 0) if add is None else L._bct(add, "add")
    L.call("sty_prologue_bwd_apply", dxp.data_ptr(), x.data_ptr(), x_bs, x_cs, L.ptr(scale), L.ptr(shift),
           L.ptr(alpha), L.ptr(mask), L.ptr(c0), L.ptr(c1), L.ptr(add), a_bs, a_cs, dx.data_ptr(), Cc * T, T,
           B, Cc, T, act, L.stream_ptr())
    return sums, dx


_POST_MASKS = {}


def _post_mask(mask):
    """binary (B,T) mask -> the kernels' "zero AFTER the prologue" encoding (1 keep, -1 force 0).  The backward keeps
    using the binary mask: for 0/1 masks d/dx [m act(s x + b)] and d/dx [act(s m x + b)] coincide wherever m = 1 and
    are both 0 wherever m = 0."""
    if mask is None:
        return None
    key = (mask.data_ptr(), tuple(mask.shape))
    hit = _POST_MASKS.get(key)
    if hit is None or hit[0] is not mask:
        if len(_POST_MASKS) > 64:
            _POST_MASKS.clear()
        hit = (mask, (2.0 * mask - 1.0).contiguous())
        _POST_MASKS[key] = hit
    return hit[1]


class ConvFn(Function):
    """Conv1d / Linear with the fused prologue (mask, InstanceNorm- or BatchNorm-affine, activation) and
    epilogue (bias, mask, residual, pixel shuffle) of ``sty_conv1d_fwd``.

    norm = None | "instance" (AdaIN: gb (B,2C) = style FC output, ada_norm.py:129-140) |
           "batch" (training-mode BatchNorm1d with weight/bias = bn_w/bn_b, conformer.py:183).
    Backward: data gradient = the same conv kernel with transposed weights, weight gradient =
    ``sty_conv1d_wgrad``, prologue / statistics = ``sty_prologue_bwd_*``.
    """

    @staticmethod
    def forward(ctx, x, w, bias, gb, alpha, res, bn_w, bn_b, cfg):
        B, CI, T = x.shape
        CO, _, K = w.shape
        dil, in_act = cfg.get("dil", 1), cfg.get("in_act", ACT_NONE)
        norm, eps = cfg.get("norm"), cfg.get("eps", 1e-5)
        in_mask, out_mask = cfg.get("in_mask"), cfg.get("out_mask")
        shuffle, umma = cfg.get("shuffle", 0), cfg.get("umma", True)
        out_scale, res_scale = cfg.get("out_scale", 1.0), cfg.get("res_scale", 1.0)
        scale = shift = mean = rstd = None
        if norm == "instance":
            mean, var = row_moments(x)
            rstd = torch.rsqrt(var + eps)
            scale = ((1.0 + gb[:, :CI]) * rstd).contiguous()
            shift = (gb[:, CI:] - mean * scale).contiguous()
        elif norm == "batch":
            # bn_group = s: the conv input is a space-to-depth view (channel c*s + p = phase p of channel c), the
            # BatchNorm it carries belongs to the original channels — statistics are pooled over the s phases and
            # bn_w / bn_b / the running buffers have CI / s entries
            # bn_frac < 1: only that fraction of the T positions is data (zero gaps between windows laid end to end,
            # kept zero by the producers' out_mask): sums run over everything, counts over the valid positions
            grp, frac = cfg.get("bn_group", 1), cfg.get("bn_frac", 1.0)
            stats = cfg.get("bn_buffers")
            if stats is not None and not isinstance(stats, list):
                stats = [stats]  # several BatchNorm modules side by side (concatenated channels)
            if cfg.get("bn_eval"):
                # module.eval() under autograd: nn.BatchNorm1d normalises with the running statistics and
                # leaves them untouched; they are constants of the backward (no batch-statistics terms)
                mean_c = torch.cat([st[0].detach() for st in stats]).clone()
                var_c = torch.cat([st[1].detach() for st in stats]).clone()
                stats = None
            else:
                m_bc, v_bc = row_moments(x)
                mean_c = m_bc.mean(0)
                ex2_c = (v_bc + m_bc * m_bc).mean(0)
                if grp > 1:
                    mean_c = mean_c.view(-1, grp).mean(1)
                    ex2_c = ex2_c.view(-1, grp).mean(1)
                if frac != 1.0:
                    mean_c, ex2_c = mean_c / frac, ex2_c / frac
                var_c = ex2_c - mean_c * mean_c
            if stats is not None:  # running statistics, momentum 0.1, unbiased variance (nn.BatchNorm1d)
                n = B * T * grp * frac
                o = 0
                for rm, rv in stats:
                    k = rm.numel()
                    rm.mul_(0.9).add_(mean_c[o:o + k], alpha=0.1)
                    rv.mul_(0.9).add_(var_c[o:o + k] * (n / max(n - 1, 1)), alpha=0.1)
                    o += k
            rstd_c = torch.rsqrt(var_c + eps)
            sc = bn_w * rstd_c
            sh = bn_b - mean_c * sc
            if grp > 1:
                mean_c, rstd_c, sc, sh = (t.repeat_interleave(grp) for t in (mean_c, rstd_c, sc, sh))
            scale = sc.unsqueeze(0).expand(B, CI).contiguous()
            shift = sh.unsqueeze(0).expand(B, CI).contiguous()
            mean, rstd = mean_c.unsqueeze(0).expand(B, CI).contiguous(), rstd_c
        elif cfg.get("in_scale") is not None:
            raise ValueError("ConvFn: constant in_scale is not supported; use a norm mode")
        cw = ConvW(w.detach(), None if bias is None else bias.detach())
        al = None if alpha is None else alpha.detach().contiguous()
        y = conv1d(x, cw, dil=dil, res=res, in_scale=scale, in_shift=shift, in_alpha=al, in_act=in_act,
                   in_mask=_post_mask(in_mask) if cfg.get("in_mask_post") else in_mask, out_mask=out_mask, shuffle=shuffle, out_scale=out_scale,
                   res_scale=res_scale, umma=umma, wide=cfg.get("wide", False))
        ctx.save_for_backward(x, w, gb, al, scale, shift, mean, rstd, bn_w)
        ctx.cfg, ctx.has_bias, ctx.has_res = cfg, bias is not None, res is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, gb, al, scale, shift, mean, rstd, bn_w = ctx.saved_tensors
        cfg = ctx.cfg
        B, CI, T = x.shape
        CO, _, K = w.shape
        dil, in_act = cfg.get("dil", 1), cfg.get("in_act", ACT_NONE)
        norm = cfg.get("norm")
        in_mask, out_mask = cfg.get("in_mask"), cfg.get("out_mask")
        shuffle, umma = cfg.get("shuffle", 0), cfg.get("umma", True)
        out_scale, res_scale = cfg.get("out_scale", 1.0), cfg.get("res_scale", 1.0)
        need = list(ctx.needs_input_grad)
        mode = cfg.get("mode")  # discriminator.BackwardMode: which gradients THIS walk of the tape is run for
        if mode is not None:
            if not mode.weights:
                need[1] = need[2] = False
            if cfg.get("first") and not mode.first_input:
                need[0] = False
        dy = dy.contiguous()
        d_res = None
        if ctx.has_res and need[5]:
            d_res = dy if res_scale == 1.0 else dy * res_scale
        g = unshuffle(dy, CO, T, shuffle) if shuffle > 1 else dy
        d_bias = channel_sum(g, out_mask, out_scale) if (ctx.has_bias and need[2]) else None
        d_w = None
        if need[1]:
            d_w = wgrad(x, g, K, dil, in_scale=scale, in_shift=shift, in_alpha=al, in_act=in_act,
                        in_mask=_post_mask(in_mask) if cfg.get("in_mask_post") else in_mask, out_mask=out_mask,
                        out_scale=out_scale, umma=umma)
        d_x = d_gb = d_alpha = d_bnw = d_bnb = None
        want_stats = norm is not None or (al is not None and need[4])
        if need[0] or want_stats:
            # in_mask_post: y = m * act(s x + b), so the shift / scale statistics only see the data positions — the
            # data gradient is zeroed in the gaps before the prologue backward sums it
            dxp = conv1d(g, transposed_weight(w), dil=dil, in_mask=out_mask, out_scale=out_scale, umma=umma,
                         out_mask=in_mask if cfg.get("in_mask_post") else None, wide=cfg.get("wide", False))
            plain = scale is None and in_mask is None and in_act == ACT_NONE
            if plain:
                d_x = dxp
            else:
                sums, _ = prologue_bwd(dxp, x, scale=scale, shift=shift, alpha=al, mask=in_mask, act=in_act,
                                       center=mean, want_sums=want_stats, sums_only=True)
                c0 = c1 = None
                if sums is not None and al is not None:
                    d_alpha = sums[:, :, 2].sum(0)
                if norm == "instance":
                    s0, s1 = sums[:, :, 0], sums[:, :, 1]  # s1 is centred: sum g_a (x - mean)
                    d_gb = torch.cat([s1 * rstd, s0], 1)
                    c1 = (-(scale * rstd * rstd) * s1 / T).contiguous()
                    c0 = (-scale * s0 / T - c1 * mean).contiguous()
                elif norm == "batch":
                    s0, s1 = sums[:, :, 0].sum(0), sums[:, :, 1].sum(0)
                    grp = cfg.get("bn_group", 1)
                    if grp > 1:  # pooled statistics: the sums of a channel's phases belong together
                        s0 = s0.view(-1, grp).sum(1).repeat_interleave(grp)
                        s1 = s1.view(-1, grp).sum(1).repeat_interleave(grp)
                    d_bnw, d_bnb = s1 * rstd, s0
                    if grp > 1:
                        d_bnw, d_bnb = d_bnw.view(-1, grp)[:, 0].contiguous(), d_bnb.view(-1, grp)[:, 0].contiguous()
                    if not cfg.get("bn_eval"):
                        n = B * T * grp * cfg.get("bn_frac", 1.0)
                        sc = scale[0]
                        c1c = -(sc * rstd * rstd) * s1 / n
                        c0c = -sc * s0 / n - c1c * mean[0]
                        c1 = c1c.unsqueeze(0).expand(B, CI).contiguous()
                        c0 = c0c.unsqueeze(0).expand(B, CI).contiguous()
                if need[0]:
                    _, d_x = prologue_bwd(dxp, x, scale=scale, shift=shift, alpha=al, mask=in_mask, act=in_act,
                                          want_sums=False, c0=c0, c1=c1)
        return d_x, d_w, d_bias, d_gb, d_alpha, d_res, d_bnw, d_bnb, None


def conv(x, w, bias=None, *, gb=None, alpha=None, res=None, bn_w=None, bn_b=None, **cfg):
    return ConvFn.apply(x, w, bias, gb, alpha, res, bn_w, bn_b, cfg)


class ConvNeXtTailFn(Function):
    """pwconv1 -> Snake -> GRN -> pwconv2 -> + residual  (conv_next.py:85-93, GRN :15-18).

    The 4C-wide pre-activation is not stored: the backward recomputes it with one more
    tensor-core pointwise conv."""

    @staticmethod
    def forward(ctx, y, xres, w1, b1, alpha, gamma, w2, b2f, umma, act=ACT_SNAKE, out_mask=None):
        """act = ACT_SNAKE (generator blocks, alpha per channel) or e.g. ACT_GELU (AdaptiveConvNeXtBlock,
        conv_next.py:96-141; alpha is then ignored); out_mask (B,T) masks the block output before the residual."""
        B, Cc, T = y.shape
        J = w1.shape[0]
        cw1 = ConvW(w1.detach().unsqueeze(-1), b1.detach())
        cw2 = ConvW(w2.detach().unsqueeze(-1), b2f.detach())
        al = alpha.detach().contiguous() if act == ACT_SNAKE else torch.ones(J, device=y.device)
        gm = gamma.detach().contiguous()
        sumsq = _zeros((B, J), y)
        hb = conv1d(y, cw1, out_act=act, out_alpha=al if act == ACT_SNAKE else None, out_sumsq=sumsq, umma=umma)
        gs = _new((B, J), y)
        L.call("sty_grn_scale_fwd", sumsq.data_ptr(), gm.data_ptr(), gs.data_ptr(), B, J, L.stream_ptr())
        out = conv1d(hb, cw2, in_scale=gs, res=xres, out_mask=out_mask, umma=umma)
        ctx.save_for_backward(y, hb, sumsq, gs, w1, b1, al, gm, w2, out_mask)
        ctx.umma, ctx.act = umma, act
        return out

    @staticmethod
    def backward(ctx, dout):
        y, hb, sumsq, gs, w1, b1, al, gm, w2, out_mask = ctx.saved_tensors
        umma, act = ctx.umma, ctx.act
        B, Cc, T = y.shape
        J = w1.shape[0]
        dout = dout.contiguous()
        d_b2f = channel_sum(dout, out_mask)
        d_w2 = wgrad(hb, dout, 1, 1, in_scale=gs, out_mask=out_mask, umma=umma)[:, :, 0]
        g_u = conv1d(dout, transposed_weight(w2.unsqueeze(-1)), in_mask=out_mask, umma=umma)  # (B,J,T)
        r = _new((B, J), y)
        L.call("sty_row_dot", g_u.data_ptr(), hb.data_ptr(), r.data_ptr(), B * J, T, L.stream_ptr())
        gx = torch.sqrt(sumsq)
        M = gx.mean(1, keepdim=True) + 1e-6
        nx = gx / M
        d_nx = r * gm
        d_gamma = (r * nx).sum(0)
        d_gx = d_nx / M - (d_nx * gx).sum(1, keepdim=True) / (M * M * J)
        kc = torch.where(gx > 0, d_gx / gx.clamp_min(1e-30), torch.zeros_like(gx)).contiguous()
        h = conv1d(y, ConvW(w1.unsqueeze(-1), b1), umma=umma)  # pre-activation, recomputed
        d_alpha = _zeros((J,), y)
        L.call("sty_grn_act_bwd", g_u.data_ptr(), h.data_ptr(), gs.data_ptr(), kc.data_ptr(), al.data_ptr(),
               g_u.data_ptr(), d_alpha.data_ptr(), B, J, T, act, L.stream_ptr())
        d_h = g_u
        del h
        d_b1 = channel_sum(d_h)
        d_w1 = wgrad(y, d_h, 1, 1, umma=umma)[:, :, 0]
        d_y = conv1d(d_h, transposed_weight(w1.unsqueeze(-1)), umma=umma)
        return (d_y, dout, d_w1, d_b1, d_alpha if act == ACT_SNAKE else None, d_gamma, d_w2, d_b2f, None, None,
                None)


class ChanLNFn(Function):
    """LayerNorm over channels of (B,C,T) with optional residual input, adaptive ((1+gamma(s)), beta(s))
    or shared (gamma, beta) affine, output mask and ReLU  (text_encoder.py:24-33, ada_norm.py:203-211,
    generator.py:756-778)."""

    @staticmethod
    def forward(ctx, x, res, gamma, beta, gb, cfg):
        B, Cc, T = x.shape
        eps, mask, act = cfg["eps"], cfg.get("mask"), cfg.get("act", ACT_NONE)
        if gb is not None:  # adaptive: gb (B, 2C) rows of the style FC output
            assert gb.stride(1) == 1
            g, be, g_bs, plus_one = gb, gb[:, Cc:], gb.stride(0), True
        else:
            g, be, g_bs, plus_one = gamma.detach().contiguous(), beta.detach().contiguous(), 0, False
        x = x.contiguous()
        if res is not None:
            res = res.contiguous()
        y = chan_layernorm(x, g, be, eps=eps, res=res, g_bs=g_bs, plus_one=plus_one, mask=mask, act=act)
        ctx.save_for_backward(x, res, g, be)
        ctx.cfg, ctx.g_bs, ctx.plus_one, ctx.adaptive = cfg, g_bs, plus_one, gb is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, res, g, be = ctx.saved_tensors
        cfg = ctx.cfg
        B, Cc, T = x.shape
        dy = dy.contiguous()
        dv = _new((B, Cc, T), x)
        dgb = _zeros((B, 2 * Cc) if ctx.adaptive else (2 * Cc,), x)
        L.call("sty_chan_layernorm_bwd", x.data_ptr(), L.ptr(res), x.stride(0), g.data_ptr(), be.data_ptr(),
               ctx.g_bs, int(ctx.plus_one), dy.data_ptr(), L.ptr(cfg.get("mask")), dv.data_ptr(), dgb.data_ptr(),
               2 * Cc if ctx.adaptive else 0, B, Cc, T, cfg["eps"], cfg.get("act", ACT_NONE), L.stream_ptr())
        d_res = dv if res is not None else None
        if ctx.adaptive:
            return dv, d_res, None, None, dgb, None
        return dv, d_res, dgb[:Cc], dgb[Cc:], None, None


def chan_ln(x, *, res=None, gamma=None, beta=None, gb=None, **cfg):
    return ChanLNFn.apply(x, res, gamma, beta, gb, cfg)


class DwConvFn(Function):
    """depthwise Conv1d (conv_next.py:82, conformer.py:176, decoder.py:77-79)."""

    @staticmethod
    def forward(ctx, x, w, bias, K, pad_left):
        B, Cc, T = x.shape
        wk = w.detach().reshape(Cc, K).contiguous()
        out = _new((B, Cc, T), x)
        dwconv1d(x, wk, None if bias is None else bias.detach().contiguous(), K=K, pad_left=pad_left, out=out)
        ctx.save_for_backward(x, wk)
        ctx.K, ctx.pad_left, ctx.w_shape, ctx.has_bias = K, pad_left, w.shape, bias is not None
        return out

    @staticmethod
    def backward(ctx, dy):
        x, wk = ctx.saved_tensors
        B, Cc, T = x.shape
        dy = dy.contiguous()
        x_bs, x_cs = L._bct(x, "x")
        dx = _new((B, Cc, T), x) if ctx.needs_input_grad[0] else None
        dw, db = _zeros((Cc, ctx.K), x), _zeros((Cc,), x)
        L.call("sty_dwconv1d_bwd", dy.data_ptr(), x.data_ptr(), x_bs, x_cs, wk.data_ptr(), L.ptr(dx), Cc * T, T,
               dw.data_ptr(), db.data_ptr(), B, Cc, T, ctx.K, ctx.pad_left, L.stream_ptr())
        return dx, dw.reshape(ctx.w_shape), (db if ctx.has_bias else None), None, None


class DropoutRng:
    """Source of the training-mode masks: one uint64 seed in DEVICE memory (so a captured CUDA graph draws new
    masks on every replay) plus the per-site ids.  ``advance()`` moves to the next step's seed; the sequence is
    a host-side splitmix64 of the user seed, written with ``fill_`` on the current stream."""

    def __init__(self, seed: int = 0, device="cuda"):
        self.state = (int(seed) * 0x9E3779B97F4A7C15 + 0x1234567) & (2 ** 64 - 1)
        self.dev = torch.zeros(1, device=device, dtype=torch.int64)
        self.value = 0
        self.advance()

    def advance(self) -> int:
        self.state = (self.state + 0x9E3779B97F4A7C15) & (2 ** 64 - 1)
        z = self.state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2 ** 64 - 1)
        self.set(z ^ (z >> 31))
        return self.value

    def set(self, value: int) -> None:
        self.value = int(value) & (2 ** 64 - 1)
        self.dev.fill_(self.value - 2 ** 64 if self.value >= 2 ** 63 else self.value)

    def fork(self) -> "DropoutRng":
        """The next seed of the sequence in a FRESH device cell.  Eager forwards use one cell per forward, so a
        backward that runs after a later forward (gradient accumulation, two batches in one loss) still
        rebuilds ITS masks; only the CUDA-graph runner keeps one shared cell (``advance``)."""
        nxt = object.__new__(DropoutRng)
        nxt.state = self.state
        nxt.dev = torch.zeros_like(self.dev)
        nxt.value = 0
        nxt.advance()
        self.state = nxt.state
        return nxt

    def spec(self, site: int, p: float) -> "L.Dropout":
        return L.Dropout(self.dev.data_ptr(), int(site), float(p))

    # checkpoint / resume: the mask stream continues where it stopped (the reference saves torch's RNG state through
    # accelerate.save_state; here the whole state is two 64-bit integers)
    def state_dict(self) -> dict:
        return {"state": int(self.state), "value": int(self.value)}

    def load_state_dict(self, sd: dict) -> None:
        self.state = int(sd["state"]) & (2 ** 64 - 1)
        self.set(int(sd["value"]))


class DropoutFn(Function):
    """y = res + scale * act(x) * keep / (1-p)   (nn.Dropout / Dropout1d / DropPath, see sty_dropout_fwd)"""

    @staticmethod
    def forward(ctx, x, res, rng, site, p, group, act, scale):
        x = x.contiguous()
        y = torch.empty_like(x)
        spec = rng.spec(site, p)
        if res is not None:
            res = res.contiguous()
            assert res.shape == x.shape
        L.call("sty_dropout_fwd", x.data_ptr(), L.ptr(res), y.data_ptr(), x.numel(), group, act, scale,
               C.byref(spec), L.stream_ptr())
        ctx.save_for_backward(x)
        ctx.meta = (rng, site, p, group, act, scale, res is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        rng, site, p, group, act, scale, has_res = ctx.meta
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        spec = rng.spec(site, p)
        L.call("sty_dropout_bwd", x.data_ptr(), dy.data_ptr(), dx.data_ptr(), x.numel(), group, act, scale,
               C.byref(spec), L.stream_ptr())
        return dx, (dy if has_res else None), None, None, None, None, None, None


def dropout(x, rng, site, p, *, res=None, group=1, act=0, scale=1.0):
    """training-mode dropout site; ``rng`` None = regularisers off (plain act / scale / residual)"""
    if rng is None or p <= 0.0:
        if act == 0 and res is None and scale == 1.0:
            return x
        p = 0.0
        rng = _NULL_RNG
    return DropoutFn.apply(x, res, rng, site, p, group, act, scale)


class AdaInDropFn(Function):
    """Dropout(act(AdaIN(x))) materialised (ada_norm.py:181-186 in train() mode): InstanceNorm statistics ->
    per-(b,c) affine -> activation -> hash-masked dropout in one pass; backward = dropout, then the two-pass
    norm backward of the conv prologue (``sty_prologue_bwd_*``)."""

    @staticmethod
    def forward(ctx, x, gb, eps, act, rng, site, p):
        B, Cc, Tn = x.shape
        x = x.contiguous()
        mean, var = row_moments(x)
        rstd = torch.rsqrt(var + eps)
        scale = ((1.0 + gb[:, :Cc]) * rstd).contiguous()
        shift = (gb[:, Cc:] - mean * scale).contiguous()
        y = torch.empty_like(x)
        spec = rng.spec(site, p)
        L.call("sty_affine_act_dropout_fwd", x.data_ptr(), scale.data_ptr(), shift.data_ptr(), y.data_ptr(), B * Cc,
               Tn, act, C.byref(spec), L.stream_ptr())
        ctx.save_for_backward(x, scale, shift, mean, rstd)
        ctx.meta = (act, rng, site, p)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, scale, shift, mean, rstd = ctx.saved_tensors
        act, rng, site, p = ctx.meta
        B, Cc, Tn = x.shape
        dy = dy.contiguous()
        dxp = torch.empty_like(dy)
        spec = rng.spec(site, p)
        L.call("sty_dropout_bwd", dy.data_ptr(), dy.data_ptr(), dxp.data_ptr(), dy.numel(), 1, ACT_NONE, 1.0,
               C.byref(spec), L.stream_ptr())
        sums, _ = prologue_bwd(dxp, x, scale=scale, shift=shift, alpha=None, mask=None, act=act, center=mean,
                               want_sums=True, sums_only=True)
        s0, s1 = sums[:, :, 0], sums[:, :, 1]
        d_gb = torch.cat([s1 * rstd, s0], 1)
        c1 = (-(scale * rstd * rstd) * s1 / Tn).contiguous()
        c0 = (-scale * s0 / Tn - c1 * mean).contiguous()
        _, d_x = prologue_bwd(dxp, x, scale=scale, shift=shift, alpha=None, mask=None, act=act, want_sums=False,
                              c0=c0, c1=c1)
        return d_x, d_gb, None, None, None, None, None


class _NullRng:
    def spec(self, site, p):
        return L.Dropout(None, 0, 0.0)


_NULL_RNG = _NullRng()


def _engine():
    from . import engine
    return engine


class AttentionFn(Function):
    """attention core on a fused (B, 3*H*D, T) q|k|v tensor (text_encoder.py:233-272, conformer.py:112-131);
    ``drop`` = (DropoutRng, site, p) puts SDPA's dropout on the probabilities."""

    @staticmethod
    def forward(ctx, qkv, H, D, lengths, rope, scale, drop=None):
        B, C3, T = qkv.shape
        n = H * D
        assert C3 == 3 * n and qkv.is_contiguous()
        out = _new((B, n, T), qkv)
        lse = _new((B, H, T), qkv)
        q = qkv.data_ptr()
        rc, rs, d_rot = (None, None, 0) if rope is None else (rope[0].data_ptr(), rope[1].data_ptr(), rope[2])
        if drop is not None and drop[2] > 0.0:
            spec = drop[0].spec(drop[1], drop[2])
            L.call("sty_attention_drop_fwd", q, q + 4 * n * T, q + 8 * n * T, qkv.stride(0), out.data_ptr(),
                   out.stride(0), L.ptr(lengths), rc, rs, d_rot, B, H, D, T, scale, lse.data_ptr(), C.byref(spec),
                   L.stream_ptr())
        elif D == 64 and rope is None and lengths is None and T >= 64 and _engine().ATTENTION64:
            drop = None
            _engine().attention64(q, q + 4 * n * T, q + 8 * n * T, qkv.stride(0), out, B, H, T, scale, lse=lse)
        else:
            drop = None
            L.call("sty_attention_lse_fwd", q, q + 4 * n * T, q + 8 * n * T, qkv.stride(0), out.data_ptr(),
                   out.stride(0), L.ptr(lengths), rc, rs, d_rot, B, H, D, T, scale, lse.data_ptr(),
                   L.stream_ptr())
        ctx.save_for_backward(qkv, out, lse)
        ctx.meta = (H, D, lengths, rope, scale, drop)
        return out

    @staticmethod
    def backward(ctx, d_out):
        qkv, out, lse = ctx.saved_tensors
        H, D, lengths, rope, scale, drop = ctx.meta
        B, C3, T = qkv.shape
        n = H * D
        d_out = d_out.contiguous()
        d_qkv = _new((B, C3, T), qkv)
        delta = _new((B, H, T), qkv)
        q, dq = qkv.data_ptr(), d_qkv.data_ptr()
        rc, rs, d_rot = (None, None, 0) if rope is None else (rope[0].data_ptr(), rope[1].data_ptr(), rope[2])
        args = (q, q + 4 * n * T, q + 8 * n * T, qkv.stride(0), out.data_ptr(),
                d_out.data_ptr(), out.stride(0), lse.data_ptr(), L.ptr(lengths), rc, rs, d_rot, dq,
                dq + 4 * n * T, dq + 8 * n * T, d_qkv.stride(0), delta.data_ptr(), B, H, D, T, scale)
        if drop is not None:
            spec = drop[0].spec(drop[1], drop[2])
            L.call("sty_attention_drop_bwd", *args, C.byref(spec), L.stream_ptr())
        elif D == 64 and rope is None and lengths is None and T >= 64 and _engine().ATTENTION64:
            # tcgen05 backward over pre-split tiles (csrc/attention64.cu): dQ kernel + dK/dV kernel
            ws = torch.empty(int(L.load().sty_attention64_bwd_workspace_bytes(B, H, T)), device=qkv.device,
                             dtype=torch.uint8)
            L.call("sty_attention64_bwd", q, q + 4 * n * T, q + 8 * n * T, qkv.stride(0), out.data_ptr(),
                   d_out.data_ptr(), out.stride(0), lse.data_ptr(), dq, dq + 4 * n * T, dq + 8 * n * T,
                   d_qkv.stride(0), B, H, T, scale, ws.data_ptr(), L.stream_ptr())
        else:
            L.call("sty_attention_bwd", *args, L.stream_ptr())
        return d_qkv, None, None, None, None, None, None


class GluFn(Function):
    @staticmethod
    def forward(ctx, x):
        B, C2, T = x.shape
        x = x.contiguous()
        y = _new((B, C2 // 2, T), x)
        L.call("sty_glu_fwd", x.data_ptr(), y.data_ptr(), B, C2 // 2, T, L.stream_ptr())
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        B, C2, T = x.shape
        dx = torch.empty_like(x)
        L.call("sty_glu_bwd", x.data_ptr(), dy.contiguous().data_ptr(), dx.data_ptr(), B, C2 // 2, T,
               L.stream_ptr())
        return dx


class EmbedFn(Function):
    """emb(tokens)*sqrt(C), transposed to (B,C,T) and masked (text_encoder.py:451-453)."""

    @staticmethod
    def forward(ctx, emb, tokens, lengths, scale):
        B, T = tokens.shape
        n_tok, Cc = emb.shape
        out = _new((B, Cc, T), emb)
        L.call("sty_embed_fwd", tokens.data_ptr(), lengths.data_ptr(), emb.detach().contiguous().data_ptr(),
               out.data_ptr(), B, T, Cc, n_tok, scale, L.stream_ptr())
        ctx.save_for_backward(tokens, lengths)
        ctx.meta = (n_tok, Cc, scale)
        return out

    @staticmethod
    def backward(ctx, dx):
        tokens, lengths = ctx.saved_tensors
        n_tok, Cc, scale = ctx.meta
        B, T = tokens.shape
        d_emb = _zeros((n_tok, Cc), dx)
        L.call("sty_embed_bwd", tokens.data_ptr(), lengths.data_ptr(), dx.contiguous().data_ptr(),
               d_emb.data_ptr(), B, T, Cc, n_tok, scale, L.stream_ptr())
        return d_emb, None, None, None


class BmmAlignFn(Function):
    """text_encoding (B,C,T) @ alignment (B,T,F) (speech_predictor.py:60); the alignment is an input
    built from integer durations in the acoustic stage (stage_type.py:99-106) and gets no gradient."""

    @staticmethod
    def forward(ctx, mu, alignment):
        B, Cm, T = mu.shape
        Fr = alignment.shape[2]
        mu = mu.contiguous()
        out = _new((B, Cm, Fr), mu)
        L.call("sty_bmm_fwd", mu.data_ptr(), mu.stride(0), alignment.data_ptr(), alignment.stride(0),
               out.data_ptr(), out.stride(0), B, Cm, Fr, T, L.stream_ptr())
        ctx.save_for_backward(alignment)
        ctx.T = T
        return out

    @staticmethod
    def backward(ctx, d_out):
        (alignment,) = ctx.saved_tensors
        B, Cm, Fr = d_out.shape
        d_out = d_out.contiguous()
        d_mu = _new((B, Cm, ctx.T), d_out)
        L.call("sty_bmm_nt_fwd", d_out.data_ptr(), d_out.stride(0), alignment.data_ptr(), alignment.stride(0),
               d_mu.data_ptr(), d_mu.stride(0), B, Cm, ctx.T, Fr, L.stream_ptr())
        return d_mu, None


class LinearRowsFn(Function):
    """all style FCs of a module as one matrix: h = s @ W^T + b (ada_norm.py:136,204)."""

    @staticmethod
    def forward(ctx, s, W, bias):
        B, I = s.shape
        J = W.shape[0]
        s, Wc = s.contiguous(), W.detach().contiguous()
        h = _new((B, J), s)
        L.call("sty_linear_rows_fwd", s.data_ptr(), Wc.data_ptr(), bias.detach().contiguous().data_ptr(),
               h.data_ptr(), B, I, J, L.stream_ptr())
        ctx.save_for_backward(s, Wc)
        return h

    @staticmethod
    def backward(ctx, dh):
        s, Wc = ctx.saved_tensors
        B, I = s.shape
        J = Wc.shape[0]
        dh = dh.contiguous()
        dW, db = _new((J, I), s), _new((J,), s)
        ds = _new((B, I), s) if ctx.needs_input_grad[0] else None
        L.call("sty_linear_rows_bwd", dh.data_ptr(), s.data_ptr(), Wc.data_ptr(), dW.data_ptr(), db.data_ptr(),
               L.ptr(ds), B, I, J, L.stream_ptr())
        return ds, dW, db


class IstftHeadFn(Function):
    """exp / atan2 / cos-sin head + conv-iSTFT + tanh (generator.py:782-799,896; stft.py:138-187)."""

    @staticmethod
    def forward(ctx, logamp, ri, basis_re, basis_im, hop):
        B, Hs, S = logamp.shape
        logamp, ri = logamp.contiguous(), ri.contiguous()
        audio = _new((B, 1, S * hop), logamp)
        L.call("sty_istft_head_fwd", logamp.data_ptr(), logamp.stride(0), ri.data_ptr(),
               ri.data_ptr() + 4 * Hs * S, ri.stride(0), basis_re.data_ptr(), basis_im.data_ptr(),
               audio.data_ptr(), B, S, Hs, 64, hop, L.stream_ptr())
        ctx.save_for_backward(logamp, ri, audio, basis_re, basis_im)
        ctx.hop = hop
        return audio

    @staticmethod
    def backward(ctx, d_audio):
        logamp, ri, audio, basis_re, basis_im = ctx.saved_tensors
        B, Hs, S = logamp.shape
        d_audio = d_audio.contiguous()
        d_la = _new((B, Hs, S), logamp)
        d_ri = _new((B, 2 * Hs, S), logamp)
        L.call("sty_istft_head_bwd", d_audio.data_ptr(), audio.data_ptr(), logamp.data_ptr(), logamp.stride(0),
               ri.data_ptr(), ri.data_ptr() + 4 * Hs * S, ri.stride(0), basis_re.data_ptr(), basis_im.data_ptr(),
               d_la.data_ptr(), d_ri.data_ptr(), d_ri.data_ptr() + 4 * Hs * S, d_ri.stride(0), B, S, Hs, 64,
               ctx.hop, L.stream_ptr())
        return d_la, d_ri, None, None, None


class AttentionGenericFn(Function):
    """attention core for any head size on separate q / k / v views (prosody encoder 2 x 160 with RoPE 80,
    prosody_encoder.py:63-81; text_encoder.py:233-272).  The backward materialises the probabilities
    (B,H,T,T) — T is the token count — and uses the batched-product kernels."""

    @staticmethod
    def forward(ctx, q, k, v, H, D, lengths, rope, scale, drop=None):
        """drop = (DropoutRng, site, p): SDPA dropout_p on the probabilities — the forward then also goes through
        the materialised (B,H,T,T) matrix, masked by the elementwise dropout kernel (same element index)."""
        from .engine import attention_generic
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        if drop is not None and drop[2] <= 0.0:
            drop = None
        ctx.save_for_backward(q, k, v)
        ctx.meta = (H, D, lengths, rope, scale, drop)
        if drop is None:
            return attention_generic(q, k, v, H=H, D=D, lengths=lengths, rope=rope, scale=scale)
        B, _, T = q.shape
        rc, rs, d_rot = (None, None, 0) if rope is None else (rope[0].data_ptr(), rope[1].data_ptr(), rope[2])
        st = L.stream_ptr()

        def rows(x, use_rope, mul):
            y = _new((B, H, T, D), x)
            L.call("sty_heads_to_rows", x.data_ptr(), x.stride(0), y.data_ptr(), rc if use_rope else None,
                   rs if use_rope else None, d_rot if use_rope else 0, B, H, D, T, mul, 0, st)
            return y

        q_r, k_r, v_t = rows(q, True, scale), rows(k, True, 1.0), rows(v, False, 1.0)
        P = _new((B, H, T, T), q)
        L.call("sty_attn_probs", q_r.data_ptr(), k_r.data_ptr(), L.ptr(lengths), P.data_ptr(), B, H, D, T, st)
        spec = drop[0].spec(drop[1], drop[2])
        L.call("sty_dropout_fwd", P.data_ptr(), None, P.data_ptr(), P.numel(), 1, ACT_NONE, 1.0, C.byref(spec), st)
        o_r = _new((B, H, T, D), q)
        L.call("sty_bmm_fwd", P.data_ptr(), T * T, v_t.data_ptr(), T * D, o_r.data_ptr(), T * D, B * H, T, D, T, st)
        out = _new((B, H * D, T), q)
        L.call("sty_heads_to_rows", o_r.data_ptr(), out.stride(0), out.data_ptr(), None, None, 0, B, H, D, T, 1.0,
               1, st)
        return out

    @staticmethod
    def backward(ctx, d_out):
        q, k, v = ctx.saved_tensors
        H, D, lengths, rope, scale, drop = ctx.meta
        B, _, T = q.shape
        d_out = d_out.contiguous()
        rc, rs, d_rot = (None, None, 0) if rope is None else (rope[0].data_ptr(), rope[1].data_ptr(), rope[2])
        st = L.stream_ptr()

        def rows(x, use_rope, mul):
            y = _new((B, H, T, D), x)
            L.call("sty_heads_to_rows", x.data_ptr(), x.stride(0), y.data_ptr(), rc if use_rope else None,
                   rs if use_rope else None, d_rot if use_rope else 0, B, H, D, T, mul, 0, st)
            return y

        def heads(y, use_rope, mul):
            x = _new((B, H * D, T), y)
            L.call("sty_heads_to_rows", y.data_ptr(), x.stride(0), x.data_ptr(), rc if use_rope else None,
                   rs if use_rope else None, d_rot if use_rope else 0, B, H, D, T, mul, 1, st)
            return x

        q_r, k_r, v_t, do_t = rows(q, True, scale), rows(k, True, 1.0), rows(v, False, 1.0), rows(d_out, False, 1.0)
        P = _new((B, H, T, T), q)
        L.call("sty_attn_probs", q_r.data_ptr(), k_r.data_ptr(), L.ptr(lengths), P.data_ptr(), B, H, D, T, st)
        BH, TT, TD = B * H, T * T, T * D
        dv_t, dq_r, dk_r = _new((B, H, T, D), q), _new((B, H, T, D), q), _new((B, H, T, D), q)
        dP = _new((B, H, T, T), q)
        Pd = P
        if drop is not None:  # dV sees the dropped probabilities, dP is masked the same way
            spec = drop[0].spec(drop[1], drop[2])
            Pd = _new((B, H, T, T), q)
            L.call("sty_dropout_fwd", P.data_ptr(), None, Pd.data_ptr(), P.numel(), 1, ACT_NONE, 1.0, C.byref(spec),
                   st)
        L.call("sty_bmm_tn_fwd", Pd.data_ptr(), TT, do_t.data_ptr(), TD, dv_t.data_ptr(), TD, BH, T, D, T, st)
        L.call("sty_bmm_nt_fwd", do_t.data_ptr(), TD, v_t.data_ptr(), TD, dP.data_ptr(), TT, BH, T, T, D, st)
        if drop is not None:
            L.call("sty_dropout_bwd", dP.data_ptr(), dP.data_ptr(), dP.data_ptr(), dP.numel(), 1, ACT_NONE, 1.0,
                   C.byref(spec), st)
        L.call("sty_softmax_bwd", P.data_ptr(), dP.data_ptr(), BH * T, T, st)
        L.call("sty_bmm_fwd", dP.data_ptr(), TT, k_r.data_ptr(), TD, dq_r.data_ptr(), TD, BH, T, D, T, st)
        L.call("sty_bmm_tn_fwd", dP.data_ptr(), TT, q_r.data_ptr(), TD, dk_r.data_ptr(), TD, BH, T, D, T, st)
        return (heads(dq_r, True, scale), heads(dk_r, True, 1.0), heads(dv_t, False, 1.0), None, None, None, None,
                None, None)
