"""Mel style encoder (reference mel_style_encoder.py:121-152) on the sm_100a kernels.

``MelStyleEncoder(n_mels, style_dim, max_conv_dim, skip_downsamples)`` is the drop-in for
``speech_style_encoder`` / ``duration_style_encoder`` (models.py:49-67): same constructor arguments, same
state-dict keys (legacy ``spectral_norm``: ``weight_orig`` / ``weight_u`` / ``weight_v``), forward
``(B,1,n_mels,F) -> (B,style_dim)``, differentiable (the acoustic stage trains it, stage_type.py:393-410).

Images are kept "row-channel" — (B, H+2, C, W) with zero border rows — so every 3x3 / 5x5 Conv2d is the
stride-1 Conv1d kernel (tcgen05 path included) over R consecutive rows seen as R*C stacked channels; its
data / weight gradients reuse ``sty_conv1d_fwd`` / ``sty_conv1d_wgrad`` on the same overlapping view and
``sty_fold_rows`` folds the view's gradient back.  The learned stride-2 depthwise conv, the 2x2 average
pool and the final region mean have their own kernels (``csrc/image_ops.cu``).  Spectral normalisation acts
on the (Co, Ci*kh*kw) weight matrices only (power iteration + sigma: a few tiny matrix-vector products on
parameters, done with torch ops).  The shortcut's ``conv1x1 -> avg_pool`` is evaluated as
``avg_pool -> conv1x1`` (both linear, 4x fewer MACs).
"""
from __future__ import annotations

import math
from typing import Dict

import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd import Function
from torch.nn.utils import spectral_norm

from . import _lib as L
from ._lib import ACT_LEAKY02, ACT_NONE
from .engine import ConvW, conv1d
from .modules import Node
from . import train_ops as T

INV_SQRT2 = 1.0 / math.sqrt(2.0)
# data gradient of the 3x3 convs as the row-stacked conv of dy with reversed taps (no (N, 3C, W) intermediate, no fold)
FOLD_FREE = True


def _new(shape, like):
    return torch.empty(shape, device=like.device, dtype=torch.float32)


class RowConvFn(Function):
    """R x K Conv2d (stride 1, 'same' along W) on row-channel images: x (B,Hp,C,W) -> (B,Hp,Co,W).

    Window n covers rows n..n+R-1 of the (B*Hp)-row stack and writes output row n + out_off; windows
    whose ``row_mask`` entry is 0 (those straddling two images) produce zeros and receive no gradient."""

    @staticmethod
    def forward(ctx, x, w4, bias, res, cfg):
        B, Hp, Cc, W = x.shape
        Co, _, R, K = w4.shape
        x = x.contiguous()
        N = B * Hp - (R - 1)
        xv = x.as_strided((N, R * Cc, W), (Cc * W, W, 1))
        w3 = w4.detach().permute(0, 2, 1, 3).reshape(Co, R * Cc, K)
        off = cfg.get("out_off", (R - 1) // 2)
        if cfg.get("fold_free") and cfg.get("row_mask") is not None and off == 1 and R == 3:
            # the masked windows write every row but the first and the last one of the stack
            y = torch.empty((B, Hp, Co, W), device=x.device, dtype=torch.float32)
            y[0, 0].zero_()
            y[B - 1, Hp - 1].zero_()
        else:
            y = torch.zeros((B, Hp, Co, W), device=x.device, dtype=torch.float32)
        yv = y.as_strided((N, Co, W), (Co * W, W, 1), off * Co * W)
        resv = None
        if res is not None:
            res = res.contiguous()
            resv = res.as_strided((N, Co, W), (Co * W, W, 1), off * Co * W)
        conv1d(xv, ConvW(w3, None if bias is None else bias.detach()), in_act=cfg.get("in_act", ACT_NONE),
               out_mask=cfg.get("row_mask"), res=resv, out=yv, out_scale=cfg.get("out_scale", 1.0),
               res_scale=cfg.get("res_scale", 1.0), umma=cfg.get("umma", True), wide=cfg.get("wide", False))
        ctx.save_for_backward(x, w4)
        ctx.cfg, ctx.has_bias, ctx.has_res, ctx.off = cfg, bias is not None, res is not None, off
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w4 = ctx.saved_tensors
        cfg = ctx.cfg
        B, Hp, Cc, W = x.shape
        Co, _, R, K = w4.shape
        N = B * Hp - (R - 1)
        in_act, mask = cfg.get("in_act", ACT_NONE), cfg.get("row_mask")
        out_scale, res_scale, umma = cfg.get("out_scale", 1.0), cfg.get("res_scale", 1.0), cfg.get("umma", True)
        dy = dy.contiguous()
        gv = dy.as_strided((N, Co, W), (Co * W, W, 1), ctx.off * Co * W)
        xv = x.as_strided((N, R * Cc, W), (Cc * W, W, 1))
        w3 = w4.detach().permute(0, 2, 1, 3).reshape(Co, R * Cc, K)
        need = list(ctx.needs_input_grad)
        mode = cfg.get("mode")  # discriminator.BackwardMode: which gradients THIS backward pass is run for
        if mode is not None:
            if not mode.weights:
                need[1] = need[2] = False
            if cfg.get("first") and not mode.first_input:
                need[0] = False
        d_res = (dy * res_scale if res_scale != 1.0 else dy) if (ctx.has_res and need[3]) else None
        d_bias = T.channel_sum(gv, mask, out_scale) if (ctx.has_bias and need[2]) else None
        d_w4 = None
        if need[1]:
            d_w3 = T.wgrad(xv, gv, K, 1, in_act=in_act, out_mask=mask, out_scale=out_scale, umma=umma)
            d_w4 = d_w3.reshape(Co, R, Cc, K).permute(0, 2, 1, 3)
        d_x = None
        if need[0] and cfg.get("fold_free"):
            # the adjoint of an R x K 'same' conv is the R x K conv of dy with the taps reversed along both axes and
            # the channel roles swapped — the same row-stacked kernel, no (N, R*C, W) intermediate and no fold pass.
            # It reads dy's border rows as padding: gradients that arrive there (w.r.t. the zero padding rows, from a
            # pooling / strided / valid-conv consumer) carry no meaning for any op of these image stacks, so they are
            # cleared in place first.
            assert R == 3 and ctx.off == 1 and mask is not None
            dy[:, 0].zero_()
            dy[:, Hp - 1].zero_()
            wt = w4.detach().flip(2, 3).permute(1, 2, 0, 3).reshape(Cc, R * Co, K)
            gx = dy.as_strided((N, R * Co, W), (Co * W, W, 1))
            folded = torch.empty((B, Hp, Cc, W), device=x.device, dtype=torch.float32)
            folded[0, 0].zero_()
            folded[B - 1, Hp - 1].zero_()
            conv1d(gx, ConvW(wt, None), out_mask=mask, out=folded.as_strided((N, Cc, W), (Cc * W, W, 1), Cc * W),
                   out_scale=out_scale, umma=umma, wide=cfg.get("wide", False))
            if in_act != ACT_NONE:  # act is elementwise on x: one act' pass on the image
                x3 = x.view(B * Hp, Cc, W)
                _, d_x3 = T.prologue_bwd(folded.view(B * Hp, Cc, W), x3, scale=None, shift=None, alpha=None,
                                         mask=None, act=in_act, want_sums=False)
                d_x = d_x3.view(B, Hp, Cc, W)
            else:
                d_x = folded
        elif need[0]:
            dxp = conv1d(gv, T.transposed_weight(w3), in_mask=mask, out_scale=out_scale, umma=umma)  # (N,R*C,W)
            folded = _new((B, Hp, Cc, W), x)
            L.call("sty_fold_rows", dxp.data_ptr(), folded.data_ptr(), R, B, Hp, Cc, W, L.stream_ptr())
            if in_act != ACT_NONE:  # act is elementwise on x: fold first, then one act' pass
                x3 = x.view(B * Hp, Cc, W)
                _, d_x3 = T.prologue_bwd(folded.view(B * Hp, Cc, W), x3, scale=None, shift=None, alpha=None,
                                         mask=None, act=in_act, want_sums=False)
                d_x = d_x3.view(B, Hp, Cc, W)
            else:
                d_x = folded
        return d_x, d_w4, d_bias, d_res, None


class DwDownFn(Function):
    """LearnedDownSample('half'): depthwise Conv2d 3x3 stride 2 pad 1 (mel_style_encoder.py:28-38)"""

    @staticmethod
    def forward(ctx, x, w, bias):
        B, Hp, Cc, W = x.shape
        x = x.contiguous()
        wk = w.detach().reshape(Cc, 9).contiguous()
        Hop, Wo = (Hp - 3) // 2 + 3, (W - 1) // 2 + 1
        y = _new((B, Hop, Cc, Wo), x)
        L.call("sty_dwconv3x3s2_fwd", x.data_ptr(), wk.data_ptr(), bias.detach().contiguous().data_ptr(),
               y.data_ptr(), B, Hp, Cc, W, L.stream_ptr())
        ctx.save_for_backward(x, wk)
        ctx.w_shape = w.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wk = ctx.saved_tensors
        B, Hp, Cc, W = x.shape
        dy = dy.contiguous()
        dx = _new(x.shape, x) if ctx.needs_input_grad[0] else None
        dw = torch.zeros((Cc, 9), device=x.device, dtype=torch.float32)
        db = torch.zeros((Cc,), device=x.device, dtype=torch.float32)
        L.call("sty_dwconv3x3s2_bwd", dy.data_ptr(), x.data_ptr(), wk.data_ptr(), L.ptr(dx), dw.data_ptr(),
               db.data_ptr(), B, Hp, Cc, W, L.stream_ptr())
        return dx, dw.reshape(ctx.w_shape), db


class AvgPool2Fn(Function):
    """DownSample('half') (mel_style_encoder.py:52-61)"""

    @staticmethod
    def forward(ctx, x):
        B, Hp, Cc, W = x.shape
        x = x.contiguous()
        y = _new((B, (Hp - 2) // 2 + 2, Cc, (W + 1) // 2), x)
        L.call("sty_avgpool2_fwd", x.data_ptr(), y.data_ptr(), B, Hp, Cc, W, L.stream_ptr())
        ctx.shape = x.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        B, Hp, Cc, W = ctx.shape
        dy = dy.contiguous()
        dx = _new(ctx.shape, dy)
        L.call("sty_avgpool2_bwd", dy.data_ptr(), dx.data_ptr(), B, Hp, Cc, W, L.stream_ptr())
        return dx


class RegionMeanFn(Function):
    """mean over a row/column window of a row-channel image -> (B,C)"""

    @staticmethod
    def forward(ctx, x, r0, Rn, w0, Wn):
        B, Hp, Cc, W = x.shape
        x = x.contiguous()
        out = _new((B, Cc), x)
        L.call("sty_region_mean_fwd", x.data_ptr(), out.data_ptr(), B, Hp, Cc, W, r0, Rn, w0, Wn, L.stream_ptr())
        ctx.meta = (x.shape, r0, Rn, w0, Wn)
        return out

    @staticmethod
    def backward(ctx, g):
        (B, Hp, Cc, W), r0, Rn, w0, Wn = ctx.meta
        g = g.contiguous()
        dx = _new((B, Hp, Cc, W), g)
        L.call("sty_region_mean_bwd", g.data_ptr(), dx.data_ptr(), B, Hp, Cc, W, r0, Rn, w0, Wn, L.stream_ptr())
        return dx, None, None, None, None


def _sn(conv: nn.Module) -> nn.Module:
    return spectral_norm(conv)


def _resblk(dim_in, dim_out, down) -> Node:
    n = Node(conv1=_sn(nn.Conv2d(dim_in, dim_in, 3, 1, 1)), conv2=_sn(nn.Conv2d(dim_in, dim_out, 3, 1, 1)))
    if down == "half":
        n.put("downsample_res", Node(conv=_sn(nn.Conv2d(dim_in, dim_in, 3, 2, 1, groups=dim_in))))
    if dim_in != dim_out:
        n.put("conv1x1", _sn(nn.Conv2d(dim_in, dim_out, 1, 1, 0, bias=False)))
    return n


class MelStyleEncoder(nn.Module):
    """Drop-in for reference MelStyleEncoder (mel_style_encoder.py:121-152).  The torch Conv2d / Linear
    children only own the parameters (same names and initialisation as the reference); they are never called."""

    def __init__(self, dim_in=48, style_dim=48, max_conv_dim=384, skip_downsamples=False):
        super().__init__()
        shared = Node()
        shared.put("0", _sn(nn.Conv2d(1, dim_in, 3, 1, 1)))
        self.blocks = []
        d_in, d_out = dim_in, 0
        for i in range(4):
            d_out = min(d_in * 2, max_conv_dim)
            down = "none" if (i == 3 and skip_downsamples) else "half"
            shared.put(str(i + 1), _resblk(d_in, d_out, down))
            self.blocks.append((f"shared.{i + 1}", d_in, d_out, down))
            d_in = d_out
        shared.put("6", _sn(nn.Conv2d(d_out, d_out, 5, 1, 0)))
        self.shared = shared
        self.unshared = nn.Linear(d_out, style_dim)
        self._masks: Dict[tuple, torch.Tensor] = {}

    # ---- spectral normalisation (torch.nn.utils.spectral_norm, 1 power iteration, eps 1e-12)
    def _weight(self, P, Bf, prefix):
        W = P[prefix + ".weight_orig"]
        u, v = Bf[prefix + ".weight_u"], Bf[prefix + ".weight_v"]
        Wm = W.reshape(W.shape[0], -1)
        if self.training:
            with torch.no_grad():
                v.copy_(F.normalize(torch.mv(Wm.t(), u), dim=0, eps=1e-12))
                u.copy_(F.normalize(torch.mv(Wm, v), dim=0, eps=1e-12))
            u, v = u.clone(), v.clone()
        sigma = torch.dot(u, torch.mv(Wm, v))
        return W / sigma

    def _row_mask(self, B, Hp, W, R, device):
        key = (B, Hp, W, R, str(device))
        if key not in self._masks:
            n = torch.arange(B * Hp - (R - 1), device=device)
            ok = ((n % Hp) < (Hp - (R - 1))).float()
            self._masks[key] = ok[:, None].expand(-1, W).contiguous()
        return self._masks[key]

    def forward(self, x):
        L.require_cuda(x, "the mel of MelStyleEncoder")
        L.load()
        return self.encode(x[:, 0])

    def encode(self, mel):
        """mel (B, n_mels, F) -> (B, style_dim)"""
        P, Bf = dict(self.named_parameters()), dict(self.named_buffers())
        B, H, W = mel.shape
        dev = mel.device
        img = torch.zeros((B, H + 2, 1, W), device=dev, dtype=torch.float32)
        img[:, 1:H + 1, 0, :] = mel.to(torch.float32)
        conv3 = lambda t, pre, **kw: RowConvFn.apply(
            t, self._weight(P, Bf, pre), P.get(pre + ".bias"), kw.pop("res", None),
            dict(row_mask=self._row_mask(t.shape[0], t.shape[1], t.shape[3], 3, dev), fold_free=FOLD_FREE, **kw))
        h = conv3(img, "shared.0")
        for pre, d_in, d_out, down in self.blocks:
            r = conv3(h, pre + ".conv1", in_act=ACT_LEAKY02)
            s = h
            if down == "half":
                r = DwDownFn.apply(r, self._weight(P, Bf, pre + ".downsample_res.conv"),
                                   P[pre + ".downsample_res.conv.bias"])
                s = AvgPool2Fn.apply(s)
            if d_in != d_out:  # 1x1 conv on the (B*Hp, C, W) row stack; no bias, so border rows stay zero
                Bq, Hq, _, Wq = s.shape
                w1 = self._weight(P, Bf, pre + ".conv1x1").reshape(d_out, d_in, 1)
                s = T.conv(s.reshape(Bq * Hq, d_in, Wq), w1, None).reshape(Bq, Hq, d_out, Wq)
            h = conv3(r, pre + ".conv2", in_act=ACT_LEAKY02, res=s, out_scale=INV_SQRT2, res_scale=INV_SQRT2)
        Bq, Hq, Cq, Wq = h.shape
        if Hq - 2 < 5 or Wq < 5:
            raise RuntimeError("stylish_tts_b200: mel too short for the style encoder's 5x5 valid conv")
        z = RowConvFn.apply(h, self._weight(P, Bf, "shared.6"), P["shared.6.bias"], None,
                            dict(in_act=ACT_LEAKY02, out_off=0))
        pooled = RegionMeanFn.apply(z, 1, Hq - 2 - 4, 2, Wq - 4)  # valid 5x5 region, then AdaptiveAvgPool2d(1)
        out = T.conv(pooled.unsqueeze(-1), self.unshared.weight.unsqueeze(-1), self.unshared.bias,
                     in_act=ACT_LEAKY02)
        return out.squeeze(-1)


class PitchStyleEncoder(MelStyleEncoder):
    """Drop-in for reference PitchStyleEncoder (mel_style_encoder.py:155-206): pitch and energy are appended
    to the mel as two extra rows, a weight-normed Conv1d(n_mels+2, n_mels, k=1, padding=1) maps them back to
    n_mels rows (the padding=1 of a k=1 conv is literal: two bias-only columns are added), then the same image
    encoder.  forward(mel (B,n_mels,F), pitch (B,F), energy (B,F)) -> (B, style_dim)."""

    def __init__(self, dim_in=48, style_dim=48, max_conv_dim=384, skip_downsamples=False, coarse_multiplier=4):
        super().__init__(dim_in, style_dim, max_conv_dim, skip_downsamples)
        from torch.nn.utils.parametrizations import weight_norm
        self.coarse_multiplier = coarse_multiplier
        self.preconv = weight_norm(nn.Conv1d(dim_in + 2, dim_in, 1, 1, 1))

    def forward(self, x, pitch, energy):
        L.require_cuda(x, "the mel of PitchStyleEncoder")
        pitch, energy = pitch.unsqueeze(1).to(torch.float32), energy.unsqueeze(1).to(torch.float32)
        if self.coarse_multiplier != 1:
            # two (B,1,F) curves resampled to the coarse frame rate (mel_style_encoder.py:189-197): index plumbing
            n = pitch.shape[2] // self.coarse_multiplier
            pitch = F.interpolate(pitch, size=n, mode="linear")
            energy = F.interpolate(energy, size=n, mode="linear")
        L.load()
        P = dict(self.named_parameters())
        w = torch._weight_norm(P["preconv.parametrizations.weight.original1"],
                               P["preconv.parametrizations.weight.original0"], 0)
        xcat = torch.cat([x.to(torch.float32), pitch, energy], dim=1)
        y = T.conv(F.pad(xcat, (1, 1)).contiguous(), w, P["preconv.bias"], umma=False)
        return self.encode(y)
